"""torchrun --nproc-per-node 2 tools/solve_sharded_check.py : Problem.solve_batch sharded over the
ranks (NCCL gather at the end) == the same starts solved by one rank alone, bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import OpenGoddard.optimize as api
from opengoddard_b200 import batch, workloads

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = workloads.build("cfg1_brachistochrone20", api)
wl.prob.compile(wl.obj, device="cuda:%d" % local)
P0 = np.vstack([np.array(wl.prob.p)[None], workloads.make_batch(wl, 6)])
res = wl.prob.solve_batch(P0, wl.obj, ftol=1e-6, maxiter=25)
# the same starts on this rank alone (a group of one rank = no sharding)
solo_group = dist.new_group([rank]) if False else None
lo, hi = batch.shard_range(len(P0), rank, world)
from opengoddard_b200 import sqp
alone = wl.prob._solve_rows(wl.prob._engine, P0, wl.obj, 1e-6, 25, None, 1, sqp)
ok = np.array_equal(alone["x"], res["x"]) and np.array_equal(alone["status"], res["status"]) and np.array_equal(alone["fun"], res["fun"])
print("rank %d: shard [%d, %d) of %d, sharded == single-rank: %s, converged %d / %d, best start %d, t_f %.8f" % (
    rank, lo, hi, len(P0), ok, int((res["status"] == 0).sum()), len(P0), batch.best_instance(res),
    res["x"][batch.best_instance(res), -1]), flush=True)
assert ok
dist.destroy_process_group()
