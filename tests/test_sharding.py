"""Instance-batch sharding and the result gather, world_size 2 on the gloo backend (CPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from opengoddard_b200 import batch, workloads
from oracle import og_numpy


def test_shard_ranges_partition_the_batch():
    for total in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [batch.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == batch.shard_sizes(total, world)
    with pytest.raises(ValueError):
        batch.shard_range(4, 2, 2)


def test_seeded_batch_is_shard_invariant():
    wl = workloads.build("cfg2_goddard50", og_numpy)
    full = workloads.make_batch(wl, 11)
    lo, hi = batch.shard_range(11, 1, 2)
    assert np.array_equal(workloads.make_batch(wl, hi - lo, first=lo), full[lo:hi])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wl = workloads.build("cfg1_brachistochrone20", og_numpy)
        lo, hi = batch.shard_range(total, rank, world)
        P = workloads.make_batch(wl, hi - lo, first=lo)
        lbs, ubs = og_numpy.bounds_arrays(wl.prob)
        # per-instance result of this shard (the oracle stands in for the device evaluation here:
        # the test is about the sharding / gather plumbing)
        cost = torch.tensor([og_numpy.eval_c(wl.prob, wl.obj, np.clip(p, lbs, ubs))[-1] for p in P])
        gathered_p = batch.gather_rows(torch.from_numpy(P), total)
        gathered_cost = batch.gather_rows(cost, total)
        if rank == 0:
            out.put((gathered_p.numpy(), gathered_cost.numpy()))
    finally:
        dist.destroy_process_group()


def test_gather_rows_gloo_world2():
    total, world = 7, 2                                   # uneven shards: 4 + 3
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, out)) for r in range(world)]
    for p in procs:
        p.start()
    got_p, got_cost = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    wl = workloads.build("cfg1_brachistochrone20", og_numpy)
    full = workloads.make_batch(wl, total)
    assert np.array_equal(got_p, full)
    lbs, ubs = og_numpy.bounds_arrays(wl.prob)
    ref = [og_numpy.eval_c(wl.prob, wl.obj, np.clip(p, lbs, ubs))[-1] for p in full]
    assert np.array_equal(got_cost, np.array(ref))


def test_gather_rows_single_process_passthrough():
    x = torch.arange(12.0).reshape(4, 3)
    assert batch.gather_rows(x, 4) is x


def _solve_worker(rank, world, port, total, out):
    import torch.distributed as dist
    from opengoddard_b200 import sqp
    from tests.test_sqp import OracleEvaluator
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wl = workloads.build("cfg1_brachistochrone20", og_numpy)
        ev = OracleEvaluator(wl)
        P = workloads.make_batch(wl, total)                # every rank holds the full start set

        def solve(rows):                                   # the oracle stands in for the device evaluator
            if len(rows) == 0:
                return {"x": rows.reshape(0, P.shape[1]), "fun": np.zeros(0), "status": np.zeros(0, dtype=int),
                        "message": []}
            r = sqp.slsqp_batch(ev, rows, ev.lb, ev.ub, 64, 41, ftol=1e-6, maxiter=4)
            return {"x": r["x"], "fun": r["fun"], "status": r["status"], "message": r["message"]}

        res = batch.run_sharded(solve, P)
        out.put((rank, res["x"], res["fun"], res["status"], res["message"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 1])
def test_run_sharded_matches_the_single_process_run(total):
    """world_size 2 over gloo: sharded multi-start == unsharded, bit for bit, on every rank
    (total = 1 leaves rank 1 with an empty shard)."""
    from opengoddard_b200 import sqp
    from tests.test_sqp import OracleEvaluator
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_solve_worker, args=(r, world, port, total, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = [out.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    wl = workloads.build("cfg1_brachistochrone20", og_numpy)
    ev = OracleEvaluator(wl)
    ref = sqp.slsqp_batch(ev, workloads.make_batch(wl, total), ev.lb, ev.ub, 64, 41, ftol=1e-6, maxiter=4)
    for rank, x, fun, status, message in got:
        assert np.array_equal(x, ref["x"]) and np.array_equal(fun, ref["fun"])
        assert np.array_equal(status, ref["status"]) and message == ref["message"]
    assert batch.best_instance({"fun": ref["fun"], "status": ref["status"]}) == int(np.argmin(ref["fun"]))


def test_best_instance_prefers_converged_starts():
    assert batch.best_instance({"fun": [3.0, 1.0, 2.0], "status": [0, 9, 0]}) == 2
    assert batch.best_instance({"fun": [3.0, 1.0, 2.0], "status": [9, 9, 9]}) == 1
