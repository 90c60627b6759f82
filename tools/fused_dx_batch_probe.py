"""Round-2 probe: ogb_eval_fd as K1 + K2 (two launches) vs D.X inside K2 (OGB_OPT_FUSED_DX, one launch) across batch
sizes -- K1 is latency-bound (~25-35 us whatever the batch), so below some batch the single launch should win.
    python tools/fused_dx_batch_probe.py [workload]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import OpenGoddard.optimize as api  # noqa: E402
from opengoddard_b200 import workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_goddard50"
wl = workloads.build(name, api)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for jit in (True, False):
    eng = wl.prob.compile(wl.obj, jit=jit)
    for B in (1, 16, 128, 512, 1024, 2048, 4096):
        P = torch.from_numpy(workloads.make_batch(wl, B)).cuda()
        c = torch.empty((B, eng.nrows), dtype=torch.float64, device="cuda")
        J = torch.empty((B, eng.nvars, eng.nrows), dtype=torch.float64, device="cuda")
        res = {}
        for fused in (0, 1):
            eng.set_option(4, fused)
            for _ in range(3):
                eng.eval_fd(P, out_c=c, out_J=J)
            best = 1e9
            for _ in range(9):
                flush.zero_()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.eval_fd(P, out_c=c, out_J=J)
                e1.record()
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            res[fused] = best
        print("%s jit=%d B=%5d  K1+K2 %.1f us   fused D.X %.1f us   ratio %.3f" % (name, jit, B, res[0] * 1e3, res[1] * 1e3, res[1] / res[0]))
        del P, c, J
