"""Round-2 probe: tail refinement of the dense FD sweep (OGB_OPT_TAIL_REFINE = 15): the last instances of the batch
-- `pct` per cent of one wave of work items -- are cut into items of >= 64 columns so the persistent CTAs finish
together.  K2 alone, CUDA events, L2 flushed, best and median of 15.  Bit-identical.
    python tools/tail_probe.py"""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import OpenGoddard.optimize as api  # noqa: E402
from opengoddard_b200 import workloads  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, B in (("cfg2_goddard50", 4096), ("cfg2_goddard50", 2048), ("cfg3_goddard_knot30x2", 4096), ("cfg5_lowthrust128", 1024)):
    wl = workloads.build(name, api)
    eng = wl.prob.compile(wl.obj)
    P = torch.from_numpy(workloads.make_batch(wl, B)).cuda()
    n, M = eng.nvars, eng.nrows
    c = torch.empty((B, M), dtype=torch.float64, device="cuda")
    J = torch.empty((B, n, M), dtype=torch.float64, device="cuda")
    DX = eng.dx_gemm(P, clip=True)
    ref = None
    for pct in (0, 25, 50, 75, 100, 150, 200, 0):
        eng.set_option(15, pct)
        J.fill_(float("nan"))
        for _ in range(2):
            eng.sweep_fd(P, DX, c, J)
        ts = []
        for _ in range(15):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.sweep_fd(P, DX, c, J)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        if ref is None:
            ref = (c.clone(), J.clone())
        same = bool(torch.equal(J, ref[1]) and torch.equal(c, ref[0]))
        best = min(ts)
        gbs = B * (8 * n + 8 * M * (n + 1)) / (best * 1e-3) / 1e9
        print("%s B=%d tail=%3d%%  K2 best %.4f ms  median %.4f ms  %.0f GB/s  identical=%s" % (
            name, B, pct, best, statistics.median(ts), gbs, same), flush=True)
    del c, J, DX, P, ref, eng
