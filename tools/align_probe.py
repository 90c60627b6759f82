"""Round-2 probe: does the 8-byte misalignment of every other Jacobian column (M odd) cost K2 bandwidth?  Goddard with
49 / 50 / 51 / 52 nodes: M = 9 N + 7 = 448 / 457 / 466 / 475 rows.  K2 alone, CUDA events, L2 flushed.
    python tools/align_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import OpenGoddard.optimize as api  # noqa: E402
from opengoddard_b200 import workloads  # noqa: E402

B = 4096
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for N in (49, 50, 51, 52, 48, 47):
    wl = workloads.goddard(api, nodes=(N,))
    eng = wl.prob.compile(wl.obj)
    P = torch.from_numpy(workloads.make_batch(wl, B)).cuda()
    n, M = eng.nvars, eng.nrows
    c = torch.empty((B, M), dtype=torch.float64, device="cuda")
    J = torch.empty((B, n, M), dtype=torch.float64, device="cuda")
    DX = eng.dx_gemm(P, clip=True)
    for _ in range(3):
        eng.sweep_fd(P, DX, c, J)
    best = 1e9
    for _ in range(10):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.sweep_fd(P, DX, c, J)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    gbs = B * (8 * n + 8 * M * (n + 1)) / (best * 1e-3) / 1e9
    print("goddard N=%d n=%d M=%d (%s)  K2 %.4f ms  %.0f GB/s" % (N, n, M, "even" if M % 2 == 0 else "odd", best, gbs), flush=True)
    del c, J, DX, P, eng
