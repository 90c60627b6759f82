"""Round-2 probe: how the fused sweep kernel should write its zeros (OGB_OPT_ZERO_MODE): 0 = per column by
the warp that owns it (round 1), 1 = the same with st.global.cs, 4 / 8 = one / two dedicated writer warps per CTA
zero the CTA's next work item while the other warps compute.  K2 alone (ogb_sweep), CUDA events, L2 flushed.
    python tools/zero_mode_probe.py [workload] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import OpenGoddard.optimize as api  # noqa: E402
from opengoddard_b200 import workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_goddard50"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
wl = workloads.build(name, api)
eng = wl.prob.compile(wl.obj)
P = torch.from_numpy(workloads.make_batch(wl, B)).cuda()
n, M = eng.nvars, eng.nrows
c = torch.empty((B, M), dtype=torch.float64, device="cuda")
J = torch.empty((B, n, M), dtype=torch.float64, device="cuda")
DX = eng.dx_gemm(P, clip=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for threads in (256, 128, 384):
    try:
        eng.set_option(1, threads)
    except Exception as ex:
        print("threads", threads, "not possible:", str(ex)[:80])
        continue
    for mode in (0, 16):
        eng.set_option(12, mode)
        J.fill_(float("nan"))
        for _ in range(2):
            eng.sweep_fd(P, DX, c, J)
        best = 1e9
        for _ in range(7):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.sweep_fd(P, DX, c, J)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        if ref is None:
            ref = J.clone()
        same = bool(torch.equal(J, ref))
        gbs = B * (8 * n + 8 * M * (n + 1)) / (best * 1e-3) / 1e9
        print("%s B=%d threads=%d zero_mode=%d  K2 %.4f ms  %.0f GB/s  identical=%s" % (name, B, threads, mode, best, gbs, same))
