"""K2 at small batches: whole-instance work items vs OGB_OPT_AUTO_SPLIT (more, smaller items)."""
import sys
import torch
sys.path.insert(0, ".")
import OpenGoddard.optimize as api
from opengoddard_b200 import workloads
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_goddard50"
wl = workloads.build(name, api)
eng = wl.prob.compile(wl.obj)
bpe = 8 * eng.nvars + 8 * eng.nrows * (eng.nvars + 1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B in [int(x) for x in sys.argv[2:]] or [128, 365, 512, 1024, 2048]:
    P = torch.as_tensor(workloads.make_batch(wl, min(B, 512)), device="cuda")
    P = P.repeat((B + P.shape[0] - 1) // P.shape[0], 1)[:B].contiguous()
    c = torch.empty((B, eng.nrows), dtype=torch.float64, device="cuda")
    J = torch.empty((B, eng.nvars, eng.nrows), dtype=torch.float64, device="cuda")
    DX = eng.dx_gemm(P, clip=True)
    ref = None
    for split in (0, 1):
        eng.set_option(7, split)
        for _ in range(3):
            eng.sweep_fd(P, DX, c, J)
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.sweep_fd(P, DX, c, J); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sum(ts) / len(ts)
        same = True if ref is None else bool(torch.equal(ref, J))
        ref = J.clone() if ref is None else ref
        print("%s B %5d auto_split %d: %.4f ms  %.0f GB/s  identical=%s" % (name, B, split, ms, B * bpe / ms / 1e6, same))
    eng.set_option(7, 1)
