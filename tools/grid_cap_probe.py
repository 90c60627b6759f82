"""K2 time vs persistent-grid size (OGB_OPT_GRID_CAP): does an even number of items per CTA beat
the full grid with a ragged last round?"""
import sys
import torch
sys.path.insert(0, ".")
import OpenGoddard.optimize as api
from opengoddard_b200 import workloads
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_goddard50"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
wl = workloads.build(name, api)
eng = wl.prob.compile(wl.obj)
P = torch.as_tensor(workloads.make_batch(wl, min(B, 512)), device="cuda")
P = P.repeat((B + P.shape[0] - 1) // P.shape[0], 1)[:B].contiguous()
c = torch.empty((B, eng.nrows), dtype=torch.float64, device="cuda")
J = torch.empty((B, eng.nvars, eng.nrows), dtype=torch.float64, device="cuda")
DX = eng.dx_gemm(P, clip=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
bpe = 8 * eng.nvars + 8 * eng.nrows * (eng.nvars + 1)
slots = 148 * eng.info.ctas_per_sm
caps = [0, -1] + sorted({-(-B // r) for r in range(max(1, B // slots), B // slots + 8)}, reverse=True)
for cap in caps:
    eng.set_option(3, cap)
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
    print("grid cap %4d (%.2f items per CTA): avg %.3f ms min %.3f  %.0f GB/s" % (
        cap, B / (cap if cap > 0 else (slots if cap == 0 else B)), ms, min(ts), B * bpe / ms / 1e6))
