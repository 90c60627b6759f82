"""K1 (ogb_dx_gemm) timing with preallocated output (development aid)."""
import sys
import torch
sys.path.insert(0, ".")
import OpenGoddard.optimize as api
from opengoddard_b200 import workloads
for name in ("cfg2_goddard50", "cfg3_goddard_knot30x2", "cfg4_polar3x40", "cfg5_lowthrust128"):
    wl = workloads.build(name, api)
    eng = wl.prob.compile(wl.obj, jit=False)
    B = 4096
    P = torch.as_tensor(workloads.make_batch(wl, 64), device="cuda").repeat(B // 64, 1).contiguous()
    DX = torch.empty((B, eng.ndx), dtype=torch.float64, device="cuda")
    for _ in range(3):
        eng.dx_gemm(P, out=DX, clip=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.dx_gemm(P, out=DX, clip=True); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    flop = 2.0 * B * sum(ns * N * N for ns, N in zip(wl.prob.number_of_states, wl.prob.nodes))
    print("%-24s K1 min %.4f ms median %.4f ms max %.4f ms  %.2f TFLOP/s (min)" % (
        name, min(ts), sorted(ts)[10], max(ts), flop / min(ts) / 1e9))
