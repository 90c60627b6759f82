"""Round-2 probe: K1 (D.X batched FP64 DMMA GEMM): the round-1 kernel with whole-row work units (OGB_OPT_GEMM_UNIT = 8;
= 2 selects its 16-node units, measured slower earlier) against the latency-organised form (0, default: p loads first,
cp.async staging of D and the bounds, software-pipelined chunks).  Bit-identical.
    python tools/k1_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import OpenGoddard.optimize as api  # noqa: E402
from opengoddard_b200 import workloads  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, B in (("cfg2_goddard50", 4096), ("cfg2_goddard50", 1024), ("cfg3_goddard_knot30x2", 4096),
                ("cfg4_polar3x40", 512), ("cfg5_lowthrust128", 1024)):
    wl = workloads.build(name, api)
    eng = wl.prob.compile(wl.obj)
    P = torch.from_numpy(workloads.make_batch(wl, B)).cuda()
    ref = None
    for unit in (8, 0):
        eng.set_option(13, unit)
        DX = eng.dx_gemm(P, clip=True)
        best = 1e9
        for _ in range(10):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.dx_gemm(P, out=DX, clip=True)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        if ref is None:
            ref = DX.clone()
        flop = 2.0 * sum(N * N * ns for N, ns in zip(wl.prob.nodes, wl.prob.number_of_states)) * B
        print("%s B=%d unit=%d  K1 %.2f us  %.2f TFLOP/s  identical=%s" % (name, B, unit, best * 1e3, flop / (best * 1e-3) / 1e12,
                                                                          bool(torch.equal(DX, ref))))
