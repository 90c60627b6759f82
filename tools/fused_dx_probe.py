"""ogb_eval_fd as K1 + K2 (two launches, D.X through an HBM scratch) vs K2 with D.X inside
(OGB_OPT_FUSED_DX: DMMAs one item ahead, one launch): time and bit-identity."""
import sys
import torch
sys.path.insert(0, ".")
import OpenGoddard.optimize as api
from opengoddard_b200 import workloads
for spec in sys.argv[1:] or ["cfg2_goddard50:4096", "cfg2_goddard50:1024", "cfg3_goddard_knot30x2:4096",
                             "cfg4_polar3x40:512", "cfg5_lowthrust128:1024", "edge_stress_mixed:300"]:
    name, B = spec.split(":"); B = int(B)
    wl = workloads.build(name, api)
    eng = wl.prob.compile(wl.obj)
    P = torch.as_tensor(workloads.make_batch(wl, min(B, 512)), device="cuda")
    P = P.repeat((B + P.shape[0] - 1) // P.shape[0], 1)[:B].contiguous()
    c = torch.empty((B, eng.nrows), dtype=torch.float64, device="cuda")
    J = torch.empty((B, eng.nvars, eng.nrows), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    bpe = 8 * eng.nvars + 8 * eng.nrows * (eng.nvars + 1)
    ref = None
    for fused in (0, 1, 0, 1):
        eng.set_option(4, fused)
        for _ in range(3):
            eng.eval_fd(P, out_c=c, out_J=J)
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.eval_fd(P, out_c=c, out_J=J); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sum(ts) / len(ts)
        ce = eng.eval(P)
        if ref is None:
            ref = (c.clone(), J.clone(), ce.clone())
        same = torch.equal(c, ref[0]) and torch.equal(J, ref[1]) and torch.equal(ce, ref[2])
        print("%-24s B %5d fused_dx %d: %.4f ms per eval_fd  %.0f GB/s  identical=%s" % (name, B, fused, ms, B * bpe / ms / 1e6, same))
    eng.set_option(4, 0)
