"""Round-2 probe: ogb_eval_fd (K1 + K2) with the sweep kernel launched behind K1 by programmatic dependent launch
(OGB_OPT_PDL = 14, default 1) against a plain stream-ordered launch (0).  CUDA events, L2 flushed between calls,
median and best of 30.  Bit-identical.
    python tools/pdl_probe.py"""
import os
import statistics
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
    c = torch.empty((B, eng.nrows), dtype=torch.float64, device="cuda")
    J = torch.empty((B, eng.nvars, eng.nrows), dtype=torch.float64, device="cuda")
    ref = None
    for rep in range(2):
        for pdl in (0, 1):
            eng.set_option(14, pdl)
            for _ in range(3):
                eng.eval_fd(P, out_c=c, out_J=J)
            ts = []
            for _ in range(30):
                flush.zero_()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.eval_fd(P, out_c=c, out_J=J)
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            if ref is None:
                ref = (c.clone(), J.clone())
            same = bool(torch.equal(c, ref[0]) and torch.equal(J, ref[1]))
            print("%s B=%d pdl=%d  eval_fd median %.4f ms  best %.4f ms  identical=%s" % (
                name, B, pdl, statistics.median(ts), min(ts), same), flush=True)
    del c, J, ref
