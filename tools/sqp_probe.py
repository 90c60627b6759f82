"""Device SQP timing probe: multi-start SLSQP with the QP on the GPU against the SciPy-core process pool.
    python tools/sqp_probe.py [workload] [starts] [maxiter]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import OpenGoddard.optimize as api  # noqa: E402
from opengoddard_b200 import workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_goddard50"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
maxiter = int(sys.argv[3]) if len(sys.argv) > 3 else 6
wl = workloads.build(name, api)
eng = wl.prob.compile(wl.obj)
P0 = workloads.make_batch(wl, S)
with eng.device_sqp(S, 1e-6, maxiter) as dq:
    print("device SQP memory: %.2f GB for %d instances" % (dq.bytes / 1e9, S))
    dq.solve(P0[:8])
    torch.cuda.synchronize()
    log = []
    t0 = time.perf_counter()
    res = dq.solve(P0, callback=lambda r, mode: log.append((time.perf_counter(), int((mode == 1).sum()), int((mode == -1).sum()))))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    clk = dq.k.scalars(S)
    print("LSEI phase 1 (triangularise C) share of LSEI: %.2f" % (float(clk["clk_lsei_phase1"].sum()) / float(clk["clk_lsei"].sum())))
    clk = {k: v for k, v in clk.items() if k != "clk_lsei_phase1"}
    tot = sum(float(clk[k].sum()) for k in clk if k.startswith("clk_"))
    print("SM cycles per phase (share):", {k[4:]: "%.1f%%" % (100 * float(clk[k].sum()) / tot) for k in clk if k.startswith("clk_")},
          " mean cycles per instance-iteration: %.0f" % (tot / max(1, int(res["nit"].sum()))))
its = int(res["nit"].sum())
print("%s x %d, maxiter %d: %.3f s, %d rounds, %d instance-iterations -> %.0f it/s; status counts %s" % (
    name, S, maxiter, dt, res["rounds"], its, its / dt, dict(zip(*np.unique(res["status"], return_counts=True)))))
prev = t0
for i, (t, n1, ng) in enumerate(log[:12]):
    print("  round %2d: %.1f ms  (after: %d in line search, %d need gradients)" % (i + 1, (t - prev) * 1e3, n1, ng))
    prev = t
sc = None
print("fun: min %.6f median %.6f max %.6f" % (res["fun"].min(), np.median(res["fun"]), res["fun"].max()))
