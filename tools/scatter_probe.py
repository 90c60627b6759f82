"""Round-2 probe: the host half of the SQP transport (ogb_host_eval_fd_scatter) and of keep_zeros with different
prefetch distances ($OGB200_HOST_PREFETCH is read once per process: run once per value).
    OGB200_HOST_PREFETCH=48 python tools/scatter_probe.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import OpenGoddard.optimize as api  # noqa: E402
from opengoddard_b200 import workloads  # noqa: E402

B = 4096
wl = workloads.build("cfg2_goddard50", api)
eng = wl.prob.compile(wl.obj)
P = torch.from_numpy(workloads.make_batch(wl, B)).pin_memory()
n, M = eng.nvars, eng.nrows
m = M - 1
sess = eng.host_session(B)
hc = np.empty((B, M))
Cb = np.zeros((B, n, m))
Gb = np.zeros((B, n))
Cp = [Cb.ctypes.data + b * Cb.strides[0] for b in range(B)]
Gp = [Gb.ctypes.data + b * Gb.strides[0] for b in range(B)]
for label, fn in (("scatter", lambda: sess.eval_fd_scatter(P, hc, Cp, m, m, Gp)),):
    fn()
    t0 = time.perf_counter()
    for _ in range(4):
        fn()
    dt = (time.perf_counter() - t0) / 4
    print("prefetch=%s %s: %.2f ms per 4096 evals = %.0f evals/s" % (os.environ.get("OGB200_HOST_PREFETCH", "default"), label, dt * 1e3, B / dt))
del Cb
hJ = np.zeros((B, n, M))
sess.eval_fd(P, hc, hJ, mode="dense")
t0 = time.perf_counter()
for _ in range(4):
    sess.eval_fd(P, hc, hJ, mode="keep_zeros")
dt = (time.perf_counter() - t0) / 4
print("prefetch=%s keep_zeros: %.2f ms = %.0f evals/s" % (os.environ.get("OGB200_HOST_PREFETCH", "default"), dt * 1e3, B / dt))
