"""Single-instance latency of the SciPy-facing device callbacks (development aid)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import OpenGoddard.optimize as api
from opengoddard_b200 import workloads

t0 = time.time()
wl = workloads.build("cfg2_goddard50", api)
t1 = time.time()
eng = wl.prob.compile(wl.obj)
torch.cuda.synchronize()
t2 = time.time()
x = np.array(wl.prob.p)
eng.eval_fd_host(x)
t3 = time.time()
print("build %.3fs compile(trace+create+jit) %.3fs first eval_fd_host %.3fs" % (t1 - t0, t2 - t1, t3 - t2))
for name, fn in (("eval_host", eng.eval_host), ("eval_fd_host", eng.eval_fd_host)):
    for _ in range(5):
        fn(x)
    t = time.time()
    for _ in range(200):
        fn(x)
    print("%-14s %.1f us/call" % (name, (time.time() - t) / 200 * 1e6))
fun, cons, jac = wl.prob._device_callables(wl.obj)
t = time.time()
for k in range(200):
    x[0] += 1e-9
    cons[0]["jac"](x); cons[1]["jac"](x); jac(x); fun(x); cons[0]["fun"](x); cons[1]["fun"](x)
print("6 callbacks at a new x: %.1f us" % ((time.time() - t) / 200 * 1e6))
