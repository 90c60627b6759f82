"""Where does a device-backed Problem.solve spend its time? (development aid)"""
import cProfile
import contextlib
import io
import pstats
import sys
import time

sys.path.insert(0, ".")
import torch
import OpenGoddard.optimize as api
from opengoddard_b200 import workloads

torch.zeros(1, device="cuda")
wl = workloads.build("cfg2_goddard50", api)
wl.prob.maxIterator = 2
pr = cProfile.Profile()
buf = io.StringIO()
t0 = time.time()
with contextlib.redirect_stdout(buf):
    pr.enable()
    wl.prob.solve(wl.obj, ftol=1e-10)
    pr.disable()
print("solve wall %.2fs" % (time.time() - t0), "launches", wl.prob._engine.launches)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22)
print(s.getvalue()[-3800:])
