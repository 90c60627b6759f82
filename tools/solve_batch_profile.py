"""Where a batched multi-start (Problem.solve_batch) spends its time: device evaluations through the
host session vs SciPy's SLSQP steps on the host, for 1 / 8 / 16 stepping threads."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
import OpenGoddard.optimize as api
from opengoddard_b200 import workloads, sqp

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_goddard50"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
maxiter = int(sys.argv[3]) if len(sys.argv) > 3 else 10
wl = workloads.build(name, api)
eng = wl.prob.compile(wl.obj)
P0 = workloads.make_batch(wl, B)
lb, ub = wl.prob.bounds_arrays()


class Timed:
    def __init__(self, ev):
        self.ev, self.t, self.calls, self.rows = ev, 0.0, 0, 0

    def eval(self, X):
        t0 = time.perf_counter(); r = self.ev.eval(X); self.t += time.perf_counter() - t0
        self.calls += 1; self.rows += len(X); return r

    def eval_fd(self, X):
        t0 = time.perf_counter(); r = self.ev.eval_fd(X); self.t += time.perf_counter() - t0
        self.calls += 1; self.rows += len(X); return r


for threads, procs in ((1, 0), (8, 0), (1, 4), (1, 8), (1, 16)):
    ev = Timed(eng.host_evaluator())
    ev.eval_fd(P0[:2])                                  # warm
    ev.t = 0.0; ev.calls = 0; ev.rows = 0
    t0 = time.perf_counter()
    res = sqp.slsqp_batch(ev, P0, lb, ub, eng.meq, eng.mineq, ftol=1e-6, maxiter=maxiter, threads=threads, processes=procs)
    dt = time.perf_counter() - t0
    nit = int(res["nit"].sum())
    print("%s B %d maxiter %d threads %2d processes %2d: %.2f s total, device evaluations %.3f s in %d calls (%d rows), "
          "host SLSQP %.2f s = %.1f ms per instance-iteration (%d iterations)" % (
              name, B, maxiter, threads, procs, dt, ev.t, ev.calls, ev.rows, dt - ev.t, 1e3 * (dt - ev.t) / max(1, nit), nit))
