"""Device SQP on the GPU against the same code run by one serial host thread (tests/emu) on the host-compiled
evaluator, a few iterations from the first make_batch starts.    python tools/sqp_vs_emu.py workload starts maxiter"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import OpenGoddard.optimize as api  # noqa: E402
from opengoddard_b200 import tape, workloads  # noqa: E402
from oracle import og_numpy  # noqa: E402
from tests.emu.emu import EmuProblem, EmuSqp  # noqa: E402

name, S, maxiter = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
wl = workloads.build(name, api)
wo = workloads.build(name, og_numpy)
lb, ub = og_numpy.bounds_arrays(wo.prob)
eng = wl.prob.compile(wl.obj)
P0 = np.clip(workloads.make_batch(wl, S), lb, ub)
t0 = time.perf_counter()
with eng.device_sqp(S, 1e-6, maxiter) as dq:
    dev = dq.solve(P0)
    sc = dq.k.scalars(S)
print("gpu: %.1f s  status %s nit %s nfev %s h4 %s" % (time.perf_counter() - t0, dev["status"], dev["nit"], dev["nfev"], np.round(sc["h4"], 6)))
lin = eng.jac_pattern().astype(np.int64)
n, M = eng.nvars, eng.nrows
colptr = np.searchsorted(lin, np.arange(n + 1) * M).astype(np.int32)
prow = (lin % M).astype(np.int32)
ep = EmuProblem(tape.build_ir(wl.prob, wl.obj), lb, ub)
emu = EmuSqp(n, M - 1, eng.meq, colptr, prow, lb, ub, 1e-6, maxiter, S)
X = P0.copy()
t0 = time.perf_counter()
for _ in range(400):
    cc, JJ = ep.eval_fd(X)
    emu.step(X, cc, np.ascontiguousarray(JJ.reshape(S, -1)[:, lin]))
    modes = np.array([emu.scalars(b)["mode"] for b in range(S)])
    if not (np.abs(modes) == 1).any():
        break
print("emu: %.1f s  status %s nit %s nfev %s h4 %s" % (time.perf_counter() - t0, modes.astype(int), [int(emu.scalars(b)["iter"]) for b in range(S)],
                                                     [int(emu.scalars(b)["nfev"]) for b in range(S)], [round(float(emu.scalars(b)["h4"]), 6) for b in range(S)]))
print("max |x_gpu - x_emu| per instance:", np.abs(dev["x"] - X).max(axis=1))
