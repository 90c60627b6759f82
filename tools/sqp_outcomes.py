"""Outcome-level comparison of the two multi-start paths on the reference's problem families: the reference's own
initial guess plus make_batch starts, `solve_batch` with SciPy's core on the host (default) and with the SLSQP
iteration on the device (qp="device"), same ftol / maxiter / outer passes.   python tools/sqp_outcomes.py [starts]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import OpenGoddard.optimize as api  # noqa: E402
from opengoddard_b200 import workloads  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 7
names = sys.argv[2:] or ["cfg1_brachistochrone20", "cfg2_goddard50", "ex05_goddard_knot25x2", "cfg3_goddard_knot30x2",
                         "ex09_polar_tsto20x2"]


def violation(c, meq):
    v = np.maximum(-c, np.where(np.arange(c.shape[1]) < meq, c, 0.0))
    return v.sum(axis=1)


for name in names:
    wl = workloads.build(name, api)
    eng = wl.prob.compile(wl.obj)
    lb, ub = wl.prob.bounds_arrays()
    P0 = np.vstack([np.asarray(wl.prob.p, dtype=float)[None], workloads.make_batch(wl, S)])
    out = {}
    for qp in ("scipy", "device"):
        t0 = time.perf_counter()
        res = wl.prob.solve_batch(P0, wl.obj, maxiter=25, max_outer=3, qp=qp, processes=8 if qp == "scipy" else 0)
        dt = time.perf_counter() - t0
        c = eng.eval(np.clip(res["x"], lb, ub)).cpu().numpy()
        out[qp] = (res, violation(c[:, :-1], eng.meq), dt)
    print("== %s (n = %d, m = %d)" % (name, eng.nvars, eng.nrows - 1))
    for qp in ("scipy", "device"):
        res, viol, dt = out[qp]
        print("  %-6s %.1f s  status %s  nit %s" % (qp, dt, res["status"].tolist(), res["nit"].tolist()))
        print("         cost %s" % np.array2string(res["fun"], precision=6, max_line_width=200))
        print("         viol %s" % np.array2string(viol, precision=1, max_line_width=200))
