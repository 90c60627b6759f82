"""Small device-SQP run for compute-sanitizer (memcheck / racecheck / initcheck): brachistochrone-20 and Goddard-50,
a few starts, three iterations, FD and exact Jacobians, plus the relaxed (slack-variable) QP path.
    compute-sanitizer --tool memcheck python tools/sanitize_sqp.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import OpenGoddard.optimize as api
from opengoddard_b200 import workloads

for name in ("cfg1_brachistochrone20", "cfg2_goddard50", "ex09_polar_tsto20x2"):
    wl = workloads.build(name, api)
    eng = wl.prob.compile(wl.obj)
    P0 = np.vstack([np.asarray(wl.prob.p)[None], workloads.make_batch(wl, 5)])
    with eng.device_sqp(6, 1e-6, 3) as dq:
        a = dq.solve(P0)
        b = dq.solve(P0, exact=True)
        sc = dq.k.scalars(6)
    torch.cuda.synchronize()
    print(name, "status", a["status"], b["status"], "nit", a["nit"], "relaxed QPs used:", int((sc["h4"] != 1.0).sum()))
