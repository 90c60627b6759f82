"""Device-backed vs host-backed Problem.solve on Goddard-50 (development aid)."""
import contextlib
import io
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import OpenGoddard.optimize as api
from opengoddard_b200 import workloads
from oracle import og_numpy

outer = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for backend in ("cuda", "host"):
    os.environ["OGB200_BACKEND"] = backend
    wl = workloads.build("cfg2_goddard50", api)
    wl.prob.maxIterator = outer
    buf = io.StringIO()
    t0 = time.time()
    with contextlib.redirect_stdout(buf):
        wl.prob.solve(wl.obj, ftol=1e-10)
    dt = time.time() - t0
    wo = workloads.build("cfg2_goddard50", og_numpy)
    p = np.array(wl.prob.p)
    ceq = wo.prob.eval_equality(p.copy(), wo.obj)
    cin = wo.prob.eval_inequality(p.copy(), wo.obj)
    msgs = [l for l in buf.getvalue().splitlines() if "Current function value" in l or "Iterations" in l or "terminated" in l or "limit" in l]
    if backend == "cuda":
        print("device launches", wl.prob._engine.launches)
    print(backend, "time %.2fs" % dt, "h(tf)=%.7f" % wl.prob.states_all_section(0)[-1], "max|ceq|=%.2e" % np.abs(ceq).max(),
          "min cineq=%.2e" % cin.min(), msgs[-3:])
