"""A one-iteration device-SQP run small enough for compute-sanitizer --tool racecheck."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import OpenGoddard.optimize as api
from opengoddard_b200 import workloads

wl = workloads.build("cfg1_brachistochrone20", api)
eng = wl.prob.compile(wl.obj)
P0 = np.vstack([np.asarray(wl.prob.p)[None], workloads.make_batch(wl, 1)])
with eng.device_sqp(2, 1e-6, 1) as dq:
    a = dq.solve(P0)
torch.cuda.synchronize()
print("status", a["status"], "nit", a["nit"], "nfev", a["nfev"])
