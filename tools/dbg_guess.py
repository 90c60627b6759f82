import sys; sys.path.insert(0, '/root/repo')
import numpy as np
import OpenGoddard.optimize as api
from opengoddard_b200 import workloads
from oracle import og_numpy
wl = workloads.build("cfg2_goddard50", api)
prob = wl.prob
eng = prob.compile(wl.obj)
params = np.array([[[0.3, 2.7, 0, 0]]])
P = prob.guess_batch([("linear", ("state", 0), None)], params, wl.obj).cpu().numpy()
t = prob.time_all_section
ref = og_numpy.Guess.linear(t, 0.3, 2.7)
got = P[0][:50]
d = got - ref
print("max diff", np.abs(d).max(), "n diff", (d != 0).sum())
i = np.argmax(np.abs(d)); print(i, repr(got[i]), repr(ref[i]), repr(t[i]), repr(t[0]), repr(t[-1]))
slope = (2.7 - 0.3) / (t[-1] - t[0]); print("host formula", repr(slope * (t[i] - t[0]) + 0.3))
import math
print("fma formula", repr(math.fma(slope, t[i] - t[0], 0.3)) if hasattr(math, "fma") else "n/a")
