"""DeviceProblem.autotune over a wider set of CTA sizes (development aid)."""
import sys
sys.path.insert(0, ".")
import OpenGoddard.optimize as api
from opengoddard_b200 import workloads
for spec in sys.argv[1:] or ["cfg2_goddard50:4096", "cfg2_goddard50:1024", "cfg3_goddard_knot30x2:4096"]:
    name, B = spec.split(":")
    wl = workloads.build(name, api)
    eng = wl.prob.compile(wl.obj)
    P = workloads.make_batch(wl, min(int(B), 512))
    import numpy as np
    P = np.tile(P, ((int(B) + len(P) - 1) // len(P), 1))[:int(B)]
    t = eng.autotune(P, candidates=(256, 128, 192, 384, 64), min_gain=0.0, reps=8)
    print(name, B, {k: round(v, 4) for k, v in t.items()}, "->", eng.tuned_threads)
