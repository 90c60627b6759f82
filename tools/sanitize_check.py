"""Small run for compute-sanitizer: device path (JIT, interpreter, generic columns), K3 and the host
session on a few problems.   compute-sanitizer --tool memcheck python tools/sanitize_check.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import OpenGoddard.optimize as api
from opengoddard_b200 import workloads

for name in ("cfg2_goddard50", "cfg3_goddard_knot30x2", "edge_stress_small", "cfg5_lowthrust128"):
    wl = workloads.build(name, api)
    eng = wl.prob.compile(wl.obj)
    P = workloads.make_batch(wl, 5)
    c1, J1 = eng.eval_fd(P)
    eng.set_option(2, 0)
    c0, J0 = eng.eval_fd(P)
    eng.set_option(0, 1)
    cg, Jg = eng.eval_fd(P)
    eng.set_option(0, 0); eng.set_option(2, 1)
    S = eng.host_session(8, chunk=2, threads=2)
    ch, Jh = S.eval_fd(P, mode="dense")
    cp, V = S.eval_fd(P, mode="packed")
    S.close()
    torch.cuda.synchronize()
    print(name, bool(torch.equal(J0, J1)), bool(torch.equal(Jg, J1)), bool((Jh == J1.cpu().numpy()).all()),
          bool((V == J1.cpu().numpy().reshape(5, -1)[:, eng.jac_pattern()]).all()))
