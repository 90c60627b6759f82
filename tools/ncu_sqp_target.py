"""Target for one ncu capture of the SLSQP step kernel: Goddard-50 x 592 starts (two blocks per SM, one wave), the
first round (every instance solves its first QP).
    ncu --set full --clock-control none --import-source on -k regex:ogb_sqp_step -c 1 -o gpurun_out/r2_sqp python tools/ncu_sqp_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import OpenGoddard.optimize as api  # noqa: E402
from opengoddard_b200 import workloads  # noqa: E402

wl = workloads.build("cfg2_goddard50", api)
eng = wl.prob.compile(wl.obj)
S = 592
P0 = workloads.make_batch(wl, S)
with eng.device_sqp(S, 1e-6, 2) as dq:
    res = dq.solve(P0, max_rounds=1)
    torch.cuda.synchronize()
print("rounds", res["rounds"])
