"""Target of the round-2 `ncu --set full` capture: two warm-ups of every device path, then ONE launch of each
on Goddard-50 x 4096 -- K1, K2 (dense), K1, K2a (packed FD), K1, K2 exact, K2b (densify).
    ncu --set full --clock-control none --import-source on -k regex:ogb_ -s 14 -c 7 -o gpurun_out/r2_kernels python tools/ncu_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import OpenGoddard.optimize as api  # noqa: E402
from opengoddard_b200 import workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_goddard50"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
wl = workloads.build(name, api)
eng = wl.prob.compile(wl.obj)
eng.jac_pattern()                      # (the structure probe launches K1 + K2 once: before the counted launches)
P = torch.from_numpy(workloads.make_batch(wl, B)).cuda()
n, M = eng.nvars, eng.nrows
c = torch.empty((B, M), dtype=torch.float64, device="cuda")
J = torch.empty((B, n, M), dtype=torch.float64, device="cuda")
vals = torch.empty((B, eng.nnz), dtype=torch.float64, device="cuda")
for rep in range(3):                   # 2 warm-ups + the captured pass: 7 ogb_ launches per pass
    eng.eval_fd(P, out_c=c, out_J=J)
    eng.eval_sparse(P, out_c=c, out_vals=vals)
    eng.eval_exact(P, out_c=c, out_vals=vals)
    eng.densify(vals, out_J=J)
    torch.cuda.synchronize()
print("done", name, B)
