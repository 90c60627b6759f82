"""Quick device timing of the hot path (development aid; bench.py is the contract)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import OpenGoddard.optimize as api
from opengoddard_b200 import workloads

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_goddard50"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
wl = workloads.build(name, api)
eng = wl.prob.compile(wl.obj)
P = torch.as_tensor(workloads.make_batch(wl, min(B, 512)), device="cuda")
P = P.repeat((B + P.shape[0] - 1) // P.shape[0], 1)[:B].contiguous()
c = torch.empty((B, eng.nrows), dtype=torch.float64, device="cuda")
J = torch.empty((B, eng.nvars, eng.nrows), dtype=torch.float64, device="cuda")
info = eng.info
print(name, "B", B, "n", eng.nvars, "M", eng.nrows, "G", info.group_cols, "TC", info.tile_cols,
      "smem", info.smem_bytes, "ctas/sm", info.ctas_per_sm)
for _ in range(3):
    eng.eval_fd(P, out_c=c, out_J=J)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record()
for _ in range(K):
    eng.eval_fd(P, out_c=c, out_J=J)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
bytes_per_eval = 8 * eng.nvars + 8 * eng.nrows * (eng.nvars + 1)
print("eval_fd: %.3f ms/step  %.3e evals/s  %.1f GB/s algorithmic (%.3f of 6551.4)" % (
    ms, B / ms * 1e3, B * bytes_per_eval / ms / 1e6, B * bytes_per_eval / ms / 1e6 / 6551.4))
e0.record()
for _ in range(K):
    eng.dx_gemm(P)
e1.record()
torch.cuda.synchronize()
print("dx_gemm: %.4f ms" % (e0.elapsed_time(e1) / K))
e0.record()
for _ in range(K):
    eng.eval(P, out=c)
e1.record()
torch.cuda.synchronize()
print("eval   : %.4f ms" % (e0.elapsed_time(e1) / K))
# write-only bandwidth of this GPU for reference (memset of J)
e0.record()
for _ in range(K):
    J.zero_()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print("memset of J: %.3f ms  %.1f GB/s" % (ms, J.numel() * 8 / ms / 1e6))
