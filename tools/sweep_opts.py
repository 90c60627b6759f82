"""Timing sweep over kernel options (development aid)."""
import sys

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
bpe = 8 * eng.nvars + 8 * eng.nrows * (eng.nvars + 1)


def timeit(label):
    for _ in range(3):
        eng.eval_fd(P, out_c=c, out_J=J)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.eval_fd(P, out_c=c, out_J=J)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%-28s %.3f ms  %.3e evals/s  frac %.3f" % (label, ms, B / ms * 1e3, B * bpe / ms / 1e6 / 6551.4))


print(name, B, "smem", eng.info.smem_bytes, "ctas", eng.info.ctas_per_sm, "warps", eng.info.tile_cols)
timeit("default (jit=%d)" % eng.info.jit)
eng.set_option(2, 0); timeit("interpreter kernel"); eng.set_option(2, 1)
eng.set_option(4, 1); timeit("D.X fused into K2"); eng.set_option(4, 0)
eng.set_option(5, 0); timeit("static item assignment"); eng.set_option(5, 1)
eng.set_option(7, 0); timeit("no auto split"); eng.set_option(7, 1)
eng.set_option(0, 1); timeit("generic columns"); eng.set_option(0, 0)
for thr in (64, 128, 192, 256):
    try:
        eng.set_option(1, thr)
    except Exception as e:
        print("threads", thr, "->", e)
        continue
    timeit("threads %d" % thr)
