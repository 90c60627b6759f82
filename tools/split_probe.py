"""Round-2 probe: the fused sweep kernel vs the split pipeline (K1 -> K2a packed sweep | K2b densify),
chunk sizes, streaming stores, and the pieces alone.  CUDA events, L2 flushed between repetitions.

    python tools/split_probe.py [workload] [batch]
"""
import json
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
P = torch.from_numpy(workloads.make_batch(wl, B)).cuda()
n, M = eng.nvars, eng.nrows
c = torch.empty((B, M), dtype=torch.float64, device="cuda")
J = torch.empty((B, n, M), dtype=torch.float64, device="cuda")
vals = torch.empty((B, eng.nnz), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
bytes_eval = 8 * n + 8 * M * (n + 1)


def timed(fn, reps=7):
    for _ in range(2):
        fn()
    best, tot = 1e9, 0.0
    for _ in range(reps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        best = min(best, ms)
        tot += ms
    return best, tot / reps


out = {"workload": name, "B": B, "n": n, "M": M, "nnz": eng.nnz, "GB": B * bytes_eval / 1e9}
eng.set_option(9, 0)
out["fused"] = timed(lambda: eng.eval_fd(P, out_c=c, out_J=J))
out["sparse(K1+K2a)"] = timed(lambda: eng.eval_sparse(P, out_c=c, out_vals=vals))
DX = eng.dx_gemm(P, clip=True)
out["K1"] = timed(lambda: eng.dx_gemm(P, out=DX, clip=True))
for streaming in (1, 0):
    eng.set_option(11, streaming)
    out["densify(streaming=%d)" % streaming] = timed(lambda: eng.densify(vals, out_J=J))
eng.set_option(11, 1)
eng.set_option(9, 1)
for chunk in (0, 64, 128, 256, 512, 1024, 2048):
    if chunk > B:
        continue
    eng.set_option(10, chunk)
    out["split(chunk=%d)" % chunk] = timed(lambda: eng.eval_fd(P, out_c=c, out_J=J))
eng.set_option(11, 0)
eng.set_option(10, 0)
out["split(auto, plain stores)"] = timed(lambda: eng.eval_fd(P, out_c=c, out_J=J))
for k, v in out.items():
    if isinstance(v, tuple):
        gbs = B * bytes_eval / (v[0] * 1e-3) / 1e9
        print("%-32s best %.4f ms  avg %.4f ms   %.0f GB/s (algorithmic, best)" % (k, v[0], v[1], gbs))
    else:
        print(k, v)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "split_probe_%s_b%d.json" % (name, B)), "w") as f:
    json.dump(out, f, indent=1)
