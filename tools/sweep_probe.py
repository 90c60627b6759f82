"""Where K2's time goes: the sweep kernel timed with parts of the column phase switched off
(OGB_OPT_PROBE_MODE).  Development aid; results of probe modes are not Jacobians."""
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
DX = eng.dx_gemm(P, clip=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
bpe = 8 * eng.nvars + 8 * eng.nrows * (eng.nvars + 1)


def timeit(label, K=20):
    for _ in range(3):
        eng.sweep_fd(P, DX, c, J)
    ts = []
    for _ in range(K):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.sweep_fd(P, DX, c, J); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    print("%-34s avg %.3f ms  min %.3f ms   %.0f GB/s algorithmic" % (label, ms, min(ts), B * bpe / ms / 1e6))


print(name, B, "smem", eng.info.smem_bytes, "ctas/sm", eng.info.ctas_per_sm)
timeit("K2 full")
for mode, label in ((2, "no zero stream (values only)"), (3, "zero stream only"), (4, "no column output (phases 1-3)"), (5, "zero stream, no tapes/assembly")):
    eng.set_option(8, mode)
    timeit(label)
eng.set_option(8, 0)
timeit("K2 full again")
for _ in range(3): J.zero_()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); J.zero_(); e1.record(); torch.cuda.synchronize()
print("torch fill of J: %.3f ms %.0f GB/s" % (e0.elapsed_time(e1), J.numel() * 8 / e0.elapsed_time(e1) / 1e6))
