"""What a pure-write kernel reaches on this GPU (the sweep kernel's traffic is ~100 % writes, while
MEASURED_PEAKS.json's hbm_gbs is a copy: half reads).  Times torch's fill kernel and a copy on
buffers the size of the Goddard-50 x 4096 Jacobian."""
import torch
n = 3_031_531_520 // 8
a = torch.empty(n, dtype=torch.float64, device="cuda")
b = torch.empty(n, dtype=torch.float64, device="cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: a.zero_())
print("fill (write only) 3.03 GB: %.3f ms = %.0f GB/s" % (ms, n * 8 / ms / 1e6))
ms = t(lambda: a.fill_(1.5))
print("fill_(1.5)        3.03 GB: %.3f ms = %.0f GB/s" % (ms, n * 8 / ms / 1e6))
ms = t(lambda: b.copy_(a))
print("copy (read+write) 6.06 GB: %.3f ms = %.0f GB/s" % (ms, 2 * n * 8 / ms / 1e6))
