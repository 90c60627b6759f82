"""Work-item size sweep (OGB_OPT_GROUP_COLS) -- development aid."""
import sys
import torch
sys.path.insert(0, ".")
import OpenGoddard.optimize as api
from opengoddard_b200 import workloads, capi
import ctypes as C
name = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
wl = workloads.build(name, api)
eng = wl.prob.compile(wl.obj)
P = torch.as_tensor(workloads.make_batch(wl, 256), device="cuda").repeat((B + 255) // 256, 1)[:B].contiguous()
c = torch.empty((B, eng.nrows), dtype=torch.float64, device="cuda")
J = torch.empty((B, eng.nvars, eng.nrows), dtype=torch.float64, device="cuda")
bpe = 8 * eng.nvars + 8 * eng.nrows * (eng.nvars + 1)
n = eng.nvars
caps = sorted(set([224] + [-(-n // k) for k in (1, 2, 3, 4, 5, 6, 8)]))
for cap in caps:
    try:
        eng.set_option(6, cap)
    except Exception as e:
        print("cap", cap, "->", str(e)[:80]); continue
    eng.b.problem_info_get(eng.h, C.byref(eng.info))
    for _ in range(3):
        eng.eval_fd(P, out_c=c, out_J=J)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.eval_fd(P, out_c=c, out_J=J)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%s cap %4d -> G %4d smem %6d ctas %d : %.3f ms frac %.3f" % (name, cap, eng.info.group_cols, eng.info.smem_bytes,
          eng.info.ctas_per_sm, ms, B * bpe / ms / 1e6 / 6551.4))
