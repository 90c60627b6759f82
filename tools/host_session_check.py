"""Host-buffer session: bit-exactness of every transport mode against the device-resident
result, and wall time per mode.   python tools/host_session_check.py [workload] [B] [chunk] [threads]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import OpenGoddard.optimize as api
from opengoddard_b200 import workloads

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2_goddard50"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
threads = int(sys.argv[4]) if len(sys.argv) > 4 else 0
wl = workloads.build(cfg, api)
eng = wl.prob.compile(wl.obj, device="cuda:0")
n, M = eng.nvars, eng.nrows
P = workloads.make_batch(wl, B)
c_d, J_d = eng.eval_fd(torch.from_numpy(P).cuda())
torch.cuda.synchronize()
c_ref, J_ref = c_d.cpu().numpy(), J_d.cpu().numpy()
lin = eng.jac_pattern()
nz = np.flatnonzero((J_ref.reshape(B, -1) != 0).any(axis=0))
print(cfg, "B", B, "n", n, "M", M, "nnz", len(lin), "of", n * M, "observed nonzero positions", len(nz),
      "subset:", bool(np.isin(nz, lin).all()))
S = eng.host_session(B, chunk=chunk, threads=threads)
Pp = torch.from_numpy(P).pin_memory()
c = np.empty((B, M)); J = np.full((B, n, M), np.nan)
for mode in ("dense", "keep_zeros", "packed", "dma", "dense"):
    Jbuf = np.empty((B, len(lin))) if mode == "packed" else J
    if mode == "dma":
        Jbuf = torch.empty((B, n, M), dtype=torch.float64).pin_memory()
    for src, name in ((P, "pageable p"), (Pp, "pinned p")):
        ts = []
        for rep in range(4):
            t0 = time.perf_counter()
            S.eval_fd(src, c, Jbuf, mode=mode)
            ts.append(time.perf_counter() - t0)
        st = S.stats()
        Jh = Jbuf.numpy() if mode == "dma" else Jbuf
        ok = (c == c_ref).all() and ((Jh == J_ref.reshape(B, -1)[:, lin]).all() if mode == "packed" else (Jh == J_ref).all())
        print("%-10s %-10s exact=%s best %.2f ms (%.0f evals/s) first-chunk %.2f ms chunks %d x %d threads %d d2h %.1f MB" % (
            mode, name, ok, min(ts) * 1e3, B / min(ts), st.ms_first_chunk, st.nchunks, st.chunk, st.threads, st.d2h_bytes / 1e6))
