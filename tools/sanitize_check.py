"""Small run for compute-sanitizer: every device path of rounds 1 and 2 on a few problems -- the fused sweep kernel
(JIT, interpreter, generic columns, D.X fused and not, writer-warp mode, tail refinement, with and without
programmatic dependent launch), both forms of K1, the packed sweep (K2a), the exact mode,
densify (K2b), the split pipeline, K3, the host session (dense / packed / scatter), guess / jitter / trajectories.
    compute-sanitizer --tool memcheck python tools/sanitize_check.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import OpenGoddard.optimize as api
from opengoddard_b200 import workloads

for name in ("cfg2_goddard50", "cfg3_goddard_knot30x2", "edge_stress_small", "edge_nonautonomous", "cfg5_lowthrust128"):
    wl = workloads.build(name, api)
    eng = wl.prob.compile(wl.obj)
    B = 5
    P = workloads.make_batch(wl, B)
    lin = torch.from_numpy(eng.jac_pattern().astype(np.int64)).cuda()
    eng.set_option(4, 0)
    c1, J1 = eng.eval_fd(P)                       # K1 + K2
    eng.set_option(4, 1)
    cf, Jf = eng.eval_fd(P)                       # D.X inside K2
    eng.set_option(4, -1)
    eng.set_option(2, 0)
    c0, J0 = eng.eval_fd(P)                       # interpreter
    _, v0 = eng.eval_sparse(P)
    _, e0 = eng.eval_exact(P)
    eng.set_option(0, 1)
    cg, Jg = eng.eval_fd(P)                       # generic columns
    eng.set_option(0, 0); eng.set_option(2, 1)
    eng.set_option(12, 4)
    cw, Jw = eng.eval_fd(P)                       # writer warps
    eng.set_option(12, 0)
    eng.set_option(3, 1); eng.set_option(15, 200)
    ct, Jt = eng.eval_fd(P)                       # tail refinement (one persistent CTA: the last two instances refined)
    eng.set_option(3, 0); eng.set_option(15, 100)
    eng.set_option(13, 8)
    DXo = eng.dx_gemm(P, clip=True).clone()       # the round-1 K1
    eng.set_option(13, 0)
    DXn = eng.dx_gemm(P, clip=True)               # the latency-organised K1
    eng.set_option(14, 0)
    cq, Jq = eng.eval_fd(P)                       # plain launch of the sweep behind K1 (no PDL)
    eng.set_option(14, 1)
    _, v1 = eng.eval_sparse(P)                    # K2a (JIT)
    _, e1 = eng.eval_exact(P)                     # exact (JIT)
    Jd = eng.densify(v1)                          # K2b
    eng.set_option(9, 1); eng.set_option(10, 2)
    cs, Js = eng.eval_fd(P)                       # split pipeline, 3 chunks
    eng.set_option(9, 0)
    vk = eng.pack(J1)                             # K3
    S = eng.host_session(8, chunk=2, threads=2)
    ch, Jh = S.eval_fd(P, mode="dense")
    cp, V = S.eval_fd(P, mode="packed")
    m = eng.nrows - 1
    Cs = [np.zeros((m, eng.nvars), order="F") for _ in range(B)]
    gs = [np.zeros(eng.nvars) for _ in range(B)]
    S.eval_fd_scatter(P, np.empty((B, eng.nrows)), [a.ctypes.data for a in Cs], m, m, [g.ctypes.data for g in gs])
    S.close()
    X = wl.prob.make_starts(7, wl.obj, seed=3)
    T = eng.trajectories(X)
    G = eng.guess_batch([("linear", 0, None), ("cubic", 1, 0)], np.ones((3, 2, 4)), wl.prob.time_all_section)
    torch.cuda.synchronize()
    Jn = J1.cpu().numpy()
    ok = [torch.equal(Jt, J1), torch.equal(DXo, DXn), torch.equal(Jq, J1), torch.equal(J0, J1), torch.equal(Jf, J1), torch.equal(Jg, J1), torch.equal(Jw, J1), torch.equal(Jd, J1),
          torch.equal(Js, J1), torch.equal(v0, v1), torch.equal(v1, J1.reshape(B, -1)[:, lin]), torch.equal(vk, v1),
          bool((Jh == Jn).all()), bool((V == v1.cpu().numpy()).all()),
          all((Cs[b] == Jn[b, :, :m].T).all() and (gs[b] == Jn[b, :, m]).all() for b in range(B)),
          bool(((e0 - e1).abs() <= 1e-12 * e1.abs().max()).all()), bool(torch.isfinite(T).all() and torch.isfinite(G).all())]
    print(name, all(bool(v) for v in ok), [bool(v) for v in ok])
