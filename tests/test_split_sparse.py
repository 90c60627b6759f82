"""Round-2 device paths, all through the C ABI and all bit-identical to the fused sweep kernel:

* ogb_eval_sparse: c + the packed non-zeros of the FD Jacobian (sweep kernel, packed output),
* ogb_densify (K2b): packed values -> dense J,
* ogb_eval_fd as the split pipeline (K1 -> K2a on an internal stream, K2b on the caller's, chunked),
* ogb_host_eval_fd_scatter: packed values scattered straight into an SQP driver's C / g buffers,
* the ring of work-item tickets (many launches in flight), one-warp CTAs with odd sizes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CFGS = ["cfg1_brachistochrone20", "cfg2_goddard50", "cfg3_goddard_knot30x2", "cfg4_polar3x40",
        "cfg5_lowthrust128", "ex05_goddard_knot25x2", "ex09_polar_tsto20x2", "ex10_lowthrust100",
        "edge_table_lookup", "edge_stress_mixed", "edge_all_ops", "edge_nonautonomous", "edge_nonautonomous_big", "edge_picked_dynamics"]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch


@pytest.mark.parametrize("name", CFGS)
def test_sparse_densify_and_split_are_bit_identical_to_the_fused_kernel(torch_cuda, api, name):
    t = torch_cuda
    from opengoddard_b200 import workloads
    wl = workloads.build(name, api)
    eng = wl.prob.compile(wl.obj)
    B = 45
    P = workloads.make_batch(wl, B)
    eng.set_option(9, 0)                                        # the fused sweep kernel (round 1)
    c0, J0 = eng.eval_fd(P)
    lin = t.from_numpy(eng.jac_pattern().astype(np.int64)).to(J0.device)
    for jit in (1, 0):                                          # NVRTC build and the tape interpreter
        eng.set_option(2, jit)
        c1, vals = eng.eval_sparse(P)
        assert t.equal(c1, c0)
        assert t.equal(vals, J0.reshape(B, -1)[:, lin])
        assert t.equal(eng.densify(vals), J0)
        for streaming in (1, 0):
            eng.set_option(11, streaming)
            assert t.equal(eng.densify(vals), J0)
        for chunk in (0, 7, 16, 45, 64):                        # ragged last chunk, one chunk, chunk > batch
            eng.set_option(9, 1)
            eng.set_option(10, chunk)
            J2 = t.full_like(J0, float("nan"))
            c2, J2 = eng.eval_fd(P, out_J=J2)
            assert t.equal(c2, c0) and t.equal(J2, J0), (jit, chunk)
        eng.set_option(9, 0)
    eng.set_option(2, 1)
    eng.set_option(0, 1)                                        # generic column code, packed output
    _, v3 = eng.eval_sparse(P)
    assert t.equal(v3, J0.reshape(B, -1)[:, lin])


def test_split_pipeline_full_size_on_a_side_stream(torch_cuda, api):
    """Goddard-50 x 4096 (auto chunks, auto split), launched on a non-default stream, twice back to back."""
    t = torch_cuda
    from opengoddard_b200 import workloads
    wl = workloads.build("cfg2_goddard50", api)
    eng = wl.prob.compile(wl.obj)
    P = t.from_numpy(workloads.make_batch(wl, 4096)).cuda()
    eng.set_option(9, 0)
    c0, J0 = eng.eval_fd(P)
    eng.set_option(9, 1)                                        # the split pipeline, automatic chunk size
    side = t.cuda.Stream()
    side.wait_stream(t.cuda.current_stream())
    with t.cuda.stream(side):
        c1, J1 = eng.eval_fd(P)
        c2, J2 = eng.eval_fd(P, out_c=t.empty_like(c0), out_J=t.full_like(J0, 1.0))
    side.synchronize()
    assert t.equal(c1, c0) and t.equal(J1, J0) and t.equal(c2, c0) and t.equal(J2, J0)
    launches0 = eng.launches
    eng.eval_fd(P, out_c=c1, out_J=J1)
    assert eng.launches - launches0 >= 6                       # several chunks x (K1, K2a, K2b)


def test_many_launches_in_flight_share_the_ticket_ring(torch_cuda, api):
    """More launches than ticket slots, alternating between two streams without host synchronisation."""
    t = torch_cuda
    from opengoddard_b200 import workloads
    wl = workloads.build("cfg3_goddard_knot30x2", api)
    eng = wl.prob.compile(wl.obj)
    eng.set_option(9, 0)
    P = t.from_numpy(workloads.make_batch(wl, 700)).cuda()      # more work items than resident CTAs
    c0, J0 = eng.eval_fd(P)
    t.cuda.synchronize()
    s1, s2 = t.cuda.Stream(), t.cuda.Stream()
    outs = []
    for i in range(80):
        with t.cuda.stream(s1 if i % 2 == 0 else s2):
            if i % 3 == 0:
                outs.append(("c", eng.eval(P)))
            else:
                outs.append(("v", eng.eval_sparse(P, out_c=t.empty_like(c0))[1]))
    t.cuda.synchronize()
    lin = t.from_numpy(eng.jac_pattern().astype(np.int64)).cuda()
    ce = eng.eval(P)
    for kind, val in outs:
        if kind == "c":
            assert t.equal(val, ce)
        else:
            assert t.equal(val, J0.reshape(700, -1)[:, lin])


@pytest.mark.parametrize("name", ["cfg1_brachistochrone20", "cfg2_goddard50", "cfg3_goddard_knot30x2"])
def test_one_warp_ctas(torch_cuda, api, name):
    """OGB_OPT_THREADS = 32: thread 0 issues the TMA loads AND fetches the odd head / tail doubles
    (n = 81 / 201 / 242: odd and even decision-vector lengths, odd rows of p start 8-byte aligned)."""
    t = torch_cuda
    from opengoddard_b200 import workloads
    wl = workloads.build(name, api)
    eng = wl.prob.compile(wl.obj)
    eng.set_option(9, 0)
    P = workloads.make_batch(wl, 11)
    c0, J0 = eng.eval_fd(P)
    for jit in (1, 0):
        eng.set_option(2, jit)
        eng.set_option(1, 32)
        c1, J1 = eng.eval_fd(P)
        assert t.equal(c1, c0) and t.equal(J1, J0)
        eng.set_option(1, 256)


@pytest.mark.parametrize("name", ["cfg2_goddard50", "cfg3_goddard_knot30x2", "ex10_lowthrust100"])
def test_host_scatter_into_sqp_buffers(torch_cuda, api, name):
    """ogb_host_eval_fd_scatter: per instance a zero-initialised Fortran (ld, n) matrix and a gradient
    vector; after the call they hold exactly J[:, :m].T and J[:, M-1] of the device-resident result."""
    from opengoddard_b200 import workloads
    wl = workloads.build(name, api)
    eng = wl.prob.compile(wl.obj)
    B, n, M = 19, eng.nvars, eng.nrows
    m = M - 1
    P = workloads.make_batch(wl, B)
    c_d, J_d = eng.eval_fd(P)
    torch_cuda.cuda.synchronize()
    c_ref, J_ref = c_d.cpu().numpy(), J_d.cpu().numpy()
    S = eng.host_session(32, chunk=5, threads=2)
    for ld in (m, m + 3):
        Cs = [np.zeros((ld, n), order="F") for _ in range(B)]
        gs = [np.zeros(n) for _ in range(B)]
        c = np.empty((B, M))
        S.eval_fd_scatter(P, c, [a.ctypes.data for a in Cs], ld, m, [g.ctypes.data for g in gs])
        assert (c == c_ref).all()
        for b in range(B):
            assert (Cs[b][:m] == J_ref[b, :, :m].T).all() and (Cs[b][m:] == 0).all()
            assert (gs[b] == J_ref[b, :, m]).all()
    # constraint rows only, no gradient target; a second call into the same buffers rewrites the pattern
    Cs = [np.zeros((m, n), order="F") for _ in range(B)]
    for _ in range(2):
        S.eval_fd_scatter(P, c, [a.ctypes.data for a in Cs], m, m, None)
    assert all((Cs[b] == J_ref[b, :, :m].T).all() for b in range(B))
    S.close()


def test_solve_batch_uses_the_scatter_transport(torch_cuda, api):
    """The batched SQP driver on the device evaluator: identical iterates whether the Jacobians reach the
    SLSQP states through the scatter transport (default) or through dense host arrays."""
    from opengoddard_b200 import sqp, workloads
    wl = workloads.build("cfg1_brachistochrone20", api)
    eng = wl.prob.compile(wl.obj)
    lb, ub = wl.prob.bounds_arrays()
    P = workloads.make_batch(wl, 6)
    ev = eng.host_evaluator()
    a = sqp.slsqp_batch(ev, P, lb, ub, eng.meq, eng.mineq, maxiter=8)

    class DenseOnly:
        eval = staticmethod(ev.eval)
        eval_fd = staticmethod(lambda X: ev.eval_fd(X))
    b = sqp.slsqp_batch(DenseOnly(), P, lb, ub, eng.meq, eng.mineq, maxiter=8)
    for key in ("x", "fun", "status", "nit"):
        assert np.array_equal(a[key], b[key]), key
    c2 = sqp.slsqp_batch(ev, P, lb, ub, eng.meq, eng.mineq, maxiter=8, processes=2)
    from threadpoolctl import threadpool_limits
    with threadpool_limits(1):
        d = sqp.slsqp_batch(ev, P, lb, ub, eng.meq, eng.mineq, maxiter=8)
    assert np.array_equal(c2["x"], d["x"])
