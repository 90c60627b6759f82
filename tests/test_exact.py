"""Exact Jacobian mode: analytic D-block + forward-mode tangents of the traced
tapes, sparse (packed) output.  Opt-in and NOT reference parity -- the reference differentiates by forward
differences -- so the checks are: against complex-step differentiation of the oracle (exact to rounding;
central 3-point differences where a callback compares values, which complex numbers cannot), against the
FD mode at the FD tolerance, and structural (same pattern as the FD Jacobian).  CPU: the g++ emulation of
the device arithmetic; GPU: the kernels through the C ABI."""
import numpy as np
import pytest

from opengoddard_b200 import tape, workloads
from oracle import og_numpy

EXACT_RTOL_CS = 1e-10        # |J - J_cs| <= EXACT_RTOL_CS * rowmax|J_cs|   (complex step: no truncation error)
EXACT_RTOL_3PT = 2e-7        # ... against central differences (truncation + cancellation ~1e-8)
FD_VS_EXACT_RTOL = 1e-4      # the forward-difference Jacobian against the exact one, row-scaled: the FD truncation
#                              error h |f''| / 2 with h = 1.5e-8 reaches ~1e-5 of the row maximum on rows with
#                              curvature (Goddard's exp / 1/h^2 terms: 7e-6 measured) -- this is the noise the
#                              exact mode removes, not a defect of either

CFGS = ["cfg1_brachistochrone20", "cfg2_goddard50", "cfg3_goddard_knot30x2", "cfg4_polar3x40",
        "cfg5_lowthrust128", "ex09_polar_tsto20x2", "edge_table_lookup", "edge_all_ops", "edge_nonautonomous",
        "edge_nonautonomous_big", "edge_picked_dynamics"]


def oracle_jacobian(wo, x, lb, ub):
    """(J (M, n), method): derivative of [c_eq; c_ineq; cost] of the numpy oracle at x.  Complex step with
    h = 1e-30 (no truncation error: the 'cs' of SciPy uses h = 1.5e-8, whose h^2 f''' / 6 term reaches 1e-9 of
    the row maximum on the exp(-altitude / 7 km) rows of the polar problems); central differences where a
    callback is not complex-analytic (comparisons, masks)."""
    from scipy.optimize._numdiff import approx_derivative
    prob, obj = wo.prob, wo.obj

    def stacked(v):
        return np.concatenate((prob.eval_equality(v, obj), prob.eval_inequality(v, obj),
                               np.atleast_1d(prob.eval_cost(v, obj))))
    try:
        cols = []
        with np.errstate(all="ignore"):
            for j in range(x.size):
                z = x.astype(complex)
                z[j] += 1e-30j
                cols.append(np.imag(stacked(z)) / 1e-30)
        J = np.stack(cols, axis=1)
        if np.isfinite(J).all():
            return J, "cs"
    except TypeError:                                   # a comparison / mask on complex values
        pass
    return np.atleast_2d(approx_derivative(stacked, x, method="3-point", bounds=(lb, ub))), "3-point"


def check_exact(J, wo, x, lb, ub, J_fd=None):
    J_ref, how = oracle_jacobian(wo, x, lb, ub)
    rowmax = np.abs(J_ref).max(axis=1, keepdims=True)
    tol = EXACT_RTOL_CS if how == "cs" else EXACT_RTOL_3PT
    err = np.abs(J - J_ref) / np.maximum(rowmax, 1e-300)
    assert err.max() <= tol, "exact Jacobian vs %s oracle: row-scaled error %g" % (how, err.max())
    if J_fd is not None:
        fd_err = np.abs(J_fd - J) / np.maximum(np.abs(J).max(axis=1, keepdims=True), 1e-300)
        assert fd_err.max() <= FD_VS_EXACT_RTOL, "FD vs exact: row-scaled difference %g" % fd_err.max()
    return how


def interior_point(wl, seed):
    """A seeded instance strictly inside its bounds (at an active bound FD flips its step; the exact
    derivative does not care, but the central-difference oracle does)."""
    lb, ub = wl.prob.bounds_arrays()
    p = workloads.make_batch(wl, 1, first=seed)[0]
    span = np.where(np.isfinite(ub - lb), ub - lb, 1.0)
    p = np.where(np.isfinite(lb), np.maximum(p, lb + 1e-3 * span), p)
    p = np.where(np.isfinite(ub), np.minimum(p, ub - 1e-3 * span), p)
    # ... and away from the kinks of |v|-type terms: the shipped guesses hold exact zeros (velocities at rest),
    # where d|v|/dv is one-sided for FD, +-1 for the complex step and v_r/|v| for the chain rule
    rng = np.random.default_rng(1000 + seed)
    small = np.abs(p) < 1e-3
    p = np.where(small, 0.01 * (1.0 + rng.random(p.size)), p)
    p = np.clip(p, np.where(np.isfinite(lb), lb + 1e-3 * span, p), np.where(np.isfinite(ub), ub - 1e-3 * span, p))
    return p, lb, ub


@pytest.mark.parametrize("name", CFGS)
def test_emulated_exact_jacobian_vs_oracle_derivatives(api, name):
    from tests.emu.emu import EmuProblem
    wl, wo = workloads.build(name, api), workloads.build(name, og_numpy)
    ir = tape.build_ir(wl.prob, wl.obj)
    p, lb, ub = interior_point(wl, 3)
    emu = EmuProblem(ir, lb, ub)
    c, J = emu.eval_exact(p)
    c_fd, J_fd = emu.eval_fd(p)
    assert np.array_equal(c, c_fd)                      # the same constraint vector, bit for bit
    check_exact(J[0].T, wo, p, lb, ub, J_fd[0].T)
    # structure: wherever the FD Jacobian is structurally zero the exact one is exactly zero
    lin_fd = (J_fd[0] != 0)
    assert not (J[0] != 0)[~lin_fd & (np.abs(J[0]) > 1e-9 * np.abs(J[0]).max())].any()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CFGS)
def test_device_exact_jacobian(api, name):
    import torch
    wl, wo = workloads.build(name, api), workloads.build(name, og_numpy)
    eng = wl.prob.compile(wl.obj)
    B = 23
    P = np.stack([interior_point(wl, s)[0] for s in range(B)])
    lb, ub = wl.prob.bounds_arrays()
    lin = eng.jac_pattern().astype(np.int64)
    c_fd, J_fd = eng.eval_fd(P)
    res = {}
    for jit in (1, 0):                                  # NVRTC-generated dual programs / the dual tape interpreter
        eng.set_option(2, jit)
        c, vals = eng.eval_exact(P)
        assert torch.equal(c, c_fd)
        J = eng.densify(vals)
        assert torch.equal(J.reshape(B, -1)[:, torch.from_numpy(lin).to(J.device)], vals)
        res[jit] = J.cpu().numpy()
    err = np.abs(res[1] - res[0]) / np.maximum(np.abs(res[0]).max(axis=1, keepdims=True), 1e-300)
    assert err.max() <= 1e-13, err.max()                # same formulas; the compiler may contract differently
    Jf = J_fd.cpu().numpy()
    for b in (0, B - 1):
        check_exact(res[1][b].T, wo, P[b], lb, ub, Jf[b].T)
    eng.set_option(0, 1)                                # generic column code
    _, v2 = eng.eval_exact(P)
    assert torch.equal(eng.densify(v2), torch.from_numpy(res[0]).to(v2.device))


@pytest.mark.gpu
def test_exact_mode_through_the_host_session_and_the_sqp_driver(api):
    """jacobian="exact" end to end: the host session's packed values equal the device ones; SLSQP fed with the
    exact Jacobians converges on the brachistochrone to the same optimum as with forward differences."""
    import torch
    wl = workloads.build("cfg1_brachistochrone20", api)
    prob, obj = wl.prob, wl.obj
    P = workloads.make_batch(wl, 5)
    c_d, v_d = prob.evaluate_batch(P, obj, jacobian="exact")
    c_h, v_h = prob.evaluate_batch(P, obj, jacobian="exact", host=True)
    assert (c_h == c_d.cpu().numpy()).all() and (v_h == v_d.cpu().numpy()).all()
    c_s, v_s = prob.evaluate_batch(P, obj, jacobian="sparse")
    _, J = prob.evaluate_batch(P, obj)
    lin = torch.from_numpy(prob._engine.jac_pattern().astype(np.int64)).to(J.device)
    assert torch.equal(v_s, J.reshape(5, -1)[:, lin]) and torch.equal(c_s, c_d)
    fd = prob.solve_batch(P, obj, ftol=1e-8, maxiter=60, max_outer=6)
    ex = prob.solve_batch(P, obj, ftol=1e-8, maxiter=60, max_outer=6, jacobian="exact")
    assert (ex["status"] == 0).all() and (fd["status"] == 0).all()
    assert np.abs(ex["fun"] - np.sqrt(np.pi)).max() < 1e-4 and np.abs(ex["fun"] - fd["fun"]).max() < 1e-5


@pytest.mark.parametrize("name", ["cfg2_goddard50", "edge_all_ops", "edge_table_lookup"])
def test_exact_and_packed_kernel_variants_compile_without_a_gpu(api, name):
    from opengoddard_b200 import capi
    wl = workloads.build(name, api)
    ir = tape.build_ir(wl.prob, wl.obj)
    for variant in (1, 2):
        size, src = capi.jit_check(ir, variant)
        assert size > 10000 and "ogb_jit_node_dual" in src
