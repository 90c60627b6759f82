"""CPU-side check of everything except the GPU launch itself: the facade traces the
callbacks, tape.py lowers them, the C++ table builder lays them out and the SAME
__host__ __device__ arithmetic the kernels run (csrc/ogb_core.h, compiled by g++ into
tests/emu/libogb_emu.so -- test infrastructure, never loaded by the product) reproduces
the reference's c and FD Jacobian within the tolerances of tests/helpers.py."""
import numpy as np
import pytest

from opengoddard_b200 import tape, workloads
from tests.emu.emu import EmuProblem, binding
from tests.helpers import (assert_c_close, assert_J_close, assert_lgl_close, golden, stacked_reference)

CFGS = [k for k in workloads.CONFIGS if not k.startswith("edge_")]
EDGE = [k for k in workloads.CONFIGS if k.startswith("edge_")]


@pytest.mark.parametrize("N", [3, 4, 5, 8, 20, 25, 30, 40, 50, 64, 100, 128])
def test_lgl_core_vs_reference(N):
    import ctypes as C
    g = golden("lgl")
    tau, w, D = np.empty(N), np.empty(N), np.empty((N, N))
    ptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    binding().lib.emu_lgl_build(N, ptr(tau), ptr(w), ptr(D))
    assert_lgl_close(tau, g["tau_%d" % N])
    assert_lgl_close(w, g["w_%d" % N])
    assert_lgl_close(D, g["D_%d" % N])
    inner = np.arange(1, N - 1)
    assert (D[inner, inner] == 0.0).all()


@pytest.mark.parametrize("name", CFGS)
def test_traced_tables_reproduce_reference(api, name):
    g = golden("workload_" + name)
    wl = workloads.build(name, api)
    ir = tape.build_ir(wl.prob, wl.obj)
    emu = EmuProblem(ir, *wl.prob.bounds_arrays())
    c_ref, J_ref = stacked_reference(g)
    assert emu.info.nvars == g["P"].shape[1]
    assert emu.info.meq == g["c_eq"].shape[1] and emu.info.mineq == g["c_ineq"].shape[1]
    c, J = emu.eval_fd(g["P"])
    assert_c_close(c, c_ref, J_ref, g["P"])
    assert_J_close(J.transpose(0, 2, 1), J_ref)
    assert_c_close(emu.eval(g["P"]), c_ref, J_ref, g["P"])


def test_bounds_drive_fd_steps(api):
    """Variables sitting on / near their bounds get the flipped or shortened steps SciPy
    would use; compare with the oracle on the polar problem (162 bounded variables)."""
    from oracle import og_numpy
    name = "ex09_polar_tsto20x2"
    wl = workloads.build(name, api)
    wo = workloads.build(name, og_numpy)
    lb, ub = wl.prob.bounds_arrays()
    P = workloads.make_batch(wl, 3, first=40)
    fin = np.isfinite(ub)
    P[0, fin] = ub[fin]                              # on the upper bound -> backward step
    finl = np.isfinite(lb)
    P[1, finl] = lb[finl] + 1e-9                     # closer than h to the lower bound
    P[2] = P[2] + 10.0                               # violates bounds -> clipped first
    emu = EmuProblem(tape.build_ir(wl.prob, wl.obj), lb, ub)
    c, J = emu.eval_fd(P)
    for b in range(3):
        c_ref, J_ref = og_numpy.eval_fd(wo.prob, wo.obj, P[b], lb, ub)
        assert_c_close(c[b], c_ref, J_ref, np.clip(P[b], lb, ub))
        assert_J_close(J[b].T, J_ref)


@pytest.mark.parametrize("name", EDGE)
def test_edge_case_problems_vs_oracle(api, name):
    """Empty inequality, 3..130 nodes per phase (130 > register-cached limit), state-count change
    across a knot, smooth knot rows, running cost, selects, atan2/tanh/cosh/abs, non-local rows."""
    from oracle import og_numpy
    wl = workloads.build(name, api)
    wo = workloads.build(name, og_numpy)
    lb, ub = wl.prob.bounds_arrays()
    emu = EmuProblem(tape.build_ir(wl.prob, wl.obj), lb, ub)
    P = workloads.make_batch(wl, 2, first=3)
    c, J = emu.eval_fd(P)
    for b in range(2):
        c_ref, J_ref = og_numpy.eval_fd(wo.prob, wo.obj, P[b], lb, ub)
        assert emu.info.mineq == wo.prob.eval_inequality(P[b].copy(), wo.obj).size
        assert_c_close(c[b], c_ref, J_ref, np.clip(P[b], lb, ub))
        assert_J_close(J[b].T, J_ref)
