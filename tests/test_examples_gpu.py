"""The reference's eleven shipped example scripts on the CUDA kernels.

The scripts themselves live under /root/reference (absent on the GPU box), so what travels is
generated data: the IR our tracer extracted from each UNCHANGED script (tests/golden/example_XX_ir.npz,
oracle/example_trace.py) and what the REFERENCE computed for it (tests/golden/example_XX.npz,
oracle/make_golden.py: c, the SciPy FD Jacobians, and for 01 / 04 / 05 the iterates of the reference's
own SLSQP run).  GPU tests: rebuild the engine from the IR, run both kernel builds and every device
path, compare.  CPU test: the same solve-level comparison through the g++ emulation of the device
arithmetic (tests/emu)."""
import numpy as np
import pytest

from opengoddard_b200 import tape
from tests.helpers import assert_c_close, assert_J_close, golden

TAGS = ["01", "02", "03", "04", "05", "06", "07", "08", "09", "10", "11"]


def _reference(tag):
    e = golden("example_" + tag)
    c_ref = np.concatenate((e["c_eq"], e["c_ineq"], [e["cost"]]))
    J_ref = np.vstack((e["J_eq"], e["J_ineq"], e["g_cost"][None]))
    return e, c_ref, J_ref


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_example_on_the_device_matches_the_reference(tag):
    import torch
    from opengoddard_b200 import engine
    assert torch.cuda.is_available()
    e, c_ref, J_ref = _reference(tag)
    f = golden("example_%s_ir" % tag)
    ir = tape.ir_from_arrays(f)
    assert np.array_equal(f["lb"], e["lb"]) and np.array_equal(f["ub"], e["ub"])
    assert np.allclose(f["x0"], e["x0"], rtol=1e-12, atol=1e-15)
    x = np.clip(e["x0"], e["lb"], e["ub"])
    X = np.stack([x, x, x])
    for jit in (True, False):                          # NVRTC-specialised kernel, tape interpreter
        eng = engine.DeviceProblem(ir, (f["lb"], f["ub"]), jit=jit)
        assert eng.info.jit == int(jit)
        assert eng.meq == e["c_eq"].size and eng.mineq == e["c_ineq"].size
        eng.set_option(9, 0)
        c, J = eng.eval_fd(X)                          # fused sweep kernel
        c, J = c.cpu().numpy(), J.cpu().numpy()
        assert (c[1] == c[0]).all() and (J[2] == J[0]).all()
        assert_c_close(c[0], c_ref, J_ref, x)
        assert_J_close(J[0].T, J_ref)
        assert_c_close(eng.eval(X).cpu().numpy()[0], c_ref, J_ref, x)
        eng.set_option(9, 1)                           # split pipeline: packed sweep + densify
        eng.set_option(10, 2)
        c2, J2 = eng.eval_fd(X)
        assert (c2.cpu().numpy() == c).all() and (J2.cpu().numpy() == J).all()
        ch, Jh = eng.host_evaluator().eval_fd(X)       # host-buffer session
        assert (ch == c).all() and (Jh == J).all()


def _callables(evaluator, e):
    from opengoddard_b200 import engine
    grad = None
    if bool(e["has_cost_derivative"]):
        g = np.array(e["cost_derivative"], dtype=float)     # examples 01 / 05: a constant vector
        grad = lambda x: g
    bounds = [(None if not np.isfinite(a) else float(a), None if not np.isfinite(b) else float(b))
              for a, b in zip(e["lb"], e["ub"])]
    return engine.scipy_callables(evaluator, cost_derivative=grad), bounds


def _minimize(evaluator, e, x0, maxiter):
    from scipy import optimize
    (fun, cons, jac), bounds = _callables(evaluator, e)
    return optimize.minimize(fun, np.array(x0, dtype=float), bounds=bounds, constraints=cons, jac=jac, method="SLSQP",
                             options={"disp": False, "maxiter": maxiter, "ftol": float(e["solve_ftol"])})


def _check_trajectory_point(evaluator, e):
    """c and J of the device arithmetic at the reference's own iterate x_k (k = traj_k) against what the
    reference computed there -- the same tolerances as at the initial guess."""
    k = int(e["traj_k"])
    x = np.clip(e["solve_x_%d" % k], e["lb"], e["ub"])
    c_ref = np.concatenate((e["traj_c_eq"], e["traj_c_ineq"], [e["traj_cost"]]))
    J_ref = np.vstack((e["traj_J_eq"], e["traj_J_ineq"], e["traj_g_cost"][None]))
    c, J = evaluator.eval_fd_host(x)
    assert_c_close(c, c_ref, J_ref, x)
    assert_J_close(J.T, J_ref)
    assert_c_close(evaluator.eval_host(x), c_ref, J_ref, x)


def _check_converges_like_the_reference(evaluator, e):
    """Example 01 through the reference's outer loop (optimize.py:738-755: restart SLSQP from where it
    stopped until exit mode 0): converges, to the reference's converged cost within 10 x ftol."""
    x = np.array(e["x0"], dtype=float)
    for outer in range(1, 31):
        opt = _minimize(evaluator, e, x, 25)
        x = opt.x
        if opt.status == 0:
            break
    ref = float(e["solve_final_fun"])
    assert opt.status == 0 == int(e["solve_final_status"]) and outer <= 2 * int(e["solve_final_outer"])
    assert abs(opt.fun - ref) <= 10.0 * float(e["solve_ftol"]) * abs(ref), (opt.fun, ref)
    assert abs(opt.fun - np.sqrt(np.pi)) <= 1e-4                      # the analytic optimum


# Iterate-level parity ("x after k SLSQP iterations equals the reference's") is NOT asserted, because it is
# not well-posed for these problems: SLSQP's LSQ step (Lawson-Hanson with pseudo-rank decisions) is
# discontinuous in its inputs at the shipped guesses.  The goldens carry the evidence, produced from the
# reference alone: its own first iterate moves by O(1) when its own FD Jacobians are perturbed by 1e-12
# relative.  Two implementations that agree to 1e-8 on J (summation order, libm ulps -- SURVEY.md section 7)
# cannot agree on x_1.  What is well-posed and tested instead: the evaluations at the reference's iterates
# (trajectory points), the SLSQP driver given identical evaluations (tests/test_sqp.py: bit-identical to
# scipy.minimize), and the converged result where the reference converges (example 01).
def test_iterate_level_parity_is_ill_posed_in_the_reference_itself():
    moved = {tag: float(np.abs(golden("example_" + tag)["solve_x_1_perturbed_1e-12"] - golden("example_" + tag)["solve_x_1"]).max())
             for tag in ("01", "04", "05")}
    assert moved["04"] > 1e-3 and moved["05"] > 1e-3, moved          # amplification > 1e9
    assert all(np.isfinite(v) for v in moved.values())


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["01", "04", "05"])
def test_device_evaluations_along_the_reference_trajectory(tag):
    from opengoddard_b200 import engine
    e = golden("example_" + tag)
    f = golden("example_%s_ir" % tag)
    for jit in (False, True):
        _check_trajectory_point(engine.DeviceProblem(tape.ir_from_arrays(f), (f["lb"], f["ub"]), jit=jit), e)


@pytest.mark.gpu
def test_device_backed_solve_converges_like_the_reference():
    from opengoddard_b200 import engine
    e = golden("example_01")
    f = golden("example_01_ir")
    eng = engine.DeviceProblem(tape.ir_from_arrays(f), (f["lb"], f["ub"]), jit=False)
    _check_converges_like_the_reference(eng, e)
    for k in (1, 2, 3, 6):                       # same major-iteration counts / exit modes as the reference's run
        opt = _minimize(eng, e, e["x0"], k)
        assert opt.nit == int(e["solve_nit_%d" % k]) and opt.status == int(e["solve_status_%d" % k])


class _EmuEngine:
    """eval_host / eval_fd_host on the g++ build of the device arithmetic (test infrastructure)."""

    def __init__(self, ir, lb, ub):
        from tests.emu.emu import EmuProblem
        self.emu = EmuProblem(ir, lb, ub)
        self.meq, self.mineq, self.nrows = self.emu.info.meq, self.emu.info.mineq, self.emu.info.nrows

    def eval_host(self, x):
        return self.emu.eval(x)[0]

    def eval_fd_host(self, x):
        c, J = self.emu.eval_fd(x)
        return c[0], J[0]


@pytest.mark.parametrize("tag", ["01", "04", "05"])
def test_emulated_device_arithmetic_along_the_reference_trajectory(tag):
    e = golden("example_" + tag)
    f = golden("example_%s_ir" % tag)
    _check_trajectory_point(_EmuEngine(tape.ir_from_arrays(f), f["lb"], f["ub"]), e)


def test_emulated_device_arithmetic_converges_like_the_reference():
    e = golden("example_01")
    f = golden("example_01_ir")
    _check_converges_like_the_reference(_EmuEngine(tape.ir_from_arrays(f), f["lb"], f["ub"]), e)


@pytest.mark.parametrize("tag", TAGS)
def test_example_ir_fixture_compiles(tag):
    """Every shipped IR fixture lowers to a specialised sweep kernel with NVRTC (no GPU needed)."""
    from opengoddard_b200 import capi
    f = golden("example_%s_ir" % tag)
    size, src = capi.jit_check(tape.ir_from_arrays(f))
    assert size > 10000 and "ogb_jit_node" in src


# ------------------------------------------------------------------ the scripts themselves, on the hardware
def _rundir():
    from oracle import example_trace
    return example_trace.RUNDIR


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_shipped_script_runs_unchanged_on_the_cuda_backend(tag, monkeypatch):
    """exec() the reference's example script as shipped (a copy sits beside the pip-installed reference under
    baseline/_ref, which travels to the GPU box) against the drop-in facade with the default cuda backend:
    SciPy's SLSQP is driven by the CUDA kernels (values and FD Jacobians), the script's own post-processing
    runs on the result.  Outer restarts are capped ($OGB200_MAX_OUTER) to keep the test short; example 01 runs
    to convergence."""
    import os
    from oracle import example_trace
    if not os.path.isdir(_rundir()):
        pytest.skip("no copy of the example scripts on this machine")
    monkeypatch.delenv("OGB200_BACKEND", raising=False)
    monkeypatch.setenv("OGB200_MAX_OUTER", "0" if tag == "01" else "1" if tag in ("10", "11") else "2")
    e = golden("example_" + tag)
    box, glb, text = example_trace.run_script(tag, intercept=False, exdir=_rundir())
    prob = glb["prob"]
    eng = prob._engine
    assert eng is not None and eng.launches >= 3, "the solve did not go through the device engine"
    assert "---- iteration : 1 ----" in text and "Exit mode" in text
    # feasibility improved from the shipped guess, measured with the device's own c at the final iterate
    c = eng.eval_host(np.clip(prob.p, e["lb"], e["ub"]))
    meq = e["c_eq"].size
    viol0 = max(np.abs(e["c_eq"]).max(), max(0.0, -e["c_ineq"].min()) if e["c_ineq"].size else 0.0)
    viol1 = max(np.abs(c[:meq]).max(), max(0.0, -c[meq:-1].min()) if c.size > meq + 1 else 0.0)
    assert np.isfinite(c).all() and np.isfinite(prob.p).all()
    # Examples 03 / 10 / 11 sit on SLSQP's knife edge at their shipped guesses: the REFERENCE's own run of 10 stops
    # in its first iteration ("Inequality constraints incompatible", exit mode 4), and its runs of 03 and 11 flip
    # from 25 iterations to exit mode 4 / 8 after a few when its FD Jacobians are perturbed by 1e-9 of the row
    # maximum -- the size of FD rounding noise (measured in the build container).  For
    # them only "ran on the device, finite" is asserted; everywhere else feasibility must improve.
    if tag in ("03", "10", "11"):
        assert viol1 <= viol0 * (1.0 + 1e-9) + 1e-6, (viol0, viol1)
    else:
        assert eng.launches > 10 and viol1 < max(0.5 * viol0, 1e-6), (viol0, viol1)
    if tag == "01":
        assert "Optimization terminated successfully" in text
        assert abs(prob.time_final(-1) - np.sqrt(np.pi)) < 1e-4
