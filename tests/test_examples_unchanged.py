"""The reference's shipped example scripts, UNCHANGED, against the drop-in facade.

Container-only (needs /root/reference/examples; skipped on the GPU box).  Every script
is exec()'d with `from OpenGoddard.optimize import ...` resolving to the facade and
`Problem.solve` intercepted at the point where the reference would call SciPy; the
callbacks the script defined are then traced and lowered, and the device arithmetic (CPU
emulation, tests/emu) must reproduce the c vector -- and for the BASELINE examples the
FD Jacobian -- that the reference itself produced for that script (tests/golden/example_XX).  Example 11
interpolates data tables (scipy interp1d) inside its dynamics: those become device lookup tables.
Example 01 is additionally solved end to end on the explicit host backend."""
import os

import numpy as np
import pytest

from oracle import ref_loader
from oracle.example_trace import run_script
from tests.helpers import assert_c_close, assert_J_close, golden

EXDIR = os.path.join(ref_loader.REFERENCE_ROOT, "examples")
pytestmark = pytest.mark.skipif(not os.path.isdir(EXDIR), reason="reference examples only in the build container")


@pytest.mark.parametrize("tag", ["01", "02", "03", "04", "05", "06", "07", "08", "09", "10", "11"])
def test_example_traces_and_matches_reference(tag):
    from opengoddard_b200 import tape
    from tests.emu.emu import EmuProblem
    e = golden("example_" + tag)
    box, glb, _ = run_script(tag)
    prob, obj = box["prob"], box["obj"]
    assert np.allclose(prob.p, e["x0"], rtol=1e-12, atol=1e-15)   # guess built through the facade's LGL times
    ir = tape.build_ir(prob, obj)
    lb, ub = prob.bounds_arrays()
    assert np.array_equal(lb, e["lb"]) and np.array_equal(ub, e["ub"])
    emu = EmuProblem(ir, lb, ub)
    assert emu.info.meq == e["c_eq"].size and emu.info.mineq == e["c_ineq"].size
    x = np.clip(e["x0"], lb, ub)
    c_ref = np.concatenate((e["c_eq"], e["c_ineq"], [e["cost"]]))
    if "J_eq" in e.files:
        J_ref = np.vstack((e["J_eq"], e["J_ineq"], e["g_cost"][None]))
        c, J = emu.eval_fd(x)
        assert_c_close(c[0], c_ref, J_ref, x)
        assert_J_close(J[0].T, J_ref)
    else:
        c = emu.eval(x)
        assert np.abs(c[0] - c_ref).max() <= 1e-9 * max(1.0, np.abs(c_ref).max())


def test_example_01_runs_unchanged_end_to_end_on_host_backend():
    box, glb, text = run_script("01", intercept=False, env_backend="host")
    assert "Optimization terminated successfully" in text
    prob = glb["prob"]
    assert abs(prob.time_final(-1) - 1.7724608832526498) < 1e-5
