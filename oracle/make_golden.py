"""Generate the golden vectors under tests/golden/ from the REAL reference.

TEST INFRASTRUCTURE ONLY; runs only where /root/reference exists (the build
container):   python -m oracle.make_golden

Three families (all float64, produced by the unmodified reference module loaded
through oracle/ref_loader.py plus SciPy 1.18.1's own `approx_derivative`, i.e.
exactly what `scipy/optimize/_slsqp_py.py:353-367` runs):

  lgl.npz              tau / w / D for a set of node counts
                       (OpenGoddard/optimize.py:183-213)
  example_XX.npz       the shipped example scripts executed UNCHANGED with
                       `optimize.minimize` replaced by a recorder: decision vector
                       at the shipped guess, bounds, c_eq, c_ineq, cost (and the FD
                       Jacobians for the examples BASELINE.json names)
  workload_<cfg>.npz   `opengoddard_b200.workloads` builders run against the
                       reference module on seeded jittered instances: p, c_eq,
                       c_ineq, cost, J_eq, J_ineq, grad(cost)
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

ABS_STEP = float(np.sqrt(np.finfo(np.float64).eps))   # scipy/optimize/_slsqp_py.py:34

LGL_NODES = (3, 4, 5, 8, 20, 25, 30, 40, 50, 64, 100, 128)
EXAMPLES_FD = ("01", "02", "03", "04", "05", "06", "07", "08", "09", "10", "11")
EXAMPLES_C = ()
# solve-level goldens: the reference's own SLSQP run (its closures, its FD Jacobians, SciPy's minimize
# exactly as optimize.py:740-749 calls it) stopped after k major iterations, from the shipped guess
EXAMPLES_SOLVE = {"01": (1, 2, 3, 6), "04": (1, 2, 3), "05": (1, 2, 3)}
WORKLOAD_INSTANCES = {
    "cfg1_brachistochrone20": 3,
    "cfg2_goddard50": 4,
    "cfg3_goddard_knot30x2": 3,
    "cfg4_polar3x40": 2,
    "cfg5_lowthrust128": 2,
    "ex05_goddard_knot25x2": 2,
    "ex09_polar_tsto20x2": 2,
    "ex10_lowthrust100": 1,
    "edge_nonautonomous": 2,
    "edge_picked_dynamics": 2,
}


def new_bounds(bounds):
    lb = np.array([-np.inf if b[0] is None else float(b[0]) for b in bounds])
    ub = np.array([np.inf if b[1] is None else float(b[1]) for b in bounds])
    return lb, ub


def fd_jac(fun, x, args, lb, ub):
    from scipy.optimize._numdiff import approx_derivative
    return np.atleast_2d(approx_derivative(fun, x, method="2-point", abs_step=ABS_STEP,
                                           args=args, bounds=(lb, ub)))


def golden_lgl(mod):
    out = {}
    prob = mod.Problem([0.0, 1.0], [3], [1], [1])
    for n in LGL_NODES:
        out["tau_%d" % n] = prob._nodes_LGL(n)
        out["w_%d" % n] = prob._weight_LGL(n)
        out["D_%d" % n] = prob._differentiation_matrix_LGL(n)
    np.savez_compressed(os.path.join(GOLD, "lgl.npz"), **out)
    print("lgl.npz", len(out))


def run_example(mod, tag):
    """exec() one shipped example script unchanged; `OpenGoddard.optimize` resolves
    to the reference module, `minimize` to a recorder."""
    exdir = os.path.join(ref_loader.REFERENCE_ROOT, "examples")
    script = [f for f in sorted(os.listdir(exdir)) if f.startswith(tag) and f.endswith(".py")][0]
    pkg = types.ModuleType("OpenGoddard")
    pkg.__path__ = []
    pkg.optimize = mod
    saved = {k: sys.modules.get(k) for k in ("OpenGoddard", "OpenGoddard.optimize")}
    sys.modules["OpenGoddard"] = pkg
    sys.modules["OpenGoddard.optimize"] = mod
    box = {}

    def recorder(fun, x0, args=(), bounds=None, constraints=(), jac=None, method=None,
                 options=None, **kw):
        box.setdefault("cap", ref_loader.Captured(fun, np.array(x0, dtype=float), args,
                                                  bounds, constraints, jac, options))
        return types.SimpleNamespace(message="captured", status=0, x=x0)

    real_opt = mod.optimize
    mod.optimize = types.SimpleNamespace(minimize=recorder, root=real_opt.root)
    cwd = os.getcwd()
    os.chdir(exdir)
    try:
        src = open(script).read()
        glb = {"__name__": "__main__", "__file__": script}
        with contextlib.redirect_stdout(io.StringIO()):
            try:
                exec(compile(src, script, "exec"), glb)
            except Exception as e:  # post-processing may fail (removed scipy APIs) -- after solve
                if "cap" not in box:
                    raise
                box["post_error"] = repr(e)
    finally:
        os.chdir(cwd)
        mod.optimize = real_opt
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return script, box["cap"], glb


def golden_example(mod, tag, with_fd):
    script, cap, glb = run_example(mod, tag)
    x0 = cap.x0
    lb, ub = new_bounds(cap.bounds)
    x = np.clip(x0, lb, ub)                         # scipy/optimize/_slsqp_py.py:322
    out = dict(script=np.array(script), x0=x0, lb=lb, ub=ub,
               c_eq=np.atleast_1d(cap.eq(x)).astype(float),
               c_ineq=np.atleast_1d(cap.ineq(x)).astype(float),
               cost=np.float64(cap.cost(x)), has_cost_derivative=np.array(cap.jac is not None))
    prob = glb["prob"]
    out["nodes"] = np.array(prob.nodes)
    out["nstates"] = np.array(prob.number_of_states)
    out["ncontrols"] = np.array(prob.number_of_controls)
    if cap.jac is not None:
        out["cost_derivative"] = np.asarray(cap.jac(x, *cap.args), dtype=float)
    if with_fd:
        out["J_eq"] = fd_jac(cap.constraints[0]["fun"], x, cap.constraints[0]["args"], lb, ub)
        out["J_ineq"] = fd_jac(cap.constraints[1]["fun"], x, cap.constraints[1]["args"], lb, ub)
        out["g_cost"] = fd_jac(cap.fun, x, cap.args, lb, ub).ravel()
    if tag in EXAMPLES_SOLVE:
        from scipy import optimize
        ftol = (cap.options or {}).get("ftol", 1e-6)
        out["solve_ftol"] = np.float64(ftol)

        def run(k, cons=cap.constraints, start=x0):
            with contextlib.redirect_stdout(io.StringIO()):
                return optimize.minimize(cap.fun, np.array(start, dtype=float), args=cap.args, bounds=cap.bounds,
                                         constraints=cons, jac=cap.jac, method="SLSQP",
                                         options={"disp": False, "maxiter": k, "ftol": ftol})
        for k in EXAMPLES_SOLVE[tag]:
            opt = run(k)
            out["solve_x_%d" % k] = np.array(opt.x, dtype=float)
            out["solve_fun_%d" % k] = np.float64(opt.fun)
            out["solve_status_%d" % k] = np.int64(opt.status)
            out["solve_nit_%d" % k] = np.int64(opt.nit)
        # the reference's c and FD Jacobians at its own last recorded iterate (a point reached by the
        # optimiser: active bounds, non-smooth guesses gone) -- evaluation parity along the trajectory
        k = EXAMPLES_SOLVE[tag][-1]
        xk = np.clip(out["solve_x_%d" % k], lb, ub)
        out["traj_k"] = np.int64(k)
        out["traj_c_eq"] = np.atleast_1d(cap.eq(xk.copy())).astype(float)
        out["traj_c_ineq"] = np.atleast_1d(cap.ineq(xk.copy())).astype(float)
        out["traj_cost"] = np.float64(cap.cost(xk.copy()))
        out["traj_J_eq"] = fd_jac(cap.constraints[0]["fun"], xk.copy(), cap.constraints[0]["args"], lb, ub)
        out["traj_J_ineq"] = fd_jac(cap.constraints[1]["fun"], xk.copy(), cap.constraints[1]["args"], lb, ub)
        out["traj_g_cost"] = fd_jac(cap.fun, xk.copy(), cap.args, lb, ub).ravel()
        # how well-posed is "the iterate after k iterations"?  The reference's own first iterate when its FD
        # Jacobians are perturbed by 1e-12 relative (seeded): SLSQP's LSQ step on these problems is
        # discontinuous in its inputs, so iterate-level parity cannot be asserted (tests/test_examples_gpu.py)
        rng = np.random.default_rng(20261017)
        noisy = tuple(dict(c, jac=(lambda x, *a, c=c: (lambda J: J * (1.0 + 1e-12 * rng.standard_normal(J.shape)))(
            fd_jac(c["fun"], x, a, lb, ub)))) for c in cap.constraints)
        out["solve_x_1_perturbed_1e-12"] = np.array(run(1, cons=noisy).x, dtype=float)
        if tag == "01":                              # converges: the reference's outer loop (optimize.py:738-755)
            x, outer = np.array(x0, dtype=float), 0
            for outer in range(1, 31):
                opt = run(25, start=x)
                x = opt.x
                if opt.status == 0:
                    break
            out["solve_final_x"] = np.array(x, dtype=float)
            out["solve_final_fun"] = np.float64(opt.fun)
            out["solve_final_status"] = np.int64(opt.status)
            out["solve_final_outer"] = np.int64(outer)
    np.savez_compressed(os.path.join(GOLD, "example_%s.npz" % tag), **out)
    print("example", tag, script, "n=%d meq=%d mineq=%d" % (x0.size, out["c_eq"].size,
                                                            out["c_ineq"].size))


def golden_workload(mod, name, ninst):
    from opengoddard_b200 import workloads
    wl = workloads.build(name, mod)
    cap = ref_loader.capture_solve(mod, wl.prob, wl.obj)
    lb, ub = new_bounds(cap.bounds)
    guess = np.array(cap.x0, dtype=float)
    P = np.vstack([np.clip(guess, lb, ub)[None, :], workloads.make_batch(wl, ninst)])
    out = dict(P=P, lb=lb, ub=ub, guess=guess)
    ceq, cin, cost, jeq, jin, gc = [], [], [], [], [], []
    for p in P:
        ceq.append(np.atleast_1d(cap.eq(p.copy())))
        cin.append(np.atleast_1d(cap.ineq(p.copy())))
        cost.append(float(cap.cost(p.copy())))
        jeq.append(fd_jac(cap.constraints[0]["fun"], p.copy(), cap.constraints[0]["args"], lb, ub))
        jin.append(fd_jac(cap.constraints[1]["fun"], p.copy(), cap.constraints[1]["args"], lb, ub))
        gc.append(fd_jac(cap.fun, p.copy(), cap.args, lb, ub).ravel())
    out.update(c_eq=np.array(ceq), c_ineq=np.array(cin), cost=np.array(cost),
               J_eq=np.array(jeq), J_ineq=np.array(jin), g_cost=np.array(gc))
    if cap.jac is not None:
        out["cost_derivative"] = np.asarray(cap.jac(P[0].copy(), *cap.args), dtype=float)
    np.savez_compressed(os.path.join(GOLD, "workload_%s.npz" % name), **out)
    print("workload", name, "P", P.shape, "meq", out["c_eq"].shape[1], "mineq",
          out["c_ineq"].shape[1])


def main():
    os.makedirs(GOLD, exist_ok=True)
    mod = ref_loader.load_reference()
    golden_lgl(mod)
    for tag in EXAMPLES_FD:
        golden_example(mod, tag, True)
    for tag in EXAMPLES_C:
        golden_example(mod, tag, False)
    for name, k in WORKLOAD_INSTANCES.items():
        golden_workload(mod, name, k)


if __name__ == "__main__":
    main()
