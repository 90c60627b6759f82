"""Seeded random expression trees as user dynamics / constraints / costs: the traced + tape-compiled
path (executed by the host-compiled device arithmetic, tests/emu) must reproduce what numpy computes
eagerly from the very same Python callbacks through the oracle.  Exercises hash-consing, register
allocation with last-use reuse, constants, selects and mixed scalar / node-vector operands far beyond
the hand-written workloads."""
import numpy as np
import pytest

from opengoddard_b200 import tape
from oracle import og_numpy
from tests.emu.emu import EmuProblem
from tests.helpers import assert_c_close, assert_J_close

UNARY = [np.sin, np.cos, np.tanh, np.arctan, lambda v: np.exp(-np.square(v)), lambda v: np.sqrt(1.0 + v * v),
         lambda v: np.log(2.0 + np.cos(v)), lambda v: -v, np.sinh, lambda v: 1.0 / (2.0 + np.sin(v)),
         lambda v: v ** 3, lambda v: np.abs(v) + 0.5, lambda v: np.cosh(0.3 * v), lambda v: np.clip(v, -0.7, 0.9)]
BINARY = [lambda a, b: a + b, lambda a, b: a - b, lambda a, b: a * b, lambda a, b: a / (1.5 + b * b),
          np.minimum, np.maximum, lambda a, b: np.arctan2(a, 1.0 + b * b), lambda a, b: np.hypot(a, b),
          lambda a, b: np.where(a > b, a, 0.5 * b), lambda a, b: np.where(np.logical_and(a > 0.0, b < 0.3), a * b, a - b)]


def random_expr(rng, leaves, depth):
    """A random function of the leaves (callables returning arrays / traced values)."""
    if depth == 0 or rng.random() < 0.15:
        k = int(rng.integers(len(leaves) + 1))
        if k == len(leaves):
            cst = float(np.round(rng.normal(), 3))
            return lambda env: cst
        return lambda env: env[k]
    if rng.random() < 0.45:
        f = UNARY[int(rng.integers(len(UNARY)))]
        a = random_expr(rng, leaves, depth - 1)
        return lambda env: f(a(env))
    f = BINARY[int(rng.integers(len(BINARY)))]
    a, b = random_expr(rng, leaves, depth - 1), random_expr(rng, leaves, depth - 1)
    return lambda env: f(a(env), b(env))


def build(api, seed):
    rng = np.random.default_rng(seed)
    nsec = int(rng.integers(1, 3))
    nodes = [int(rng.integers(4, 9)) for _ in range(nsec)]
    ns, ncn = int(rng.integers(2, 4)), int(rng.integers(1, 3))
    prob = api.Problem(list(np.linspace(0.0, 2.0, nsec + 1)), nodes, [ns] * nsec, [ncn] * nsec, 3)
    if rng.random() < 0.5:
        prob.set_unit_states_all_section(0, 2.0)
        prob.set_unit_time(1.5)
    names = list(range(ns + ncn))
    dyn_exprs = [random_expr(rng, names, 4) for _ in range(ns)]
    ineq_expr, run_expr = random_expr(rng, names, 3), random_expr(rng, names, 3)
    eq_expr, cost_expr = random_expr(rng, names, 3), random_expr(rng, names, 3)

    def env_of(prob, section):
        return [prob.states(a, section) for a in range(ns)] + [prob.controls(u, section) for u in range(ncn)]

    def dynamics(prob, obj, section):
        env = env_of(prob, section)
        d = api.Dynamics(prob, section)
        for a in range(ns):
            d[a] = dyn_exprs[a](env) + 0.0 * env[0]                # (a bare constant still becomes a vector)
        return d()

    def equality(prob, obj):
        r = api.Condition()
        env0 = [v[0] for v in env_of(prob, 0)]
        envf = [v[-1] for v in env_of(prob, nsec - 1)]
        r.equal(env0[0], 0.2)
        r.equal(eq_expr(envf) + envf[1], 0.1)
        for s in range(nsec - 1):
            for a in range(ns):
                r.equal(prob.states(a, s)[-1], prob.states(a, s + 1)[0])
        return r()

    def inequality(prob, obj):
        r = api.Condition()
        for s in range(nsec):
            r.upper_bound(ineq_expr(env_of(prob, s)) + 0.0 * prob.states(0, s), 50.0)
        r.lower_bound(prob.time_final(-1), 0.1)
        return r()

    def cost(prob, obj):
        return cost_expr([v[-1] for v in env_of(prob, nsec - 1)]) + prob.time_final(-1)

    def running(prob, obj):
        env = [prob.states_all_section(a) for a in range(ns)] + [prob.controls_all_section(u) for u in range(ncn)]
        return run_expr(env) + 0.0 * env[0]

    for s in range(nsec):
        for a in range(ns):
            prob.set_states(a, s, rng.uniform(-0.8, 0.8, nodes[s]))
        for u in range(ncn):
            prob.set_controls(u, s, rng.uniform(-0.8, 0.8, nodes[s]))
    prob.set_states_bounds_all_section(0, -2.0, 2.0)
    prob.dynamics = [dynamics] * nsec
    prob.knot_states_smooth = [False] * (nsec - 1)
    prob.cost = cost
    prob.running_cost = running
    prob.equality = equality
    prob.inequality = inequality
    return prob


@pytest.mark.parametrize("seed", range(24))
def test_random_callbacks_trace_and_match_numpy(api, seed):
    pa, po = build(api, seed), build(og_numpy, seed)
    assert np.array_equal(np.asarray(pa.p), np.asarray(po.p))
    lb, ub = og_numpy.bounds_arrays(po)
    emu = EmuProblem(tape.build_ir(pa, None), lb, ub)
    rng = np.random.default_rng(1000 + seed)
    P = np.asarray(po.p)[None] * (1.0 + 0.05 * rng.standard_normal((3, len(po.p))))
    c, J = emu.eval_fd(P)
    for b in range(3):
        c_ref, J_ref = og_numpy.eval_fd(po, None, P[b], lb, ub)
        if not (np.isfinite(c_ref).all() and np.isfinite(J_ref).all()):
            continue
        assert_c_close(c[b], c_ref, J_ref, np.clip(P[b], lb, ub))
        # value tolerance as everywhere (1e-6 of the row maximum); the zero-pattern check allows dust up
        # to that same size here: the rows of these random problems are small numbers, and sqrt(a^2+b^2)
        # vs libm's hypot differ by an ulp, which a forward difference turns into ~1e-7 of such a row
        assert_J_close(J[b].T, J_ref, dust=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 3, 7, 11, 21, 23])
def test_random_callbacks_on_the_device(api, seed):
    """The same random problems through the real kernels: NVRTC-compiled tapes == tape interpreter bit
    for bit, and both match numpy."""
    import torch
    pa, po = build(api, seed), build(og_numpy, seed)
    lb, ub = og_numpy.bounds_arrays(po)
    eng = pa.compile(None)
    assert eng.info.jit == 1
    rng = np.random.default_rng(1000 + seed)
    P = np.asarray(po.p)[None] * (1.0 + 0.05 * rng.standard_normal((5, len(po.p))))
    c1, J1 = eng.eval_fd(P)
    eng.set_option(2, 0)
    c0, J0 = eng.eval_fd(P)
    assert torch.equal(c0, c1) and torch.equal(J0, J1)
    c, J = c1.cpu().numpy(), J1.cpu().numpy()
    for b in range(5):
        c_ref, J_ref = og_numpy.eval_fd(po, None, P[b], lb, ub)
        if not (np.isfinite(c_ref).all() and np.isfinite(J_ref).all()):
            continue
        assert_c_close(c[b], c_ref, J_ref, np.clip(P[b], lb, ub))
        assert_J_close(J[b].T, J_ref, dust=1e-6)


def _small_problem(mod, f, g):
    prob = mod.Problem([0.0, 1.0], [6], [2], [1], 3)

    def dyn(prob, obj, section):
        x, y, u = prob.states(0, section), prob.states(1, section), prob.controls(0, section)
        d = mod.Dynamics(prob, section)
        d[0] = f(x, y, u)
        d[1] = u
        return d()

    def eq(p, o):
        r = mod.Condition()
        r.equal(p.states(0, 0)[0], 0.3)
        return r()

    prob.dynamics = [dyn]
    prob.knot_states_smooth = []
    prob.cost = lambda p, o: p.time_final(-1) + g(p.states(0, 0), p.states(1, 0))
    prob.equality = eq
    prob.inequality = lambda p, o: mod.Condition()()
    prob.set_states(0, 0, np.linspace(0.3, 0.9, 6))
    prob.set_states(1, 0, np.linspace(1.2, 2.1, 6))
    prob.set_controls(0, 0, np.linspace(-0.4, 0.5, 6))
    return prob


IDIOMS = {
    "like": (lambda x, y, u: np.zeros_like(x) + y * np.ones_like(x) + np.full_like(x, 2.5) * u, lambda x, y: 0.0),
    "expm1_log1p": (lambda x, y, u: np.expm1(x) + np.log1p(y * y), lambda x, y: 0.0),
    "isnan_heaviside": (lambda x, y, u: np.where(np.isnan(x), 0.0, x) + np.heaviside(u, 0.5) * y, lambda x, y: 0.0),
    "mod_fmod": (lambda x, y, u: np.mod(3.7 * x + u, 0.8) + np.fmod(5.3 * u - 0.1, 0.7) + np.remainder(y, 0.9),
                 lambda x, y: 0.0),
    "reductions_in_cost": (lambda x, y, u: x * y,
                           lambda x, y: np.mean(x) + 0.1 * np.max(y) - 0.2 * np.min(x) + 0.01 * np.dot(x, y)
                           + 0.05 * np.linalg.norm(y) + 0.001 * np.prod(x)),
    "python_builtins": (lambda x, y, u: abs(x) + pow(y, 2) + x.copy() * len(x) + x.shape[0] * u, lambda x, y: 0.0),
}


@pytest.mark.parametrize("name", sorted(IDIOMS))
def test_numpy_idioms_trace_and_match(api, name):
    """Common numpy idioms in user callbacks (constructors, expm1 / log1p, isnan, heaviside, mod / fmod,
    reductions over the nodes inside a cost) lower to tape operations that reproduce numpy."""
    f, g = IDIOMS[name]
    pa, po = _small_problem(api, f, g), _small_problem(og_numpy, f, g)
    lb, ub = og_numpy.bounds_arrays(po)
    emu = EmuProblem(tape.build_ir(pa, None), lb, ub)
    P = np.asarray(po.p)[None] * (1.0 + 0.03 * np.random.default_rng(1).standard_normal((2, len(po.p))))
    c, J = emu.eval_fd(P)
    for b in range(2):
        c_ref, J_ref = og_numpy.eval_fd(po, None, P[b], lb, ub)
        assert_c_close(c[b], c_ref, J_ref, P[b])
        assert_J_close(J[b].T, J_ref, dust=1e-6)


def test_node_reductions_inside_dynamics_become_dense_columns(api):
    """A reduction over the nodes inside `dynamics` (here sum(x)) is not node-local.  Round 1 refused it; since
    round 2 every element the reduction reads is a global input of the node program and the Jacobian columns of
    those variables are dense in the phase -- checked against the oracle (FD) and against complex-step
    differentiation (exact mode)."""
    from tests.test_exact import check_exact
    f, g = (lambda x, y, u: y + np.sum(x) * 0.1 + 0.2 * np.max(u)), (lambda x, y: 0.0)
    pa, po = _small_problem(api, f, g), _small_problem(og_numpy, f, g)
    ir = tape.build_ir(pa, None)
    assert len(ir.node_tapes[0].globals) == 12               # the 6 x's and the 6 u's
    lb, ub = og_numpy.bounds_arrays(po)
    emu = EmuProblem(ir, lb, ub)
    P = np.asarray(po.p)[None] * (1.0 + 0.03 * np.random.default_rng(2).standard_normal((2, len(po.p))))
    c, J = emu.eval_fd(P)
    for b in range(2):
        c_ref, J_ref = og_numpy.eval_fd(po, None, P[b], lb, ub)
        assert_c_close(c[b], c_ref, J_ref, P[b])
        assert_J_close(J[b].T, J_ref, dust=1e-6)
    class W:                                                 # (check_exact wants a workload-like holder)
        prob, obj = po, None
    ce, Je = emu.eval_exact(P[0])
    assert np.array_equal(ce[0], c[0])
    x = np.clip(P[0], lb, ub)
    x[12:18] += np.linspace(0.0, 0.05, 6)                    # (make max(u) unique: the arg-max is then differentiable)
    ce, Je = emu.eval_exact(x)
    cf, Jf = emu.eval_fd(x)
    check_exact(Je[0].T, W, x, lb, ub, Jf[0].T)
