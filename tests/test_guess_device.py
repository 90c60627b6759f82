"""Device-side guess generation, multi-start jitter and trajectory unpacking.

CPU: the __host__ __device__ formulas (g++ build, tests/emu) against the REFERENCE-restating numpy Guess
(oracle/og_numpy.py) and against the numpy Philox oracle, which is pinned by the Random123 known-answer
vectors.  GPU: the kernels through the C ABI against the same oracles and against the facade's eager
setters / time_update / *_all_section."""
import ctypes as C

import numpy as np
import pytest

from opengoddard_b200 import workloads
from oracle import og_numpy, og_rng

GUESS_RTOL_LINEAR = 0.0        # Guess.linear: the same three operations as numpy.interp -> bit-identical
GUESS_RTOL_CUBIC = 1e-11       # Guess.cubic: |y - y_ref| <= 1e-11 * max|y_ref|; the reference inverts a monomial 4x4
#                                system (cond ~1e3-1e6 -> its own error is cond * eps), the device uses the Hermite basis
JITTER_RTOL = 1e-12            # log / sqrt / cos differ by ulps between CUDA libm and glibc

# Random123 known-answer vectors for Philox4x32-10 (counter, key) -> output
KAT = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
       ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
       ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


def test_philox_oracle_and_device_function_match_the_known_answers():
    from tests.emu.emu import binding
    lib = binding().lib
    for ctr, key, expect in KAT:
        assert tuple(int(v) for v in og_rng.philox4x32(*ctr, *key)) == expect
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        lib.emu_philox4x32(c, k, o)
        assert tuple(o) == expect
    rng = np.random.default_rng(1)
    for a, b in rng.integers(0, 2 ** 32, size=(50, 2), dtype=np.uint64):
        assert lib.emu_u53(int(a), int(b)) == float(og_rng.u53(np.uint32(a), np.uint32(b)))


def test_guess_formulas_match_the_reference_guess():
    from tests.emu.emu import binding
    lib = binding().lib
    rng = np.random.default_rng(7)
    for N, (t0, tf) in ((20, (0.0, 2.0)), (50, (0.0, 0.3)), (128, (3.0, 250.0)), (5, (-1.0, 1.0))):
        tau = og_numpy.lgl_nodes(N)
        time = (tf - t0) / 2.0 * tau + (tf + t0) / 2.0
        for _ in range(5):
            q = rng.normal(size=4) * 10.0
            qa = (C.c_double * 4)(*q)
            lin = np.array([lib.emu_guess_value(2, float(t), float(time[0]), float(time[-1]), qa) for t in time])
            assert np.array_equal(lin, og_numpy.Guess.linear(time, q[0], q[1]))
            cub = np.array([lib.emu_guess_value(3, float(t), float(time[0]), float(time[-1]), qa) for t in time])
            ref = og_numpy.Guess.cubic(time, q[0], q[1], q[2], q[3])
            assert np.abs(cub - ref).max() <= GUESS_RTOL_CUBIC * np.abs(ref).max()
            con = np.array([lib.emu_guess_value(1, float(t), 0.0, 1.0, qa) for t in time])
            assert np.array_equal(con, og_numpy.Guess.constant(time, q[0]))
            assert lib.emu_guess_value(0, 0.5, 0.0, 1.0, qa) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg2_goddard50", "cfg3_goddard_knot30x2", "ex09_polar_tsto20x2"])
def test_device_guess_batch_matches_the_reference_setters(api, name):
    """Per instance: Guess.linear / cubic / constant on prob.time_all_section (or one phase's time) stored with
    set_*_all_section / set_states / set_controls / set_time_final, exactly as the shipped examples build p."""
    wl, wo = workloads.build(name, api), workloads.build(name, og_numpy)
    prob, ref = wl.prob, wo.prob
    eng = prob.compile(wl.obj)
    nsec, ns, nc = prob.number_of_section, prob.number_of_states[0], prob.number_of_controls[0]
    B = 6
    rng = np.random.default_rng(5)
    # (distinct variables: the specs are written concurrently; state ns - 1 >= 2 in these problems)
    specs = [("linear", ("state", 0), None), ("cubic", ("state", 1), None), ("constant", ("control", 0), None),
             ("linear", ("state", ns - 1), nsec - 1)] + ([("zeros", ("control", nc - 1), 0)] if nc > 1 else [])
    params = rng.normal(size=(B, len(specs), 4)) * 3.0
    tfinal = np.sort(rng.uniform(0.5, 3.0, size=(B, nsec)), axis=1) * prob.unit_time
    P = prob.guess_batch(specs, params, wl.obj, tfinal=tfinal).cpu().numpy()
    G = og_numpy.Guess
    for b in range(B):
        ref.p = np.array(prob.p, dtype=float)
        t_all = prob.time_all_section        # (the facade's LGL nodes: the library's, ulps from SciPy's)
        ref.set_states_all_section(0, G.linear(t_all, params[b, 0, 0], params[b, 0, 1]))
        ref.set_states_all_section(1, G.cubic(t_all, *params[b, 1]))
        ref.set_controls_all_section(0, G.constant(t_all, params[b, 2, 0]))
        ref.set_states(ns - 1, nsec - 1, G.linear(prob.time[nsec - 1], params[b, 3, 0], params[b, 3, 1]))
        if nc > 1:
            ref.set_controls(nc - 1, 0, G.zeros(ref.time[0]))
        for s in range(nsec):
            ref.set_time_final(s, tfinal[b, s])
        scale = np.abs(ref.p).max()
        assert np.abs(P[b] - ref.p).max() <= GUESS_RTOL_CUBIC * scale
        lin_idx = np.arange(prob.index_states(0, 0), prob.index_states(0, 0) + prob.nodes[0])
        assert np.array_equal(P[b][lin_idx], ref.p[lin_idx])           # the linear block: bit-identical
        assert np.array_equal(P[b][-nsec:], ref.p[-nsec:])
    from opengoddard_b200 import capi
    with pytest.raises(capi.OgbError):                                    # overlapping specs are refused
        prob.guess_batch([("linear", ("state", 0), None), ("zeros", ("state", 0), 0)], params[:, :2], wl.obj)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg2_goddard50", "cfg5_lowthrust128"])
def test_device_jitter_matches_the_numpy_oracle_and_is_shard_invariant(api, name):
    import torch
    wl = workloads.build(name, api)
    prob = wl.prob
    eng = prob.compile(wl.obj)
    lb, ub = prob.bounds_arrays()
    B, seed = 64, 20261017
    P = prob.make_starts(B, wl.obj, seed=seed)
    ref = og_rng.jitter(np.tile(prob.p, (B, 1)), prob.number_of_section, seed, 0, 0.01, 0.05, lb, ub)
    got = P.cpu().numpy()
    assert np.abs(got - ref).max() <= JITTER_RTOL * np.abs(ref).max()
    assert (got >= lb).all() and (got <= ub).all()
    z = (got[:, :-prob.number_of_section] / np.where(prob.p[:-1] == 0, 1, prob.p[:-prob.number_of_section]) - 1.0) / 0.01
    moved = prob.p[:-prob.number_of_section] != 0
    assert abs(z[:, moved].mean()) < 0.05 and abs(z[:, moved].std() - 1.0) < 0.05      # N(0, 1) draws
    part = prob.make_starts(16, wl.obj, seed=seed, first=40)              # a shard: instances 40..55
    assert torch.equal(part, P[40:56])
    assert not torch.equal(prob.make_starts(B, wl.obj, seed=seed + 1), P)
    c, _ = eng.eval_fd(P)                                                 # and they are usable starts
    assert torch.isfinite(c).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg2_goddard50", "cfg3_goddard_knot30x2", "ex09_polar_tsto20x2"])
def test_device_trajectories_match_the_facade_accessors(api, name):
    wl = workloads.build(name, api)
    prob = wl.prob
    P = workloads.make_batch(wl, 5)
    T = prob.trajectories(P, wl.obj).cpu().numpy()
    ns, nc = prob.number_of_states[0], prob.number_of_controls[0]
    assert T.shape == (5, sum(prob.nodes), 1 + ns + nc)
    for b in range(5):
        prob.p = P[b].copy()
        assert np.array_equal(T[b, :, 0], prob.time_update())
        for a in range(ns):
            assert np.array_equal(T[b, :, 1 + a], prob.states_all_section(a))
        for k in range(nc):
            assert np.array_equal(T[b, :, 1 + ns + k], prob.controls_all_section(k))
