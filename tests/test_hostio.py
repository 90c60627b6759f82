"""The host-buffer entry point (ogb_host_session_* / ogb_host_eval_fd / ogb_host_expand,
include/ogb200.h).  CPU part: the host half of the packed Jacobian transport against numpy.
GPU part: every transport mode returns, in HOST memory, exactly the c and J the device-resident
path computes, and those match the reference goldens / the oracle."""
import ctypes

import numpy as np
import pytest

from opengoddard_b200 import capi
from tests.helpers import assert_c_close, assert_J_close, golden, stacked_reference


# ------------------------------------------------------------------ CPU: ogb_host_expand
@pytest.mark.parametrize("nM,nnz,B,threads", [(1, 1, 3, 1), (7, 0, 2, 2), (511, 40, 5, 3), (512, 512, 2, 1),
                                               (201 * 457, 9349, 9, 4), (1537, 100, 1, 8)])
def test_expand_dense_matches_numpy(nM, nnz, B, threads):
    rng = np.random.default_rng(nM + nnz)
    lin = np.sort(rng.choice(nM, nnz, replace=False)).astype(np.uint32)
    vals = rng.standard_normal((B, nnz))
    ref = np.zeros((B, nM))
    ref[:, lin] = vals
    for shift in (0, 1, 3):                                   # destination not 64-byte aligned
        raw = np.full(B * nM + shift + 8, 7.0)
        out = raw[shift:shift + B * nM].reshape(B, nM)
        capi.host_expand(vals, lin, nM, out=out, mode="dense", threads=threads)
        assert (out == ref).all()
        assert (raw[:shift] == 7.0).all() and (raw[shift + B * nM:] == 7.0).all()   # no overrun


def test_expand_keep_zeros_only_touches_the_pattern():
    rng = np.random.default_rng(5)
    nM, nnz, B = 1000, 77, 4
    lin = np.sort(rng.choice(nM, nnz, replace=False)).astype(np.uint32)
    vals = rng.standard_normal((B, nnz))
    out = np.full((B, nM), -3.0)
    capi.host_expand(vals, lin, nM, out=out, mode="keep_zeros", threads=2)
    mask = np.zeros(nM, dtype=bool)
    mask[lin] = True
    assert (out[:, mask] == vals).all() and (out[:, ~mask] == -3.0).all()


def test_expand_rejects_a_bad_pattern():
    vals = np.zeros((1, 2))
    with pytest.raises(capi.OgbError):
        capi.host_expand(vals, np.array([3, 3], dtype=np.uint32), 10)
    with pytest.raises(capi.OgbError):
        capi.host_expand(vals, np.array([3, 10], dtype=np.uint32), 10)


def test_host_stats_layout():
    assert ctypes.sizeof(capi.OgbHostStats) == 56


# ------------------------------------------------------------------ GPU: the session
GPU_CFGS = ["cfg1_brachistochrone20", "cfg2_goddard50", "cfg3_goddard_knot30x2", "cfg4_polar3x40",
            "cfg5_lowthrust128", "ex05_goddard_knot25x2", "ex09_polar_tsto20x2", "ex10_lowthrust100",
            "edge_table_lookup", "edge_stress_mixed", "edge_all_ops", "edge_nonautonomous", "edge_picked_dynamics"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", GPU_CFGS)
def test_host_session_modes_are_bit_identical_to_the_device_path(api, name):
    import torch
    from opengoddard_b200 import workloads
    wl = workloads.build(name, api)
    eng = wl.prob.compile(wl.obj)
    B, n, M = 37, eng.nvars, eng.nrows
    P = workloads.make_batch(wl, B)
    c_d, J_d = eng.eval_fd(P)
    torch.cuda.synchronize()
    c_ref, J_ref = c_d.cpu().numpy(), J_d.cpu().numpy()
    lin = eng.jac_pattern()
    assert (np.diff(lin.astype(np.int64)) > 0).all() and lin[-1] < n * M
    outside = np.ones(n * M, dtype=bool)
    outside[lin] = False
    assert (J_ref.reshape(B, -1)[:, outside] == 0.0).all()       # nothing non-zero outside the pattern
    S = eng.host_session(64, chunk=8, threads=3)                  # 5 chunks, the last one ragged
    c = np.empty((B, M))
    J = np.full((B, n, M), np.nan)
    S.eval_fd(P, c, J, mode="dense")
    assert (c == c_ref).all() and (J == J_ref).all()
    c[:] = 0
    J.reshape(B, -1)[:, lin] = np.nan                             # zero background stays, pattern refilled
    S.eval_fd(P, c, J, mode="keep_zeros")
    assert (c == c_ref).all() and (J == J_ref).all()
    c2, V = S.eval_fd(P, mode="packed")
    assert (c2 == c_ref).all() and (V == J_ref.reshape(B, -1)[:, lin]).all()
    assert (capi.host_expand(V, lin, n * M).reshape(B, n, M) == J_ref).all()
    c3, J3 = S.eval_fd(torch.from_numpy(P).pin_memory(), mode="dma")
    assert (c3 == c_ref).all() and (J3 == J_ref).all()
    st = S.stats()
    assert st.nnz == len(lin) and st.chunk == 8 and st.nchunks == 5 and st.launches in (5, 10)    # (one launch per chunk when D.X is fused: small chunks)
    S.eval_fd(P[:1], c[:1], J[:1], mode="dense")                  # a batch of one through the same session
    assert (J[0] == J_ref[0]).all()
    with pytest.raises(capi.OgbError):
        S.eval_fd(np.zeros((65, n)), mode="packed")               # larger than max_batch
    S.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg2_goddard50", "cfg3_goddard_knot30x2", "ex09_polar_tsto20x2"])
def test_host_session_vs_reference_golden(api, name):
    from opengoddard_b200 import workloads
    g = golden("workload_" + name)
    wl = workloads.build(name, api)
    eng = wl.prob.compile(wl.obj)
    c_ref, J_ref = stacked_reference(g)
    P = np.ascontiguousarray(g["P"])
    S = eng.host_session(len(P))
    c, J = S.eval_fd(P, mode="dense")
    assert_c_close(c, c_ref, J_ref, P)
    assert_J_close(J.transpose(0, 2, 1), J_ref)
    S.close()


@pytest.mark.gpu
def test_host_session_full_size_roundtrip(api):
    """Goddard-50 x 4096 (BASELINE size) through the packed transport: dense host J == device J."""
    import torch
    from opengoddard_b200 import workloads
    wl = workloads.build("cfg2_goddard50", api)
    eng = wl.prob.compile(wl.obj)
    B = 4096
    P = workloads.make_batch(wl, B)
    c_d, J_d = eng.eval_fd(P)
    S = eng.host_session(B)
    c, J = S.eval_fd(P, mode="dense")
    Jt = torch.from_numpy(J).to(J_d.device)
    assert bool((Jt == J_d).all()) and (c == c_d.cpu().numpy()).all()
    S.close()


@pytest.mark.gpu
def test_facade_evaluate_batch_host(api):
    import torch
    from opengoddard_b200 import workloads
    wl = workloads.build("cfg3_goddard_knot30x2", api)
    P = workloads.make_batch(wl, 9)
    c_d, J_d = wl.prob.evaluate_batch(P, wl.obj)
    c_h, J_h = wl.prob.evaluate_batch(P, wl.obj, host=True)
    assert isinstance(c_h, np.ndarray) and (c_h == c_d.cpu().numpy()).all() and (J_h == J_d.cpu().numpy()).all()
    assert (wl.prob.evaluate_batch(P, wl.obj, jacobian=False, host=True) == c_h).all()


@pytest.mark.parametrize("isa", ["0", "1", "2"])
def test_expand_dense_every_store_flavour(monkeypatch, isa):
    """SSE2 / AVX2 / AVX-512 non-temporal store paths of the host expansion ($OGB200_HOST_ISA; a flavour
    the CPU lacks falls back to the next one) give the same bytes, for aligned and unaligned rows and
    for patterns with long zero runs, dense runs and values in the first / last line."""
    monkeypatch.setenv("OGB200_HOST_ISA", isa)
    rng = np.random.default_rng(int(isa) + 11)
    for nM, runs in ((4099, [(0, 3), (17, 40), (4000, 99)]), (1000, [(5, 1), (999, 1)]), (65, [(0, 65)]), (9, [])):
        lin = np.array(sorted({p for a, k in runs for p in range(a, min(nM, a + k))}), dtype=np.uint32)
        B = 3
        vals = rng.standard_normal((B, len(lin)))
        ref = np.zeros((B, nM))
        ref[:, lin] = vals
        for shift in (0, 1, 5):
            raw = np.full(B * nM + shift + 8, -9.0)
            out = raw[shift:shift + B * nM].reshape(B, nM)
            capi.host_expand(vals, lin, nM, out=out, mode="dense", threads=2)
            assert (out == ref).all() and (raw[:shift] == -9.0).all() and (raw[shift + B * nM:] == -9.0).all()
