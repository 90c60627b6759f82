"""numpy restatement of the device-side multi-start generator (ogb_jitter, csrc/ogb_guess.cuh).

TEST INFRASTRUCTURE ONLY.  The device draws, for instance i and variable j, 128 bits from
Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) with
key = the 64-bit seed and counter = (j, 0, i_lo, i_hi), turns them into two 53-bit uniforms and applies
    state / control entries  p * (1 + rel_x * z),  z = sqrt(-2 ln(1 - u1)) cos(2 pi u2)      (Box-Muller)
    final times              p * (1 + rel_t * (2 u1 - 1))
then clips into the bounds -- the multi-start perturbation of SURVEY.md section 8(d), made counter-based so
that any sub-range of instances is reproducible on any number of GPUs.  (bench.py's workloads.make_batch keeps
numpy's default_rng: that is the benchmark's input definition; this is the device-side equivalent.)
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32: uint32 arrays (or scalars) in, four uint32 arrays out."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint32) for v in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.asarray(k0, dtype=np.uint32)
    k1 = np.asarray(k1, dtype=np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(rounds):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n1 = p1.astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            n3 = p0.astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return c0, c1, c2, c3


def u53(a, b):
    return ((a >> np.uint32(5)).astype(np.float64) * 67108864.0 + (b >> np.uint32(6)).astype(np.float64)) / 9007199254740992.0


def jitter(P, nsec, seed, first=0, rel_x=0.01, rel_t=0.05, lb=None, ub=None):
    """The perturbed copy of P (B, n) ogb_jitter produces in place."""
    P = np.array(P, dtype=np.float64)
    B, n = P.shape
    j = np.arange(n, dtype=np.uint64)[None, :]
    inst = (np.uint64(first) + np.arange(B, dtype=np.uint64))[:, None]
    seed = np.uint64(seed)
    r = philox4x32((j & np.uint64(0xFFFFFFFF)).astype(np.uint32), np.uint32(0),
                   (inst & np.uint64(0xFFFFFFFF)).astype(np.uint32), (inst >> np.uint64(32)).astype(np.uint32),
                   np.uint32(seed & np.uint64(0xFFFFFFFF)), np.uint32(seed >> np.uint64(32)))
    u1, u2 = u53(r[0], r[1]), u53(r[2], r[3])
    z = np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(6.283185307179586477 * u2)
    out = P * (1.0 + rel_x * z)
    out[:, n - nsec:] = (P * (1.0 + rel_t * (2.0 * u1 - 1.0)))[:, n - nsec:]
    if lb is not None:
        out = np.maximum(out, lb)
    if ub is not None:
        out = np.minimum(out, ub)
    return out
