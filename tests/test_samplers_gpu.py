"""samplea (scalar and batched) and batched sampleb on the GPU against the compiled reference
(slice-sampler build), chain by chain under identical 48-bit streams.

Bars: every evaluated log-posterior within 1e-12 relative (abs floor 1) of the same
formula evaluated with the REFERENCE's table and libm; the streams stay in step (same number of uniforms
consumed) and the slice-sampler draws are then BIT-EQUAL to the reference's: a proposal is a function of the
stream and of earlier proposals only, the density decides nothing but accept / reject.  (sampleb with a = 0
is a closed-form draw divided by the device-summed Q: 1e-12.)"""
import ctypes as C
import math
import os

import numpy as np
import pytest

import libstb_b200 as stb
from tests import harness

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(harness.REF_SLICE_SO), reason="reference build not present")]

libc = C.CDLL(None)
libc.drand48.restype = C.c_double
libc.srand48.argtypes = [C.c_long]


def _ref():
    R = harness.ref(slice_build=True)
    d, u32p = C.c_double, C.POINTER(C.c_uint32)
    R.samplea.restype = d
    R.samplea.argtypes = [d, C.c_int, C.POINTER(C.c_int), u32p, C.POINTER(u32p), C.POINTER(C.POINTER(C.c_uint16)),
                          C.c_void_p, C.POINTER(d), C.c_void_p, C.c_int, C.c_int]
    R.sampleb.restype, R.sampleb.argtypes = d, [d, C.c_int, d, d, u32p, u32p, d, C.c_void_p, C.c_int, C.c_int]
    return R


def _counts(seed, I, K, nmax):
    """The C4 recipe of SURVEY.md 8d at reduced size: n = 1 + floor(u^3 nmax), t <= n^0.6."""
    libc.srand48(seed)
    n_rows, t_rows = [], []
    for _ in range(I):
        nr, tr = [], []
        for _ in range(K):
            n = 1 + int(libc.drand48() ** 3 * nmax)
            t = min(n, 1 + int(libc.drand48() * n ** 0.6))
            nr.append(n)
            tr.append(t)
        n_rows.append(nr)
        t_rows.append(tr)
    return stb.Counts(n_rows, t_rows)


def _aterms_reference(R, cts, bpar, x):
    """lib/samplea.c:46-83 evaluated with the reference's own table and libm's lgamma."""
    maxn = max(int(r.max()) for r in cts.n_rows) + 1
    maxt = max(int(r.max()) for r in cts.t_rows) + 1
    sp = R.S_make(maxn, maxt, maxn, maxt, x, 1)
    val = 0.0
    for i in range(cts.I):
        Ti = int(cts.T[i])
        val += Ti * math.log(x) + math.lgamma(Ti + bpar[i] / x) - math.lgamma(bpar[i] / x)
        for n, t in zip(cts.n_rows[i], cts.t_rows[i]):
            if n > 1:
                val += R.S_S(sp, int(n), int(t))
    R.S_free(sp)
    return val


def test_scalar_samplea_matches_reference():
    L, R = stb.lib(), _ref()
    cts = _counts(42, I=12, K=10, nmax=400)
    bpar = np.full(cts.I, 10.0)
    dp = C.POINTER(C.c_double)
    a_ref = a_our = 0.5
    for step in range(4):
        libc.srand48(500 + step)
        a_ref = R.samplea(a_ref, *cts.args(), None, bpar.ctypes.data_as(dp), None, 2, 0)
        s_ref = libc.drand48()
        libc.srand48(500 + step)
        a_our = L.samplea(a_our, *cts.args(), None, bpar.ctypes.data_as(dp), None, 2, 0)
        assert libc.drand48() == s_ref, "different numbers of draws consumed"
        assert a_our == a_ref
        assert 0.01 <= a_our <= 0.98
        a_our = a_ref


def test_batched_samplea_matches_reference_chain_by_chain():
    L, R = stb.lib(), _ref()
    cts = _counts(7, I=10, K=12, nmax=600)
    bpar = np.full(cts.I, 10.0)
    Cn = 24
    a0 = 0.05 + 0.9 * (np.arange(Cn) + 0.5) / Cn
    seeds = 12345 + np.arange(Cn)
    rng0 = np.array([L.stb_rng48_state(int(s)) for s in seeds], dtype=np.uint64)
    a1, rng1, st = stb.samplea_batch(a0, cts, bpar, rng0, loops=2, trace_cap=64)
    tx, tv, tn = st["trace"]
    assert st["evals"] >= int(tn.sum()) and st["rounds"] >= 3  # evals counts speculative proposals too
    dp = C.POINTER(C.c_double)
    for c in range(Cn):
        libc.srand48(int(seeds[c]))
        a_ref = R.samplea(float(a0[c]), *cts.args(), None, bpar.ctypes.data_as(dp), None, 2, 0)
        # stream in step: the state glibc is left in equals the chain's state
        nxt = libc.drand48()
        s = C.c_uint64(int(rng1[c]))
        assert L.stb_rng48_drand(C.byref(s)) == nxt, c
        assert a1[c] == a_ref, c
    # every evaluated log-posterior of a few chains against the reference's table + libm
    for c in (0, Cn // 2, Cn - 1):
        for k in range(min(int(tn[c]), 8)):
            ref = _aterms_reference(R, cts, bpar, float(tx[c, k]))
            assert abs(tv[c, k] - ref) <= 1e-12 * max(1.0, abs(ref)), (c, k, tv[c, k], ref)


def test_batched_samplea_per_chain_bpar_and_errors():
    L = stb.lib()
    cts = _counts(3, I=6, K=8, nmax=200)
    Cn = 5
    bpar = np.tile(np.linspace(2.0, 30.0, Cn)[:, None], (1, cts.I))
    rng0 = np.array([L.stb_rng48_state(100 + c) for c in range(Cn)], dtype=np.uint64)
    a1, _, _ = stb.samplea_batch(np.full(Cn, 0.4), cts, bpar, rng0, loops=1, bpar_per_chain=True)
    # chain c equals a one-chain run with its own bpar row
    for c in range(Cn):
        a_c, _, _ = stb.samplea_batch([0.4], cts, bpar[c], rng0[c:c + 1], loops=1)
        assert a_c[0] == a1[c]
    # start outside the slice bounds: the scalar sampler exits, the batched one names the chain
    with pytest.raises(RuntimeError, match=r"\(3\)"):
        stb.samplea_batch([0.5, 0.5, 0.99], cts, bpar[0], rng0[:3], loops=1)


@pytest.mark.parametrize("I,tmax", [(40, 30), (300, 80)])
def test_batched_sampleb_matches_reference_chain_by_chain(I, tmax):
    L, R = stb.lib(), _ref()
    g = np.random.default_rng(I)
    T = g.integers(1, tmax, size=I)
    N = T + g.integers(0, 500, size=I)
    N[1] = 0
    cts = stb.Counts([[int(x)] for x in N], [[int(min(x, 65535))] for x in T])
    cts.T = T.astype(np.uint32)
    cts.N = N.astype(np.uint32)
    Cn = 20
    apar = np.where(np.arange(Cn) % 5 == 0, 0.0, 0.05 + 0.9 * (np.arange(Cn) + 0.5) / Cn)
    b0 = np.linspace(1.0, 50.0, Cn)
    seeds = 777 + np.arange(Cn)
    rng0 = np.array([L.stb_rng48_state(int(s)) for s in seeds], dtype=np.uint64)
    b1, rng1, st = stb.sampleb_batch(b0, cts, 1.1, 20.0, apar, rng0, loops=2)
    u32p = C.POINTER(C.c_uint32)
    for c in range(Cn):
        libc.srand48(int(seeds[c]))
        b_ref = R.sampleb(float(b0[c]), I, 1.1, 20.0, cts.N.ctypes.data_as(u32p), cts.T.ctypes.data_as(u32p),
                          float(apar[c]), None, 2, 0)
        nxt = libc.drand48()
        s = C.c_uint64(int(rng1[c]))
        assert L.stb_rng48_drand(C.byref(s)) == nxt, c
        if apar[c] == 0:
            assert b1[c] == pytest.approx(b_ref, rel=1e-12), c
        else:
            assert b1[c] == b_ref, (c, apar[c])


def test_config4_recipe_small_slice():
    """BASELINE config 4 statistics (100 000 nodes: I=1000 x K=100, n <= 5000) with a few chains: the
    internal table is 5001 x <=167 per evaluation; results stay in bounds and chain 0 matches the
    reference."""
    L, R = stb.lib(), _ref()
    libc.srand48(12345)
    n_rows, t_rows = [], []
    for _ in range(1000):
        nr, tr = [], []
        for _ in range(100):
            n = 1 + int(libc.drand48() ** 3 * 5000)
            t = min(n, 1 + int(libc.drand48() * n ** 0.6))
            nr.append(n)
            tr.append(t)
        n_rows.append(nr)
        t_rows.append(tr)
    cts = stb.Counts(n_rows, t_rows)
    bpar = np.full(cts.I, 10.0)
    Cn = 6
    a0 = 0.05 + 0.9 * (np.arange(Cn) + 0.5) / Cn
    rng0 = np.array([L.stb_rng48_state(12345 + c) for c in range(Cn)], dtype=np.uint64)
    a1, rng1, st = stb.samplea_batch(a0, cts, bpar, rng0, loops=1)
    assert ((a1 >= 0.01) & (a1 <= 0.98)).all() and st["evals"] >= 2 * Cn
    b1, _, _ = stb.sampleb_batch(np.full(Cn, 10.0), cts, 1.1, 20.0, a1, rng1, loops=1)
    assert ((b1 >= 0.01) & (b1 <= 2000)).all()
    dp, u32p = C.POINTER(C.c_double), C.POINTER(C.c_uint32)
    for c in range(Cn):  # every chain, a and b, against the reference on one core (0.2 s per chain)
        libc.srand48(12345 + c)
        a_ref = R.samplea(float(a0[c]), *cts.args(), None, bpar.ctypes.data_as(dp), None, 1, 0)
        b_ref = R.sampleb(10.0, cts.I, 1.1, 20.0, cts.N.ctypes.data_as(u32p), cts.T.ctypes.data_as(u32p), a_ref, None, 1, 0)
        assert a1[c] == a_ref and b1[c] == b_ref, c


def _config4_counts():
    import bench

    return bench.config4_counts()


def test_config4_scale_64_chains_against_reference():
    """SURVEY.md 8d C4 at C = 64: 100 000 nodes, chain c started at a_c = 0.05 + 0.9 (c + 1/2) / C on the stream of
    srand48(12345 + c); every chain's a and b draw equals the reference's (one core, ~15 s)."""
    L, R = stb.lib(), _ref()
    cts = _config4_counts()
    bpar = np.full(cts.I, 10.0)
    Cn = 64
    a0 = 0.05 + 0.9 * (np.arange(Cn) + 0.5) / Cn
    rng0 = np.array([L.stb_rng48_state(12345 + c) for c in range(Cn)], dtype=np.uint64)
    a1, rng1, _ = stb.samplea_batch(a0, cts, bpar, rng0, loops=1)
    b1, rng2, _ = stb.sampleb_batch(np.full(Cn, 10.0), cts, 1.1, 20.0, a1, rng1, loops=1)
    dp, u32p = C.POINTER(C.c_double), C.POINTER(C.c_uint32)
    for c in range(Cn):
        libc.srand48(12345 + c)
        a_ref = R.samplea(float(a0[c]), *cts.args(), None, bpar.ctypes.data_as(dp), None, 1, 0)
        b_ref = R.sampleb(10.0, cts.I, 1.1, 20.0, cts.N.ctypes.data_as(u32p), cts.T.ctypes.data_as(u32p), a_ref, None, 1, 0)
        nxt = libc.drand48()
        s = C.c_uint64(int(rng2[c]))
        assert L.stb_rng48_drand(C.byref(s)) == nxt, c
        assert a1[c] == a_ref and b1[c] == b_ref, c


@pytest.mark.parametrize("Cn", [1, 4096])
def test_config4_scale_chain_counts(Cn):
    """C = 1 and C = 4096 chains of the config-4 statistics: in bounds, and a chain's draw does not depend on how many
    chains share its call (chains 0 and C-1 against one-chain calls)"""
    L = stb.lib()
    cts = _config4_counts()
    bpar = np.full(cts.I, 10.0)
    a0 = 0.05 + 0.9 * (np.arange(Cn) + 0.5) / Cn
    rng0 = np.array([L.stb_rng48_state(12345 + c) for c in range(Cn)], dtype=np.uint64)
    a1, rng1, st = stb.samplea_batch(a0, cts, bpar, rng0, loops=1)
    b1, _, _ = stb.sampleb_batch(np.full(Cn, 10.0), cts, 1.1, 20.0, a1, rng1, loops=1)
    assert ((a1 >= 0.01) & (a1 <= 0.98)).all() and ((b1 >= 0.01) & (b1 <= 2000)).all()
    for c in {0, Cn - 1}:
        a_c, r_c, _ = stb.samplea_batch(a0[c:c + 1], cts, bpar, rng0[c:c + 1], loops=1)
        b_c, _, _ = stb.sampleb_batch([10.0], cts, 1.1, 20.0, a_c, r_c, loops=1)
        assert a_c[0] == a1[c] and b_c[0] == b1[c], c


@pytest.mark.skipif(not os.path.exists(harness.REF_SO), reason="reference build not present")
def test_scalar_samplea_ars_mode_matches_reference_default_build():
    """STB_SAMPLER_ARS: samplea runs arms_simple over the GPU table density, like the reference's
    default (ARS) build (lib/samplea.c:209-215).  Same srand() seed: same number of uniforms
    consumed and draws within 1e-9 (the density's last bits differ: device table vs libm)."""
    L = stb.lib()
    R = C.CDLL(harness.REF_SO)  # default build = ARS configuration
    d, u32p = C.c_double, C.POINTER(C.c_uint32)
    R.samplea.restype = d
    R.samplea.argtypes = [d, C.c_int, C.POINTER(C.c_int), u32p, C.POINTER(u32p), C.POINTER(C.POINTER(C.c_uint16)),
                          C.c_void_p, C.POINTER(d), C.c_void_p, C.c_int, C.c_int]
    libc.srand.argtypes = [C.c_uint]
    libc.rand.restype = C.c_int
    cts = _counts(43, I=12, K=10, nmax=400)
    bpar = np.full(cts.I, 10.0)
    dp = C.POINTER(C.c_double)
    old = L.stb_set_sampler(1)
    try:
        a_ref = a_our = 0.5
        for step in range(6):
            libc.srand(900 + step)
            a_ref = R.samplea(a_ref, *cts.args(), None, bpar.ctypes.data_as(dp), None, 1, 0)
            r_ref = libc.rand()
            libc.srand(900 + step)
            a_our = L.samplea(a_our, *cts.args(), None, bpar.ctypes.data_as(dp), None, 1, 0)
            assert libc.rand() == r_ref, "different numbers of uniforms consumed"
            assert a_our == pytest.approx(a_ref, rel=1e-9)
            assert 0.01 <= a_our <= 0.98
            a_our = a_ref
    finally:
        L.stb_set_sampler(old)


def _ref_default():
    R = C.CDLL(harness.REF_SO)  # default build = ARS configuration
    d, u32p = C.c_double, C.POINTER(C.c_uint32)
    R.samplea.restype = d
    R.samplea.argtypes = [d, C.c_int, C.POINTER(C.c_int), u32p, C.POINTER(u32p), C.POINTER(C.POINTER(C.c_uint16)),
                          C.c_void_p, C.POINTER(d), C.c_void_p, C.c_int, C.c_int]
    R.sampleb.restype, R.sampleb.argtypes = d, [d, C.c_int, d, d, u32p, u32p, d, C.c_void_p, C.c_int, C.c_int]
    libc.srand.argtypes = [C.c_uint]
    libc.rand.restype = C.c_int
    return R


@pytest.mark.skipif(not os.path.exists(harness.REF_SO), reason="reference build not present")
def test_batched_samplea_ars_matches_reference_chain_by_chain():
    """stb_samplea_batch_ars: C chains of arms_simple in lock-step, one device batch per round,
    against the reference's default (ARS) build run chain by chain from the same srand() seeds:
    draws 1e-9, and every chain's rand() stream ends where the reference's does."""
    L, R = stb.lib(), _ref_default()
    cts = _counts(44, I=10, K=12, nmax=300)
    bpar = np.full(cts.I, 7.0)
    dp = C.POINTER(C.c_double)
    Cn = 24
    a0 = 0.05 + 0.9 * (np.arange(Cn) + 0.5) / Cn
    seeds = [700 + c for c in range(Cn)]
    a1, rnd, stats = stb.samplea_batch_ars(a0, cts, bpar, stb.rand31_states(seeds))
    assert stats["evals"] >= 3 * Cn and stats["rounds"] >= 4
    for c in range(Cn):
        libc.srand(seeds[c])
        a_ref = R.samplea(float(a0[c]), *cts.args(), None, bpar.ctypes.data_as(dp), None, 1, 0)
        assert a1[c] == pytest.approx(a_ref, rel=1e-9), c
        assert L.stb_rand31_next(rnd.ctypes.data + c * stb.RAND31_DTYPE.itemsize) == libc.rand(), c
        assert max(0.01, a0[c] - 0.2) <= a1[c] <= min(0.98, a0[c] + 0.2)


@pytest.mark.skipif(not os.path.exists(harness.REF_SO), reason="reference build not present")
def test_batched_sampleb_ars_matches_reference_chain_by_chain():
    L, R = stb.lib(), _ref_default()
    cts = _counts(45, I=60, K=8, nmax=200)
    u32p = C.POINTER(C.c_uint32)
    Cn = 20
    apar = np.where(np.arange(Cn) % 5 == 4, 0.0, 0.1 + 0.8 * np.arange(Cn) / Cn)  # every fifth chain is a DP (a = 0)
    b0 = np.linspace(0.5, 40.0, Cn)
    s48 = [300 + c for c in range(Cn)]
    s31 = [400 + c for c in range(Cn)]
    r48 = np.array([L.stb_rng48_state(s) for s in s48], dtype=np.uint64)
    b1, r48b, rnd, stats = stb.sampleb_batch_ars(b0, cts, 1.1, 20.0, apar, r48, stb.rand31_states(s31))
    for c in range(Cn):
        libc.srand48(s48[c])
        libc.srand(s31[c])
        b_ref = R.sampleb(float(b0[c]), cts.I, 1.1, 20.0, cts.N.ctypes.data_as(u32p), cts.T.ctypes.data_as(u32p),
                          float(apar[c]), None, 1, 0)
        assert b1[c] == pytest.approx(b_ref, rel=1e-9), (c, apar[c])
        assert L.stb_rand31_next(rnd.ctypes.data + c * stb.RAND31_DTYPE.itemsize) == libc.rand(), c
        assert 0.01 <= b1[c] <= 2000


@pytest.mark.skipif(not os.path.exists(harness.REF_SO), reason="reference build not present")
def test_batched_ars_at_config4_scale():
    """At config-4 scale (100 000 nodes) the discount's log-posterior is so peaked (|dy/da| ~ 1e6)
    that on [a - 0.2, a + 0.2] it is numerically a straight line: whether ARS finds it concave
    (gl < grl on nearly collinear chords, lib/arms.c:776-795) is decided by the last bits of the
    density, which differ between libm sums and the device's (1e-12 relative).  A chain then either
    lands at the end of its interval or stops with code 2000 and keeps its value (the reference
    ignores arms_simple's return value, lib/samplea.c:210-215) -- in EITHER library.  What is
    well-conditioned is checked: every chain consumes exactly the uniforms the reference consumes,
    stays inside its interval, and interior draws agree to 1e-5.  (Bit-identity of the sampler
    itself, failing chains included: tests/test_ars_cpu.py.)"""
    import bench

    L, R = stb.lib(), _ref_default()
    cts = bench.config4_counts()
    bpar = np.full(cts.I, 10.0)
    dp = C.POINTER(C.c_double)
    idx = [2, 300, 500, 913]
    a0 = np.array([0.05 + 0.9 * (c + 0.5) / 1024 for c in idx])
    seeds = [777 + c for c in idx]
    a1, rnd, _ = stb.samplea_batch_ars(a0, cts, bpar, stb.rand31_states(seeds))
    for j in range(len(idx)):
        libc.srand(seeds[j])
        a_ref = R.samplea(float(a0[j]), *cts.args(), None, bpar.ctypes.data_as(dp), None, 1, 0)
        assert L.stb_rand31_next(rnd.ctypes.data + j * stb.RAND31_DTYPE.itemsize) == libc.rand(), idx[j]
        lo, hi = max(0.01, a0[j] - 0.2), min(0.98, a0[j] + 0.2)
        assert lo <= a1[j] <= hi
        ends = (a0[j], lo, hi)
        interior = lambda v: all(abs(v - e) > 1e-5 for e in ends)
        if interior(a1[j]) and interior(a_ref):
            assert a1[j] == pytest.approx(a_ref, rel=1e-5), idx[j]
