"""the device-side consumer of the V table -- table-indicator Gibbs sweeps (stb_ti_gibbs,
csrc/gibbs_cuda.cu) against the oracle's restatement of the reference's per-token update (test/demo.c:405-434;
oracle/stirling_oracle.c::orc_ti_gibbs: the reference's arithmetic on its own V table, uniforms from glibc's
erand48).  Bar: every table count and every stream state equal."""
import ctypes as C

import numpy as np
import pytest

import libstb_b200 as stb
from tests import harness

pytestmark = pytest.mark.gpu


def _restaurants(R, D, ntok_mean, seed):
    """R restaurants over D dishes: Zipf-ish dish popularity, every restaurant's tokens in arrival order"""
    rng = np.random.default_rng(seed)
    pop = 1.0 / np.arange(1, D + 1) ** 0.8
    pop /= pop.sum()
    tok, off = [], [0]
    n = np.zeros((R, D), dtype=np.uint32)
    for j in range(R):
        k = max(1, int(rng.poisson(ntok_mean)))
        d = rng.choice(D, size=k, p=pop).astype(np.uint32)
        tok.append(d)
        off.append(off[-1] + k)
        n[j] = np.bincount(d, minlength=D)
    t = (n > 0).astype(np.uint16)  # demo.c:391-401: every dish starts at one table
    T = t.sum(axis=1).astype(np.uint32)
    H = (pop * D).astype(np.float32)
    return np.concatenate(tok), np.array(off, dtype=np.uint32), n, t, T, H


def _oracle_sweeps(a, bpar, N, M, tok, off, n, t, T, H, rng, shared, sweeps):
    O = harness.oracle()
    tb = O.orc_make(N, M, N, M, a, 1 | 2)
    t, T, rng = t.copy(), T.copy(), rng.copy()
    u32p = C.POINTER(C.c_uint32)
    O.orc_ti_gibbs(tb, a, bpar, n.shape[0], off.ctypes.data_as(u32p), tok.ctypes.data_as(u32p),
                   H.ctypes.data_as(C.POINTER(C.c_float)), n.shape[1], n.ctypes.data_as(u32p),
                   t.ctypes.data_as(C.POINTER(C.c_uint16)), T.ctypes.data_as(u32p), rng.ctypes.data_as(C.POINTER(C.c_uint64)),
                   int(shared), sweeps)
    O.orc_free(tb)
    return t, T, rng


def _seed48(s):
    return np.uint64((int(s) & 0xFFFFFFFF) << 16 | 0x330E)


@pytest.mark.parametrize("a,bpar,shared", [(0.5, 10.0, True), (0.5, 10.0, False), (0.0, 3.0, False), (0.9, 0.5, True)])
def test_ti_gibbs_equals_reference_arithmetic(a, bpar, shared):
    """Mirror-order tables hold the reference's V bit for bit (IEEE + - x / only), so every acceptance test sees the
    float the reference sees: counts and streams must be EQUAL after several sweeps, in the demo's one-stream
    schedule and with one stream per restaurant."""
    R, D = 40, 12
    tok, off, n, t, T, H = _restaurants(R, D, 150, seed=5)
    N = int(n.max())
    M = N
    tab = stb.Table(N, M, N, M, a, stb.S_STABLE | stb.S_UVTABLE | stb.S_MIRROR_ORDER)
    rng = np.array([_seed48(77)] if shared else [_seed48(1000 + j) for j in range(R)], dtype=np.uint64)
    t1, T1, r1 = tab.ti_gibbs(bpar, off, tok, H, n, t, T, rng, shared_stream=shared, sweeps=3)
    t2, T2, r2 = _oracle_sweeps(a, bpar, N, M, tok, off, n, t, T, H, rng, shared, 3)
    assert np.array_equal(t1, t2) and np.array_equal(T1, T2) and np.array_equal(r1, r2)
    assert np.array_equal(T1, t1.sum(axis=1)) and (t1 <= n).all() and (t1[n > 0] >= 1).all()
    assert (t1 != t).any()  # the sweeps did move indicators
    # three calls of one sweep continue where the last one stopped: same result as one call of three
    ta, Ta, ra = t, T, rng
    for _ in range(3):
        ta, Ta, ra = tab.ti_gibbs(bpar, off, tok, H, n, ta, Ta, ra, shared_stream=shared, sweeps=1)
    assert np.array_equal(ta, t1) and np.array_equal(ra, r1)
    tab.free()


def test_ti_gibbs_on_the_throughput_table_and_growth():
    """The strip kernel's V agrees with the reference's to ~1e-15, far below the float `one` is rounded to: the
    sampled counts equal the reference arithmetic's here too (a flipped acceptance would need a uniform within
    1e-15 of its threshold).  The table starts small and is grown by the call to cover the counts."""
    R, D = 300, 20
    tok, off, n, t, T, H = _restaurants(R, D, 400, seed=9)
    N = int(n.max())
    a, bpar = 0.7, 5.0
    tab = stb.Table(50, 50, N + 100, N + 100, a, stb.S_STABLE | stb.S_UVTABLE)
    rng = np.array([_seed48(31 * j + 7) for j in range(R)], dtype=np.uint64)
    t1, T1, r1 = tab.ti_gibbs(bpar, off, tok, H, n, t, T, rng, sweeps=2)
    assert tab.last_gibbs_ms > 0
    t2, T2, r2 = _oracle_sweeps(a, bpar, N, N, tok, off, n, t, T, H, rng, False, 2)
    assert np.array_equal(r1, r2), "streams out of step: an acceptance test went the other way"
    assert np.array_equal(t1, t2) and np.array_equal(T1, T2)
    tab.free()


def test_ti_gibbs_argument_checks():
    tok, off, n, t, T, H = _restaurants(3, 4, 30, seed=1)
    N = int(n.max())
    noV = stb.Table(N, N, N, N, 0.5, stb.S_STABLE)
    rng = np.array([_seed48(j) for j in range(3)], dtype=np.uint64)
    with pytest.raises(RuntimeError, match="no V"):
        noV.ti_gibbs(1.0, off, tok, H, n, t, T, rng)
    noV.free()
    tab = stb.Table(N, N, N, N, 0.5, stb.S_STABLE | stb.S_UVTABLE)
    with pytest.raises(RuntimeError, match="sum of t"):
        tab.ti_gibbs(1.0, off, tok, H, n, t, T + 1, rng)
    small = stb.Table(4, 4, 4, 4, 0.5, stb.S_STABLE | stb.S_UVTABLE)
    if N > 4:
        with pytest.raises(RuntimeError, match="maximum size"):
            small.ti_gibbs(1.0, off, tok, H, n, t, T, rng)
    small.free()
    tab.free()
