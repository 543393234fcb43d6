"""Parity of the CUDA table fill (through the C ABI of libstb_b200.so) with the CPU oracle.

Bars:  FP64 log S, V, U, UV: |x-y| <= 1e-12*max(1,|y|)  (measured ~1e-14);
S_MIRROR_ORDER V: bit-identical;  S_FLOAT: one float ulp of the FP64 value (3e-7 stated bound).
"""
import json
import math
import os

import numpy as np
import pytest

import libstb_b200 as stb
from tests import harness

pytestmark = pytest.mark.gpu
FLAGS = stb.S_STABLE | stb.S_UVTABLE
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tables.json")))

SHAPES = [(10, 10, 0.0), (10, 10, 0.3), (64, 33, 0.5), (300, 300, 0.98), (300, 40, 0.01), (1500, 200, 0.7),
          (777, 129, 0.25), (4000, 1000, 0.5), (2500, 2500, 0.9), (3000, 70, 0.0)]


def _compare_full(t, N, M, a, rel):
    S, V = harness.oracle_tables(N, M, a)
    gS = t.rows(0, 1, N)[:, :M]
    gV = t.rows(1, 1, N)[:, :M]
    mS, mV = harness.valid_mask(N, M), harness.valid_mask(N, M, for_V=True)
    okS = harness.close(gS[mS], S[mS], rel)
    okV = harness.close(gV[mV], V[mV], rel)
    assert okS.all(), f"S: {np.count_nonzero(~okS)} cells off, max err {harness.max_err(gS[mS], S[mS]):.3g}"
    assert okV.all(), f"V: {np.count_nonzero(~okV)} cells off, max err {harness.max_err(gV[mV], V[mV]):.3g}"
    return gS, gV, S, V


@pytest.mark.parametrize("N,M,a", SHAPES)
def test_linear_fill_matches_oracle(N, M, a):
    t = stb.Table(N, M, N, M, a, FLAGS)
    assert (t.usedN, t.usedM) == (N, M)
    gS, gV, S, V = _compare_full(t, N, M, a, 1e-12)
    # the diagonal S^n_n = 1 is stored as exactly +0.0
    for n in range(2, min(N, M) + 1):
        assert gS[n - 1, n - 1] == 0.0
    t.free()


@pytest.mark.parametrize("N,M,a", SHAPES[:8])
def test_mirror_order_fill(N, M, a):
    t = stb.Table(N, M, N, M, a, FLAGS | stb.S_MIRROR_ORDER)
    S, V = harness.oracle_tables(N, M, a)
    gS = t.rows(0, 1, N)[:, :M]
    gV = t.rows(1, 1, N)[:, :M]
    mS, mV = harness.valid_mask(N, M), harness.valid_mask(N, M, for_V=True)
    assert np.array_equal(gV[mV], V[mV]), "mirror-order V must be bit-identical to the reference"
    assert harness.close(gS[mS], S[mS], 1e-13).all(), harness.max_err(gS[mS], S[mS])
    # S1 column is the host running sum: bit-identical
    assert np.array_equal(gS[:, 0], S[:, 0])
    t.free()


@pytest.mark.parametrize("flags", [FLAGS | stb.S_FLOAT, FLAGS | stb.S_FLOAT | stb.S_MIRROR_ORDER])
def test_float_storage(flags):
    N, M, a = 1200, 150, 0.6
    t = stb.Table(N, M, N, M, a, flags)
    S, V = harness.oracle_tables(N, M, a)
    gS = t.rows(0, 1, N)[:, :M]
    gV = t.rows(1, 1, N)[:, :M]
    mS, mV = harness.valid_mask(N, M), harness.valid_mask(N, M, for_V=True)
    # stored value is a float; it must be the FP64 value rounded, give or take one float ulp
    assert np.array_equal(gS[mS], gS[mS].astype(np.float32).astype(np.float64))
    assert harness.close(gS[mS], S[mS].astype(np.float32).astype(np.float64), 1.2e-7).all()
    assert harness.close(gV[mV], V[mV].astype(np.float32).astype(np.float64), 1.2e-7).all()
    frac_exact = np.mean(gS[mS] == S[mS].astype(np.float32).astype(np.float64))
    assert frac_exact > 0.99, frac_exact
    t.free()


def test_float_table_at_config1_size_fresh_and_grown():
    """S_FLOAT at the config-1 shape (the precision_test.c question): every stored cell is
    the FP64 value rounded to float give or take one float ulp, and a table GROWN to that size by
    look-ups equals the fresh one bit for bit (the reference's incremental float table drifts by up to
    2.3e-7 because its column extension re-reads float-rounded rows, lib/stable.c:399-400; a growth here
    refills the extent from FP64 state)."""
    N, M, a = 10000, 1000, 0.5
    fl = stb.S_STABLE | stb.S_UVTABLE | stb.S_FLOAT
    fresh = stb.Table(N, M, N, M, a, fl)
    grown = stb.Table(300, 60, N, M, a, fl)
    for n, m in ((900, 200), (4000, 700), (N - 1, M - 1)):
        grown.S(n, m), grown.V(n, m)
    assert (grown.usedN, grown.usedM) == (N, M)
    S, V = harness.oracle_tables(N, M, a)
    for which, ref in ((0, S), (1, V)):
        mask = harness.valid_mask(N, M, for_V=bool(which))
        g = fresh.rows(which, 1, N)[:, :M]
        assert np.array_equal(g[mask], grown.rows(which, 1, N)[:, :M][mask])
        want = ref.astype(np.float32).astype(np.float64)
        assert harness.close(g[mask], want[mask], 1.2e-7).all()
        assert np.mean(g[mask] == want[mask]) > 0.99
    fresh.free(), grown.free()


def test_float_table_answers_the_first_column_in_fp64():
    """S_S(n,1) is S_S1(n) (lib/stable.c:946-947): FP64 also for S_FLOAT tables -- scalar call, batched
    gather and sweep gather alike."""
    N, M, a = 900, 60, 0.45
    t = stb.Table(N, M, N, M, a, stb.S_STABLE | stb.S_FLOAT)
    S, _ = harness.oracle_tables(N, M, a, want_V=False)
    n = np.arange(2, N + 1, dtype=np.uint32)
    one = np.ones_like(n)
    got = t.S_batch(n, one)
    assert harness.close(got, S[n - 1, 0]).all()
    assert not np.array_equal(got, got.astype(np.float32).astype(np.float64)), "not float-rounded"
    assert all(t.S(int(v), 1) == g for v, g in zip(n[::97], got[::97]))
    t.free()
    w = stb.Sweep(N, M, stb.S_STABLE | stb.S_FLOAT)
    w.set_pairs(n, one)
    g, _, _ = w.run(np.array([a]), gather=True, sums=False)
    w.free()
    assert harness.close(g[0], S[n - 1, 0]).all()


@pytest.mark.parametrize("flags", [stb.S_STABLE, stb.S_UVTABLE])
def test_single_table_flags(flags):
    N, M, a = 900, 100, 0.35
    t = stb.Table(N, M, N, M, a, flags)
    S, V = harness.oracle_tables(N, M, a)
    if flags & stb.S_STABLE:
        g = t.rows(0, 1, N)[:, :M]
        m = harness.valid_mask(N, M)
        assert harness.close(g[m], S[m]).all()
        assert t.V(500, 20) == 0.0  # no V table: 0 (lib/stable.c:901-902)
    else:
        g = t.rows(1, 1, N)[:, :M]
        m = harness.valid_mask(N, M, for_V=True)
        assert harness.close(g[m], V[m]).all()
        assert t.S(500, 20) == -math.inf  # no S table: log 0 (lib/stable.c:942-943)
        # S1 still exists without the S table (lib/stable.c:338-348)
        assert harness.close(t.S1(500), S[499, 0]).all()
    t.free()


def test_scalar_lookups_and_conventions():
    """S_S/S_V/S_U/S_UV conventions of lib/stable.c:875-974 at the edge cells the survey lists."""
    O = harness.oracle()
    for extra in (0, stb.S_ASYMPT):
        t = stb.Table(30, 12, 30, 12, 0.3, FLAGS | extra)
        o = O.orc_make(30, 12, 30, 12, 0.3, FLAGS | extra)
        cells = [(2, 2), (2, 1), (0, 0), (5, 0), (3, 5), (5, 1), (5, 6), (5, 5), (31, 3), (20, 13), (30, 12),
                 (12, 12), (13, 12), (29, 11), (100, 4), (100000, 7), (17, 9), (30, 2)]
        for n, m in cells:
            assert harness.close(t.S(n, m), O.orc_S(o, n, m)).all(), ("S", n, m)
            if m >= 2:
                assert harness.close(t.V(n, m), O.orc_V(o, n, m)).all(), ("V", n, m)
            if m >= 1:
                assert harness.close(t.UV(n, m), O.orc_UV(o, n, m)).all(), ("UV", n, m)
                assert harness.close(t.U(n, m), O.orc_U(o, n, m)).all(), ("U", n, m)
            if n >= 2 and m >= 2:
                assert t.asympt(n, m) == O.orc_asympt(0.3, n, m)
        O.orc_free(o)
        t.free()


def test_golden_spots():
    for ent in GOLD["tables"] + GOLD["big"]:
        N, M, a = ent["N"], ent["M"], ent["a"]
        t = stb.Table(N, M, N, M, a, FLAGS)
        for sp in ent["spots"]:
            n, m = sp["n"], sp["m"]
            assert harness.close(t.S(n, m), float(sp["S"])).all(), (N, M, a, n, m)
            if sp.get("V") is not None:
                assert harness.close(t.V(n, m), float(sp["V"])).all(), (N, M, a, n, m)
            if sp.get("U") is not None:
                assert harness.close(t.U(n, m), float(sp["U"])).all()
                assert harness.close(t.UV(n, m), float(sp["UV"])).all()
        t.free()


def test_remake_changes_discount():
    N, M = 800, 90
    t = stb.Table(N, M, N, M, 0.2, FLAGS)
    for a in (0.75, 0.0, 0.5):
        t.remake(a)
        _compare_full(t, N, M, a, 1e-12)
    t.free()


def test_growth_on_lookup_matches_fresh_build():
    """S_S/S_V past the filled extent grow the table (lib/stable.c:564-815: >=10%, >=+50, capped);
    an extended table equals a freshly built one bit for bit."""
    a = 0.45
    t = stb.Table(100, 20, 5000, 400, a, FLAGS)
    assert (t.usedN, t.usedM) == (100, 20)
    v = t.S(150, 10)  # N request 151 -> +1 -> 152 >= max(110, 150) ; M unchanged
    assert t.usedN == 152 and t.usedM == 20
    v2 = t.S(152, 30)  # M grows: request 31 -> 32 -> at least 20+50
    assert t.usedM == 70
    t.S(3000, 300)
    assert t.usedN >= 3001 and t.usedM >= 301
    N, M = t.usedN, t.usedM
    fresh = stb.Table(N, M, N, M, a, FLAGS)
    for which in (0, 1):
        assert np.array_equal(np.nan_to_num(t.rows(which, 1, N)[:, :M][harness.valid_mask(N, M, which == 1)]),
                              np.nan_to_num(fresh.rows(which, 1, N)[:, :M][harness.valid_mask(N, M, which == 1)]))
    S, _ = harness.oracle_tables(152, 70, a, want_V=False)
    assert harness.close(v, S[149, 9]).all() and harness.close(v2, S[151, 29]).all()
    # capped by the maxima; beyond them: log 0 / 0
    t.S(5000, 400)
    assert (t.usedN, t.usedM) == (5000, 400)
    assert t.S(5001, 3) == -math.inf and t.V(5001, 3) == 0.0 and t.S(4000, 401) == -math.inf
    # deviation from the reference, on purpose: m beyond the old usedN is served
    t2 = stb.Table(10, 10, 6000, 400, a, FLAGS)
    S2, _ = harness.oracle_tables(5000, 300, a, want_V=False)
    assert harness.close(t2.S(5000, 300), S2[4999, 299]).all()
    t.free(), fresh.free(), t2.free()


def test_batch_api_matches_scalar():
    N, M, a = 3000, 250, 0.55
    t = stb.Table(N, M, N, M, a, FLAGS)
    rng = np.random.default_rng(7)
    n = rng.integers(0, N + 20, size=20000).astype(np.uint32)
    m = rng.integers(0, M + 5, size=20000).astype(np.uint32)
    gs, gv = t.S_batch(n, m), t.V_batch(n, m)
    for i in range(0, 20000, 37):
        ni, mi = int(n[i]), int(m[i])
        es = t.S(ni, mi) if mi != 1 else None  # m==1 reads S1 host cache; device reads column 1: same values
        if es is not None:
            assert gs[i] == es or (math.isinf(es) and math.isinf(gs[i])), (ni, mi, gs[i], es)
        if mi >= 2:
            assert gv[i] == t.V(ni, mi), (ni, mi)
    k = (m == 1) & (n >= 1) & (n <= N)
    assert np.array_equal(gs[k], np.array([t.S1(int(x)) for x in n[k]]))
    t.free()


@pytest.mark.parametrize("K", [1, 3, 5, 7])
def test_bit_reproducible_across_geometries(K, monkeypatch):
    """Scaling by powers of two is exact, so the strip width (columns per lane) cannot change
    a single bit of the result."""
    N, M, a = 2000, 700, 0.65
    monkeypatch.delenv("STB_STRIP_K", raising=False)
    base = stb.Table(N, M, N, M, a, FLAGS)
    monkeypatch.setenv("STB_STRIP_K", str(K))
    t = stb.Table(N, M, N, M, a, FLAGS)
    for which in (0, 1):
        mask = harness.valid_mask(N, M, which == 1)
        assert np.array_equal(t.rows(which, 1, N)[:, :M][mask], base.rows(which, 1, N)[:, :M][mask])
    t.free(), base.free()


def test_config1_full_table():
    """BASELINE config 1: S_make N=10,000 M=1,000 a=0.5 with U/V, every cell vs the oracle."""
    N, M, a = 10000, 1000, 0.5
    t = stb.Table(N, M, N, M, a, FLAGS)
    gS, gV, S, V = _compare_full(t, N, M, a, 1e-12)
    assert harness.close(t.S(10000, 1000), 76855.523932738957).all()
    assert harness.close(t.V(9998, 998), 0.0019008494057812999).all()
    # U and UV derived from V
    n = np.arange(2, N + 1)
    for m in (2, 17, 500, 1000):
        nn = n[n > m]
        u_or = nn - m * a + 1.0 / V[nn - 1, m - 1]
        u_g = np.array([t.U(int(x), m) for x in nn[:: max(1, len(nn) // 200)]])
        assert harness.close(u_g, u_or[:: max(1, len(nn) // 200)]).all()
    t.free()


def test_lazy_mirror_reads(monkeypatch):
    """Large tables are not mirrored eagerly; scalar look-ups fetch row blocks on demand."""
    monkeypatch.setenv("STB_MIRROR_EAGER_BYTES", "0")
    N, M, a = 6000, 300, 0.4
    t = stb.Table(N, M, N, M, a, FLAGS)
    S, V = harness.oracle_tables(N, M, a)
    rng = np.random.default_rng(3)
    for _ in range(300):
        n = int(rng.integers(3, N + 1))
        m = int(rng.integers(2, min(n - 1, M) + 1))
        assert harness.close(t.S(n, m), S[n - 1, m - 1]).all()
        assert harness.close(t.V(n, m), V[n - 1, m - 1]).all()
    t.free()


# Column counts around the planner's shape boundaries (one strip of 5 x 32 = 160 or 7 x 32 = 224
# columns per CTA, lane counts rounded up to 32-byte sectors), tiny tables, M == N, a at both ends.
EDGE_SHAPES = [(2, 1, 0.5), (2, 2, 0.5), (3, 3, 0.0), (40, 4, 0.3), (400, 31, 0.7), (400, 32, 0.2), (400, 33, 0.9),
               (700, 159, 0.5), (700, 160, 0.01), (700, 161, 0.98), (900, 223, 0.6), (900, 224, 0.0), (900, 225, 0.4),
               (1000, 320, 0.75), (1000, 321, 0.33), (1300, 1121, 0.5), (1125, 1125, 0.15), (5000, 483, 0.66)]


@pytest.mark.parametrize("N,M,a", EDGE_SHAPES)
def test_plan_boundaries_match_oracle(N, M, a):
    """every stored cell against the oracle at the widths where the launch geometry changes"""
    t = stb.Table(N, M, N, M, a, FLAGS)
    assert (t.usedN, t.usedM) == (N, M) or M < 10 or N < 10  # (S_make raises tiny extents, lib/stable.c:118-129)
    Nu, Mu = t.usedN, t.usedM
    _compare_full(t, Nu, Mu, a, 1e-12)
    t.free()


def test_random_shapes_match_oracle():
    """twenty seeded random extents / discounts, S+V, every cell"""
    rng = np.random.default_rng(20261018)
    for _ in range(20):
        M = int(rng.integers(10, 700))
        N = M + int(rng.integers(0, 1500))
        a = float(rng.choice([0.0, rng.uniform(0.01, 0.98)]))
        fl = int(rng.choice([FLAGS, FLAGS | stb.S_FLOAT]))
        t = stb.Table(N, M, N, M, a, fl)
        if fl & stb.S_FLOAT:
            S, V = harness.oracle_tables(N, M, a)
            gS, gV = t.rows(0, 1, N)[:, :M], t.rows(1, 1, N)[:, :M]
            mS, mV = harness.valid_mask(N, M), harness.valid_mask(N, M, for_V=True)
            assert harness.close(gS[mS], S[mS].astype(np.float32).astype(np.float64), 1.2e-7).all(), (N, M, a)
            assert harness.close(gV[mV], V[mV].astype(np.float32).astype(np.float64), 1.2e-7).all(), (N, M, a)
        else:
            _compare_full(t, N, M, a, 1e-12)
        t.free()


@pytest.mark.parametrize("slots", [1, 2, 3])
def test_several_passes_over_the_columns(slots, monkeypatch):
    """Tables wider than one launch's CTAs can hold are filled in passes, the boundary column
    carried between passes in a linear buffer.  STB_STRIP_SLOTS shrinks a launch to `slots` CTAs so
    that a small table needs 2..5 passes: every cell against the oracle, and bit-identical to the
    single-pass fill."""
    N, M, a = 1500, 1000, 0.6
    monkeypatch.delenv("STB_STRIP_SLOTS", raising=False)
    base = stb.Table(N, M, N, M, a, FLAGS)
    monkeypatch.setenv("STB_STRIP_SLOTS", str(slots))
    t = stb.Table(N, M, N, M, a, FLAGS)
    gS, gV, _, _ = _compare_full(t, N, M, a, 1e-12)
    for which, g in ((0, gS), (1, gV)):
        mask = harness.valid_mask(N, M, which == 1)
        assert np.array_equal(g[mask], base.rows(which, 1, N)[:, :M][mask])
    t.remake(0.3)  # and again through S_remake
    _compare_full(t, N, M, 0.3, 1e-12)
    t.free()
    base.free()


def test_concurrent_lookups_with_growth_S_THREADS():
    """S_THREADS (lib/stable.h:43): look-ups from several threads while the table grows on demand
    (growth refills the table and replaces the host mirror, so readers hold a shared lock).
    ctypes releases the GIL during the C calls, so the eight threads really overlap."""
    import threading

    N0, M0, Nmax, Mmax, a = 200, 40, 6000, 300, 0.45
    t = stb.Table(N0, M0, Nmax, Mmax, a, FLAGS | stb.S_THREADS)
    S, V = harness.oracle_tables(Nmax, Mmax, a)
    errs = []

    def work(seed):
        rng = np.random.default_rng(seed)
        top = 300
        for it in range(400):
            top = min(Nmax - 2, int(top * 1.02) + 5)  # reach further and further: growth on the way
            n = int(rng.integers(3, top))
            m = int(rng.integers(2, min(n, Mmax - 2)))
            if seed % 2 and it % 8 == 0:  # batched look-ups (one kernel on the handle's stream) in between
                nn = rng.integers(3, top, size=64).astype(np.uint32)
                mm = np.minimum(rng.integers(2, Mmax - 2, size=64), nn - 1).astype(np.uint32)
                if not harness.close(t.S_batch(nn, mm), S[nn - 1, mm - 1]).all():
                    errs.append(("batch", int(nn[0]), int(mm[0])))
            s, v = t.S(n, m), t.V(n, m)
            if not (abs(s - S[n - 1, m - 1]) <= 1e-12 * max(1.0, abs(S[n - 1, m - 1])) and
                    abs(v - V[n - 1, m - 1]) <= 1e-12 * max(1.0, abs(V[n - 1, m - 1]))):
                errs.append((n, m, s, S[n - 1, m - 1], v, V[n - 1, m - 1]))

    threads = [threading.Thread(target=work, args=(100 + i,)) for i in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errs, errs[:3]
    assert t.usedN > N0 and t.usedM > M0
    t.free()


def test_batched_U_and_UV_match_the_scalar_calls():
    """stb_U_batch / stb_UV_batch (one gather kernel on the V table) against S_U / S_UV, including
    their special cases (lib/stable.c:875-897): m == 1, m == n, m == n + 1"""
    N, M, a = 900, 120, 0.4
    t = stb.Table(N, M, N, M, a, FLAGS)
    rng = np.random.default_rng(8)
    n = rng.integers(3, N - 2, size=3000).astype(np.uint32)
    m = np.minimum(rng.integers(1, M - 2, size=3000), n).astype(np.uint32)
    n[:4], m[:4] = [50, 60, 70, 80], [1, 60, 71, 2]  # m == 1, m == n, m == n + 1, m == 2
    U, UV = t.U_batch(n, m), t.UV_batch(n, m)
    for i in range(len(n)):
        ni, mi = int(n[i]), int(m[i])
        if mi <= ni:
            assert U[i] == t.U(ni, mi), (ni, mi)
        assert UV[i] == t.UV(ni, mi), (ni, mi)
    t.free()
