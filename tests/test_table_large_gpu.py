"""BASELINE configs at full size, checked through size-independent properties and the spot
values the reference produced for them.  The CPU oracle needs ~150 s and 31 GB
for config 2, so the full-size check uses: golden spot cells; the column-prefix property (a
table's first M' columns do not depend on M, so they must equal the oracle run at M'=256);
identities S^n_n=1, S^n_{n-1}=n(n-1)(1-a)/2, S(n,1)=lgamma(n-a)-lgamma(1-a), U=S(n+1)-S(n)."""
import math

import numpy as np
import pytest

import libstb_b200 as stb
from tests import harness

pytestmark = pytest.mark.gpu


def test_config2_large_single_table():
    N, M, a = 200000, 20000, 0.7
    t = stb.Table(N, M, N, M, a, stb.S_STABLE)
    # spot values from the reference build
    assert harness.close(t.S(200000, 20000), 2070256.428328451).all()
    assert harness.close(t.S(200000, 2), 2241200.0617941185).all()
    assert harness.close(t.S(200000, 10000), 2162666.1833218271).all()
    # column prefix vs the oracle at M'=256
    Mp = 256
    S, _ = harness.oracle_tables(N, Mp, a, want_V=False)
    rng = np.random.default_rng(11)
    rows = np.unique(np.concatenate([rng.integers(2, N + 1, size=400), [N, N - 1, 257, 256, 255]]))
    for n in rows:
        g = t.rows(0, int(n), 1)[0, :Mp]
        top = min(int(n), Mp)
        assert harness.close(g[:top], S[n - 1, :top]).all(), n
    # identities on cells the prefix does not reach
    nn = rng.integers(3, N + 1, size=200000).astype(np.uint32)
    mm = np.minimum(rng.integers(2, M + 1, size=200000), nn - 1).astype(np.uint32)
    s0 = t.S_batch(nn, mm)
    assert np.isfinite(s0).all()
    # monotone in n for fixed m: S^{n+1}_m > S^n_m  (U > 1 for n >= 2)
    k = nn < N
    s1 = t.S_batch(nn[k] + 1, mm[k])
    assert (s1 > s0[k]).all()
    for n in (5, 100, 19999, 20000):
        assert t.S(n, n) == 0.0
        assert harness.close(t.S(n, n - 1), math.log(n * (n - 1) * (1 - a) / 2)).all()
    for n in (2, 1000, 200000):
        assert harness.close(t.S(n, 1), math.lgamma(n - a) - math.lgamma(1 - a)).all()
    t.free()


def _golden_cells(count, seed, N, M):
    """the seeded (n, m) of tests/golden/make_golden_config2.py"""
    rng = np.random.default_rng(seed)
    n = rng.integers(2, N + 1, size=count).astype(np.uint32)
    m = (2 + (rng.random(count) * (np.minimum(n, M) - 1)).astype(np.uint32)).astype(np.uint32)
    return n, np.minimum(m, np.minimum(n, M)).astype(np.uint32)


def test_config2_against_reference_golden():
    """SURVEY.md 8d C2 / VERDICT r1 #5: the full config-2 table against values the UNMODIFIED reference produced
    at that size (tests/golden/config2_S.npz: last row, last column, 10^6 seeded cells of log S, read with the
    reference's S_S) and, for V, the oracle's restatement (config2_O.npz; the reference's own V-only table of this
    size crashes in S_V).  Bar: 1e-12 relative (abs floor 1)."""
    import os

    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    gs = np.load(os.path.join(gdir, "config2_S.npz"))
    N, M, a, seed = int(gs["N"]), int(gs["M"]), float(gs["a"]), int(gs["seed"])
    assert (N, M, a) == (200000, 20000, 0.7)
    have_v = os.path.exists(os.path.join(gdir, "config2_O.npz"))
    t = stb.Table(N, M, N, M, a, stb.S_STABLE | (stb.S_UVTABLE if have_v else 0) | stb.S_NOMIRROR)
    mm = np.arange(1, M + 1, dtype=np.uint32)
    nn = np.arange(M, N + 1, dtype=np.uint32)
    worst = {}
    for name, got, want in (
        ("S last row", t.S_batch(np.full(M, N, dtype=np.uint32), mm), gs["S_lastrow"]),
        ("S last column", t.S_batch(nn, np.full(nn.shape[0], M, dtype=np.uint32)), gs["S_lastcol"]),
        ("S cells", t.S_batch(*_golden_cells(gs["S_cells"].shape[0], seed, N, M)), gs["S_cells"]),
    ):
        assert harness.close(got, want).all(), name
        worst[name] = harness.max_err(got[np.isfinite(want)], want[np.isfinite(want)])
    if have_v:
        gv = np.load(os.path.join(gdir, "config2_O.npz"))
        for name, got, want in (
            ("V last row", t.V_batch(np.full(M, N, dtype=np.uint32), mm), gv["V_lastrow"]),
            ("V last column", t.V_batch(nn, np.full(nn.shape[0], M, dtype=np.uint32)), gv["V_lastcol"]),
            ("V cells", t.V_batch(*_golden_cells(gv["V_cells"].shape[0], seed + 1, N, M)), gv["V_cells"]),
        ):
            assert harness.close(got, want).all(), name
            worst[name] = harness.max_err(got, want)
    print("config 2 vs golden, max relative error:", worst)
    t.free()


def test_config3_shape_with_V():
    N, M, a = 50000, 5000, 0.7
    t = stb.Table(N, M, N, M, a, stb.S_STABLE | stb.S_UVTABLE)
    assert harness.close(t.S(50000, 5000), 455174.29062118637).all()
    rng = np.random.default_rng(5)
    nn = rng.integers(3, N, size=50000).astype(np.uint32)
    mm = np.minimum(rng.integers(2, M + 1, size=50000), nn - 1).astype(np.uint32)
    s0, s1 = t.S_batch(nn, mm), t.S_batch(nn + 1, mm)
    v = t.V_batch(nn, mm)
    # U^n_m = n - m a + 1/V^n_m = S^{n+1}_m / S^n_m    (lib/stable.c:875-883)
    u = nn - mm * a + 1.0 / v
    assert np.allclose(np.log(u), s1 - s0, rtol=0, atol=2e-9)
    # V^n_m = S^n_m / S^n_{m-1}
    sm = t.S_batch(nn, mm - 1)
    k = mm >= 3
    assert np.allclose(np.log(v[k]), (s0 - sm)[k], rtol=0, atol=2e-9)
    t.free()


def test_table_wider_than_one_launch():
    """M = 36 000 > 148 x 224 columns: two passes over the columns on a real B200.  Its first
    20 000 columns are the columns of a single-pass M = 20 000 table (a table's columns do not
    depend on M) bit for bit; across the pass boundary and in the last column the stored values
    satisfy the recurrence S^n_m = logadd(log(n-1-m a) + S^{n-1}_m, S^{n-1}_{m-1}) to 1e-12."""
    N, Mw, Ms, a = 37000, 36000, 20000, 0.55
    wide = stb.Table(N, Mw, N, Mw, a, stb.S_STABLE | stb.S_NOMIRROR)
    ref = stb.Table(N, Ms, N, Ms, a, stb.S_STABLE | stb.S_NOMIRROR)
    for n in (2, 777, 20000, 20001, 36999, 37000):
        rw = wide.rows(0, n, 1)[0]
        rr = ref.rows(0, n, 1)[0]
        top = min(n, Ms)
        assert np.array_equal(rw[:top], rr[:top]), n
    last = wide.rows(0, 36990, 11)
    assert np.isfinite(last[:, :36000]).all()
    assert wide.S(36000, 36000) == 0.0 and wide.S(37000, 36000) > 0
    r0, r1 = wide.rows(0, 36999, 1)[0], wide.rows(0, 37000, 1)[0]
    m = np.array([30000, 33152, 33153, 33154, 35990, 36000])  # 33152 = 148 x 224: the pass boundary
    lhs = r1[m - 1]
    rhs = np.logaddexp(np.log(36999 - m * a) + r0[m - 1], r0[m - 2])
    assert np.allclose(lhs, rhs, rtol=1e-12, atol=0)
    wide.free()
    ref.free()


def test_config2_recurrence_holds_across_every_cta_boundary():
    """Full-size config 2, S and V: two million stored cells -- half of them on the two columns either
    side of EVERY strip boundary of the launch (multiples of 160 columns), the rest uniform -- must
    satisfy the defining recurrences from the stored row above, to 1e-12:
        S^n_m = logadd(log(n-1-m a) + S^{n-1}_m, S^{n-1}_{m-1})              (lib/stable.c:380-388)
        V^n_m = (1 + (n-1-m a) V^{n-1}_m) / (1/V^{n-1}_{m-1} + (n-1-(m-1)a))  (lib/stable.c:475-482)
    A hand-off error between CTAs (boundary ring, exponents, batch phase) would break them."""
    N, M, a = 200000, 20000, 0.7
    t = stb.Table(N, M, N, M, a, stb.S_STABLE | stb.S_UVTABLE | stb.S_NOMIRROR)
    rng = np.random.default_rng(5)
    cnt = 500_000
    edges = np.arange(160, M, 160)
    m_edge = rng.choice(np.concatenate([edges, edges + 1]), size=cnt)
    m_any = rng.integers(3, M + 1, size=cnt)
    for m in (m_edge, m_any):
        m = m.astype(np.int64)
        n = np.maximum(m + 2, rng.integers(4, N + 1, size=cnt)).astype(np.int64)  # n-1 > m: all three cells off the diagonal's right
        n = np.minimum(n, N)
        ok = n - 1 > m
        n, m = n[ok], m[ok]
        u = lambda x: x.astype(np.uint32)
        s_nm, s_up, s_left = t.S_batch(u(n), u(m)), t.S_batch(u(n - 1), u(m)), t.S_batch(u(n - 1), u(m - 1))
        rhs = np.logaddexp(np.log(n - 1 - m * a) + s_up, s_left)
        assert np.all(np.abs(s_nm - rhs) <= 1e-12 * np.maximum(1.0, np.abs(rhs))), np.max(np.abs(s_nm - rhs) / np.maximum(1.0, np.abs(rhs)))
        v_nm, v_up, v_left = t.V_batch(u(n), u(m)), t.V_batch(u(n - 1), u(m)), t.V_batch(u(n - 1), u(m - 1))
        k = m > 2  # V^n_2 has its own form (no column 1 in the V table)
        rhs_v = (1 + (n - 1 - m * a) * v_up) / (1 / v_left + (n - 1 - (m - 1) * a))
        assert np.all(np.abs(v_nm[k] - rhs_v[k]) <= 1e-12 * np.abs(rhs_v[k])), np.max(np.abs(v_nm[k] - rhs_v[k]) / np.abs(rhs_v[k]))
    t.free()
