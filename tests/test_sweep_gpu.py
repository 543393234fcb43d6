"""Discount sweep (stb_sweep_*: many tables per launch, streamed through resident slabs) against the
CPU oracle and against the single-table path.  Scaling by powers of two is exact, so a table filled
inside a multi-table launch must equal the same table filled alone BIT FOR BIT."""
import math

import numpy as np
import pytest

import libstb_b200 as stb
from tests import harness

pytestmark = pytest.mark.gpu


def _pairs(rng, N, M, count):
    n = rng.integers(1, N + 1, size=count).astype(np.uint32)
    m = np.minimum(rng.integers(1, M + 1, size=count), n).astype(np.uint32)
    # the conventions: n == m -> 0, m == 0 / n < m / out of range -> -inf
    n[:4] = [5, 7, N + 1, 3]
    m[:4] = [5, 0, 2, 9]
    return n, m


@pytest.mark.parametrize("N,M,na", [(600, 80, 37), (2000, 167, 300)])
def test_sweep_matches_oracle_and_single_table(N, M, na):
    rng = np.random.default_rng(N + na)
    a = np.concatenate([[0.0, 0.003, 0.5, 0.98, 0.9995], rng.uniform(0.01, 0.99, size=na - 5)])
    n, m = _pairs(rng, N, M, 5000)
    w = stb.Sweep(N, M)
    w.set_pairs(n, m)
    g, s, last = w.run(a, gather=True, sums=True, lastrow=True)
    assert w.tables_in_flight >= 1
    ok_pair = (n >= m) & (m >= 1) & (n <= N)
    for j in list(range(6)) + [na // 2, na - 1]:
        S, _ = harness.oracle_tables(N, M, float(a[j]), want_V=False)
        ref = np.where(n == m, 0.0, np.where(ok_pair, S[np.minimum(n, N) - 1, np.maximum(m, 1) - 1], -math.inf))
        assert harness.close(g[j], ref).all(), (j, a[j])
        assert harness.close(last[j], S[N - 1, :M]).all()
        fin = np.isfinite(ref)
        assert abs(g[j][fin].sum() - ref[fin].sum()) <= 1e-12 * abs(ref[fin].sum())
        # same table filled alone: bit-identical
        t = stb.Table(N, M, N, M, float(a[j]), stb.S_STABLE)
        alone = t.S_batch(n, m)
        assert np.array_equal(alone, g[j]), (j, a[j])
        t.free()
    # sums: the fixed-order tree over ALL pairs (including the -inf conventions)
    assert np.all(np.isneginf(s))
    n2, m2 = n[4:], m[4:]
    w.set_pairs(n2, m2)
    g2, s2, _ = w.run(a[:9], gather=True, sums=True)
    assert np.allclose(s2, g2.sum(axis=1), rtol=1e-13, atol=0)
    w.free()


def test_sweep_config3_shape():
    """BASELINE config 3 shape (N=50 000, M=5 000) for a handful of discounts: golden spot, identities."""
    N, M = 50000, 5000
    a = np.array([0.7, (0 + 0.5) / 4096, (4095 + 0.5) / 4096, 0.25, 0.5, 0.9, 0.33, 0.61, 0.05])
    rng = np.random.default_rng(3)
    n = rng.integers(3, N + 1, size=100000).astype(np.uint32)
    m = np.minimum(rng.integers(2, M + 1, size=100000), n - 1).astype(np.uint32)
    w = stb.Sweep(N, M)
    w.set_pairs(n, m)
    g, s, last = w.run(a, gather=True, sums=True, lastrow=True)
    assert harness.close(last[0, M - 1], 455174.29062118637).all()
    assert np.isfinite(g).all() and np.isfinite(s).all()
    for j, aj in enumerate(a):
        # S(N,1) = lgamma(N-a) - lgamma(1-a)   (lib/stable.c:822-873)
        assert harness.close(last[j, 0], math.lgamma(N - aj) - math.lgamma(1 - aj)).all()
    # column prefix against the oracle for two of the discounts
    for j in (0, 2):
        S, _ = harness.oracle_tables(N, 64, float(a[j]), want_V=False)
        assert harness.close(last[j, :64], S[N - 1, :64]).all()
    w.free()
