"""BASELINE config 5: the Dirichlet-process limit a = 0 (unsigned Stirling numbers of the first
kind), growth on look-up from N = 10 000 to 500 000 in >= x1.1 steps (lib/stable.c:564-815) and
the closed-form asymptote past maxN (lib/stable.c:1057-1084, :905-911)."""
import math

import numpy as np
import pytest

import libstb_b200 as stb
from tests import harness

pytestmark = pytest.mark.gpu


def _stirling1_unsigned(nmax):
    """c(n, m): c(n+1, m) = n c(n, m) + c(n, m-1), exact integers."""
    c = [[0] * (nmax + 2) for _ in range(nmax + 2)]
    c[0][0] = 1
    for n in range(nmax + 1):
        for m in range(1, n + 2):
            c[n + 1][m] = n * c[n][m] + c[n][m - 1]
    return c


def test_a0_exact_stirling_numbers():
    t = stb.Table(30, 30, 30, 30, 0.0, stb.S_STABLE | stb.S_UVTABLE)
    c = _stirling1_unsigned(30)
    assert math.exp(t.S(10, 3)) == pytest.approx(1172700, rel=1e-13)
    for n in range(2, 21):
        for m in range(1, n + 1):
            assert math.exp(t.S(n, m)) == pytest.approx(c[n][m], rel=1e-12), (n, m)
            if m >= 2:
                assert t.V(n, m) == pytest.approx(c[n][m] / c[n][m - 1], rel=1e-12), (n, m)
    t.free()


def test_growth_to_500k_and_asymptote():
    M, maxN = 2000, 500000
    flags = stb.S_STABLE | stb.S_UVTABLE | stb.S_ASYMPT | stb.S_NOMIRROR
    t = stb.Table(10000, M, maxN, M, 0.0, flags)
    O = harness.oracle()
    Mp = 48  # column prefix the oracle can follow to N = 500 000 in a second
    S, V = harness.oracle_tables(maxN, Mp, 0.0)
    n, extends, used = 10000, 0, t.usedN
    while n < maxN:
        n = min(maxN, int(n * 1.1) + 1)
        got = t.S(n, M)  # grows the table when n is past the filled extent
        assert math.isfinite(got)
        if t.usedN != used:
            extends += 1
            assert t.usedN >= min(maxN, max(int(used * 1.1), used + 50))  # growth policy
            used = t.usedN
        for m in (2, 7, Mp):
            assert harness.close(t.S(n, m), S[n - 1, m - 1]).all(), (n, m)
            assert harness.close(t.V(n - 2, m), V[n - 3, m - 1]).all(), (n, m)
        assert harness.close(t.S(n, 1), math.lgamma(n)).all()  # S^n_1 = (n-1)! at a = 0
    assert t.usedN == maxN and 20 <= extends <= 45
    # monotone in m near the mode and U = S(n+1,m)/S(n,m)
    n0 = 400000
    u = t.U(n0, 1000)
    assert math.log(u) == pytest.approx(t.S(n0 + 1, 1000) - t.S(n0, 1000), abs=5e-9)
    # past maxN: the asymptote, identical to the reference's closed form
    for nn in (600000, 1000000, 10000000):
        for m in (2, 50, 2000):
            assert t.S(nn, m) == O.orc_asympt(0.0, nn, m)
            assert t.V(nn, m) == O.orc_V_asympt(0.0, nn, m)
    # without S_ASYMPT: log 0 / 0 past the bound
    t2 = stb.Table(100, 20, 200, 20, 0.0, stb.S_STABLE | stb.S_UVTABLE)
    assert t2.S(201, 5) == -math.inf and t2.V(201, 5) == 0.0
    t.free(), t2.free()
