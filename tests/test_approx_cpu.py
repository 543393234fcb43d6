"""S_approx / S_approx_da (include/sapprox.h), gammadiff / psidiff / g-p-q caches (include/lgamma.h)
and the polygamma functions against the reference's own outputs: the committed golden vectors
(tests/golden/approx.json, generated from oracle/_ref/libstb_ref_slice.so by
tests/golden/make_golden_approx.py) and, when the compiled reference is present, live calls.

These are host C closed forms in both libraries (same glibc lgamma/log/exp), so everything that
does not pass through digamma / trigamma / tetragamma is BIT-IDENTICAL; where the polygamma
functions enter (series + recurrence here, Amos 610 in the reference) the bar is 1e-11 relative
with an absolute floor of 1e-13 (differences of nearly equal digammas)."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

import libstb_b200 as stb
from tests import harness

def dec(v):
    """floats travel as repr strings (exact round trip, and inf / nan survive JSON)"""
    if isinstance(v, str) and v not in "gpq":
        return float(v)
    return v


GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "approx.json")))
for _k, _rows in GOLD.items():
    if isinstance(_rows, list):
        GOLD[_k] = [[dec(v) for v in row] for row in _rows]


class GCache(C.Structure):
    _fields_ = [("par", C.c_double), ("lgpar", C.c_double), ("cache", C.c_double * 100)]


def same(x, y):
    return x == y or (math.isnan(x) and math.isnan(y))


def near(x, y, rel=1e-11, floor=1e-13):
    if math.isnan(x) or math.isnan(y) or math.isinf(x) or math.isinf(y):
        return same(x, y)
    return abs(x - y) <= rel * abs(y) + floor


def test_S_approx_golden():
    L = stb.lib()
    bad = []
    for n, m, a, want in GOLD["S_approx"]:
        want = dec(want)
        got = L.S_approx(n, m, a)
        # a >= 0.001: lgamma/log/exp only -> bit-identical; below: polygamma differences
        ok = near(got, want) if a < 0.001 else same(got, want)
        if not ok:
            bad.append((n, m, a, got, want))
    assert not bad, bad[:5]


def test_S_approx_da_golden():
    L = stb.lib()
    bad = []
    for n, m, a, want in GOLD["S_approx_da"]:
        want = dec(want)
        got = L.S_approx_da(n, m, a)
        # The sum of weighted digamma differences cancels to O(a^(m-1)) and is then divided by
        # a^(m-1): the last bits of digamma (different algorithms) are amplified by 1/a^(m-1).  At
        # a = 0.0005 that leaves ~7 digits in EITHER library; elsewhere the bar is 1e-9.
        rel = 1e-6 if a < 0.001 else 1e-9
        if not near(got, want, rel=rel, floor=1e-12):
            bad.append((n, m, a, got, want))
    assert not bad, bad[:5]


def test_S_approx_conventions_and_table():
    """0 on the diagonal, -inf above it and for m > 4 (lib/sapprox.c:35-38,76); and the closed form
    agrees with the recurrence table to the accuracy its float arguments and alternating sum allow (1e-4)."""
    L = stb.lib()
    assert L.S_approx(5, 5, 0.3) == 0.0 and L.S_approx_da(5, 5, 0.3) == 0.0
    assert L.S_approx(3, 4, 0.3) == -math.inf and L.S_approx(10, 5, 0.3) == -math.inf
    # (m a < 1 here: past a pole of Gamma(1 - j a) the alternating sum's signs are wrong in the
    # reference's formula and it returns NaN / garbage -- reproduced, see the golden vectors)
    S, _ = harness.oracle_tables(60, 4, 0.2)
    for n in (5, 17, 60):
        for m in (1, 2, 3, 4):
            assert abs(L.S_approx(n, m, 0.2) - S[n - 1, m - 1]) <= 1e-4 * max(1.0, abs(S[n - 1, m - 1]))


def test_gammadiff_psidiff_golden():
    L = stb.lib()
    bad = []
    for n, al, lga, want in GOLD["gammadiff"]:
        got = L.gammadiff(n, al, lga)
        ok = same(got, dec(want)) if (n <= 3 or al > 0.5 or n >= 2000 and lga != 0) else near(got, dec(want))
        if not ok:
            bad.append(("g", n, al, lga, got, dec(want)))
    for n, al, pa, want in GOLD["psidiff"]:
        got = L.psidiff(n, al, pa)
        if not near(got, dec(want)):
            bad.append(("p", n, al, pa, got, dec(want)))
    assert not bad, bad[:5]


def test_psidiff_large_N_is_a_digamma_difference():
    """Deviation (documented in csrc/approx.c): for N >= 2000, alpha <= 1/2, pa == 0 the reference
    adds lgamma(N+alpha) where digamma is meant (lib/lgamma.c:224-227)."""
    L = stb.lib()
    for al in (0.05, 0.2, 0.5):
        want = L.MLdigamma(2500 + al) - L.MLdigamma(al)
        assert abs(L.psidiff(2500, al, 0.0) - want) <= 1e-3  # (third-order expansion about 3: ~alpha^4 accuracy)


def test_caches_golden():
    L = stb.lib()
    state = {}
    bad = []
    for kind, p, j, want in GOLD["cache"]:
        key = (kind, p)
        if key not in state:
            state[key] = GCache()
            getattr(L, kind + "cache_init")(C.byref(state[key]), p)
        got = getattr(L, kind + "cache_value")(C.byref(state[key]), j)
        ok = same(got, dec(want)) if kind == "g" else near(got, dec(want))
        if not ok:
            bad.append((kind, p, j, got, dec(want)))
    assert not bad, bad[:5]


def test_polygamma_golden():
    L = stb.lib()
    for x, d0, d1, d2, d3 in GOLD["polygamma"]:
        for got, want in ((L.MLdigamma(x), d0), (L.MLtrigamma(x), d1), (L.MLtetragamma(x), d2), (L.MLpentagamma(x), d3)):
            assert abs(got - dec(want)) <= 1e-12 * abs(dec(want)) + 1e-15, (x, got, want)
        for k, want in enumerate((d0, d1, d2, d3)):
            assert L.MLpsigamma(x, float(k)) == (L.MLdigamma, L.MLtrigamma, L.MLtetragamma, L.MLpentagamma)[k](x)
    # orders beyond 3 (the reference's Amos 610 code serves any order, lib/polygamma.c:502-523): Hurwitz zeta form
    from scipy.special import polygamma

    for n in (4, 5, 7, 10):
        for x in (0.3, 1.0, 2.5, 17.0, 123.4):
            want = float(polygamma(n, x))
            assert abs(L.MLpsigamma(x, float(n)) - want) <= 1e-12 * abs(want), (n, x)
    if os.path.exists(harness.REF_SLICE_SO):  # and the reference itself (its polygamma.c is in the slice build)
        R = C.CDLL(harness.REF_SLICE_SO)
        R.MLpsigamma.restype, R.MLpsigamma.argtypes = C.c_double, [C.c_double, C.c_double]
        for n in (4, 6):
            for x in (0.7, 2.5, 40.0):
                want = R.MLpsigamma(x, float(n))
                assert abs(L.MLpsigamma(x, float(n)) - want) <= 1e-10 * abs(want), (n, x, want)
    assert math.isnan(L.MLpsigamma(1.0, -1.0))


@pytest.mark.skipif(not os.path.exists(harness.REF_SLICE_SO), reason="reference build not present")
def test_against_live_reference_random():
    from tests.golden.make_golden_approx import declare

    R = declare(C.CDLL(harness.REF_SLICE_SO))
    L = stb.lib()
    rng = np.random.default_rng(7)
    for _ in range(400):
        n = int(rng.integers(2, 5000))
        m = int(rng.integers(1, 5))
        a = float(np.float32(rng.uniform(0.002, 0.98)))
        if any(abs(j * a - round(j * a)) < 1e-3 for j in range(1, 5)):
            continue  # at a pole of lgamma(1 - j a) both libraries return garbage
        assert same(L.S_approx(n, m, a), R.S_approx(n, m, a)), (n, m, a)
        x, y = L.S_approx_da(n, m, a), R.S_approx_da(n, m, a)
        assert near(x, y, rel=1e-8, floor=1e-10), (n, m, a, x, y)
    for _ in range(400):
        n = int(rng.integers(0, 3000))
        al = float(rng.uniform(0.01, 2.0))
        lga = math.lgamma(al) if rng.integers(2) else 0.0
        assert near(L.gammadiff(n, al, lga), R.gammadiff(n, al, lga)), (n, al, lga)
