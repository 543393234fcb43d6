"""The CPU oracle against the committed golden vectors (tests/golden/tables.json, generated
from the reference by tests/golden/make_golden.py) and against mathematical known answers
the reference satisfies.  Runs anywhere, no GPU, no /root/reference."""
import hashlib
import json
import math
import os

import numpy as np
import pytest

from tests import harness

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tables.json")))
FLAGS = harness.S_STABLE | harness.S_UVTABLE


def _f(s):
    return None if s is None else float(s)


@pytest.mark.parametrize("ent", GOLD["tables"], ids=lambda e: f"N{e['N']}_M{e['M']}_a{e['a']}")
def test_tables_hash_and_spots(ent):
    N, M, a = ent["N"], ent["M"], ent["a"]
    S, V = harness.oracle_tables(N, M, a)
    hS, hV = hashlib.sha256(), hashlib.sha256()
    for n in range(2, N + 1):
        top = min(n, M)
        hS.update(S[n - 1, :top].astype("<f8").tobytes())
        if top >= 2:
            hV.update(V[n - 1, 1:top].astype("<f8").tobytes())
    assert hS.hexdigest() == ent["sha256_S"]
    assert hV.hexdigest() == ent["sha256_V"]
    O = harness.oracle()
    t = O.orc_make(N, M, N, M, a, FLAGS)
    try:
        for sp in ent["spots"]:
            n, m = sp["n"], sp["m"]
            assert O.orc_S(t, n, m) == _f(sp["S"])
            if sp["V"] is not None:
                assert O.orc_V(t, n, m) == _f(sp["V"])
            assert O.orc_U(t, n, m) == _f(sp["U"])
            assert O.orc_UV(t, n, m) == _f(sp["UV"])
    finally:
        O.orc_free(t)


def test_survey_spot_values():
    """Spot values recorded in SURVEY.md 8c / BASELINE.md from the reference build."""
    O = harness.oracle()
    t = O.orc_make(10000, 1000, 10000, 1000, 0.5, FLAGS)
    try:
        assert O.orc_S(t, 10000, 1000) == 76855.523932738957
        assert O.orc_S(t, 10000, 2) == 82095.23314599508
        assert O.orc_V(t, 9998, 998) == 0.0019008494057812999
    finally:
        O.orc_free(t)
    t = O.orc_make(10, 10, 10, 10, 0.0, FLAGS)
    try:
        assert O.orc_S(t, 10, 3) == 13.974819340449155
    finally:
        O.orc_free(t)


def test_asymptote_and_edges_golden():
    O = harness.oracle()
    for e in GOLD["asympt"]:
        t = O.orc_make(20, 10, 20, 10, e["a"], FLAGS | harness.S_ASYMPT)
        try:
            assert O.orc_S(t, e["n"], e["m"]) == _f(e["S"])
            assert O.orc_V(t, e["n"], e["m"]) == _f(e["V"])
            assert O.orc_asympt(e["a"], e["n"], e["m"]) == _f(e["asympt"])
        finally:
            O.orc_free(t)
    g = GOLD["edge"]
    t = O.orc_make(g["N"], g["M"], g["N"], g["M"], g["a"], FLAGS)
    try:
        for c in g["cells"]:
            n, m = c["n"], c["m"]
            assert O.orc_S(t, n, m) == _f(c["S"]), (n, m)
            if c["V"] is not None:
                assert O.orc_V(t, n, m) == _f(c["V"]), (n, m)
            if c["UV"] is not None:
                assert O.orc_UV(t, n, m) == _f(c["UV"]), (n, m)
            if c["U"] is not None:
                assert O.orc_U(t, n, m) == _f(c["U"]), (n, m)
    finally:
        O.orc_free(t)


def _stirling1_unsigned(nmax):
    """Exact unsigned Stirling numbers of the first kind: c(n+1,k) = n c(n,k) + c(n,k-1)."""
    c = [[0] * (nmax + 2) for _ in range(nmax + 2)]
    c[0][0] = 1
    for n in range(nmax + 1):
        for k in range(1, n + 2):
            c[n + 1][k] = n * c[n][k] + c[n][k - 1]
    return c


def test_known_answers():
    """a=0 gives unsigned Stirling numbers of the first kind; S^n_n=1; S^n_{n-1}=n(n-1)(1-a)/2;
    S_U*S_V==S_UV; S_S(n,1)=lgamma(n-a)-lgamma(1-a)."""
    N = 20
    c = _stirling1_unsigned(N)
    S, _ = harness.oracle_tables(N, N, 0.0, want_V=False)
    for n in range(2, N + 1):
        for m in range(1, n + 1):
            assert math.isclose(math.exp(S[n - 1, m - 1]), c[n][m], rel_tol=1e-12), (n, m)
    assert round(math.exp(S[9, 2])) == 1172700
    for a in (0.1, 0.5, 0.9):
        S, V = harness.oracle_tables(60, 60, a)
        for n in range(3, 61):
            assert S[n - 1, n - 1] == 0.0
            assert math.isclose(math.exp(S[n - 1, n - 2]), n * (n - 1) * (1 - a) / 2, rel_tol=1e-12)
            assert math.isclose(S[n - 1, 0], math.lgamma(n - a) - math.lgamma(1 - a), rel_tol=1e-13, abs_tol=1e-13)
        O = harness.oracle()
        t = O.orc_make(60, 60, 60, 60, a, FLAGS)
        try:
            for n, m in [(30, 7), (59, 20), (40, 39)]:
                assert math.isclose(O.orc_U(t, n, m) * O.orc_V(t, n, m), O.orc_UV(t, n, m), rel_tol=1e-13)
                # U^n_m = S^{n+1}_m / S^n_m
                assert math.isclose(O.orc_U(t, n, m), math.exp(O.orc_S(t, n + 1, m) - O.orc_S(t, n, m)), rel_tol=1e-12)
        finally:
            O.orc_free(t)


def test_cell_counts():
    O = harness.oracle()
    assert O.orc_cells_S(10000, 1000) == 9489501
    assert O.orc_cells_V(10000, 1000) == 9490500
    assert O.orc_cells_S(200000, 20000) == 3799790001
    assert O.orc_cells_S(50000, 5000) == 237447501
