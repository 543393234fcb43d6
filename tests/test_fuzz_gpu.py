"""Fuzz of the table fill against the CPU oracle: seeded random extents, discounts, storage flags
and launch geometries (STB_STRIP_K; STB_STRIP_SLOTS -> several passes over the columns), every
stored cell compared.  STB_FUZZ_CASES / STB_FUZZ_SEED widen it (400 cases ran clean on a B200)."""
import os

import numpy as np
import pytest

import libstb_b200 as stb
from tests import harness

pytestmark = pytest.mark.gpu


def test_fuzz_against_oracle(monkeypatch):
    cases = int(os.environ.get("STB_FUZZ_CASES", "40"))
    rng = np.random.default_rng(int(os.environ.get("STB_FUZZ_SEED", "1")))
    S, V, F = stb.S_STABLE, stb.S_UVTABLE, stb.S_FLOAT
    bad = []
    for _ in range(cases):
        M = int(rng.choice([rng.integers(1, 40), rng.integers(40, 400), rng.integers(400, 1400)]))
        N = M + int(rng.choice([0, 1, rng.integers(2, 60), rng.integers(60, 3000)]))
        N, M = max(N, 2), max(M, 1)
        a = float(rng.choice([0.0, 0.01, 0.98, rng.uniform(0.01, 0.98)]))
        fl = int(rng.choice([S, V, S | V, S | V | F, S | F]))
        monkeypatch.delenv("STB_STRIP_K", raising=False)
        monkeypatch.delenv("STB_STRIP_SLOTS", raising=False)
        geo = int(rng.integers(0, 4))
        if geo == 1:
            monkeypatch.setenv("STB_STRIP_K", str(rng.choice([1, 3, 5, 7])))
        elif geo == 2:
            monkeypatch.setenv("STB_STRIP_SLOTS", str(rng.integers(1, 4)))
        t = stb.Table(N, M, N, M, a, fl)
        Nu, Mu = t.usedN, t.usedM
        So, Vo = harness.oracle_tables(Nu, Mu, a)
        rel = 1.2e-7 if fl & F else 1e-12
        ok = True
        if fl & S:
            g = t.rows(0, 1, Nu)[:, :Mu]
            m = harness.valid_mask(Nu, Mu)
            ref = So[m].astype(np.float32).astype(np.float64) if fl & F else So[m]
            ok &= bool(harness.close(g[m], ref, rel).all())
        if fl & V:
            g = t.rows(1, 1, Nu)[:, :Mu]
            m = harness.valid_mask(Nu, Mu, for_V=True)
            ref = Vo[m].astype(np.float32).astype(np.float64) if fl & F else Vo[m]
            ok &= bool(harness.close(g[m], ref, rel).all())
        t.free()
        if not ok:
            bad.append((N, M, a, fl, os.environ.get("STB_STRIP_K"), os.environ.get("STB_STRIP_SLOTS")))
    assert not bad, bad[:5]
