"""One-off GPU fuzz of the table fill against the CPU oracle (development aid; the permanent tests
are in tests/): random extents, discounts, storage flags, launch geometries (STB_STRIP_K,
STB_STRIP_SLOTS -> several passes).  usage: python tools/fuzz_gpu.py [cases] [seed]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import libstb_b200 as stb  # noqa: E402
from tests import harness  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
S, V, F = stb.S_STABLE, stb.S_UVTABLE, stb.S_FLOAT
bad = 0
for c in range(cases):
    M = int(rng.choice([rng.integers(1, 40), rng.integers(40, 400), rng.integers(400, 1400)]))
    N = M + int(rng.choice([0, 1, rng.integers(2, 60), rng.integers(60, 3000)]))
    N, M = max(N, 2), max(M, 1)
    a = float(rng.choice([0.0, 0.01, 0.98, rng.uniform(0.01, 0.98)]))
    fl = int(rng.choice([S, V, S | V, S | V | F, S | F]))
    for k in ("STB_STRIP_K", "STB_STRIP_SLOTS"):
        os.environ.pop(k, None)
    geo = rng.integers(0, 4)
    if geo == 1:
        os.environ["STB_STRIP_K"] = str(rng.choice([1, 3, 5, 7]))
    elif geo == 2:
        os.environ["STB_STRIP_SLOTS"] = str(rng.integers(1, 4))
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
        bad += 1
        print("MISMATCH", N, M, a, fl, dict((k, os.environ.get(k)) for k in ("STB_STRIP_K", "STB_STRIP_SLOTS")), flush=True)
print(f"fuzz: {cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
