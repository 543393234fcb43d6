"""the reference's precision study (test/precision_test.c:322-357) on the engine's tables.
The program fills, at N = 10001, M = 4000, a = 0.5, the log table and the ratio table in double and in float and
prints, along row n = 10000 for m = 10, 20, ..., the ratio S^n_m / S^n_{m-1} four ways: exp of a difference of two
double logs, the same from float logs, the double ratio table, the float ratio table -- its point (README, "ratio of
Stirling numbers"): a float LOG table is useless for ratios (a difference of two floats of size ~1e4 keeps no
digits), a float RATIO table is as good as a float gets.  The same four numbers here, from S_STABLE / S_UVTABLE
tables with and without S_FLOAT, with the double ratio table as the truth."""
import numpy as np
import pytest

import libstb_b200 as stb

pytestmark = pytest.mark.gpu


def test_float_vs_double_log_and_ratio_tables():
    N, M, a = 10001, 4000, 0.5
    n = N - 1
    mm = np.arange(10, M, 10, dtype=np.uint32)
    nn = np.full(mm.shape[0], n, dtype=np.uint32)
    out = {}
    for name, flags in (("f64", 0), ("f32", stb.S_FLOAT)):
        t = stb.Table(N, M, N, M, a, stb.S_STABLE | stb.S_UVTABLE | flags)
        out["S" + name] = np.exp(t.S_batch(nn, mm) - t.S_batch(nn, mm - 1))
        out["V" + name] = t.V_batch(nn, mm)
        t.free()
    truth = out["Vf64"]
    err = {k: float(np.max(np.abs(v - truth) / truth)) for k, v in out.items() if k != "Vf64"}
    print("max relative error of S^n_m/S^n_(m-1), n = 10000, m = 10..3990, against the FP64 ratio table:", err)
    assert err["Sf64"] < 1e-9   # two FP64 logs of size ~1e4..1e5: ~1e-16 * 1e5 absolute in the exponent
    assert err["Vf32"] < 2e-7   # a float ratio table is good to a float ulp
    assert err["Sf32"] > 1e-4   # a float log table is not (the study's point): digits lost in the difference
    assert err["Sf32"] > 100 * err["Vf32"]
