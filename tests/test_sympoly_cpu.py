"""sympoly / sympoly_sample (include/sympoly.h, csrc/sympoly.c) against the reference's
lib/sympoly.c compiled unmodified (oracle/_ref/libstb_ref.so): the same elementary symmetric polynomials bit
for bit (same recursion, same scaling), the same subsets from the same drand48 stream, and the polynomials
against their definition (sums of products over subsets) on small cases."""
import ctypes as C
import itertools
import math
import os

import numpy as np
import pytest

import libstb_b200 as stb
from tests import harness

needs_ref = pytest.mark.skipif(not harness.have_ref(), reason="reference build not present")
dp = C.POINTER(C.c_double)


def _bind(L):
    L.sympoly.restype, L.sympoly.argtypes = C.c_int, [C.c_int, C.c_int, dp, dp, dp]
    L.sympoly_sample.restype, L.sympoly_sample.argtypes = C.c_uint32, [C.c_int, C.c_int, dp, C.c_void_p]
    return L


def _sympoly(L, val, BK):
    K = len(val)
    v = np.array(val, dtype=np.float64)
    res = np.full(K + 3, np.nan)
    ov = C.c_double(0)
    assert L.sympoly(K, BK, v.ctypes.data_as(dp), res.ctypes.data_as(dp), C.byref(ov)) == 0
    return res, ov.value


def _cases():
    rng = np.random.default_rng(3)
    out = []
    for K in (1, 2, 3, 5, 9, 17, 32):
        for scale in (0.3, 1.0, 4.0, 200.0):
            out.append(rng.gamma(1.0, scale, size=K))
    out.append(np.array([3.0, 0.5, 7.0, 0.25, 1.0, 1e6, 2.0, 9.0, 1e5, 30.0, 4e4, 11.0]))  # log overflow > 15: kept
    return out


def test_sympoly_is_the_definition():
    L = _bind(stb.lib())
    for val in ([0.5, 2.0, 3.0], [0.1, 0.2, 0.3, 0.4, 0.9], [4.0, 5.0, 0.5, 6.0, 1.5, 0.2]):
        K = len(val)
        res, ov = _sympoly(L, val, K)
        for h in range(1, K + 1):
            want = sum(math.prod(c) for c in itertools.combinations(val, h))
            assert res[h] * math.exp(ov) == pytest.approx(want, rel=1e-13)
        assert res[0] == 1.0


@needs_ref
def test_sympoly_equals_the_reference_bit_for_bit():
    L, R = _bind(stb.lib()), _bind(C.CDLL(harness.REF_SO))
    for val in _cases():
        K = len(val)
        for BK in sorted({1, 2, max(1, K // 2), K}):
            (a, oa), (b, ob) = _sympoly(L, val, BK), _sympoly(R, val, BK)
            top = min(BK, K)
            assert oa == ob
            assert np.array_equal(a[: top + 1], b[: top + 1]), (K, BK)
    big = _sympoly(L, _cases()[-1], 12)
    assert big[1] > 15  # the scaled form was kept


@needs_ref
def test_sympoly_sample_equals_the_reference_draw_for_draw():
    L, R = _bind(stb.lib()), _bind(C.CDLL(harness.REF_SO))
    libc = C.CDLL(None)
    libc.srand48.argtypes = [C.c_long]
    libc.drand48.restype = C.c_double
    rng = np.random.default_rng(8)
    for K, H in ((1, 1), (4, 1), (4, 4), (5, 2), (8, 3), (12, 7), (20, 5), (31, 30), (32, 16)):
        for scale in (0.4, 3.0):
            val = rng.gamma(1.0, scale, size=K)
            got, want = [], []
            for lib, out in ((L, got), (R, want)):
                libc.srand48(1000 + K * 37 + H)
                for _ in range(200):
                    out.append(lib.sympoly_sample(K, H, val.ctypes.data_as(dp), None))
                out.append(libc.drand48())  # where the stream stands afterwards
            assert got == want, (K, H, scale)
            assert all(bin(int(w)).count("1") == H for w in got[:-1])
    assert L.sympoly_sample(3, 5, val.ctypes.data_as(dp), None) == 0 and L.sympoly_sample(3, 0, val.ctypes.data_as(dp), None) == 0


def test_sympoly_sample_frequencies():
    """H = 2 of K = 4: P(subset) = product of its values / e_2"""
    L = _bind(stb.lib())
    libc = C.CDLL(None)
    libc.srand48.argtypes = [C.c_long]
    libc.srand48(5)
    val = np.array([0.5, 2.0, 1.0, 3.0])
    n = 40000
    counts = {}
    for _ in range(n):
        w = L.sympoly_sample(4, 2, val.ctypes.data_as(dp), None)
        counts[w] = counts.get(w, 0) + 1
    e2 = sum(a * b for a, b in itertools.combinations(val, 2))
    for (i, a), (j, b) in itertools.combinations(enumerate(val), 2):
        p = a * b / e2
        assert counts.get((1 << i) | (1 << j), 0) / n == pytest.approx(p, abs=4 * math.sqrt(p * (1 - p) / n))
