"""Host side of the samplers against glibc and the compiled reference (slice-sampler build,
oracle/_ref/libstb_ref_slice.so).  No GPU: random streams, distributions, special functions,
SliceSimple and the scalar sampleb are plain C host code in both libraries.

Bars: the 48-bit streams and everything drawn from them are BIT-IDENTICAL to the reference (the
ziggurat tables are constructed, not copied, and still reproduce the published constants);
digamma / trigamma / digammaInv agree to 1e-12 relative (different algorithms: series+recurrence
here, Amos 610 there); SliceSimple on a shared density is bit-identical; scalar sampleb draws
agree to 1e-9 (their densities differ by the digamma warm-up's last bits)."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import libstb_b200 as stb
from tests import harness

libc = C.CDLL(None)
libc.drand48.restype = C.c_double
libc.lrand48.restype = C.c_long
libc.srand48.argtypes = [C.c_long]

needs_ref = pytest.mark.skipif(not os.path.exists(harness.REF_SLICE_SO), reason="reference build not present")


def _ref():
    R = C.CDLL(harness.REF_SLICE_SO)
    d = C.c_double
    for name in ("gsl_rng_gaussian_ziggurat", "gsl_rng_gamma", "digammaRN", "MLdigamma", "MLtrigamma", "digammaInv"):
        f = getattr(R, name)
        f.restype, f.argtypes = d, [d]
    R.gsl_rng_beta.restype, R.gsl_rng_beta.argtypes = d, [d, d]
    POST = C.CFUNCTYPE(d, d, C.c_void_p)
    R.SliceSimple.restype = C.c_int
    R.SliceSimple.argtypes = [C.POINTER(d), POST, C.POINTER(d), C.c_void_p, C.c_int, C.c_void_p]
    u32p = C.POINTER(C.c_uint32)
    R.sampleb.restype, R.sampleb.argtypes = d, [d, C.c_int, d, d, u32p, u32p, d, C.c_void_p, C.c_int, C.c_int]
    return R


@pytest.mark.parametrize("seed", [0, 1, 12345, 2**31 - 1, -7])
def test_rng48_is_glibc_drand48(seed):
    L = stb.lib()
    st = C.c_uint64(L.stb_rng48_state(seed))
    libc.srand48(seed)
    for _ in range(2000):
        assert L.stb_rng48_drand(C.byref(st)) == libc.drand48()
    for _ in range(2000):
        assert L.stb_rng48_lrand48(C.byref(st)) == libc.lrand48()


@needs_ref
def test_distributions_bit_identical_to_reference():
    """Same global stream, same draws: Gaussian ziggurat, gamma (both branches), beta."""
    L, R = stb.lib(), _ref()
    n = 60000
    libc.srand48(99)
    ref = [R.gsl_rng_gaussian_ziggurat(1.0) for _ in range(n)]
    ref_state = libc.drand48()
    libc.srand48(99)
    our = [L.gsl_rng_gaussian_ziggurat(1.0) for _ in range(n)]
    assert our == ref and libc.drand48() == ref_state
    assert abs(np.mean(our)) < 0.02 and abs(np.std(our) - 1) < 0.02 and max(map(abs, our)) > 3.5  # tail reached
    for a in (0.3, 0.999, 1.0, 2.5, 10.0, 777.0):
        libc.srand48(5)
        ref = [R.gsl_rng_gamma(a) for _ in range(5000)]
        libc.srand48(5)
        our = [L.gsl_rng_gamma(a) for _ in range(5000)]
        assert our == ref, a
    libc.srand48(11)
    ref = [R.gsl_rng_beta(10.0, float(k % 50 + 1)) for k in range(5000)]
    libc.srand48(11)
    our = [L.gsl_rng_beta(10.0, float(k % 50 + 1)) for k in range(5000)]
    assert our == ref
    # the per-chain form draws the same numbers from its own state
    st = C.c_uint64(L.stb_rng48_state(11))
    assert [L.stb_rng48_beta(C.byref(st), 10.0, float(k % 50 + 1)) for k in range(300)] == ref[:300]


@needs_ref
def test_special_functions_vs_reference():
    L, R = stb.lib(), _ref()
    xs = np.concatenate([np.linspace(0.01, 12, 400), np.logspace(1, 7, 200)])
    for x in xs:
        x = float(x)
        assert L.digammaRN(x) == pytest.approx(R.digammaRN(x), rel=2e-16, abs=2e-16)
        d_ref = R.MLdigamma(x)
        assert abs(L.MLdigamma(x) - d_ref) <= 1e-12 * max(1.0, abs(d_ref)), x
        assert L.MLtrigamma(x) == pytest.approx(R.MLtrigamma(x), rel=1e-12)
    for y in np.linspace(-8, 12, 200):
        assert L.digammaInv(float(y)) == pytest.approx(R.digammaInv(float(y)), rel=1e-10)
    # known values
    assert L.MLdigamma(1.0) == pytest.approx(-0.5772156649015329, rel=1e-14)
    assert L.MLtrigamma(1.0) == pytest.approx(math.pi**2 / 6, rel=1e-14)


@needs_ref
def test_slice_simple_bit_identical():
    L, R = stb.lib(), _ref()
    POST = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)
    calls = []

    def post(x, _):  # log density of gamma(shape 3, rate 0.5): unimodal
        calls.append(x)
        return 2 * math.log(x) - 0.5 * x

    cb = POST(post)
    out = {}
    for name, lib_ in (("ref", R), ("our", L)):
        calls.clear()
        libc.srand48(2024)
        x = C.c_double(4.0)
        bounds = (C.c_double * 2)(0.01, 60.0)
        rc = lib_.SliceSimple(C.byref(x), cb, bounds, None, 25, None)
        out[name] = (rc, x.value, list(calls), libc.drand48())
    assert out["ref"] == out["our"]
    # error convention: start outside the bounds -> 1 (lib/sslice.c:40-45)
    x = C.c_double(100.0)
    assert L.SliceSimple(C.byref(x), cb, (C.c_double * 2)(0.01, 60.0), None, 1, None) == 1


@needs_ref
@pytest.mark.parametrize("apar,I,tmax", [(0.0, 3, 40), (0.0, 60, 40), (0.3, 40, 30), (0.7, 200, 80)])
def test_scalar_sampleb_matches_reference(apar, I, tmax):
    """lib/sampleb.c:79-159: a==0 closed form (gamma and Gaussian branches), a>0 warm-up + slice sampler."""
    L, R = stb.lib(), _ref()
    rng = np.random.default_rng(I)
    T = rng.integers(1, tmax, size=I).astype(np.uint32)
    N = (T + rng.integers(0, 500, size=I)).astype(np.uint32)
    N[0] = 0  # skipped restaurant (lib/sampleb.c:91-92)
    u32p = C.POINTER(C.c_uint32)
    b_ref = b_our = 10.0
    for step in range(6):
        libc.srand48(1000 + step)
        b_ref = R.sampleb(b_ref, I, 1.1, 20.0, N.ctypes.data_as(u32p), T.ctypes.data_as(u32p), apar, None, 2, 0)
        s_ref = libc.drand48()
        libc.srand48(1000 + step)
        b_our = L.sampleb(b_our, I, 1.1, 20.0, N.ctypes.data_as(u32p), T.ctypes.data_as(u32p), apar, None, 2, 0)
        assert libc.drand48() == s_ref, "the two samplers consumed different numbers of draws"
        assert b_our == pytest.approx(b_ref, rel=1e-9), step
        b_our = b_ref  # keep the chains together
