"""arms / arms_simple (include/arms.h, libstb_b200/csrc/ars.c) against the reference's ARMS
(lib/arms.c, compiled unmodified into oracle/_ref/libstb_ref.so): the same log-density callback is
handed to both libraries, both draw their uniforms from glibc rand() after the same srand(seed),
and the draws, the number of density evaluations and the return codes must be IDENTICAL -- the
envelope arithmetic is restated operation for operation (on a sorted array instead of the
reference's linked pool).  Also the sampler switch: sampleb in ARS mode against the reference's
default build (which is the ARS configuration)."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import libstb_b200 as stb
from tests import harness

needs_ref = pytest.mark.skipif(not os.path.exists(harness.REF_SO), reason="reference build not present")
libc = C.CDLL(None)
libc.srand.argtypes = [C.c_uint]
d, dp, ip = C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int)
POST = C.CFUNCTYPE(d, d, C.c_void_p)


def _ref():
    R = C.CDLL(harness.REF_SO)
    R.arms_simple.restype = C.c_int
    R.arms_simple.argtypes = [C.c_int, dp, dp, POST, C.c_void_p, C.c_int, dp, dp]
    R.arms.restype = C.c_int
    R.arms.argtypes = [dp, C.c_int, dp, dp, POST, C.c_void_p, dp, C.c_int, C.c_int, dp, dp, C.c_int, dp, dp, C.c_int, ip]
    u32p = C.POINTER(C.c_uint32)
    R.sampleb.restype, R.sampleb.argtypes = d, [d, C.c_int, d, d, u32p, u32p, d, C.c_void_p, C.c_int, C.c_int]
    return R


DENSITIES = {
    "normal": (lambda x: -0.5 * (x - 0.3) ** 2 / 0.04, -2.0, 3.0),
    "gamma": (lambda x: 3.5 * math.log(x) - 2.0 * x, 0.01, 20.0),
    "beta": (lambda x: 2.0 * math.log(x) + 5.0 * math.log1p(-x), 0.001, 0.999),
    "steep": (lambda x: -40.0 * abs(x - 1.0) ** 1.5, 0.0, 4.0),
    "flat": (lambda x: -1e-3 * x * x, -1.0, 1.0),
}


def _simple(L, f, xl, xr, seed, n, metro=0, xprev=None):
    cb = POST(lambda x, _: f(x))
    a, b = C.c_double(xl), C.c_double(xr)
    prev = C.c_double((xl + xr) / 2 if xprev is None else xprev)
    out = C.c_double(0.0)
    libc.srand(seed)
    res = []
    for _ in range(n):
        rc = L.arms_simple(3, C.byref(a), C.byref(b), cb, None, metro, C.byref(prev), C.byref(out))
        res.append((rc, out.value))
        if metro:
            prev.value = out.value
    return res


@needs_ref
@pytest.mark.parametrize("name", sorted(DENSITIES))
def test_arms_simple_bit_identical(name):
    f, xl, xr = DENSITIES[name]
    for seed in (1, 7, 12345):
        mine = _simple(stb.lib(), f, xl, xr, seed, 200)
        ref = _simple(_ref(), f, xl, xr, seed, 200)
        assert mine == ref
        assert all(rc == 0 and xl <= x <= xr for rc, x in mine)


@needs_ref
def test_arms_metropolis_nonconcave_bit_identical():
    """a bimodal log-density: plain ARS reports the violation (2000), ARMS samples it"""
    f = lambda x: math.log(0.6 * math.exp(-0.5 * (x + 1.5) ** 2 / 0.3) + 0.4 * math.exp(-0.5 * (x - 1.0) ** 2 / 0.2))
    mine = _simple(stb.lib(), f, -4.0, 4.0, 99, 300, metro=1, xprev=0.1)
    ref = _simple(_ref(), f, -4.0, 4.0, 99, 300, metro=1, xprev=0.1)
    assert mine == ref and all(rc == 0 for rc, _ in mine)
    assert _simple(stb.lib(), f, -4.0, 4.0, 5, 20) == _simple(_ref(), f, -4.0, 4.0, 5, 20)


def _full(L, f, xinit, xl, xr, seed, nsamp, npoint=50, convex=1.0, metro=0, xprev=0.0, cent=(5.0, 50.0, 95.0)):
    cb = POST(lambda x, _: f(x))
    xi = (d * len(xinit))(*xinit)
    a, b, cv, prev = C.c_double(xl), C.c_double(xr), C.c_double(convex), C.c_double(xprev)
    xs = (d * nsamp)()
    qc = (d * len(cent))(*cent)
    xc = (d * len(cent))()
    nev = C.c_int(-1)
    libc.srand(seed)
    rc = L.arms(xi, len(xinit), C.byref(a), C.byref(b), cb, None, C.byref(cv), npoint, metro, C.byref(prev), xs, nsamp,
                qc, xc, len(cent), C.byref(nev))
    return rc, list(xs), list(xc), nev.value


@needs_ref
def test_arms_many_samples_centiles_and_full_envelope():
    """one envelope, many draws: the envelope fills up (npoint reached: further points are ignored),
    centiles of the final envelope, the evaluation count"""
    f, xl, xr = DENSITIES["gamma"]
    for npoint in (9, 20, 50):
        mine = _full(stb.lib(), f, [0.5, 2.0, 6.0, 12.0], xl, xr, 2024, 60, npoint=npoint)
        ref = _full(_ref(), f, [0.5, 2.0, 6.0, 12.0], xl, xr, 2024, 60, npoint=npoint)
        assert mine == ref
        assert mine[0] in (0, 2001)


@needs_ref
def test_arms_error_codes():
    f, xl, xr = DENSITIES["normal"]
    L, R = stb.lib(), _ref()
    cases = [
        dict(xinit=[0.0, 1.0], ),                       # 1001
        dict(xinit=[-1.0, 0.0, 1.0], npoint=6),         # 1002
        dict(xinit=[-2.0, 0.0, 1.0]),                   # 1003 (on the bound)
        dict(xinit=[0.0, 0.0, 1.0]),                    # 1004
        dict(xinit=[-1.0, 0.0, 1.0], cent=(101.0,)),    # 1005
        dict(xinit=[-1.0, 0.0, 1.0], convex=-1.0),      # 1008
        dict(xinit=[-1.0, 0.0, 1.0], metro=1, xprev=9.0),  # 1007
    ]
    want = [1001, 1002, 1003, 1004, 1005, 1008, 1007]
    for kw, code in zip(cases, want):
        xinit = kw.pop("xinit")
        assert _full(L, f, xinit, xl, xr, 1, 1, **kw)[0] == code
        assert _full(R, f, xinit, xl, xr, 1, 1, **kw)[0] == code


@needs_ref
def test_sampleb_ars_mode_matches_reference_default_build():
    """sampleb with STB_SAMPLER_ARS against the reference's default (ARS) build: the auxiliary
    beta draws come from drand48/lrand48, the ARS uniforms from rand(); same seeds, same draws.
    (Host code in both: bterms is lgamma sums.)"""
    L, R = stb.lib(), _ref()
    libc.srand48.argtypes = [C.c_long]
    rng = np.random.default_rng(3)
    I = 40
    N = rng.integers(5, 400, size=I).astype(np.uint32)
    T = np.minimum(N, rng.integers(1, 60, size=I)).astype(np.uint32)
    u32p = C.POINTER(C.c_uint32)
    old = L.stb_set_sampler(1)
    try:
        for apar in (0.3, 0.7):
            b_m, b_r = 5.0, 5.0
            for it in range(20):
                libc.srand48(100 + it)
                libc.srand(200 + it)
                b_m = L.sampleb(b_m, I, 1.1, 20.0, N.ctypes.data_as(u32p), T.ctypes.data_as(u32p), apar, None, 1, 0)
                libc.srand48(100 + it)
                libc.srand(200 + it)
                b_r = R.sampleb(b_r, I, 1.1, 20.0, N.ctypes.data_as(u32p), T.ctypes.data_as(u32p), apar, None, 1, 0)
                assert b_m == b_r, (apar, it, b_m, b_r)
    finally:
        L.stb_set_sampler(old)


def test_arms_simple_distribution():
    """no reference needed: draws from a N(0.3, 0.2^2) log-density have its mean and spread"""
    f, xl, xr = DENSITIES["normal"]
    xs = np.array([x for rc, x in _simple(stb.lib(), f, xl, xr, 42, 4000) if rc == 0])
    assert len(xs) == 4000
    assert abs(xs.mean() - 0.3) < 0.02 and abs(xs.std() - 0.2) < 0.02


def test_rand31_is_glibc_rand():
    """stb_rand31_t (csrc/rand31.h) reproduces srand()/rand() bit for bit: the per-chain streams of
    the batched ARS samplers are the streams the reference's global generator would produce"""
    L = stb.lib()
    libc.rand.restype = C.c_int
    for seed in (0, 1, 2, 777, 12345, 2**31 - 1, 2**32 - 5):
        g = stb.rand31_states([seed])
        libc.srand(seed)
        for _ in range(2000):
            assert L.stb_rand31_next(g.ctypes.data) == libc.rand()


@needs_ref
def test_lockstep_ars_is_the_scalar_sampler_chain_by_chain():
    """stb_arms_simple_batch: the lock-step driver of the batched ARS samplers, here with a host
    density.  Chain c on the stream of srand(seed_c) must make exactly the draw the reference's
    arms_simple makes after srand(seed_c) -- including chains whose log-density is not concave on
    their interval: the sampler stops with code 2000 and the value stays what it was (the
    reference's samplea / sampleb ignore the return value, lib/samplea.c:210-215)."""
    L, R = stb.lib(), _ref()
    bimodal = lambda x: math.log(0.6 * math.exp(-0.5 * (x + 1.5) ** 2 / 0.3) + 0.4 * math.exp(-0.5 * (x - 1.0) ** 2 / 0.2))
    for name, f in (("gamma", DENSITIES["gamma"][0]), ("bimodal", bimodal)):
        Cn = 64
        rng = np.random.default_rng(5)
        if name == "gamma":
            lo = rng.uniform(0.01, 1.0, Cn)
            hi = lo + rng.uniform(2.0, 15.0, Cn)
        else:
            lo = rng.uniform(-4.0, -2.0, Cn)
            hi = rng.uniform(-1.0, 4.0, Cn)  # some intervals see one mode (concave), some both
        seeds = [3000 + c for c in range(Cn)]
        x = np.full(Cn, -77.0)
        cb = POST(lambda v, _: f(v))
        nfailed = C.c_size_t(0)
        rnd = stb.rand31_states(seeds)
        rc = L.stb_arms_simple_batch(x.ctypes.data_as(dp), Cn, lo.ctypes.data_as(dp), hi.ctypes.data_as(dp),
                                     rnd.ctypes.data, cb, None, C.byref(nfailed))
        assert rc == 0
        fails = 0
        for c in range(Cn):
            a, b, prev, out = C.c_double(lo[c]), C.c_double(hi[c]), C.c_double(0.0), C.c_double(-77.0)
            libc.srand(seeds[c])
            code = R.arms_simple(3, C.byref(a), C.byref(b), cb, None, 0, C.byref(prev), C.byref(out))
            assert x[c] == out.value, (name, c, code)
            libc.rand.restype = C.c_int
            assert L.stb_rand31_next(rnd.ctypes.data + c * stb.RAND31_DTYPE.itemsize) == libc.rand(), (name, c)
            fails += code != 0
        assert nfailed.value == fails
        if name == "bimodal":
            assert fails > 0
