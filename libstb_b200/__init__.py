"""libstb_b200 -- B200-native generalised Stirling-number engine (stable.h drop-in).

The product is the C-ABI shared library ``libstb_b200/lib/libstb_b200.so`` (C host layer +
hand-written sm_100a CUDA kernels).  This Python module is only the ctypes binding the tests
and the benchmark use to call that C ABI -- there is no Python or CPU implementation of the
table fill here, and importing it without the built library fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# STB_B200_LIB: another build of the same library (development: profiling builds); never a fallback
LIB_PATH = os.environ.get("STB_B200_LIB") or os.path.join(_HERE, "lib", "libstb_b200.so")

# flag bits, include/stable.h
S_STABLE, S_UVTABLE, S_FLOAT, S_VERBOSE, S_QUITONBOUND, S_THREADS, S_ASYMPT = 1, 2, 4, 8, 16, 32, 64
S_MIRROR_ORDER, S_NOMIRROR = 1 << 16, 1 << 17
STB_SAMPLER_SLICE, STB_SAMPLER_ARS = 0, 1          # include/psample.h
STB_PARTITION_REFERENCE, STB_PARTITION_EXACT = 0, 1  # include/stb_b200.h

_lib = None


def lib() -> C.CDLL:
    """Load the C-ABI library and declare every prototype of include/*.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    u, d, vp, u32p, dp = C.c_uint, C.c_double, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_double)
    L.S_make.restype, L.S_make.argtypes = vp, [u, u, u, u, d, C.c_uint32]
    L.S_tag.restype, L.S_tag.argtypes = None, [vp, C.c_char_p]
    L.S_remake.restype, L.S_remake.argtypes = C.c_int, [vp, d]
    L.S_free.restype, L.S_free.argtypes = None, [vp]
    for name in ("S_S", "S_U", "S_UV", "S_V", "S_asympt"):
        f = getattr(L, name)
        f.restype, f.argtypes = d, [vp, u, u]
    L.S_S1.restype, L.S_S1.argtypes = d, [vp, u]
    L.S_report.restype, L.S_report.argtypes = None, [vp, vp]
    for name in ("stb_S_batch", "stb_V_batch", "stb_U_batch", "stb_UV_batch"):
        f = getattr(L, name)
        f.restype, f.argtypes = C.c_int, [vp, u32p, u32p, dp, C.c_size_t]
    for name in ("stb_S_batch_device", "stb_V_batch_device", "stb_U_batch_device", "stb_UV_batch_device"):
        f = getattr(L, name)
        f.restype, f.argtypes = C.c_int, [vp, vp, vp, vp, C.c_size_t]
    L.stb_extend.restype, L.stb_extend.argtypes = C.c_int, [vp, u, u]
    L.stb_read_rows.restype, L.stb_read_rows.argtypes = C.c_int, [vp, C.c_int, u, u, dp]
    L.stb_sweep_create.restype, L.stb_sweep_create.argtypes = vp, [u, u, C.c_uint32]
    L.stb_sweep_set_pairs.restype, L.stb_sweep_set_pairs.argtypes = C.c_int, [vp, u32p, u32p, C.c_size_t]
    L.stb_sweep_run.restype, L.stb_sweep_run.argtypes = C.c_int, [vp, dp, C.c_size_t, dp, dp, dp]
    L.stb_sweep_last_fill_ms.restype, L.stb_sweep_last_fill_ms.argtypes = d, [vp]
    L.stb_sweep_tables_in_flight.restype, L.stb_sweep_tables_in_flight.argtypes = C.c_int, [vp]
    L.stb_sweep_tables_per_launch.restype, L.stb_sweep_tables_per_launch.argtypes = C.c_int, [vp]
    L.stb_sweep_free.restype, L.stb_sweep_free.argtypes = None, [vp]
    L.stb_sweep_multi_create.restype = vp
    L.stb_sweep_multi_create.argtypes = [C.POINTER(C.c_int), C.c_int, u, u, C.c_uint32]
    L.stb_sweep_multi_set_pairs.restype, L.stb_sweep_multi_set_pairs.argtypes = C.c_int, [vp, u32p, u32p, C.c_size_t]
    L.stb_sweep_multi_run.restype, L.stb_sweep_multi_run.argtypes = C.c_int, [vp, dp, C.c_size_t, dp, dp, dp]
    L.stb_sweep_multi_last_fill_ms.restype, L.stb_sweep_multi_last_fill_ms.argtypes = d, [vp]
    L.stb_sweep_multi_device_ms.restype, L.stb_sweep_multi_device_ms.argtypes = d, [vp, C.c_int]
    L.stb_sweep_multi_devices.restype, L.stb_sweep_multi_devices.argtypes = C.c_int, [vp]
    L.stb_sweep_multi_free.restype, L.stb_sweep_multi_free.argtypes = None, [vp]
    L.stb_release_caches.restype, L.stb_release_caches.argtypes = None, []
    # samplers (include/psample.h, srng.h, digamma.h)
    u64p, ip = C.POINTER(C.c_uint64), C.POINTER(C.c_int)
    POST = C.CFUNCTYPE(d, d, vp)
    L.SliceSimple.restype, L.SliceSimple.argtypes = C.c_int, [dp, POST, dp, vp, C.c_int, vp]
    L.sampleb.restype, L.sampleb.argtypes = d, [d, C.c_int, d, d, u32p, u32p, d, vp, C.c_int, C.c_int]
    L.samplea.restype = d
    L.samplea.argtypes = [d, C.c_int, ip, u32p, C.POINTER(u32p), C.POINTER(C.POINTER(C.c_uint16)), vp, dp, vp,
                          C.c_int, C.c_int]
    L.samplea2.restype = d
    L.samplea2.argtypes = [d, vp, C.c_int, ip, u32p, C.POINTER(u32p), C.POINTER(C.POINTER(C.c_uint16)), vp, dp, vp,
                           C.c_int, C.c_int]
    L.logminus.restype, L.logminus.argtypes = d, [d, d]
    L.stb_ti_gibbs.restype = C.c_int
    L.stb_ti_gibbs.argtypes = [vp, d, C.c_size_t, u32p, u32p, C.POINTER(C.c_float), C.c_uint32, u32p,
                               C.POINTER(C.c_uint16), u32p, C.POINTER(C.c_uint64), C.c_int, C.c_int]
    L.stb_last_gibbs_ms.restype, L.stb_last_gibbs_ms.argtypes = d, [vp]
    L.stb_partition_sample.restype = C.c_int
    L.stb_partition_sample.argtypes = [vp, d, u32p, C.POINTER(C.c_uint16), dp, u32p, C.c_size_t,
                                       C.POINTER(C.c_uint16), C.c_size_t, C.c_int]
    L.stb_set_partition_mode.restype, L.stb_set_partition_mode.argtypes = C.c_int, [C.c_int]
    for name in ("gsl_rng_gaussian_ziggurat", "gsl_rng_gamma", "digammaRN", "MLdigamma", "MLtrigamma", "digammaInv",
                 "MLtetragamma", "MLpentagamma"):
        f = getattr(L, name)
        f.restype, f.argtypes = d, [d]
    L.MLpsigamma.restype, L.MLpsigamma.argtypes = d, [d, d]
    # adaptive rejection sampling (include/arms.h) and the sampler switch (include/psample.h)
    L.arms_simple.restype = C.c_int
    L.arms_simple.argtypes = [C.c_int, dp, dp, POST, vp, C.c_int, dp, dp]
    L.arms.restype = C.c_int
    L.arms.argtypes = [dp, C.c_int, dp, dp, POST, vp, dp, C.c_int, C.c_int, dp, dp, C.c_int, dp, dp, C.c_int, ip]
    L.stb_set_sampler.restype, L.stb_set_sampler.argtypes = C.c_int, [C.c_int]
    # closed forms beside the tables (include/sapprox.h, lgamma.h)
    L.S_approx.restype, L.S_approx.argtypes = d, [C.c_int, C.c_int, C.c_float]
    L.S_approx_da.restype, L.S_approx_da.argtypes = d, [C.c_int, C.c_int, C.c_float]
    L.gammadiff.restype, L.gammadiff.argtypes = d, [C.c_int, d, d]
    L.psidiff.restype, L.psidiff.argtypes = d, [C.c_int, d, d]
    for kind in "gpq":
        getattr(L, kind + "cache_init").restype = None
        getattr(L, kind + "cache_init").argtypes = [vp, d]
        getattr(L, kind + "cache_value").restype = d
        getattr(L, kind + "cache_value").argtypes = [vp, C.c_int]
    L.gsl_rng_beta.restype, L.gsl_rng_beta.argtypes = d, [d, d]
    L.stb_rng48_state.restype, L.stb_rng48_state.argtypes = C.c_uint64, [C.c_long]
    L.stb_rng48_drand.restype, L.stb_rng48_drand.argtypes = d, [u64p]
    L.stb_rng48_lrand48.restype, L.stb_rng48_lrand48.argtypes = C.c_long, [u64p]
    L.stb_rng48_gaussian.restype, L.stb_rng48_gaussian.argtypes = d, [u64p, d]
    L.stb_rng48_gamma.restype, L.stb_rng48_gamma.argtypes = d, [u64p, d]
    L.stb_rng48_beta.restype, L.stb_rng48_beta.argtypes = d, [u64p, d, d]
    L.stb_rand31_seed.restype, L.stb_rand31_seed.argtypes = None, [vp, C.c_uint]
    L.stb_rand31_next.restype, L.stb_rand31_next.argtypes = C.c_int, [vp]
    L.stb_arms_simple_batch.restype = C.c_int
    L.stb_arms_simple_batch.argtypes = [dp, C.c_size_t, dp, dp, vp, POST, vp, C.POINTER(C.c_size_t)]
    L.stb_samplea_batch_ars.restype = C.c_int
    L.stb_samplea_batch_ars.argtypes = [dp, C.c_size_t, C.c_int, ip, u32p, C.POINTER(u32p),
                                        C.POINTER(C.POINTER(C.c_uint16)), dp, C.c_int, vp, vp]
    L.stb_sampleb_batch_ars.restype = C.c_int
    L.stb_sampleb_batch_ars.argtypes = [dp, C.c_size_t, C.c_int, d, d, u32p, u32p, dp, u64p, vp, vp]
    L.stb_samplea_batch.restype = C.c_int
    L.stb_samplea_batch.argtypes = [dp, C.c_size_t, C.c_int, ip, u32p, C.POINTER(u32p),
                                    C.POINTER(C.POINTER(C.c_uint16)), dp, C.c_int, u64p, C.c_int, vp]
    L.stb_samplea2_batch.restype = C.c_int
    L.stb_samplea2_batch.argtypes = L.stb_samplea_batch.argtypes
    L.stb_sampleb_batch.restype = C.c_int
    L.stb_sampleb_batch.argtypes = [dp, C.c_size_t, C.c_int, d, d, u32p, u32p, dp, u64p, C.c_int, vp]
    L.stb_samplea_batch_multi.restype = C.c_int
    L.stb_samplea_batch_multi.argtypes = [ip, C.c_int] + L.stb_samplea_batch.argtypes
    L.stb_sampleb_batch_multi.restype = C.c_int
    L.stb_sampleb_batch_multi.argtypes = [ip, C.c_int] + L.stb_sampleb_batch.argtypes
    L.stb_last_fill_ms.restype, L.stb_last_fill_ms.argtypes = d, [vp]
    L.stb_last_partition_ms.restype, L.stb_last_partition_ms.argtypes = d, [vp]
    L.stb_device_table.restype, L.stb_device_table.argtypes = vp, [vp, C.c_int, C.POINTER(C.c_size_t)]
    L.stb_device_count.restype, L.stb_device_count.argtypes = C.c_int, []
    L.stb_last_error.restype, L.stb_last_error.argtypes = C.c_char_p, []
    _lib = L
    return L


class _Header(C.Structure):
    """Leading public fields of stable_t (include/stable.h)."""

    _fields_ = [
        ("maxM", C.c_uint), ("maxN", C.c_uint), ("usedM", C.c_uint), ("usedN", C.c_uint),
        ("usedN1", C.c_uint), ("S1", C.POINTER(C.c_double)), ("lga", C.c_double), ("a", C.c_double),
        ("flags", C.c_uint32), ("memalloced", C.c_uint32),
    ]


class Table:
    """A stable_t handle.  Every method is one C-ABI call."""

    def __init__(self, initN, initM, maxN=None, maxM=None, a=0.5, flags=S_STABLE | S_UVTABLE):
        L = lib()
        self._L = L
        self.sp = L.S_make(initN, initM, maxN or initN, maxM or initM, a, flags)
        if not self.sp:
            raise RuntimeError("S_make failed: " + L.stb_last_error().decode())
        self.flags = flags

    # -- scalar API ------------------------------------------------------------------------
    def S(self, n, m): return self._L.S_S(self.sp, n, m)
    def V(self, n, m): return self._L.S_V(self.sp, n, m)
    def U(self, n, m): return self._L.S_U(self.sp, n, m)
    def UV(self, n, m): return self._L.S_UV(self.sp, n, m)
    def S1(self, n): return self._L.S_S1(self.sp, n)
    def asympt(self, n, m): return self._L.S_asympt(self.sp, n, m)

    def remake(self, a):
        if self._L.S_remake(self.sp, a):
            raise RuntimeError("S_remake failed: " + self._L.stb_last_error().decode())

    def extend(self, N, M):
        if self._L.stb_extend(self.sp, N, M):
            raise RuntimeError("stb_extend failed: " + self._L.stb_last_error().decode())

    @property
    def hdr(self): return _Header.from_address(self.sp)
    @property
    def usedN(self): return self.hdr.usedN
    @property
    def usedM(self): return self.hdr.usedM
    @property
    def last_fill_ms(self): return self._L.stb_last_fill_ms(self.sp)

    @property
    def last_partition_ms(self): return self._L.stb_last_partition_ms(self.sp)

    @property
    def ld(self):
        ld = C.c_size_t()
        self._L.stb_device_table(self.sp, 0, C.byref(ld))
        return ld.value

    # -- batched API -----------------------------------------------------------------------
    def _batch(self, fn, n, m):
        n = np.ascontiguousarray(n, dtype=np.uint32)
        m = np.ascontiguousarray(m, dtype=np.uint32)
        out = np.empty(n.shape[0], dtype=np.float64)
        u32p, dp = C.POINTER(C.c_uint32), C.POINTER(C.c_double)
        if fn(self.sp, n.ctypes.data_as(u32p), m.ctypes.data_as(u32p), out.ctypes.data_as(dp), n.shape[0]):
            raise RuntimeError("batch look-up failed: " + self._L.stb_last_error().decode())
        return out

    def S_batch(self, n, m): return self._batch(self._L.stb_S_batch, n, m)
    def V_batch(self, n, m): return self._batch(self._L.stb_V_batch, n, m)
    def U_batch(self, n, m): return self._batch(self._L.stb_U_batch, n, m)
    def UV_batch(self, n, m): return self._batch(self._L.stb_UV_batch, n, m)

    def partition_sample(self, a, n, t, logu, exact=False):
        """stb_partition_sample: table sizes of nodes (n[j], t[j]), 1 < t < n.  logu: one log-uniform per node
        (reference mode) or t[j]-1 per node laid out like the result (exact mode).  Returns (m, off):
        node j's sizes are m[off[j] : off[j] + t[j] - 1], entry M-1 drawn by round M."""
        n = np.ascontiguousarray(n, dtype=np.uint32)
        t = np.ascontiguousarray(t, dtype=np.uint16)
        logu = np.ascontiguousarray(logu, dtype=np.float64)
        sizes = t.astype(np.int64) - 1
        off = np.zeros(n.shape[0], dtype=np.uint32)
        if n.shape[0]:
            off[1:] = np.cumsum(sizes)[:-1]
        n_m = int(sizes.sum())
        if logu.shape[0] != (n_m if exact else n.shape[0]):
            raise ValueError("logu: one entry per node (reference mode) or per sampled size (exact mode)")
        m = np.zeros(max(n_m, 1), dtype=np.uint16)
        u32p, u16p, dp = C.POINTER(C.c_uint32), C.POINTER(C.c_uint16), C.POINTER(C.c_double)
        if self._L.stb_partition_sample(self.sp, float(a), n.ctypes.data_as(u32p), t.ctypes.data_as(u16p),
                                        logu.ctypes.data_as(dp), off.ctypes.data_as(u32p), n.shape[0],
                                        m.ctypes.data_as(u16p), n_m, int(bool(exact))):
            raise RuntimeError("stb_partition_sample failed: " + self._L.stb_last_error().decode())
        return m[:n_m], off

    def ti_gibbs(self, bpar, tok_off, tok_dish, H, n, t, T, rng, shared_stream=False, sweeps=1):
        """stb_ti_gibbs: table-indicator Gibbs sweeps over R restaurants on the device (test/demo.c:405-434).
        n, t: (R, D) counts / table counts, T: (R,) sums of t, tok_off: (R+1,) offsets into tok_dish,
        rng: 48-bit stream states, one per restaurant (one in all with shared_stream).  Returns new (t, T, rng)."""
        n = np.ascontiguousarray(n, dtype=np.uint32)
        t = np.array(t, dtype=np.uint16, order="C")
        T = np.array(T, dtype=np.uint32)
        rng = np.array(rng, dtype=np.uint64)
        tok_off = np.ascontiguousarray(tok_off, dtype=np.uint32)
        tok_dish = np.ascontiguousarray(tok_dish, dtype=np.uint32)
        H = np.ascontiguousarray(H, dtype=np.float32)
        R, D = n.shape
        if t.shape != (R, D) or T.shape != (R,) or tok_off.shape != (R + 1,) or H.shape != (D,):
            raise ValueError("ti_gibbs: shapes")
        if rng.shape[0] != (1 if shared_stream else R):
            raise ValueError("ti_gibbs: one stream per restaurant (one in all with shared_stream)")
        u32p = C.POINTER(C.c_uint32)
        if self._L.stb_ti_gibbs(self.sp, float(bpar), R, tok_off.ctypes.data_as(u32p), tok_dish.ctypes.data_as(u32p),
                                H.ctypes.data_as(C.POINTER(C.c_float)), D, n.ctypes.data_as(u32p),
                                t.ctypes.data_as(C.POINTER(C.c_uint16)), T.ctypes.data_as(u32p),
                                rng.ctypes.data_as(C.POINTER(C.c_uint64)), int(bool(shared_stream)), int(sweeps)):
            raise RuntimeError("stb_ti_gibbs failed: " + self._L.stb_last_error().decode())
        return t, T, rng

    @property
    def last_gibbs_ms(self):
        return self._L.stb_last_gibbs_ms(self.sp)

    def rows(self, which_V, n0, nrows):
        """Rows n0..n0+nrows-1 as an (nrows, ld) float64 array; column j holds m=j+1."""
        ld = self.ld
        out = np.empty((nrows, ld), dtype=np.float64)
        if self._L.stb_read_rows(self.sp, int(which_V), n0, nrows, out.ctypes.data_as(C.POINTER(C.c_double))):
            raise RuntimeError("stb_read_rows failed: " + self._L.stb_last_error().decode())
        return out

    def free(self):
        if self.sp:
            self._L.S_free(self.sp)
            self.sp = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Sweep:
    """A discount sweep context (include/stb_b200.h: stb_sweep_*).  Every method is one C-ABI call."""

    def __init__(self, N, M, flags=0):
        self._L = lib()
        self.N, self.M = N, M
        self.w = self._L.stb_sweep_create(N, M, flags)
        if not self.w:
            raise RuntimeError("stb_sweep_create failed: " + self._L.stb_last_error().decode())
        self.npairs = 0

    def set_pairs(self, n, m):
        n = np.ascontiguousarray(n, dtype=np.uint32)
        m = np.ascontiguousarray(m, dtype=np.uint32)
        u32p = C.POINTER(C.c_uint32)
        if self._L.stb_sweep_set_pairs(self.w, n.ctypes.data_as(u32p), m.ctypes.data_as(u32p), n.shape[0]):
            raise RuntimeError("stb_sweep_set_pairs failed: " + self._L.stb_last_error().decode())
        self.npairs = n.shape[0]

    def run(self, a, gather=True, sums=True, lastrow=False):
        a = np.ascontiguousarray(a, dtype=np.float64)
        dp = C.POINTER(C.c_double)
        g = np.empty((a.shape[0], self.npairs)) if gather else None
        s = np.empty(a.shape[0]) if sums else None
        r = np.empty((a.shape[0], self.M)) if lastrow else None
        ptr = lambda x: x.ctypes.data_as(dp) if x is not None else None
        if self._L.stb_sweep_run(self.w, a.ctypes.data_as(dp), a.shape[0], ptr(g), ptr(s), ptr(r)):
            raise RuntimeError("stb_sweep_run failed: " + self._L.stb_last_error().decode())
        return g, s, r

    @property
    def last_fill_ms(self): return self._L.stb_sweep_last_fill_ms(self.w)
    @property
    def tables_in_flight(self): return self._L.stb_sweep_tables_in_flight(self.w)

    @property
    def tables_per_launch(self): return self._L.stb_sweep_tables_per_launch(self.w)

    def free(self):
        if self.w:
            self._L.stb_sweep_free(self.w)
            self.w = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _device_list(devices):
    """(int*, ndev) for the multi-device entry points; None = every visible device"""
    if devices is None:
        return None, 0
    arr = (C.c_int * len(devices))(*[int(x) for x in devices])
    return arr, len(devices)


class SweepMulti(Sweep):
    """The sweep over several devices of this process (stb_sweep_multi_*): table j on devices[j % ndev]."""

    def __init__(self, N, M, devices=None, flags=0):
        self._L = lib()
        self.N, self.M = N, M
        arr, nd = _device_list(devices)
        self.w = self._L.stb_sweep_multi_create(arr, nd, N, M, flags)
        if not self.w:
            raise RuntimeError("stb_sweep_multi_create failed: " + self._L.stb_last_error().decode())
        self.npairs = 0

    def set_pairs(self, n, m):
        n = np.ascontiguousarray(n, dtype=np.uint32)
        m = np.ascontiguousarray(m, dtype=np.uint32)
        u32p = C.POINTER(C.c_uint32)
        if self._L.stb_sweep_multi_set_pairs(self.w, n.ctypes.data_as(u32p), m.ctypes.data_as(u32p), n.shape[0]):
            raise RuntimeError("stb_sweep_multi_set_pairs failed: " + self._L.stb_last_error().decode())
        self.npairs = n.shape[0]

    def run(self, a, gather=True, sums=True, lastrow=False):
        a = np.ascontiguousarray(a, dtype=np.float64)
        dp = C.POINTER(C.c_double)
        g = np.empty((a.shape[0], self.npairs)) if gather else None
        s = np.empty(a.shape[0]) if sums else None
        r = np.empty((a.shape[0], self.M)) if lastrow else None
        ptr = lambda x: x.ctypes.data_as(dp) if x is not None else None
        if self._L.stb_sweep_multi_run(self.w, a.ctypes.data_as(dp), a.shape[0], ptr(g), ptr(s), ptr(r)):
            raise RuntimeError("stb_sweep_multi_run failed: " + self._L.stb_last_error().decode())
        return g, s, r

    @property
    def last_fill_ms(self): return self._L.stb_sweep_multi_last_fill_ms(self.w)
    @property
    def ndev(self): return self._L.stb_sweep_multi_devices(self.w)
    @property
    def device_ms(self): return [self._L.stb_sweep_multi_device_ms(self.w, g) for g in range(self.ndev)]
    @property
    def tables_in_flight(self): raise AttributeError("per device: see Sweep.tables_in_flight")

    def free(self):
        if self.w:
            self._L.stb_sweep_multi_free(self.w)
            self.w = None


class SampleStats(C.Structure):
    """stb_sample_stats of include/psample.h."""

    _fields_ = [("evals", C.c_uint64), ("rounds", C.c_uint64), ("eval_ms", C.c_double),
                ("trace_x", C.POINTER(C.c_double)), ("trace_v", C.POINTER(C.c_double)),
                ("trace_n", C.POINTER(C.c_uint32)), ("trace_cap", C.c_uint32)]


class Counts:
    """Ragged count arrays n[i][k] (uint32), t[i][k] (uint16) laid out the way samplea takes them."""

    def __init__(self, n_rows, t_rows):
        self.I = len(n_rows)
        self.n_rows = [np.ascontiguousarray(r, dtype=np.uint32) for r in n_rows]
        self.t_rows = [np.ascontiguousarray(r, dtype=np.uint16) for r in t_rows]
        self.K = np.array([len(r) for r in self.n_rows], dtype=np.int32)
        self.T = np.array([int(r.sum()) for r in self.t_rows], dtype=np.uint32)
        self.N = np.array([int(r.sum()) for r in self.n_rows], dtype=np.uint32)
        u32p, u16p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint16)
        self.n_pp = (u32p * self.I)(*[r.ctypes.data_as(u32p) for r in self.n_rows])
        self.t_pp = (u16p * self.I)(*[r.ctypes.data_as(u16p) for r in self.t_rows])

    def args(self):
        ip, u32p = C.POINTER(C.c_int), C.POINTER(C.c_uint32)
        return self.I, self.K.ctypes.data_as(ip), self.T.ctypes.data_as(u32p), self.n_pp, self.t_pp


def _stats(C_chains, trace_cap):
    st = SampleStats()
    keep = None
    if trace_cap:
        tx = np.zeros((C_chains, trace_cap))
        tv = np.zeros((C_chains, trace_cap))
        tn = np.zeros(C_chains, dtype=np.uint32)
        st.trace_x, st.trace_v = tx.ctypes.data_as(C.POINTER(C.c_double)), tv.ctypes.data_as(C.POINTER(C.c_double))
        st.trace_n, st.trace_cap = tn.ctypes.data_as(C.POINTER(C.c_uint32)), trace_cap
        keep = (tx, tv, tn)
    return st, keep


def samplea_batch(a, counts, bpar, rng, loops=1, bpar_per_chain=False, trace_cap=0, devices=False):
    """stb_samplea_batch (devices=False) or stb_samplea_batch_multi (devices = a list of device ordinals, or
    None for every visible device): returns (a_new, rng_new, stats dict)."""
    L = lib()
    a = np.array(a, dtype=np.float64)
    rng = np.array(rng, dtype=np.uint64)
    bpar = np.ascontiguousarray(bpar, dtype=np.float64)
    st, keep = _stats(a.shape[0], trace_cap if devices is False else 0)
    dp, u64p = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    args = (a.ctypes.data_as(dp), a.shape[0], *counts.args(), bpar.ctypes.data_as(dp), int(bpar_per_chain),
            rng.ctypes.data_as(u64p), loops, C.byref(st))
    rc = L.stb_samplea_batch(*args) if devices is False else L.stb_samplea_batch_multi(*_device_list(devices), *args)
    if rc:
        raise RuntimeError(f"stb_samplea_batch failed ({rc}): " + L.stb_last_error().decode())
    return a, rng, {"evals": st.evals, "rounds": st.rounds, "eval_ms": st.eval_ms, "trace": keep}


def samplea2_batch(a, counts, bpar, rng, loops=1, bpar_per_chain=False, trace_cap=0):
    """stb_samplea2_batch: the table-free discount update (seat partitions against one table per chain, then one
    slice step) for C chains; returns (a_new, rng_new, stats dict)."""
    L = lib()
    a = np.array(a, dtype=np.float64)
    rng = np.array(rng, dtype=np.uint64)
    bpar = np.ascontiguousarray(bpar, dtype=np.float64)
    st, keep = _stats(a.shape[0], trace_cap)
    dp, u64p = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    rc = L.stb_samplea2_batch(a.ctypes.data_as(dp), a.shape[0], *counts.args(), bpar.ctypes.data_as(dp), int(bpar_per_chain),
                              rng.ctypes.data_as(u64p), loops, C.byref(st))
    if rc:
        raise RuntimeError(f"stb_samplea2_batch failed ({rc}): " + L.stb_last_error().decode())
    return a, rng, {"evals": st.evals, "rounds": st.rounds, "eval_ms": st.eval_ms, "trace": keep}


def sampleb_batch(b, counts, shape, scale, apar, rng, loops=1, trace_cap=0, devices=False):
    """stb_sampleb_batch / stb_sampleb_batch_multi (see samplea_batch): returns (b_new, rng_new, stats dict)."""
    L = lib()
    b = np.array(b, dtype=np.float64)
    rng = np.array(rng, dtype=np.uint64)
    apar = np.ascontiguousarray(apar, dtype=np.float64)
    st, keep = _stats(b.shape[0], trace_cap if devices is False else 0)
    dp, u64p, u32p = C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    args = (b.ctypes.data_as(dp), b.shape[0], counts.I, shape, scale, counts.N.ctypes.data_as(u32p),
            counts.T.ctypes.data_as(u32p), apar.ctypes.data_as(dp), rng.ctypes.data_as(u64p), loops, C.byref(st))
    rc = L.stb_sampleb_batch(*args) if devices is False else L.stb_sampleb_batch_multi(*_device_list(devices), *args)
    if rc:
        raise RuntimeError(f"stb_sampleb_batch failed ({rc}): " + L.stb_last_error().decode())
    return b, rng, {"evals": st.evals, "rounds": st.rounds, "eval_ms": st.eval_ms, "trace": keep}


RAND31_DTYPE = np.dtype([("r", np.int32, 31), ("f", np.int32), ("b", np.int32)])  # stb_rand31_t


def rand31_states(seeds):
    """per-chain rand() streams: element c is in the state srand(seeds[c]) puts glibc's generator in"""
    L = lib()
    g = np.zeros(len(seeds), dtype=RAND31_DTYPE)
    for c, s in enumerate(seeds):
        L.stb_rand31_seed(g.ctypes.data + c * RAND31_DTYPE.itemsize, int(s))
    return g


def samplea_batch_ars(a, counts, bpar, rnd, bpar_per_chain=False, trace_cap=0):
    """stb_samplea_batch_ars: returns (a_new, rnd_new, stats dict)."""
    L = lib()
    a = np.array(a, dtype=np.float64)
    rnd = np.array(rnd, dtype=RAND31_DTYPE)
    bpar = np.ascontiguousarray(bpar, dtype=np.float64)
    st, keep = _stats(a.shape[0], trace_cap)
    dp = C.POINTER(C.c_double)
    rc = L.stb_samplea_batch_ars(a.ctypes.data_as(dp), a.shape[0], *counts.args(), bpar.ctypes.data_as(dp),
                                 int(bpar_per_chain), rnd.ctypes.data, C.byref(st))
    if rc:
        raise RuntimeError(f"stb_samplea_batch_ars failed ({rc}): " + L.stb_last_error().decode())
    return a, rnd, {"evals": st.evals, "rounds": st.rounds, "eval_ms": st.eval_ms, "trace": keep}


def sampleb_batch_ars(b, counts, shape, scale, apar, rng, rnd, trace_cap=0):
    """stb_sampleb_batch_ars: returns (b_new, rng_new, rnd_new, stats dict)."""
    L = lib()
    b = np.array(b, dtype=np.float64)
    rng = np.array(rng, dtype=np.uint64)
    rnd = np.array(rnd, dtype=RAND31_DTYPE)
    apar = np.ascontiguousarray(apar, dtype=np.float64)
    st, keep = _stats(b.shape[0], trace_cap)
    dp, u64p, u32p = C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    rc = L.stb_sampleb_batch_ars(b.ctypes.data_as(dp), b.shape[0], counts.I, shape, scale,
                                 counts.N.ctypes.data_as(u32p), counts.T.ctypes.data_as(u32p), apar.ctypes.data_as(dp),
                                 rng.ctypes.data_as(u64p), rnd.ctypes.data, C.byref(st))
    if rc:
        raise RuntimeError(f"stb_sampleb_batch_ars failed ({rc}): " + L.stb_last_error().decode())
    return b, rng, rnd, {"evals": st.evals, "rounds": st.rounds, "eval_ms": st.eval_ms, "trace": keep}
