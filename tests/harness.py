"""ctypes access to the CPU oracle (oracle/liboracle.so) and, when it was built in this
container, the unmodified reference (oracle/_ref/libstb_ref*.so).  Test infrastructure."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libstb_ref.so")
REF_SLICE_SO = os.path.join(ORACLE_DIR, "_ref", "libstb_ref_slice.so")
REF_SLICE_M_SO = os.path.join(ORACLE_DIR, "_ref", "libstb_ref_slice_m.so")

S_STABLE, S_UVTABLE, S_FLOAT, S_ASYMPT = 1, 2, 4, 64

_oracle = None


def oracle() -> C.CDLL:
    global _oracle
    if _oracle is None:
        src_newer = not os.path.exists(ORACLE_SO) or any(
            os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(ORACLE_SO)
            for f in os.listdir(ORACLE_DIR)
            if f.endswith((".c", ".h"))
        )
        if src_newer:
            subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(ORACLE_SO)
        u, d, dp, vp = C.c_uint, C.c_double, C.POINTER(C.c_double), C.c_void_p
        L.orc_fill_S1.restype, L.orc_fill_S1.argtypes = None, [u, d, dp]
        L.orc_fill_S.restype, L.orc_fill_S.argtypes = None, [u, u, d, dp, C.c_size_t]
        L.orc_fill_V.restype, L.orc_fill_V.argtypes = None, [u, u, d, dp, C.c_size_t]
        L.orc_make.restype, L.orc_make.argtypes = vp, [u, u, u, u, d, C.c_uint32]
        L.orc_free.restype, L.orc_free.argtypes = None, [vp]
        for n in ("orc_S", "orc_V", "orc_U", "orc_UV"):
            f = getattr(L, n)
            f.restype, f.argtypes = d, [vp, u, u]
        L.orc_S1.restype, L.orc_S1.argtypes = d, [vp, u]
        L.orc_asympt.restype, L.orc_asympt.argtypes = d, [d, u, u]
        L.orc_V_asympt.restype, L.orc_V_asympt.argtypes = d, [d, u, u]
        L.orc_logminus.restype, L.orc_logminus.argtypes = d, [d, d]
        u32p = C.POINTER(C.c_uint32)
        L.orc_ti_gibbs.restype = None
        L.orc_ti_gibbs.argtypes = [vp, d, d, C.c_size_t, u32p, u32p, C.POINTER(C.c_float), C.c_uint32, u32p,
                                   C.POINTER(C.c_uint16), u32p, C.POINTER(C.c_uint64), C.c_int, C.c_int]
        L.orc_partition_node.restype = None
        L.orc_partition_node.argtypes = [vp, d, u, u, dp, C.POINTER(C.c_uint16), C.c_int]
        L.orc_partition_logp.restype, L.orc_partition_logp.argtypes = d, [vp, d, u, u, u]
        L.orc_cells_S.restype, L.orc_cells_S.argtypes = C.c_uint64, [C.c_uint64, C.c_uint64]
        L.orc_cells_V.restype, L.orc_cells_V.argtypes = C.c_uint64, [C.c_uint64, C.c_uint64]
        _oracle = L
    return _oracle


def oracle_tables(N, M, a, want_S=True, want_V=True):
    """Dense (N, M) float64 arrays from the oracle; cell (n,m) at [n-1, m-1], NaN where unset."""
    L = oracle()
    dp = C.POINTER(C.c_double)
    S = V = None
    if want_S:
        S = np.full((N, M), np.nan)
        L.orc_fill_S(N, M, a, S.ctypes.data_as(dp), M)
    if want_V:
        V = np.full((N, M), np.nan)
        L.orc_fill_V(N, M, a, V.ctypes.data_as(dp), M)
    return S, V


def valid_mask(N, M, for_V=False):
    """Cells a table stores: S: 1<=m<=min(n,M) ; V: 2<=m<=min(n,M)."""
    n = np.arange(1, N + 1)[:, None]
    m = np.arange(1, M + 1)[None, :]
    mask = m <= n
    if for_V:
        mask &= m >= 2
    return mask


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _declare_ref(L):
    u, d, vp = C.c_uint, C.c_double, C.c_void_p
    L.S_make.restype, L.S_make.argtypes = vp, [u, u, u, u, d, C.c_uint32]
    L.S_remake.restype, L.S_remake.argtypes = C.c_int, [vp, d]
    L.S_free.restype, L.S_free.argtypes = None, [vp]
    for n in ("S_S", "S_U", "S_UV", "S_V", "S_asympt"):
        f = getattr(L, n)
        f.restype, f.argtypes = d, [vp, u, u]
    L.S_S1.restype, L.S_S1.argtypes = d, [vp, u]
    return L


_ref = {}


def ref(slice_build=False) -> C.CDLL:
    """The unmodified reference, compiled by oracle/build_ref.sh (container only)."""
    key = bool(slice_build)
    if key not in _ref:
        _ref[key] = _declare_ref(C.CDLL(REF_SLICE_SO if slice_build else REF_SO))
    return _ref[key]


def close(x, y, rel=1e-12):
    """|x-y| <= rel*max(1,|y|) elementwise; NaN/inf must match exactly."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    fin = np.isfinite(y)
    ok = np.where(fin, np.abs(x - y) <= rel * np.maximum(1.0, np.abs(y)), (x == y) | (np.isnan(x) & np.isnan(y)))
    return ok


def max_err(x, y):
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    return float(np.max(np.abs(x - y) / np.maximum(1.0, np.abs(y))))
