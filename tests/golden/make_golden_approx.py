"""Generates tests/golden/approx.json from the UNMODIFIED reference (slice/polygamma build,
oracle/_ref/libstb_ref_slice.so): S_approx, S_approx_da (lib/sapprox.c), gammadiff, psidiff and the
g/p/q caches (lib/lgamma.c), the polygamma functions (lib/polygamma.c).  Run in the build container:

    python tests/golden/make_golden_approx.py
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import harness  # noqa: E402


class GCache(C.Structure):
    _fields_ = [("par", C.c_double), ("lgpar", C.c_double), ("cache", C.c_double * 100)]


def declare(L):
    d, i, f = C.c_double, C.c_int, C.c_float
    L.S_approx.restype, L.S_approx.argtypes = d, [i, i, f]
    L.S_approx_da.restype, L.S_approx_da.argtypes = d, [i, i, f]
    L.gammadiff.restype, L.gammadiff.argtypes = d, [i, d, d]
    L.psidiff.restype, L.psidiff.argtypes = d, [i, d, d]
    for c in "gpq":
        getattr(L, c + "cache_init").restype = None
        getattr(L, c + "cache_init").argtypes = [C.POINTER(GCache), d]
        getattr(L, c + "cache_value").restype = d
        getattr(L, c + "cache_value").argtypes = [C.POINTER(GCache), i]
    for name in ("MLdigamma", "MLtrigamma", "MLtetragamma", "MLpentagamma"):
        getattr(L, name).restype, getattr(L, name).argtypes = d, [d]
    L.MLpsigamma.restype, L.MLpsigamma.argtypes = d, [d, d]
    return L


APPROX_A = [0.0005, 0.01, 0.1, 0.3, 0.45, 0.7, 0.9]   # (0.5 is singular for m >= 2: lgamma(1-2a) at a pole)
APPROX_N = [2, 3, 4, 5, 7, 10, 50, 1000, 100000]
DIFF_N = [0, 1, 2, 3, 4, 5, 17, 99, 100, 1999, 2000, 5000]
DIFF_ALPHA = [0.05, 0.2, 0.5, 0.7, 3.25]
CACHE_P = [0.3, 0.015, 0.9]
POLY_X = [0.01, 0.3, 1.0, 2.5, 9.99, 10.0, 14.9, 40.0, 1234.5]

if __name__ == "__main__":
    R = declare(C.CDLL(harness.REF_SLICE_SO))
    out = {"generator": "tests/golden/make_golden_approx.py", "source": "oracle/_ref/libstb_ref_slice.so",
           "S_approx": [], "S_approx_da": [], "gammadiff": [], "psidiff": [], "cache": [], "polygamma": []}
    for a in APPROX_A:
        for n in APPROX_N:
            for m in range(1, 6):
                out["S_approx"].append([n, m, a, R.S_approx(n, m, a)])
                out["S_approx_da"].append([n, m, a, R.S_approx_da(n, m, a)])
    for al in DIFF_ALPHA:
        for n in DIFF_N:
            for known in (0, 1):
                lga = float(__import__("math").lgamma(al)) if known else 0.0
                out["gammadiff"].append([n, al, lga, R.gammadiff(n, al, lga)])
                pa = R.MLdigamma(al) if known else 0.0
                if n >= 2000 and not known and al <= 0.5:
                    continue  # the reference's slip (lgamma for digamma, lib/lgamma.c:224-227): not a golden value
                out["psidiff"].append([n, al, pa, R.psidiff(n, al, pa)])
    for p in CACHE_P:
        for kind in "gpq":
            c = GCache()
            getattr(R, kind + "cache_init")(C.byref(c), p)
            for j in (-1, 0, 1, 2, 3, 4, 10, 99, 100, 250, 4, 10):
                out["cache"].append([kind, p, j, getattr(R, kind + "cache_value")(C.byref(c), j)])
    for x in POLY_X:
        out["polygamma"].append([x, R.MLdigamma(x), R.MLtrigamma(x), R.MLtetragamma(x), R.MLpentagamma(x)])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "approx.json")

    def enc(v):
        if isinstance(v, float):
            if v != v:
                return "nan"
            if v in (float("inf"), float("-inf")):
                return "inf" if v > 0 else "-inf"
            return repr(v)
        return v

    for k in ("S_approx", "S_approx_da", "gammadiff", "psidiff", "cache", "polygamma"):
        out[k] = [[enc(v) for v in row] for row in out[k]]
    json.dump(out, open(path, "w"), indent=0)
    print("wrote", path, {k: len(out[k]) for k in out if isinstance(out[k], list)})
