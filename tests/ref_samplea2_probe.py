"""Runs the REFERENCE's samplea2 (oracle/_ref/libstb_ref_slice_m.so, built with -DSAMPLEA_M) in a
process of its own and prints what it did as JSON.  Test infrastructure, started by
tests/test_samplea2_gpu.py -- never imported by the product.

The table sizes samplea2 samples are private to it (lib/samplea.c:262-321); they are recovered from
the gcache_value(size - 1) calls of its first aterms2 evaluation, recorded by oracle/_ref/shim_gcache.so
(loaded RTLD_GLOBAL first so that the reference's calls bind to it).

usage: ref_samplea2_probe.py seed I K nmax a0 draw_seed loops
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(os.path.dirname(HERE), "oracle", "_ref")


def counts(libc, seed, I, K, nmax):
    """The C4 recipe of SURVEY.md 8d at reduced size (the same as tests/test_samplers_gpu.py)."""
    libc.srand48(seed)
    n_rows, t_rows = [], []
    for _ in range(I):
        nr, tr = [], []
        for _ in range(K):
            n = 1 + int(libc.drand48() ** 3 * nmax)
            t = min(n, 1 + int(libc.drand48() * n ** 0.6))
            nr.append(n)
            tr.append(t)
        n_rows.append(np.array(nr, dtype=np.uint32))
        t_rows.append(np.array(tr, dtype=np.uint16))
    return n_rows, t_rows


def main():
    seed, I, K, nmax = (int(v) for v in sys.argv[1:5])
    a0, draw_seed, loops = float(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])
    libc = C.CDLL(None)
    libc.drand48.restype = C.c_double
    libc.srand48.argtypes = [C.c_long]
    shim = C.CDLL(os.path.join(REF_DIR, "shim_gcache.so"), mode=C.RTLD_GLOBAL)
    R = C.CDLL(os.path.join(REF_DIR, "libstb_ref_slice_m.so"))
    cap = 1 << 24
    rec = np.zeros(cap, dtype=np.int32)
    shim.shim_set.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    shim.shim_count.restype = C.c_size_t
    shim.shim_set(C.cast(R.gcache_value, C.c_void_p), rec.ctypes.data, cap)

    n_rows, t_rows = counts(libc, seed, I, K, nmax)
    d, u32p, u16p = C.c_double, C.POINTER(C.c_uint32), C.POINTER(C.c_uint16)
    Kc = np.full(I, K, dtype=np.int32)
    T = np.array([int(r.sum()) for r in t_rows], dtype=np.uint32)
    n_pp = (u32p * I)(*[r.ctypes.data_as(u32p) for r in n_rows])
    t_pp = (u16p * I)(*[r.ctypes.data_as(u16p) for r in t_rows])
    bpar = np.full(I, 10.0)
    maxn = max(int(r.max()) for r in n_rows) + 1
    maxt = max(int(r.max()) for r in t_rows) + 1
    R.S_make.restype = C.c_void_p
    R.S_make.argtypes = [C.c_uint] * 4 + [d, C.c_uint32]
    sp = R.S_make(maxn, maxt, maxn, maxt, a0, 1)
    R.samplea2.restype = d
    R.samplea2.argtypes = [d, C.c_void_p, C.c_int, C.POINTER(C.c_int), u32p, C.POINTER(u32p), C.POINTER(u16p),
                           C.c_void_p, C.POINTER(d), C.c_void_p, C.c_int, C.c_int]
    libc.srand48(draw_seed)
    a1 = R.samplea2(a0, sp, I, Kc.ctypes.data_as(C.POINTER(C.c_int)), T.ctypes.data_as(u32p), n_pp, t_pp, None,
                    bpar.ctypes.data_as(C.POINTER(d)), None, loops, 0)
    nxt = libc.drand48()
    total = shim.shim_count()
    json.dump({"a": repr(a1), "next_u": repr(nxt), "calls": int(total), "seq": rec[:min(total, cap)].tolist()},
              sys.stdout)


if __name__ == "__main__":
    main()
