"""Golden values of BASELINE config 2 (N=200 000, M=20 000, a=0.7) from the UNMODIFIED reference library
(oracle/_ref/libstb_ref.so, built by oracle/build_ref.sh): the last row, the last column and seeded random
cells of the log S table (S_make ... S_STABLE) and of the V table (S_make ... S_UVTABLE), read with the
reference's own S_S / S_V.  Needs ~31 GB of host memory and ~4 minutes on one core; run once in the build
container, the result is committed (tests/golden/config2.npz) and compared on the GPU by
tests/test_table_large_gpu.py.

    python tests/golden/make_golden_config2.py
"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
N, M, A = 200_000, 20_000, 0.7
SEED, N_S, N_V = 20260218, 1_000_000, 250_000


def cells(count, seed):
    """seeded (n, m) with 2 <= m <= min(n, M): the same on the generating and on the checking side"""
    rng = np.random.default_rng(seed)
    n = rng.integers(2, N + 1, size=count).astype(np.uint32)
    m = (2 + (rng.random(count) * (np.minimum(n, M) - 1)).astype(np.uint32)).astype(np.uint32)
    return n, np.minimum(m, np.minimum(n, M)).astype(np.uint32)


def main():
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libstb_ref.so"))
    u, d, vp = C.c_uint, C.c_double, C.c_void_p
    R.S_make.restype, R.S_make.argtypes = vp, [u, u, u, u, d, C.c_uint32]
    R.S_free.restype, R.S_free.argtypes = None, [vp]
    for f in (R.S_S, R.S_V):
        f.restype, f.argtypes = d, [vp, u, u]
    out = {}
    which = sys.argv[1] if len(sys.argv) > 1 else "SV"
    for name, flag, fn, count, seed in (("S", 1, R.S_S, N_S, SEED), ("V", 2, R.S_V, N_V, SEED + 1)):
        if name not in which:
            continue
        t0 = time.time()
        # two rows and columns to spare: S_V treats look-ups within one cell of the table's edge as a request to grow
        # (lib/stable.c:903) and, at the maximum extent, reads past its rows; cells with n <= N, m <= M do not depend on
        # the extent of the table they are part of
        sp = R.S_make(N + 2, M + 2, N + 2, M + 2, A, flag)
        assert sp, "S_make failed (needs ~31 GB)"
        print(f"{name}: reference S_make took {time.time() - t0:.0f} s", file=sys.stderr)
        out[name + "_lastrow"] = np.array([fn(sp, N, m) for m in range(1, M + 1)])
        out[name + "_lastcol"] = np.array([fn(sp, n, M) for n in range(M, N + 1)])
        n, m = cells(count, seed)
        out[name + "_cells"] = np.array([fn(sp, int(a), int(b)) for a, b in zip(n, m)])
        R.S_free(sp)
    if "O" in which:
        # V from the oracle's restatement (oracle/stirling_oracle.c, bit-identical to the reference wherever both run,
        # tests/test_oracle_vs_reference.py): the reference's own V-only table of this size dies in its first S_V call
        # (its V rows are addressed with 32-bit products: 200 002 x 20 002 > 2^32)
        from tests import harness

        t0 = time.time()
        V = np.empty((N, M))
        harness.oracle().orc_fill_V(N, M, A, V.ctypes.data_as(C.POINTER(C.c_double)), M)
        print(f"V: oracle fill took {time.time() - t0:.0f} s", file=sys.stderr)
        out["V_lastrow"] = V[N - 1, :].copy()
        out["V_lastrow"][0] = 0.0  # m = 1: S_V answers 0
        out["V_lastcol"] = V[M - 1:, M - 1].copy()
        n, m = cells(N_V, SEED + 1)
        out["V_cells"] = V[n.astype(np.int64) - 1, m.astype(np.int64) - 1].copy()
        del V
    dst = os.path.join(HERE, "config2.npz" if which == "SV" else f"config2_{which}.npz")
    np.savez_compressed(dst, N=N, M=M, a=A, seed=SEED, **out)
    print("wrote", dst, file=sys.stderr)


if __name__ == "__main__":
    main()
