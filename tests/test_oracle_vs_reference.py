"""Pins the CPU oracle (oracle/stirling_oracle.c, our restatement) against the UNMODIFIED
reference compiled from /root/reference by oracle/build_ref.sh.  Same libm, same operation
order => every cell must be bit-identical.  Skipped where oracle/_ref was not built (it is
built in the development container and travels to the GPU box as a .so)."""
import numpy as np
import pytest

from tests import harness

pytestmark = pytest.mark.skipif(not harness.have_ref(), reason="oracle/_ref not built")

FLAGS = harness.S_STABLE | harness.S_UVTABLE


@pytest.mark.parametrize("N,M,a", [(10, 10, 0.0), (10, 10, 0.3), (64, 33, 0.5), (300, 40, 0.01),
                                    (300, 300, 0.98), (1500, 200, 0.7), (777, 129, 0.25)])
def test_tables_bit_identical(N, M, a):
    L = harness.ref()
    S, V = harness.oracle_tables(N, M, a)
    sp = L.S_make(N, M, N, M, a, FLAGS)
    try:
        for n in range(2, N + 1):
            top = min(n, M)
            refS = np.array([L.S_S(sp, n, m) for m in range(1, top + 1)])
            assert np.array_equal(refS, S[n - 1, :top]), f"S row {n}"
            if top >= 2:
                refV = np.array([L.S_V(sp, n, m) for m in range(2, top + 1)])
                assert np.array_equal(refV, V[n - 1, 1:top]), f"V row {n}"
    finally:
        L.S_free(sp)


def test_lookup_conventions_match_reference():
    """n==m, m==1, n<m, m==0, beyond bounds, S_ASYMPT: the oracle's table object answers like
    the reference's S_S/S_V/S_U/S_UV (lib/stable.c:875-974)."""
    L, O = harness.ref(), harness.oracle()
    for flags in (FLAGS, FLAGS | harness.S_ASYMPT):
        sp = L.S_make(30, 12, 30, 12, 0.3, flags)
        t = O.orc_make(30, 12, 30, 12, 0.3, flags)
        try:
            cells = [(2, 2), (2, 1), (0, 0), (5, 0), (3, 5), (5, 1), (5, 6), (5, 5), (31, 3), (20, 13),
                     (30, 12), (12, 12), (13, 12), (29, 11), (100, 4), (100000, 7)]
            for n, m in cells:
                assert L.S_S(sp, n, m) == O.orc_S(t, n, m), ("S", n, m, flags)
                if m >= 2:
                    assert L.S_V(sp, n, m) == O.orc_V(t, n, m), ("V", n, m, flags)
                if m >= 1:
                    assert L.S_UV(sp, n, m) == O.orc_UV(t, n, m), ("UV", n, m, flags)
                    assert L.S_U(sp, n, m) == O.orc_U(t, n, m), ("U", n, m, flags)
                if n >= 2 and m >= 2:
                    assert L.S_asympt(sp, n, m) == O.orc_asympt(0.3, n, m)
        finally:
            L.S_free(sp)
            O.orc_free(t)


def test_float_storage_is_rounded_fp64():
    """A freshly built S_FLOAT table equals (float) of the FP64 table."""
    L, O = harness.ref(), harness.oracle()
    N, M, a = 400, 60, 0.6
    sp = L.S_make(N, M, N, M, a, FLAGS | harness.S_FLOAT)
    t = O.orc_make(N, M, N, M, a, FLAGS | harness.S_FLOAT)
    try:
        for n in range(3, N + 1, 7):
            for m in range(2, min(n - 1, M) + 1, 3):
                assert L.S_S(sp, n, m) == O.orc_S(t, n, m)
                assert L.S_V(sp, n, m) == O.orc_V(t, n, m)
    finally:
        L.S_free(sp)
        O.orc_free(t)
