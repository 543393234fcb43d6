"""The oracle's restatement of samplea2's seat-partition sampler (lib/samplea.c:290-321), on the CPU:
pinned against the REFERENCE itself (a -DSAMPLEA_M build, observed through its gcache_value calls),
and the exact mode against the identity and the frequencies that define it."""
import ctypes as C
import json
import math
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import harness

HERE = os.path.dirname(os.path.abspath(__file__))
PROBE = os.path.join(HERE, "ref_samplea2_probe.py")
needs_ref = pytest.mark.skipif(not os.path.exists(harness.REF_SLICE_M_SO), reason="reference build not present")


def run_probe(seed, I, K, nmax, a0, draw_seed, loops):
    out = subprocess.check_output([sys.executable, PROBE] + [str(v) for v in (seed, I, K, nmax, a0, draw_seed, loops)])
    return json.loads(out)


def probe_counts(seed, I, K, nmax):
    sys.path.insert(0, HERE)
    import ref_samplea2_probe as P
    libc = C.CDLL(None)
    libc.drand48.restype = C.c_double
    libc.srand48.argtypes = [C.c_long]
    return P.counts(libc, seed, I, K, nmax), libc


def gcache_sequence(n_rows, t_rows, sizes_of):
    """The j arguments of aterms2's gcache_value calls (lib/samplea.c:118-143) given the sampled sizes."""
    seq = []
    for i, (nr, tr) in enumerate(zip(n_rows, t_rows)):
        for k, (n, t) in enumerate(zip(nr, tr)):
            n, t = int(n), int(t)
            if n == 0 or t == n:
                continue
            if t == 1:
                seq.append(n - 1)
                continue
            m = sizes_of(i, k)
            for l in range(t - 2, -1, -1):
                if m[l] > 1:
                    seq.append(int(m[l]) - 1)
                n -= int(m[l])
            if n > 0:
                seq.append(n - 1)
    return seq


def oracle_sizes(n_rows, t_rows, a, libc, draw_seed, exact=False):
    """Every node's sizes from the oracle, uniforms drawn like samplea2 draws them."""
    O = harness.oracle()
    maxn = max(int(r.max()) for r in n_rows) + 1
    maxt = max(int(r.max()) for r in t_rows) + 1
    tb = O.orc_make(maxn, maxt, maxn, maxt, a, 1)
    libc.srand48(draw_seed)
    out = {}
    for i, (nr, tr) in enumerate(zip(n_rows, t_rows)):
        for k, (n, t) in enumerate(zip(nr, tr)):
            n, t = int(n), int(t)
            if 1 < t < n:
                cnt = t - 1 if exact else 1
                logu = np.zeros(t - 1)
                if exact:
                    for M in range(t - 1, 0, -1):
                        logu[M - 1] = math.log(libc.drand48())
                else:
                    logu[0] = math.log(libc.drand48())
                m = np.zeros(t - 1, dtype=np.uint16)
                O.orc_partition_node(tb, a, n, t, logu.ctypes.data_as(C.POINTER(C.c_double)),
                                     m.ctypes.data_as(C.POINTER(C.c_uint16)), int(exact))
                out[(i, k)] = m
    O.orc_free(tb)
    return out


@needs_ref
@pytest.mark.parametrize("seed,I,K,nmax,a0", [(42, 6, 8, 300, 0.5), (7, 10, 12, 1500, 0.2), (3, 4, 30, 60, 0.8)])
def test_oracle_partition_matches_the_reference(seed, I, K, nmax, a0):
    ref = run_probe(seed, I, K, nmax, a0, 900 + seed, 1)
    (n_rows, t_rows), libc = probe_counts(seed, I, K, nmax)
    sizes = oracle_sizes(n_rows, t_rows, a0, libc, 900 + seed)
    want = gcache_sequence(n_rows, t_rows, lambda i, k: sizes[(i, k)])
    E = len(want)
    assert E > 0 and ref["calls"] % E == 0, "whole evaluations only"
    assert ref["seq"][:E] == want
    # what the reference's walk produces whenever S(n,t) is large (rem starts at S(n,t) + log u while the
    # terms are normalised by S(n,t)): one big table and singletons, whatever u
    for (i, k), m in sizes.items():
        n, t = int(n_rows[i][k]), int(t_rows[i][k])
        if n >= 40:
            assert m[t - 2] == n - (t - 1) and (m[:t - 2] == 1).all()


def test_logminus():
    O = harness.oracle()
    assert O.orc_logminus(1.0, 1.0) == -math.inf and O.orc_logminus(0.0, 2.0) == -math.inf
    assert O.orc_logminus(3.0, 1.0) == 3.0 + math.log(1 - math.exp(-2.0))
    assert O.orc_logminus(100.0, 0.0) == 100.0 - math.exp(-100.0)


@pytest.mark.parametrize("a", [0.0, 0.3, 0.7, 0.95])
def test_exact_mode_probabilities_sum_to_one(a):
    """sum_l C(N-1,l-1) (1-a)_{l-1} S^{N-l}_M = S^N_{M+1}: the table a given customer sits at has some size."""
    O = harness.oracle()
    tb = O.orc_make(400, 60, 400, 60, a, 1)
    for N, M in [(5, 1), (9, 3), (50, 7), (400, 59), (399, 2), (120, 119)]:
        lp = [O.orc_partition_logp(tb, a, N, M, l) for l in range(1, N - M + 1)]
        assert math.fsum(math.exp(v) for v in lp) == pytest.approx(1.0, abs=1e-11)
    O.orc_free(tb)


def test_exact_mode_follows_its_distribution():
    """Inverse-CDF on a uniform grid of u: the share of each size equals its probability to 1/grid."""
    O = harness.oracle()
    a, N, t = 0.4, 12, 2
    tb = O.orc_make(20, 10, 20, 10, a, 1)
    grid = 4000
    hist = np.zeros(N, dtype=np.int64)
    m = np.zeros(1, dtype=np.uint16)
    for j in range(grid):
        logu = np.array([math.log((j + 0.5) / grid)])
        O.orc_partition_node(tb, a, N, t, logu.ctypes.data_as(C.POINTER(C.c_double)),
                             m.ctypes.data_as(C.POINTER(C.c_uint16)), 1)
        hist[int(m[0])] += 1
    # the walk subtracts masses from the top: size l is chosen for u in (1 - cdf(l), 1 - cdf(l-1)]
    p = np.array([0.0] + [math.exp(O.orc_partition_logp(tb, a, N, 1, l)) for l in range(1, N)])
    assert hist.sum() == grid and np.abs(hist / grid - p).max() <= 1.5 / grid
    O.orc_free(tb)
