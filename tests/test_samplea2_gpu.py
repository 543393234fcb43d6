"""samplea2 (lib/samplea.c:227-341, the reference's -DSAMPLEA_M build) on the GPU: the partition kernel
against the oracle's restatement (itself pinned to the reference in tests/test_partition_cpu.py),
bit for bit, and the whole call against the reference run in a process of its own."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import libstb_b200 as stb
from tests import harness
from tests.test_partition_cpu import gcache_sequence, oracle_sizes, probe_counts, run_probe

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not os.path.exists(harness.REF_SLICE_M_SO), reason="reference build not present")


def _nodes(n_rows, t_rows):
    idx = [(i, k) for i, (nr, tr) in enumerate(zip(n_rows, t_rows)) for k, (n, t) in enumerate(zip(nr, tr))
           if 1 < int(t) < int(n)]
    n = np.array([n_rows[i][k] for i, k in idx], dtype=np.uint32)
    t = np.array([t_rows[i][k] for i, k in idx], dtype=np.uint16)
    return idx, n, t


def _logu(libc, t, draw_seed, exact):
    libc.srand48(draw_seed)
    if not exact:
        return np.array([math.log(libc.drand48()) for _ in t])
    out = []
    for tj in t:
        block = np.zeros(int(tj) - 1)
        for M in range(int(tj) - 1, 0, -1):
            block[M - 1] = math.log(libc.drand48())
        out.append(block)
    return np.concatenate(out) if out else np.zeros(0)


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("seed,I,K,nmax,a", [(42, 6, 8, 300, 0.5), (7, 10, 12, 1500, 0.2), (3, 4, 30, 60, 0.8),
                                             (11, 20, 25, 3000, 0.0)])
def test_partition_kernel_matches_oracle(seed, I, K, nmax, a, exact):
    (n_rows, t_rows), libc = probe_counts(seed, I, K, nmax)
    idx, n, t = _nodes(n_rows, t_rows)
    maxn, maxt = int(n.max()) + 1, int(t.max()) + 1
    want = oracle_sizes(n_rows, t_rows, a, libc, 900 + seed, exact=exact)
    tab = stb.Table(30, 12, maxn, maxt, a, stb.S_STABLE)  # grown by the call to cover the nodes
    m, off = tab.partition_sample(a, n, t, _logu(libc, t, 900 + seed, exact), exact=exact)
    tab.free()
    bad = [(i, k) for j, (i, k) in enumerate(idx)
           if not np.array_equal(m[off[j]:off[j] + int(t[j]) - 1], want[(i, k)])]
    assert not bad, f"{len(bad)} of {len(idx)} nodes differ, first {bad[0]}"
    for j in range(len(idx)):  # sizes are a composition of n into t parts
        mm = m[off[j]:off[j] + int(t[j]) - 1].astype(np.int64)
        assert (mm >= 1).all() and mm.sum() <= int(n[j]) - 1
    if exact:
        assert any((want[key][:-1] > 1).any() for key in want if len(want[key]) > 1), "exact mode spreads customers"


def test_partition_rejects_bad_nodes():
    tab = stb.Table(50, 10, 100, 20, 0.5, stb.S_STABLE)
    with pytest.raises(RuntimeError):
        tab.partition_sample(0.5, [200], [5], [-0.1])  # beyond maxN
    with pytest.raises(RuntimeError):
        tab.partition_sample(0.5, [30], [30], [-0.1])  # t == n has no free size
    with pytest.raises(RuntimeError):
        tab.partition_sample(0.5, [30], [1], [-0.1])
    m, off = tab.partition_sample(0.5, np.zeros(0), np.zeros(0), np.zeros(0))
    assert m.shape == (0,)
    tab.free()


def stb_squeeze():
    return 0.2  # SQUEEZEA, include/psample.h


def _call_samplea2(L, tab, a0, n_rows, t_rows, loops):
    I = len(n_rows)
    cts = stb.Counts(n_rows, t_rows)
    bpar = np.full(I, 10.0)
    return L.samplea2(a0, tab.sp, *cts.args(), None, bpar.ctypes.data_as(C.POINTER(C.c_double)), None, loops, 0)


@needs_ref
@pytest.mark.parametrize("seed,I,K,nmax,a0,loops", [(42, 6, 8, 300, 0.5, 1), (7, 10, 12, 1500, 0.2, 2),
                                                    (3, 4, 30, 60, 0.8, 1)])
def test_samplea2_matches_the_reference(seed, I, K, nmax, a0, loops):
    """Same 48-bit stream: the same discount, bit for bit (the density is summed on the host in the
    reference's order with the same libm), and the same number of uniforms consumed."""
    ref = run_probe(seed, I, K, nmax, a0, 900 + seed, loops)
    (n_rows, t_rows), libc = probe_counts(seed, I, K, nmax)
    L = stb.lib()
    maxn = max(int(r.max()) for r in n_rows) + 1
    maxt = max(int(r.max()) for r in t_rows) + 1
    tab = stb.Table(maxn, maxt, maxn, maxt, a0, stb.S_STABLE)
    libc.srand48(900 + seed)
    a1 = _call_samplea2(L, tab, a0, n_rows, t_rows, loops)
    nxt = libc.drand48()
    tab.free()
    assert repr(nxt) == ref["next_u"], "different numbers of draws consumed"
    assert repr(a1) == ref["a"]
    assert 0.01 <= a1 <= 0.98


def test_samplea2_exact_mode_and_ars_mode():
    (n_rows, t_rows), libc = probe_counts(5, 8, 10, 400)
    libc.srand.argtypes = [C.c_uint]
    L = stb.lib()
    maxn = max(int(r.max()) for r in n_rows) + 1
    maxt = max(int(r.max()) for r in t_rows) + 1
    tab = stb.Table(maxn, maxt, maxn, maxt, 0.5, stb.S_STABLE)
    _, _, t = _nodes(n_rows, t_rows)
    assert L.stb_set_partition_mode(stb.STB_PARTITION_EXACT) == stb.STB_PARTITION_REFERENCE
    try:
        libc.srand48(77)
        a1 = _call_samplea2(L, tab, 0.5, n_rows, t_rows, 1)
        after = libc.drand48()
        # the partition step consumed one uniform per sampled size, before the slice sampler's own
        libc.srand48(77)
        for _ in range(int((t.astype(np.int64) - 1).sum())):
            libc.drand48()
        assert libc.drand48() != after
        assert 0.5 - stb_squeeze() - 1e-12 <= a1 <= 0.98  # the slice sampler's bounds: [a - SQUEEZEA, A_MAX]
        # ARS mode (the reference's own call passes a NULL data pointer and cannot run)
        old = L.stb_set_sampler(stb.STB_SAMPLER_ARS)
        try:
            libc.srand(5)
            a2 = _call_samplea2(L, tab, 0.5, n_rows, t_rows, 1)
            assert 0.5 - stb_squeeze() - 1e-12 <= a2 <= 0.5 + stb_squeeze() + 1e-12
        finally:
            L.stb_set_sampler(old)
    finally:
        L.stb_set_partition_mode(stb.STB_PARTITION_REFERENCE)
        tab.free()


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("seed,I,K,nmax,loops", [(42, 6, 8, 300, 1), (7, 10, 12, 1500, 2)])
def test_samplea2_batch_equals_the_scalar_call_chain_by_chain(seed, I, K, nmax, loops, exact):
    """samplea2 batched over chains (one table per chain from the sweep, the partition kernel over
    chains x nodes with the uniforms made on the device from each chain's stream, the likelihood a device
    reduction over the chain's histogram of sizes).  Chain c's discount and stream must equal the scalar
    samplea2's on a table at a0[c] with the stream of the same seed -- and the scalar call equals the reference's
    -DSAMPLEA_M build bit for bit (test above)."""
    (n_rows, t_rows), libc = probe_counts(seed, I, K, nmax)
    L = stb.lib()
    cts = stb.Counts(n_rows, t_rows)
    bpar = np.full(I, 10.0)
    a0 = np.array([0.15, 0.3, 0.5, 0.5, 0.7, 0.9])
    seeds = [900 + seed + 13 * c for c in range(a0.shape[0])]
    maxn = max(int(r.max()) for r in n_rows) + 1
    maxt = max(int(r.max()) for r in t_rows) + 1
    old = L.stb_set_partition_mode(stb.STB_PARTITION_EXACT if exact else stb.STB_PARTITION_REFERENCE)
    try:
        want_a, want_next = [], []
        for c, (ac, sd) in enumerate(zip(a0, seeds)):
            tab = stb.Table(maxn, maxt, maxn, maxt, float(ac), stb.S_STABLE)
            libc.srand48(sd)
            want_a.append(_call_samplea2(L, tab, float(ac), n_rows, t_rows, loops))
            want_next.append(libc.drand48())
            tab.free()
        rng = np.array([L.stb_rng48_state(sd) for sd in seeds], dtype=np.uint64)
        a1, rng1, stats = stb.samplea2_batch(a0, cts, bpar, rng, loops=loops)
        got_next = [L.stb_rng48_drand(C.byref(C.c_uint64(int(x)))) for x in rng1]
        assert got_next == want_next, "a chain's stream is out of step with the scalar call's"
        assert [float(x) for x in a1] == [float(x) for x in want_a]
        assert stats["evals"] >= a0.shape[0] * loops
    finally:
        L.stb_set_partition_mode(old)


def test_samplea2_batch_many_chains_in_bounds():
    """more chains than one wave of tables holds; every draw inside the slice sampler's bounds"""
    (n_rows, t_rows), _ = probe_counts(9, 12, 20, 800)
    L = stb.lib()
    cts = stb.Counts(n_rows, t_rows)
    Cn = 700
    a0 = 0.05 + 0.9 * (np.arange(Cn) + 0.5) / Cn
    rng = np.array([L.stb_rng48_state(5000 + c) for c in range(Cn)], dtype=np.uint64)
    a1, rng1, stats = stb.samplea2_batch(a0, cts, np.full(12, 10.0), rng, loops=1)
    assert ((a1 >= np.maximum(0.01, a0 - stb_squeeze()) - 1e-12) & (a1 <= 0.98)).all()
    assert (rng1 != rng).all() and (a1 != a0).any()
    # the same chains in two halves: a chain's result does not depend on who runs beside it
    a2, rng2, _ = stb.samplea2_batch(a0[:300], cts, np.full(12, 10.0), rng[:300], loops=1)
    assert np.array_equal(a2, a1[:300]) and np.array_equal(rng2, rng1[:300])
