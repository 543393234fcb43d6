"""Several devices from one process (include/stb_b200.h, include/psample.h: *_multi; csrc/multi.c).

SURVEY.md 8(e): the discount sweep (the reference's unit: one S_remake per log-posterior evaluation,
lib/samplea.c:57-60) and the batched chains (lib/samplea.c:155-225, lib/sampleb.c:79-159) shard over
devices as independent units, no traffic during the work.  The bar: whatever the device list, every
unit's result equals the single-device result BIT FOR BIT.  A device named twice exercises the dealing,
the per-device worker threads and the result placement on a one-GPU box; the two-device tests run when
the box has two."""
import threading

import numpy as np
import pytest

import libstb_b200 as stb
from tests import test_samplers_gpu as tsg

pytestmark = pytest.mark.gpu


def _ndev():
    return stb.lib().stb_device_count()


def _sweep_inputs(N, M, na, npairs=5000, seed=5):
    rng = np.random.default_rng(seed)
    n = rng.integers(1, N + 1, size=npairs).astype(np.uint32)
    m = np.minimum(rng.integers(1, M + 1, size=npairs), n).astype(np.uint32)
    a = (np.arange(na) + 0.5) / na
    return n, m, a


def _single(N, M, n, m, a):
    w = stb.Sweep(N, M)
    w.set_pairs(n, m)
    out = w.run(a, gather=True, sums=True, lastrow=True)
    w.free()
    return out


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0], None])
def test_sweep_multi_equals_single_device(devices):
    N, M, na = 3000, 400, 13  # 13 units over 2 or 3 shares: uneven deal
    n, m, a = _sweep_inputs(N, M, na)
    g1, s1, r1 = _single(N, M, n, m, a)
    w = stb.SweepMulti(N, M, devices)
    assert w.ndev == (len(devices) if devices else _ndev())
    w.set_pairs(n, m)
    g, s, r = w.run(a, gather=True, sums=True, lastrow=True)
    assert np.array_equal(g, g1) and np.array_equal(s, s1) and np.array_equal(r, r1)
    # sums only (the samplers' form: distinct pairs with multiplicities), and fewer units than shares
    _, s2, _ = w.run(a[:1], gather=False, sums=True)
    assert np.array_equal(s2, s1[:1])
    assert w.last_fill_ms > 0 and len(w.device_ms) == w.ndev
    w.free()


@pytest.mark.skipif("_ndev() < 2", reason="needs two GPUs (run with gpurun --gpus 2)")
def test_sweep_two_devices_equal_one_device():
    N, M, na = 20000, 2000, 64
    n, m, a = _sweep_inputs(N, M, na, npairs=20000)
    g1, s1, r1 = _single(N, M, n, m, a)
    for devices in ([0, 1], [1, 0], None):
        w = stb.SweepMulti(N, M, devices)
        w.set_pairs(n, m)
        g, s, r = w.run(a, gather=True, sums=True, lastrow=True)
        assert np.array_equal(g, g1) and np.array_equal(s, s1) and np.array_equal(r, r1), devices
        w.free()


def test_sweep_multi_errors_name_the_device():
    with pytest.raises(RuntimeError, match="does not exist"):
        stb.SweepMulti(100, 20, [0, 999])
    with pytest.raises(RuntimeError, match="bad extent"):
        stb.SweepMulti(10, 20, [0, 0])  # M > N: refused by every worker, the caller's thread has the text


@pytest.mark.parametrize("devices", [[0, 0], None])
def test_chains_multi_equal_single_device(devices):
    cts = tsg._counts(7, 30, 40, 400)
    C = 11
    L = stb.lib()
    a0 = 0.1 + 0.8 * (np.arange(C) + 0.5) / C
    r0 = np.array([L.stb_rng48_state(500 + c) for c in range(C)], dtype=np.uint64)
    bpar = np.full(cts.I, 5.0)
    a1, r1, s1 = stb.samplea_batch(a0, cts, bpar, r0, loops=2)
    a2, r2, s2 = stb.samplea_batch(a0, cts, bpar, r0, loops=2, devices=devices)
    assert np.array_equal(a1, a2) and np.array_equal(r1, r2)
    assert s2["evals"] >= s1["evals"] - 0 and s2["rounds"] > 0  # speculation fills different slots per share
    # per-chain concentrations travel with their chain
    bp = np.tile(bpar, (C, 1)) * (1 + 0.1 * np.arange(C))[:, None]
    a3, r3, _ = stb.samplea_batch(a0, cts, bp, r0, loops=1, bpar_per_chain=True)
    a4, r4, _ = stb.samplea_batch(a0, cts, bp, r0, loops=1, bpar_per_chain=True, devices=devices)
    assert np.array_equal(a3, a4) and np.array_equal(r3, r4)
    ap = a1.copy()
    ap[::4] = 0.0  # the closed-form branch of sampleb on some chains
    b1, q1, _ = stb.sampleb_batch(np.full(C, 7.0), cts, 1.1, 20.0, ap, r1, loops=1)
    b2, q2, _ = stb.sampleb_batch(np.full(C, 7.0), cts, 1.1, 20.0, ap, r1, loops=1, devices=devices)
    assert np.array_equal(b1, b2) and np.array_equal(q1, q2)


@pytest.mark.skipif("_ndev() < 2", reason="needs two GPUs (run with gpurun --gpus 2)")
def test_chains_two_devices_equal_one_device():
    cts = tsg._counts(7, 30, 40, 400)
    C = 37
    L = stb.lib()
    a0 = 0.1 + 0.8 * (np.arange(C) + 0.5) / C
    r0 = np.array([L.stb_rng48_state(900 + c) for c in range(C)], dtype=np.uint64)
    bpar = np.full(cts.I, 5.0)
    a1, r1, _ = stb.samplea_batch(a0, cts, bpar, r0, loops=1)
    a2, r2, _ = stb.samplea_batch(a0, cts, bpar, r0, loops=1, devices=[0, 1])
    assert np.array_equal(a1, a2) and np.array_equal(r1, r2)
    b1, q1, _ = stb.sampleb_batch(np.full(C, 7.0), cts, 1.1, 20.0, a1, r1, loops=1)
    b2, q2, _ = stb.sampleb_batch(np.full(C, 7.0), cts, 1.1, 20.0, a1, r1, loops=1, devices=[0, 1])
    assert np.array_equal(b1, b2) and np.array_equal(q1, q2)


def test_current_device_is_restored():
    """every entry point runs on its handle's device and gives the caller's back"""
    import torch

    if _ndev() < 2:
        pytest.skip("one device: nothing to restore")
    torch.cuda.set_device(1)
    t = stb.Table(500, 50, 500, 50, 0.5, stb.S_STABLE)  # created on device 1
    torch.cuda.set_device(0)
    t.remake(0.6)
    assert stb.lib().stb_cuda_current_device() == 0
    w = stb.SweepMulti(500, 50, [1, 0])
    w.run(np.array([0.3, 0.4, 0.5]), gather=False, sums=False, lastrow=True)
    assert stb.lib().stb_cuda_current_device() == 0
    w.free()
    t.free()


def test_last_error_is_per_thread():
    """two threads fail at the same time with different messages and each reads its own"""
    L = stb.lib()
    barrier = threading.Barrier(2)
    seen = {}

    def work(tag, N, M):
        for _ in range(50):
            assert not L.stb_sweep_create(N, M, 0)  # M > N: refused with the extent in the message
            barrier.wait()  # both have failed; neither has read yet
            seen.setdefault(tag, set()).add(L.stb_last_error().decode())
            barrier.wait()

    th = [threading.Thread(target=work, args=("a", 10, 20)), threading.Thread(target=work, args=("b", 30, 40))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert seen["a"] == {"stb_cuda_sweep_create: bad extent 10x20"}
    assert seen["b"] == {"stb_cuda_sweep_create: bad extent 30x40"}
