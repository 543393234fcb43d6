"""The multi-GPU plumbing on CPU: world_size 2 (and 3) over gloo.  Units are dealt to ranks, worked
on independently, and reassembled by one all_gather.  The per-unit work here is the CPU oracle (a
stand-in for the CUDA engine, which needs a GPU) so that the test checks real numbers end to end."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from libstb_b200 import shard
from tests import harness


class OracleSweep:
    """Same interface as libstb_b200.Sweep, computed by the oracle (test stand-in)."""

    def __init__(self, N, M):
        self.N, self.M, self.last_fill_ms = N, M, 0.0

    def set_pairs(self, n, m):
        self.n, self.m = np.asarray(n), np.asarray(m)

    def run(self, a, gather=False, sums=True):
        out = []
        for aj in a:
            S, _ = harness.oracle_tables(self.N, self.M, float(aj), want_V=False)
            out.append(S[self.n - 1, self.m - 1].sum())
        return None, np.array(out), None

    def free(self):
        pass


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_units, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. reassembly of uneven shards, vector-valued units
        mine = shard.my_units(n_units, rank, world)
        local = np.stack([mine * 1.5, mine ** 2.0], axis=1) if mine.size else np.empty((0, 2))
        full = shard.gather_units(local, n_units, rank, world, dist)
        # 2. a sharded discount sweep with the oracle engine
        a_all = (np.arange(n_units) + 0.5) / n_units
        rng = np.random.default_rng(0)
        n = rng.integers(2, 120, size=200)
        m = np.minimum(rng.integers(1, 30, size=200), n)
        sums, _ = shard.sweep_sharded(120, 30, a_all, n, m, rank, world, dist, sweep_factory=OracleSweep)
        q.put((rank, full, sums))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_units", [(2, 7), (2, 8), (3, 5)])
def test_sharded_units_reassemble(world, n_units):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(n_units)
    expect = np.stack([idx * 1.5, idx ** 2.0], axis=1)
    # single-process answer of the sweep
    a_all = (idx + 0.5) / n_units
    rng = np.random.default_rng(0)
    n = rng.integers(2, 120, size=200)
    m = np.minimum(rng.integers(1, 30, size=200), n)
    ref, _ = shard.sweep_sharded(120, 30, a_all, n, m, sweep_factory=OracleSweep)
    for rank, full, sums in got:
        assert np.array_equal(full, expect), rank
        assert np.array_equal(sums, ref), rank  # every rank ends with every unit's result


def test_my_units_partition():
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 4096):
            allidx = np.concatenate([shard.my_units(n, r, world) for r in range(world)])
            assert sorted(allidx.tolist()) == list(range(n))
            sizes = [shard.my_units(n, r, world).size for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
