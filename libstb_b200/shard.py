"""Sharding of independent units (discounts of a sweep, posterior chains) over ranks, one process
per GPU.  The hot path has NO inter-GPU traffic: rank r works on units r, r+W, r+2W, ...; the only
collective is the final all_gather of the per-unit results (a few doubles per unit).  A single table
does not shard: rows are sequential and every row needs its left neighbour.

`dist` is torch.distributed (NCCL on GPUs, gloo in the CPU tests) or None for a single process."""
from __future__ import annotations

import numpy as np


def my_units(n_units: int, rank: int, world: int) -> np.ndarray:
    """Indices of the units rank `rank` owns (strided deal: keeps the ranks' loads level when the
    cost of a unit drifts with its index, as it does across a discount sweep)."""
    return np.arange(rank, n_units, world, dtype=np.int64)


def gather_units(local: np.ndarray, n_units: int, rank: int, world: int, dist=None, device="cpu") -> np.ndarray:
    """Reassemble per-unit results: `local[k]` belongs to unit rank + k*world.  Every rank gets the
    full (n_units, ...) array.  Uneven shards are padded for the collective and trimmed after it."""
    local = np.asarray(local, dtype=np.float64)
    tail = local.shape[1:]
    if world == 1 or dist is None:
        assert local.shape[0] == n_units
        return local.copy()
    import torch

    per = (n_units + world - 1) // world
    buf = torch.full((per,) + tail, float("nan"), dtype=torch.float64, device=device)
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(local).to(device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = np.empty((n_units,) + tail, dtype=np.float64)
    for r in range(world):
        idx = my_units(n_units, r, world)
        out[idx] = parts[r][: idx.shape[0]].cpu().numpy()
    return out


def sweep_sharded(N, M, a_all, n, m, rank=0, world=1, dist=None, device="cpu", sweep_factory=None):
    """Discount sweep over all ranks: each fills the tables of its own discounts and reduces them to
    the sum over the (n, m) pairs; returns (sums for ALL discounts, device ms of this rank's fills).
    `sweep_factory(N, M)` builds the per-rank engine (default: libstb_b200.Sweep, i.e. the CUDA path)."""
    a_all = np.asarray(a_all, dtype=np.float64)
    if sweep_factory is None:
        import libstb_b200 as stb

        sweep_factory = lambda N_, M_: stb.Sweep(N_, M_)  # noqa: E731
    w = sweep_factory(N, M)
    w.set_pairs(n, m)
    mine = my_units(a_all.shape[0], rank, world)
    _, sums, _ = w.run(a_all[mine], gather=False, sums=True)
    ms = w.last_fill_ms
    w.free()
    return gather_units(sums, a_all.shape[0], rank, world, dist, device), ms
