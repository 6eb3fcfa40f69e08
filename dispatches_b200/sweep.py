"""Scenario sharding across the GPUs of one box (one process per GPU, torch.distributed).

The LPs of a sweep are independent: rank r solves the interleaved shard  i = r (mod world)
(interleaving balances the per-LP iteration counts when scenario difficulty is ordered, e.g. by season) and
ONE all_gather at the end assembles objective / status / iteration arrays on every rank.  This replaces the
reference's ``multiprocessing.Pool(35).starmap(run_design, ...)`` fan-out (run_pricetaker_wind_PEM.py:106-107).
"""
from __future__ import annotations

import numpy as np


def shard_indices(N, rank, world):
    return np.arange(rank, N, world)


def solve_sharded(solve_fn, N, group=None):
    """solve_fn(idx) -> dict of equal-length 1-D torch tensors for the global problem indices ``idx``.
    Returns the same dict with full-length [N] tensors, identical on every rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return solve_fn(np.arange(N))
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    idx = shard_indices(N, rank, world)
    per = (N + world - 1) // world
    local = solve_fn(idx)
    out = {}
    for k, v in local.items():
        pad = torch.zeros(per, dtype=v.dtype, device=v.device)
        pad[: v.numel()] = v
        gathered = torch.empty(per * world, dtype=v.dtype, device=v.device)
        dist.all_gather_into_tensor(gathered, pad, group=group) if v.is_cuda else \
            dist.all_gather(list(gathered.view(world, per).unbind(0)), pad, group=group)
        # shard r holds the indices r, r+world, r+2*world, ...: transposing [world, per] interleaves them back
        full = gathered.view(world, per).transpose(0, 1).reshape(-1)[:N].contiguous()
        out[k] = full
    return out
