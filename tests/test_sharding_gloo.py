"""world_size-2 gloo test of the scenario sharding + final all_gather (dispatches_b200/sweep.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dispatches_b200 import sweep


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seen = []

    def fake_solve(idx):          # stands in for the CUDA solver: a deterministic function of the global index
        seen.append(np.asarray(idx))
        return dict(obj=torch.tensor(idx, dtype=torch.float64) * 1.5 + 1.0,
                    status=torch.zeros(len(idx), dtype=torch.int32),
                    iters=torch.tensor(idx % 7, dtype=torch.int32))

    out = sweep.solve_sharded(fake_solve, N)
    q.put((rank, out["obj"].numpy(), out["iters"].numpy(), seen[0]))
    dist.destroy_process_group()


def test_interleaved_shards_cover_everything():
    for N in (0, 1, 7, 10):
        for world in (1, 2, 4, 8):
            parts = [sweep.shard_indices(N, r, world) for r in range(world)]
            assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(N))


def test_two_rank_gather_is_complete_and_identical():
    N, world = 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    for rank, obj, iters, idx in res:
        assert np.array_equal(idx, np.arange(rank, N, world))
        assert np.allclose(obj, np.arange(N) * 1.5 + 1.0)
        assert np.array_equal(iters, np.arange(N) % 7)
