"""The N>1 host path on CPU: world_size-2 gloo run of the sharding + metric gather (no GPU needed)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dvae_b200.shard import gather_metrics, max_over_ranks, shard_range


def test_shard_range_matches_array_split():
    for n in (0, 1, 7, 512, 4096, 4097):
        for world in (1, 2, 3, 8):
            ref = np.array_split(np.arange(n), world)
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert list(range(lo, hi)) == list(ref[r])
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    local = torch.stack([torch.arange(lo, hi, dtype=torch.float64), torch.arange(lo, hi, dtype=torch.float64) ** 2], dim=1)
    full = gather_metrics(local, n_total)
    t = max_over_ranks(1.0 + rank, "cpu")
    q.put((rank, full.numpy(), t))
    dist.destroy_process_group()


def test_gather_metrics_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_total, world = 7, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.stack([np.arange(n_total, dtype=np.float64), np.arange(n_total, dtype=np.float64) ** 2], axis=1)
    for rank, full, t in res:
        assert np.array_equal(full, expect)
        assert t == 2.0
