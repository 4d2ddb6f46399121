"""Utterance sharding over the GPUs of one box and the only collective of the path: gathering per-utterance metrics.

The reference splits the file list with ``np.array_split`` over ``2 x nb_devices`` spawned workers
(``scripts/evaluate_ntcd_M1.py:249-259``).  Here: one process per GPU (torchrun), rank r takes a contiguous block of the
utterance index list; no data-path collective exists (utterances are independent end to end), only a gather of a few
bytes per utterance and a MAX-reduce of the elapsed time.  Works with the ``nccl`` (GPU) and ``gloo`` (CPU tests) backends.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block ``[lo, hi)`` of rank ``rank``: sizes differ by at most one, like ``np.array_split``."""
    if world < 1 or not 0 <= rank < world or n_items < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_metrics(local: torch.Tensor, n_total: int):
    """All-gather ragged per-utterance metric rows ``[n_local][m]`` into ``[n_total][m]`` on every rank (index order)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    m = local.shape[1]
    cap = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    buf = torch.zeros((cap, m), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        parts.append(out[r][: hi - lo])
    return torch.cat(parts, dim=0)


def max_over_ranks(value: float, device) -> float:
    """MAX all-reduce of a scalar (elapsed device time)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
