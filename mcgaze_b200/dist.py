"""Multi-GPU sharding of clips.

Clips are independent units, so the forward needs no collective: rank r of W takes clip indices
r, r+W, r+2W, ... (the reference's DistributedSampler(shuffle=False) order,
mmdet/datasets/samplers/distributed_sampler.py:53) and the per-rank results are gathered ONCE at
the end and re-interleaved like mmdet/apis/test.py:170-173,204-206.  Backend: NCCL on GPUs,
gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Indices of the items rank `rank` processes (round-robin, no padding)."""
    return list(range(rank, n_items, world))


def padded_shard(n_items: int, rank: int, world: int) -> List[int]:
    """DistributedSampler semantics: every rank gets ceil(n/W) items, wrapping around at the end."""
    per = -(-n_items // world)
    idx = list(range(n_items)) + [i % n_items for i in range(per * world - n_items)]
    return idx[rank:per * world:world]


def interleave(parts: Sequence[np.ndarray], n_items: int) -> np.ndarray:
    """Inverse of round-robin sharding: parts[r][j] is item r + j*W.  Drops the wrap-around padding."""
    world = len(parts)
    out = np.empty((n_items,) + tuple(parts[0].shape[1:]), dtype=parts[0].dtype)
    for r, p in enumerate(parts):
        k = len(range(r, n_items, world))
        out[r::world] = p[:k]
    return out


def gather_results(local, n_items: int, group=None):
    """All-gather per-rank result tensors [n_local_padded, ...] and return them in item order on every
    rank.  `local` must have the same shape on all ranks (use padded_shard)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    bufs = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(bufs, local.contiguous(), group=group)
    parts = [b.cpu().numpy() for b in bufs]
    return torch.from_numpy(interleave(parts, n_items))
