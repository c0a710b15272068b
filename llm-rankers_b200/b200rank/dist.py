"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the rendezvous.

The hot path shards embarrassingly — every (query, passage) prompt is independent — so there is exactly
one collective in the system: the broadcast of the device weight arena from rank 0 at load (NCCL over NVLink). Work is
split contiguously over ranks; scores return by an all-gather of small host-side arrays after the loop. No collective
runs inside the scoring loop.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced split: the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def arena_as_tensor(engine, device_index: int):
    """View the engine's device weight arena as a uint8 torch tensor (no copy) through __cuda_array_interface__."""
    import torch
    ptr, nbytes = engine.weights_blob()

    class _Arena:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Arena(), device=torch.device("cuda", device_index))


def broadcast_weights(engine, device_index: int, src: int = 0) -> None:
    """The one collective: rank `src` has loaded the weights; every other rank receives the arena bytes over NCCL."""
    import torch
    import torch.distributed as dist
    t = arena_as_tensor(engine, device_index)
    dist.broadcast(t, src=src)
    torch.cuda.synchronize(device_index)
    if dist.get_rank() != src:
        engine.mark_weights_loaded()


def all_gather_variable(local: np.ndarray, group=None) -> np.ndarray:
    """Concatenate per-rank 1-D float arrays of different lengths in rank order (works on gloo and nccl)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes) if sizes else 0
    buf = torch.zeros((m,), dtype=torch.float32, device=dev)
    buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float32)).to(dev)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return np.concatenate([p[:s].cpu().numpy() for p, s in zip(parts, sizes)]) if sizes else np.zeros((0,), np.float32)


def score_sharded(score_fn, rows: Sequence, group=None) -> np.ndarray:
    """Score rows[lo:hi] on this rank with score_fn(list_of_rows) -> 1-D array, then all-gather in rank order.
    The result is identical on every rank and identical to score_fn(rows) on one rank (no cross-row arithmetic)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(len(rows), rank, world)
    local = np.asarray(score_fn(list(rows[lo:hi])), dtype=np.float32) if hi > lo else np.zeros((0,), np.float32)
    return all_gather_variable(local, group)


def gather_lists(local: List, group=None) -> List:
    """Concatenate per-rank Python lists in rank order on every rank (picklable items; gloo or nccl)."""
    import torch.distributed as dist
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, list(local), group=group)
    return [x for part in parts for x in part]
