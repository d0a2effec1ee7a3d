"""Sharding a batch of images over the GPUs of one box: one process per GPU, no data-path collective.

Images are independent units, so rank g of G simply takes the contiguous range
``[g*B//G, (g+1)*B//G)`` of a batch of B images and runs the single-GPU pipeline on it.  The only
communication is the host-side gather of the finished flat records to rank 0 (and, in bench.py, a barrier
and a max-reduction of the device time), done through ``torch.distributed`` -- NCCL on the GPU box, gloo in
the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced, order-preserving partition: sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return n_items * rank // world, n_items * (rank + 1) // world


def shard_sizes(n_items: int, world: int) -> list[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def gather_records(local: np.ndarray, n_total: int, rank: int, world: int, group=None, device=None):
    """Gathers per-rank record blocks [n_local, record_bytes] (uint8) to rank 0 in rank order.

    Returns the [n_total, record_bytes] array on rank 0 and None elsewhere.  Uses torch.distributed when
    world > 1 (tensors staged on `device` for NCCL, on the host for gloo)."""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    sizes = shard_sizes(n_total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} records, expected {sizes[rank]}")
    rb = local.shape[1]
    pad = max(sizes)
    buf = torch.zeros((pad, rb), dtype=torch.uint8)
    buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local))
    if device is not None:
        buf = buf.to(device)
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0, group=group)
    if rank != 0:
        return None
    return np.concatenate([o[: sizes[r]].cpu().numpy() for r, o in enumerate(out)], axis=0)
