"""Sharding of independent streams across the GPUs of one box.

Streams never interact at inference (BN folded, attention within one stream's frame,
GRU rows never mix -- /root/reference/models/fastenhancer/default/model.py:270-272, 283-285), so the batch is cut
into contiguous slices, one per rank, and there is no collective inside a hop.  ``scatter_streams`` /
``gather_streams`` exist only for callers that hold the whole batch on one rank (trivial NCCL/gloo
scatter/gather of ``[B/G, samples]`` audio); state never moves.
"""
from __future__ import annotations

import typing as tp

import torch
import torch.distributed as dist


def shard_range(n_streams: int, world: int, rank: int) -> tp.Tuple[int, int]:
    """(first stream, number of streams) of ``rank``: contiguous, sizes differ by at most one, ragged tail allowed."""
    if not (0 <= rank < world) or n_streams < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(n_streams, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def scatter_streams(full: tp.Optional[torch.Tensor], n_streams: int, samples: int, src: int = 0,
                    device: tp.Union[str, torch.device] = "cpu", group=None) -> torch.Tensor:
    """Rank ``src`` holds ``full`` [n_streams, samples]; every rank gets its slice [count, samples]."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    start, count = shard_range(n_streams, world, rank)
    # ragged slices: pad to the largest shard so that scatter sees equal shapes
    biggest = shard_range(n_streams, world, 0)[1]
    buf = torch.zeros((biggest, samples), dtype=torch.float32, device=device)
    chunks = None
    if rank == src:
        chunks = []
        for r in range(world):
            s, c = shard_range(n_streams, world, r)
            piece = torch.zeros((biggest, samples), dtype=torch.float32, device=device)
            piece[:c] = full[s:s + c].to(device)
            chunks.append(piece)
    dist.scatter(buf, chunks, src=src, group=group)
    return buf[:count].clone()


def gather_streams(local: torch.Tensor, n_streams: int, dst: int = 0, group=None) -> tp.Optional[torch.Tensor]:
    """Inverse of :func:`scatter_streams`: rank ``dst`` receives [n_streams, samples]."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    biggest = shard_range(n_streams, world, 0)[1]
    buf = torch.zeros((biggest, local.shape[1]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, outs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([outs[r][:shard_range(n_streams, world, r)[1]] for r in range(world)], dim=0)
