"""Batch sharding of the LPSR forward across the GPUs of one box.

Every crop is independent (the only cross-pixel coupling, CSAR's global average pool, is per sample: reference
lpsr.py:124), so the path shards by batch with replicated weights and NO per-layer collective.  The only exchange is
the final gather of the fp32 ``[B/G,1,H,W]`` outputs (NCCL all-gather over NVLink/NVSwitch; gloo on CPU in tests).
One process per GPU, launched with torchrun.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of a batch: rank g gets crops [g*B/G, (g+1)*B/G), the first B % G ranks one extra."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_outputs(y_local: torch.Tensor, batch: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather the per-rank output shards back into the full ``[batch, ...]`` tensor on every rank.

    Shards may be ragged (batch % world != 0) or empty: each rank pads its shard to the largest shard so that one
    ``all_gather_into_tensor`` (a single NCCL collective) suffices, then the padding rows are dropped."""
    world = dist.get_world_size(group)
    if world == 1:
        return y_local
    sizes = [shard_bounds(batch, world, r)[1] - shard_bounds(batch, world, r)[0] for r in range(world)]
    mx = max(sizes)
    tail = tuple(y_local.shape[1:])
    send = y_local
    if y_local.shape[0] != mx:
        send = y_local.new_zeros((mx,) + tail)
        send[: y_local.shape[0]] = y_local
    recv = y_local.new_empty((world * mx,) + tail)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if all(s == mx for s in sizes):
        return recv
    return torch.cat([recv[r * mx: r * mx + sizes[r]] for r in range(world)], dim=0)


def forward_sharded(forward: Callable[[torch.Tensor], torch.Tensor], x: torch.Tensor, gather: bool = True,
                    group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Run ``forward`` on this rank's shard of the full batch ``x`` (every rank holds or can slice the same ``x``) and
    optionally gather the outputs.  ``forward`` is ``LPSR.__call__`` on the GPU box; tests pass a CPU stand-in."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(x.shape[0], world, rank)
    y_local = forward(x[lo:hi])
    if not gather or world == 1:
        return y_local
    return gather_outputs(y_local, x.shape[0], group)
