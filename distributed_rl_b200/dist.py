"""Replay-sharded data parallelism helpers.

The reference has a single learner process and no collective of any kind
(SURVEY §2 row 17); this is new work.  One process per GPU, slots sharded, every
rank samples locally; the ONLY inter-GPU traffic per step is
  * one all-reduce (mean) of the gradients, kept in one flat bucket, and
  * one all-reduce (max) of the shard's max IS weight ("priority-max reduction").
The helpers are backend-agnostic so the logic is covered by gloo tests on CPU.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class FlatGradBucket:
    """All gradients of `params` as views into ONE contiguous buffer -> one collective per step."""

    def __init__(self, params, device=None):
        self.params = list(params)
        device = device or self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), device=device, dtype=self.params[0].dtype)
        off = 0
        for p in self.params:
            # same element order as the parameter (channels_last conv weights keep their strides), so that
            # elementwise fused optimizers can walk param and grad storage together
            p.grad = self.flat[off:off + p.numel()].as_strided(p.shape, p.stride())
            off += p.numel()

    def all_reduce_mean(self) -> None:
        if world() == 1:
            return
        if dist.get_backend() == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:                                   # gloo has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(world())


def all_reduce_max_(x: torch.Tensor) -> torch.Tensor:
    """In-place MAX all-reduce of the shard-local max IS weight."""
    if world() > 1:
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
    return x


def shard_slots(total_slots: int, rank: int, world_size: int) -> range:
    """Contiguous slot range owned by `rank` (SURVEY §8e 'contiguous ranges of N/G')."""
    per = (total_slots + world_size - 1) // world_size
    return range(min(rank * per, total_slots), min((rank + 1) * per, total_slots))
