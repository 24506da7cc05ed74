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

    def _reduce(self, t: torch.Tensor, async_op: bool = False):
        if dist.get_backend() == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op)
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=async_op)     # gloo has no AVG
        if not async_op:
            t.div_(world())
        return w

    def all_reduce_mean(self) -> None:
        if world() == 1:
            return
        self._reduce(self.flat)

    # -- overlap: reduce the gradients that are ready EARLY in the backward pass (the big head
    #    matrices, 97 % of the bytes) while the backward of the convolution stack still runs ----
    def enable_overlap(self, early_params) -> None:
        """`early_params`: parameters whose gradients are complete first (must be a prefix or any
        subset; they are moved to the front of the bucket).  Their all-reduce is launched from a
        post-accumulate hook on the LAST of them; call finish() after backward."""
        early = [p for p in self.params if any(p is q for q in early_params)]
        late = [p for p in self.params if not any(p is q for q in early_params)]
        self.params = early + late
        off = 0
        for p in self.params:                       # re-lay the bucket: early params first
            p.grad = self.flat[off:off + p.numel()].as_strided(p.shape, p.stride())
            off += p.numel()
        n_early = sum(p.numel() for p in early)
        self._early, self._late = self.flat[:n_early], self.flat[n_early:]
        self._pending, self._work = len(early), None
        self._ready = set()                 # ids of the early params whose gradient is complete this step
        self._sink_attached = False
        gloo = dist.is_initialized() and dist.get_backend() != "nccl"

        def mark(p):
            self._ready.add(id(p))          # a set: a parameter reported twice still counts once
            if len(self._ready) == self._pending and self._work is None and world() > 1:
                self._work = self._reduce(self._early, async_op=True)
                self._early_needs_div = gloo

        def autograd_hook(p):
            # torch >= 2.x runs post-accumulate-grad hooks even when a custom Function returned None for the
            # parameter (nothing was accumulated).  With a WeightGradSink attached the early gradients never come
            # through autograd — the sink reports them AFTER its own accumulation — so this path must stay silent,
            # or the all-reduce would start before the late-arriving gradients exist.
            if not self._sink_attached:
                mark(p)

        self._early_hook, self._early_params = mark, early
        for p in early:
            p.register_post_accumulate_grad_hook(autograd_hook)

    def attach_sink(self, sink) -> None:
        """Weight gradients that bypass autograd's AccumulateGrad (linear.WeightGradSink) report here instead:
        the early all-reduce is then launched from the sink's side stream as soon as the last early gradient
        has been accumulated there."""
        self._sink_attached = True
        for p in getattr(self, "_early_params", []):
            sink.on_ready[id(p)] = self._early_hook

    def finish(self) -> None:
        """After backward: reduce the late (small) part, then wait for the early part."""
        if world() == 1:
            self._ready.clear() if hasattr(self, "_ready") else None
            return
        if self._late.numel():
            self._reduce(self._late)
        if self._work is not None:
            self._work.wait()
            if getattr(self, "_early_needs_div", False):
                self._early.div_(world())
        else:                                       # hooks did not fire (no early grads): reduce now
            self._reduce(self._early)
        self._work = None
        self._ready.clear()


def all_reduce_max_(x: torch.Tensor, async_op: bool = False):
    """In-place MAX all-reduce of the shard-local max IS weight.  async_op=True returns the work
    handle (None at world size 1) so the tiny collective can hide behind the learner step."""
    if world() > 1:
        w = dist.all_reduce(x, op=dist.ReduceOp.MAX, async_op=async_op)
        return w if async_op else x
    return None if async_op else x


def shard_slots(total_slots: int, rank: int, world_size: int) -> range:
    """Contiguous slot range owned by `rank` (SURVEY §8e 'contiguous ranges of N/G')."""
    per = (total_slots + world_size - 1) // world_size
    return range(min(rank * per, total_slots), min((rank + 1) * per, total_slots))
