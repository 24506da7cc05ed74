"""Asynchronous parameter publication.

The reference serialises `{k: v.cpu()}` -> pickle -> Redis on the learner's own thread every 50
steps (every step for IMPALA) — `APE_X/Learner.py:207-216`, `IMPALA/Learner.py:286-287`; each
`.cpu()` is a synchronous device->host copy that stalls the step.  Here the weights are snapshot
into ONE flat device buffer on the learner stream (a few microseconds, graph-friendly), copied to
pinned host memory on a side stream, and pickled + SET by whoever calls `poll()` once the copy's
event has fired — the learner stream never waits for PCIe.  Keys and payload format are the
reference's (`state_dict` / `target_state_dict` / `count`; a dict of CPU tensors), so unmodified
actors (`APE_X/Player.py:113-133`) keep working.
"""
from __future__ import annotations

import pickle

import torch


class ParamPublisher:
    def __init__(self, model: torch.nn.Module, connect, key: str | None = "state_dict",
                 count_key: str | None = "count", wrap=None, on_ready=None):
        """`wrap(sd)`: payload put under `key` (IMPALA publishes the 1-tuple `(sd,)`, IMPALA/Learner.py:268-272);
        `on_ready(sd, step)`: extra consumer of a landed snapshot (e.g. the checkpoint writer)."""
        self.model, self.connect, self.key, self.count_key = model, connect, key, count_key
        self.wrap, self.on_ready = wrap, on_ready
        sd = model.state_dict()
        self.names = list(sd.keys())
        self.shapes = [tuple(v.shape) for v in sd.values()]
        self.numels = [v.numel() for v in sd.values()]
        dev = next(model.parameters()).device
        n = sum(self.numels)
        self.flat_dev = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_host = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.stream = torch.cuda.Stream(dev)
        self.snap_done = torch.cuda.Event()
        self.copy_done = [torch.cuda.Event(), torch.cuda.Event()]
        self.pending = None          # (buffer index, step) of a copy in flight
        self.slot = 0
        self.published = 0

    def snapshot(self, step: int) -> None:
        """Enqueue: weights -> flat device buffer (learner stream) -> pinned host (side stream)."""
        self.poll()
        if self.pending is not None:
            return                   # previous publication still in flight: skip this one (actors lag by design)
        views = torch.split(self.flat_dev, self.numels)
        srcs = [v.detach().reshape(-1) if v.is_contiguous() else v.detach().contiguous().reshape(-1)
                for v in self.model.state_dict().values()]
        torch._foreach_copy_(list(views), srcs)
        self.snap_done.record(torch.cuda.current_stream(self.flat_dev.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.snap_done)
            self.flat_host[self.slot].copy_(self.flat_dev, non_blocking=True)
            self.copy_done[self.slot].record(self.stream)
        self.pending = (self.slot, step)
        self.slot ^= 1

    def poll(self, block: bool = False) -> bool:
        """If the pending copy has landed, pickle it and SET it (same payload as the reference)."""
        if self.pending is None:
            return False
        slot, step = self.pending
        if block:
            self.copy_done[slot].synchronize()
        elif not self.copy_done[slot].query():
            return False
        parts = torch.split(self.flat_host[slot], self.numels)
        sd = {k: p.view(s).clone() for k, p, s in zip(self.names, parts, self.shapes)}
        if self.connect is not None and self.key:
            self.connect.set(self.key, pickle.dumps(self.wrap(sd) if self.wrap else sd))
            if self.count_key:
                self.connect.set(self.count_key, pickle.dumps(step))
        if self.on_ready is not None:
            self.on_ready(sd, step)
        self.last = sd
        self.pending = None
        self.published += 1
        return True
