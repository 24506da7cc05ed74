"""Ape-X learner side: `Replay` and `Learner` with the reference's surface
(APE_X/ReplayMemory.py:19-167, APE_X/Learner.py:20-272) on top of the
HBM-resident replay and the fused kernels.

What changed relative to the reference, and why it is still a drop-in:
  * replay contents + priorities live in HBM (DeviceReplay); `Replay.sample()`
    returns the same 7-list `[s, a, r, s', done, w, idx]`, but as CUDA tensors,
    so `Learner.train` never copies frames host->device (the reference converts
    to fp32 on the CPU and ships 4x the bytes, APE_X/Learner.py:61-67).
  * `Learner.train` has no device->host sync: double-DQN argmax, target, clipped
    TD, priority, loss and dLoss/dQ come from one kernel (b2rl_apex_target) and
    the network backward is seeded with dLoss/dQ directly.
  * priorities are written back immediately (`Replay.update` enqueues the tree
    update on the stream) instead of after >1000 pending entries
    (APE_X/ReplayMemory.py:147-150); eviction is the ring's FIFO overwrite, so
    the `lock` flag handshake (:151-160, APE_X/Learner.py:189-197) is a no-op.
  * `Learner.fused_step()` runs sample -> gather -> 3 forwards -> target ->
    backward -> RMSprop -> priority write-back as one CUDA graph.
"""
from __future__ import annotations

import pickle
import threading
import time
from dataclasses import dataclass, field

import numpy as np
import torch

from . import replay as R
from .agent import GraphAgent


@dataclass
class ApexConfig:
    """The globals the reference lifts out of cfg/ape_x.json (configuration.py:39-98)."""
    BATCHSIZE: int = 32
    ACTION_SIZE: int = 6
    ALPHA: float = 0.6
    BETA: float = 0.4
    GAMMA: float = 0.99
    UNROLL_STEP: int = 3
    REPLAY_MEMORY_LEN: int = 100000
    BUFFER_SIZE: int = 50000
    TARGET_FREQUENCY: int = 2500
    LEARNER_DEVICE: str = "cuda:0"
    REDIS_SERVER: str = "localhost"
    LOG_W: str | None = None        # ./weight/<ALG>/<time> (configuration.py:101-109); None: no checkpoints
    OPTIM_INFO: dict = field(default_factory=lambda: {
        "name": "rmsprop", "lr": 0.0000625, "eps": 1.5e-7, "decay": 0, "alpha": 0.95, "momentum": 0,
        "centered": True})
    MODEL: dict = field(default_factory=lambda: default_apex_model())
    CHANNELS_LAST: bool = True      # NHWC activations/weights: cuDNN's TF32 kernels skip their layout transposes
    FUSED_CONV1: bool = True        # gather + conv_1 on the tcgen05 tensor cores (csrc/conv1.cu) in fused_step
    FUSED_OPTIM: bool = True        # RMSprop + zero_grad + grad-norm in one launch (csrc/optim.cu)
    EARLY_HEAD_UPDATE: bool = True  # RMSprop of the dense heads as soon as their gradients are final (no clipping in :123-138)
    CUDNN_BENCHMARK: bool = True    # let cuDNN time its conv_2/conv_3 algorithms once (no precision change)
    DEFERRED_WGRAD: bool = True      # weight gradients on a side stream, off the critical path of backward
    PARALLEL_FORWARDS: bool = True   # the three forward passes of a step on three streams (fork/join inside the graph)
    FUSED_DUELING_TAIL: bool = True  # heads' second layers + dueling combine in one kernel (csrc/dueling.cu)
    DENSE_3XTF32: bool = True       # dense heads as 3xTF32 tcgen05 GEMMs at fp32 accuracy (csrc/gemm.cu)
    BATCHED_ONLINE: bool = True     # the online net's two passes (s with grad, s' without) as ONE B = 2*BATCHSIZE call

    @staticmethod
    def from_configuration():
        import configuration as C  # the drop-in module (dropin/configuration.py) or the user's own
        kw = {k: getattr(C, k) for k in ("BATCHSIZE", "ACTION_SIZE", "ALPHA", "BETA", "GAMMA", "UNROLL_STEP",
                                         "REPLAY_MEMORY_LEN", "BUFFER_SIZE", "TARGET_FREQUENCY",
                                         "LEARNER_DEVICE", "REDIS_SERVER", "OPTIM_INFO", "MODEL")}
        kw["LOG_W"] = getattr(C, "LOG_W", None)
        return ApexConfig(**kw)


def default_apex_model() -> dict:
    """The dueling DQN of cfg/ape_x.json:37-88 (values are configuration, not code)."""
    return {
        "module00": {"netCat": "CNN2D", "iSize": 4, "nLayer": 4, "fSize": [8, 4, 3, -1], "nUnit": [32, 64, 64],
                     "padding": [0, 0, 0], "stride": [4, 2, 1], "act": ["relu", "relu", "relu"],
                     "BN": [False] * 4, "linear": True, "input": [0], "prior": 0},
        "module02": {"netCat": "MLP", "iSize": 3136, "nLayer": 2, "fSize": [512, 6], "act": ["relu", "linear"],
                     "BN": [False] * 3, "prior": 1, "prevNodeNames": ["module00"]},
        "module02_1": {"netCat": "MLP", "iSize": 3136, "nLayer": 2, "fSize": [512, 1], "act": ["relu", "linear"],
                       "BN": [False] * 3, "prior": 1, "prevNodeNames": ["module00"]},
        "module03": {"netCat": "Add", "prior": 2, "prevNodeNames": ["module02", "module02_1"]},
        "module03_1": {"netCat": "Mean", "prior": 2, "prevNodeNames": ["module02"]},
        "module04": {"netCat": "Substract", "prior": 3, "prevNodeNames": ["module03", "module03_1"],
                     "output": True},
    }


def make_optimizer(info: dict, params, capturable: bool = True):
    """baseline/utils.py getOptim (:78-132) for the optimisers the shipped configs name."""
    name = info["name"]
    lr, decay, eps = info["lr"], info.get("decay", 0), info.get("eps", 1e-5)
    if name == "rmsprop":
        return torch.optim.RMSprop(params, lr=lr, weight_decay=decay, eps=eps, momentum=info.get("momentum", 0),
                                   alpha=info.get("alpha", 0.99), centered=info.get("centered", False),
                                   capturable=capturable, foreach=True)
    if name == "adam":
        return torch.optim.Adam(params, lr=lr, weight_decay=decay, eps=eps,
                                betas=(info.get("beta1", 0.9), info.get("beta2", 0.99)),
                                capturable=capturable, foreach=True)
    if name == "sgd":
        return torch.optim.SGD(params, lr=lr, weight_decay=decay, momentum=info.get("momentum", 0))
    raise ValueError(f"unknown optimizer {name!r}")


class _MemoryView:
    """What the learner reads from `Replay.memory` (APE_X/Learner.py:143,241):
    len() and .max_weight (baseline/PER.py:80-81,129-133)."""

    def __init__(self, dev_replay: R.DeviceReplay, beta: float):
        self._r, self._beta = dev_replay, beta

    def __len__(self):
        return len(self._r)

    @property
    def max_weight(self) -> float:
        return float(self._r.stats(self._beta)[2].item())


class Replay(threading.Thread):
    """APE_X/ReplayMemory.py Replay (:19-167): same methods and attributes."""

    def __init__(self, cfg: ApexConfig | None = None, connect=None):
        super().__init__(daemon=True)
        self.cfg = cfg or ApexConfig.from_configuration()
        self.device = torch.device(self.cfg.LEARNER_DEVICE)
        self.store = R.DeviceReplay(self.cfg.REPLAY_MEMORY_LEN, R.APEX_FIELDS, self.device)
        self.memory = _MemoryView(self.store, self.cfg.BETA)
        self.connect = connect
        self.cond = False
        self.lock = False          # eviction handshake flag: kept for API compatibility, unused
        self.deque = []            # pre-assembled minibatches (filled on demand)
        self.total_frame = 0
        self._lock = threading.Lock()
        self._stop_evt = threading.Event()     # NOT `_stop`: that name is threading.Thread's own method

    # -- ingest: records are [s, a, R_n, s', done, prio] pickled by the actors ----
    def push_records(self, blobs) -> None:
        """PER.push (baseline/PER.py:69-75) for a list of pickled actor records
        (APE_X/Player.py:252-261): decoded once on the host, then one batched
        H2D copy + fused leaf write / path refresh."""
        if not blobs:
            return
        from .wire import decode_apex
        recs = [pickle.loads(b) for b in blobs]
        n = len(recs)
        st = self._staging(n)        # pinned, on the GPU's NUMA node: the H2D copy is a straight DMA
        decode_apex(recs, {k: st[k][:n].numpy() for k in ("s", "ns", "a", "r", "d", "p")})
        with self._lock:
            self.store.push([st[k][:n] for k in ("s", "ns", "a", "r", "d")], st["p"][:n])
            st["event"].record(torch.cuda.current_stream(self.device))
        self.total_frame += n

    def _staging(self, n: int) -> dict:
        """One of two pinned staging sets (alternating), grown on demand; reused only after the copy that
        last read it has completed."""
        from .hostmem import pinned_empty
        if not hasattr(self, "_stages"):
            self._stages, self._stage_i = [None, None], 0
        self._stage_i ^= 1
        st = self._stages[self._stage_i]
        if st is not None:
            st["event"].synchronize()
        if st is None or st["cap"] < n:
            cap = max(n, 2 * (st["cap"] if st else 0), 64)
            shape = tuple(self.store.fields[0].shape)
            st = {"cap": cap, "event": torch.cuda.Event(),
                  "s": pinned_empty((cap, *shape), torch.uint8, self.device),
                  "ns": pinned_empty((cap, *shape), torch.uint8, self.device),
                  "a": pinned_empty((cap,), torch.int32, self.device),
                  "r": pinned_empty((cap,), torch.float32, self.device),
                  "d": pinned_empty((cap,), torch.uint8, self.device),
                  "p": pinned_empty((cap,), torch.float32, self.device)}
            self._stages[self._stage_i] = st
        return st

    def push_arrays(self, s, ns, a, r, d, p) -> None:
        """Same ingest for already-decoded arrays (host pinned or device)."""
        with self._lock:
            self.store.push([s, ns, a, r, d], p)
        self.total_frame += int(torch.as_tensor(p).numel())

    def begin_ingest(self, s, ns, a, r, d) -> None:
        """Pipelined ingest, phase 1: retire the slots about to be overwritten and start the
        host->device copy on the ingest stream (overlaps the learner step in flight)."""
        with self._lock:
            self.store.push_begin([s, ns, a, r, d], int(torch.as_tensor(a).numel()))

    def commit_ingest(self, p) -> None:
        """Phase 2: wait for the copy, publish the new priorities (records become sampleable)."""
        with self._lock:
            self.store.push_commit(p)
        self.total_frame += int(torch.as_tensor(p).numel())

    def stop(self) -> None:
        """Ask the ingest thread to leave its loop (the reference's daemon thread can only die with the process)."""
        self._stop_evt.set()

    def ingest(self, s, ns, a, r, d, p) -> None:
        """Steady-state ingest, one call per learner iteration (b2rl_replay_ingest_pipelined): the batch handed
        over by the previous call becomes sampleable, this one's host->device copy starts on the library's copy
        stream and overlaps the learner step that follows.  All arguments pinned host (or device) tensors."""
        with self._lock:
            self.store.ingest_pipelined([s, ns, a, r, d], p)
        self.total_frame += int(p.numel())

    def run(self):
        """Poll the actors' Redis list like APE_X/ReplayMemory.py:118-161: drain `experience`, push, honour the
        learner's eviction request (`lock`, :151-160).  Minibatches are assembled on demand by sample()."""
        if self.connect is None:
            return
        from .wire import drain
        while not self._stop_evt.is_set():
            data = drain(self.connect, "experience")
            if data:
                self.push_records(data)
                self.cond = len(self.store) > self.cfg.BUFFER_SIZE
            if self.lock:
                self._evict_on_request()
            if not data:
                time.sleep(0.002)

    def _evict_on_request(self) -> None:
        """The `lock` handshake (APE_X/ReplayMemory.py:151-160, APE_X/Learner.py:189-197): once the memory is full,
        drop queued minibatches and trim to REPLAY_MEMORY_LEN (PER.remove_to_fit, baseline/PER.py:118-127).  The
        ring already overwrites its oldest slot on push, so there is normally nothing to trim."""
        if len(self.store) >= self.cfg.REPLAY_MEMORY_LEN:
            with self._lock:
                self.deque.clear()
                over = len(self.store) - self.cfg.REPLAY_MEMORY_LEN
                if over > 0:
                    self.store.evict(over)
        self.lock = False

    # -- sampling -------------------------------------------------------------------
    def buffer(self, m: int = 1) -> None:
        """Replay.buffer (:61-116): sample m*BATCHSIZE, IS weights, assemble minibatches."""
        B = self.cfg.BATCHSIZE
        with self._lock:
            idx, _, w = self.store.sample(B * m, beta=self.cfg.BETA)
            batch = self.store.gather(idx)
        for k in range(m):
            sl = slice(k * B, (k + 1) * B)
            self.deque.append([batch["state"][sl], batch["action"][sl], batch["reward"][sl],
                               batch["next_state"][sl], batch["done"][sl], w[sl], idx[sl]])

    def sample(self):
        if len(self.deque) == 0:
            if len(self.store) <= self.cfg.BUFFER_SIZE:
                return False
            self.buffer(1)
        return self.deque.pop(0)

    # -- priority write-back ----------------------------------------------------------
    def update(self, idx, vals) -> None:
        """Replay.update (:43-47) + _update (:49-59) -> PER.update: applied at once."""
        if isinstance(idx, (list, tuple)):
            idx = torch.stack([torch.as_tensor(i) for i in idx]) if len(idx) and torch.is_tensor(idx[0]) \
                else torch.as_tensor(np.asarray(idx, np.int64))
        vals = torch.as_tensor(vals)
        with self._lock:
            self.store.update(idx.to(self.device), vals.to(self.device))

    def _update(self):
        """Replay._update (:49-59) applies the pending write-backs; here update() has already enqueued them on
        the stream, so all that is left is to make them visible to the host."""
        torch.cuda.current_stream(self.device).synchronize()


class _Conv1Gathered(torch.autograd.Function):
    """conv_1 over rows `idx` of a uint8 frame table (a replay field or an explicit batch).
    Forward: fused gather+conv on the tensor cores (or a precomputed output of the same kernel).
    Backward: only dL/dW is needed (the input is data); cuDNN computes it from a gathered fp32
    copy of the same rows — the one place the sampled frames are staged — or, with `fused_wgrad`
    (default), libb2rl's fused gather + wgrad kernel computes it from the uint8 rows directly."""

    fused_wgrad = True

    @staticmethod
    def forward(ctx, weight, frames, idx, pack, mem_format, store=None, y_pre=None, relu=False):
        """relu=True: the kernel's epilogue applies the ReLU that follows conv_1 and backward applies its mask
        inside the wgrad kernel (the caller must then skip the network's own ReLU: forward_from_conv1(y, True))."""
        ctx.frames, ctx.store, ctx.mem_format, ctx.wshape = frames, store, mem_format, weight.shape
        ctx.weight_param = weight
        ctx.has_idx, ctx.relu = idx is not None, bool(relu)
        idx_t = idx if idx is not None else torch.empty(0, dtype=torch.int64, device=frames.device)
        if y_pre is not None:
            y = y_pre.view_as(y_pre)
        else:
            y = R.conv1_fused(frames, idx, pack, relu=bool(relu))[0]
        if relu:
            ctx.save_for_backward(idx_t, y)
        else:
            ctx.save_for_backward(idx_t)
        return y

    @staticmethod
    def backward(ctx, gy):
        idx = ctx.saved_tensors[0]
        y = ctx.saved_tensors[1] if ctx.relu else None
        if _Conv1Gathered.fused_wgrad:
            # fused gather + wgrad on the tensor cores: the sampled rows are never staged (csrc/conv1_wgrad.cu)
            w = ctx.weight_param
            from . import linear as _lin
            if _lin._SINK is not None and w.grad is not None and w.grad.is_contiguous():
                # deferred-gradient mode (grads pre-allocated, zeroed by the optimizer): the kernel's reduction adds
                # straight into .grad — no temporary, no AccumulateGrad launch at the very end of backward
                R.conv1_wgrad(ctx.frames, idx if ctx.has_idx else None, gy, out=w.grad, accumulate=True, relu_y=y)
                return (None,) * len(ctx.needs_input_grad)
            gw = R.conv1_wgrad(ctx.frames, idx if ctx.has_idx else None, gy, relu_y=y)
            return (gw,) + (None,) * (len(ctx.needs_input_grad) - 1)
        if y is not None:
            gy = gy * (y > 0)
        if not ctx.has_idx:
            x = ctx.frames
        elif ctx.store is not None:    # TMA bulk gather straight from the replay payload
            x = ctx.store.gather(idx, ctx.store.alloc_batch(idx.numel(), ("state",)))["state"]
        else:
            x = ctx.frames.index_select(0, idx)
        xf = (x.to(torch.float32) / 255.0).contiguous(memory_format=ctx.mem_format)
        gw = torch.nn.grad.conv2d_weight(xf, ctx.wshape, gy.contiguous(memory_format=ctx.mem_format), stride=4)
        return (gw,) + (None,) * (len(ctx.needs_input_grad) - 1)


class Learner:
    """APE_X/Learner.py Learner (:20-272): train / step / run / state_dict."""

    def __init__(self, cfg: ApexConfig | None = None, connect=None, start_replay: bool = True,
                 writer=None):
        self.cfg = cfg or ApexConfig.from_configuration()
        self.device = torch.device(self.cfg.LEARNER_DEVICE)
        if self.cfg.CUDNN_BENCHMARK and self.device.type == "cuda":
            torch.backends.cudnn.benchmark = True
        self.build_model()
        self.build_optim()
        self.connect = connect
        self.memory = Replay(self.cfg, connect)
        if start_replay and connect is not None:
            self.memory.start()
        self.writer = writer
        if connect is not None:                  # :41-43 — whatever a previous run left behind is dropped
            from .wire import wipe_stale_keys
            wipe_stale_keys(connect)
        self.gamma_n = float(np.float32(0.99 ** self.cfg.UNROLL_STEP))  # hard-coded 0.99, :103
        self._graph = None
        self._world = 1
        self.launches_per_step = None   # libb2rl kernels per fused step (bench.py's gpu_launches)

    def enable_data_parallel(self):
        """Replay-sharded data parallelism: every rank owns a replay shard and
        samples locally; per step one NCCL all-reduce (AVG) of the gradients, kept in ONE flat
        bucket so it is a single collective, and one MAX all-reduce of the max IS weight."""
        from . import dist as D
        self._D = D
        self._world = D.world()
        self._bucket = D.FlatGradBucket(self.model.getParameters(), self.device,
                                        symmetric=D.PeerAllReduce.available(self.device))
        # the dense heads hold 97 % of the parameters and their gradients are complete first in the
        # backward pass: their all-reduce overlaps the backward of the convolution stack
        import os
        heads = [p for p in self.model.getParameters() if p.dim() == 2]
        if os.environ.get("B2RL_NO_OVERLAP"):
            heads = []
        self._bucket.enable_overlap(heads)
        # What is left after backward (the convolution stack's 0.3 MB) is the collective on the critical path: one
        # libb2rl kernel over NVLink peer memory instead of an NCCL launch (csrc/peer.cu; NCCL stays where peer
        # mapping is not available).  (A second overlapped NCCL group for conv_2 / conv_3 was tried: its kernel
        # cannot get SMs while the SM-filling conv_1 weight-gradient kernel runs, so it only added a launch.)
        self.peer_allreduce = self._bucket.enable_peer_allreduce()
        self.peer_allreduce_heads = getattr(self._bucket, "_peer_big", None) is not None
        self._max_w = torch.empty(1, dtype=torch.float32, device=self.device)       # being reduced this step
        self._max_w_use = None                                                      # reduced last step, used now

    def build_model(self):
        self.model = GraphAgent(self.cfg.MODEL).to(self.device)
        self.target_model = GraphAgent(self.cfg.MODEL).to(self.device)
        self._mf = torch.channels_last if self.cfg.CHANNELS_LAST else torch.contiguous_format
        if self.cfg.CHANNELS_LAST:
            self.model.to(memory_format=torch.channels_last)
            self.target_model.to(memory_format=torch.channels_last)
            if self.cfg.FUSED_CONV1:
                # conv_1 runs in libb2rl's kernels, which take the plain (c_out, 4, 8, 8) layout: keeping that weight
                # channels_last cost one re-layout copy per pack (three launches per step)
                for m in (self.model, self.target_model):
                    name = m.first_conv_node()
                    if name is not None and getattr(m, name).is_atari_conv1():
                        w = getattr(m, name).conv_1.weight
                        w.data = w.data.contiguous(memory_format=torch.contiguous_format)
        self.model.dense_3xtf32 = self.target_model.dense_3xtf32 = bool(self.cfg.DENSE_3XTF32)
        self.model.fused_dueling_tail = self.target_model.fused_dueling_tail = bool(self.cfg.FUSED_DUELING_TAIL)

    def build_optim(self):
        info = self.cfg.OPTIM_INFO
        fusable = (self.cfg.FUSED_OPTIM and info["name"] == "rmsprop" and not info.get("momentum", 0)
                   and not info.get("decay", 0) and self.device.type == "cuda")
        if fusable:
            from .optim import FusedRMSprop, flat_grads
            self._grad_flat = flat_grads(self.model.getParameters())     # adjacent head gradients: see linear._stacked_rows
            self.optim = FusedRMSprop(self.model.getParameters(), lr=info["lr"], alpha=info.get("alpha", 0.99),
                                      eps=info.get("eps", 1e-5), centered=info.get("centered", False))
        else:
            self.optim = make_optimizer(info, self.model.getParameters())
        self._fused_optim = fusable

    # -- one training step on an explicit minibatch (reference signature) -----------
    def _to_dev(self, x, dtype):
        if torch.is_tensor(x):
            return x.to(device=self.device, dtype=dtype, non_blocking=True)
        if isinstance(x, np.ndarray) and x.dtype == object:
            x = x.astype(np.float64 if dtype.is_floating_point else np.int64)
        return torch.as_tensor(x).to(device=self.device, dtype=dtype, non_blocking=True)

    def _forward_backward(self, state, action, reward, next_state, done, weight):
        s = (state.to(torch.float32) / 255.0).contiguous(memory_format=self._mf)        # :61-63, on the device
        ns = (next_state.to(torch.float32) / 255.0).contiguous(memory_format=self._mf)  # :65-67
        q = self.model.forward([s])[0]                 # :78
        with torch.no_grad():
            qn_target = self.target_model.forward([ns])[0]   # :85
            qn_online = self.model.forward([ns])[0]          # :87
        notdone = 1.0 - done.to(torch.float32)         # :76
        out = R.apex_target(q.detach(), qn_online, qn_target, action, reward, notdone, weight,
                            self.gamma_n, self.cfg.ALPHA)
        q.backward(out["grad_q"])                      # == loss.backward(), :112-115
        if self._world > 1:
            self._bucket.finish()
        return out

    def train(self, transition, t=0):
        state, action, reward, next_state, done, weight, idx = transition
        state = self._to_dev(state, torch.uint8)
        next_state = self._to_dev(next_state, torch.uint8)
        action = self._to_dev(action, torch.int64)
        reward = self._to_dev(reward, torch.float32)
        done = self._to_dev(done, torch.uint8) if not (torch.is_tensor(done) and done.dtype == torch.bool) \
            else done.to(self.device, torch.uint8)
        weight = self._to_dev(weight, torch.float32)
        out = self._forward_backward(state, action, reward, next_state, done, weight)
        info = self.step()
        info["mean_value"] = out["scalars"][1]
        info["loss"] = out["scalars"][0]
        return info, out["prio"], idx, out["scalars"][2]

    def step(self):
        """Learner.step (:123-138): 'norm' = sqrt(sum_i ||g_i||_2) (sic), RMSprop, zero_grad."""
        if self._fused_optim:
            return {"p_norm": self.optim.step(want_norm=True)[0]}    # update + zero_grad + norm: one launch
        grads = [p.grad for p in self.model.parameters() if p.grad is not None]
        p_norm = torch.stack(torch._foreach_norm(grads, 2)).sum().sqrt()
        self.optim.step()
        self.optim.zero_grad(set_to_none=False)
        return {"p_norm": p_norm}

    # -- fused gather + conv_1 path ------------------------------------------------------------
    def _conv1_ready(self) -> bool:
        if not self.cfg.FUSED_CONV1 or self.model.first_conv_node() is None:
            return False
        if not hasattr(self, "_pack2"):
            self._conv_name = self.model.first_conv_node()
            c_out = getattr(self.model, self._conv_name).conv_1.out_channels
            self._pack1 = R.Conv1Pack(1, self.device, c_out)     # online net (grad pass on s)
            self._pack2 = R.Conv1Pack(2, self.device, c_out)     # online + target in one pass over s'
        return True

    def _streams(self):
        if not hasattr(self, "_fork"):
            # s1: operand packs (needed much later: default priority); s2: the target network's pass — as critical
            # as the main branch (both feed the target kernel), so it gets the main branch's priority in the graph
            self._fork = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device, priority=-2),
                          torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event())
            self._ev_pk, self._ev_tg, self._ev_upd = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
            self._ev_hp, self._ev_mw = torch.cuda.Event(), torch.cuda.Event()
        return self._fork

    def _pack_conv1(self):
        w_on = getattr(self.model, self._conv_name).conv_1.weight
        w_tg = getattr(self.target_model, self._conv_name).conv_1.weight
        R.conv1_pack_jobs([(self._pack1, 0, w_on), (self._pack2, 0, w_on), (self._pack2, 1, w_tg)])   # one launch

    def _pack_conv1_async(self):
        """The conv_1 weight packs of this step on a side stream (they only depend on the weights): they overlap
        the tree sample + scalar gather at the start of the step.  Returns the event the conv_1 launch waits for."""
        s1, s2 = self._streams()[:2]
        cur = torch.cuda.current_stream(self.device)
        self._ev_pk.record(cur)
        s1.wait_event(self._ev_pk)
        s2.wait_event(self._ev_pk)
        with torch.cuda.stream(s2):          # the high-priority side stream: conv_1 waits for these three launches
            self._pack_conv1()
            self._ev_pk.record(s2)
        # The heads' forward operands (online and target: four 3136 x 512 matrices -> two packed images) are needed
        # ~120 us into the step: built on the low-priority side stream now instead of on the main / target branches
        # in front of their GEMMs.
        self._head_packs = None
        if self.cfg.PARALLEL_FORWARDS and self.cfg.BATCHED_ONLINE:
            with torch.cuda.stream(s1):
                with self.model.packed_heads_cache():
                    self.model.prepack_heads()
                    self._head_packs = dict(self.model._pack_cache)
                with self.target_model.packed_heads_cache():      # was ~11 us in front of the target pass's GEMM
                    self.target_model.prepack_heads()
                    self._head_packs_tg = dict(self.target_model._pack_cache)
                self._ev_hp.record(s1)
        return self._ev_pk

    def _forward_backward_fused(self, idx, action, reward, done, weight, packs_done=None, update_tree=False,
                                early_update=False):
        """Same maths as _forward_backward, but s and s' are never staged as uint8/fp32 batches:
        conv_1 reads the sampled rows straight from the replay payload (b2rl_conv1_fused).
        `packs_done`: event after which the conv_1 weight packs are valid (None: pack here).
        `update_tree`: write the new priorities back on a side stream as soon as the target kernel has produced
        them (overlapping backward); the caller then waits for `self._ev_upd` instead of calling store.update.
        `early_update`: the caller WILL call self.step() next; the heads' part of that optimizer step may then be
        issued here, behind their weight gradients on the sink's lane, while the conv stack's backward still runs."""
        st = self.memory.store
        w_on = getattr(self.model, self._conv_name).conv_1.weight
        notdone = None
        if packs_done is None:
            self._pack_conv1()
        else:
            torch.cuda.current_stream(self.device).wait_event(packs_done)
        with self.model.packed_heads_cache():     # the online weights are packed once for both passes
            if self.cfg.PARALLEL_FORWARDS and self.cfg.BATCHED_ONLINE:
                # Q(s) (with grad) and Q_online(s') (without) share the online weights: every kernel after conv_1 is
                # launch / set-up bound at B = 512, so both run as ONE B = 2*BATCHSIZE pass
                # (conv_2, conv_3, heads' GEMM, dueling tail: one launch each instead of two).  The pass is recorded
                # on an OutputTape; the autograd graph of the s half is then built by replaying the recorded per-op
                # outputs (first B rows: views, no kernels) through the same forward code.  Q_target(s') runs
                # beside it on a second stream.
                from .linear import OutputTape
                s1, s2, e0, e1, e2 = self._streams()
                cur = torch.cuda.current_stream(self.device)
                early_packs = packs_done is not None and bool(getattr(self, "_head_packs", None))
                if early_packs:
                    self.model._pack_cache.update(self._head_packs)       # built on s1 at the start of the step
                else:
                    self.model.prepack_heads()
                B = idx.numel()
                c_out = self._pack1.c_out
                if getattr(self, "_y_big", None) is None or self._y_big.shape[1] != B:
                    self._y_big = torch.empty((3, B, 20, 20, c_out), dtype=torch.float32, device=self.device)
                big = self._y_big          # [0] conv_1(s) online, [1] conv_1(s') online, [2] conv_1(s') target
                with torch.no_grad():
                    R.conv1_fused(st.field_view("state"), idx, self._pack1, relu=True, out=big[0:1])
                    y_tg = R.conv1_fused(st.field_view("next_state"), idx, self._pack2, relu=True, out=big[1:3])[1]
                if early_packs:
                    cur.wait_event(self._ev_hp)      # long done; the heads' GEMM is ~100 us away
                e0.record(cur)
                s1.wait_event(e0)
                s2.wait_event(e0)
                with torch.no_grad():
                    with torch.cuda.stream(s2):
                        with self.target_model.packed_heads_cache():
                            if early_packs:
                                self.target_model._pack_cache.update(self._head_packs_tg)
                            qn_target = self.target_model.forward_from_conv1(y_tg, True)[0]  # :85
                        e2.record(s2)
                    with torch.cuda.stream(s1):
                        if action.dtype != torch.int64:            # int32 replay field -> the target kernel's int64:
                            action = action.to(torch.int64)        # off the main branch (it sat in front of conv_1)
                        notdone = 1.0 - done.to(torch.float32)     # likewise (two launches in front of the target kernel)
                        self.model.prepack_heads(transposed=True)   # W^T operand of the heads' dgrad, off the main branch
                        e1.record(s1)
                    y_both = big[0:2].view(2 * B, 20, 20, c_out).permute(0, 3, 1, 2)     # logical NCHW, physical NHWC
                    with OutputTape.record() as tape:
                        q_all = self.model.forward_from_conv1(y_both, True)[0]          # :78 and :87 in one pass
                qn_online = q_all[B:]
                y = _Conv1Gathered.apply(w_on, st.field_view("state"), idx, self._pack1, self._mf, st,
                                         big[0].permute(0, 3, 1, 2), True)
                with OutputTape.replay(tape.half(B)):
                    q = self.model.forward_from_conv1(y, True)[0]                       # graph only: outputs replayed
                cur.wait_event(e1)
                cur.wait_event(e2)
            elif self.cfg.PARALLEL_FORWARDS:
                # The three passes are independent until the target kernel: fork them onto three streams
                # (captured as parallel branches of the step's CUDA graph) so their small kernels overlap.
                s1, s2, e0, e1, e2 = self._streams()
                cur = torch.cuda.current_stream(self.device)
                self.model.prepack_heads()
                with torch.no_grad():
                    y_on, y_tg = R.conv1_fused(st.field_view("next_state"), idx, self._pack2, relu=True)
                e0.record(cur)
                s1.wait_event(e0)
                s2.wait_event(e0)
                with torch.no_grad():
                    with torch.cuda.stream(s1):
                        qn_online = self.model.forward_from_conv1(y_on, True)[0]        # :87
                        self.model.prepack_heads(transposed=True)   # W^T operand of the heads' dgrad, off the main branch
                        e1.record(s1)
                    with torch.cuda.stream(s2):
                        qn_target = self.target_model.forward_from_conv1(y_tg, True)[0]  # :85
                        e2.record(s2)
                y = _Conv1Gathered.apply(w_on, st.field_view("state"), idx, self._pack1, self._mf, st, None, True)
                q = self.model.forward_from_conv1(y, True)[0]                        # :78 (ReLU in the conv_1 epilogue)
                cur.wait_event(e1)
                cur.wait_event(e2)
            else:
                with torch.no_grad():
                    y_on, y_tg = R.conv1_fused(st.field_view("next_state"), idx, self._pack2, relu=True)
                    qn_online = self.model.forward_from_conv1(y_on, True)[0]        # :87
                    qn_target = self.target_model.forward_from_conv1(y_tg, True)[0]  # :85
                y = _Conv1Gathered.apply(w_on, st.field_view("state"), idx, self._pack1, self._mf, st, None, True)
                q = self.model.forward_from_conv1(y, True)[0]                        # :78 (ReLU in the conv_1 epilogue)
        if notdone is None:
            notdone = 1.0 - done.to(torch.float32)
        out = R.apex_target(q.detach(), qn_online, qn_target, action, reward, notdone, weight,
                            self.gamma_n, self.cfg.ALPHA)
        if update_tree:      # priority write-back (one CTA) next to backward instead of after the optimizer
            s2 = self._streams()[1]
            cur = torch.cuda.current_stream(self.device)
            self._ev_tg.record(cur)
            s2.wait_event(self._ev_tg)
            with torch.cuda.stream(s2):
                st.update(idx, out["prio"])
                self._ev_upd.record(s2)
        if self.cfg.DEFERRED_WGRAD and self._fused_optim:      # grads are pre-allocated and zeroed by the optimizer
            if not hasattr(self, "_sink"):
                from .linear import WeightGradSink
                self._sink = WeightGradSink(self.device)
                getattr(self.model, self._conv_name).split_backward = True
                if self._world > 1:
                    self._bucket.attach_sink(self._sink)
                # Learner.step (:123-138) has no clipping: a parameter can be stepped once its own gradient is final
                # (after its all-reduce when data parallel).  The heads (97 % of the elements) are final ~150 us
                # before the conv stack's.
                self._early_params = [p for n, p in self.model.named_parameters()
                                      if not n.startswith(self._conv_name + ".")]
                self._early_ok = (self.cfg.EARLY_HEAD_UPDATE and bool(self._early_params)
                                  and self.optim.set_early(self._early_params))
            self._sink.grads_are_zero = True      # this branch: grads pre-allocated and zeroed by the fused optimizer
            with self._sink.active():
                q.backward(out["grad_q"])
            if early_update and self._early_ok and \
                    all(id(p) in self._sink.accumulated[0] for p in self._early_params):
                if self._world > 1:
                    # data parallel: the heads' all-reduce was launched from this lane when their last gradient
                    # landed; the lane waits for it, then steps them — still beside the conv stack's backward
                    self._sink.run_on_lane(lambda: self._bucket.wait_group(0) and self.optim.step_early(), 0)
                else:
                    self._sink.run_on_lane(self.optim.step_early, 0)
            self._sink.join()
        else:
            q.backward(out["grad_q"])
        if self._world > 1:
            self._bucket.finish()
        return out

    # -- the whole hot loop iteration as one CUDA graph -----------------------------------
    def fused_step(self, use_graph: bool = True):
        """sample -> gather -> forwards -> target -> backward -> RMSprop -> priority
        write-back (APE_X/Learner.py:165-197) with no host round trip."""
        if self._graph is not None:
            self._graph.replay()
            return self._static
        B = self.cfg.BATCHSIZE
        st = self.memory.store
        fused_conv1 = self._conv1_ready()

        def body():
            max_w, mw_work = None, None
            if self._world > 1:
                # priority-max reduction: IS weights are normalised by the GLOBAL max weight.  The MAX
                # all-reduce of this step's local value runs behind the step and is used by the next one
                # (the reference's own max_weight is up to 16 minibatches stale, APE_X/ReplayMemory.py:61-67).
                mw_work = self._D.all_reduce_max_(st.max_weight(self.cfg.BETA, out=self._max_w), async_op=True)
                max_w = self._max_w_use
            side = fused_conv1 and self.cfg.PARALLEL_FORWARDS
            packs_done = self._pack_conv1_async() if side else None
            if fused_conv1:
                # ONE launch draws the minibatch: indices + IS weights from the sum-tree and the sampled slots'
                # scalar fields (a, r, done); the frames are read in place by the conv_1 kernels.
                # (Drawing the NEXT minibatch at the end of the step would hide these ~5 us too, but a ring slot
                # overwritten by the ingest between the draw and its use would pair new frames with the old
                # record's a / r / done — DESIGN.md §4.2.)
                if not hasattr(self, "_cur"):
                    self._cur = dict(st.alloc_batch(B, ("action", "reward", "done")),
                                     idx=torch.empty(B, dtype=torch.int64, device=self.device),
                                     w=torch.empty(B, dtype=torch.float32, device=self.device))
                c = self._cur
                st.sample_fetch(B, self.cfg.BETA, c["idx"], c["w"], {k: c[k] for k in ("action", "reward", "done")},
                                max_w=max_w)
                idx = c["idx"]
                batched = self.cfg.PARALLEL_FORWARDS and self.cfg.BATCHED_ONLINE      # converts the action itself
                out = self._forward_backward_fused(idx, c["action"] if batched else c["action"].to(torch.int64),
                                                   c["reward"], c["done"], c["w"],
                                                   packs_done=packs_done, update_tree=side, early_update=True)
            else:
                idx, _, w = st.sample(B, beta=self.cfg.BETA, want_prob=False, max_w=max_w)
                b = st.gather(idx)
                out = self._forward_backward(b["state"], b["action"].to(torch.int64), b["reward"],
                                             b["next_state"], b["done"], w)
            info = self.step()
            if side:
                torch.cuda.current_stream(self.device).wait_event(self._ev_upd)
            else:
                st.update(idx, out["prio"])
            if mw_work is not None:
                # the reduced maximum becomes the NEXT step's normaliser: wait + copy on a side stream (this step's
                # draw has long read the old value), joined here at no cost instead of 5 us at the end of the step
                s1 = self._streams()[0]
                with torch.cuda.stream(s1):
                    mw_work.wait()
                    self._max_w_use.copy_(self._max_w)
                    self._ev_mw.record(s1)
                torch.cuda.current_stream(self.device).wait_event(self._ev_mw)
            return {"scalars": out["scalars"], "p_norm": info["p_norm"], "prio": out["prio"], "idx": idx}

        lib = st.lib
        if self._world > 1 and self._max_w_use is None:      # first step: reduce synchronously once
            self._max_w_use = self._D.all_reduce_max_(st.max_weight(self.cfg.BETA)).clone()
        if not use_graph:
            c0 = lib.b2rl_launch_count()
            r = body()
            self.launches_per_step = lib.b2rl_launch_count() - c0
            return r
        self.optim.zero_grad(set_to_none=False)
        # The step's main branch is captured on a HIGH-priority stream (kernel nodes inherit it): the side branches
        # (weight gradients, early optimizer step, operand packs) only fill SMs the critical chain leaves idle.
        side = torch.cuda.Stream(self.device, priority=-2)
        # The ingest thread keeps pushing on the same replay handle: hold its lock so that no cudaMalloc /
        # cudaHostAlloc / copy of that thread lands inside the warm-up or the (global-mode) capture.
        with self.memory._lock:
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(3):   # warm-up: lazy inits (cuDNN plans, optimizer state) happen outside capture
                    body()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            c0 = lib.b2rl_launch_count()
            with torch.cuda.graph(g, stream=side):
                self._static = body()
            self.launches_per_step = lib.b2rl_launch_count() - c0   # recorded into the graph, replayed each step
        self._graph = g
        g.replay()
        return self._static

    # -- parameter publication / main loop -------------------------------------------------
    @property
    def state_dict(self):
        return {k: v.cpu() for k, v in self.model.state_dict().items()}

    @property
    def target_state_dict(self):
        return {k: v.cpu() for k, v in self.target_model.state_dict().items()}

    def run(self, max_steps: int | None = None, log_every: int = 500):
        """Learner.run (:140-262) with the reference's cadence: wait for BUFFER_SIZE records, announce `Start`,
        hard target sync every TARGET_FREQUENCY steps, parameter publication every 50, and every `log_every`
        (500) steps the eviction request (`memory.lock`, :189-191), the `reward` drain + log line (:219-253) and a
        checkpoint of the online weights (:256-262).  Publication and checkpoints go through ParamPublisher
        (async D2H into pinned memory), so none of them stalls the learner stream."""
        from .publish import ParamPublisher
        from . import wire
        while len(self.memory.memory) <= self.cfg.BUFFER_SIZE:
            time.sleep(0.05)
        if self.connect is not None:
            self.connect.set("state_dict", pickle.dumps(self.state_dict))
            self.connect.set("count", pickle.dumps(1))
            self.connect.set("target_state_dict", pickle.dumps(self.target_state_dict))
            self.connect.set("Start", pickle.dumps(True))
        pub = ParamPublisher(self.model, self.connect, "state_dict", "count")
        pub_t = ParamPublisher(self.target_model, self.connect, "target_state_dict", None)
        ckpt_path = wire.checkpoint_path(self.cfg.LOG_W)
        ckpt = ParamPublisher(self.model, None, None, None,
                              on_ready=lambda sd, step: torch.save(sd, ckpt_path)) if ckpt_path else None
        self._publishers = (pub, pub_t) + ((ckpt,) if ckpt else ())
        step = 0
        t0 = time.time()
        acc = None
        self.last_log = None
        while max_steps is None or step < max_steps:
            out = self.fused_step()
            step += 1
            tot = torch.cat([out["scalars"], out["p_norm"].reshape(1)])
            acc = tot.clone() if acc is None else acc + tot
            if step % self.cfg.TARGET_FREQUENCY == 0:
                self.target_model.updateParameter(self.model, 1)
                pub_t.snapshot(step)                 # async D2H; published by a later poll()
            if step % 50 == 0:
                pub.snapshot(step - 50)              # :212-216, without stalling the learner stream
            for p in self._publishers:
                p.poll()
            if step % log_every == 0:
                self.memory.lock = True              # :189-191 eviction request, served by the ingest thread
                if self.connect is None or not self.memory.is_alive():
                    self.memory._evict_on_request()
                reward, n_rew = wire.drain_rewards(self.connect) if self.connect is not None else (-21.0, 0)
                loss, mean_value, mean_w, norm = (acc / log_every).tolist()
                dt = (time.time() - t0) / log_every
                self.last_log = {"step": step, "mean_value": mean_value, "norm": norm, "reward": reward,
                                 "loss": loss, "mean_weight": mean_w, "time_per_step": dt}
                print(f"step:{step} // mean_value:{mean_value:.3f} // norm: {norm:.3f} // REWARD:{reward:.3f} // "
                      f"NUM_MEMORY:{len(self.memory.memory)} // Mean_Weight:{mean_w:.3f} // "
                      f"MAX_WEIGHT:{self.memory.memory.max_weight:.3f} // TIME:{dt:.5f} // loss:{loss:.5f}")
                if self.writer is not None:
                    if n_rew:
                        self.writer.add_scalar("Reward", reward, step)
                    self.writer.add_scalar("value", mean_value, step)
                    self.writer.add_scalar("norm", norm, step)
                if ckpt is not None:
                    ckpt.snapshot(step)
                acc, t0 = None, time.time()
        return step
