"""DeviceReplay — one HBM-resident replay shard (payload SoA + fp64 sum-tree).

Thin, typed wrapper over the C ABI (include/b2rl.h).  All heavy lifting happens
in the CUDA kernels of libb2rl.so; this file only marshals torch tensors'
device pointers and the current CUDA stream.  It is the object the reference-
facing mirrors (per.py: PER / PrioritizedMemory, apex.py: Replay) are built on.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Sequence

import numpy as np
import torch

from . import _lib
from ._lib import ReplayDesc, check

FRAME_STACK_BYTES = 4 * 84 * 84  # one (4,84,84) uint8 observation, 28 224 B


@dataclass(frozen=True)
class Field:
    name: str
    dtype: torch.dtype
    shape: tuple  # per-slot shape

    @property
    def nbytes(self) -> int:
        n = 1
        for s in self.shape:
            n *= int(s)
        return n * torch.empty((), dtype=self.dtype).element_size()


# Record layouts of the three learners.
APEX_FIELDS = (  # [s, a, R_n, s', done, prio]  APE_X/Player.py:252-261
    Field("state", torch.uint8, (4, 84, 84)),
    Field("next_state", torch.uint8, (4, 84, 84)),
    Field("action", torch.int32, ()),
    Field("reward", torch.float32, ()),
    Field("done", torch.uint8, ()),
)


def r2d2_fields(T: int = 80, hidden: int = 512):
    """[(h0,h1), (s,a,r) x T, done, prio]  R2D2/ReplayMemory.py:70-88."""
    return (
        Field("state", torch.uint8, (T, 4, 84, 84)),
        Field("action", torch.int32, (T,)),
        Field("reward", torch.float32, (T,)),
        Field("h0", torch.float32, (hidden,)),
        Field("h1", torch.float32, (hidden,)),
        Field("notdone", torch.float32, ()),
    )


def impala_fields(T: int = 20):
    """(s[T+1], a[T], mu[T], r[T], done)  IMPALA/ReplayMemory.py:34-43."""
    return (
        Field("state", torch.uint8, (T + 1, 4 * 84 * 84)),
        Field("action", torch.int32, (T,)),
        Field("mu", torch.float32, (T,)),
        Field("reward", torch.float32, (T,)),
        Field("done", torch.float32, ()),
    )


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _CudaView:
    """Expose library-owned device memory to torch via __cuda_array_interface__."""

    def __init__(self, ptr: int, shape, typestr: str, owner):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}
        self._owner = owner


_TYPESTR = {torch.uint8: "|u1", torch.int32: "<i4", torch.int64: "<i8", torch.float32: "<f4",
            torch.float64: "<f8", torch.int8: "|i1", torch.int16: "<i2", torch.float16: "<f2"}


class DeviceReplay:
    def __init__(self, capacity: int, fields: Sequence[Field] = APEX_FIELDS, device="cuda:0"):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.B2RLError("DeviceReplay lives in GPU HBM; there is no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.fields = tuple(fields)
        if len(self.fields) > _lib.MAX_FIELDS:
            raise ValueError("too many payload fields")
        self.capacity = int(capacity)
        d = ReplayDesc()
        d.capacity = self.capacity
        d.n_fields = len(self.fields)
        d.device = self.device.index
        for i, f in enumerate(self.fields):
            d.field_bytes[i] = f.nbytes
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)  # make sure the primary context exists
        h = C.c_void_p()
        check(self.lib.b2rl_replay_create(C.byref(d), C.byref(h)))
        self._h = h

    # -- bookkeeping -----------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.b2rl_replay_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _sizes(self):
        s, c, hd = C.c_int64(), C.c_int64(), C.c_int64()
        check(self.lib.b2rl_replay_size(self._h, C.byref(s), C.byref(c), C.byref(hd)))
        return s.value, c.value, hd.value

    def __len__(self) -> int:
        return self._sizes()[0]

    @property
    def head(self) -> int:
        return self._sizes()[2]

    def _st(self) -> int:
        return _stream_ptr(self.device)

    def field_view(self, name_or_idx) -> torch.Tensor:
        """Zero-copy torch view (capacity, *shape) of a library-owned payload field."""
        i = name_or_idx if isinstance(name_or_idx, int) else [f.name for f in self.fields].index(name_or_idx)
        f = self.fields[i]
        p = C.c_void_p()
        check(self.lib.b2rl_replay_field_ptr(self._h, i, C.byref(p)))
        view = _CudaView(p.value, (self.capacity,) + tuple(f.shape), _TYPESTR[f.dtype], self)
        with torch.cuda.device(self.device):
            return torch.as_tensor(view, device=self.device)

    # -- ingest ----------------------------------------------------------------
    def push(self, fields: Sequence, priorities) -> None:
        """fields[i]: tensor / ndarray (n, *shape_i), host (pinned preferred) or device."""
        keep = []
        ptrs = (C.c_void_p * _lib.MAX_FIELDS)()
        n = None
        for i, (f, x) in enumerate(zip(self.fields, fields)):
            if x is None:
                ptrs[i] = None
                continue
            t = torch.as_tensor(x)
            if t.dtype != f.dtype:
                t = t.to(f.dtype)
            t = t.contiguous()
            if n is None:
                n = t.shape[0]
            assert t.shape[0] == n and t.numel() * t.element_size() == n * f.nbytes, f"bad shape for {f.name}"
            keep.append(t)
            ptrs[i] = t.data_ptr()
        pr = torch.as_tensor(priorities).to(torch.float32).contiguous()
        if n is None:
            n = pr.shape[0]
        assert pr.numel() == n
        keep.append(pr)
        check(self.lib.b2rl_replay_push(self._h, ptrs, pr.data_ptr(), n, self._st()))
        # host buffers must outlive the async copies on this stream
        if any(not t.is_cuda and not t.is_pinned() for t in keep):
            pass  # pageable memory: cudaMemcpyAsync already staged it synchronously
        else:
            self._inflight = keep

    # -- pipelined ingest: H2D of the next batch overlaps the current learner step ------------
    def push_begin(self, fields: Sequence, n: int) -> None:
        """reserve (current stream) + payload copy on a private ingest stream.  Pair with push_commit."""
        if not hasattr(self, "_ingest_stream"):
            self._ingest_stream = torch.cuda.Stream(self.device)
            self._ev_reserved = torch.cuda.Event()
            self._ev_copied = torch.cuda.Event()
        start = C.c_int64()
        check(self.lib.b2rl_replay_reserve(self._h, int(n), C.byref(start), self._st()))
        self._ev_reserved.record(torch.cuda.current_stream(self.device))
        keep = []
        ptrs = (C.c_void_p * _lib.MAX_FIELDS)()
        for i, (f, x) in enumerate(zip(self.fields, fields)):
            if x is None:
                ptrs[i] = None
                continue
            t = torch.as_tensor(x)
            assert t.dtype == f.dtype and t.is_contiguous() and t.numel() * t.element_size() == n * f.nbytes
            keep.append(t)
            ptrs[i] = t.data_ptr()
        with torch.cuda.stream(self._ingest_stream):
            self._ingest_stream.wait_event(self._ev_reserved)
            check(self.lib.b2rl_replay_copy_payload(self._h, ptrs, start.value, int(n),
                                                    self._ingest_stream.cuda_stream))
            self._ev_copied.record(self._ingest_stream)
        self._pending = (keep, int(n))

    def push_commit(self, priorities) -> None:
        """Make the records copied by push_begin sampleable (current stream waits for the copy)."""
        keep, n = self._pending
        pr = torch.as_tensor(priorities).to(torch.float32).contiguous()
        assert pr.numel() == n
        torch.cuda.current_stream(self.device).wait_event(self._ev_copied)
        check(self.lib.b2rl_replay_commit(self._h, pr.data_ptr(), n, self._st()))
        self._inflight = keep + [pr]
        self._pending = None

    def ingest_pipelined(self, fields: Sequence | None, priorities=None) -> None:
        """One C call per iteration (b2rl_replay_ingest_pipelined): publish the batch copied during the previous
        step, retire the next batch's slots, start its host->device copy on the library's copy stream.
        `fields`: pinned host (or device) tensors in field order, `priorities`: fp32[n]; fields None = flush."""
        if fields is None:
            check(self.lib.b2rl_replay_ingest_pipelined(self._h, None, None, 0, self._st()))
            self._pipe_keep = None
            return
        key = tuple(t.data_ptr() for t in fields) + (priorities.data_ptr(),)
        cache = getattr(self, "_pipe_cache", None)
        if cache is None or cache[0] != key:          # pointer array built once per set of staging buffers
            ptrs = (C.c_void_p * _lib.MAX_FIELDS)()
            n = None
            for i, (f, t) in enumerate(zip(self.fields, fields)):
                assert t.dtype == f.dtype and t.is_contiguous()
                n = t.shape[0] if n is None else n
                assert t.numel() * t.element_size() == n * f.nbytes, f"bad shape for {f.name}"
                ptrs[i] = t.data_ptr()
            assert priorities.dtype == torch.float32 and priorities.numel() == n and priorities.is_contiguous()
            cache = self._pipe_cache = (key, ptrs, int(n))
        _, ptrs, n = cache
        check(self.lib.b2rl_replay_ingest_pipelined(self._h, ptrs, priorities.data_ptr(), n, self._st()))
        self._pipe_keep = (fields, priorities)        # host buffers stay alive until the next call

    def evict(self, delta: int) -> None:
        check(self.lib.b2rl_replay_evict(self._h, int(delta), self._st()))

    def fill_hash(self, n: int, seed: int = 0xB200) -> None:
        check(self.lib.b2rl_replay_fill_hash(self._h, int(n), int(seed) & 0xFFFFFFFF, self._st()))

    # -- tree ------------------------------------------------------------------
    def build(self, priorities: torch.Tensor) -> None:
        p = priorities.to(device=self.device, dtype=torch.float32).contiguous()
        check(self.lib.b2rl_tree_build(self._h, p.data_ptr(), p.numel(), self._st()))

    def seed(self, seed: int, counter: int = 0) -> None:
        """(Re)seed the device-resident Philox stream used when no uniforms are passed."""
        check(self.lib.b2rl_replay_seed(self._h, int(seed), int(counter), self._st()))

    def sample(self, n: int, beta: float = 0.4, u01: torch.Tensor | None = None,
               want_prob: bool = True, out=None, max_w: torch.Tensor | None = None):
        """-> (idx int64[n], prob fp32[n], weight fp32[n]) on the device.
        With explicit fp64 uniforms `u01` (parity runs) or, if None, from the
        handle's device-resident Philox stream (graph-replayable)."""
        if out is None:
            idx = torch.empty(n, dtype=torch.int64, device=self.device)
            prob = torch.empty(n, dtype=torch.float32, device=self.device) if want_prob else None
            w = torch.empty(n, dtype=torch.float32, device=self.device)
        else:
            idx, prob, w = out
        pp = prob.data_ptr() if prob is not None else None
        mw = max_w.data_ptr() if max_w is not None else None   # all-reduced max IS weight (multi-GPU)
        if u01 is not None:
            u01 = u01.to(device=self.device, dtype=torch.float64).contiguous()
            assert u01.numel() == n
            check(self.lib.b2rl_tree_sample(self._h, u01.data_ptr(), 0, 0, n, float(beta), mw,
                                            idx.data_ptr(), pp, w.data_ptr(), self._st()))
        else:
            check(self.lib.b2rl_tree_sample_stream(self._h, n, float(beta), mw, idx.data_ptr(), pp,
                                                   w.data_ptr(), self._st()))
        return idx, prob, w

    def sample_fetch(self, n: int, beta: float, idx: torch.Tensor, w: torch.Tensor, small: dict,
                     max_w: torch.Tensor | None = None, prob: torch.Tensor | None = None):
        """sample() from the device-resident Philox stream into the given buffers AND, in the same launch, copy
        the sampled slots' scalar fields (1/2/4/8-byte rows, e.g. action / reward / done) into `small[name]`.
        Everything is written in place, so a captured graph can prefetch the next minibatch's indices."""
        ptrs = (C.c_void_p * _lib.MAX_FIELDS)()
        for i, f in enumerate(self.fields):
            t = small.get(f.name)
            ptrs[i] = t.data_ptr() if t is not None else None
        check(self.lib.b2rl_tree_sample_fetch(self._h, int(n), float(beta), max_w.data_ptr() if max_w is not None else None,
                                              idx.data_ptr(), prob.data_ptr() if prob is not None else None,
                                              w.data_ptr(), ptrs, self._st()))
        return idx, prob, w

    def sample_counter(self, seed: int, counter: int, n: int, beta: float = 0.4):
        """Stateless Philox draw: uniform k = philox(seed, counter + k)."""
        idx = torch.empty(n, dtype=torch.int64, device=self.device)
        prob = torch.empty(n, dtype=torch.float32, device=self.device)
        w = torch.empty(n, dtype=torch.float32, device=self.device)
        check(self.lib.b2rl_tree_sample(self._h, None, int(seed), int(counter), n, float(beta), None,
                                        idx.data_ptr(), prob.data_ptr(), w.data_ptr(), self._st()))
        return idx, prob, w

    def philox_uniforms(self, seed: int, offset: int, n: int) -> torch.Tensor:
        out = torch.empty(n, dtype=torch.float64, device=self.device)
        check(self.lib.b2rl_philox_uniforms(seed, offset, n, out.data_ptr(), self._st()))
        return out

    def update(self, idx: torch.Tensor, vals: torch.Tensor) -> None:
        idx = idx.to(device=self.device, dtype=torch.int64).contiguous()
        vals = vals.to(device=self.device, dtype=torch.float32).contiguous()
        assert idx.numel() == vals.numel()
        check(self.lib.b2rl_tree_update(self._h, idx.data_ptr(), vals.data_ptr(), idx.numel(), self._st()))

    def stats(self, beta: float = 0.4) -> torch.Tensor:
        """device tensor fp64[3] = {sum(p), min p, max IS weight}."""
        out = torch.empty(3, dtype=torch.float64, device=self.device)
        check(self.lib.b2rl_tree_stats(self._h, float(beta), out.data_ptr(), None, self._st()))
        return out

    def max_weight(self, beta: float = 0.4, out: torch.Tensor | None = None) -> torch.Tensor:
        """device fp32[1]: this shard's max IS weight (operand of the multi-GPU MAX all-reduce)."""
        out = torch.empty(1, dtype=torch.float32, device=self.device) if out is None else out
        check(self.lib.b2rl_tree_stats(self._h, float(beta), None, out.data_ptr(), self._st()))
        return out

    def priorities(self, start: int = 0, n: int | None = None) -> torch.Tensor:
        n = self.capacity - start if n is None else n
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        check(self.lib.b2rl_tree_leaves(self._h, start, n, out.data_ptr(), self._st()))
        return out

    # -- gather ----------------------------------------------------------------
    def alloc_batch(self, n: int, names: Sequence[str] | None = None):
        names = [f.name for f in self.fields] if names is None else list(names)
        return {f.name: torch.empty((n,) + tuple(f.shape), dtype=f.dtype, device=self.device)
                for f in self.fields if f.name in names}

    def gather(self, idx: torch.Tensor, out: dict | None = None) -> dict:
        n = idx.numel()
        if out is None:
            out = self.alloc_batch(n)
        ptrs = (C.c_void_p * _lib.MAX_FIELDS)()
        for i, f in enumerate(self.fields):
            t = out.get(f.name)
            ptrs[i] = t.data_ptr() if t is not None else None
        check(self.lib.b2rl_replay_gather(self._h, idx.data_ptr(), n, ptrs, self._st()))
        return out


# ---- stateless target kernels -------------------------------------------------
def _p(t):
    return None if t is None else t.data_ptr()


def apex_target(q_s, qn_online, qn_target, action, reward, notdone, weight, gamma_n, alpha,
                want_grad=True, out=None):
    """APE_X/Learner.py:85-121 in one launch.  All inputs device tensors."""
    lib = _lib.load()
    B, A = q_s.shape
    dev = q_s.device
    if out is None:
        out = {"target": torch.empty(B, device=dev), "td": torch.empty(B, device=dev),
               "prio": torch.empty(B, device=dev),
               "grad_q": torch.empty(B, A, device=dev) if want_grad else None,
               "scalars": torch.empty(3, device=dev)}
    action = action.to(torch.int64)
    for t in (q_s, qn_online, qn_target, action, reward, notdone, weight):
        assert t.is_cuda and t.is_contiguous()
    check(lib.b2rl_apex_target(_p(q_s), _p(qn_online), _p(qn_target), _p(action), _p(reward), _p(notdone),
                               _p(weight), B, A, float(np.float32(gamma_n)), float(alpha),
                               _p(out["target"]), _p(out["td"]), _p(out["prio"]), _p(out.get("grad_q")),
                               _p(out["scalars"]), _stream_ptr(dev)))
    return out


def r2d2_target(q, q_target, action, reward, notdone, weight, n_step, gamma, alpha, rescale=True,
                want_grad=True):
    """R2D2/Learner.py:110-198 in one launch (+ a 1-thread finisher for the scalars)."""
    lib = _lib.load()
    L, B, A = q.shape
    dev = q.device
    out = {"target": torch.empty(L - 1, B, device=dev), "td": torch.empty(L - 1, B, device=dev),
           "prio": torch.empty(B, device=dev),
           "grad_q": torch.empty(L, B, A, device=dev) if want_grad else None,
           "scalars": torch.empty(2, device=dev)}
    action = action.to(torch.int64).contiguous()
    check(lib.b2rl_r2d2_target(_p(q), _p(q_target), _p(action), _p(reward), _p(notdone), _p(weight),
                               L, B, A, int(n_step), float(gamma), float(alpha), int(bool(rescale)),
                               _p(out["target"]), _p(out["td"]), _p(out["prio"]), _p(out["grad_q"]),
                               _p(out["scalars"]), _stream_ptr(dev)))
    return out


def vtrace(pi_a, mu_a, value, bootstrap, reward, gamma, c_lambda, c_bar, p_bar):
    """IMPALA/Learner.py:141-215 in one launch.  (T, B) time-major."""
    lib = _lib.load()
    T, B = value.shape
    dev = value.device
    vt = torch.empty(T, B, device=dev)
    adv = torch.empty(T, B, device=dev)
    check(lib.b2rl_vtrace(_p(pi_a), _p(mu_a), _p(value), _p(bootstrap), _p(reward), T, B,
                          float(np.float32(gamma)), float(c_lambda), float(c_bar), float(p_bar),
                          _p(vt), _p(adv), _stream_ptr(dev)))
    return vt, adv


# ---- fused gather + first convolution (tcgen05) ---------------------------------
class Conv1Pack:
    """Packed conv_1 weights of 1 or 2 networks for b2rl_conv1_fused (re-pack after every
    optimizer step of the online net / target sync: one tiny launch per network)."""

    def __init__(self, n_nets: int, device, c_out: int = 32):
        assert n_nets in (1, 2) and c_out in (16, 32)
        self.n_nets, self.c_out = n_nets, c_out
        self.device = torch.device(device)
        self.bq = torch.empty(n_nets * 4 * c_out * 256, dtype=torch.int8, device=self.device)
        self.scale = torch.empty(n_nets * c_out, dtype=torch.float32, device=self.device)

    def pack(self, net: int, weight: torch.Tensor) -> None:
        """weight: conv_1.weight fp32 (c_out, 4, 8, 8) (any memory format)."""
        w = weight.detach().to(torch.float32).contiguous(memory_format=torch.contiguous_format)
        assert w.shape == (self.c_out, 4, 8, 8)
        check(_lib.load().b2rl_conv1_pack(w.data_ptr(), net, self.n_nets, self.c_out, self.bq.data_ptr(),
                                          self.scale.data_ptr(), _stream_ptr(self.device)))


def conv1_pack_jobs(jobs) -> None:
    """jobs: up to 4 (pack, net, weight) triples -> ONE launch of all the packs (b2rl_conv1_pack_jobs)."""
    import ctypes as C
    n = len(jobs)
    ws = [w.detach().to(torch.float32).contiguous(memory_format=torch.contiguous_format) for _, _, w in jobs]
    c_out = jobs[0][0].c_out
    assert all(p.c_out == c_out and w.shape == (c_out, 4, 8, 8) for (p, _, _), w in zip(jobs, ws))
    arr = C.c_void_p * n
    check(_lib.load().b2rl_conv1_pack_jobs(
        arr(*[w.data_ptr() for w in ws]), (C.c_int32 * n)(*[net for _, net, _ in jobs]),
        (C.c_int32 * n)(*[p.n_nets for p, _, _ in jobs]), arr(*[p.bq.data_ptr() for p, _, _ in jobs]),
        arr(*[p.scale.data_ptr() for p, _, _ in jobs]), n, c_out, _stream_ptr(jobs[0][0].device)))


def conv1_fused(frames: torch.Tensor, idx, pack: Conv1Pack, relu: bool = False, out=None):
    """frames: uint8 (rows, 4, 84, 84) contiguous (e.g. DeviceReplay.field_view("state"));
    idx: int64[n] rows to take (None: all rows in order).
    -> list of n_nets tensors (n, c_out, 20, 20) fp32 in channels_last memory format."""
    assert frames.dtype == torch.uint8 and frames.is_contiguous() and frames[0].numel() == FRAME_STACK_BYTES
    n = frames.shape[0] if idx is None else idx.numel()
    dev = frames.device
    if out is None:
        out = torch.empty((pack.n_nets, n, 20, 20, pack.c_out), dtype=torch.float32, device=dev)
    check(_lib.load().b2rl_conv1_fused(frames.data_ptr(), frames.shape[0], None if idx is None else idx.data_ptr(),
                                       n, pack.bq.data_ptr(), pack.scale.data_ptr(), pack.n_nets, pack.c_out,
                                       out.data_ptr(),
                                       int(bool(relu)), _stream_ptr(dev)))
    return [out[i].permute(0, 3, 1, 2) for i in range(pack.n_nets)]   # logical NCHW, physical NHWC


_wgrad_ws = {}


def conv1_wgrad(frames: torch.Tensor, idx, gy: torch.Tensor, out: torch.Tensor | None = None,
                accumulate: bool = False, relu_y: torch.Tensor | None = None) -> torch.Tensor:
    """dL/dW of conv_1 from the sampled uint8 rows and dL/dy, without staging the rows (b2rl_conv1_wgrad).
    frames: uint8 (rows, 4, 84, 84) contiguous; idx: int64[n] or None; gy: (n, c_out, 20, 20) fp32
    (made channels_last if it is not) -> (c_out, 4, 8, 8) fp32.  relu_y: the post-ReLU output of
    conv1_fused(relu=True) for the same rows; gy is then dL/d(relu output) and is masked by (y > 0) in the kernel."""
    assert frames.dtype == torch.uint8 and frames.is_contiguous() and frames[0].numel() == FRAME_STACK_BYTES
    n = frames.shape[0] if idx is None else idx.numel()
    c_out = gy.shape[1]
    assert gy.shape == (n, c_out, 20, 20) and gy.dtype == torch.float32
    gy = gy.contiguous(memory_format=torch.channels_last)
    if relu_y is not None:
        assert relu_y.shape == gy.shape and relu_y.dtype == torch.float32
        relu_y = relu_y.contiguous(memory_format=torch.channels_last)
    dev = frames.device
    key = (dev, c_out)
    if key not in _wgrad_ws:
        _wgrad_ws[key] = torch.empty(_lib.load().b2rl_conv1_wgrad_workspace_floats(c_out), dtype=torch.float32,
                                     device=dev)
    if out is None:
        out = torch.empty((c_out, 4, 8, 8), dtype=torch.float32, device=dev)
        accumulate = False
    assert out.is_contiguous() and out.numel() == c_out * 256
    check(_lib.load().b2rl_conv1_wgrad(frames.data_ptr(), frames.shape[0], None if idx is None else idx.data_ptr(), n,
                                       gy.data_ptr(), None if relu_y is None else relu_y.data_ptr(), c_out,
                                       _wgrad_ws[key].data_ptr(), out.data_ptr(), int(bool(accumulate)), _stream_ptr(dev)))
    return out
