"""FusedRMSprop — torch.optim.RMSprop's update for the learners' configs (cfg/ape_x.json centered,
cfg/impala.json plain) as one libb2rl launch that also zeroes the gradients and produces the
reference's diagnostic gradient "norm" (APE_X/Learner.py:123-138).  momentum and weight_decay are
0 in every shipped config and are not supported here (falls back to torch.optim.RMSprop)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check


class FusedRMSprop:
    def __init__(self, params, lr, alpha=0.99, eps=1e-8, centered=False):
        self.params = [p for p in params]
        assert 1 <= len(self.params) <= 24
        self.lr, self.alpha, self.eps, self.centered = float(lr), float(alpha), float(eps), bool(centered)
        dev = self.params[0].device
        self.device = dev
        for p in self.params:
            assert p.is_cuda and p.dtype == torch.float32 and _dense(p), "dense fp32 CUDA parameters only"
            if p.grad is None or p.grad.stride() != p.stride():
                # gradients must share the parameter's element order (channels_last weights!): autograd
                # accumulates in place into this buffer from then on
                old = p.grad
                p.grad = torch.zeros_like(p, memory_format=torch.preserve_format)
                if old is not None:
                    p.grad.copy_(old)
        self.square_avg = [torch.zeros_like(p, memory_format=torch.preserve_format) for p in self.params]
        self.grad_avg = [torch.zeros_like(p, memory_format=torch.preserve_format) for p in self.params] \
            if self.centered else None
        n = len(self.params)
        self._scratch = torch.zeros(n, dtype=torch.float64, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        arr = C.c_void_p * n
        self._p = arr(*[p.data_ptr() for p in self.params])
        self._sq = arr(*[t.data_ptr() for t in self.square_avg])
        self._ga = arr(*[t.data_ptr() for t in self.grad_avg]) if self.centered else None
        self._numel = (C.c_int64 * n)(*[p.numel() for p in self.params])
        self._arr = arr

    def step(self, want_norm: bool = True) -> torch.Tensor:
        """Update + zero_grad.  Returns the device scalar sqrt(sum_i ||g_i||_2) (of the pre-step grads)."""
        g = self._arr(*[p.grad.data_ptr() for p in self.params])
        for p in self.params:
            assert p.grad.stride() == p.stride()
        check(_lib.load().b2rl_rmsprop_step(
            self._p, g, self._sq, self._ga, self._numel, len(self.params), self.lr, self.alpha, self.eps,
            int(self.centered), self._scratch.data_ptr(), self.grad_norm.data_ptr() if want_norm else None,
            torch.cuda.current_stream(self.device).cuda_stream))
        return self.grad_norm

    def zero_grad(self, set_to_none: bool = False) -> None:
        for p in self.params:
            p.grad.zero_()


def _dense(t: torch.Tensor) -> bool:
    """True if t's storage is a dense permutation (contiguous in some dim order)."""
    return t.is_contiguous() or t.is_contiguous(memory_format=torch.channels_last) or \
        (t.numel() == t.untyped_storage().nbytes() // t.element_size())
