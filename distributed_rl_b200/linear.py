"""Bias-free nn.Linear at fp32 accuracy on the tensor cores (3xTF32, csrc/gemm.cu).

The reference's dense heads are `nn.Linear(..., bias=False)` (baseline/baseNetwork.py:77-79); on the
GPU PyTorch runs them as fp32 SIMT GEMMs.  `linear3x(x, w)` computes the same `x @ w.T` — forward,
input gradient and weight gradient — with every operand split into two TF32 terms and three
tcgen05 products per term pair, fp32 accumulation in TMEM (error ~2^-22 relative per product,
the same order as an fp32 FMA chain; tests/test_gpu_gemm.py pins it against fp64).
"""
from __future__ import annotations

import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def split_pack(x: torch.Tensor, transpose: bool, b_role: bool) -> torch.Tensor:
    """{hi, lo} TF32 operand image of a 2-D fp32 CUDA matrix (or of its transpose)."""
    if x.dim() != 2 or x.dtype != torch.float32 or not x.is_cuda:
        raise ValueError("split_pack expects a 2-D fp32 CUDA tensor")
    if x.stride(1) != 1:
        x = x.contiguous()
    rows, k = (x.shape[1], x.shape[0]) if transpose else (x.shape[0], x.shape[1])
    n = _lib.load().b2rl_gemm_packed_floats(rows, k, int(b_role))
    out = torch.empty(n, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().b2rl_gemm_split_pack(x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), int(transpose),
                                               int(b_role), out.data_ptr(), _stream()))
    return out


def gemm_packed(a: torch.Tensor, b: torch.Tensor, M: int, N: int, K: int, out: torch.Tensor | None = None,
                accumulate: bool = False) -> torch.Tensor:
    """C[M][N] = A[M][K] @ B[N][K]^T from packed operand images (A-role, B-role)."""
    if out is None:
        n_pad = (N + 3) // 4 * 4
        full = torch.empty(M, n_pad, dtype=torch.float32, device=a.device)
        out = full[:, :N]
        accumulate = False
    if out.stride(1) != 1 or out.stride(0) % 4:
        raise ValueError("output rows must be contiguous with a leading dimension divisible by 4")
    _lib.check(_lib.load().b2rl_gemm_tf32x3(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, out.stride(0),
                                           0 if accumulate else 1, _stream()))
    return out


class _Linear3x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        M, K = x.shape
        N = w.shape[0]
        y = gemm_packed(split_pack(x, False, False), split_pack(w, False, True), M, N, K)
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        M, K = x.shape
        N = w.shape[0]
        gy = gy.contiguous()
        gx = gw = None
        if ctx.needs_input_grad[0]:
            # dx[M][K] = gy[M][N] @ w[N][K]: contraction over N, the B operand is w^T ([K rows][N])
            gx = gemm_packed(split_pack(gy, False, False), split_pack(w, True, True), M, K, N)
        if ctx.needs_input_grad[1]:
            # dw[N][K] = gy^T[N][M] @ x[M][K]: contraction over M
            gw = gemm_packed(split_pack(gy, True, False), split_pack(x, True, True), N, K, M)
        return gx, gw


def linear3x(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """`torch.nn.functional.linear(x, w)` for 2-D fp32 CUDA `x` ([M][K]) and `w` ([N][K]), bias-free."""
    return _Linear3x.apply(x, w)
