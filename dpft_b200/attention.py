"""Host wrapper of ``dpft_self_attention_forward`` (csrc/attention.cu): the scaled-dot-product core of the decoder's
nn.MultiheadAttention (reference src/dprt/models/fusers/mpfusion.py:56-57, :122-148) as a tcgen05 flash-attention kernel.

Forward only (inference): under autograd the decoder keeps torch's differentiable attention.  CUDA only, no fallback.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

from . import native

MAX_HEAD_DIM = 64


def _strides(t: torch.Tensor, H: int, D: int):
    """t viewed as (B, N, H*D) with the last dimension contiguous -> (row stride, batch stride) in elements."""
    if t.dim() != 3 or t.shape[-1] != H * D or t.stride(-1) != 1:
        raise RuntimeError("self_attention: q/k/v must be (B, N, H*D) tensors with a contiguous last dimension")
    return t.stride(1), t.stride(0)


def self_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, n_heads: int, scale: Optional[float] = None,
                   precise: Optional[bool] = None) -> torch.Tensor:
    """q, k, v (B, N, C) (may be strided slices of a packed projection), heads interleaved as C = n_heads * D
    -> softmax(scale * q k^T) v per head, (B, N, C) contiguous, same dtype.

    ``precise`` (fp32 inputs only, default True): split-f16 operands, three MMAs per product, fp32-grade results."""
    native.require_cuda(q, k, v)
    B, N, C = q.shape
    if C % n_heads:
        raise RuntimeError(f"self_attention: {C} channels do not divide into {n_heads} heads")
    D = C // n_heads
    if D > MAX_HEAD_DIM:
        raise RuntimeError(f"self_attention: head dimension {D} > {MAX_HEAD_DIM}")
    if not (q.dtype == k.dtype == v.dtype) or q.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        raise RuntimeError("self_attention: q, k, v must share one of float32 / float16 / bfloat16")
    if k.shape != q.shape or v.shape != q.shape:
        raise RuntimeError("self_attention: q, k, v must have the same shape (self-attention)")
    if precise is None:
        precise = q.dtype == torch.float32
    if precise and q.dtype != torch.float32:
        raise RuntimeError("self_attention: precise mode takes float32 inputs")
    (qr, qb), (kr, kb), (vr, vb) = _strides(q, n_heads, D), _strides(k, n_heads, D), _strides(v, n_heads, D)
    out = torch.empty((B, N, C), dtype=q.dtype, device=q.device)
    scale = 1.0 / math.sqrt(D) if scale is None else float(scale)
    with torch.cuda.device(q.device):
        st = native.load_library().dpft_self_attention_forward(
            native.ptr(q), native.ptr(k), native.ptr(v), native.ptr(out), B, n_heads, N, D, qr, kr, vr, qb, kb, vb,
            scale, native.dtype_code(q), int(bool(precise)), native.stream_ptr(q.device))
    native.check(st, "dpft_self_attention_forward")
    native.count_launch()
    return out


def mha_eligible(mha: torch.nn.MultiheadAttention, x: torch.Tensor) -> bool:
    """True when ``multihead_self_attention`` reproduces ``mha(q, k, v, need_weights=False)`` for this call."""
    return (x.is_cuda and not torch.is_grad_enabled() and mha._qkv_same_embed_dim and mha.batch_first
            and mha.in_proj_bias is not None and mha.bias_k is None and not mha.add_zero_attn
            and (not mha.training or mha.dropout == 0.0) and mha.embed_dim // mha.num_heads <= MAX_HEAD_DIM
            and x.dtype in (torch.float32, torch.float16, torch.bfloat16))


def multihead_self_attention(mha: torch.nn.MultiheadAttention, qk: torch.Tensor, value: torch.Tensor) -> torch.Tensor:
    """``mha(query=qk, key=qk, value=value, need_weights=False)[0]`` (the call at reference mpfusion.py:139): packed q/k
    in-projection and the value in-projection as two GEMMs, the attention core in the tcgen05 kernel on strided views of
    their outputs (no head split / transpose copies), then the output projection."""
    C = mha.embed_dim
    w, b = mha.in_proj_weight, mha.in_proj_bias
    qk_p = F.linear(qk, w[:2 * C], b[:2 * C])                   # (B, N, 2C): q | k
    v_p = F.linear(value, w[2 * C:], b[2 * C:])                 # (B, N, C)
    ctx = self_attention(qk_p[..., :C], qk_p[..., C:], v_p, mha.num_heads)
    return mha.out_proj(ctx)
