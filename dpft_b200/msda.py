"""Python face of the multi-scale deformable attention kernels.

Mirrors the interface of the extension DPFT imports as ``MultiScaleDeformableAttention``
(reference src/dprt/models/layers/ms_deform_attn.py:24): the two module-level functions
``ms_deform_attn_forward`` / ``ms_deform_attn_backward`` with the same argument order, the same
contiguity / CUDA-only contract, and errors raised as ``RuntimeError``.  ``MSDeformAttnFunction`` is the
autograd wrapper with the reference's ``apply`` signature (ms_deform_attn.py:27-68).
"""
from __future__ import annotations

import ctypes
from typing import Tuple

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import native


def _check_inputs(value, spatial_shapes, level_start_index, sampling_loc, attn_weight):
    native.require_cuda(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    for name, t in (("value", value), ("spatial_shapes", spatial_shapes),
                    ("level_start_index", level_start_index), ("sampling_loc", sampling_loc),
                    ("attn_weight", attn_weight)):
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
    if value.dim() != 4 or sampling_loc.dim() != 6 or attn_weight.dim() != 5:
        raise RuntimeError("expected value (B,S,M,D), sampling_loc (B,N,M,L,P,2), attn_weight (B,N,M,L,P)")
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError("spatial_shapes and level_start_index must be int64 tensors")
    if not (value.dtype == sampling_loc.dtype == attn_weight.dtype):
        raise RuntimeError("value, sampling_loc and attn_weight must share one floating dtype")
    B, S, M, D = value.shape
    Bl, N, Ml, L, P, two = sampling_loc.shape
    if (Bl, Ml, two) != (B, M, 2) or tuple(attn_weight.shape) != (B, N, M, L, P):
        raise RuntimeError("sampling_loc / attn_weight shapes do not match value")
    if tuple(spatial_shapes.shape) != (L, 2) or tuple(level_start_index.shape) != (L,):
        raise RuntimeError("spatial_shapes must be (L,2) and level_start_index (L,)")
    return B, S, M, D, N, L, P


def ms_deform_attn_forward(value: torch.Tensor, spatial_shapes: torch.Tensor,
                           level_start_index: torch.Tensor, sampling_loc: torch.Tensor,
                           attn_weight: torch.Tensor, im2col_step: int = 64) -> torch.Tensor:
    """(B,S,M,D), (L,2), (L,), (B,N,M,L,P,2), (B,N,M,L,P) -> (B,N,M*D).  ``im2col_step`` is ignored."""
    B, S, M, D, N, L, P = _check_inputs(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    lib = native.load_library()
    out = torch.empty((B, N, M * D), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device):
        st = lib.dpft_msda_forward(native.ptr(value), native.ptr(spatial_shapes), native.ptr(level_start_index),
                                   native.ptr(sampling_loc), native.ptr(attn_weight), native.ptr(out),
                                   B, S, M, D, N, L, P, native.dtype_code(value),
                                   native.stream_ptr(value.device))
    native.check(st, "dpft_msda_forward")
    native.count_launch(1 if B * N else 0)
    return out


def ms_deform_attn_backward(value: torch.Tensor, spatial_shapes: torch.Tensor,
                            level_start_index: torch.Tensor, sampling_loc: torch.Tensor,
                            attn_weight: torch.Tensor, grad_output: torch.Tensor,
                            im2col_step: int = 64) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Returns (grad_value, grad_sampling_loc, grad_attn_weight) in the dtype of ``value``."""
    B, S, M, D, N, L, P = _check_inputs(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    native.require_cuda(grad_output)
    grad_output = grad_output.contiguous()
    if tuple(grad_output.shape) != (B, N, M * D) or grad_output.dtype != value.dtype:
        raise RuntimeError("grad_output must be (B,N,M*D) in the dtype of value")
    lib = native.load_library()
    acc_dtype = torch.float64 if value.dtype == torch.float64 else torch.float32
    grad_value = torch.zeros(value.shape, dtype=acc_dtype, device=value.device)
    grad_loc = torch.empty_like(sampling_loc)
    grad_attn = torch.empty_like(attn_weight)
    with torch.cuda.device(value.device):
        st = lib.dpft_msda_backward(native.ptr(value), native.ptr(spatial_shapes), native.ptr(level_start_index),
                                    native.ptr(sampling_loc), native.ptr(attn_weight), native.ptr(grad_output),
                                    native.ptr(grad_value), native.ptr(grad_loc), native.ptr(grad_attn),
                                    B, S, M, D, N, L, P, native.dtype_code(value),
                                    native.stream_ptr(value.device))
    native.check(st, "dpft_msda_backward")
    native.count_launch(1 if B * N else 0)
    if grad_value.dtype != value.dtype:
        grad_value = grad_value.to(value.dtype)
    return grad_value, grad_loc, grad_attn


class MSDeformAttnFunction(Function):
    """Same ``apply(value, shapes, level_start_index, sampling_locations, attention_weights, im2col_step)``
    as the reference's autograd shim (ms_deform_attn.py:27-68)."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step=64):
        ctx.im2col_step = im2col_step
        value = value.contiguous()
        sampling_locations = sampling_locations.contiguous()
        attention_weights = attention_weights.contiguous()
        out = ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                                     sampling_locations, attention_weights, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        gv, gl, ga = ms_deform_attn_backward(value, shapes, lsi, loc, attn, grad_output.contiguous(),
                                             ctx.im2col_step)
        return gv, None, None, gl, ga, None


def install_plugin(name: str = "MultiScaleDeformableAttention") -> None:
    """Registers this module under the name DPFT imports (ms_deform_attn.py:24) so that the unmodified
    reference package runs on these kernels:  ``dpft_b200.msda.install_plugin(); import dprt``."""
    import sys
    sys.modules[name] = sys.modules[__name__]
