"""Host side of the sm_100a convolution kernels: BatchNorm folding, weight re-layout, launch wrappers.

Inference form of the torchvision ResNet blocks the reference runs (src/dprt/models/backbones/resnet.py:101):
``conv -> BatchNorm2d(eval) [-> + identity] [-> ReLU]`` becomes one call of ``dpft_conv2d_nhwc`` with
``w' = w * gamma / sqrt(var + eps)`` and ``bias' = beta - mean * gamma / sqrt(var + eps)``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from . import native

# bench.py sets this to a list to collect (flops, bytes, start_event, end_event) of every conv launch of one step
PROFILE = None

# Cap on the persistent grid of the launches issued inside `cta_budget(n)` (0 = one CTA per SM): the engine runs the small
# side views (radar) under it so that their latency-bound layers do not take every SM from the view on the critical path.
_MAX_CTAS = 0


class cta_budget:
    def __init__(self, n: int):
        self.n = int(n or 0)

    def __enter__(self):
        global _MAX_CTAS
        self.saved, _MAX_CTAS = _MAX_CTAS, self.n
        return self

    def __exit__(self, *exc):
        global _MAX_CTAS
        _MAX_CTAS = self.saved
        return False


def fold_conv_bn(conv: nn.Conv2d, bn: Optional[nn.Module]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (weight (Cout, R, S, Cin) fp32, bias (Cout,) fp32) of conv followed by eval-mode BatchNorm."""
    w = conv.weight.detach().float()
    cout = w.shape[0]
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(cout, device=w.device)
    if bn is not None:
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
        w = w * scale.view(-1, 1, 1, 1)
        b = (b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
    return w.permute(0, 2, 3, 1).contiguous(), b.contiguous()


class FoldedConv:
    """One ``conv+bn`` of the backbone prepared for the tcgen05 kernel (weights bf16 [Cout,R,S,Cin], bias fp32)."""

    def __init__(self, conv: nn.Conv2d, bn: Optional[nn.Module], device, dtype: torch.dtype = torch.bfloat16):
        w, b = fold_conv_bn(conv, bn)
        self.weight = w.to(device=device, dtype=dtype).contiguous()
        self.bias = b.to(device=device, dtype=torch.float32).contiguous()
        self.cout, self.r, self.s, self.cin = self.weight.shape
        self.stride = conv.stride[0]
        self.pad = conv.padding[0]

    def out_hw(self, h: int, w: int) -> Tuple[int, int]:
        return ((h + 2 * self.pad - self.r) // self.stride + 1, (w + 2 * self.pad - self.s) // self.stride + 1)

    def __call__(self, x: torch.Tensor, relu: bool, residual: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None, block_n: int = 0) -> torch.Tensor:
        return conv2d_nhwc(x, self.weight, self.bias, self.stride, self.pad, relu, residual, out, block_n)


def conv2d_nhwc(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, stride: int, pad: int, relu: bool,
                     residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                     block_n: int = 0, cluster_mode: int = 0) -> torch.Tensor:
    """x (B,H,W,Cin) bf16|f16 contiguous, weight (Cout,R,S,Cin) same dtype, bias (Cout,) fp32 -> (B,P,Q,Cout)."""
    native.require_cuda(x, weight, bias)
    if x.dtype not in (torch.bfloat16, torch.float16) or weight.dtype != x.dtype or bias.dtype != torch.float32:
        raise RuntimeError("conv2d_nhwc: x and weight must share bfloat16 or float16, bias must be float32")
    if not (x.is_contiguous() and weight.is_contiguous() and bias.is_contiguous()):
        raise RuntimeError("conv2d_nhwc: tensors have to be contiguous")
    B, H, W, Cin = x.shape
    Cout, R, S, Cin_w = weight.shape
    if Cin_w != Cin:
        raise RuntimeError(f"conv2d_nhwc: weight expects {Cin_w} input channels, got {Cin}")
    P, Q = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    if out is None:
        out = torch.empty((B, P, Q, Cout), dtype=x.dtype, device=x.device)
    if residual is not None and (residual.shape != out.shape or not residual.is_contiguous()
                                 or residual.dtype != x.dtype):
        raise RuntimeError("conv2d_nhwc: residual must be a contiguous tensor of the output shape and dtype")
    lib = native.load_library()
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    with torch.cuda.device(x.device):
        st = lib.dpft_conv2d_nhwc_ex(native.ptr(x), native.ptr(weight), native.ptr(bias), native.ptr(residual),
                                     native.ptr(out), B, H, W, Cin, Cout, R, S, stride, pad, int(relu), block_n, cluster_mode,
                                     native.dtype_code(x), _MAX_CTAS, native.stream_ptr(x.device))
    native.check(st, "dpft_conv2d_nhwc")
    native.count_launch()
    if prof is not None:
        e1.record()
        M = B * P * Q
        prof.append((2.0 * M * Cout * R * S * Cin,
                     2.0 * (x.numel() + weight.numel() + M * Cout * (2 if residual is not None else 1)), e0, e1))
    return out


def conv2d_nhwc_pair(xs, weights, biases, stride: int, pad: int, relu: bool, residuals=(None, None)):
    """Two convolutions of IDENTICAL shape (different inputs / weights / residuals) in ONE launch (``dpft_conv2d_nhwc_pair``):
    -> (y0, y1).  Falls back to two launches for a layer the pair entry does not serve (halo / weight-stationary kernels)."""
    x0, x1 = xs
    w0, w1 = weights
    if x0.shape != x1.shape or w0.shape != w1.shape or x0.dtype != x1.dtype or (residuals[0] is None) != (residuals[1] is None):
        raise RuntimeError("conv2d_nhwc_pair: the two problems must have identical shapes and types")
    native.require_cuda(x0, x1, w0, w1)
    B, H, W, Cin = x0.shape
    Cout, R, S, _ = w0.shape
    P, Q = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    outs = [torch.empty((B, P, Q, Cout), dtype=x0.dtype, device=x0.device) for _ in range(2)]
    lib = native.load_library()
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    with torch.cuda.device(x0.device):
        st = lib.dpft_conv2d_nhwc_pair(native.ptr(x0), native.ptr(w0), native.ptr(biases[0]), native.ptr(residuals[0]), native.ptr(outs[0]),
                                       native.ptr(x1), native.ptr(w1), native.ptr(biases[1]), native.ptr(residuals[1]), native.ptr(outs[1]),
                                       B, H, W, Cin, Cout, R, S, stride, pad, int(relu), native.dtype_code(x0), _MAX_CTAS,
                                       native.stream_ptr(x0.device))
    if st == -2:                                  # DPFT_ERR_UNSUPPORTED: not a layer of the generic kernel
        y0 = conv2d_nhwc(x0, w0, biases[0], stride, pad, relu, residuals[0], outs[0])
        y1 = conv2d_nhwc(x1, w1, biases[1], stride, pad, relu, residuals[1], outs[1])
        return y0, y1
    native.check(st, "dpft_conv2d_nhwc_pair")
    native.count_launch()
    if prof is not None:
        e1.record()
        M = B * P * Q
        prof.append((2 * 2.0 * M * Cout * R * S * Cin,
                     2 * 2.0 * (x0.numel() + w0.numel() + M * Cout * (2 if residuals[0] is not None else 1)), e0, e1))
    return outs[0], outs[1]
