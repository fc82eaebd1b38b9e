"""Host wrappers of the training kernels in libdpft_b200.so (csrc/conv_wgrad.cu, csrc/train_ops.cu).

Autograd pieces of the torchvision Bottleneck blocks the reference trains (src/dprt/models/backbones/resnet.py:54-55,101;
src/dprt/training/trainer.py:125-133): weight gradient, data gradient (through ``conv2d_nhwc`` on re-laid-out weights),
BatchNorm with batch statistics forward/backward, max-pool backward.  NHWC, 16-bit activations, CUDA only.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import native
from .conv import conv2d_nhwc


def _lib():
    return native.load_library()


def _check16(*ts: torch.Tensor) -> None:
    native.require_cuda(*ts)
    for t in ts:
        if t.dtype not in (torch.bfloat16, torch.float16) or not t.is_contiguous():
            raise RuntimeError("dpft_b200.train_ops: activations must be contiguous bfloat16 / float16 CUDA tensors")


def conv2d_wgrad(x: torch.Tensor, dy: torch.Tensor, R: int, S: int, stride: int, pad: int,
                 out: Optional[torch.Tensor] = None, splits: int = 0) -> torch.Tensor:
    """x (B,H,W,Cin), dy (B,P,Q,Cout) 16-bit -> fp32 (Cout,R,S,Cin), ADDED into ``out`` (zeros if not given)."""
    _check16(x, dy)
    B, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    P, Q = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    if tuple(dy.shape) != (B, P, Q, Cout) or dy.dtype != x.dtype:
        raise RuntimeError(f"conv2d_wgrad: dy has shape {tuple(dy.shape)}, expected {(B, P, Q, Cout)} of {x.dtype}")
    if out is None:
        out = torch.zeros((Cout, R, S, Cin), dtype=torch.float32, device=x.device)
    elif out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != Cout * R * S * Cin:
        raise RuntimeError("conv2d_wgrad: out must be a contiguous fp32 tensor of Cout*R*S*Cin elements")
    with torch.cuda.device(x.device):
        st = _lib().dpft_conv2d_wgrad(native.ptr(x), native.ptr(dy), native.ptr(out), B, H, W, Cin, Cout, R, S, stride, pad,
                                      splits, native.dtype_code(x), native.stream_ptr(x.device))
    native.check(st, "dpft_conv2d_wgrad")
    native.count_launch()
    return out


def zero_insert2(src: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """(B,P,Q,C) -> (B,H,W,C) with src at the even positions and zeros elsewhere."""
    _check16(src)
    B, P, Q, C = src.shape
    up = torch.empty((B, H, W, C), dtype=src.dtype, device=src.device)
    with torch.cuda.device(src.device):
        st = _lib().dpft_zero_insert2_nhwc(native.ptr(src), native.ptr(up), B, H, W, C, P, Q, native.stream_ptr(src.device))
    native.check(st, "dpft_zero_insert2_nhwc")
    native.count_launch()
    return up


def conv2d_dgrad(dy: torch.Tensor, w_dgrad: torch.Tensor, zero_bias: torch.Tensor, in_hw: Tuple[int, int], stride: int, pad: int,
                 residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Data gradient of a convolution: dy (B,P,Q,Cout), w_dgrad (Cin,R,S,Cout) (taps flipped) -> (B,H,W,Cin) (+ residual).

    stride 1: one ``conv2d_nhwc`` with padding R-1-pad.  stride 2: 1x1 -> GEMM at the coarse resolution then zero
    insertion; 3x3 -> zero insertion of dy, then the stride-1 form."""
    Cin, R, S, Cout = w_dgrad.shape
    H, W = in_hw
    if stride == 1:
        return conv2d_nhwc(dy, w_dgrad, zero_bias, 1, R - 1 - pad, False, residual)
    if stride != 2:
        raise NotImplementedError("conv2d_dgrad: stride must be 1 or 2")
    if R == 1:
        small = conv2d_nhwc(dy, w_dgrad, zero_bias, 1, 0, False, None)
        up = zero_insert2(small, H, W)
        if residual is not None:
            up += residual
        return up
    up = zero_insert2(dy, H, W)
    return conv2d_nhwc(up, w_dgrad, zero_bias, 1, R - 1 - pad, False, residual)


BN_MAX_PARTS = 296          # DPFT_BN_MAX_PARTS
_workspaces = {}


def bn_workspace(device) -> torch.Tensor:
    """Partial-sum scratch of the BatchNorm calls of ONE stream (calls on a stream are ordered; the views of a training step
    run on forked streams and must not share it)."""
    dev = torch.device(device)
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        ws = _workspaces[key] = torch.empty(BN_MAX_PARTS * 2 * 2048, dtype=torch.float32, device=device)
    return ws


class BNState:
    """fp32 per-channel buffers of one BatchNorm call: [scale, shift, mean, invstd] rows of one (4, C) tensor."""

    def __init__(self, buf: torch.Tensor):
        self.buf = buf
        self.scale, self.shift, self.mean, self.invstd = buf.unbind(0)


def bn_forward(y: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, running_mean: Optional[torch.Tensor],
               running_var: Optional[torch.Tensor], momentum: float, eps: float, relu: bool,
               residual: Optional[torch.Tensor] = None, state_buf: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, BNState]:
    """z = [relu](batch_norm_train(y) (+ residual)); y (..., C) 16-bit NHWC.  Returns (z, state for the backward)."""
    _check16(y)
    C = y.shape[-1]
    M = y.numel() // C
    if C > 2048:
        raise RuntimeError("bn_forward: at most 2048 channels")
    if state_buf is None:
        state_buf = torch.empty((4, C), dtype=torch.float32, device=y.device)
    st8 = BNState(state_buf)
    lib = _lib()
    code, stream = native.dtype_code(y), native.stream_ptr(y.device)
    with torch.cuda.device(y.device):
        native.check(lib.dpft_bn_forward_stats(native.ptr(y), native.ptr(bn_workspace(y.device)), native.ptr(gamma), native.ptr(beta),
                                               native.ptr(running_mean), native.ptr(running_var), float(momentum), float(eps), M, C,
                                               native.ptr(st8.scale), native.ptr(st8.shift), native.ptr(st8.mean), native.ptr(st8.invstd),
                                               code, stream), "dpft_bn_forward_stats")
        z = torch.empty_like(y)
        native.check(lib.dpft_bn_apply(native.ptr(y), native.ptr(st8.scale), native.ptr(st8.shift), native.ptr(residual), native.ptr(z),
                                       M, C, int(relu), code, stream), "dpft_bn_apply")
    native.count_launch(3)
    return z, st8


def bn_backward(dz: torch.Tensor, z: Optional[torch.Tensor], y: torch.Tensor, state: BNState, gamma: torch.Tensor, relu: bool,
                dgamma: torch.Tensor, dbeta: torch.Tensor, want_g: bool = False,
                sums: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Backward of ``bn_forward``: returns (dy, g) with g = dz * [z > 0] when ``want_g`` (gradient of the residual branch);
    dgamma / dbeta (fp32) are added to."""
    _check16(dz, y)
    C = y.shape[-1]
    M = y.numel() // C
    if sums is None:
        sums = torch.empty((2, C), dtype=torch.float32, device=y.device)
    dy = torch.empty_like(y)
    g = torch.empty_like(y) if want_g else None
    lib = _lib()
    code, stream = native.dtype_code(y), native.stream_ptr(y.device)
    with torch.cuda.device(y.device):
        native.check(lib.dpft_bn_backward_reduce(native.ptr(dz), native.ptr(z), native.ptr(y), native.ptr(state.mean),
                                                 native.ptr(state.invstd), native.ptr(bn_workspace(y.device)), native.ptr(sums[0]),
                                                 native.ptr(sums[1]), native.ptr(dgamma), native.ptr(dbeta), M, C, int(relu),
                                                 code, stream), "dpft_bn_backward_reduce")
        native.check(lib.dpft_bn_backward_apply(native.ptr(dz), native.ptr(z), native.ptr(y), native.ptr(state.mean),
                                                native.ptr(state.invstd), native.ptr(gamma), native.ptr(sums[0]), native.ptr(sums[1]),
                                                native.ptr(dy), native.ptr(g), M, C, int(relu), code, stream), "dpft_bn_backward_apply")
    native.count_launch(3)
    return dy, g


def stem_wgrad(x: torch.Tensor, dy: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x (B,H,W,Cin) fp32, dy (B,P,Q,64) 16-bit -> fp32 [7][7][Cin][64], ADDED into ``out`` (zeros if not given)."""
    _check16(dy)
    native.require_cuda(x)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise RuntimeError("stem_wgrad: x must be a contiguous fp32 tensor")
    B, H, W, Cin = x.shape
    if out is None:
        out = torch.zeros((7, 7, Cin, 64), dtype=torch.float32, device=x.device)
    n = 49 * Cin * 64
    sms = torch.cuda.get_device_properties(x.device).multi_processor_count
    ws = torch.empty(2 * sms * n, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        st = _lib().dpft_stem_conv7x7_wgrad(native.ptr(x), native.ptr(dy), native.ptr(ws), ws.numel(), native.ptr(out), B, H, W, Cin,
                                            native.dtype_code(dy), native.stream_ptr(x.device))
    native.check(st, "dpft_stem_conv7x7_wgrad")
    native.count_launch(2)
    return out


def maxpool_backward(x: torch.Tensor, pooled: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    """x (B,H,W,C) input of the 3x3/2 max-pool, pooled (B,P,Q,C) its output, dy (B,P,Q,C) -> dx (B,H,W,C)."""
    _check16(x, pooled, dy)
    B, H, W, C = x.shape
    dx = torch.empty_like(x)
    with torch.cuda.device(x.device):
        st = _lib().dpft_maxpool3x3s2_backward(native.ptr(x), native.ptr(pooled), native.ptr(dy), native.ptr(dx), B, H, W, C, native.dtype_code(x),
                                               native.stream_ptr(x.device))
    native.check(st, "dpft_maxpool3x3s2_backward")
    native.count_launch()
    return dx


# ---- batched weight re-layout -------------------------------------------------------------------------------------------
_ENTRY = np.dtype([("src", np.uint64), ("fwd", np.uint64), ("dgrad", np.uint64), ("Cout", np.int32), ("Cin", np.int32),
                   ("R", np.int32), ("S", np.int32), ("offset", np.int64)])
assert _ENTRY.itemsize == 48


class WeightPacker:
    """All convolution weights of a model: fp32 masters (Cout,Cin,R,S) -> one 16-bit buffer with the forward layout
    (Cout,R,S,Cin) and one with the data-gradient layout (Cin,R,S,Cout, taps flipped), refreshed by ONE launch per step."""

    def __init__(self, weights: Sequence[torch.Tensor], dtype: torch.dtype, need_dgrad: Sequence[bool]):
        self.weights = list(weights)
        self.dtype = dtype
        dev = self.weights[0].device
        sizes = [w.numel() for w in self.weights]
        self.total = int(sum(sizes))
        pad = lambda n: (n + 63) // 64 * 64                    # keep every layer's operand 128-byte aligned
        offs, o = [], 0
        for n in sizes:
            offs.append(o)
            o += pad(n)
        self.fwd_flat = torch.empty(o, dtype=dtype, device=dev)
        self.dgrad_flat = torch.empty(o, dtype=dtype, device=dev)
        self.fwd: List[torch.Tensor] = []
        self.dgrad: List[Optional[torch.Tensor]] = []
        for w, n, off, nd in zip(self.weights, sizes, offs, need_dgrad):
            co, ci, r, s = w.shape
            self.fwd.append(self.fwd_flat[off:off + n].view(co, r, s, ci))
            self.dgrad.append(self.dgrad_flat[off:off + n].view(ci, r, s, co) if nd else None)
        self._ptrs = None
        self._table = None

    def _build_table(self) -> None:
        tab = np.zeros(len(self.weights), dtype=_ENTRY)
        o = 0
        for i, w in enumerate(self.weights):
            if w.dtype != torch.float32 or not w.is_contiguous():
                raise RuntimeError("WeightPacker: master weights must be contiguous fp32")
            co, ci, r, s = w.shape
            tab[i] = (w.data_ptr(), self.fwd[i].data_ptr(), 0 if self.dgrad[i] is None else self.dgrad[i].data_ptr(), co, ci, r, s, o)
            o += w.numel()
        self._table = torch.from_numpy(tab.view(np.uint8).copy()).to(self.weights[0].device)
        self._ptrs = [w.data_ptr() for w in self.weights]

    def refresh(self) -> None:
        if self._ptrs != [w.data_ptr() for w in self.weights]:
            self._build_table()
        dev = self.weights[0].device
        with torch.cuda.device(dev):
            st = _lib().dpft_pack_conv_weights(native.ptr(self._table), len(self.weights), self.total,
                                               native._DTYPE_CODE[self.dtype], native.stream_ptr(dev))
        native.check(st, "dpft_pack_conv_weights")
        native.count_launch()
