"""Host side of the fused decoder kernels (dpft_b200/csrc/decoder.cu): weight packing and launch wrappers.

``pack_layer`` lays one ``MLFusion`` layer (reference src/dprt/models/fusers/mpfusion.py:16-263) out as the
flat fp32 image the kernel stages into shared memory; ``pack_head`` does the same for the view-reduction
Linear (mpfusion.py:393-395) and one ``LinearDetectionHead`` (src/dprt/models/heads/detection.py:149-275).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch

from . import native

C, NH = 16, 8
ACTIVATIONS = {"ReLU": 0, "Mish": 1, "GELU": 2}
REDUCTIONS = {"linear": 0, "mean": 1, "max": 2}


class DecoderView(ctypes.Structure):
    """Mirror of ``dpft_decoder_view`` in include/dpft_b200.h."""
    _fields_ = [("pyramid", ctypes.c_void_p), ("weights", ctypes.c_void_p), ("transform", ctypes.c_void_p),
                ("projection", ctypes.c_void_p), ("shape_hw", ctypes.c_void_p), ("use_transform", ctypes.c_void_p),
                ("S", ctypes.c_longlong), ("level_h", ctypes.c_int * 8), ("level_w", ctypes.c_int * 8),
                ("level_start", ctypes.c_longlong * 8)]


def layer_eligible(layer, d_model: int) -> Optional[str]:
    """None if the fused kernel covers this MLFusion layer, else the reason it does not."""
    att = layer.ms_deform_attn
    if d_model != C or layer.n_heads != NH:
        return f"d_model={d_model}, n_heads={layer.n_heads} (fused path: 16 / 8)"
    if att.n_points != 4 or not 1 <= att.n_levels <= 5:
        return f"n_points={att.n_points}, n_levels={att.n_levels} (fused path: 4 points, 1..5 levels)"
    if not layer.norm:
        return "norm=False"
    if layer.activation not in ACTIVATIONS:
        return f"activation {layer.activation}"
    if layer.d_ffn % 4 or layer.d_ffn > 256:
        return f"d_ffn={layer.d_ffn}"
    return None


def pack_layer(layer) -> torch.Tensor:
    """Flat fp32 image of one MLFusion layer; offsets must match ``LayerImage`` in decoder.cu."""
    att = layer.ms_deform_attn
    LP = att.n_levels * att.n_points
    f = lambda t: t.detach().float().reshape(-1).cpu()
    pad4 = torch.zeros(4)
    parts: List[torch.Tensor] = [f(layer.self_attn.in_proj_weight), f(layer.self_attn.in_proj_bias),
                                 f(layer.self_attn.out_proj.weight), f(layer.self_attn.out_proj.bias),
                                 f(layer.norm1.weight), f(layer.norm1.bias)]
    off_w = att.sampling_offsets.weight.detach().float().cpu().view(NH, LP * 2 * C)     # rows (m, l, p, xy)
    for m in range(NH):
        parts += [off_w[m], pad4]
    parts.append(f(att.sampling_offsets.bias))
    att_w = att.attention_weights.weight.detach().float().cpu().view(NH, LP * C)        # rows (m, l, p)
    for m in range(NH):
        parts += [att_w[m], pad4]
    parts += [f(att.attention_weights.bias), f(att.value_proj.weight), f(att.value_proj.bias),
              f(att.output_proj.weight), f(att.output_proj.bias), f(layer.norm2.weight), f(layer.norm2.bias),
              f(layer.norm3.weight), f(layer.norm3.bias), f(layer.ffn1.weight), f(layer.ffn1.bias),
              f(layer.ffn2.weight), f(layer.ffn2.bias)]
    img = torch.cat(parts)
    assert img.numel() % 4 == 0
    return img


HEAD_LANES16 = 0x100      # include/dpft_b200.h: DPFT_HEAD_LANES16
HEAD_LANES1 = 0x200       # include/dpft_b200.h: DPFT_HEAD_LANES1


def pack_head(reduction_layer, head, reduction: str) -> torch.Tensor:
    """Reduction weight (16, V*16) then the centre / size / angle / class branches (3 Linear weights each)."""
    f = lambda t: t.detach().float().reshape(-1).cpu()
    parts = []
    if reduction == "linear":
        parts.append(f(reduction_layer.weight))
    for name in ("center", "size", "angle", "class"):
        seq = head.layers[f"{name}_head"]
        linears = [m for m in seq if isinstance(m, torch.nn.Linear)]
        if len(linears) != 3 or any(l.bias is not None for l in linears):
            raise ValueError("fused head expects three bias-free Linear layers per branch")
        parts += [f(l.weight) for l in linears]
    return torch.cat(parts)


def head_eligible(head) -> Optional[str]:
    from .models.head import LinearDetectionHead
    if not isinstance(head, LinearDetectionHead):
        return f"head type {type(head).__name__}"
    if head.in_channels != C or head.num_reg_layers != 3 or head.num_cls_layers != 3 or head.bias:
        return "head is not 3 bias-free 16-wide layers per branch"
    if not 1 <= head.num_classes <= 8:
        return f"num_classes={head.num_classes}"
    return None


# bench.py sets this to a list to collect (gathered bytes, start_event, end_event) of every decoder-layer launch of one step
PROFILE = None


def layer_forward(views: Sequence[DecoderView], query: torch.Tensor, pos: torch.Tensor, center: torch.Tensor,
                  out: torch.Tensor, B: int, N: int, L: int, P: int, d_ffn: int, act: int, weight_floats: int,
                  pyramid_dtype: torch.dtype = torch.float32) -> None:
    lib = native.load_library()
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    V = len(views)
    arr = (DecoderView * V)(*views)
    qs = 0 if query.dim() == 2 else N * C
    cs = 0 if center.dim() == 2 else N * 3
    st = lib.dpft_decoder_layer_forward(ctypes.cast(arr, ctypes.c_void_p), V, native.ptr(query), qs, native.ptr(pos),
                                        native.ptr(center), cs, native.ptr(out), B, N, L, P, d_ffn, act, weight_floats,
                                        native._DTYPE_CODE[pyramid_dtype], native.stream_ptr(out.device))
    native.check(st, "dpft_decoder_layer_forward")
    native.count_launch()
    if prof is not None:
        e1.record()
        # per sample 4 bilinear corners x one 16-channel pyramid row (32 B in f16, 64 B in f32), M = 8 heads
        row = 16 * (2 if pyramid_dtype == torch.float16 else 4)
        prof.append((float(B) * len(views) * N * 8 * L * P * 4 * row, e0, e1))


def head_forward(views: torch.Tensor, weights: torch.Tensor, center_in: torch.Tensor, query_out: torch.Tensor,
                 center_out: torch.Tensor, size_out, angle_out, class_out, B: int, V: int, N: int, n_cls: int,
                 reduction: int, lanes16: Optional[bool] = None) -> None:
    """``lanes16``: None = the library's default kernel; True / False force the sixteen-lanes-per-query / one-thread-per-query
    kernel (DPFT_HEAD_LANES16 / DPFT_HEAD_LANES1 in include/dpft_b200.h; bit-identical results)."""
    lib = native.load_library()
    if lanes16 is not None:
        reduction |= HEAD_LANES16 if lanes16 else HEAD_LANES1
    cs = 0 if center_in.dim() == 2 else N * 3
    st = lib.dpft_decoder_head_forward(native.ptr(views), native.ptr(weights), native.ptr(center_in), cs,
                                       native.ptr(query_out), native.ptr(center_out), native.ptr(size_out),
                                       native.ptr(angle_out), native.ptr(class_out), B, V, N, n_cls, reduction,
                                       weights.numel(), native.stream_ptr(views.device))
    native.check(st, "dpft_decoder_head_forward")
    native.count_launch()
