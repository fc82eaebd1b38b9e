"""Native (sm_100a) feature path of one view: ResNet backbone -> FPN neck -> positional embedding -> pyramid.

Replaces, for inference, what ``DPRT.forward`` does per input at reference src/dprt/models/dprt.py:219-231
(``backbones[input]`` resnet.py:80-107, skip link, ``necks[input]`` fpn.py:70-83, ``embeddings[input]``
sinusoidal.py:137-153) and the per-layer ``torch.cat`` of mpfusion.py:179: the result is ONE buffer
(B, S, 16) fp32 per view, finest level first.

Arithmetic: activations NHWC bf16 with fp32 accumulation (tcgen05), BatchNorm folded into the convolutions,
FPN / embedding / pyramid in fp32.  Weights are re-laid-out once when the engine is built.
"""
from __future__ import annotations

import os

from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import native
from .conv import FoldedConv, fold_conv_bn
from .models.backbone import Backbone
from .models.embedding import MultiLevelSinusoidalEmbedding
from .models.neck import FPN

FC = 16


def _lib():
    return native.load_library()


def stem_pack_weights(w: torch.Tensor) -> torch.Tensor:
    """[7][7][Cin][64] fp32 -> the f16 operand images of the tcgen05 stem kernels: 57344 bytes in the 8-channels-per-tap
    layout, followed (Cin == 3) by 28672 bytes in the 4-channels-per-tap layout of the streaming kernel."""
    packed = torch.empty(57344 + 28672, dtype=torch.uint8, device=w.device)
    st = _lib().dpft_stem_pack_weights(native.ptr(w), native.ptr(packed), w.shape[2], native.stream_ptr(w.device))
    native.check(st, "dpft_stem_pack_weights")
    return packed


def fpn_pack_weights(w: torch.Tensor) -> torch.Tensor:
    """[3][3][16][16] fp32 -> the f16 operand image of the tcgen05 FPN output kernel (4608 bytes)."""
    packed = torch.empty(4608, dtype=torch.uint8, device=w.device)
    st = _lib().dpft_fpn_pack_weights(native.ptr(w), native.ptr(packed), native.stream_ptr(w.device))
    native.check(st, "dpft_fpn_pack_weights")
    return packed


def stem_forward(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, dtype: torch.dtype = torch.bfloat16,
                 impl: int = 0, w_packed: Optional[torch.Tensor] = None, relu: bool = True) -> torch.Tensor:
    B, H, W, Cin = x.shape
    y = torch.empty((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, 64), dtype=dtype, device=x.device)
    if w_packed is None and impl == 2:
        w_packed = stem_pack_weights(w)
    # uint8 frames (include/dpft_b200.h: DPFT_RAW_U8) are converted on load inside the row-streaming kernel
    cin_code = Cin | RAW_U8 if x.dtype == torch.uint8 else Cin
    st = _lib().dpft_stem_conv7x7_forward_ex(native.ptr(x), native.ptr(w), native.ptr(w_packed), native.ptr(bias), native.ptr(y), B, H, W,
                                             cin_code, native.dtype_code(y), impl, int(relu), native.stream_ptr(x.device))
    native.check(st, "dpft_stem_conv7x7_forward_ex")
    native.count_launch()
    return y


def maxpool_forward(x: torch.Tensor) -> torch.Tensor:
    B, H, W, Cc = x.shape
    y = torch.empty((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, Cc), dtype=x.dtype, device=x.device)
    st = _lib().dpft_maxpool3x3s2_nhwc(native.ptr(x), native.ptr(y), B, H, W, Cc, native.dtype_code(x),
                                       native.stream_ptr(x.device))
    native.check(st, "dpft_maxpool3x3s2_nhwc")
    native.count_launch()
    return y


def lateral_forward(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, coarse: Optional[torch.Tensor]) -> torch.Tensor:
    B, H, W, Cin = x.shape
    out = torch.empty((B, H, W, FC), dtype=torch.float32, device=x.device)
    hc, wc = (coarse.shape[1], coarse.shape[2]) if coarse is not None else (0, 0)
    st = _lib().dpft_fpn_lateral_forward(native.ptr(x), native.ptr(w), native.ptr(bias), native.ptr(coarse), hc, wc,
                                         native.ptr(out), B, H, W, Cin, native.dtype_code(x), native.stream_ptr(x.device))
    native.check(st, "dpft_fpn_lateral_forward")
    native.count_launch()
    return out


def fpn_output_forward(pyramid: torch.Tensor, start: int, H: int, W: int, w: torch.Tensor, bias: torch.Tensor,
                       pos_y: torch.Tensor, pos_x: torch.Tensor, inner: Optional[torch.Tensor] = None,
                       raw: Optional[torch.Tensor] = None, lat_w: Optional[torch.Tensor] = None,
                       lat_b: Optional[torch.Tensor] = None, coarse: Optional[torch.Tensor] = None, impl: int = 0,
                       w_packed: Optional[torch.Tensor] = None, lat_w_host: Optional[torch.Tensor] = None) -> None:
    """``lat_w_host``: contiguous float32 CPU copy of ``lat_w`` (raw level) — handed to the kernel as parameters."""
    B, S, _ = pyramid.shape
    if w_packed is None and impl in (2, 3):
        w_packed = fpn_pack_weights(w)
    hc, wc = (coarse.shape[1], coarse.shape[2]) if coarse is not None else (0, 0)
    raw_c = raw.shape[-1] if raw is not None else 0
    if raw is not None and raw.dtype == torch.uint8:
        raw_c |= RAW_U8
    if lat_w_host is not None and (lat_w_host.is_cuda or lat_w_host.dtype != torch.float32 or not lat_w_host.is_contiguous()):
        raise RuntimeError("fpn_output_forward: lat_w_host must be a contiguous float32 CPU tensor")
    st = _lib().dpft_fpn_output_forward_ex(native.ptr(inner), native.ptr(raw), raw_c, native.ptr(lat_w), native.ptr(lat_w_host),
                                           native.ptr(lat_b), native.ptr(coarse), hc, wc, native.ptr(w), native.ptr(w_packed),
                                           native.ptr(bias), native.ptr(pos_y), native.ptr(pos_x), native.ptr(pyramid),
                                           native.dtype_code(pyramid), S, start, B, H, W, impl, native.stream_ptr(pyramid.device))
    native.check(st, "dpft_fpn_output_forward_ex")
    native.count_launch()


RAW_U8 = 0x100            # include/dpft_b200.h: DPFT_RAW_U8


# Forked FPN output launches (NativeView._pyramid_forked): default since round 2 (bit-identical on B200, bench step
# 4.841 -> 4.775 ms, profiles/r02_ab_validated_paths.txt); DPFT_FPN_FORK=0 keeps everything on one stream for A/B timing
_FPN_FORK = os.environ.get("DPFT_FPN_FORK", "1") == "1"


class NativeView:
    """Prepared weights + launch sequence of one input view."""

    @staticmethod
    def ineligible_reason(backbone, neck, embedding, skiplink: bool) -> Optional[str]:
        if not isinstance(backbone, Backbone):
            return f"backbone {type(backbone).__name__}"
        if backbone.in_channels not in (3, 6):
            return f"in_channels={backbone.in_channels}"
        if not isinstance(backbone.body.bn1, nn.BatchNorm2d):
            return "norm layer is not BatchNorm2d"
        if not isinstance(neck, FPN) or neck.out_channels != FC:
            return "neck is not a 16-channel FPN"
        if not isinstance(embedding, MultiLevelSinusoidalEmbedding):
            return "embedding is not sinusoidal"
        if any(e.num_feats != FC for e in embedding.embedding_layers.values()):
            return "embedding num_feats != 16"
        n_levels = backbone.multi_scale + (1 if skiplink else 0)
        if len(neck.in_channels_list) != n_levels or embedding.n_levels < n_levels:
            return "neck / embedding level count does not match the backbone"
        return None

    def __init__(self, backbone: Backbone, neck: FPN, embedding: MultiLevelSinusoidalEmbedding, skiplink: bool, device,
                 dtype: torch.dtype = torch.bfloat16, pyramid_dtype: torch.dtype = torch.float32):
        self.device = device
        self.dtype = dtype                                                 # activation / weight type of the backbone
        self.pyramid_dtype = pyramid_dtype                                 # storage type of the (B, S, 16) pyramid
        self.skiplink = skiplink
        self.cin = backbone.in_channels
        body = backbone.body
        # stem: fold bn1, and the 1x1 adjustment conv (linear, bias-free, so it commutes with zero padding)
        w, b = fold_conv_bn(body.conv1, body.bn1)                          # (64, 7, 7, 3)
        if self.cin != 3:
            adj = backbone.adjustment_layer.weight.detach().float()[:, :, 0, 0].to(w.device)   # (3, Cin)
            w = torch.einsum("orsc,cd->orsd", w, adj)
        self.stem_w = w.permute(1, 2, 3, 0).contiguous().to(device)        # [7][7][Cin][64]
        self.stem_b = b.to(device)
        self.stem_w_packed = stem_pack_weights(self.stem_w)
        self.stages: List[List[Tuple[FoldedConv, FoldedConv, FoldedConv, Optional[FoldedConv]]]] = []
        for s in range(body.n_stages):
            blocks = []
            for blk in getattr(body, f"layer{s + 1}"):
                ds = (FoldedConv(blk.downsample[0], blk.downsample[1], device, dtype)
                      if blk.downsample is not None else None)
                blocks.append((FoldedConv(blk.conv1, blk.bn1, device, dtype), FoldedConv(blk.conv2, blk.bn2, device, dtype),
                               FoldedConv(blk.conv3, blk.bn3, device, dtype), ds))
            self.stages.append(blocks)
        # FPN
        fpn = neck.fpn
        self.n_levels = len(neck.in_channels_list)
        self.out_w, self.out_b, self.lat_w, self.lat_b, self.out_w_packed = [], [], [], [], []
        for i in range(self.n_levels):
            lat = fpn.inner_blocks[i][0]
            out = fpn.layer_blocks[i][0]
            self.out_w.append(out.weight.detach().float().permute(2, 3, 0, 1).contiguous().to(device))   # [3][3][o][c]
            self.out_b.append(out.bias.detach().float().contiguous().to(device))
            self.out_w_packed.append(fpn_pack_weights(self.out_w[-1]))
            if skiplink and i == 0:
                self.lat_w.append(lat.weight.detach().float()[:, :, 0, 0].contiguous().to(device))        # [16][Cin]
                self.lat_b.append(lat.bias.detach().float().contiguous().to(device))
                # host copy: travels as kernel parameters (DPFT_FPN_LAT_PARAMS=0 keeps the shared-memory path for A/B timing)
                self.lat_w_host = (lat.weight.detach().float()[:, :, 0, 0].contiguous().cpu().clone()
                                   if os.environ.get("DPFT_FPN_LAT_PARAMS", "1") == "1" else None)
            else:
                cin = lat.weight.shape[1]
                wpad = torch.zeros(64, cin, dtype=torch.float32)
                wpad[:FC] = lat.weight.detach().float()[:, :, 0, 0].cpu()
                bpad = torch.zeros(64, dtype=torch.float32)
                bpad[:FC] = lat.bias.detach().float().cpu()
                self.lat_w.append(wpad.to(device=device, dtype=dtype).contiguous())
                self.lat_b.append(bpad.to(device))
        self.embeddings = list(embedding.embedding_layers.values())
        self._pos: Dict[Tuple[int, int, int], Tuple[torch.Tensor, torch.Tensor]] = {}

    def _tables(self, level: int, H: int, W: int):
        key = (level, H, W)
        if key not in self._pos:
            py, px = self.embeddings[level].tables(H, W, torch.float32, self.device)
            self._pos[key] = (py.contiguous(), px.contiguous())
        return self._pos[key]

    def backbone(self, x: torch.Tensor) -> List[torch.Tensor]:
        """x (B,H,W,Cin) fp32 -> [layer1, ...] NHWC bf16."""
        y = maxpool_forward(stem_forward(x, self.stem_w, self.stem_b, self.dtype, w_packed=self.stem_w_packed))
        feats = []
        for blocks in self.stages:
            for c1, c2, c3, ds in blocks:
                identity = ds(y, relu=False) if ds is not None else y
                y = c3(c2(c1(y, relu=True), relu=True), relu=True, residual=identity)
            feats.append(y)
        return feats

    def _pyramid_forked(self, x, feats, pyr, shapes, starts, first):
        """Same kernels, same arguments, different order: the top-down chain of lateral GEMMs (each needs only the coarser
        INNER map) runs first; the small output-stage launches of the backbone levels (latency-bound, ~15 us each) then go to a
        forked stream and overlap the one large launch of the raw level instead of standing in front of it.  Results are
        bit-identical (no kernel sees different inputs)."""
        dev = x.device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_fork_stream", None) is None:
            self._fork_stream = torch.cuda.Stream(device=dev)
        inners, coarse = {}, None
        for li in range(self.n_levels - 1, first - 1, -1):
            coarse = inners[li] = lateral_forward(feats[li - first], self.lat_w[li], self.lat_b[li], coarse)
        fork = torch.cuda.Event()
        fork.record(main)
        self._fork_stream.wait_event(fork)
        with torch.cuda.stream(self._fork_stream):
            for li in range(self.n_levels - 1, first - 1, -1):
                H, W = shapes[li]
                py, px = self._tables(li, H, W)
                fpn_output_forward(pyr, starts[li], H, W, self.out_w[li], self.out_b[li], py, px, inner=inners[li],
                                   w_packed=self.out_w_packed[li])
            done = torch.cuda.Event()
            done.record(self._fork_stream)
        if self.skiplink:
            H, W = shapes[0]
            py, px = self._tables(0, H, W)
            fpn_output_forward(pyr, 0, H, W, self.out_w[0], self.out_b[0], py, px, raw=x, lat_w=self.lat_w[0],
                               lat_b=self.lat_b[0], coarse=inners[first], w_packed=self.out_w_packed[0],
                               lat_w_host=getattr(self, "lat_w_host", None))
        main.wait_event(done)
        del inners                      # alive until the join: their memory is not reused while the forked stream reads it
        return pyr, shapes

    def accepts_uint8(self, x: torch.Tensor) -> bool:
        """uint8 frames are read directly by the stem's row-streaming kernel and the FPN raw-level kernels when the view is a
        3-channel image at least 127 pixels wide (the camera); anything else is converted to float32 first (plumbing)."""
        return self.cin == 3 and (x.shape[2] - 1) // 2 + 1 >= 64 and self.stem_w_packed is not None

    def same_architecture(self, other: "NativeView") -> bool:
        """Every convolution of the two backbones has the same shape (different weights): they can share launches."""
        if self.cin != other.cin or self.dtype != other.dtype or len(self.stages) != len(other.stages):
            return False
        for ba, bb in zip(self.stages, other.stages):
            if len(ba) != len(bb):
                return False
            for ca, cb in zip(ba, bb):
                for fa, fb in zip(ca, cb):
                    if (fa is None) != (fb is None):
                        return False
                    if fa is not None and (fa.weight.shape != fb.weight.shape or fa.stride != fb.stride or fa.pad != fb.pad):
                        return False
        return True

    @staticmethod
    def backbone_pair(va: "NativeView", vb: "NativeView", xa: torch.Tensor, xb: torch.Tensor):
        """``backbone`` of two views with the same architecture on equally shaped inputs: stem and max-pool per view, every
        Bottleneck convolution of the two as ONE launch (``conv2d_nhwc_pair``).  Bit-identical to two ``backbone`` calls."""
        from .conv import conv2d_nhwc_pair
        ya = maxpool_forward(stem_forward(xa, va.stem_w, va.stem_b, va.dtype, w_packed=va.stem_w_packed))
        yb = maxpool_forward(stem_forward(xb, vb.stem_w, vb.stem_b, vb.dtype, w_packed=vb.stem_w_packed))
        fa, fb = [], []

        def both(ca, cb, ia, ib, relu, ra=None, rb=None):
            return conv2d_nhwc_pair((ia, ib), (ca.weight, cb.weight), (ca.bias, cb.bias), ca.stride, ca.pad, relu, (ra, rb))

        for blocks_a, blocks_b in zip(va.stages, vb.stages):
            for (a1, a2, a3, ad), (b1, b2, b3, bd) in zip(blocks_a, blocks_b):
                ida, idb = both(ad, bd, ya, yb, False) if ad is not None else (ya, yb)
                ta, tb = both(a1, b1, ya, yb, True)
                ta, tb = both(a2, b2, ta, tb, True)
                ya, yb = both(a3, b3, ta, tb, True, ida, idb)
            fa.append(ya)
            fb.append(yb)
        return fa, fb

    def pyramid(self, x: torch.Tensor, feats: Optional[List[torch.Tensor]] = None) -> Tuple[torch.Tensor, List[Tuple[int, int]]]:
        """``feats``: the stage outputs when the backbone already ran (``backbone_pair``)."""
        x = x.contiguous()
        if x.dtype == torch.uint8 and not self.accepts_uint8(x):
            x = x.float()
        B = x.shape[0]
        if feats is None:
            feats = self.backbone(x)
        shapes = ([(x.shape[1], x.shape[2])] if self.skiplink else []) + [(f.shape[1], f.shape[2]) for f in feats]
        sizes = [h * w for h, w in shapes]
        starts = [sum(sizes[:i]) for i in range(len(sizes))]
        S = sum(sizes)
        pyr = torch.empty((B, S, FC), dtype=self.pyramid_dtype, device=x.device)
        first = 1 if self.skiplink else 0
        coarse = None
        if _FPN_FORK and x.is_cuda and self.n_levels - first >= 2:
            return self._pyramid_forked(x, feats, pyr, shapes, starts, first)
        # top-down: coarsest level first
        for li in range(self.n_levels - 1, first - 1, -1):
            f = feats[li - first]
            inner = lateral_forward(f, self.lat_w[li], self.lat_b[li], coarse)
            H, W = shapes[li]
            py, px = self._tables(li, H, W)
            fpn_output_forward(pyr, starts[li], H, W, self.out_w[li], self.out_b[li], py, px, inner=inner,
                               w_packed=self.out_w_packed[li])
            coarse = inner
        if self.skiplink:
            H, W = shapes[0]
            py, px = self._tables(0, H, W)
            fpn_output_forward(pyr, 0, H, W, self.out_w[0], self.out_b[0], py, px, raw=x, lat_w=self.lat_w[0],
                               lat_b=self.lat_b[0], coarse=coarse, w_packed=self.out_w_packed[0],
                               lat_w_host=getattr(self, "lat_w_host", None))
        return pyr, shapes
