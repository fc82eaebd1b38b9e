"""Native (sm_100a) training path of the ResNet stages: forward with batch-statistics BatchNorm and the whole backward,
as ONE autograd node per backbone.

Replaces, under ``model.train()``, the cuDNN convolution / batch-norm / ReLU forward+backward launches autograd records for
the torchvision Bottleneck blocks the reference trains (src/dprt/models/backbones/resnet.py:54-55,101; the backward is
driven from src/dprt/training/trainer.py:125-133).  Activations and activation gradients are NHWC 16-bit (float16 with a
power-of-two gradient scale, or bfloat16), accumulation and every parameter gradient fp32.

Per block (x = block input, 16-bit):
    y1 = conv1(x)            z1 = relu(bn1(y1))
    y2 = conv2(z1)           z2 = relu(bn2(y2))
    y3 = conv3(z2)           z3 = relu(bn3(y3) + idn)         idn = x  or  bn_d(conv_d(x))
Kernels: ``dpft_conv2d_nhwc`` (tcgen05 implicit GEMM; also the data gradients, on flipped/transposed weights),
``dpft_conv2d_wgrad`` (tcgen05 GEMM over the pixels), ``dpft_bn_*`` (streaming passes), ``dpft_zero_insert2_nhwc``,
``dpft_pack_conv_weights`` (one launch per step re-lays out every weight after the optimiser update).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import nn

from . import train_ops as T
from .conv import conv2d_nhwc


class _ConvSpec:
    __slots__ = ("conv", "bn", "idx", "stride", "pad", "r", "cin", "cout", "bn_off")

    def __init__(self, conv: nn.Conv2d, bn: nn.BatchNorm2d, idx: int, bn_off: int):
        self.conv, self.bn, self.idx, self.bn_off = conv, bn, idx, bn_off
        self.stride, self.pad, self.r = conv.stride[0], conv.padding[0], conv.kernel_size[0]
        self.cout, self.cin = conv.weight.shape[0], conv.weight.shape[1]


class NativeStages:
    """Prepared launch sequence of ``body.layer1 .. layer<n>`` for training."""

    @staticmethod
    def ineligible_reason(body) -> Optional[str]:
        for m in body.modules():
            if isinstance(m, nn.Conv2d) and m is not body.conv1:
                if m.bias is not None or m.groups != 1 or m.dilation[0] != 1 or m.weight.shape[0] % 64 or m.weight.shape[1] % 64:
                    return "convolution outside the bottleneck family"
        if not isinstance(body.bn1, nn.BatchNorm2d):
            return "norm layer is not BatchNorm2d"
        for m in body.modules():
            if isinstance(m, nn.BatchNorm2d) and (m.momentum is None or not m.affine or not m.track_running_stats):
                return "BatchNorm2d variant without momentum / affine / running statistics"
        return None

    def __init__(self, body, dtype: torch.dtype = torch.bfloat16):
        self.dtype = dtype
        self.blocks: List[Tuple[_ConvSpec, _ConvSpec, _ConvSpec, Optional[_ConvSpec]]] = []
        self.stage_ends: List[int] = []
        self.specs: List[_ConvSpec] = []
        bn_off = 0

        def spec(conv, bn):
            nonlocal bn_off
            s = _ConvSpec(conv, bn, len(self.specs), bn_off)
            bn_off += s.cout
            self.specs.append(s)
            return s

        for st in range(body.n_stages):
            for blk in getattr(body, f"layer{st + 1}"):
                ds = spec(blk.downsample[0], blk.downsample[1]) if blk.downsample is not None else None
                self.blocks.append((spec(blk.conv1, blk.bn1), spec(blk.conv2, blk.bn2), spec(blk.conv3, blk.bn3), ds))
            self.stage_ends.append(len(self.blocks) - 1)
        self.bn_channels = bn_off
        self.device = self.specs[0].conv.weight.device
        self.packer = T.WeightPacker([s.conv.weight for s in self.specs], dtype, [True] * len(self.specs))
        self.wgrad_offsets, o = [], 0
        for s in self.specs:
            self.wgrad_offsets.append(o)
            o += s.conv.weight.numel()
        self.wgrad_total = o
        self.zero_bias = torch.zeros(4096, dtype=torch.float32, device=self.device)
        self._packed_version = None

    def _wgrad_stream(self):
        import os
        if os.environ.get("DPFT_WGRAD_STREAM", "1") != "1":
            return None
        from .streams import single_process
        if not single_process():                     # see streams.single_process
            return None
        if getattr(self, "_ws", None) is None:
            from .streams import new_stream
            self._ws = new_stream(self.device)
        return self._ws

    # parameters in the order the autograd node receives them / returns their gradients
    def parameters(self) -> List[torch.Tensor]:
        out = []
        for s in self.specs:
            out += [s.conv.weight, s.bn.weight, s.bn.bias]
        return out

    def _refresh_weights(self) -> None:
        version = tuple(s.conv.weight._version for s in self.specs)
        # under CUDA-graph capture the host cannot know whether the replayed step follows an optimiser update: always re-pack
        if (version != self._packed_version or self.packer._ptrs != [w.data_ptr() for w in self.packer.weights]
                or torch.cuda.is_current_stream_capturing()):
            self.packer.refresh()
            self._packed_version = version

    # ---- forward ---------------------------------------------------------------------------------------------------
    def _conv(self, s: _ConvSpec, x: torch.Tensor) -> torch.Tensor:
        return conv2d_nhwc(x, self.packer.fwd[s.idx], self.zero_bias, s.stride, s.pad, False)

    def _bn(self, s: _ConvSpec, y: torch.Tensor, states: torch.Tensor, relu: bool, residual=None):
        bn = s.bn
        buf = states[4 * s.bn_off: 4 * (s.bn_off + s.cout)].view(4, s.cout)
        return T.bn_forward(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps, relu, residual, buf)

    def forward(self, x: torch.Tensor):
        """x (B,H,W,64) 16-bit (the max-pooled stem output) -> (stage outputs, tape)."""
        self._refresh_weights()
        states = torch.empty(4 * self.bn_channels, dtype=torch.float32, device=x.device)
        tape = []
        outs = []
        for bi, (c1, c2, c3, ds) in enumerate(self.blocks):
            y1 = self._conv(c1, x)
            z1, s1 = self._bn(c1, y1, states, True)
            y2 = self._conv(c2, z1)
            z2, s2 = self._bn(c2, y2, states, True)
            y3 = self._conv(c3, z2)
            if ds is not None:
                yd = self._conv(ds, x)
                idn, sd = self._bn(ds, yd, states, False)
            else:
                yd, sd, idn = None, None, x
            z3, s3 = self._bn(c3, y3, states, True, residual=idn)
            tape.append((x, y1, z1, s1, y2, z2, s2, y3, z3, s3, yd, sd))
            x = z3
            if bi in self.stage_ends:
                outs.append(z3)
        nbt = [s.bn.num_batches_tracked for s in self.specs if s.bn.num_batches_tracked is not None]
        if nbt:
            torch._foreach_add_(nbt, 1)
        return outs, tape

    # ---- backward --------------------------------------------------------------------------------------------------
    def backward(self, tape, grad_outs: List[Optional[torch.Tensor]]):
        dev = self.device
        wg = torch.zeros(self.wgrad_total, dtype=torch.float32, device=dev)
        bng = torch.zeros(2 * self.bn_channels, dtype=torch.float32, device=dev)       # dgamma | dbeta
        sums = torch.empty(2 * self.bn_channels, dtype=torch.float32, device=dev)

        def bn_bwd(s: _ConvSpec, dz, z, y, state, relu, want_g=False):
            sm = sums[2 * s.bn_off: 2 * (s.bn_off + s.cout)].view(2, s.cout)
            return T.bn_backward(dz, z, y, state, s.bn.weight, relu, bng[s.bn_off: s.bn_off + s.cout],
                                 bng[self.bn_channels + s.bn_off: self.bn_channels + s.bn_off + s.cout], want_g, sm)

        # Weight gradients are leaves of the backward sweep (only the optimiser reads them): they go to a stream of their own and
        # fill the tails of the data-gradient / BatchNorm chain, which is the dependency chain (DPFT_WGRAD_STREAM=0: same stream).
        cur = torch.cuda.current_stream(dev)
        ws = self._wgrad_stream() if dev.type == "cuda" else None

        def wgrad(s: _ConvSpec, x, dy):
            o = self.wgrad_offsets[s.idx]
            if ws is None:
                T.conv2d_wgrad(x, dy, s.r, s.r, s.stride, s.pad, out=wg[o: o + s.conv.weight.numel()])
                return
            ready = torch.cuda.Event()
            ready.record(cur)
            ws.wait_event(ready)
            with torch.cuda.stream(ws):
                T.conv2d_wgrad(x, dy, s.r, s.r, s.stride, s.pad, out=wg[o: o + s.conv.weight.numel()])
            x.record_stream(ws)
            dy.record_stream(ws)

        def dgrad(s: _ConvSpec, dy, x, residual=None):
            return T.conv2d_dgrad(dy, self.packer.dgrad[s.idx], self.zero_bias, (x.shape[1], x.shape[2]), s.stride, s.pad, residual)

        dz = None
        stage_of = {b: i for i, b in enumerate(self.stage_ends)}
        for bi in range(len(self.blocks) - 1, -1, -1):
            c1, c2, c3, ds = self.blocks[bi]
            x, y1, z1, s1, y2, z2, s2, y3, z3, s3, yd, sd = tape[bi]
            if bi in stage_of and grad_outs[stage_of[bi]] is not None:
                g_out = grad_outs[stage_of[bi]]
                dz = g_out if dz is None else dz.add_(g_out)
            if dz is None:                                   # nothing downstream of this block needs a gradient
                dz = torch.zeros_like(z3)
            dy3, g = bn_bwd(c3, dz, z3, y3, s3, True, want_g=True)
            wgrad(c3, z2, dy3)
            dz2 = dgrad(c3, dy3, z2)
            dy2, _ = bn_bwd(c2, dz2, z2, y2, s2, True)
            wgrad(c2, z1, dy2)
            dz1 = dgrad(c2, dy2, z1)
            dy1, _ = bn_bwd(c1, dz1, z1, y1, s1, True)
            wgrad(c1, x, dy1)
            if ds is not None:
                dyd, _ = bn_bwd(ds, g, None, yd, sd, False)
                wgrad(ds, x, dyd)
                res = dgrad(ds, dyd, x)
            else:
                res = g
            dz = dgrad(c1, dy1, x, residual=res)
            tape[bi] = None                                  # release the block's activations as the sweep passes
        if ws is not None:
            cur.wait_stream(ws)                              # join: the gradients below are read on the sweep's stream
        grads = []
        for s in self.specs:
            o = self.wgrad_offsets[s.idx]
            co, ci, r = s.cout, s.cin, s.r
            g = wg[o: o + co * ci * r * r]
            # 1x1 filters: (co, 1, 1, ci) and torch's (co, ci, 1, 1) are the same memory — hand autograd a CONTIGUOUS gradient so
            # that its accumulation into the flat bucket is a vectorised add, not a strided one (2/3 of the 210 conv weights)
            grads.append(g.view(co, ci, 1, 1) if r == 1 else g.view(co, r, r, ci).permute(0, 3, 1, 2))
            grads.append(bng[s.bn_off: s.bn_off + co])
            grads.append(bng[self.bn_channels + s.bn_off: self.bn_channels + s.bn_off + co])
        return dz, grads, (wg, bng)


GRAD_TARGET_AMAX = 64.0      # f16 gradients: the incoming gradients are scaled so that their largest element is ~2^6


class NativeStem:
    """[adjustment 1x1 (radar, C -> 3) folded into] conv1 7x7/2 -> bn1 (batch statistics) -> ReLU -> max-pool, forward and
    backward (reference src/dprt/models/backbones/resnet.py:98-101 in train()).  The input needs no gradient; the adjustment
    layer is linear and bias-free, so the kernels see the composed 7x7 weight W'[o,c] = sum_k W1[o,k] A[k,c] and the chain
    rule back to W1 and A is two small einsums on the 7x7 gradient."""

    @staticmethod
    def ineligible_reason(backbone) -> Optional[str]:
        if backbone.in_channels not in (3, 6):
            return f"in_channels={backbone.in_channels}"
        adj = backbone.adjustment_layer
        if backbone.in_channels != 3 and not (isinstance(adj, nn.Conv2d) and adj.kernel_size == (1, 1) and adj.bias is None):
            return "adjustment layer is not a bias-free 1x1 convolution"
        c1 = backbone.body.conv1
        if c1.kernel_size != (7, 7) or c1.stride != (2, 2) or c1.padding != (3, 3) or c1.bias is not None or c1.weight.shape[0] != 64:
            return "stem is not conv 7x7 / stride 2 / pad 3 -> 64"
        return None

    def __init__(self, backbone, dtype: torch.dtype):
        self.backbone, self.dtype = backbone, dtype
        self.cin = backbone.in_channels
        self.zero_bias = torch.zeros(64, dtype=torch.float32, device=backbone.body.conv1.weight.device)

    def parameters(self) -> List[torch.Tensor]:
        body = self.backbone.body
        ps = [body.conv1.weight, body.bn1.weight, body.bn1.bias]
        if self.cin != 3:
            ps.append(self.backbone.adjustment_layer.weight)
        return ps

    def _composed_weight(self) -> torch.Tensor:
        w1 = self.backbone.body.conv1.weight.detach()                       # (64, 3, 7, 7)
        if self.cin == 3:
            return w1.permute(2, 3, 1, 0).contiguous()                      # [7][7][3][64]
        a = self.backbone.adjustment_layer.weight.detach()[:, :, 0, 0]      # (3, Cin)
        return torch.einsum("okrs,kc->rsco", w1, a).contiguous()

    def forward(self, x: torch.Tensor):
        from . import features as Fe
        bn = self.backbone.body.bn1
        w = self._composed_weight()
        y0 = Fe.stem_forward(x, w, self.zero_bias, self.dtype, w_packed=Fe.stem_pack_weights(w), relu=False)
        z0, s0 = T.bn_forward(y0, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps, True)
        if bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        pooled = Fe.maxpool_forward(z0)
        return pooled, (x, y0, z0, s0, pooled)

    def backward(self, tape, dpooled: torch.Tensor):
        """-> (gradients in ``parameters()`` order, flat buffers that carry the gradient scale)."""
        x, y0, z0, s0, pooled = tape
        bn = self.backbone.body.bn1
        dz0 = T.maxpool_backward(z0, pooled, dpooled)
        bng = torch.zeros(2, 64, dtype=torch.float32, device=x.device)
        dy0, _ = T.bn_backward(dz0, z0, y0, s0, bn.weight, True, bng[0], bng[1])
        dw = T.stem_wgrad(x, dy0)                                           # [7][7][Cin][64]
        return dw, bng

    def chain_to_parameters(self, dw: torch.Tensor, bng: torch.Tensor) -> List[torch.Tensor]:
        if self.cin == 3:
            return [dw.permute(3, 2, 0, 1), bng[0], bng[1]]
        w1 = self.backbone.body.conv1.weight.detach()
        a = self.backbone.adjustment_layer.weight.detach()[:, :, 0, 0]
        dw1 = torch.einsum("rsco,kc->okrs", dw, a)
        da = torch.einsum("rsco,okrs->kc", dw, w1)
        return [dw1, bng[0], bng[1], da[:, :, None, None]]


class _BackboneFn(torch.autograd.Function):
    """fp32 in / fp32 out at the autograd boundary; 16-bit inside.  With f16 the activation gradients are carried with a
    power-of-two scale S (computed on the device from the incoming gradients, no host sync) so that they stay inside
    f16's exponent range; every backward kernel is linear in the gradient, so the parameter gradients and the input
    gradient are multiplied by 1/S at the end.

    ``stem`` given: x is the raw input (B,H,W,Cin) fp32 (no gradient) and the node covers the whole backbone;
    otherwise x is the max-pooled stem output (B,h,w,64) fp32 computed (and differentiated) by torch."""

    @staticmethod
    def forward(ctx, runner: NativeStages, stem: Optional[NativeStem], x: torch.Tensor, *params):
        if stem is not None:
            pooled, stem_tape = stem.forward(x.contiguous())
        else:
            pooled, stem_tape = x.to(runner.dtype).contiguous(), None
        outs, tape = runner.forward(pooled)
        ctx.runner, ctx.stem, ctx.tape, ctx.stem_tape = runner, stem, tape, stem_tape
        return tuple(o.float() for o in outs)

    @staticmethod
    def backward(ctx, *grad_outs):
        runner, stem = ctx.runner, ctx.stem
        scale = None
        if runner.dtype == torch.float16:
            amax = torch.stack([g.abs().max() for g in grad_outs if g is not None]).max()
            scale = torch.exp2(torch.floor(torch.log2(GRAD_TARGET_AMAX / amax.clamp_min(1e-30)))).clamp(2.0 ** -40, 2.0 ** 40)
            gos = [None if g is None else (g * scale).to(runner.dtype).contiguous() for g in grad_outs]
        else:
            gos = [None if g is None else g.to(runner.dtype).contiguous() for g in grad_outs]
        dpooled, grads, flats = runner.backward(ctx.tape, gos)
        ctx.tape = None
        flats = list(flats)
        stem_flats = None
        if stem is not None:
            stem_flats = stem.backward(ctx.stem_tape, dpooled)
            ctx.stem_tape = None
            flats += list(stem_flats)
            dx = None
        else:
            dx = dpooled.float()
        if scale is not None:
            inv = 1.0 / scale
            if dx is not None:
                dx = dx * inv
            for f in flats:
                f.mul_(inv)
        stem_grads = stem.chain_to_parameters(*stem_flats) if stem is not None else []
        return (None, None, dx, *stem_grads, *grads)


def stages_forward(runner: NativeStages, pooled: torch.Tensor) -> Tuple[torch.Tensor, ...]:
    """pooled (B,H,W,64) fp32 NHWC (autograd-tracked) -> tuple of stage outputs (B,h,w,C) fp32 NHWC."""
    return _BackboneFn.apply(runner, None, pooled, *runner.parameters())


def backbone_forward(runner: NativeStages, stem: NativeStem, x: torch.Tensor) -> Tuple[torch.Tensor, ...]:
    """x (B,H,W,Cin) fp32 raw input -> tuple of stage outputs (B,h,w,C) fp32 NHWC; the whole backbone is one autograd node."""
    return _BackboneFn.apply(runner, stem, x, *stem.parameters(), *runner.parameters())
