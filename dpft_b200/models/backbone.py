"""ResNet backbones of the DPRT model (mirror of reference src/dprt/models/backbones/resnet.py).

The parameter tree reproduces torchvision's ResNet-50/101/152 (v1.5 bottlenecks) under ``body.*`` plus the
optional 1x1 ``adjustment_layer`` (C -> 3, no bias, resnet.py:47-51), so a reference ``state_dict`` loads
with ``strict=True``.  Inputs and outputs are channel-last (B, H, W, C) like the reference (resnet.py:94-105).

Two execution paths, both on the GPU:
  * ``forward``          — layer by layer through torch (cuDNN); autograd-capable; used for training.
  * ``forward_folded``   — inference: BatchNorm folded into the convolutions, NHWC, run by the
                           sm_100a convolution kernels in libdpft_b200.so (dpft_b200/conv.py).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Dict, Optional

import torch
import torch.nn.functional as F
from torch import nn

STAGE_BLOCKS = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3), "resnet152": (3, 8, 36, 3)}
STAGE_PLANES = (64, 128, 256, 512)
EXPANSION = 4


def _norm(name: Optional[str]):
    if name is None or name == "BatchNorm2d":
        return nn.BatchNorm2d
    if name == "FrozenBatchNorm2d":
        import torchvision
        return torchvision.ops.FrozenBatchNorm2d
    return getattr(nn, name)


class Bottleneck(nn.Module):
    """1x1 -> 3x3(stride) -> 1x1 with identity / projected shortcut (torchvision Bottleneck, v1.5)."""

    def __init__(self, inplanes: int, planes: int, stride: int, norm, project: bool):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = norm(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = norm(planes)
        self.conv3 = nn.Conv2d(planes, planes * EXPANSION, 1, bias=False)
        self.bn3 = norm(planes * EXPANSION)
        self.stride = stride
        if project:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes * EXPANSION, 1, stride=stride, bias=False),
                                            norm(planes * EXPANSION))
        else:
            self.downsample = None

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = F.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        if self.downsample is not None:
            x = self.downsample(x)
        return F.relu(out + x)


class ResNetBody(nn.Module):
    """Stem + the first ``multi_scale`` stages; returns {'1': layer1, ...} (NCHW) like IntermediateLayerGetter."""

    def __init__(self, arch: str, multi_scale: int, norm):
        super().__init__()
        blocks = STAGE_BLOCKS[arch]
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = norm(64)
        inplanes = 64
        self.n_stages = multi_scale
        for s in range(multi_scale):
            planes, stride = STAGE_PLANES[s], (1 if s == 0 else 2)
            layer = []
            for i in range(blocks[s]):
                layer.append(Bottleneck(inplanes, planes, stride if i == 0 else 1, norm, project=(i == 0)))
                inplanes = planes * EXPANSION
            setattr(self, f"layer{s + 1}", nn.Sequential(*layer))
        for m in self.modules():                          # torchvision's ResNet initialisation
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(x)))
        x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
        out = OrderedDict()
        for s in range(self.n_stages):
            x = getattr(self, f"layer{s + 1}")(x)
            out[str(s + 1)] = x
        return out


class Backbone(nn.Module):
    def __init__(self, name: str, weights: str = "", norm_layer: Optional[str] = None, in_channels: int = 3,
                 multi_scale: int = 1, channel_last: bool = True, **kwargs):
        super().__init__()
        arch = name.lower()
        if arch not in STAGE_BLOCKS:
            raise NotImplementedError(f"backbone {name!r}: only ResNet50/101/152 are on the accelerated path")
        if not 1 <= multi_scale <= 4:
            raise ValueError(f"multi_scale must be in [1, 4], got {multi_scale}")
        self.arch, self.in_channels, self.multi_scale, self.channel_last = arch, in_channels, multi_scale, channel_last
        if in_channels == 3:
            self.adjustment_layer = nn.Identity()
        else:
            self.adjustment_layer = nn.Conv2d(in_channels, 3, 1, bias=False)
        self.body = ResNetBody(arch, multi_scale, _norm(norm_layer))
        self.native_train = False             # train(): run layer1.. through the sm_100a training kernels (16-bit activations)
        self.train_dtype = torch.float16      # float16 (gradients carried with a power-of-two scale) or bfloat16
        self.native_stem = True               # ... and the stem (conv 7x7 + bn1 + relu + max-pool) as well
        self._stages = None
        self._stem = None
        if weights:
            self._load_weights(name, weights)

    def _load_weights(self, name: str, weights: str) -> None:
        """``weights`` is a torchvision weight-enum member (downloaded by torchvision, as the reference does at
        resnet.py:157-167) or a path to a state dict of this module."""
        import os
        if os.path.exists(weights):
            self.load_state_dict(torch.load(weights, map_location="cpu"))
            return
        import torchvision
        enum = torchvision.models.get_weight(f"{name}_Weights.{weights}")
        sd = enum.get_state_dict(progress=False)
        own = self.body.state_dict()
        self.body.load_state_dict({k: v for k, v in sd.items() if k in own})

    @classmethod
    def from_config(cls, config: Dict[str, Any]):
        return cls(**config)

    def __getstate__(self):            # torch.save(model) must not pickle the device-side launch plan
        state = self.__dict__.copy()
        state["_stages"] = None
        state["_stem"] = None
        return state

    def _native_stages(self):
        """The native training plan of layer1.. (dpft_b200/train_backbone.py), or None when it does not apply."""
        from ..train_backbone import NativeStages
        st = self._stages
        if st is not None and st.device == self.body.conv1.weight.device and st.dtype == self.train_dtype:
            return st
        if NativeStages.ineligible_reason(self.body) is not None:
            return None
        self._stages = NativeStages(self.body, self.train_dtype)
        return self._stages

    def forward_native_train(self, batch: torch.Tensor, stages) -> "OrderedDict[str, torch.Tensor]":
        """The whole backbone as one native autograd node (stem through torch autograd when the stem kernels do not cover
        it, or when the raw input itself needs a gradient); same outputs as ``forward`` (fp32)."""
        from ..train_backbone import NativeStem, backbone_forward, stages_forward
        if self.native_stem and not batch.requires_grad and NativeStem.ineligible_reason(self) is None:
            x = batch if self.channel_last else batch.movedim(1, -1)
            if self._stem is None or self._stem.dtype != self.train_dtype or self._stem.zero_bias.device != x.device:
                self._stem = NativeStem(self, self.train_dtype)
            outs = backbone_forward(stages, self._stem, x)
        else:
            x = batch.movedim(-1, 1) if self.channel_last else batch
            body = self.body
            x = F.relu(body.bn1(body.conv1(self.adjustment_layer(x))))
            x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
            outs = stages_forward(stages, x.permute(0, 2, 3, 1))
        feats = OrderedDict((str(i + 1), o) for i, o in enumerate(outs))
        if not self.channel_last:
            feats = OrderedDict((k, v.movedim(-1, 1)) for k, v in feats.items())
        return feats

    def forward(self, batch: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
        if self.native_train and self.training and batch.is_cuda and torch.is_grad_enabled():
            stages = self._native_stages()
            if stages is not None:
                return self.forward_native_train(batch, stages)
        x = batch.movedim(-1, 1) if self.channel_last else batch
        feats = self.body(self.adjustment_layer(x))
        if self.channel_last:
            feats = OrderedDict((k, v.movedim(1, -1)) for k, v in feats.items())
        return feats


def build_backbone(name: str, config: Dict[str, Any], *args, **kwargs):
    if "resnet" in name.lower():
        return Backbone.from_config(config)
    raise NotImplementedError(f"backbone {name!r} is outside the accelerated hot path (SURVEY.md §2 row 2b)")
