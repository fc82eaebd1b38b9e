"""Feature pyramid neck (mirror of reference src/dprt/models/necks/fpn.py over torchvision's
FeaturePyramidNetwork): per level a 1x1 lateral conv and a 3x3 output conv (both with bias, no norm, no
activation), top-down nearest-neighbour upsample-add.  Parameter names: ``fpn.inner_blocks.{i}.0.*``,
``fpn.layer_blocks.{i}.0.*``.  Channel-last in, channel-last out, finest level first.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Dict, List

import torch
import torch.nn.functional as F
from torch import nn


class _Pyramid(nn.Module):
    def __init__(self, in_channels_list: List[int], out_channels: int):
        super().__init__()
        self.inner_blocks = nn.ModuleList(nn.Sequential(nn.Conv2d(c, out_channels, 1)) for c in in_channels_list)
        self.layer_blocks = nn.ModuleList(nn.Sequential(nn.Conv2d(out_channels, out_channels, 3, padding=1))
                                          for _ in in_channels_list)
        for m in self.modules():                          # torchvision FPN initialisation
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, a=1)
                nn.init.constant_(m.bias, 0)


class FPN(nn.Module):
    def __init__(self, in_channels_list: List[int], out_channels: int, norm_layer=None, channel_last: bool = True,
                 **kwargs):
        super().__init__()
        if norm_layer is not None:
            raise NotImplementedError("FPN norm layers are not used by any shipped DPFT config")
        self.in_channels_list, self.out_channels, self.channel_last = list(in_channels_list), out_channels, channel_last
        self.fpn = _Pyramid(self.in_channels_list, out_channels)

    @classmethod
    def from_config(cls, config: Dict[str, Any]):
        return cls(config["in_channels_list"], config["out_channels"], config.get("norm_layer"))

    def forward(self, batch: "OrderedDict[str, torch.Tensor]") -> "OrderedDict[str, torch.Tensor]":
        names = list(batch.keys())
        xs = [v.movedim(-1, 1) if self.channel_last else v for v in batch.values()]
        inner, layer = self.fpn.inner_blocks, self.fpn.layer_blocks
        last = inner[-1](xs[-1])
        outs = [layer[-1](last)]
        for i in range(len(xs) - 2, -1, -1):
            lateral = inner[i](xs[i])
            last = lateral + F.interpolate(last, size=lateral.shape[-2:], mode="nearest")
            outs.insert(0, layer[i](last))
        if self.channel_last:
            outs = [o.movedim(1, -1) for o in outs]
        return OrderedDict(zip(names, outs))


def build_neck(name: str, config: Dict[str, Any], *args, **kwargs):
    if "fpn" in name.lower():
        return FPN.from_config(config)
    return None
