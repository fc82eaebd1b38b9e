"""Detection head (mirror of reference src/dprt/models/heads/detection.py, LinearDetectionHead :149-275).

Parameter names match the reference (``layers.{center,size,angle,class}_head.{0,3,6}.weight``): each branch
is Linear -> ReLU -> Dropout repeated, ending in a Linear; all bias-free by default.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Dict

import torch
from torch import nn

ACTIVATIONS = OrderedDict(center="Identity", size="ReLU", angle="Tanh")
ACTIVATIONS["class"] = "Identity"
OUT_DIMS = {"center": 3, "size": 3, "angle": 2}


class LinearDetectionHead(nn.Module):
    def __init__(self, in_channels: int, num_classes: int, num_reg_layers: int = 1, num_cls_layers: int = 1,
                 bias: bool = False, dropout: float = 0.0, channels_last: bool = True, **kwargs):
        super().__init__()
        self.in_channels, self.num_classes = in_channels, num_classes
        self.num_reg_layers, self.num_cls_layers = num_reg_layers, num_cls_layers
        self.bias, self.dropout, self.channels_last = bias, dropout, channels_last
        self.activations = dict(ACTIVATIONS)
        self.layers = nn.ModuleDict({
            f"{k}_head": self._branch(OUT_DIMS.get(k, num_classes), num_cls_layers if k == "class" else num_reg_layers)
            for k in ACTIVATIONS
        })
        self.activation_fn = nn.ModuleDict({k: getattr(nn, v)() for k, v in ACTIVATIONS.items()})

    def _branch(self, out_channels: int, depth: int) -> nn.Sequential:
        mods = []
        for _ in range(depth - 1):
            mods += [nn.Linear(self.in_channels, self.in_channels, bias=self.bias), nn.ReLU(), nn.Dropout(self.dropout)]
        mods.append(nn.Linear(self.in_channels, out_channels, bias=self.bias))
        return nn.Sequential(*mods)

    @classmethod
    def from_config(cls, config: Dict[str, Any]):
        return cls(config["in_channels"], config["num_classes"], config.get("num_reg_layers", 1),
                   config.get("num_cls_layers", 1), config.get("bias", False), config.get("dropout", 0.0),
                   config.get("channels_last", True))

    def forward(self, batch: torch.Tensor, ref: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
        out = OrderedDict((k, self.activation_fn[k](self.layers[f"{k}_head"](batch))) for k in ACTIVATIONS)
        center = out["center"]
        out["center"] = torch.cat((center[..., :3] + ref["center"][..., :3], center[..., 3:]), -1)  # detection.py:273
        return out


def build_head(name: str, config: Dict[str, Any], *args, **kwargs):
    low = name.lower()
    if "detection" in low and "linear" in low:
        return LinearDetectionHead.from_config(config)
    if "detection" in low and "unary" in low:
        raise NotImplementedError("UnaryDetectionHead is outside the accelerated hot path (SURVEY.md §2 row 7b)")
    return None
