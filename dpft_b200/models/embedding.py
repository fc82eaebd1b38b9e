"""Sinusoidal positional embedding of the feature maps (mirror of reference
src/dprt/models/embeddings/sinusoidal.py).  The embedding depends only on (H, W, num_feats, normalize), so
the per-axis tables are computed once per level shape and cached; the reference recomputes them each
forward and makes two read-modify-write passes over every level (sinusoidal.py:107-108).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Any, Dict, Tuple

import torch
from torch import nn


def axis_tables(H: int, W: int, num_feats: int, normalize: bool, temperature: float = 10000.0,
                scale: float = 2 * math.pi, eps: float = 1e-6, offset: float = 0.0,
                dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (pos_y (H, num_feats), pos_x (W, num_feats)) with sin on even and cos on odd channels,
    computed with the reference's operation order (sinusoidal.py:83-104)."""
    y = torch.arange(1, H + 1, dtype=dtype)
    x = torch.arange(1, W + 1, dtype=dtype)
    if normalize:
        y = (y + offset) / (y[-1:] + eps) * scale
        x = (x + offset) / (x[-1:] + eps) * scale
    dim_t = torch.arange(num_feats, dtype=dtype)
    dim_t = temperature ** (2 * (dim_t // 2) / num_feats)

    def enc(v):
        p = v[:, None] / dim_t
        return torch.stack((p[:, 0::2].sin(), p[:, 1::2].cos()), dim=2).reshape(v.shape[0], -1)

    return enc(y), enc(x)


class SinusoidalEmbedding(nn.Module):
    def __init__(self, num_feats: int, temperature: int = 10000, normalize: bool = False,
                 scale: float = 2 * math.pi, eps: float = 1e-6, offset: float = 0.0, **kwargs):
        super().__init__()
        self.num_feats, self.temperature, self.normalize = num_feats, temperature, normalize
        self.scale, self.eps, self.offset = scale, eps, offset
        self._cache: Dict[Any, Tuple[torch.Tensor, torch.Tensor]] = {}

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_cache"] = {}
        return state

    def tables(self, H: int, W: int, dtype, device) -> Tuple[torch.Tensor, torch.Tensor]:
        key = (H, W, dtype, str(device))
        if key not in self._cache:
            py, px = axis_tables(H, W, self.num_feats, self.normalize, self.temperature, self.scale, self.eps,
                                 self.offset, dtype)
            self._cache[key] = (py.to(device), px.to(device))
        return self._cache[key]

    def forward(self, batch: torch.Tensor) -> torch.Tensor:
        _, H, W, _ = batch.shape
        py, px = self.tables(H, W, batch.dtype, batch.device)
        return batch + px[None, None, :, :] + py[None, :, None, :]


class MultiLevelSinusoidalEmbedding(nn.Module):
    def __init__(self, n_levels: int = 1, **kwargs):
        super().__init__()
        self.n_levels = n_levels
        self.embedding_layers = nn.ModuleDict({f"embedding{i}": SinusoidalEmbedding(**kwargs) for i in range(n_levels)})

    @classmethod
    def from_config(cls, config: Dict[str, Any]):
        return cls(**config)

    def forward(self, batches: "OrderedDict[str, torch.Tensor]") -> "OrderedDict[str, torch.Tensor]":
        return OrderedDict((k, layer(b)) for (k, b), layer in zip(batches.items(), self.embedding_layers.values()))


def build_embedding(name: str, config: Dict[str, Any], *args, **kwargs):
    if "sinusoidal" in name:
        return MultiLevelSinusoidalEmbedding.from_config(config)
    return None
