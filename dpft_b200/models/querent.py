"""Static query reference points (mirror of reference src/dprt/models/queries/data_agnostic.py).

The grid depends only on the config, so it is computed once per (dtype, device) and cached; the reference
rebuilds it every forward and syncs the host on ``torch.isclose`` (data_agnostic.py:121).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Dict, List, Optional, Union

import torch
from torch import nn

from .geometry import build_transformation


class DataAgnosticStaticQueries(nn.Module):
    def __init__(self, resolution: List[int] = None, minimum: List[float] = None, maximum: List[float] = None,
                 transformation: Optional[nn.Module] = None, distribution: Union[str, List[str], None] = None,
                 **kwargs):
        super().__init__()
        self.resolution = list(resolution or [])
        self.minimum = list(minimum or [])
        self.maximum = list(maximum or [])
        self.transformation = transformation if transformation is not None else nn.Identity()
        if distribution is None:
            distribution = ["linear"] * len(self.resolution)
        elif isinstance(distribution, str):
            distribution = [distribution] * len(self.resolution)
        self.distribution = list(distribution)
        assert len(self.resolution) == len(self.minimum) == len(self.maximum) == len(self.distribution)
        self._cache: Dict[Any, torch.Tensor] = {}

    @classmethod
    def from_config(cls, config: Dict[str, Any]):
        return cls(config["resolution"], config["minimum"], config["maximum"],
                   transformation=build_transformation(config.get("transformation")),
                   distribution=config.get("distribution"))

    def __getstate__(self):  # the cache holds device tensors; keep pickles (torch.save(model)) lean
        state = self.__dict__.copy()
        state["_cache"] = {}
        return state

    def grid(self, dtype: torch.dtype, device) -> torch.Tensor:
        """(N, dim) query points; data_agnostic.py:147-169 for one sample."""
        key = (dtype, str(device))
        if key not in self._cache:
            axes = []
            for res, lo, hi, dist in zip(self.resolution, self.minimum, self.maximum, self.distribution):
                q = torch.linspace(0.0, 1.0, res, dtype=dtype)
                q = q * 1 if dist == "linear" else getattr(torch, dist)(q)
                span = float(q.max() - q.min())
                if abs(span) <= 1e-8:                       # torch.isclose(den, 0) -> den = 1 (data_agnostic.py:120-122)
                    span = 1.0
                axes.append((q - q.min()) / span * (hi - lo) + lo)
            mesh = torch.meshgrid(*axes, indexing="ij")
            pts = torch.stack([m.flatten() for m in mesh], dim=-1)
            pts = self.transformation(pts.unsqueeze(0)).squeeze(0)
            self._cache[key] = pts.to(device)
        return self._cache[key]

    def forward(self, batch) -> "OrderedDict[str, torch.Tensor]":
        first = batch
        while not isinstance(first, torch.Tensor):
            first = first[next(iter(first))] if isinstance(first, dict) else first[0]
        pts = self.grid(first.dtype, first.device)
        return OrderedDict(center=pts.unsqueeze(0).repeat(first.shape[0], 1, 1))


class DataAgnosticLinearQueries(DataAgnosticStaticQueries):
    def __init__(self, resolution=None, minimum=None, maximum=None, transformation=None, **kwargs):
        super().__init__(resolution, minimum, maximum, transformation=transformation, distribution="linear")


def build_querent(name: str, config: Dict[str, Any], *args, **kwargs):
    """Registry by substring as reference queries/__init__.py:5-9 and data_agnostic.py:203-207."""
    low = name.lower()
    if "data_agnostic" in low:
        if "static" in low:
            return DataAgnosticStaticQueries.from_config(config)
        if "linear" in low:
            return DataAgnosticLinearQueries.from_config(config)
    if "learnable" in low:
        raise NotImplementedError("learnable querents are outside the accelerated hot path (SURVEY.md §2 row 5b)")
    return None
