"""Coordinate transforms of the DPRT query path (mirror of reference src/dprt/models/utils/transformations.py).

Angles are in degrees as in the reference (``degrees=True`` is what every shipped config uses); phi is the
azimuth from the x-axis, roh the elevation from the x-y plane.
"""
from __future__ import annotations

import torch
from torch import nn


def spher2cart(pts: torch.Tensor, degrees: bool = True) -> torch.Tensor:
    """(…,3) (r, phi, roh) -> (…,3) (x, y, z); reference transformations.py:212-255."""
    r, phi, roh = pts.unbind(-1)
    if degrees:
        phi, roh = torch.deg2rad(phi), torch.deg2rad(roh)
    cr = torch.cos(roh)
    return torch.stack((r * torch.cos(phi) * cr, r * torch.sin(phi) * cr, r * torch.sin(roh)), -1)


def cart2spher(pts: torch.Tensor, degrees: bool = True) -> torch.Tensor:
    """(…,3) (x, y, z) -> (…,3) (r, phi, roh) with roh = 0 where r = 0; reference transformations.py:71-120."""
    x, y, z = pts.unbind(-1)
    r = torch.sqrt(x * x + y * y + z * z)
    nz = r != 0
    roh = torch.asin(torch.where(nz, z / torch.where(nz, r, torch.ones_like(r)), torch.zeros_like(z)))
    phi = torch.atan2(y, x)
    if degrees:
        phi, roh = torch.rad2deg(phi), torch.rad2deg(roh)
    return torch.stack((r, phi, roh), -1)


def polar2cart(pts: torch.Tensor, degrees: bool = True) -> torch.Tensor:
    r, phi = pts.unbind(-1)
    if degrees:
        phi = torch.deg2rad(phi)
    return torch.stack((r * torch.cos(phi), r * torch.sin(phi)), -1)


def cart2polar(pts: torch.Tensor, degrees: bool = True) -> torch.Tensor:
    x, y = pts.unbind(-1)
    phi = torch.atan2(y, x)
    return torch.stack((torch.sqrt(x * x + y * y), torch.rad2deg(phi) if degrees else phi), -1)


class _Transform(nn.Module):
    fn = None

    def __init__(self, dim: int = -1, degrees: bool = True, **kwargs):
        super().__init__()
        self.dim, self.degrees = dim, degrees

    def forward(self, batch: torch.Tensor) -> torch.Tensor:
        return type(self).fn(batch.movedim(self.dim, -1), self.degrees).movedim(-1, self.dim)


class Spher2Cart(_Transform):
    fn = staticmethod(spher2cart)


class Cart2Spher(_Transform):
    fn = staticmethod(cart2spher)


class Polar2Cart(_Transform):
    fn = staticmethod(polar2cart)


class Cart2Polar(_Transform):
    fn = staticmethod(cart2polar)


def build_transformation(name, *args, **kwargs):
    """Same substring registry as reference transformations.py:284-294."""
    if name is None:
        return None
    low = name.lower()
    for key, cls in (("polar2cart", Polar2Cart), ("spher2cart", Spher2Cart), ("cart2polar", Cart2Polar),
                     ("cart2spher", Cart2Spher)):
        if key in low:
            return cls(*args, **kwargs)
    return None
