"""Oracle for the multi-scale deformable attention op (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Two independent restatements of the op DPFT reaches through ``MSDA.ms_deform_attn_forward/backward``
(reference src/dprt/models/layers/ms_deform_attn.py:32-39, :58-66):

* ``msda_forward_torch`` / ``msda_backward_torch`` — the op's documented PyTorch equivalent
  (``grid_sample(2*loc-1, bilinear, zeros, align_corners=False)`` then the attention-weighted sum);
  works in fp32 and fp64 and gives the three gradients through autograd.
* ``msda_forward_c`` / ``msda_backward_c`` — ctypes binding of ``oracle/msda_oracle.c`` (explicit corner
  arithmetic with per-corner bounds checks), fp32 and fp64.

The external extension is un-pinned and absent: parity for this op is anchored on the
reference's call contract (ms_deform_attn.py:27-68, :145-161) and on these two agreeing with each other
and with the HuggingFace restatement shipped in this image (tests/test_oracle.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Sequence, Tuple

import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_oracle(force: bool = False) -> str:
    """Compiles oracle/msda_oracle.c with gcc (recipe: oracle/Makefile) and returns the .so path."""
    so = os.path.join(_HERE, "_build", "libmsda_oracle.so")
    src = os.path.join(_HERE, "msda_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_build/libmsda_oracle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c_oracle())
    return _LIB


def _shapes_list(shapes) -> Sequence[Tuple[int, int]]:
    if isinstance(shapes, torch.Tensor):
        return [(int(h), int(w)) for h, w in shapes.tolist()]
    return [(int(h), int(w)) for h, w in shapes]


def level_start_index(shapes) -> torch.Tensor:
    sizes = [h * w for h, w in _shapes_list(shapes)]
    out, acc = [], 0
    for s in sizes:
        out.append(acc)
        acc += s
    return torch.tensor(out, dtype=torch.int64)


def msda_forward_torch(value: torch.Tensor, shapes, loc: torch.Tensor, attn: torch.Tensor) -> torch.Tensor:
    """value (B,S,M,D), loc (B,N,M,L,P,2) as (x,y), attn (B,N,M,L,P) -> (B,N,M*D)."""
    B, S, M, D = value.shape
    _, N, _, L, P, _ = loc.shape
    out = value.new_zeros(B, M, D, N)
    start = 0
    for lvl, (H, W) in enumerate(_shapes_list(shapes)):
        v = value[:, start:start + H * W].permute(0, 2, 3, 1).reshape(B * M, D, H, W)
        grid = (2.0 * loc[:, :, :, lvl] - 1.0).permute(0, 2, 1, 3, 4).reshape(B * M, N, P, 2)
        sampled = F.grid_sample(v, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
        w = attn[:, :, :, lvl].permute(0, 2, 1, 3).reshape(B * M, 1, N, P)
        out = out + (sampled * w).sum(-1).view(B, M, D, N)
        start += H * W
    return out.permute(0, 3, 1, 2).reshape(B, N, M * D).contiguous()


def msda_backward_torch(value, shapes, loc, attn, grad_out):
    """Returns (grad_value, grad_loc, grad_attn) of ``msda_forward_torch`` through autograd."""
    with torch.enable_grad():  # callers may sit inside a once_differentiable backward
        v = value.detach().clone().requires_grad_(True)
        lo = loc.detach().clone().requires_grad_(True)
        a = attn.detach().clone().requires_grad_(True)
        out = msda_forward_torch(v, shapes, lo, a)
        gv, gl, ga = torch.autograd.grad(out, (v, lo, a), grad_out)
    return gv, gl, ga


def _ptr(t: torch.Tensor):
    return ctypes.c_void_p(t.data_ptr())


def _suffix(dtype):
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise TypeError(f"C oracle handles float32/float64 only, got {dtype}")


def msda_forward_c(value, shapes, loc, attn):
    value, loc, attn = value.contiguous().cpu(), loc.contiguous().cpu(), attn.contiguous().cpu()
    B, S, M, D = value.shape
    _, N, _, L, P, _ = loc.shape
    sh = torch.tensor(_shapes_list(shapes), dtype=torch.int64).contiguous()
    lsi = level_start_index(shapes)
    out = torch.empty(B, N, M * D, dtype=value.dtype)
    fn = getattr(_lib(), "msda_oracle_fwd_" + _suffix(value.dtype))
    fn(_ptr(value), _ptr(sh), _ptr(lsi), _ptr(loc), _ptr(attn), _ptr(out),
       *[ctypes.c_int(x) for x in (B, S, M, D, N, L, P)])
    return out


def msda_backward_c(value, shapes, loc, attn, grad_out):
    value, loc, attn = value.contiguous().cpu(), loc.contiguous().cpu(), attn.contiguous().cpu()
    grad_out = grad_out.contiguous().cpu()
    B, S, M, D = value.shape
    _, N, _, L, P, _ = loc.shape
    sh = torch.tensor(_shapes_list(shapes), dtype=torch.int64).contiguous()
    lsi = level_start_index(shapes)
    gv = torch.zeros_like(value)
    gl = torch.empty_like(loc)
    ga = torch.empty_like(attn)
    fn = getattr(_lib(), "msda_oracle_bwd_" + _suffix(value.dtype))
    fn(_ptr(value), _ptr(sh), _ptr(lsi), _ptr(loc), _ptr(attn), _ptr(grad_out),
       _ptr(gv), _ptr(gl), _ptr(ga), *[ctypes.c_int(x) for x in (B, S, M, D, N, L, P)])
    return gv, gl, ga


def random_problem(B, N, M, D, shapes, P, dtype=torch.float32, seed=0, spread=0.15):
    """Seeded synthetic op inputs.  Locations deliberately leave [0,1] (``spread``) so the zero-padding
    branches are exercised; attention weights are a softmax over L*P as in ms_deform_attn.py:179-182."""
    g = torch.Generator().manual_seed(seed)
    shapes = _shapes_list(shapes)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    value = torch.randn(B, S, M, D, generator=g, dtype=torch.float64).to(dtype)
    loc = (torch.rand(B, N, M, L, P, 2, generator=g, dtype=torch.float64) * (1 + 2 * spread) - spread).to(dtype)
    attn = torch.softmax(torch.randn(B, N, M, L * P, generator=g, dtype=torch.float64), -1)
    attn = attn.view(B, N, M, L, P).to(dtype)
    grad_out = torch.randn(B, N, M * D, generator=g, dtype=torch.float64).to(dtype)
    return value, shapes, loc, attn, grad_out
