"""Iterative multi-perspective fusion (mirror of reference src/dprt/models/fusers/mpfusion.py and
src/dprt/models/layers/ms_deform_attn.py).

Module / parameter names follow the reference so its checkpoints load unchanged:
``IMPFusion{query, query_embedding, mpfusion.fusion{i}, heads.{i}}``,
``MPFusion{ml_fusion_layers.ms_deform_attn{v}, reduction_layer}``,
``MLFusion{self_attn, norm1..3, ms_deform_attn, ffn1, ffn2}``,
``MSDeformAttn{sampling_offsets, attention_weights, value_proj, output_proj}``.

Differences in mechanism (not in results):
  * each view's pyramid is flattened ONCE per forward into a ``FeaturePyramid`` (the reference re-flattens and
    ``torch.cat``s all levels in every layer of every iteration, mpfusion.py:172-187);
  * no host synchronisation: the level-shape tensors are cached per shape, the ``transformation.any()`` switch
    (mpfusion.py:647) stays on the device;
  * the sampling core runs in libdpft_b200.so (dpft_b200/msda.py); in ``eval()`` the whole decoder runs in the
    fused kernels of dpft_b200/decoder.py when the configuration is eligible.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from copy import deepcopy
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn.functional as F
from torch import nn

from .. import attention as attention_ops
from .. import msda as msda_ops
from .geometry import cart2spher


class FeaturePyramid:
    """All levels of one view as one (B, S, C) tensor, finest level first, plus the level table."""

    _tables: Dict[Any, Tuple[torch.Tensor, torch.Tensor]] = {}

    def __init__(self, flat: torch.Tensor, shapes: Sequence[Tuple[int, int]]):
        self.flat = flat
        self.shapes = [(int(h), int(w)) for h, w in shapes]
        key = (tuple(self.shapes), str(flat.device))
        if key not in FeaturePyramid._tables:
            sizes = [h * w for h, w in self.shapes]
            starts = [sum(sizes[:i]) for i in range(len(sizes))]
            FeaturePyramid._tables[key] = (torch.tensor(self.shapes, dtype=torch.int64, device=flat.device),
                                           torch.tensor(starts, dtype=torch.int64, device=flat.device))
        self.shapes_t, self.lsi_t = FeaturePyramid._tables[key]

    @classmethod
    def from_levels(cls, levels: "Union[FeaturePyramid, Dict[str, torch.Tensor]]") -> "FeaturePyramid":
        if isinstance(levels, FeaturePyramid):
            return levels
        maps = list(levels.values())
        return cls(torch.cat([m.flatten(1, 2) for m in maps], dim=1), [m.shape[1:3] for m in maps])

    @property
    def n_levels(self) -> int:
        return len(self.shapes)


class MSDeformAttn(nn.Module):
    """Multi-scale deformable attention module; interface of reference ms_deform_attn.py:77-217."""

    def __init__(self, d_model: int = 256, n_levels: int = 4, n_heads: int = 8, n_points: int = 4):
        super().__init__()
        if d_model % n_heads:
            raise ValueError(f"d_model must be divisible by n_heads, but got {d_model} and {n_heads}")
        self.im2col_step = 64
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self.gather_then_project = True      # composed path: project the sampled features instead of all S positions
        self._reset_parameters()

    def _reset_parameters(self) -> None:
        # Deformable-DETR initialisation: zero offset weights, offsets biased along n_heads compass directions
        # and growing with the point index; uniform attention; Xavier projections (ms_deform_attn.py:117-136).
        with torch.no_grad():
            self.sampling_offsets.weight.zero_()
            theta = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
            d = torch.stack((theta.cos(), theta.sin()), -1)
            d = d / d.abs().max(-1, keepdim=True)[0]
            grid = d.view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
            grid = grid * torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, -1, 1)
            self.sampling_offsets.bias.copy_(grid.reshape(-1))
            self.attention_weights.weight.zero_()
            self.attention_weights.bias.zero_()
            nn.init.xavier_uniform_(self.value_proj.weight)
            self.value_proj.bias.zero_()
            nn.init.xavier_uniform_(self.output_proj.weight)
            self.output_proj.bias.zero_()

    def sampling(self, query: torch.Tensor, reference_points: torch.Tensor, shapes_t: torch.Tensor):
        """Returns (sampling_locations (B,N,M,L,P,2), attention_weights (B,N,M,L,P)); ms_deform_attn.py:177-204."""
        B, N, _ = query.shape
        M, L, P = self.n_heads, self.n_levels, self.n_points
        off = self.sampling_offsets(query).view(B, N, M, L, P, 2)
        aw = F.softmax(self.attention_weights(query).view(B, N, M, L * P), -1).view(B, N, M, L, P)
        if reference_points.shape[-1] == 2:
            normalizer = torch.stack((shapes_t[:, 1], shapes_t[:, 0]), -1).to(query.dtype)   # (W, H)
            loc = reference_points[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            loc = reference_points[:, :, None, :, None, :2] + off / P * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError(f"Last dim of reference_points must be 2 or 4, but get {reference_points.shape[-1]} instead.")
        return loc, aw

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        B, S, _ = input_flatten.shape
        if not (reference_points.shape[2] == input_spatial_shapes.shape[0] == input_level_start_index.shape[0]
                == self.n_levels):
            raise AssertionError("reference_points / spatial_shapes / level_start_index do not match n_levels")
        loc, aw = self.sampling(query, reference_points, input_spatial_shapes)
        if self.gather_then_project and input_padding_mask is None:
            return self.output_proj(self._gather_then_project(input_flatten, input_spatial_shapes, input_level_start_index,
                                                              loc, aw))
        value = self.value_proj(input_flatten)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], 0.0)
        value = value.view(B, S, self.n_heads, self.d_model // self.n_heads)
        out = msda_ops.MSDeformAttnFunction.apply(value, input_spatial_shapes, input_level_start_index, loc, aw,
                                                  self.im2col_step)
        return self.output_proj(out)

    def _gather_then_project(self, x, shapes_t, lsi_t, loc, aw):
        """Same result as sampling ``value_proj(x)`` (ms_deform_attn.py:172 then :206-214), without the dense pass over all
        S positions and its weight-gradient reduction over B*S rows: sampling is linear, so
            out[b,q,m,:] = Wv[m] @ G[b,q,m,:] + bv[m] * mass[b,q,m]
        with G = the attention-weighted bilinear samples of the UNPROJECTED features (one launch of the op: the M heads
        become extra queries over a single shared 'head' of all C channels) and mass = the attention-weighted sum of the
        in-bounds bilinear weights (zero padding drops the bias where a corner falls outside)."""
        B, S, C = x.shape
        _, N, M, L, P, _ = loc.shape
        D = self.d_model // M
        G = msda_ops.MSDeformAttnFunction.apply(x.reshape(B, S, 1, C), shapes_t, lsi_t, loc.reshape(B, N * M, 1, L, P, 2),
                                                aw.reshape(B, N * M, 1, L, P), self.im2col_step).view(B, N, M, C)
        wh = shapes_t.to(loc.dtype).flip(-1).view(1, 1, 1, L, 1, 2)            # (W, H) per level
        pix = loc * wh - 0.5
        lo = torch.floor(pix)
        frac = pix - lo
        inb0 = ((lo >= 0) & (lo <= wh - 1)).to(loc.dtype)                      # corner lo inside the map
        inb1 = ((lo + 1 >= 0) & (lo + 1 <= wh - 1)).to(loc.dtype)              # corner lo + 1 inside the map
        axis = (1 - frac) * inb0 + frac * inb1                                 # in-bounds weight along x and y
        mass = (aw * axis[..., 0] * axis[..., 1]).sum((-1, -2))                # (B, N, M)
        wv = self.value_proj.weight.view(M, D, C)
        out = torch.einsum("bnmc,mdc->bnmd", G, wv) + self.value_proj.bias.view(1, 1, M, D) * mass.unsqueeze(-1)
        return out.reshape(B, N, M * D)


class MLFusion(nn.Module):
    """One decoder layer for one view: self-attention -> deformable cross-attention -> FFN, post-norm
    (reference mpfusion.py:16-263)."""

    def __init__(self, d_model: int = 256, d_ffn: int = 1024, n_levels: int = 1, n_heads: int = 1, n_points: int = 1,
                 ffn_layer: str = "Linear", activation: str = "ReLU", dropout: float = 0.0, norm: bool = False,
                 **kwargs):
        super().__init__()
        if ffn_layer != "Linear":
            raise NotImplementedError("only Linear feed-forward layers are on the accelerated path")
        self.d_model, self.d_ffn, self.n_levels, self.n_heads, self.n_points = d_model, d_ffn, n_levels, n_heads, n_points
        self.ffn_layer, self.activation, self.dropout, self.norm = ffn_layer, activation, dropout, norm
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout, batch_first=True)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.ms_deform_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.ffn1 = nn.Linear(d_model, d_ffn)
        self.activation1 = getattr(nn, activation)()
        self.dropout3 = nn.Dropout(dropout)
        self.ffn2 = nn.Linear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)

    @classmethod
    def from_config(cls, config: Dict[str, Any]):
        return cls(**config)

    native_self_attn = True      # inference on CUDA: attention core in the tcgen05 flash-attention kernel (dpft_b200/attention.py)

    def forward_self_attn(self, query, query_positions=None):
        qk = query if query_positions is None else query + query_positions
        if self.native_self_attn and attention_ops.mha_eligible(self.self_attn, query):
            attn = attention_ops.multihead_self_attention(self.self_attn, qk, query)
        else:
            attn = self.self_attn(query=qk, key=qk, value=query, need_weights=False)[0]
        out = query + self.dropout1(attn)
        return self.norm1(out) if self.norm else out

    def forward_cross_attn(self, query, batch, reference_points, query_positions=None):
        pyr = FeaturePyramid.from_levels(batch)
        refs = reference_points.unsqueeze(2).expand(-1, -1, pyr.n_levels, -1)
        q = query if query_positions is None else query + query_positions
        out = query + self.dropout2(self.ms_deform_attn(q, refs, pyr.flat, pyr.shapes_t, pyr.lsi_t))
        return self.norm2(out) if self.norm else out

    def forward_ffn(self, query):
        out = query + self.dropout4(self.ffn2(self.dropout3(self.activation1(self.ffn1(query)))))
        return self.norm3(out) if self.norm else out

    def forward(self, query, batch, reference_points, query_positions=None):
        out = self.forward_self_attn(query, query_positions)
        out = self.forward_cross_attn(out, batch, reference_points, query_positions)
        return self.forward_ffn(out)


class MPFusion(nn.Module):
    """The per-view decoder layers of one iteration plus the view reduction (reference mpfusion.py:266-514)."""

    SUPPORTED = ("mean", "max", "linear")

    def __init__(self, m_views: int, d_model: int = 256, d_ffn: int = 1024, n_levels: List[int] = None,
                 n_heads: List[int] = None, n_points: List[int] = None, ffn_layer: str = "Linear",
                 activation: str = "ReLU", dropout: float = 0.0, norm: bool = False, reduction: str = "mean", **kwargs):
        super().__init__()
        if reduction not in {"mean", "max", "unary", "linear", "cross-attn", "ffn"}:
            raise ValueError("The reduction mode must be one of either 'mean', 'max', 'unary', 'linear', "
                             f"'cross-attn' or 'ffn' but {reduction} was given!")
        if reduction not in self.SUPPORTED:
            raise NotImplementedError(f"reduction {reduction!r} is outside the accelerated hot path")
        self.m_views, self.d_model, self.d_ffn = m_views, d_model, d_ffn
        self.n_levels = list(n_levels) if n_levels is not None else [1] * m_views
        self.n_heads = list(n_heads) if n_heads is not None else [1] * m_views
        self.n_points = list(n_points) if n_points is not None else [1] * m_views
        self.ffn_layer, self.activation, self.dropout, self.norm, self.reduction = ffn_layer, activation, dropout, norm, reduction
        self.ml_fusion_layers = nn.ModuleDict({
            f"ms_deform_attn{v}": MLFusion(d_model, d_ffn, self.n_levels[v], self.n_heads[v], self.n_points[v],
                                           ffn_layer, activation, dropout, norm)
            for v in range(m_views)})
        if reduction == "linear":
            self.reduction_layer = nn.Linear(m_views * d_model, d_model, bias=False)
        else:
            self.reduction_layer = None

    @classmethod
    def from_config(cls, config: Dict[str, Any]):
        return cls(**config)

    def reduce(self, query, queries, query_positions=None):
        """queries (B, N, d_model, m_views) -> (B, N, d_model); 'linear' flattens channel-major, view-minor
        exactly like ``queries.view(B, N, d_model * m_views)`` at reference mpfusion.py:438."""
        if self.reduction == "mean":
            return queries.mean(-1)
        if self.reduction == "max":
            return queries.max(-1)[0]
        B, N = queries.shape[:2]
        return self.reduction_layer(queries.reshape(B, N, self.d_model * self.m_views))

    def forward(self, query, batch, reference_points, query_positions):
        layers = list(zip(self.ml_fusion_layers.values(), batch, reference_points))
        if (self.training and len(layers) > 1 and query.is_cuda and torch.is_grad_enabled()
                and getattr(self, "train_parallel_views", True)):
            # training: the views' layers are independent until the reduction — one stream per view, forward and (because
            # autograd follows the forward's streams) backward; issue order = view order, so dropout masks are unchanged
            from ..streams import fork_map
            outs = fork_map([(lambda l=layer, v=levels, r=ref: l(query, v, r() if callable(r) else r, query_positions))
                             for layer, levels, ref in layers], query.device,
                            reads=[query, query_positions, [getattr(v, "flat", v) for _, v, _ in layers],
                                   getattr(self, "_fork_reads", ())])
        else:
            outs = [layer(query, levels, ref() if callable(ref) else ref, query_positions) for layer, levels, ref in layers]
        return self.reduce(query, torch.stack(outs, dim=-1), query_positions)


class IMPFusion(nn.Module):
    """Iterative refinement: reference points from the current centres -> fusion -> head, ``i_iter`` times
    (reference mpfusion.py:517-745)."""

    def __init__(self, i_iter: int = 1, m_views: int = 1, d_model: int = 256, d_ffn: int = 1024, n_queries: int = 100,
                 n_levels: List[int] = None, n_heads: List[int] = None, n_points: List[int] = None,
                 q_init: str = "uniform_", ffn_layer: str = "Linear", activation: str = "ReLU", dropout: float = 0.0,
                 norm: bool = False, reduction: str = "mean", head: Optional[nn.Module] = None, **kwargs):
        super().__init__()
        self.i_iter, self.m_views, self.d_model, self.d_ffn, self.n_queries = i_iter, m_views, d_model, d_ffn, n_queries
        self.n_levels = list(n_levels) if n_levels is not None else [1] * m_views
        self.n_heads = list(n_heads) if n_heads is not None else [1] * m_views
        self.n_points = list(n_points) if n_points is not None else [1] * m_views
        self.ffn_layer, self.activation, self.dropout, self.norm, self.reduction = ffn_layer, activation, dropout, norm, reduction
        self.q_init = q_init
        if head is None:
            head = nn.Identity()
        self.mpfusion = nn.ModuleDict({
            f"fusion{i}": MPFusion(m_views, d_model, d_ffn, self.n_levels, self.n_heads, self.n_points, ffn_layer,
                                   activation, dropout, norm, reduction)
            for i in range(i_iter)})
        self.heads = nn.ModuleList(deepcopy(head) for _ in range(i_iter))
        self.query_embedding = nn.Embedding(n_queries, d_model)
        self.query = nn.Parameter(torch.empty(n_queries, d_model))
        self.reset_parameters()

    @classmethod
    def from_config(cls, config: Dict[str, Any], **kwargs):
        return cls(**config, **kwargs)

    def reset_parameters(self) -> None:
        getattr(nn.init, self.q_init)(self.query)

    @staticmethod
    def get_reference_points(query: torch.Tensor, transformation: torch.Tensor, projection: torch.Tensor,
                             shape: torch.Tensor) -> torch.Tensor:
        """Projects query centres (B,N,3) into a view: optional rigid transform + Cartesian->spherical when the
        transformation is non-zero (radar), then the (3|4)x4 projection, perspective division where w != 0,
        normalisation by the ORIGINAL input (W, H), clip to [0,1].  Returns (B,N,2) as (u, v) = (x, y).
        Reference mpfusion.py:617-696; computed without leaving the device."""
        cart = query[..., :3]
        ones = torch.ones_like(cart[..., :1])
        use_t = transformation.any()
        eye = torch.eye(4, dtype=transformation.dtype, device=transformation.device).expand_as(transformation)
        t_safe = torch.where(use_t, transformation, eye)          # keeps the unused branch finite for autograd
        moved = torch.einsum("bij,bkj->bki", t_safe, torch.cat((cart, ones), -1))[..., :3]
        pts = torch.where(use_t, cart2spher(moved, degrees=True), cart)
        proj = torch.einsum("bij,bkj->bki", projection, torch.cat((pts, ones), -1))
        w = proj[..., 2]
        nz = w != 0
        safe = torch.where(nz, w, torch.ones_like(w))
        u = torch.where(nz, proj[..., 0] / safe, proj[..., 0]) / shape[:, 1].unsqueeze(1)
        v = torch.where(nz, proj[..., 1] / safe, proj[..., 1]) / shape[:, 0].unsqueeze(1)
        return torch.clip(torch.stack((u, v), -1), min=0.0, max=1.0)

    def forward(self, batch, shape, projection, out):
        B = out["center"].shape[0]
        query = self.query.unsqueeze(0).expand(B, -1, -1)
        query_pos = self.query_embedding.weight.unsqueeze(0).expand(B, -1, -1)
        pyramids = [FeaturePyramid.from_levels(levels) for levels in batch]
        for layer, head in zip(self.mpfusion.values(), self.heads):
            center = out["center"][..., :3]
            # handed over as thunks: MPFusion evaluates each view's projection inside that view's (possibly forked) branch
            refs = [(lambda t=t, p=p, s=s: self.get_reference_points(center, t, p, s)) for (t, p), s in zip(projection, shape)]
            layer._fork_reads = [center, [list(tp) for tp in projection], list(shape)]      # what the thunks read (streams.fork_map)
            query = layer(query, pyramids, refs, query_pos)
            layer._fork_reads = ()
            out = head(query, out)
        return out


def build_fuser(name: str, config: Dict[str, Any], *args, **kwargs):
    if "impfusion" in name.lower():
        return IMPFusion.from_config(config, **kwargs)
    return None
