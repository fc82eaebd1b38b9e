"""CPU oracle of the DPRT model forward (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

A functional, straight-line restatement in plain PyTorch (fp32 or fp64, CPU) of what
``DPRT.forward(batch)`` computes in ``eval()`` mode (reference src/dprt/models/dprt.py:200-244), working
directly on a reference ``state_dict`` and the reference JSON config.  It exists because the reference
package cannot travel to the GPU box; it is pinned against the real reference (imported from
/root/reference in the build container) by tools/make_golden.py -> tests/golden/*.pt and
tests/test_oracle_model.py.

Every function cites the reference lines it follows.  The deformable sampling op itself comes from
oracle/msda.py (the external dependency, see its header).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import msda as msda_oracle

RESNET_BLOCKS = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3), "resnet152": (3, 8, 36, 3)}


# ------------------------------------------------------------------------------------------------ backbone
def _bn(sd, p, x, eps=1e-5):
    # torchvision BatchNorm2d in eval mode: running statistics (backbones/resnet.py:169-176 picks nn.BatchNorm2d)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, eps)


def _bottleneck(sd, p, x, stride):
    # torchvision.models.resnet.Bottleneck (v1.5: the stride sits on the 3x3 conv)
    out = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"])))
    out = F.relu(_bn(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1)))
    out = _bn(sd, p + ".bn3", F.conv2d(out, sd[p + ".conv3.weight"]))
    if (p + ".downsample.0.weight") in sd:
        x = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride))
    return F.relu(out + x)


def backbone(sd, prefix: str, cfg: dict, x_nhwc: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
    """backbones/resnet.py:80-107: NHWC -> NCHW, optional 1x1 adjustment conv, ResNet stem + stages,
    returns {'1'..str(multi_scale)} as NHWC."""
    x = x_nhwc.movedim(-1, 1)
    if (prefix + ".adjustment_layer.weight") in sd:                      # resnet.py:47-51
        x = F.conv2d(x, sd[prefix + ".adjustment_layer.weight"])
    b = prefix + ".body"
    x = F.relu(_bn(sd, b + ".bn1", F.conv2d(x, sd[b + ".conv1.weight"], stride=2, padding=3)))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    out = OrderedDict()
    blocks = RESNET_BLOCKS[cfg["name"].lower()]
    for stage in range(cfg.get("multi_scale", 1)):                       # resnet.py:54-55
        for i in range(blocks[stage]):
            stride = 2 if (i == 0 and stage > 0) else 1
            x = _bottleneck(sd, f"{b}.layer{stage + 1}.{i}", x, stride)
        out[str(stage + 1)] = x.movedim(1, -1)                           # resnet.py:104-105
    return out


# ---------------------------------------------------------------------------------------------------- neck
def fpn(sd, prefix: str, feats: "OrderedDict[str, torch.Tensor]") -> "OrderedDict[str, torch.Tensor]":
    """necks/fpn.py:70-83 -> torchvision FeaturePyramidNetwork.forward: 1x1 lateral (+bias), top-down
    nearest upsample-add, 3x3 output conv (+bias); no norm, no activation."""
    names = list(feats.keys())
    xs = [f.movedim(-1, 1) for f in feats.values()]
    p = prefix + ".fpn"
    n = len(xs)

    def inner(i, x):
        return F.conv2d(x, sd[f"{p}.inner_blocks.{i}.0.weight"], sd[f"{p}.inner_blocks.{i}.0.bias"])

    def layer(i, x):
        return F.conv2d(x, sd[f"{p}.layer_blocks.{i}.0.weight"], sd[f"{p}.layer_blocks.{i}.0.bias"], padding=1)

    last = inner(n - 1, xs[-1])
    results = [layer(n - 1, last)]
    for i in range(n - 2, -1, -1):
        lateral = inner(i, xs[i])
        last = lateral + F.interpolate(last, size=lateral.shape[-2:], mode="nearest")
        results.insert(0, layer(i, last))
    return OrderedDict((k, r.movedim(1, -1)) for k, r in zip(names, results))


# ----------------------------------------------------------------------------------------------- embedding
def sine_embedding(feat: torch.Tensor, num_feats: int, normalize: bool = False, temperature: float = 10000,
                   scale: float = 2 * math.pi, eps: float = 1e-6, offset: float = 0.0) -> torch.Tensor:
    """embeddings/sinusoidal.py:63-110; returns feat + pos_x + pos_y (the reference adds in place)."""
    B, H, W, _ = feat.shape
    dt = feat.dtype
    y = torch.arange(1, H + 1, dtype=dt).view(1, H, 1).expand(B, H, W)     # cumsum of ones, :83-84
    x = torch.arange(1, W + 1, dtype=dt).view(1, 1, W).expand(B, H, W)
    if normalize:                                                          # :86-90
        y = (y + offset) / (y[:, -1:, :] + eps) * scale
        x = (x + offset) / (x[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_feats, dtype=dt)
    dim_t = temperature ** (2 * (dim_t // 2) / num_feats)                  # :92-94
    px = x[..., None] / dim_t
    py = y[..., None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).reshape(B, H, W, -1)  # :99-104
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).reshape(B, H, W, -1)
    return feat + px + py                                                  # :107-108


# ------------------------------------------------------------------------------------------------- querent
def static_queries(cfg: dict, B: int, dtype=torch.float32) -> torch.Tensor:
    """queries/data_agnostic.py:126-172 with the 'linear' distribution and the config's transformation."""
    axes = []
    for res, mi, ma in zip(cfg["resolution"], cfg["minimum"], cfg["maximum"]):
        q = torch.linspace(0.0, 1.0, res, dtype=dtype) * 1                 # :147-153
        den = q.max() - q.min()                                            # _min_max_scaling :117-124
        if torch.isclose(den, torch.zeros_like(den)):
            den = 1.0
        axes.append((q - q.min()) / den * (ma - mi) + mi)
    grid = torch.meshgrid(*axes, indexing="ij")                            # :162
    pts = torch.stack([g.flatten() for g in grid], dim=-1)                 # :163
    pts = pts.unsqueeze(0).repeat(B, 1, 1)                                 # :166
    name = (cfg.get("transformation") or "").lower()
    if "spher2cart" in name:                                               # utils/transformations.py:212-281
        r, phi, roh = pts[..., 0], pts[..., 1], pts[..., 2]
        phi, roh = torch.deg2rad(phi), torch.deg2rad(roh)
        pts = torch.stack((r * torch.cos(phi) * torch.cos(roh), r * torch.sin(phi) * torch.cos(roh),
                           r * torch.sin(roh)), dim=-1)
    elif name:
        raise NotImplementedError(f"oracle: querent transformation {name!r}")
    return pts


# -------------------------------------------------------------------------------------------------- fuser
def reference_points(center: torch.Tensor, T: torch.Tensor, P: torch.Tensor, shape_hw: torch.Tensor) -> torch.Tensor:
    """fusers/mpfusion.py:617-696 (+ cart2spher, utils/transformations.py:71-120).  Returns (B,N,2) as (u,v)."""
    ones = torch.ones_like(center[..., :1])
    if bool(T.any()):                                                      # :647
        q = torch.einsum("bij,bkj->bki", T, torch.cat((center[..., :3], ones), -1))  # :649-653
        x, y, z = q[..., 0], q[..., 1], q[..., 2]
        r = torch.sqrt(x * x + y * y + z * z)                              # transformations.py:104
        phi = torch.atan2(y, x)
        c = torch.where(r != 0, z / torch.where(r != 0, r, torch.ones_like(r)), torch.zeros_like(z))  # :108-110
        roh = torch.asin(c)
        pts = torch.stack((r, torch.rad2deg(phi), torch.rad2deg(roh)), -1)  # :114-116, mpfusion.py:663
    else:
        pts = center[..., :3]                                              # :665-666
    proj = torch.einsum("bij,bkj->bki", P, torch.cat((pts, ones), -1))     # :669-673
    w = proj[..., 2]
    nz = w != 0                                                            # :676
    safe = torch.where(nz, w, torch.ones_like(w))
    u = torch.where(nz, proj[..., 0] / safe, proj[..., 0])                 # :679-680
    v = torch.where(nz, proj[..., 1] / safe, proj[..., 1])                 # :683-684
    u = u / shape_hw[:, 1].unsqueeze(1)                                    # :687
    v = v / shape_hw[:, 0].unsqueeze(1)                                    # :688
    return torch.clip(torch.stack((u, v), -1), 0.0, 1.0)                   # :691-694


def _layer_norm(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def self_attention(sd, p: str, x: torch.Tensor, pos: torch.Tensor, n_heads: int) -> torch.Tensor:
    """nn.MultiheadAttention(batch_first) with q = k = x + pos, v = x (mpfusion.py:122-148), eval mode."""
    B, N, E = x.shape
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    qk = x + pos
    q = F.linear(qk, w[:E], b[:E]).view(B, N, n_heads, E // n_heads).transpose(1, 2)
    k = F.linear(qk, w[E:2 * E], b[E:2 * E]).view(B, N, n_heads, E // n_heads).transpose(1, 2)
    v = F.linear(x, w[2 * E:], b[2 * E:]).view(B, N, n_heads, E // n_heads).transpose(1, 2)
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(E // n_heads), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, N, E)
    return F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def deformable_attention(sd, p: str, query, ref_points, flat, shapes: Sequence[Tuple[int, int]],
                         n_heads: int, n_points: int) -> torch.Tensor:
    """layers/ms_deform_attn.py:138-217 with 2-d reference points."""
    B, N, C = query.shape
    L = len(shapes)
    S = flat.shape[1]
    assert sum(h * w for h, w in shapes) == S                              # :166
    value = F.linear(flat, sd[p + ".value_proj.weight"], sd[p + ".value_proj.bias"])      # :172
    value = value.view(B, S, n_heads, C // n_heads)
    off = F.linear(query, sd[p + ".sampling_offsets.weight"], sd[p + ".sampling_offsets.bias"])
    off = off.view(B, N, n_heads, L, n_points, 2)                          # :177-178
    aw = F.linear(query, sd[p + ".attention_weights.weight"], sd[p + ".attention_weights.bias"])
    aw = torch.softmax(aw.view(B, N, n_heads, L * n_points), -1).view(B, N, n_heads, L, n_points)  # :179-182
    norm = torch.tensor([[w, h] for h, w in shapes], dtype=query.dtype)    # :186-188  (W, H)
    loc = ref_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]     # :189-191
    out = msda_oracle.msda_forward_torch(value, shapes, loc, aw)           # :206-213
    return F.linear(out, sd[p + ".output_proj.weight"], sd[p + ".output_proj.bias"])      # :215


def ml_fusion(sd, p: str, fcfg: dict, view: int, query, pos, levels: "OrderedDict[str, torch.Tensor]", ref,
              taps: dict = None, tap_key: str = ""):
    """MLFusion.forward (mpfusion.py:231-263), eval mode (dropouts are identities), norm optional."""
    n_heads, n_points = fcfg["n_heads"][view], fcfg["n_points"][view]
    norm = fcfg.get("norm", False)
    x = query + self_attention(sd, p + ".self_attn", query, pos, n_heads)  # :253, :122-148
    if norm:
        x = _layer_norm(sd, p + ".norm1", x)
    shapes = [(int(l.shape[1]), int(l.shape[2])) for l in levels.values()]  # :172-176
    flat = torch.cat([l.flatten(1, 2) for l in levels.values()], dim=1)    # :179
    refs = ref.unsqueeze(2).repeat(1, 1, len(shapes), 1)                   # :190
    cross = deformable_attention(sd, p + ".ms_deform_attn", x + pos, refs, flat, shapes, n_heads, n_points)
    if taps is not None:
        taps[tap_key] = cross.clone()
    x = x + cross
    if norm:
        x = _layer_norm(sd, p + ".norm2", x)                               # :202-206
    act = getattr(torch.nn, fcfg.get("activation", "ReLU"))()
    y = F.linear(act(F.linear(x, sd[p + ".ffn1.weight"], sd[p + ".ffn1.bias"])),
                 sd[p + ".ffn2.weight"], sd[p + ".ffn2.bias"])             # :221
    x = x + y
    if norm:
        x = _layer_norm(sd, p + ".norm3", x)
    return x


def detection_head(sd, p: str, hcfg: dict, query, center_prev):
    """LinearDetectionHead.forward (heads/detection.py:252-275): four bias-free MLPs + activations."""
    acts = OrderedDict(center=lambda t: t, size=F.relu, angle=torch.tanh)
    acts["class"] = lambda t: t
    n_layers = {"center": hcfg.get("num_reg_layers", 1), "size": hcfg.get("num_reg_layers", 1),
                "angle": hcfg.get("num_reg_layers", 1), "class": hcfg.get("num_cls_layers", 1)}
    out = OrderedDict()
    for k, act in acts.items():
        h = query
        for i in range(n_layers[k]):
            w = sd[f"{p}.layers.{k}_head.{3 * i}.weight"]
            b = sd.get(f"{p}.layers.{k}_head.{3 * i}.bias")
            h = F.linear(h, w, b)
            if i < n_layers[k] - 1:
                h = F.relu(h)                                              # :239-241 (Dropout(0.0) follows)
        out[k] = act(h)
    out["center"] = out["center"].clone()
    out["center"][..., :3] += center_prev[..., :3]                         # :273
    return out


def fuser(sd, cfg: dict, feats: List["OrderedDict[str, torch.Tensor]"], shapes_hw: List[torch.Tensor],
          projections: List[Tuple[torch.Tensor, torch.Tensor]], center: torch.Tensor, taps: dict = None):
    """IMPFusion.forward (mpfusion.py:698-745) over MPFusion.forward (:472-514) with reduction 'linear'."""
    fcfg, hcfg = cfg["model"]["fuser"], cfg["model"]["head"]
    B = center.shape[0]
    query = sd["fuser.query"].unsqueeze(0).repeat(B, 1, 1)                 # :727
    pos = sd["fuser.query_embedding.weight"].unsqueeze(0).repeat(B, 1, 1)  # :730
    out = OrderedDict(center=center)
    V, C = fcfg["m_views"], fcfg["d_model"]
    for it in range(fcfg["i_iter"]):                                       # :732
        refs = [reference_points(out["center"][..., :3], T, P, s) for (T, P), s in zip(projections, shapes_hw)]
        per_view = []
        for v in range(V):                                                 # :496-509
            p = f"fuser.mpfusion.fusion{it}.ml_fusion_layers.ms_deform_attn{v}"
            per_view.append(ml_fusion(sd, p, fcfg, v, query, pos, feats[v], refs[v], taps, f"msda_out_{it}_{v}"))
        stacked = torch.stack(per_view, dim=-1)                            # (B,N,C,V)
        red = fcfg.get("reduction", "mean")
        if red == "linear":                                                # :438: view(B,N,C*V), channel-major
            query = F.linear(stacked.reshape(B, -1, C * V), sd[f"fuser.mpfusion.fusion{it}.reduction_layer.weight"])
        elif red == "mean":
            query = stacked.mean(-1)
        else:
            raise NotImplementedError(f"oracle: reduction {red!r}")
        out = detection_head(sd, f"fuser.heads.{it}", hcfg, query, out["center"])   # :743
        if taps is not None:
            taps[f"ref_points_{it}"] = [r.clone() for r in refs]
            taps[f"query_{it}"] = query.clone()
            taps[f"center_{it}"] = out["center"].clone()
    return out


# --------------------------------------------------------------------------------------------------- model
def forward(sd: Dict[str, torch.Tensor], cfg: dict, batch: Dict[str, torch.Tensor], taps: dict = None):
    """DPRT.forward in eval mode (models/dprt.py:200-244).  Returns OrderedDict(center,size,angle,class)."""
    m = cfg["model"]
    inputs = m["inputs"]
    feats = []
    for name in inputs:
        x = batch[name]
        f = backbone(sd, f"backbones.{name}", m["backbones"][name], x)     # dprt.py:219
        if m.get("skiplinks", {}).get(name, False):                        # :222-225
            f["0"] = x
            f.move_to_end("0", last=False)
        f = fpn(sd, f"necks.{name}", f)                                    # :228
        ecfg = m["embeddings"][name]
        f = OrderedDict((k, sine_embedding(t, ecfg["num_feats"], ecfg.get("normalize", False)))
                        for k, t in f.items())                            # :231
        feats.append(f)
        if taps is not None:
            taps[f"features_{name}"] = [t.clone() for t in f.values()]
    B = batch[inputs[0]].shape[0]
    center = static_queries(m["querent"], B, batch[inputs[0]].dtype)       # :234
    shapes_hw = [batch[f"{n}_shape"][:, :2] for n in inputs]               # :216, :239
    projections = [(batch[f"label_to_{n}_t"], batch[f"label_to_{n}_p"]) for n in inputs]  # :188-198
    return fuser(sd, cfg, feats, shapes_hw, projections, center, taps)     # :237-242
