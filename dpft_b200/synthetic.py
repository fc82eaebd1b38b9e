"""Seeded synthetic inputs and weights for the DPRT hot path.

``synthetic_batch`` produces the dictionary ``KRadarDataset.__getitem__`` + ``listed_collating`` would hand to
``DPRT.forward`` (reference src/dprt/datasets/kradar/dataset.py:120-181; projection matrices :259-293, the
zero camera transformation :205, ``*_shape`` taken before the camera resize :165-169).  There is no dataset
on the box (16 TB), so tests and bench.py use this.  ``seeded_state_dict`` fills a state-dict template with
well-conditioned values that depend only on (key, shape, seed), so the same weights can be rebuilt on any
machine without shipping them.
"""
from __future__ import annotations

import copy
import json
import os
import zlib
from typing import Dict, Optional, Tuple

import torch

from .configs import VIEWS, make_config

# K-Radar raster constants (reference src/dprt/datasets/kradar/utils/radar_info.py:3-13,33-37):
# 256 range bins up to 118.037 m, 107 azimuth bins (1 degree), 37 elevation bins (1 degree).
RANGE_BINS, RANGE_MAX, AZIMUTH_BINS, ELEVATION_BINS = 256, 118.03710938, 107, 37

DEFAULT_SIZES = {"camera_mono": (512, 910, 3), "radar_bev": (256, 107, 6), "radar_front": (37, 107, 6)}
# BASELINE.json: "synthetic 1280x720 camera + 256x256 radar-cube projections"
BASELINE_SIZES = {"camera_mono": (720, 1280, 3), "radar_bev": (256, 256, 6), "radar_front": (256, 256, 6)}


def load_config(name_or_path: str) -> dict:
    """A shipped config by name (``kradar``, ``kradar_radar_bev`` ...; dpft_b200/configs.py) or any DPFT JSON file."""
    if name_or_path in VIEWS:
        return make_config(name_or_path)
    with open(name_or_path) as f:
        return json.load(f)


def offline_config(cfg: dict, n_queries: Optional[Tuple[int, int, int]] = None, dropout: Optional[float] = None,
                   multi_scale: Optional[int] = None, d_model: Optional[int] = None) -> dict:
    """Config edits SURVEY.md §8d lists: no ImageNet download, optional query-grid / level-count / width change."""
    cfg = copy.deepcopy(cfg)
    m = cfg["model"]
    for b in m["backbones"].values():
        b["weights"] = ""                                  # resnet.py:151-152 random-init path (no network)
    if n_queries is not None:
        m["querent"]["resolution"] = list(n_queries)
        m["fuser"]["n_queries"] = int(n_queries[0] * n_queries[1] * n_queries[2])
    if dropout is not None:
        m["fuser"]["dropout"] = dropout
    if multi_scale is not None:                            # 4 levels = raw + 3 stages (SURVEY §8d cfg 5)
        chans = [256, 512, 1024, 2048][:multi_scale]
        for name in m["inputs"]:
            m["backbones"][name]["multi_scale"] = multi_scale
            first = m["necks"][name]["in_channels_list"][0]
            m["necks"][name]["in_channels_list"] = [first] + chans
            m["embeddings"][name]["n_levels"] = multi_scale + 1
        m["fuser"]["n_levels"] = [multi_scale + 1] * m["fuser"]["m_views"]
    if d_model is not None:                                # SURVEY §8d cfg 5 sweeps d_model in {16, 64, 256}
        for name in m["inputs"]:
            m["necks"][name]["out_channels"] = d_model
            m["embeddings"][name]["num_feats"] = d_model
        m["fuser"]["d_model"], m["fuser"]["d_ffn"] = d_model, 2 * d_model
        m["head"]["in_channels"] = d_model
    return cfg


def _camera_projection(h: int, w: int) -> torch.Tensor:
    # pinhole with lidar axes (x forward, y left, z up): u = cx - f*y/x, v = cy - f*z/x
    f, cx, cy = 0.55 * w, w / 2.0, h / 2.0
    return torch.tensor([[cx, -f, 0.0, 0.0], [cy, 0.0, -f, 0.0], [1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]])


def synthetic_batch(cfg: dict, batch_size: int, seed: int = 0, sizes: Optional[Dict[str, Tuple[int, int, int]]] = None,
                    dtype: torch.dtype = torch.float32, device="cpu") -> Dict[str, torch.Tensor]:
    """Random 0..255 inputs plus the calibration tensors of the dataset contract, for cfg['model']['inputs']."""
    sizes = {**DEFAULT_SIZES, **(sizes or {})}
    g = torch.Generator().manual_seed(seed)
    batch: Dict[str, torch.Tensor] = {}
    for name in cfg["model"]["inputs"]:
        H, W, C = sizes[name]
        batch[name] = (torch.rand(batch_size, H, W, C, generator=g) * 255.0).to(dtype)
        batch[f"{name}_shape"] = torch.tensor([[H, W, C]] * batch_size, dtype=torch.int64)
        if name.startswith("camera"):
            t = torch.zeros(4, 4)                                          # dataset.py:205
            p = _camera_projection(H, W)
        else:
            t = torch.eye(4)
            t[:3, 3] = (torch.rand(3, generator=g) - 0.5) * 0.2            # small calibration offset
            if name == "radar_bev":                                        # dataset.py:277-293
                p = torch.tensor([[0.0, -1.0, 0.0, (W - 1) / 2.0], [H / RANGE_MAX, 0.0, 0.0, 0.0],
                                  [0.0, 0.0, 0.0, 1.0]])
            else:                                                          # dataset.py:259-275
                p = torch.tensor([[0.0, -1.0, 0.0, (W - 1) / 2.0], [0.0, 0.0, 1.0, (H - 1) / 2.0],
                                  [0.0, 0.0, 0.0, 1.0]])
        batch[f"label_to_{name}_t"] = t.to(dtype).unsqueeze(0).repeat(batch_size, 1, 1)
        batch[f"label_to_{name}_p"] = p.to(dtype).unsqueeze(0).repeat(batch_size, 1, 1)
    if device != "cpu":
        batch = {k: v.to(device) for k, v in batch.items()}
    return batch


def _gen(key: str, seed: int) -> torch.Generator:
    return torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def seeded_state_dict(template: Dict[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Returns a state dict with the template's keys/shapes/dtypes and values drawn per key from (key, seed).

    Scales keep activations O(1) through ~100 residual layers on 0..255 inputs in eval mode, and make every
    learned quantity (offsets, attention logits, biases, norms) non-trivial so parity checks exercise them, while
    keeping the decoder well conditioned: sampling offsets and centre refinements react mildly to the query state
    (as in a trained model), so a relative feature error is not amplified by re-sampling spatially white feature
    maps at shifted locations.  (With O(1) offset gains the same network turns a 0.1 % feature error into 5 %.)
    """
    out = {}
    for key, t in template.items():
        g = _gen(key, seed)
        shape = tuple(t.shape)
        if key.endswith("num_batches_tracked"):
            v = torch.zeros(shape, dtype=t.dtype)
        elif key.endswith("running_var"):
            v = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif key.endswith("running_mean"):
            v = torch.randn(shape, generator=g) * 0.1
        elif ".bn" in key or "downsample.1" in key or ".norm" in key:       # affine norm parameters
            if key.endswith("weight"):
                v = torch.rand(shape, generator=g) * 0.5 + 0.75
                if ".bn3." in key:
                    v = v * 0.25                                           # keep the residual branches modest
            else:
                v = torch.randn(shape, generator=g) * 0.1
        elif key.endswith("adjustment_layer.weight"):
            v = torch.randn(shape, generator=g) / (shape[1] ** 0.5)
        elif key.endswith("body.conv1.weight"):                            # stem sees 0..255 inputs
            fan_in = shape[1] * shape[2] * shape[3]
            v = torch.randn(shape, generator=g) / (fan_in ** 0.5 * 100.0)
        elif len(shape) == 4:                                              # conv weights
            fan_in = shape[1] * shape[2] * shape[3]
            if "inner_blocks.0.0" in key:                                  # FPN lateral on the raw 0..255 input
                v = torch.randn(shape, generator=g) / (fan_in ** 0.5 * 100.0)
            else:
                v = torch.randn(shape, generator=g) * (1.5 / fan_in) ** 0.5
        elif key.endswith("sampling_offsets.weight"):
            v = torch.randn(shape, generator=g) * 0.05                     # offsets move by a fraction of a pixel per unit query
        elif key.endswith("sampling_offsets.bias"):
            v = torch.randn(shape, generator=g) * 2.0
        elif key in ("fuser.query",):
            v = torch.rand(shape, generator=g)
        elif key.endswith("query_embedding.weight"):
            v = torch.randn(shape, generator=g)
        elif len(shape) == 2:                                              # linear weights
            v = torch.randn(shape, generator=g) / (shape[1] ** 0.5)
            if ".center_head." in key and shape[0] == 3:
                v = v * 0.25                                               # small additive refinements of the centres
        elif len(shape) == 1:                                              # linear / conv biases
            v = torch.randn(shape, generator=g) * 0.1
        else:
            v = torch.randn(shape, generator=g) * 0.1
        out[key] = v.to(t.dtype)
    return out
