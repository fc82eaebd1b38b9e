"""Programmatic equivalents of the reference's shipped model configs (reference config/*.json).

``make_config(name)`` returns a dict whose ``computing`` and ``model`` sections equal those of the
reference file of the same name (tests/test_reference_parity.py checks this against /root/reference when
it is present).  Any user-supplied DPFT JSON works too: ``DPRT.from_config`` reads the same keys
(``config['computing']``, ``config['model']``; reference src/dprt/models/dprt.py:114-133).
"""
from __future__ import annotations

from typing import Dict, List

VIEWS: Dict[str, List[str]] = {
    "kradar": ["camera_mono", "radar_bev", "radar_front"],
    "kradar_camera_mono": ["camera_mono"],
    "kradar_radar": ["radar_bev", "radar_front"],
    "kradar_radar_bev": ["radar_bev"],
    "kradar_radar_front": ["radar_front"],
}


def _backbone(view: str) -> dict:
    if view.startswith("camera"):
        return {"name": "ResNet101", "weights": "IMAGENET1K_V2", "multi_scale": 4, "norm_layer": "BatchNorm2d"}
    return {"name": "ResNet50", "weights": "IMAGENET1K_V2", "in_channels": 6, "multi_scale": 4,
            "norm_layer": "BatchNorm2d"}


def make_config(name: str = "kradar") -> dict:
    views = VIEWS[name]
    n = len(views)
    raw = {v: (3 if v.startswith("camera") else 6) for v in views}
    return {
        "dataset": "kradar",
        "computing": {"dtype": "float32", "seed": 42, "workers": 16, "device": "cuda"},
        "train": {"batch_size": 4, "optimizer": {"name": "AdamW", "lr": 0.0001}},
        "model": {
            "name": "dprt",
            "inputs": list(views),
            "skiplinks": {v: True for v in views},
            "backbones": {v: _backbone(v) for v in views},
            "necks": {v: {"name": "FPN", "in_channels_list": [raw[v], 256, 512, 1024, 2048], "out_channels": 16}
                      for v in views},
            "embeddings": {v: {"name": "sinusoidal_embedding", "num_feats": 16, "n_levels": 5, "normalize": True}
                           for v in views},
            "querent": {"name": "data_agnostic_static_querent", "transformation": "spher2cart",
                        "resolution": [20, 20, 1], "minimum": [4, -50, 0], "maximum": [72, 50, 0]},
            "fuser": {"name": "IMPFusion", "i_iter": 4, "m_views": n, "d_model": 16, "d_ffn": 32,
                      "n_queries": 400, "n_levels": [5] * n, "n_heads": [8] * n, "n_points": [4] * n,
                      "norm": True, "dropout": 0.1, "reduction": "linear", "activation": "Mish"},
            "head": {"name": "linear_detection_head", "in_channels": 16, "num_classes": 2,
                     "num_reg_layers": 3, "num_cls_layers": 3},
        },
    }
