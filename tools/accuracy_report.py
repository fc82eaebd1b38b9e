"""Accuracy of the native 16-bit feature path against the reference golden vectors and the fp32 torch features:
per-output relative error (max-norm) for each golden case and activation dtype, plus per-stage / per-level feature errors."""
import json
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden  # noqa: E402
from helpers import case_setup, rel_err  # noqa: E402
from dpft_b200 import models, synthetic  # noqa: E402
from dpft_b200.models.fuser import FeaturePyramid  # noqa: E402

CASES = ["radar_bev_native", "radar_bev_256", "radar_front_native", "camera_mono_small", "fusion_small_300q", "fusion_native_1"]
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"

for name in CASES:
    rec = load_golden(name)
    cfg, batch = case_setup(rec)
    model = models.build("dprt", cfg).eval()
    model.load_state_dict(synthetic.seeded_state_dict(model.state_dict(), seed=rec["weight_seed"]))
    model = model.to(dev)
    gb = {k: v.to(dev) for k, v in batch.items()}
    row = {"case": name}
    with torch.no_grad():
        for label, native_feats, dt in (("fp32_fused", False, torch.bfloat16), ("bf16", True, torch.bfloat16),
                                        ("f16", True, torch.float16)):
            model.native_features, model.feature_dtype = native_feats, dt
            out = model(gb)
            row[label] = {k: round(rel_err(out[k].cpu(), w), 6) for k, w in rec["outputs"].items()}
        # feature-level errors
        feats = model.extract_features(gb)
        for dt, label in ((torch.bfloat16, "bf16"), (torch.float16, "f16")):
            model.feature_dtype = dt
            model(gb)
            eng = model._engine
            for vname, nv in zip(model.inputs, eng.views):
                want_levels = list(feats[vname].values())
                flat, shapes = nv.pyramid(gb[vname])
                off = 0
                errs = []
                for (h, w), wl in zip(shapes, want_levels):
                    got = flat[:, off:off + h * w].reshape(wl.shape)
                    errs.append(round(rel_err(got.cpu(), wl.cpu()), 5))
                    off += h * w
                row[f"pyr_{label}_{vname}"] = errs
                bb = model.backbones[vname](gb[vname])
                stage = [round(rel_err(g.float().cpu(), w.cpu()), 5) for g, w in zip(nv.backbone(gb[vname]), bb.values())]
                row[f"stages_{label}_{vname}"] = stage
    print(json.dumps(row), flush=True)
