"""tests/golden/refinit_*.pt: the UNMODIFIED reference with ITS OWN initialisation (``torch.manual_seed(42)`` as in
config/kradar.json:6, then the reference constructor — no dpft_b200.synthetic.seeded_state_dict conditioning) on seeded
synthetic inputs.  The fixture holds the outputs, the largest activation of every backbone stage (the range a 16-bit
pipeline has to hold) and a digest of the weights; the weights themselves (100-400 MB) are rebuilt where the test runs by
the same constructor call through the installed copy of the reference (baseline/_ref) and checked against the digest.

Run in the build container only:  python tools/make_golden_refinit.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))

import reference_shim  # noqa: E402
from dpft_b200 import configs, synthetic  # noqa: E402

GOLDEN = os.path.join(HERE, "..", "tests", "golden")
SEED = 42
CASES = [
    dict(name="refinit_radar_bev", config="kradar_radar_bev", batch=2, sizes={"radar_bev": (128, 107, 6)}, n_queries=None),
    dict(name="refinit_fusion_small", config="kradar", batch=1,
         sizes={"camera_mono": (96, 160, 3), "radar_bev": (64, 48, 6), "radar_front": (37, 48, 6)}, n_queries=(20, 15, 1)),
]


def weight_digest(sd):
    """Per-tensor (sum, |.|-sum) in float64 of every parameter / buffer: cheap, order-independent, catches any RNG drift."""
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in sd.items()}


def build_reference_init(ref_models, cfg):
    torch.manual_seed(SEED)
    return ref_models.build("dprt", cfg).eval()


def main():
    ref = reference_shim.import_reference_models()
    for i, case in enumerate(CASES):
        cfg = synthetic.offline_config(configs.make_config(case["config"]), n_queries=case["n_queries"])
        model = build_reference_init(ref, cfg)
        batch = synthetic.synthetic_batch(cfg, case["batch"], seed=700 + i, sizes=case["sizes"])
        ranges = {}
        hooks = []
        for view, bb in model.backbones.items():
            for lname in ("conv1", "layer1", "layer2", "layer3", "layer4"):
                mod = getattr(bb.body, lname, None)
                if mod is not None:
                    hooks.append(mod.register_forward_hook(
                        lambda m, a, out, key=f"{view}.{lname}": ranges.__setitem__(key, float(out.abs().max()))))
        with torch.no_grad():
            out = model({k: v.clone() for k, v in batch.items()})
        for h in hooks:
            h.remove()
        rec = dict(case=case, init_seed=SEED, input_seed=700 + i, torch_version=torch.__version__,
                   outputs={k: v.clone() for k, v in out.items()}, activation_max=ranges,
                   weight_digest=weight_digest(model.state_dict()))
        torch.save(rec, os.path.join(GOLDEN, case["name"] + ".pt"))
        print(case["name"], {k: float(v.abs().max()) for k, v in out.items()})
        print("   activation max per stage:", {k: f"{v:.3g}" for k, v in ranges.items()})


if __name__ == "__main__":
    main()
