"""Generates tests/golden/taps_fusion_small_300q.pt and tests/golden/grads_radar_small.pt from the UNMODIFIED reference
(imported from /root/reference; build container only):

  * the intermediate quantities SURVEY.md §8c lists (per-level FPN + embedding features, reference points per view and
    iteration, fused queries, per-iteration centres, MSDeformAttn outputs) of the `fusion_small_300q` golden case
    (same config / weight seed / input seed as tests/golden/fusion_small_300q.pt), collected by tools/model_taps.py;
  * a per-parameter digest (norm, sum, 16 sampled entries) of the gradients of the fixed scalar loss
    sum_k mean(out_k^2) in train() mode with dropout 0 on the radar-only config, plus the BatchNorm running statistics
    after that step.

  python tools/make_golden_taps.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))

import model_taps  # noqa: E402
import reference_shim  # noqa: E402
from dpft_b200 import configs, synthetic  # noqa: E402

GOLDEN = os.path.join(HERE, "..", "tests", "golden")
GRAD_SIZES = {"radar_bev": (64, 40, 6), "radar_front": (37, 40, 6)}
GRAD_SEEDS = (31, 32)


def main():
    ref = reference_shim.import_reference_models()
    base = torch.load(os.path.join(GOLDEN, "fusion_small_300q.pt"), weights_only=False)
    case = base["case"]
    cfg = synthetic.offline_config(configs.make_config(case["config"]), n_queries=case["n_queries"])
    model = ref.build("dprt", cfg).eval()
    model.load_state_dict(synthetic.seeded_state_dict(model.state_dict(), seed=base["weight_seed"]), strict=True)
    batch = synthetic.synthetic_batch(cfg, case["batch"], seed=base["input_seed"], sizes=case["sizes"])
    f = cfg["model"]["fuser"]
    out, taps = model_taps.collect(model, {k: v.clone() for k, v in batch.items()}, f["i_iter"], f["m_views"])
    for k, v in base["outputs"].items():                    # the tapped run is the run of the existing golden
        assert torch.equal(out[k], v), k
    rec = {"case": case, "weight_seed": base["weight_seed"], "input_seed": base["input_seed"], "torch_version": torch.__version__,
           "taps": {k: (model_taps.compress_features(v) if k.startswith("features_") else v) for k, v in taps.items()
                    if not k.startswith(("msda_out_1", "msda_out_2"))}}        # MSDeformAttn outputs: first and last iteration
    path = os.path.join(GOLDEN, "taps_fusion_small_300q.pt")
    torch.save(rec, path)
    print("taps", sorted(taps), os.path.getsize(path), "bytes")

    cfg = synthetic.offline_config(configs.make_config("kradar_radar"), dropout=0.0)
    model = ref.build("dprt", cfg).train()
    model.load_state_dict(synthetic.seeded_state_dict(model.state_dict(), seed=GRAD_SEEDS[0]))
    batch = synthetic.synthetic_batch(cfg, 2, seed=GRAD_SEEDS[1], sizes=GRAD_SIZES)
    loss = sum((v ** 2).mean() for v in model({k: v.clone() for k, v in batch.items()}).values())
    loss.backward()
    grads = model_taps.gradient_digest({k: p.grad for k, p in model.named_parameters()})
    running = {k: (float(v.double().norm()), float(v.double().sum())) for k, v in model.state_dict().items() if "running_" in k}
    rec = {"config": "kradar_radar", "dropout": 0.0, "batch": 2, "sizes": GRAD_SIZES, "weight_seed": GRAD_SEEDS[0],
           "input_seed": GRAD_SEEDS[1], "loss": float(loss.detach()), "grads": grads, "running": running,
           "torch_version": torch.__version__}
    path = os.path.join(GOLDEN, "grads_radar_small.pt")
    torch.save(rec, path)
    print("grads", len(grads["names"]), "parameters with,", len(grads["none"]), "without gradient;", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
