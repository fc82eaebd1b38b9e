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


def edge_case_inputs():
    """Reference-point inputs that hit every branch of mpfusion.py:617-696 / transformations.py:71-120: r = 0 after the
    rigid transform (elevation forced to 0), points on the axes (azimuth +-90 / 180 degrees, elevation +-90), w = 0 and
    w < 0 in the perspective division, results outside [0, 1] (clipped), 3x4 and 4x4 projections, zero transformation."""
    t = torch.eye(4).repeat(2, 1, 1)
    t[0, :3, 3] = torch.tensor([0.5, -0.25, 0.125])
    t[1, :3, 3] = torch.tensor([-1.0, 2.0, 0.0])
    pts = torch.tensor([[-0.5, 0.25, -0.125], [9.5, 0.25, -0.125], [-10.5, 0.25, -0.125], [-0.5, 7.25, -0.125],
                        [-0.5, -6.75, -0.125], [-0.5, 0.25, 3.875], [-0.5, 0.25, -4.125], [30.0, 10.0, 2.0],
                        [70.0, -40.0, -1.0], [4.0, 4.0, 4.0], [1e-6, 1e-6, 1e-6], [118.0, 0.0, 0.0]])
    radar_pts = torch.stack((pts, pts - t[1, :3, 3] + t[0, :3, 3]))          # sample 1 hits r = 0 as well
    H, W = 256, 107
    p_bev = torch.tensor([[0.0, -1.0, 0.0, (W - 1) / 2.0], [H / 118.037, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]]).repeat(2, 1, 1)
    p_front = torch.tensor([[0.0, -1.0, 0.0, 53.0], [0.0, 0.0, 1.0, 18.0], [0.0, 0.0, 0.0, 1.0]]).repeat(2, 1, 1)
    cam_pts = torch.tensor([[10.0, 0.0, 0.0], [10.0, 3.0, -1.0], [0.0, 1.0, 1.0], [-5.0, 1.0, 1.0], [2.0, 30.0, 0.0],
                            [2.0, -30.0, 5.0], [50.0, -6.0, 2.0], [1e-3, 0.0, 0.0], [0.0, 0.0, 0.0], [72.0, 6.4, 6.0],
                            [4.0, -6.4, -2.0], [20.0, 0.5, 0.25]]).repeat(2, 1, 1)
    f, cx, cy = 700.0, 640.0, 360.0
    p_cam = torch.tensor([[cx, -f, 0.0, 0.0], [cy, 0.0, -f, 0.0], [1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]]).repeat(2, 1, 1)
    return [dict(name="radar_bev", query=radar_pts, t=t, p=p_bev, shape=torch.tensor([[H, W]] * 2)),
            dict(name="radar_front", query=radar_pts, t=t, p=p_front, shape=torch.tensor([[37, 107]] * 2)),
            dict(name="camera", query=cam_pts, t=torch.zeros(2, 4, 4), p=p_cam, shape=torch.tensor([[720, 1280]] * 2))]


def reference_point_edge_cases():
    reference_shim.import_reference_models()
    from dprt.models.fusers.mpfusion import IMPFusion
    cases = edge_case_inputs()
    for c in cases:
        c["out"] = IMPFusion.get_reference_points(None, c["query"].clone(), c["t"].clone(), c["p"].clone(), c["shape"].clone())
        print(c["name"], c["out"][0, :4].tolist())
    path = os.path.join(GOLDEN, "refpoints_edge_cases.pt")
    torch.save({"cases": cases, "torch_version": torch.__version__}, path)
    print("refpoints", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    if "--refpoints-only" not in sys.argv:
        main()
    reference_point_edge_cases()
