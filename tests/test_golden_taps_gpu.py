"""The reference's intermediate quantities and gradient digests (tests/golden/taps_fusion_small_300q.pt,
grads_radar_small.pt; tools/make_golden_taps.py) against the GPU paths: module-by-module forward (torch dense layers with TF32
off, dpft_msda_forward, tcgen05 flash-attention) and the training step through dpft_msda_forward / dpft_msda_backward.
"""
import os

import pytest
import torch

import model_taps
from conftest import load_golden
from helpers import case_setup, rel_err
from dpft_b200 import configs, models, synthetic

pytestmark = [pytest.mark.gpu]
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _strict_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_composed_gpu_path_intermediates_match_reference_taps():
    from test_golden_taps import _check_taps
    rec = load_golden("taps_fusion_small_300q")
    cfg, batch = case_setup(rec)
    model = models.build("dprt", cfg).eval()
    model.load_state_dict(synthetic.seeded_state_dict(model.state_dict(), seed=rec["weight_seed"]), strict=True)
    model = model.to(DEV)
    model.use_fused = False
    f = cfg["model"]["fuser"]
    out, taps = model_taps.collect(model, {k: v.to(DEV) for k, v in batch.items()}, f["i_iter"], f["m_views"])
    _check_taps(taps, rec["taps"], 1e-3)                                   # north_star: 1e-3 rel fp32
    for k, w in load_golden("fusion_small_300q")["outputs"].items():
        assert rel_err(out[k].cpu(), w) < 1e-3, k


def _digest_run(rec, native_train, tf32):
    cfg = synthetic.offline_config(configs.make_config(rec["config"]), dropout=rec["dropout"])
    model = models.build("dprt", cfg).train()
    model.load_state_dict(synthetic.seeded_state_dict(model.state_dict(), seed=rec["weight_seed"]))
    model = model.to(DEV)
    model.native_train = native_train          # False: torch fp32 dense layers + the native deformable-attention fwd/bwd
    batch = synthetic.synthetic_batch(cfg, rec["batch"], seed=rec["input_seed"], sizes=rec["sizes"], device=DEV)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = tf32
    try:
        loss = sum((v ** 2).mean() for v in model(batch).values())
        loss.backward()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    details = []
    worst_norm, worst_val = model_taps.digest_errors({k: p.grad for k, p in model.named_parameters()}, rec["grads"], details)
    details.sort(key=lambda r: -max(r[1], r[2]))
    return float(loss.detach()), worst_norm, worst_val, details


def test_gpu_training_gradients_match_reference_digest():
    """Gradient digests of the unmodified reference's train-mode step (CPU, fp32) against three GPU runs of the product:
    torch fp32 dense layers (TF32 off) + the native deformable-attention backward — held to the digest directly; the same with
    PyTorch's default TF32 convolutions — the yardstick (what the reference's own GPU training computes; 10-bit operands like
    float16); and the native 16-bit backbone kernels.  Train-mode BatchNorm over the few hundred samples of these small maps
    amplifies any rounding (tests/test_train_backbone_gpu.py measures 3-5x per stage), so the native path is bounded relative
    to the yardstick, as there: norms within 4x of TF32's own error, single entries within 4x, with floors."""
    rec = load_golden("grads_radar_small")
    loss32, norm32, val32, _ = _digest_run(rec, False, False)
    assert abs(loss32 - rec["loss"]) < 1e-3 * abs(rec["loss"])
    assert norm32 < 2e-2 and val32 < 1e-1, (norm32, val32)
    loss_tf, norm_tf, val_tf, _ = _digest_run(rec, False, True)
    loss_n, norm_n, val_n, details = _digest_run(rec, True, False)
    report = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(report):                  # per-parameter errors, worst first (evidence under profiles/ once promoted)
        import json
        with open(os.path.join(report, "grad_digest_native.json"), "w") as f:
            json.dump({"want_loss": rec["loss"], "fp32": [loss32, norm32, val32], "tf32_yardstick": [loss_tf, norm_tf, val_tf],
                       "native_f16": [loss_n, norm_n, val_n],
                       "native rows(name, norm_err, entry_err, rms, max_sampled)": details[:40]}, f, indent=1)
    assert abs(loss_n - rec["loss"]) < 2e-2 * abs(rec["loss"])
    assert norm_n <= max(4.0 * norm_tf, 2e-2), (norm_n, norm_tf)
    assert val_n <= max(4.0 * val_tf, 1e-1), (val_n, val_tf)


@pytest.mark.parametrize("path", ["composed_fp32", "fused_decoder_fp32", "native_f16"])
def test_two_view_radar_model_matches_reference_golden(path):
    """config/kradar_radar.json (V = 2: range-azimuth + elevation-azimuth) through the module-by-module path, the fused decoder
    on fp32 features and the native 16-bit pipeline — the shipped two-view configuration had no GPU golden case."""
    fused, native_feats, tol = {"composed_fp32": (False, False, 1e-3), "fused_decoder_fp32": (True, False, 1e-3),
                                "native_f16": (True, True, 1e-2)}[path]
    rec = load_golden("radar_two_views")
    cfg, batch = case_setup(rec)
    model = models.build("dprt", cfg).eval()
    model.load_state_dict(synthetic.seeded_state_dict(model.state_dict(), seed=rec["weight_seed"]), strict=True)
    model = model.to(DEV)
    model.use_fused, model.native_features = fused, native_feats
    with torch.no_grad():
        out = model({k: v.to(DEV) for k, v in batch.items()})
    if fused:
        assert model._engine is not None
    for k, want in rec["outputs"].items():
        assert rel_err(out[k].cpu(), want) < tol, (path, k, rel_err(out[k].cpu(), want))


def test_reference_point_edge_cases_through_the_fused_decoder():
    """The reference-point projection of the FUSED decoder kernel (decoder_layer_kernel: rigid transform, cart2spher, 3x4 / 4x4
    projection with perspective division, normalisation by the stored shape, clipping) on the edge points of
    tests/golden/refpoints_edge_cases.pt — r = 0, points on the axes, w = 0 and w < 0, far outside the field of view — with
    that fixture's calibrations.  The host logic of the module-by-module path is held to the reference's values on exactly
    these cases on the CPU (tests/test_golden_taps.py); here the fused kernel must agree with that path on the GPU, through all
    four decoder iterations (the refined centres of later iterations start from these points)."""
    rec = load_golden("refpoints_edge_cases")
    cases = {c["name"]: c for c in rec["cases"]}
    pts = torch.cat((cases["radar_bev"]["query"][0], cases["camera"]["query"][0]), dim=0)          # 24 query centres
    cfg = synthetic.offline_config(configs.make_config("kradar"), n_queries=(24, 1, 1))
    model = models.build("dprt", cfg).eval()
    model.load_state_dict(synthetic.seeded_state_dict(model.state_dict(), seed=21))
    model = model.to(DEV)
    model.querent.grid = lambda dtype, device: pts.to(device=device, dtype=dtype)
    sizes = {"camera_mono": (96, 160, 3), "radar_bev": (64, 48, 6), "radar_front": (37, 48, 6)}
    batch = synthetic.synthetic_batch(cfg, 2, seed=22, sizes=sizes)
    for view, name in (("camera_mono", "camera"), ("radar_bev", "radar_bev"), ("radar_front", "radar_front")):
        c = cases[name]
        batch[f"label_to_{view}_t"] = c["t"].clone()
        batch[f"label_to_{view}_p"] = c["p"].clone()
        batch[f"{view}_shape"] = torch.cat((c["shape"], torch.full((2, 1), sizes[view][2])), dim=1).to(torch.int64)
    gb = {k: v.to(DEV) for k, v in batch.items()}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            model.use_fused, model.native_features = False, False
            want = model(gb)                               # module by module: IMPFusion.get_reference_points in torch
            model.use_fused = True
            got = model(gb)                                # fused decoder on the same fp32 features
            assert model._engine is not None
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    for k in want:
        assert torch.isfinite(got[k]).all(), k
        assert rel_err(got[k].cpu(), want[k].cpu()) < 1e-3, (k, rel_err(got[k].cpu(), want[k].cpu()))
