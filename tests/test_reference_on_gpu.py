"""The UNMODIFIED reference package on the GPU box (the copy __graft_entry__.build() installs into baseline/_ref):

* it runs on cuda:0 with THIS repo's deformable-attention kernels behind its ``import MultiScaleDeformableAttention``
  (the native-op plugin boundary) and agrees with the product model on the same weights and inputs;
* golden cases with the reference's OWN initialisation (torch.manual_seed(42) + its constructor, config/kradar.json:6 —
  no seeded_state_dict conditioning): the weights are rebuilt here by the reference constructor, checked against the
  committed digest, and the GPU paths are held to the committed outputs of the reference's CPU forward.
"""
import json
import os

import pytest
import torch

import reference_shim
from conftest import load_golden
from helpers import rel_err
from dpft_b200 import configs, models, native, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref_models():
    if not reference_shim.available():
        pytest.fail("baseline/_ref is missing: run __graft_entry__.build() in the build container (it travels with gpurun)")
    return reference_shim.import_reference_models()


@pytest.fixture(autouse=True)
def _strict_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _report(name, rows):
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, name), "w") as f:
            json.dump(rows, f, indent=1)


def test_unmodified_reference_runs_on_these_kernels_and_matches_the_product(ref_models):
    cfg = synthetic.offline_config(configs.make_config("kradar"), n_queries=(20, 15, 1))
    sizes = {"camera_mono": (96, 160, 3), "radar_bev": (64, 48, 6), "radar_front": (37, 48, 6)}
    ours = models.build("dprt", cfg).eval()
    sd = synthetic.seeded_state_dict(ours.state_dict(), seed=11)
    ours.load_state_dict(sd)
    ours = ours.to(DEV)
    ref = ref_models.build("dprt", cfg).eval()
    ref.load_state_dict(sd, strict=True)                       # the reference's 1652 state-dict names
    ref = ref.to(DEV)
    batch = {k: v.to(DEV) for k, v in synthetic.synthetic_batch(cfg, 2, seed=12, sizes=sizes).items()}
    with torch.no_grad():
        l0 = native.launches()
        want = ref(batch)
        used = native.launches() - l0
        assert used >= 12, f"the reference forward launched {used} kernels of libdpft_b200.so (4 iterations x 3 views expected)"
        rows = {}
        for path, (fused, feats, tol) in {"composed_fp32": (False, False, 1e-3), "fused_decoder_fp32": (True, False, 1e-3),
                                          "native_f16": (True, True, 1e-2)}.items():
            ours.use_fused, ours.native_features = fused, feats
            got = ours(batch)
            rows[path] = {k: rel_err(got[k].cpu(), want[k].cpu()) for k in want}
            for k in want:
                assert rows[path][k] < tol, (path, k, rows[path][k])
    _report("reference_on_gpu_vs_product.json", rows)


def _reference_init_model(ref_models, rec):
    case = rec["case"]
    cfg = synthetic.offline_config(configs.make_config(case["config"]), n_queries=case["n_queries"])
    torch.manual_seed(rec["init_seed"])
    ref = ref_models.build("dprt", cfg).eval()
    sd = ref.state_dict()
    for k, (s, a) in rec["weight_digest"].items():             # the constructor reproduced the fixture's weights on this machine
        v = sd[k].double()
        assert abs(float(v.sum()) - s) <= 1e-9 * max(1.0, a) and abs(float(v.abs().sum()) - a) <= 1e-9 * max(1.0, a), k
    ours = models.build("dprt", cfg).eval()
    ours.load_state_dict(sd, strict=True)
    batch = synthetic.synthetic_batch(cfg, case["batch"], seed=rec["input_seed"], sizes=case["sizes"])
    return ours.to(DEV), {k: v.to(DEV) for k, v in batch.items()}


@pytest.mark.parametrize("path", ["composed_fp32", "fused_decoder_fp32", "native_f16", "native_bf16"])
def test_reference_initialisation_radar_bev(ref_models, path):
    """BASELINE config 1's model (kradar_radar_bev, ResNet-50) as the reference initialises it.  Untrained BatchNorm statistics
    are the identity, so activations grow through the residual stages (to 1.2e4 in layer4 here: inside the f16 range)."""
    rec = load_golden("refinit_radar_bev")
    model, batch = _reference_init_model(ref_models, rec)
    fused, feats, dtype, tol = {"composed_fp32": (False, False, None, 1e-3), "fused_decoder_fp32": (True, False, None, 1e-3),
                                "native_f16": (True, True, torch.float16, 1e-2),
                                "native_bf16": (True, True, torch.bfloat16, None)}[path]
    model.use_fused, model.native_features = fused, feats
    if dtype is not None:
        model.feature_dtype = dtype
    with torch.no_grad():
        out = model(batch)
    errs = {k: rel_err(out[k].cpu(), w) for k, w in rec["outputs"].items()}
    _report(f"refinit_radar_bev_{path}.json", errs)
    if tol is None:                  # bf16 (7-bit mantissa) is measured and reported, not held to north_star's 1e-2 bar
        assert all(e == e and e < 0.5 for e in errs.values()), errs
        return
    for k, e in errs.items():
        assert e < tol, (path, k, e)


@pytest.mark.parametrize("path", ["composed_fp32", "fused_decoder_fp32", "native_bf16", "native_f16"])
def test_reference_initialisation_full_fusion(ref_models, path):
    """config/kradar.json (camera ResNet-101 + two radar ResNet-50) as the reference initialises it.  With identity BatchNorm
    statistics the camera trunk reaches 1.2e7 in layer3 (fixture `activation_max`), 190x beyond the largest finite f16 value:
    the fp32 paths meet 1e-3; in f16 the epilogue saturates at 65504 by design (cvt.satfinite) and in bf16 the range is held but
    8 mantissa bits are not enough for this un-normalised depth: both native runs give finite outputs that are NOT the
    reference's (measured errors are written to gpurun_out/ and quoted in DESIGN §2) — the 16-bit pipeline is for
    BatchNorm-normalised (trained) networks; the ResNet-50 radar case above is inside its range and meets 1e-2 in f16."""
    rec = load_golden("refinit_fusion_small")
    assert max(v for k, v in rec["activation_max"].items() if k.startswith("camera_mono")) > 65504.0
    model, batch = _reference_init_model(ref_models, rec)
    fused, feats, dtype, tol = {"composed_fp32": (False, False, None, 1e-3), "fused_decoder_fp32": (True, False, None, 1e-3),
                                "native_bf16": (True, True, torch.bfloat16, None),
                                "native_f16": (True, True, torch.float16, None)}[path]
    model.use_fused, model.native_features = fused, feats
    if dtype is not None:
        model.feature_dtype = dtype
    with torch.no_grad():
        out = model(batch)
    errs = {k: rel_err(out[k].cpu(), w) for k, w in rec["outputs"].items()}
    _report(f"refinit_fusion_small_{path}.json", errs)
    assert all(torch.isfinite(v).all() for v in out.values()), "non-finite outputs"
    if tol is not None:
        for k, e in errs.items():
            assert e < tol, (path, k, e)
    # native_bf16 / native_f16 on this untrained 101-layer camera trunk: measured 0.2-0.6 on size / angle / class (f16 saturates,
    # bf16's 8-bit mantissa is amplified by the un-normalised depth); finite, reported in gpurun_out/, NOT a parity claim
