"""Golden vectors produced by the real reference (tools/make_golden.py) pin the CPU oracle and, with the oracle
op injected for the one CUDA-only piece, the product's host-side model logic — all on the CPU."""
import pytest
import torch

from conftest import load_golden
from helpers import case_setup, oracle_op_injected, rel_err
from dpft_b200 import models, synthetic
from oracle import dprt_oracle

CASES = ["radar_bev_native", "radar_bev_256", "radar_front_native", "camera_mono_small", "fusion_small_300q",
         "fusion_native_1"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_model_matches_reference_golden(name):
    rec = load_golden(name)
    cfg, batch = case_setup(rec)
    template = models.build("dprt", cfg).state_dict()
    sd = synthetic.seeded_state_dict(template, seed=rec["weight_seed"])
    with torch.no_grad():
        out = dprt_oracle.forward(sd, cfg, batch)
    assert list(out.keys()) == ["center", "size", "angle", "class"]
    for k, want in rec["outputs"].items():
        assert rel_err(out[k], want) < 5e-4, k  # fp32 end-to-end; north_star bar is 1e-3 rel


@pytest.mark.parametrize("name", ["radar_bev_native", "fusion_small_300q"])
def test_product_host_logic_matches_reference_golden(name):
    rec = load_golden(name)
    cfg, batch = case_setup(rec)
    model = models.build("dprt", cfg).eval()
    model.load_state_dict(synthetic.seeded_state_dict(model.state_dict(), seed=rec["weight_seed"]), strict=True)
    with oracle_op_injected(), torch.no_grad():
        out = model(batch)
    assert list(out.keys()) == ["center", "size", "angle", "class"]
    for k, want in rec["outputs"].items():
        assert rel_err(out[k], want) < 5e-4, k  # fp32 end-to-end; north_star bar is 1e-3 rel


def test_msdeformattn_module_matches_reference_golden():
    from dpft_b200.models.fuser import MSDeformAttn
    rec = torch.load(__import__("os").path.join(__import__("conftest").GOLDEN, "msdeformattn_module.pt"),
                     weights_only=False)
    mod = MSDeformAttn(d_model=16, n_levels=3, n_heads=8, n_points=4).eval()
    mod.load_state_dict(rec["state_dict"], strict=True)
    sh = torch.tensor(rec["shapes"])
    lsi = torch.tensor([0, 108, 138])
    with oracle_op_injected(), torch.no_grad():
        out = mod(rec["query"], rec["ref"], rec["flat"], sh, lsi)
    assert rel_err(out, rec["out"]) < 1e-4
    # the oracle's own module-level restatement
    sd = {"x." + k: v for k, v in rec["state_dict"].items()}
    got = dprt_oracle.deformable_attention(sd, "x", rec["query"], rec["ref"], rec["flat"], rec["shapes"], 8, 4)
    assert rel_err(got, rec["out"]) < 1e-4


def test_default_initialisation_of_msdeformattn_follows_the_reference_scheme():
    from dpft_b200.models.fuser import MSDeformAttn
    m = MSDeformAttn(16, 5, 8, 4)
    assert m.sampling_offsets.weight.abs().max() == 0 and m.attention_weights.weight.abs().max() == 0
    b = m.sampling_offsets.bias.view(8, 5, 4, 2)
    assert torch.allclose(b[0, 0, :, 0], torch.tensor([1.0, 2.0, 3.0, 4.0])) and b[0, :, :, 1].abs().max() < 1e-6
    assert torch.allclose(b[2, 3, 1], torch.tensor([0.0, 2.0]), atol=1e-6)
