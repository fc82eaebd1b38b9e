"""Golden vectors produced by the real reference (tools/make_golden.py) pin the CPU oracle and, with the oracle
op injected for the one CUDA-only piece, the product's host-side model logic — all on the CPU."""
import pytest
import torch

from conftest import load_golden
from helpers import case_setup, oracle_op_injected, rel_err
from dpft_b200 import models, synthetic
from oracle import dprt_oracle

CASES = ["radar_bev_native", "radar_bev_256", "radar_front_native", "camera_mono_small", "fusion_small_300q",
         "fusion_native_1", "stress_4level_d64_900q", "radar_two_views"]


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


@pytest.mark.parametrize("name", ["radar_bev_native", "fusion_small_300q", "stress_4level_d64_900q", "radar_two_views"])
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


def test_self_attention_sublayer_matches_reference_golden():
    """The decoder's self-attention sub-layer (reference MLFusion.forward_self_attn, mpfusion.py:122-148; fixture from
    tools/make_golden_selfattn.py): the oracle restatement and the product's host logic (torch path on the CPU) against the
    outputs of the unmodified reference."""
    import torch.nn.functional as F
    from oracle import dprt_oracle
    from dpft_b200.models.fuser import MLFusion
    rec = load_golden("mlfusion_self_attn")
    for case in rec["cases"]:
        d, h = case["d_model"], case["n_heads"]
        sd = {"l." + k: v for k, v in case["state_dict"].items()}
        with torch.no_grad():
            got = case["x"] + dprt_oracle.self_attention(sd, "l.self_attn", case["x"], case["pos"], h)
            got = F.layer_norm(got, (d,), sd["l.norm1.weight"], sd["l.norm1.bias"], 1e-5)
        assert float((got - case["out"]).abs().max()) < 2e-6
        layer = MLFusion(d_model=d, d_ffn=2 * d, n_levels=1, n_heads=h, n_points=4, norm=True, dropout=0.1, activation="Mish").eval()
        layer.load_state_dict(case["state_dict"], strict=False)
        with torch.no_grad():
            mine = layer.forward_self_attn(case["x"], case["pos"])
        assert float((mine - case["out"]).abs().max()) < 2e-6
