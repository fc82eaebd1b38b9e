"""BASELINE config 5's structure (900 queries, 4 feature levels = raw + 3 stages, d_model 64: 8 heads of 8 channels;
SURVEY.md §8d) on the GPU against the golden outputs of the unmodified reference (tools/make_golden.py).  This width is
outside the fused engine's family, so the forward runs module by module: dense layers through torch (TF32 off),
deformable attention through ``dpft_msda_forward`` (gather-then-project form, D = 64 rows) and the decoder
self-attention through the tcgen05 flash-attention kernel (head dim 8)."""
import pytest
import torch

from conftest import load_golden
from helpers import case_setup, rel_err
from dpft_b200 import models, native, synthetic

pytestmark = pytest.mark.gpu

TOL_FP32 = 1e-3      # north_star: outputs within 1e-3 rel (fp32) of the reference forward


def test_stress_structure_forward_matches_reference_golden():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        rec = load_golden("stress_4level_d64_900q")
        cfg, batch = case_setup(rec)
        assert cfg["model"]["fuser"]["d_model"] == 64 and cfg["model"]["fuser"]["n_levels"] == [4, 4, 4]
        model = models.build("dprt", cfg).eval()
        model.load_state_dict(synthetic.seeded_state_dict(model.state_dict(), seed=rec["weight_seed"]), strict=True)
        model = model.to("cuda:0")
        l0 = native.launches()
        with torch.no_grad():
            out = model({k: v.to("cuda:0") for k, v in batch.items()})
        assert model._engine is None, "d_model 64 is outside the fused engine's family: the composed path must run"
        # 4 iterations x 3 views x (flash attention + deformable-attention op)
        assert native.launches() - l0 >= 24, "the composed path did not launch the native attention / sampling kernels"
        assert list(out.keys()) == ["center", "size", "angle", "class"]
        for k, want in rec["outputs"].items():
            assert out[k].shape == want.shape
            assert rel_err(out[k].cpu(), want) < TOL_FP32, (k, rel_err(out[k].cpu(), want))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
