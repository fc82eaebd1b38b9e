"""The MSDA oracle: C restatement vs the grid_sample form vs the independent HuggingFace restatement, plus
hand-computed known answers for the edge semantics."""
import pytest
import torch

from oracle import msda as O

SHAPE_SETS = [
    [(6, 4), (3, 2)],                                   # upstream's own test sizes
    [(256 // 8, 107 // 8), (8, 4), (2, 4), (1, 1)],
    [(37, 107), (10, 27), (5, 14), (3, 7), (2, 4)],     # radar_front pyramid
]


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-5)])
@pytest.mark.parametrize("shapes", SHAPE_SETS)
@pytest.mark.parametrize("D", [2, 3, 8])
def test_c_matches_grid_sample(dtype, tol, shapes, D):
    v, sh, loc, a, go = O.random_problem(2, 9, 3, D, shapes, 4, dtype=dtype, seed=D, spread=0.3)
    assert torch.allclose(O.msda_forward_c(v, sh, loc, a), O.msda_forward_torch(v, sh, loc, a), atol=tol, rtol=tol)
    for x, y in zip(O.msda_backward_c(v, sh, loc, a, go), O.msda_backward_torch(v, sh, loc, a, go)):
        assert torch.allclose(x, y, atol=tol * 50, rtol=tol * 50)


def test_matches_huggingface_restatement():
    hf = pytest.importorskip("transformers.models.deformable_detr.modeling_deformable_detr")
    v, sh, loc, a, _ = O.random_problem(2, 7, 4, 4, SHAPE_SETS[0], 3, dtype=torch.float64, seed=5, spread=0.4)
    want = hf.MultiScaleDeformableAttention()(v, None, sh, None, loc, a, 64)
    assert torch.allclose(O.msda_forward_c(v, sh, loc, a), want, atol=1e-12)
    assert torch.allclose(O.msda_forward_torch(v, sh, loc, a), want, atol=1e-12)


def _single(value_hw, x, y, dtype=torch.float64):
    """One sample, one head, one channel on an HxW map with attention weight 1."""
    H, W = value_hw.shape
    v = value_hw.reshape(1, H * W, 1, 1).to(dtype)
    loc = torch.tensor([x, y], dtype=dtype).view(1, 1, 1, 1, 1, 2)
    a = torch.ones(1, 1, 1, 1, 1, dtype=dtype)
    return float(O.msda_forward_c(v, [(H, W)], loc, a))


def test_known_answers():
    m = torch.tensor([[1.0, 2.0, 3.0, 4.0], [5.0, 6.0, 7.0, 8.0]])      # 2x4, the radar_front level '4' size
    # pixel centres: loc = (j + 0.5)/W, (i + 0.5)/H reproduces m[i, j]
    for i in range(2):
        for j in range(4):
            assert _single(m, (j + 0.5) / 4, (i + 0.5) / 2) == pytest.approx(float(m[i, j]))
    # half-way between four pixels
    assert _single(m, 1.0 / 4, 1.0 / 2) == pytest.approx((1 + 2 + 5 + 6) / 4)
    # loc = 0 sits half a pixel outside: only the corner pixel, weight 1/4 (zero padding)
    assert _single(m, 0.0, 0.0) == pytest.approx(0.25 * 1.0)
    assert _single(m, 1.0, 1.0) == pytest.approx(0.25 * 8.0)
    # more than one pixel outside contributes nothing
    assert _single(m, -0.2, 0.5) == 0.0 and _single(m, 0.5, 1.6) == 0.0 and _single(m, 1.3, 0.5) == 0.0
    # 1x1 level
    one = torch.tensor([[3.0]])
    assert _single(one, 0.5, 0.5) == pytest.approx(3.0)
    assert _single(one, 0.0, 0.5) == pytest.approx(1.5)


def test_zero_attention_and_linearity():
    v, sh, loc, a, _ = O.random_problem(1, 5, 2, 4, SHAPE_SETS[1], 2, dtype=torch.float64, seed=9)
    assert O.msda_forward_c(v, sh, loc, torch.zeros_like(a)).abs().max() == 0
    o1 = O.msda_forward_c(v, sh, loc, a)
    assert torch.allclose(O.msda_forward_c(2.5 * v, sh, loc, a), 2.5 * o1, atol=1e-12)


def test_empty_queries():
    v, sh, loc, a, _ = O.random_problem(1, 0, 2, 4, SHAPE_SETS[0], 2, dtype=torch.float32)
    assert O.msda_forward_c(v, sh, loc, a).shape == (1, 0, 8)
