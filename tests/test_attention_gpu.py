"""tcgen05 flash-attention kernel (dpft_self_attention_forward, csrc/attention.cu) against the attention core of the
reference's decoder layer — nn.MultiheadAttention as called at src/dprt/models/fusers/mpfusion.py:139 — computed by torch in
fp64.  Floating point: tolerances are stated per mode (relative to the largest |output|):
  precise (fp32 in, split-f16 operands, 3 MMAs per product)  2e-5   (the fp32 module path is held to 1e-3 end to end)
  fast fp32 / float16 (single f16 operands)                  1e-2   (north_star: 1e-2 rel for the 16-bit tier)
  bfloat16                                                   4e-2   (8-bit mantissa: outside the 1e-2 tier, like the bf16 backbone)
"""
import math

import pytest
import torch

from dpft_b200 import attention

DEV = "cuda:0"

SHAPES = [(8, 8, 300, 2), (2, 8, 400, 2), (1, 4, 128, 16), (2, 8, 900, 32), (1, 2, 129, 64), (1, 1, 1, 8), (3, 8, 257, 8),
          (1, 3, 127, 24), (1, 2, 640, 40)]


def _ref(q, k, v, H):
    B, N, C = q.shape
    D = C // H
    qd, kd, vd = (t.double().view(B, N, H, D).transpose(1, 2) for t in (q, k, v))
    p = torch.softmax(qd @ kd.transpose(-1, -2) / math.sqrt(D), -1)
    return (p @ vd).transpose(1, 2).reshape(B, N, C)


def _err(got, want):
    return float((got.double() - want).abs().max() / want.abs().max().clamp_min(1e-12))


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,N,D", SHAPES)
@pytest.mark.parametrize("mode", ["precise", "fast32", "f16", "bf16"])
def test_attention_matches_fp64_reference(B, H, N, D, mode):
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + N + D)
    dtype = {"precise": torch.float32, "fast32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[mode]
    tol = {"precise": 2e-5, "fast32": 1e-2, "f16": 1e-2, "bf16": 4e-2}[mode]
    # scores with a spread of several units so the softmax is far from uniform (peaked rows exercise the running maximum)
    q, k, v = (torch.randn(B, N, H * D, generator=g, device=DEV) * s for s in (2.0, 2.0, 1.0))
    q, k, v = q.to(dtype), k.to(dtype), v.to(dtype)
    out = attention.self_attention(q, k, v, H, precise=(mode == "precise"))
    assert out.shape == (B, N, H * D) and out.dtype == dtype and out.is_contiguous()
    e = _err(out, _ref(q, k, v, H))
    print(f"{mode} B{B} H{H} N{N} D{D}: rel err {e:.2e}")
    assert e < tol, e


@pytest.mark.gpu
def test_attention_takes_strided_projection_slices():
    """q and k as the two halves of one packed (B, N, 2C) projection, v from a wider buffer: no copies needed."""
    g = torch.Generator(device=DEV).manual_seed(7)
    B, N, H, D = 2, 300, 8, 2
    C = H * D
    qk = torch.randn(B, N, 2 * C, generator=g, device=DEV)
    vbuf = torch.randn(B, N + 3, C + 5, generator=g, device=DEV)
    q, k, v = qk[..., :C], qk[..., C:], vbuf[:, 1:N + 1, :C]
    out = attention.self_attention(q, k, v, H)
    assert _err(out, _ref(q.contiguous(), k.contiguous(), v.contiguous(), H)) < 2e-5


@pytest.mark.gpu
def test_large_scores_and_constant_rows_stay_finite():
    """Scores of +-60 (softmax one-hot) and an all-equal row (uniform softmax): the running-max rescaling must not overflow."""
    B, N, H, D = 1, 260, 2, 4
    q = torch.zeros(B, N, H * D, device=DEV)
    k = torch.zeros(B, N, H * D, device=DEV)
    q[:, :, 0] = 30.0
    k[:, :, 0] = torch.linspace(-4, 4, N, device=DEV)
    v = torch.randn(B, N, H * D, device=DEV)
    out = attention.self_attention(q, k, v, H)
    assert torch.isfinite(out).all()
    assert _err(out, _ref(q, k, v, H)) < 2e-5
    # head 1 has all-zero q/k: uniform attention = mean of v over the keys
    assert torch.allclose(out[0, :, D:], v[0, :, D:].mean(0, keepdim=True).expand(N, -1), atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("C,H,N", [(16, 8, 300), (64, 8, 400), (256, 8, 900)])
def test_multihead_wrapper_equals_nn_multihead_attention(C, H, N):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1)
    mha = torch.nn.MultiheadAttention(C, H, dropout=0.1, batch_first=True).to(DEV).eval()
    with torch.no_grad():
        mha.in_proj_bias.normal_(0, 0.5)
    x = torch.randn(2, N, C, device=DEV)
    pos = torch.randn(2, N, C, device=DEV)
    with torch.no_grad():
        assert attention.mha_eligible(mha, x)
        got = attention.multihead_self_attention(mha, x + pos, x)
        want = mha(query=x + pos, key=x + pos, value=x, need_weights=False)[0]
    assert _err(got, want.double()) < 2e-5
    assert not attention.mha_eligible(mha.train(), x)              # attention dropout is live in train(): torch path
    with torch.enable_grad():
        assert not attention.mha_eligible(mha.eval(), x)           # autograd: torch path (the kernel is forward only)


@pytest.mark.gpu
def test_decoder_layer_uses_the_kernel_in_eval_and_matches_torch():
    from dpft_b200 import native
    from dpft_b200.models.fuser import MLFusion
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(2)
    layer = MLFusion(d_model=16, d_ffn=32, n_levels=1, n_heads=8, n_points=4, norm=True, dropout=0.1, activation="Mish").to(DEV).eval()
    x, pos = torch.randn(2, 300, 16, device=DEV), torch.randn(2, 300, 16, device=DEV)
    with torch.no_grad():
        l0 = native.launches()
        a = layer.forward_self_attn(x, pos)
        assert native.launches() == l0 + 1
        layer.native_self_attn = False
        b = layer.forward_self_attn(x, pos)
    assert _err(a, b.double()) < 2e-5


@pytest.mark.gpu
def test_decoder_self_attention_matches_the_oracle_restatement():
    """The whole self-attention sub-layer of MLFusion (in-projection, kernel, out-projection, residual, LayerNorm) against
    oracle/dprt_oracle.py::self_attention, the CPU restatement of mpfusion.py:122-148 pinned by the golden vectors."""
    from oracle import dprt_oracle
    from dpft_b200.models.fuser import MLFusion
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(5)
    layer = MLFusion(d_model=16, d_ffn=32, n_levels=1, n_heads=8, n_points=4, norm=True, dropout=0.1, activation="Mish").eval()
    with torch.no_grad():
        layer.self_attn.in_proj_bias.normal_(0, 0.3)
        layer.self_attn.out_proj.bias.normal_(0, 0.3)
    sd = {"l." + k: v for k, v in layer.state_dict().items()}
    x, pos = torch.randn(3, 400, 16), torch.rand(3, 400, 16)
    with torch.no_grad():
        want = x + dprt_oracle.self_attention(sd, "l.self_attn", x, pos, 8)
        want = torch.nn.functional.layer_norm(want, (16,), sd["l.norm1.weight"], sd["l.norm1.bias"], 1e-5)
        got = layer.to(DEV).forward_self_attn(x.to(DEV), pos.to(DEV)).cpu()
    assert float((got - want).abs().max()) < 1e-5


@pytest.mark.gpu
def test_decoder_self_attention_matches_the_reference_golden():
    """The same sub-layer through the tcgen05 kernel against the outputs of the UNMODIFIED reference (fixture
    tests/golden/mlfusion_self_attn.pt: 8 heads x 2 channels at 400 queries, 4 heads x 16 channels at 257 queries)."""
    from conftest import load_golden
    from dpft_b200 import native
    from dpft_b200.models.fuser import MLFusion
    torch.backends.cuda.matmul.allow_tf32 = False
    rec = load_golden("mlfusion_self_attn")
    for case in rec["cases"]:
        d, h = case["d_model"], case["n_heads"]
        layer = MLFusion(d_model=d, d_ffn=2 * d, n_levels=1, n_heads=h, n_points=4, norm=True, dropout=0.1, activation="Mish").eval()
        layer.load_state_dict(case["state_dict"], strict=False)
        layer = layer.to(DEV)
        with torch.no_grad():
            l0 = native.launches()
            got = layer.forward_self_attn(case["x"].to(DEV), case["pos"].to(DEV)).cpu()
            assert native.launches() == l0 + 1                       # the attention core ran in the native kernel
        assert float((got - case["out"]).abs().max()) < 2e-5, float((got - case["out"]).abs().max())


def test_rejects_cpu_tensors_and_wide_heads():
    x = torch.zeros(1, 4, 8)
    with pytest.raises(RuntimeError, match="not implemented on the CPU"):
        attention.self_attention(x, x, x, 2)
