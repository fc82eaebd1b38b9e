"""Parity of the sm_100a deformable-attention kernels (through the C ABI) with the CPU oracle."""
import pytest
import torch

from oracle import msda as O

pytestmark = pytest.mark.gpu

FRONT = [(37, 107), (10, 27), (5, 14), (3, 7), (2, 4)]
BEV = [(256, 107), (64, 27), (32, 14), (16, 7), (8, 4)]


def _cuda(value, shapes, loc, attn, grad_out=None):
    dev = "cuda:0"
    sh = torch.tensor(shapes, dtype=torch.int64, device=dev)
    lsi = O.level_start_index(shapes).to(dev)
    args = [value.to(dev), sh, lsi, loc.to(dev), attn.to(dev)]
    if grad_out is not None:
        args.append(grad_out.to(dev))
    return args


def _tols(dtype):
    return {torch.float64: 1e-11, torch.float32: 1e-5, torch.float16: 2e-3, torch.bfloat16: 1.6e-2}[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("D", [1, 2, 3, 4, 8, 16, 32, 64, 128, 24])
@pytest.mark.parametrize("shapes,P", [(FRONT, 4), ([(6, 4), (3, 2)], 2), ([(1, 1)], 1), (BEV[:4], 8)])
def test_forward_and_backward_match_oracle(dtype, D, shapes, P):
    from dpft_b200 import msda
    B, N, M = 2, 37, 3
    ref_dtype = dtype if dtype in (torch.float32, torch.float64) else torch.float32
    v, sh, loc, a, go = O.random_problem(B, N, M, D, shapes, P, dtype=dtype, seed=D * 7 + P, spread=0.3)
    want = O.msda_forward_c(v.to(ref_dtype), sh, loc.to(ref_dtype), a.to(ref_dtype))
    got = msda.ms_deform_attn_forward(*_cuda(v, sh, loc, a), 64)
    assert got.dtype == dtype and got.shape == (B, N, M * D)
    tol = _tols(dtype)
    assert torch.allclose(got.cpu().to(ref_dtype), want, atol=tol * 4, rtol=tol), (got.cpu().to(ref_dtype) - want).abs().max()

    wv, wl, wa = O.msda_backward_c(v.to(ref_dtype), sh, loc.to(ref_dtype), a.to(ref_dtype), go.to(ref_dtype))
    gv, gl, ga = msda.ms_deform_attn_backward(*_cuda(v, sh, loc, a, go), 64)
    assert gv.dtype == dtype and gl.shape == loc.shape and ga.shape == a.shape
    for name, g, w, scale in (("value", gv, wv, 4), ("loc", gl, wl, 400), ("attn", ga, wa, 40)):
        err = (g.cpu().to(ref_dtype) - w).abs().max()
        assert torch.allclose(g.cpu().to(ref_dtype), w, atol=tol * scale, rtol=tol * 4), (name, float(err))


def test_shipped_configuration_shapes():
    """B=8, N=400, M=8, D=2, L=5, P=4 on the radar pyramid: fp32 within 1e-5 of the oracle."""
    from dpft_b200 import msda
    v, sh, loc, a, go = O.random_problem(8, 400, 8, 2, BEV, 4, dtype=torch.float32, seed=1, spread=0.1)
    got = msda.ms_deform_attn_forward(*_cuda(v, sh, loc, a), 64).cpu()
    assert torch.allclose(got, O.msda_forward_c(v, sh, loc, a), atol=1e-5, rtol=1e-5)
    gv, gl, ga = msda.ms_deform_attn_backward(*_cuda(v, sh, loc, a, go), 64)
    wv, wl, wa = O.msda_backward_c(v, sh, loc, a, go)
    assert torch.allclose(gv.cpu(), wv, atol=2e-4, rtol=1e-4)      # float atomics: order-dependent rounding
    assert torch.allclose(gl.cpu(), wl, atol=2e-3, rtol=1e-4)
    assert torch.allclose(ga.cpu(), wa, atol=1e-4, rtol=1e-4)


def test_edge_locations_known_answers():
    from dpft_b200 import msda
    m = torch.tensor([[1.0, 2.0, 3.0, 4.0], [5.0, 6.0, 7.0, 8.0]])
    pts = [(0.125, 0.25, 1.0), (0.25, 0.5, 3.5), (0.0, 0.0, 0.25), (1.0, 1.0, 2.0), (-0.2, 0.5, 0.0),
           (0.5, 1.6, 0.0), (float("nan"), 0.5, 0.0), (float("inf"), 0.5, 0.0)]
    v = m.reshape(1, 8, 1, 1)
    loc = torch.tensor([[x, y] for x, y, _ in pts]).view(1, len(pts), 1, 1, 1, 2)
    a = torch.ones(1, len(pts), 1, 1, 1)
    got = msda.ms_deform_attn_forward(*_cuda(v, [(2, 4)], loc, a), 64).cpu().flatten()
    assert torch.allclose(got, torch.tensor([w for _, _, w in pts]), atol=1e-6)


def test_empty_and_error_paths():
    from dpft_b200 import msda
    v, sh, loc, a, _ = O.random_problem(2, 0, 2, 4, [(4, 4)], 2)
    assert msda.ms_deform_attn_forward(*_cuda(v, sh, loc, a), 64).shape == (2, 0, 8)
    v, sh, loc, a, _ = O.random_problem(2, 3, 2, 4, [(4, 4)], 2)
    args = _cuda(v, sh, loc, a)
    args[3] = args[3].transpose(1, 2)                       # non-contiguous
    with pytest.raises(RuntimeError, match="contiguous"):
        msda.ms_deform_attn_forward(*args, 64)
    with pytest.raises(RuntimeError):
        msda.ms_deform_attn_forward(*_cuda(v, [(4, 4)] * 17, loc, a), 64)


def test_full_size_properties():
    """BASELINE sizes (cfg 5: 900 queries, 4 levels, bf16, bs=16, d_model=256) are too big for the CPU oracle in
    seconds: check size-independent properties instead — linearity in value, weights that sum to one reproduce
    a constant map, zero weights give zero, and the backward is the adjoint of the forward."""
    from dpft_b200 import msda
    dev = "cuda:0"
    shapes = [(180, 320), (90, 160), (45, 80), (23, 40)]
    B, N, M, D, P = 16, 900, 8, 32, 4
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    g = torch.Generator(device=dev).manual_seed(3)
    sh = torch.tensor(shapes, dtype=torch.int64, device=dev)
    lsi = O.level_start_index(shapes).to(dev)
    loc = torch.rand(B, N, M, L, P, 2, generator=g, device=dev) * 0.9 + 0.05      # fully inside
    attn = torch.softmax(torch.randn(B, N, M, L * P, generator=g, device=dev), -1).view(B, N, M, L, P)
    ones = torch.ones(B, S, M, D, device=dev)
    out = msda.ms_deform_attn_forward(ones.bfloat16(), sh, lsi, loc.bfloat16(), attn.bfloat16(), 64)
    assert torch.allclose(out.float(), torch.ones_like(out.float()), atol=3e-2)
    out = msda.ms_deform_attn_forward(ones, sh, lsi, loc, attn, 64)
    assert torch.allclose(out, torch.ones_like(out), atol=1e-5)
    assert msda.ms_deform_attn_forward(ones, sh, lsi, loc, torch.zeros_like(attn), 64).abs().max() == 0
    v1 = torch.randn(B, S, M, D, generator=g, device=dev)
    v2 = torch.randn(B, S, M, D, generator=g, device=dev)
    o1 = msda.ms_deform_attn_forward(v1, sh, lsi, loc, attn, 64)
    o2 = msda.ms_deform_attn_forward(v2, sh, lsi, loc, attn, 64)
    o12 = msda.ms_deform_attn_forward(v1 + 0.5 * v2, sh, lsi, loc, attn, 64)
    assert torch.allclose(o12, o1 + 0.5 * o2, atol=1e-4)
    go = torch.randn_like(o1)
    gv, _, ga = msda.ms_deform_attn_backward(v1, sh, lsi, loc, attn, go, 64)
    lhs = (o1.double() * go.double()).sum()
    assert abs(float((gv.double() * v1.double()).sum() - lhs)) < 1e-4 * abs(float(lhs)) + 1e-2     # <J v, g> = <v, J^T g>
    assert abs(float((ga.double() * attn.double()).sum() - lhs)) < 1e-4 * abs(float(lhs)) + 1e-2


def test_autograd_function_matches_reference_signature():
    from dpft_b200 import msda
    v, sh, loc, a, go = O.random_problem(2, 5, 4, 8, [(6, 4), (3, 2)], 2, dtype=torch.float64, seed=4)
    args = _cuda(v, sh, loc, a)
    for i in (0, 3, 4):
        args[i].requires_grad_(True)
    out = msda.MSDeformAttnFunction.apply(*args, 64)
    out.backward(go.to("cuda:0"))
    wv, wl, wa = O.msda_backward_torch(v, sh, loc, a, go)
    assert torch.allclose(args[0].grad.cpu(), wv, atol=1e-10)
    assert torch.allclose(args[3].grad.cpu(), wl, atol=1e-9)
    assert torch.allclose(args[4].grad.cpu(), wa, atol=1e-10)
