"""Training kernels (weight gradient, data gradient, BatchNorm train fwd/bwd, max-pool backward, weight re-layout) vs plain
PyTorch fp32 autograd of the same op on the SAME 16-bit-rounded inputs.  Floating point: tolerances written per check
(the kernels accumulate in fp32; what remains is the 16-bit rounding of the outputs that are stored in 16 bits)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"

# (B, H, W, Cin, Cout, R, stride, pad): bottleneck shapes at small spatial sizes, ragged pixel counts
WGRAD_SHAPES = [
    (2, 16, 24, 64, 64, 1, 1, 0),        # Cout 64: upper half of the 128-row tile is out of bounds
    (1, 9, 7, 64, 256, 1, 1, 0),         # 63 pixels: one partial k-block
    (2, 23, 40, 256, 64, 1, 1, 0),       # N tile 256
    (1, 12, 20, 1024, 256, 1, 1, 0),     # four ci tiles, two co tiles
    (2, 16, 24, 64, 64, 3, 1, 1),        # 3x3 im2col
    (1, 9, 7, 128, 128, 3, 1, 1),        # N tile 128; k-block spans rows
    (3, 10, 27, 128, 128, 3, 2, 1),      # 3x3 stride 2, odd sizes, k-blocks span images
    (2, 23, 40, 256, 512, 1, 2, 0),      # 1x1 stride 2 (downsample)
    (1, 2, 4, 512, 512, 3, 1, 1),        # tiny map (driver fix-up path of the im2col map)
    (2, 8, 8, 512, 2048, 1, 1, 0),       # widest Cout
]


def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("shape", WGRAD_SHAPES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("splits", [0, 1, 3])
def test_wgrad_matches_torch(shape, dtype, splits):
    from dpft_b200 import train_ops
    _no_tf32()
    B, H, W, Cin, Cout, R, stride, pad = shape
    g = torch.Generator(device=DEV).manual_seed(Cin * 7 + Cout + R + stride)
    x = torch.randn(B, H, W, Cin, generator=g, device=DEV).to(dtype)
    P, Q = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    dy = torch.randn(B, P, Q, Cout, generator=g, device=DEV).to(dtype)
    want = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Cout, Cin, R, R), dy.float().permute(0, 3, 1, 2),
                                       stride=stride, padding=pad).permute(0, 2, 3, 1)
    got = train_ops.conv2d_wgrad(x, dy, R, R, stride, pad, splits=splits)
    torch.cuda.synchronize()
    assert got.shape == (Cout, R, R, Cin) and got.dtype == torch.float32
    scale = want.abs().max().item()
    err = (got - want).abs().max().item()
    assert err <= 2e-5 * scale * max(1.0, (B * P * Q) ** 0.5 / 8), (err, scale)     # fp32 accumulation order only
    # accumulates into an existing buffer
    got2 = train_ops.conv2d_wgrad(x, dy, R, R, stride, pad, out=got.clone(), splits=splits)
    torch.cuda.synchronize()
    assert (got2 - 2 * want).abs().max().item() <= 4e-5 * scale * max(1.0, (B * P * Q) ** 0.5 / 8)


@pytest.mark.parametrize("shape", WGRAD_SHAPES)
def test_dgrad_matches_torch(shape):
    from dpft_b200 import train_ops
    _no_tf32()
    B, H, W, Cin, Cout, R, stride, pad = shape
    dtype = torch.bfloat16
    g = torch.Generator(device=DEV).manual_seed(Cin + Cout * 3 + R)
    w = torch.randn(Cout, Cin, R, R, generator=g, device=DEV) / (R * R * Cout) ** 0.5
    P, Q = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    dy = torch.randn(B, P, Q, Cout, generator=g, device=DEV).to(dtype)
    res = torch.randn(B, H, W, Cin, generator=g, device=DEV).to(dtype)
    packer = train_ops.WeightPacker([w], dtype, [True])
    packer.refresh()
    wq = packer.fwd[0].float().permute(0, 3, 1, 2)                 # the rounded weights the kernels see, torch layout
    assert torch.equal(packer.fwd[0], w.permute(0, 2, 3, 1).to(dtype))
    want = torch.nn.grad.conv2d_input((B, Cin, H, W), wq, dy.float().permute(0, 3, 1, 2), stride=stride, padding=pad).permute(0, 2, 3, 1)
    zero_bias = torch.zeros(max(Cin, Cout), device=DEV)
    for r in (None, res):
        got = train_ops.conv2d_dgrad(dy, packer.dgrad[0], zero_bias, (H, W), stride, pad, residual=r)
        torch.cuda.synchronize()
        ref = want if r is None else want + r.float()
        scale = ref.abs().max().item()
        assert got.shape == (B, H, W, Cin)
        assert (got.float() - ref).abs().max().item() <= 2.0 ** -6 * scale


@pytest.mark.parametrize("C", [64, 128, 256, 512, 1024, 2048])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("relu,with_res", [(True, False), (True, True), (False, False)])
def test_batchnorm_train_forward_backward(C, dtype, relu, with_res):
    from dpft_b200 import train_ops
    shape = (3, 11, 13, C) if C <= 256 else (2, 5, 7, C)
    g = torch.Generator(device=DEV).manual_seed(C)
    y = (torch.randn(shape, generator=g, device=DEV) * 2 + 0.5).to(dtype)
    res = torch.randn(shape, generator=g, device=DEV).to(dtype) if with_res else None
    gamma = torch.rand(C, generator=g, device=DEV) + 0.5
    beta = torch.randn(C, generator=g, device=DEV) * 0.1
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    dz = torch.randn(shape, generator=g, device=DEV).to(dtype)

    yf = y.float().requires_grad_(True)
    gf, bf = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rf = res.float().requires_grad_(True) if with_res else None
    out = F.batch_norm(yf.permute(0, 3, 1, 2), rm_ref, rv_ref, gf, bf, training=True, momentum=0.1, eps=1e-5).permute(0, 2, 3, 1)
    if with_res:
        out = out + rf
    if relu:
        out = torch.relu(out)

    z, state = train_ops.bn_forward(y, gamma, beta, rm, rv, 0.1, 1e-5, relu, res)
    torch.cuda.synchronize()
    tol = 2.0 ** (-7 if dtype == torch.bfloat16 else -10)
    assert (z.float() - out).abs().max().item() <= tol * out.abs().max().item() + 1e-6
    assert torch.allclose(rm, rm_ref, rtol=1e-4, atol=1e-5) and torch.allclose(rv, rv_ref, rtol=1e-4, atol=1e-5)

    # backward against autograd, with the ReLU mask taken from OUR z (a value that rounds to 0 in 16 bits masks the gradient)
    mask = (z.float() > 0).float() if relu else torch.ones_like(out)
    pre = F.batch_norm(yf.permute(0, 3, 1, 2), None, None, gf, bf, training=True, eps=1e-5).permute(0, 2, 3, 1)
    (pre * (dz.float() * mask)).sum().backward()
    dgamma, dbeta = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    dy, gres = train_ops.bn_backward(dz, z if relu else None, y, state, gamma, relu, dgamma, dbeta, want_g=with_res)
    torch.cuda.synchronize()
    assert (dy.float() - yf.grad).abs().max().item() <= 2 * tol * yf.grad.abs().max().item()
    assert torch.allclose(dgamma, gf.grad, rtol=2e-3, atol=2e-3 * gf.grad.abs().max().item())
    assert torch.allclose(dbeta, bf.grad, rtol=2e-3, atol=2e-3 * bf.grad.abs().max().item())
    if with_res:
        assert torch.equal(gres.float(), (dz.float() * mask).to(dtype).float())


@pytest.mark.parametrize("shape", [(2, 12, 16, 64), (1, 9, 7, 64), (2, 37, 54, 64)])
def test_maxpool_backward_matches_torch(shape):
    from dpft_b200 import features, train_ops
    g = torch.Generator(device=DEV).manual_seed(shape[1])
    # post-ReLU-like input with many exact ties (bf16 grid + zeros) to exercise the first-maximum rule
    x = torch.relu(torch.randn(shape, generator=g, device=DEV)).mul(4).round().div(4).bfloat16()
    B, H, W, C = shape
    P, Q = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    dy = torch.randn(B, P, Q, C, generator=g, device=DEV).bfloat16()
    xf = x.float().permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    out = F.max_pool2d(xf, 3, 2, 1)
    out.backward(dy.float().permute(0, 3, 1, 2))
    assert torch.equal(features.maxpool_forward(x).float(), out.detach().permute(0, 2, 3, 1))
    got = train_ops.maxpool_backward(x, features.maxpool_forward(x), dy)
    torch.cuda.synchronize()
    want = xf.grad.permute(0, 2, 3, 1)
    assert (got.float() - want).abs().max().item() <= 2.0 ** -7 * want.abs().max().item()


@pytest.mark.parametrize("cin,shape", [(3, (2, 70, 150)), (6, (2, 37, 107)), (3, (1, 256, 320)), (6, (3, 64, 64))])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_stem_train_forward_and_wgrad(cin, shape, dtype):
    """Stem without ReLU (pre-BatchNorm activation) and its weight gradient vs torch fp32 on the same inputs.  The tensor-core
    stem rounds x and w to f16 (11-bit mantissa): forward within 2e-3 of the output scale for f16 storage (1e-2 for bf16);
    the weight gradient reads x in fp32: 1e-4 of its scale."""
    from dpft_b200 import features, train_ops
    _no_tf32()
    B, H, W = shape
    g = torch.Generator(device=DEV).manual_seed(cin + H)
    x = torch.rand(B, H, W, cin, generator=g, device=DEV) * 255
    w = torch.randn(64, cin, 7, 7, generator=g, device=DEV) / (49 * cin) ** 0.5 / 64
    w_k = w.permute(2, 3, 1, 0).contiguous()                              # [7][7][Cin][64]
    want = F.conv2d(x.permute(0, 3, 1, 2), w, stride=2, padding=3).permute(0, 2, 3, 1)
    zero = torch.zeros(64, device=DEV)
    for impl in (1, 2):
        if impl == 2 and want.shape[2] < 8:
            continue
        got = features.stem_forward(x, w_k, zero, dtype, impl=impl, w_packed=features.stem_pack_weights(w_k), relu=False)
        torch.cuda.synchronize()
        assert got.dtype == dtype and got.shape == want.shape
        assert (got.float() < 0).any(), "the ReLU was applied"
        tol = 2e-3 if dtype == torch.float16 else 1e-2
        assert (got.float() - want).abs().max().item() <= tol * want.abs().max().item(), impl
    P, Q = want.shape[1], want.shape[2]
    dy = torch.randn(B, P, Q, 64, generator=g, device=DEV).to(dtype)
    ref = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2), (64, cin, 7, 7), dy.float().permute(0, 3, 1, 2), stride=2, padding=3)
    got = train_ops.stem_wgrad(x, dy)
    torch.cuda.synchronize()
    assert got.shape == (7, 7, cin, 64)
    assert (got.permute(3, 2, 0, 1) - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    got2 = train_ops.stem_wgrad(x, dy, out=got.clone())                    # accumulates
    torch.cuda.synchronize()
    assert (got2.permute(3, 2, 0, 1) - 2 * ref).abs().max().item() <= 2e-4 * ref.abs().max().item()


def test_weight_packer_layouts():
    from dpft_b200 import train_ops
    g = torch.Generator(device=DEV).manual_seed(5)
    ws = [torch.randn(64, 64, 1, 1, generator=g, device=DEV), torch.randn(128, 64, 3, 3, generator=g, device=DEV),
          torch.randn(256, 128, 1, 1, generator=g, device=DEV)]
    for dtype in (torch.bfloat16, torch.float16):
        pk = train_ops.WeightPacker(ws, dtype, [True, True, False])
        pk.refresh()
        torch.cuda.synchronize()
        for w, f, d in zip(ws, pk.fwd, pk.dgrad):
            assert torch.equal(f, w.permute(0, 2, 3, 1).to(dtype))
            if d is not None:
                assert torch.equal(d, w.flip(2, 3).permute(1, 2, 3, 0).to(dtype))
        ws[1].mul_(2.0)                                    # an optimiser step changes the masters in place
        pk.refresh()
        torch.cuda.synchronize()
        assert torch.equal(pk.fwd[1], ws[1].permute(0, 2, 3, 1).to(dtype))
        ws[1].mul_(0.5)
