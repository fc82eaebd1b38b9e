"""tcgen05 convolution kernel vs a plain PyTorch fp32 reference of the same op (bf16-rounded operands, fp32 math).
Floating-point kernel: tolerance 1e-2 rel (north_star's bf16 bar); the fp32 accumulation itself is checked tighter
by comparing against the fp32 conv of the SAME bf16-rounded inputs (only the output rounding remains: 2^-8)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (B, H, W, Cin, Cout, R, stride, pad) — the ResNet-50/101 bottleneck shapes at small spatial sizes + tails
SHAPES = [
    (2, 16, 24, 64, 64, 1, 1, 0),        # 1x1, tiled-2D path, M multiple of 128
    (1, 9, 7, 64, 256, 1, 1, 0),         # M = 63: one partial tile
    (2, 23, 40, 256, 64, 1, 1, 0),       # K = 256 (4 k-blocks), M tail
    (1, 12, 20, 1024, 256, 1, 1, 0),     # deep K, wraps the stage ring
    (2, 16, 24, 64, 64, 3, 1, 1),        # 3x3 im2col
    (1, 9, 7, 128, 128, 3, 1, 1),        # 3x3, tiny map, tile spans rows and images
    (3, 10, 27, 128, 128, 3, 2, 1),      # 3x3 stride 2, odd sizes (radar_front)
    (2, 23, 40, 256, 512, 1, 2, 0),      # 1x1 stride 2 (downsample)
    (1, 2, 4, 512, 512, 3, 1, 1),        # radar_front layer4: 2x4 map (< 128 KiB tensor: driver fix-up path)
    (2, 8, 8, 512, 2048, 1, 1, 0),       # widest Cout
]


def _reference(x, w, bias, stride, pad, relu, residual):
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, stride=stride, padding=pad)
    y = y.permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual.float()
    return torch.relu(y) if relu else y


# every output-channel tile width that divides the layer's Cout (0 = the library's choice)
SHAPE_TILES = [(shape, bn) for shape in SHAPES for bn in (0, 64, 128, 256) if bn == 0 or shape[4] % bn == 0]


@pytest.mark.parametrize("shape,block_n", SHAPE_TILES)
@pytest.mark.parametrize("relu,with_res", [(True, False), (True, True), (False, False)])
def test_conv_matches_torch_fp32(shape, block_n, relu, with_res):
    from dpft_b200 import conv
    B, H, W, Cin, Cout, R, stride, pad = shape
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(Cin + Cout + R)
    x = torch.randn(B, H, W, Cin, generator=g, device=dev).bfloat16()
    w = (torch.randn(Cout, R, R, Cin, generator=g, device=dev) / (R * R * Cin) ** 0.5).bfloat16()
    bias = torch.randn(Cout, generator=g, device=dev)
    P, Q = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    res = torch.randn(B, P, Q, Cout, generator=g, device=dev).bfloat16() if with_res else None
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    want = _reference(x, w, bias, stride, pad, relu, res)
    scale = want.abs().max().item()
    for cluster_mode in (1, 2):      # one CTA per tile; CTA pairs (tcgen05 cta_group::2) where the tile allows it
        got = conv.conv2d_nhwc(x, w, bias, stride, pad, relu, res, block_n=block_n, cluster_mode=cluster_mode)
        torch.cuda.synchronize()
        assert got.shape == (B, P, Q, Cout) and got.dtype == torch.bfloat16
        err = (got.float() - want).abs().max().item()
        assert err <= 1e-2 * scale, (cluster_mode, err, scale)         # north_star bf16 bar
        assert err <= 2.0 ** -7 * scale, (cluster_mode, err, scale)    # only the bf16 output rounding should remain


# the pointwise "expand" convs with a residual (conv3 of every Bottleneck: the stream-bound configuration of the kernel): several
# waves of m-tiles per CTA, M tails, 1 / 2 / 4 n-tiles, grid smaller and larger than the SM count
EXPAND_SHAPES = [(2, 45, 80, 256, 1024), (3, 30, 33, 128, 512), (2, 16, 24, 64, 256), (8, 45, 80, 256, 1024), (1, 5, 5, 192, 768),
                 (4, 90, 160, 128, 512), (1, 3, 5, 64, 256), (2, 90, 160, 64, 256), (1, 2, 2, 256, 2048), (8, 23, 41, 256, 1280)]


@pytest.mark.parametrize("shape", EXPAND_SHAPES)
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("relu", [True, False])
def test_expand_conv_with_residual_matches_torch_fp32(shape, dtype, relu):
    from dpft_b200 import conv
    B, H, W, Cin, Cout = shape
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(Cin + Cout + H)
    x = torch.randn(B, H, W, Cin, generator=g, device=dev).to(dtype)
    w = (torch.randn(Cout, 1, 1, Cin, generator=g, device=dev) / Cin ** 0.5).to(dtype)
    bias = torch.randn(Cout, generator=g, device=dev)
    res = torch.randn(B, H, W, Cout, generator=g, device=dev).to(dtype)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    want = _reference(x, w, bias, 1, 0, relu, res)
    scale = want.abs().max().item()
    got = conv.conv2d_nhwc(x, w, bias, 1, 0, relu, res, block_n=256, cluster_mode=1)
    again = conv.conv2d_nhwc(x, w, bias, 1, 0, relu, res, block_n=256, cluster_mode=1)
    torch.cuda.synchronize()
    assert torch.equal(got, again)                                        # deterministic
    err = (got.float() - want).abs().max().item()
    assert err <= (2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10) * scale, (err, scale)   # output rounding only
    auto = conv.conv2d_nhwc(x, w, bias, 1, 0, relu, res)                  # whatever tile the heuristic picks agrees
    assert (auto.float() - got.float()).abs().max().item() <= 2.0 ** -7 * scale
    if Cin <= 256:
        # cluster_mode 3: the weight-stationary kernel (resident weight slice, in-place residual / output ring) — the same
        # fp32 sums in the same order as the generic kernel, so the 16-bit outputs must be bit-identical
        ws = conv.conv2d_nhwc(x, w, bias, 1, 0, relu, res, cluster_mode=3)
        ws2 = conv.conv2d_nhwc(x, w, bias, 1, 0, relu, res, cluster_mode=3)
        torch.cuda.synchronize()
        assert torch.equal(ws, ws2)
        assert torch.equal(ws, got), (ws.float() - got.float()).abs().max().item()


# 64 -> 64 channel 3x3 layers on wide maps go through the halo-tile kernel (conv3x3_halo.cu): width / height tails, a width that
# is not a multiple of the 128-column tile, odd heights (last tile has one row), several images, more tiles than SMs
HALO_SHAPES = [(1, 2, 128), (2, 7, 130), (1, 9, 96), (3, 33, 200), (2, 16, 320), (8, 45, 257), (1, 181, 131)]


@pytest.mark.parametrize("shape", HALO_SHAPES)
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("relu", [True, False])
def test_halo_conv3x3_matches_torch_fp32(shape, dtype, relu):
    from dpft_b200 import conv
    B, H, W = shape
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(H * 1000 + W)
    x = torch.randn(B, H, W, 64, generator=g, device=dev).to(dtype)
    w = (torch.randn(64, 3, 3, 64, generator=g, device=dev) / 24.0).to(dtype)
    bias = torch.randn(64, generator=g, device=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    want = _reference(x, w, bias, 1, 1, relu, None)
    scale = want.abs().max().item()
    got = conv.conv2d_nhwc(x, w, bias, 1, 1, relu)                       # heuristic path = halo kernel for these shapes
    generic = conv.conv2d_nhwc(x, w, bias, 1, 1, relu, block_n=64)       # forcing a tile keeps the im2col kernel
    torch.cuda.synchronize()
    assert got.shape == (B, H, W, 64) and got.dtype == dtype
    err = (got.float() - want).abs().max().item()
    assert err <= (2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10) * scale, (err, scale)   # output rounding only
    # both kernels accumulate the same 576 products in fp32 (in a different order): at most one 16-bit ulp apart
    assert (got.float() - generic.float()).abs().max().item() <= 2.0 ** -7 * scale


def test_folded_bottleneck_matches_torch_block():
    """conv+bn folding and the residual/ReLU epilogue reproduce a torchvision-style bottleneck in eval mode."""
    from dpft_b200 import conv
    from dpft_b200.models.backbone import Bottleneck
    from torch import nn
    dev = "cuda:0"
    torch.manual_seed(0)
    blk = Bottleneck(256, 128, stride=2, norm=nn.BatchNorm2d, project=True).to(dev).eval()
    for m in blk.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.2)
    x = torch.randn(2, 256, 18, 30, device=dev)
    with torch.no_grad():
        want = blk(x).permute(0, 2, 3, 1)
    xb = x.permute(0, 2, 3, 1).contiguous().bfloat16()
    c1, c2, c3 = (conv.FoldedConv(blk.conv1, blk.bn1, dev), conv.FoldedConv(blk.conv2, blk.bn2, dev),
                  conv.FoldedConv(blk.conv3, blk.bn3, dev))
    ds = conv.FoldedConv(blk.downsample[0], blk.downsample[1], dev)
    idn = ds(xb, relu=False)
    y = c3(c2(c1(xb, relu=True), relu=True), relu=True, residual=idn)
    torch.cuda.synchronize()
    rel = (y.float() - want).abs().max().item() / want.abs().max().item()
    assert rel < 2e-2, rel


@pytest.mark.parametrize("budget", [1, 2, 7, 40])
@pytest.mark.parametrize("case", [(2, 32, 32, 128, 512, 1, 1, 0, True), (2, 16, 16, 256, 256, 3, 1, 1, False),
                                  (1, 64, 64, 64, 64, 3, 1, 1, False), (2, 32, 32, 512, 128, 1, 1, 0, False),
                                  (8, 45, 80, 256, 1024, 1, 1, 0, True)])
def test_cta_budget_changes_only_the_grid(budget, case):
    """dpft_conv2d_nhwc_ex with a cap on the persistent grid (the engine runs the side views under it): every tile is still
    computed, by fewer CTAs, so the output is bit-identical to the uncapped launch — generic kernel, CTA pairs and the
    weight-stationary kernel (whose grid must stay a multiple of its n-tile count)."""
    from dpft_b200 import conv
    B, H, W, Cin, Cout, R, stride, pad, with_res = case
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(budget + Cin)
    x = torch.randn(B, H, W, Cin, generator=g, device=dev).half()
    w = (torch.randn(Cout, R, R, Cin, generator=g, device=dev) / (R * R * Cin) ** 0.5).half()
    bias = torch.randn(Cout, generator=g, device=dev)
    res = torch.randn(B, H, W, Cout, generator=g, device=dev).half() if with_res else None
    want = conv.conv2d_nhwc(x, w, bias, stride, pad, True, res)
    with conv.cta_budget(budget):
        got = conv.conv2d_nhwc(x, w, bias, stride, pad, True, res)
    after = conv.conv2d_nhwc(x, w, bias, stride, pad, True, res)          # the cap does not leak into later calls
    torch.cuda.synchronize()
    assert torch.equal(got, want) and torch.equal(after, want)


PAIR_CASES = [(8, 64, 64, 64, 64, 1, 1, 0, False),        # radar stage-1 conv1 (generic kernel, 64-wide tiles)
              (8, 64, 64, 64, 64, 3, 1, 1, False),        # 3x3 on a 64-wide map (below the halo kernel's width)
              (8, 64, 64, 64, 256, 1, 1, 0, True),        # expand + residual, too few tiles for the weight-stationary kernel
              (8, 32, 32, 256, 128, 1, 2, 0, False),      # strided 1x1
              (8, 16, 16, 256, 256, 3, 1, 1, False),      # deep K: CTA pairs (cta_group::2) with two problems in the grid
              (8, 8, 8, 1024, 512, 1, 1, 0, False),
              (2, 16, 100, 64, 64, 3, 1, 1, False),       # halo-kernel layer: the pair entry declines, two launches instead
              (8, 45, 80, 256, 1024, 1, 1, 0, True)]      # weight-stationary layer: declined as well


@pytest.mark.parametrize("case", PAIR_CASES)
@pytest.mark.parametrize("budget", [0, 48])
def test_pair_launch_equals_two_launches(case, budget):
    """dpft_conv2d_nhwc_pair: two problems of identical shape in one launch (gridDim.y = 2) must give, bit for bit, what two
    launches give — generic kernel, CTA pairs, with and without the persistent-grid cap; layers of the halo / weight-stationary
    kernels fall back to two launches."""
    from dpft_b200 import conv
    B, H, W, Cin, Cout, R, stride, pad, with_res = case
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(Cin * 7 + Cout + R)
    xs = [torch.randn(B, H, W, Cin, generator=g, device=dev).half() for _ in range(2)]
    ws = [(torch.randn(Cout, R, R, Cin, generator=g, device=dev) / (R * R * Cin) ** 0.5).half() for _ in range(2)]
    bs = [torch.randn(Cout, generator=g, device=dev) for _ in range(2)]
    P, Q = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    rs = [torch.randn(B, P, Q, Cout, generator=g, device=dev).half() for _ in range(2)] if with_res else [None, None]
    with conv.cta_budget(budget):
        want = [conv.conv2d_nhwc(xs[i], ws[i], bs[i], stride, pad, True, rs[i]) for i in range(2)]
        got = conv.conv2d_nhwc_pair(xs, ws, bs, stride, pad, True, rs)
    torch.cuda.synchronize()
    assert not torch.equal(want[0], want[1])
    for i in range(2):
        assert torch.equal(got[i], want[i]), i
