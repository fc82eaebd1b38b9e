"""Native feature path (stem, maxpool, tcgen05 bottlenecks, FPN + positional embedding) against the plain PyTorch fp32
modules of the same model.  bf16 activations: tolerance 1e-2 rel (north_star's bf16 bar) unless noted."""
import pytest
import torch
import torch.nn.functional as F

from dpft_b200 import configs, models, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _strict_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12))


def _model(name, seed=3):
    cfg = synthetic.offline_config(configs.make_config(name))
    m = models.build("dprt", cfg).eval()
    m.load_state_dict(synthetic.seeded_state_dict(m.state_dict(), seed=seed))
    return cfg, m.to(DEV)


@pytest.mark.parametrize("name,view,size", [("kradar_camera_mono", "camera_mono", (2, 70, 101, 3)),
                                            ("kradar_radar_bev", "radar_bev", (2, 64, 107, 6)),
                                            ("kradar_radar_front", "radar_front", (1, 37, 107, 6))])
def test_stem_and_maxpool(name, view, size):
    from dpft_b200 import features
    from dpft_b200.features import NativeView
    cfg, m = _model(name)
    nv = NativeView(m.backbones[view], m.necks[view], m.embeddings[view], True, DEV)
    x = torch.rand(*size, device=DEV) * 255.0
    bb = m.backbones[view]
    with torch.no_grad():
        t = bb.adjustment_layer(x.movedim(-1, 1))
        t = F.relu(bb.body.bn1(bb.body.conv1(t)))
        want_stem = t.movedim(1, -1)
        want_pool = F.max_pool2d(t, 3, 2, 1).movedim(1, -1)
    for impl in (1, 2):                                 # fp32 CUDA-core kernel, tcgen05 kernel with f16 operands
        for dt in (torch.float16, torch.bfloat16):
            got_stem = features.stem_forward(x, nv.stem_w, nv.stem_b, dt, impl=impl)
            assert got_stem.shape == want_stem.shape and got_stem.dtype == dt
            assert _rel(got_stem, want_stem) < (2e-3 if dt == torch.float16 else 1e-2), (impl, dt, _rel(got_stem, want_stem))
    got_pool = features.maxpool_forward(got_stem)
    assert got_pool.shape == want_pool.shape
    ref_pool = F.max_pool2d(got_stem.float().movedim(-1, 1), 3, 2, 1).movedim(1, -1)
    assert torch.equal(got_pool.float(), ref_pool)                      # max is exact


@pytest.mark.parametrize("name,view,size", [("kradar_radar_bev", "radar_bev", (2, 64, 107, 6)),
                                            ("kradar_radar_front", "radar_front", (2, 37, 107, 6)),
                                            ("kradar_camera_mono", "camera_mono", (1, 90, 160, 3))])
def test_backbone_and_pyramid(name, view, size):
    from dpft_b200.features import NativeView
    from dpft_b200.models.fuser import FeaturePyramid
    cfg, m = _model(name)
    nv = NativeView(m.backbones[view], m.necks[view], m.embeddings[view], True, DEV, torch.float16)
    x = torch.rand(*size, device=DEV) * 255.0
    with torch.no_grad():
        want_feats = m.backbones[view](x)
        want_pyr = FeaturePyramid.from_levels(m.extract_features({view: x})[view])
    got_feats = nv.backbone(x)
    for g, (k, w) in zip(got_feats, want_feats.items()):
        assert g.shape == w.shape, k
        assert _rel(g, w) < 5e-3, (k, _rel(g, w))          # f16 activations through up to 33 blocks
    flat, shapes = nv.pyramid(x)
    assert shapes == want_pyr.shapes and flat.shape == want_pyr.flat.shape and flat.dtype == torch.float32
    assert _rel(flat, want_pyr.flat) < 5e-3, _rel(flat, want_pyr.flat)
    nv16 = NativeView(m.backbones[view], m.necks[view], m.embeddings[view], True, DEV, torch.float16, torch.float16)
    flat16, _ = nv16.pyramid(x)                         # f16 storage of the pyramid: one more 2^-11 rounding
    assert flat16.dtype == torch.float16 and _rel(flat16, flat) < 1e-3, _rel(flat16, flat)


def test_fpn_stages_in_isolation():
    """lateral GEMM + top-down + 3x3 + positional embedding against torch on the SAME bf16-rounded features (tight)."""
    from dpft_b200 import features
    from dpft_b200.features import NativeView
    cfg, m = _model("kradar_radar_bev")
    view = "radar_bev"
    nv = NativeView(m.backbones[view], m.necks[view], m.embeddings[view], True, DEV)
    g = torch.Generator(device=DEV).manual_seed(0)
    B = 2
    raw = torch.rand(B, 50, 37, 6, generator=g, device=DEV) * 255
    f2 = torch.randn(B, 13, 10, 512, generator=g, device=DEV).bfloat16()
    f3 = torch.randn(B, 7, 5, 1024, generator=g, device=DEV).bfloat16()
    fpn = m.necks[view].fpn
    g3 = features.lateral_forward(f3, nv.lat_w[3], nv.lat_b[3], None)
    g2 = features.lateral_forward(f2, nv.lat_w[2], nv.lat_b[2], g3)
    with torch.no_grad():
        i3 = fpn.inner_blocks[3](f3.float().movedim(-1, 1))
        i2 = fpn.inner_blocks[2](f2.float().movedim(-1, 1)) + F.interpolate(i3, size=(13, 10), mode="nearest")
        # the 3x3 / raw-level stages are fp32: check them tightly on the kernel's own inner map
        i2k = g2.movedim(-1, 1)
        i0 = fpn.inner_blocks[0](raw.movedim(-1, 1)) + F.interpolate(i2k, size=(50, 37), mode="nearest")
        o2 = m.embeddings[view].embedding_layers["embedding2"](fpn.layer_blocks[2](i2k).movedim(1, -1).contiguous())
        o0 = m.embeddings[view].embedding_layers["embedding0"](fpn.layer_blocks[0](i0).movedim(1, -1).contiguous())
    assert _rel(g3, i3.movedim(1, -1)) < 5e-3          # lateral weights are rounded to bf16 for the tensor cores
    assert _rel(g2, i2.movedim(1, -1)) < 5e-3
    S = 13 * 10 + 50 * 37
    # impl 1 = fp32 CUDA-core kernel (tight), impl 2 = tcgen05 row-strip kernel with an f16 inner tile (f16 rounding)
    for impl, tol in ((1, 1e-4), (2, 2e-3)):
        pyr = torch.zeros(B, S, 16, device=DEV)
        py, px = nv._tables(2, 13, 10)
        features.fpn_output_forward(pyr, 50 * 37, 13, 10, nv.out_w[2], nv.out_b[2], py, px, inner=g2, impl=impl)
        py, px = nv._tables(0, 50, 37)
        features.fpn_output_forward(pyr, 0, 50, 37, nv.out_w[0], nv.out_b[0], py, px, raw=raw, lat_w=nv.lat_w[0],
                                    lat_b=nv.lat_b[0], coarse=g2, impl=impl)
        assert _rel(pyr[:, 50 * 37:], o2.flatten(1, 2)) < tol, (impl, _rel(pyr[:, 50 * 37:], o2.flatten(1, 2)))
        assert _rel(pyr[:, :50 * 37], o0.flatten(1, 2)) < tol, (impl, _rel(pyr[:, :50 * 37], o0.flatten(1, 2)))


@pytest.mark.parametrize("cin,H,W", [(3, 45, 300), (6, 17, 129), (0, 23, 256)])
def test_fpn_output_tensor_core_matches_cuda_core(cin, H, W):
    """Wide levels: the tcgen05 row-strip kernel against the fp32 CUDA-core kernel on the same inputs."""
    from dpft_b200 import features
    g = torch.Generator(device=DEV).manual_seed(cin + H)
    B = 2
    w = torch.randn(3, 3, 16, 16, generator=g, device=DEV) * 0.1
    bias = torch.randn(16, generator=g, device=DEV)
    py = torch.randn(H, 16, generator=g, device=DEV)
    px = torch.randn(W, 16, generator=g, device=DEV)
    coarse = torch.randn(B, (H + 3) // 4, (W + 3) // 4, 16, generator=g, device=DEV)
    kw = {}
    if cin:
        kw = dict(raw=torch.rand(B, H, W, cin, generator=g, device=DEV) * 255, coarse=coarse,
                  lat_w=torch.randn(16, cin, generator=g, device=DEV) * 0.01, lat_b=torch.randn(16, generator=g, device=DEV))
    else:
        kw = dict(inner=torch.randn(B, H, W, 16, generator=g, device=DEV) * 3)
    S = H * W + 7
    a = torch.zeros(B, S, 16, device=DEV)
    b = torch.zeros(B, S, 16, device=DEV)
    features.fpn_output_forward(a, 7, H, W, w, bias, py, px, impl=1, **kw)
    features.fpn_output_forward(b, 7, H, W, w, bias, py, px, impl=2, **kw)
    torch.cuda.synchronize()
    assert float(b[:, :7].abs().max()) == 0.0                      # nothing written before `start`
    assert _rel(b, a) < 2e-3, _rel(b, a)



@pytest.mark.parametrize("cin,H,W", [(3, 45, 300), (6, 17, 129), (6, 37, 107), (3, 90, 160), (6, 64, 256)])
def test_fpn_output_column_builder_matches_cuda_core(cin, H, W):
    """impl 3 (fpn_output_tc2_kernel: column-owning tile builder, staged coarse patch) against the fp32 CUDA-core kernel
    and the validated tensor-core kernel on the same inputs."""
    from dpft_b200 import features
    g = torch.Generator(device=DEV).manual_seed(cin + H)
    B = 2
    w = torch.randn(3, 3, 16, 16, generator=g, device=DEV) * 0.1
    bias = torch.randn(16, generator=g, device=DEV)
    py = torch.randn(H, 16, generator=g, device=DEV)
    px = torch.randn(W, 16, generator=g, device=DEV)
    coarse = torch.randn(B, (H + 3) // 4, (W + 3) // 4, 16, generator=g, device=DEV)
    kw = dict(raw=torch.rand(B, H, W, cin, generator=g, device=DEV) * 255, coarse=coarse,
              lat_w=torch.randn(16, cin, generator=g, device=DEV) * 0.01, lat_b=torch.randn(16, generator=g, device=DEV))
    S = H * W + 7
    outs = []
    for impl in (1, 2, 3):
        o = torch.zeros(B, S, 16, device=DEV)
        features.fpn_output_forward(o, 7, H, W, w, bias, py, px, impl=impl, **kw)
        outs.append(o)
    torch.cuda.synchronize()
    a, b, c = outs
    assert float(c[:, :7].abs().max()) == 0.0                      # nothing written before `start`
    assert _rel(c, a) < 2e-3, _rel(c, a)
    assert _rel(c, b) < 1e-3, _rel(c, b)                           # same operand image up to the rounding of the FMA chain
    # the lateral weights as kernel parameters (constant operands of the builder's FMAs) instead of shared-memory loads:
    # the same FMA chain on the same values
    d = torch.zeros(B, S, 16, device=DEV)
    features.fpn_output_forward(d, 7, H, W, w, bias, py, px, impl=3, lat_w_host=kw["lat_w"].cpu().contiguous(), **kw)
    torch.cuda.synchronize()
    assert torch.equal(d, c)


@pytest.mark.parametrize("size", [(2, 70, 131, 3), (1, 128, 228, 3), (3, 33, 300, 3)])
def test_uint8_frames_are_read_in_place_of_float32(size):
    """DPFT_RAW_U8: the camera view handed over as the uint8 frames an image decoder produces (before the reference's
    `.type(float32)`, dataset.py read_image).  The stem's row-streaming kernel and the FPN raw-level kernels convert on load;
    0..255 is exact in fp32 / f16, so every result must be BIT-IDENTICAL to the float32 path on the same values."""
    from dpft_b200 import features
    from dpft_b200.features import NativeView
    cfg, m = _model("kradar_camera_mono")
    nv = NativeView(m.backbones["camera_mono"], m.necks["camera_mono"], m.embeddings["camera_mono"], True, DEV,
                    torch.float16, torch.float16)
    g = torch.Generator(device=DEV).manual_seed(size[1])
    x8 = torch.randint(0, 256, size, generator=g, device=DEV, dtype=torch.uint8)
    xf = x8.float()
    assert nv.accepts_uint8(x8)
    for dt in (torch.float16, torch.bfloat16):
        a = features.stem_forward(x8, nv.stem_w, nv.stem_b, dt, w_packed=nv.stem_w_packed)
        b = features.stem_forward(xf, nv.stem_w, nv.stem_b, dt, w_packed=nv.stem_w_packed)
        assert torch.equal(a, b), dt
    with torch.no_grad():
        p8, shapes8 = nv.pyramid(x8)                     # stem, stages, lateral chain, every FPN output kernel incl. the raw level
        pf, shapesf = nv.pyramid(xf)
    torch.cuda.synchronize()
    assert shapes8 == shapesf and torch.equal(p8, pf)
    narrow = torch.randint(0, 256, (1, 40, 90, 3), generator=g, device=DEV, dtype=torch.uint8)     # below the streaming kernel's width
    assert not nv.accepts_uint8(narrow)
    with torch.no_grad():
        assert torch.equal(nv.pyramid(narrow)[0], nv.pyramid(narrow.float())[0])                  # converted first (plumbing)
