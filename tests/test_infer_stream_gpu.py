"""DPRT.infer_stream (FusedEngine.stream): pipelined replay of consecutive forwards must give exactly what the one-at-a-time
forward gives, in order, for device and pinned-host batches, any depth, and leave grad mode alone between items."""
import pytest
import torch

from conftest import load_golden
from helpers import case_setup
from dpft_b200 import models, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup():
    rec = load_golden("fusion_small_300q")
    cfg, _ = case_setup(rec)
    model = models.build("dprt", cfg).eval()
    model.load_state_dict(synthetic.seeded_state_dict(model.state_dict(), seed=rec["weight_seed"]), strict=True)
    model = model.to(DEV)
    batches = [synthetic.synthetic_batch(cfg, rec["case"]["batch"], seed=100 + i, sizes=rec["case"]["sizes"]) for i in range(5)]
    return model, batches


@pytest.mark.parametrize("depth", [1, 2, 3])
@pytest.mark.parametrize("host", [False, True])
def test_stream_equals_sequential_forward(depth, host):
    model, batches = _setup()
    dev_batches = [{k: v.to(DEV) for k, v in b.items()} for b in batches]
    with torch.no_grad():
        model.use_cuda_graph = False
        want = [model(b) for b in dev_batches]
        model.use_cuda_graph = True
    feed = [{k: v.pin_memory() for k, v in b.items()} for b in batches] if host else dev_batches
    order = [0, 1, 2, 3, 4, 2, 0, 4, 1]                         # more items than slots, repeats, both parities
    got = []
    for out in model.infer_stream((feed[i] for i in order), depth=depth):
        assert torch.is_grad_enabled()                           # the generator does not leak no_grad into the caller
        got.append({k: v.clone() for k, v in out.items()})
    torch.cuda.synchronize()
    assert len(got) == len(order)
    for i, out in zip(order, got):
        assert list(out) == ["center", "size", "angle", "class"]
        for k in out:
            assert torch.allclose(out[k], want[i][k], rtol=1e-5, atol=1e-5), (i, k)
    # host batches get spare slots (FusedEngine.upload_slots) so that an upload never waits for the forward that last read its buffer
    n_slots = depth + (model._engine.upload_slots if host else 0)
    assert len(model._engine._pipelines) == 1 and len(next(iter(model._engine._pipelines.values()))) == n_slots


def test_stream_rejects_changing_shapes_and_handles_empty_input():
    model, batches = _setup()
    assert list(model.infer_stream(iter(()))) == []
    b0 = {k: v.to(DEV) for k, v in batches[0].items()}
    b1 = {k: v[:1].to(DEV) for k, v in batches[1].items()}
    with pytest.raises(RuntimeError, match="same shapes"):
        list(model.infer_stream([b0, b1]))


def test_stream_falls_back_to_sequential_without_the_graph():
    model, batches = _setup()
    model.use_cuda_graph = False
    dev_batches = [{k: v.to(DEV) for k, v in b.items()} for b in batches[:2]]
    with torch.no_grad():
        want = [model(b) for b in dev_batches]
    got = list(model.infer_stream(dev_batches))
    for a, b in zip(got, want):
        for k in a:
            assert torch.equal(a[k], b[k])


def test_feeder_on_the_gpu_matches_the_cpu_and_feeds_the_stream():
    """dpft_b200.feeder on cuda (uint8 frames uploaded on the copy stream, arithmetic on the GPU) against the same feeder on
    the CPU (bit-exact against the reference dataset methods, tests/test_feeder.py), then DPRT.infer_stream fed by it."""
    from dpft_b200 import configs, feeder
    cfg = synthetic.offline_config(configs.make_config("kradar"), n_queries=(20, 15, 1))
    sizes = {"camera_mono": (96, 160, 3), "radar_bev": (64, 48, 6), "radar_front": (37, 48, 6)}
    inputs = cfg["model"]["inputs"]
    raws = [feeder.synthetic_raw_batch(inputs, 2, seed=50 + i, sizes=sizes, pin=True) for i in range(3)]
    cpu = feeder.BatchFeeder(inputs, image_size=64, device="cpu")
    gpu = feeder.BatchFeeder(inputs, image_size=64, device="cuda:0")
    for raw, got in zip(raws, gpu.stream(iter(raws))):
        want = cpu.prepare(raw)
        for k, w in want.items():
            assert got[k].shape == w.shape and got[k].dtype == w.dtype, k
            # the library resize differs in rounding between CPU and CUDA; the radar power scaling (log10 / divisions) differs
            # by one fp32 ulp at the 255 end of the range between the CPU and the CUDA math libraries (measured: 1.5e-5)
            tol = 1e-3 if k == "camera_mono" else 2 * 2.0 ** -23
            assert float((got[k].cpu().double() - w.double()).abs().max()) <= tol * 255, k
    model = models.build("dprt", cfg).eval()
    model.load_state_dict(synthetic.seeded_state_dict(model.state_dict(), seed=3))
    model = model.to("cuda:0")
    plain = feeder.BatchFeeder(inputs, image_size=None, device="cuda:0")
    with torch.no_grad():
        want = [model(plain.prepare(r)) for r in raws]
        got = list(model.infer_stream(plain.stream(iter(raws)), depth=2))
    assert len(got) == 3
    for g, w in zip(got, want):
        for k in w:
            assert torch.allclose(g[k], w[k], rtol=1e-5, atol=1e-5), k


def test_uint8_camera_frames_through_the_model_api():
    """The full fusion model fed with uint8 camera frames (device tensors, pinned host tensors, and through infer_stream) gives
    exactly the outputs of the same frames as float32 — the e2e leg of bench.py uploads a quarter of the camera bytes."""
    from dpft_b200 import configs
    cfg = synthetic.offline_config(configs.make_config("kradar"), n_queries=(20, 15, 1))
    sizes = {"camera_mono": (96, 160, 3), "radar_bev": (64, 48, 6), "radar_front": (37, 48, 6)}
    model = models.build("dprt", cfg).eval()
    model.load_state_dict(synthetic.seeded_state_dict(model.state_dict(), seed=3))
    model = model.to("cuda:0")
    batches = [synthetic.synthetic_batch(cfg, 2, seed=70 + i, sizes=sizes) for i in range(4)]
    for b in batches:
        b["camera_mono"] = b["camera_mono"].round().clamp(0, 255)
    as_u8 = [{k: (v.to(torch.uint8) if k == "camera_mono" else v) for k, v in b.items()} for b in batches]
    with torch.no_grad():
        want = [model({k: v.to("cuda:0") for k, v in b.items()}) for b in batches]
        want = [{k: v.clone() for k, v in o.items()} for o in want]
        got_dev = [model({k: v.to("cuda:0") for k, v in b.items()}) for b in as_u8]          # eager, then captured graphs
        got_dev = [{k: v.clone() for k, v in o.items()} for o in got_dev]
        got_host = [{k: v.clone() for k, v in model({k: v.pin_memory() for k, v in b.items()}).items()} for b in as_u8]
    got_stream = list(model.infer_stream([{k: v.pin_memory() for k, v in b.items()} for b in as_u8], depth=2))
    torch.cuda.synchronize()
    for w, a, b, c in zip(want, got_dev, got_host, got_stream):
        for k in w:
            assert torch.equal(a[k], w[k]) and torch.equal(b[k], w[k]) and torch.equal(c[k], w[k]), k
    # outside the native feature path the frames are converted first (plumbing): still the same model
    model.native_features = False
    with torch.no_grad():
        f32 = model({k: v.to("cuda:0") for k, v in batches[0].items()})
        u8 = model({k: v.to("cuda:0") for k, v in as_u8[0].items()})
    for k in f32:
        assert torch.equal(f32[k], u8[k]), k


def test_paired_side_views_give_identical_outputs():
    """The two radar views of the fusion config (same ResNet-50, equally shaped inputs) share the launches of their Bottleneck
    convolutions (conv2d_nhwc_pair): outputs must be bit-identical to the unpaired engine — eager, serial, graph and stream."""
    from dpft_b200 import configs
    cfg = synthetic.offline_config(configs.make_config("kradar"), n_queries=(20, 15, 1))
    sizes = {"camera_mono": (96, 160, 3), "radar_bev": (64, 64, 6), "radar_front": (64, 64, 6)}
    model = models.build("dprt", cfg).eval()
    model.load_state_dict(synthetic.seeded_state_dict(model.state_dict(), seed=3))
    model = model.to("cuda:0")
    batches = [{k: v.to("cuda:0") for k, v in synthetic.synthetic_batch(cfg, 2, seed=80 + i, sizes=sizes).items()} for i in range(3)]
    outs = {}
    for paired in (False, True):
        model.pair_side_views = paired
        with torch.no_grad():
            model.parallel_views = False
            serial = {k: v.clone() for k, v in model(batches[0]).items()}
            model.parallel_views = True
            first = {k: v.clone() for k, v in model(batches[0]).items()}             # eager (first sighting after the switch)
            replay = [{k: v.clone() for k, v in model(b).items()} for b in batches]   # captured graphs
        streamed = [{k: v.clone() for k, v in o.items()} for o in model.infer_stream(batches, depth=2)]
        outs[paired] = (serial, first, replay, streamed)
    eng = model._engine
    assert eng._paired_side_views(batches[0], [0, 1, 2]) == (1, 2)
    for a, b in zip(outs[False], outs[True]):
        for x, y in zip(a if isinstance(a, list) else [a], b if isinstance(b, list) else [b]):
            for k in x:
                assert torch.equal(x[k], y[k]), k
    # unequal radar sizes: no pairing, same code path as before
    other = {k: v.to("cuda:0") for k, v in synthetic.synthetic_batch(cfg, 2, seed=90, sizes={**sizes, "radar_front": (37, 64, 6)}).items()}
    assert eng._paired_side_views(other, [0, 1, 2]) is None
