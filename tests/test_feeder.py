"""dpft_b200.feeder (the data path between the decoded files and DPRT.forward) against the reference's own
dataset methods: a committed fixture (tools/make_golden_feeder.py) on any machine, the live reference where it exists."""
import pytest
import torch

from conftest import load_golden
from dpft_b200 import configs, feeder, models, synthetic


def _check(got, want):
    assert sorted(got) == sorted(want)
    for k, w in want.items():
        assert got[k].shape == w.shape and got[k].dtype == w.dtype, k
        assert torch.equal(got[k], w), (k, float((got[k].double() - w.double()).abs().max()))   # same ops in the same order: exact


@pytest.mark.parametrize("i", [0, 1, 2])
def test_feeder_matches_reference_dataset_fixture(i):
    rec = load_golden("feeder_small")["cases"][i]
    case = rec["case"]
    raw = feeder.synthetic_raw_batch(case["inputs"], case["batch"], seed=case["seed"], sizes=case["sizes"])
    f = feeder.BatchFeeder(case["inputs"], image_size=case["image_size"], scale=case["scale"], device="cpu")
    _check(f.prepare(raw), rec["batch"])


def test_feeder_matches_live_reference(reference_models):
    import reference_shim
    from make_golden_feeder import reference_batch
    ds = reference_shim.import_reference_dataset()
    case = dict(inputs=["camera_mono", "radar_front"], batch=2, seed=7, image_size=16, scale=True,
                sizes={"camera_mono": (40, 52, 3), "radar_front": (37, 107, 6)})
    raw = feeder.synthetic_raw_batch(case["inputs"], case["batch"], seed=case["seed"], sizes=case["sizes"])
    f = feeder.BatchFeeder(case["inputs"], image_size=16, device="cpu")
    _check(f.prepare(raw), reference_batch(ds, case))


def test_radar_projections_equal_the_synthetic_contract():
    """The raster projections the feeder attaches are the ones dpft_b200.synthetic (and the golden cases) use at native sizes."""
    cfg = configs.make_config("kradar_radar")
    batch = synthetic.synthetic_batch(cfg, 1, seed=0)
    for view in ("radar_bev", "radar_front"):
        assert torch.allclose(feeder.radar_projection(view), batch[f"label_to_{view}_p"][0], rtol=0, atol=1e-6)


def test_half_precision_cubes_and_uint8_frames_are_widened_before_the_arithmetic():
    inputs = ["camera_mono", "radar_bev"]
    sizes = {"camera_mono": (24, 32, 3), "radar_bev": (8, 107, 6)}
    raw = feeder.synthetic_raw_batch(inputs, 2, seed=1, sizes=sizes)
    f = feeder.BatchFeeder(inputs, image_size=None, device="cpu")
    full = f.prepare(raw)
    assert full["camera_mono"].dtype == torch.float32 and torch.equal(full["camera_mono"], raw["camera_mono"].float())
    assert float(full["radar_bev"].min()) == 0.0 and float(full["radar_bev"].max()) == 255.0       # clip exercised
    half = f.prepare({**raw, "radar_bev": raw["radar_bev"].half()})
    want = torch.clip((raw["radar_bev"].half().float() - 100.0) / 100.0 * 255.0, 0, 255)
    assert torch.equal(half["radar_bev"], want)


def test_stream_yields_every_batch_in_order():
    inputs = ["radar_bev"]
    f = feeder.BatchFeeder(inputs, device="cpu")
    raws = [feeder.synthetic_raw_batch(inputs, 1, seed=s, sizes={"radar_bev": (8, 107, 6)}) for s in range(4)]
    outs = list(f.stream(iter(raws)))
    assert len(outs) == 4 and list(f.stream(iter([]))) == []
    for o, r in zip(outs, raws):
        assert torch.equal(o["radar_bev"], f.prepare(r)["radar_bev"])


def test_from_config_reads_the_reference_data_section_and_feeds_the_model():
    cfg = synthetic.offline_config(configs.make_config("kradar_radar_bev"))
    cfg["data"] = {"image_size": 512, "scale": True}
    f = feeder.BatchFeeder.from_config(cfg, device="cpu")
    assert f.inputs == ["radar_bev"] and f.image_size == 512
    batch = f.prepare(feeder.synthetic_raw_batch(f.inputs, 1, seed=2, sizes={"radar_bev": (32, 107, 6)}))
    model = models.build("dprt", cfg).eval()
    assert {"radar_bev", "radar_bev_shape", "label_to_radar_bev_t", "label_to_radar_bev_p"} == set(batch)
    from helpers import oracle_op_injected
    with oracle_op_injected(), torch.no_grad():
        out = model(batch)
    assert list(out) == ["center", "size", "angle", "class"] and all(torch.isfinite(v).all() for v in out.values())
