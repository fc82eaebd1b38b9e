"""Checkpoint interchange (dpft_b200/checkpoint.py): weights-only state files, and reading a reference
whole-module pickle (src/dprt/training/trainer.py:258) without the reference package."""
import os

import pytest
import torch

from dpft_b200 import checkpoint, configs, models, synthetic


def _cfg():
    return synthetic.offline_config(configs.make_config("kradar_radar_bev"), n_queries=(6, 5, 1))


def test_state_file_round_trip_is_weights_only(tmp_path):
    cfg = _cfg()
    m = models.build("dprt", cfg)
    m.load_state_dict(synthetic.seeded_state_dict(m.state_dict(), seed=3))
    path = str(tmp_path / "20260101-000000_checkpoint_0007.pt")
    checkpoint.save_state(m, path, cfg, epoch=7, timestamp="20260101-000000")
    torch.load(path, weights_only=True)                                   # no pickled code in the file
    m2, epoch, ts = checkpoint.load_state(path)
    assert (epoch, ts) == (7, "20260101-000000")
    sd, sd2 = m.state_dict(), m2.state_dict()
    assert list(sd) == list(sd2) and all(torch.equal(sd[k], sd2[k]) for k in sd)
    m3, epoch3, ts3 = models.load(path)                                   # the reference-shaped entry point takes it too
    assert (epoch3, ts3) == (7, "20260101-000000") and torch.equal(m3.state_dict()["fuser.query"], sd["fuser.query"])


def test_whole_module_pickle_still_loads_like_the_reference(tmp_path):
    m = models.build("dprt", _cfg())
    path = str(tmp_path / "20260101-000000_checkpoint_0001.pt")
    torch.save(m, path)                                                   # what trainer.py:258 does
    m2, epoch, ts = models.load(path)
    assert isinstance(m2, type(m)) and epoch == 1 and ts == "20260101-000000"


def test_reference_pickle_is_read_without_the_reference_package(tmp_path, reference_models):
    """A checkpoint written by the UNMODIFIED reference; read back with every ``dprt.*`` and ``torchvision.*`` class forced
    to a stand-in (as on a machine where neither is importable)."""
    cfg = _cfg()
    torch.manual_seed(0)
    ref = reference_models.build("dprt", cfg)
    path = str(tmp_path / "20260101-000000_checkpoint_0003.pt")
    torch.save(ref, path)
    got = checkpoint.reference_state_dict(path, force_standin=("dprt", "torchvision", "MultiScaleDeformableAttention"))
    want = ref.state_dict()
    assert list(got) == list(want)
    assert all(torch.equal(got[k], want[k]) for k in want)
    out = str(tmp_path / "converted.pt")
    ours = checkpoint.convert_reference_checkpoint(path, cfg, out_path=out)
    assert all(torch.equal(ours.state_dict()[k], want[k]) for k in want)
    back, epoch, _ = checkpoint.load_state(out)
    assert epoch == 3 and torch.equal(back.state_dict()["fuser.query"], want["fuser.query"])
    batch = synthetic.synthetic_batch(cfg, 1, seed=2, sizes={"radar_bev": (64, 48, 6)})
    from helpers import oracle_op_injected
    with torch.no_grad(), oracle_op_injected():
        a = ref.eval()(batch)
        b = ours.eval()(batch)
    for k in a:
        assert torch.allclose(a[k], b[k], rtol=1e-4, atol=1e-5), k
