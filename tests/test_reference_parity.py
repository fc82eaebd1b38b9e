"""In-container checks against the UNMODIFIED reference package (skipped where /root/reference is absent)."""
import json
import os

import pytest
import torch

from helpers import oracle_op_injected, rel_err
from dpft_b200 import configs, models, synthetic
from oracle import dprt_oracle

SMALL = {"camera_mono": (64, 96, 3), "radar_bev": (64, 40, 6), "radar_front": (37, 40, 6)}


@pytest.mark.parametrize("name", list(configs.VIEWS))
def test_generated_configs_equal_the_shipped_json(reference_models, name):
    with open(f"/root/reference/config/{name}.json") as f:
        ref = json.load(f)
    mine = configs.make_config(name)
    assert mine["model"] == ref["model"] and mine["computing"] == ref["computing"]


@pytest.mark.parametrize("name", ["kradar_radar", "kradar"])
def test_eval_forward_three_way(reference_models, name):
    cfg = synthetic.offline_config(configs.make_config(name))
    ref = reference_models.build("dprt", cfg).eval()
    sd = synthetic.seeded_state_dict(ref.state_dict(), seed=11)
    ref.load_state_dict(sd, strict=True)
    mine = models.build("dprt", cfg).eval()
    mine.load_state_dict(sd, strict=True)                 # same names, same shapes
    batch = synthetic.synthetic_batch(cfg, 2, seed=12, sizes=SMALL)
    with torch.no_grad():
        want = ref({k: v.clone() for k, v in batch.items()})
        got_oracle = dprt_oracle.forward(sd, cfg, batch)
        with oracle_op_injected():
            got_mine = mine(batch)
    for k in want:
        assert rel_err(got_oracle[k], want[k]) < 5e-4, k  # fp32 end-to-end; north_star bar is 1e-3 rel
        assert rel_err(got_mine[k], want[k]) < 5e-4, k  # fp32 end-to-end; north_star bar is 1e-3 rel


def test_train_step_gradients_match(reference_models):
    cfg = synthetic.offline_config(configs.make_config("kradar_radar"), dropout=0.0)
    ref = reference_models.build("dprt", cfg).train()
    sd = synthetic.seeded_state_dict(ref.state_dict(), seed=21)
    ref.load_state_dict(sd)
    mine = models.build("dprt", cfg).train()
    mine.load_state_dict(sd)
    batch = synthetic.synthetic_batch(cfg, 2, seed=22, sizes=SMALL)
    sum((v ** 2).mean() for v in ref({k: v.clone() for k, v in batch.items()}).values()).backward()
    with oracle_op_injected():
        sum((v ** 2).mean() for v in mine(batch).values()).backward()
    g_ref = {k: p.grad for k, p in ref.named_parameters()}
    g_mine = {k: p.grad for k, p in mine.named_parameters()}
    missing = [k for k in g_ref if g_ref[k] is None]
    assert missing == [k for k in g_mine if g_mine[k] is None] and len(missing) == 39
    for k, g in g_ref.items():
        if g is not None:
            assert rel_err(g_mine[k], g) < 2e-2, k
    # BatchNorm running statistics updated identically
    for k, v in ref.state_dict().items():
        if "running_" in k:
            assert torch.allclose(mine.state_dict()[k], v, rtol=1e-5, atol=1e-6), k
