"""north_star: "drops into dprt.train/dprt.evaluate unchanged".  The reference's entry point imports
``from dprt.models import build/load`` (src/dprt/train.py:7-8); both drop-in routes of dpft_b200/dropin must make that
import resolve to dpft_b200.models while everything else of ``dprt`` still comes from the reference.  Runs in a fresh
interpreter against the reference source tree (build container) or its installed copy (baseline/_ref); the reference's
trainer stack needs tensorboard / deepspeed / pytorch3d, absent here, so those imports are stubbed and ``main`` is not run —
only the module-level binding of train.py is checked, plus a model built through the bound name."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

import reference_shim

CHECK = r"""
import sys, types
for name in ("pypcd", "pypcd.pypcd", "pytorch3d", "pytorch3d.ops", "deepspeed", "deepspeed.accelerator", "deepspeed.profiling",
             "deepspeed.profiling.flops_profiler", "open3d", "torch.utils.tensorboard"):
    m = types.ModuleType(name); sys.modules[name] = m
sys.modules["pytorch3d.ops"].box3d_overlap = lambda *a, **k: None
sys.modules["torch.utils.tensorboard"].SummaryWriter = object
sys.modules["deepspeed.accelerator"].get_accelerator = lambda: None
sys.modules["deepspeed.profiling.flops_profiler"].get_model_profile = lambda *a, **k: None
MODE
import dprt.train as train                       # the reference's own entry module, unmodified
import dpft_b200.models as ours
assert train.build_model is ours.build and train.load_model is ours.load, (train.build_model, train.load_model)
assert "REFDIR" in train.__file__, train.__file__          # train.py itself is the reference's file
import dprt.datasets, dprt.utils.config                # the rest of the package still resolves to the reference
assert "REFDIR" in dprt.utils.config.__file__
from dpft_b200 import configs, synthetic
model = train.build_model("dprt", synthetic.offline_config(configs.make_config("kradar_radar_bev")))
assert type(model).__module__.startswith("dpft_b200.models"), type(model)
print("OK", type(model).__module__)
"""


@pytest.mark.parametrize("mode", ["path", "hook"])
def test_reference_entry_point_binds_dpft_b200_models(mode):
    if not reference_shim.available():
        pytest.skip("no copy of the reference package on this machine")
    ref_src = reference_shim.source()
    overlay = os.path.join(ROOT, "dpft_b200", "dropin")
    env = dict(os.environ)
    if mode == "path":
        env["PYTHONPATH"] = os.pathsep.join([overlay, ref_src, ROOT])
        prelude = ""
    else:
        env["PYTHONPATH"] = os.pathsep.join([ref_src, ROOT])
        prelude = "import dpft_b200.dropin; dpft_b200.dropin.install(msda_plugin=False)"
    code = CHECK.replace("MODE", prelude).replace("REFDIR", ref_src)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK dpft_b200.models" in r.stdout, r.stdout + r.stderr
