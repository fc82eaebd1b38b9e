"""Host-side pieces added in round 2 that need no GPU: bench.py's parity report, the reference installer / locator, the
persistent-grid budget context, the drop-in hook."""
import os
import sys

import torch

from conftest import ROOT

import bench
import reference_shim


def test_parity_report_metrics():
    want = {"center": torch.tensor([[[10.0, -20.0, 0.0]]]), "size": torch.tensor([[[1.0, 2.0, 4.0]]])}
    got = {"center": want["center"] + torch.tensor([[[0.1, 0.0, 0.02]]]), "size": want["size"] * 1.01}
    grid = torch.tensor([[9.0, -19.0, 0.0]])
    rep = bench.parity_report(got, want, grid)
    assert set(rep) == {"center", "size", "center_refinement"}
    assert abs(rep["center"]["max_norm_rel"] - 0.1 / 20.0) < 1e-6                     # max |d| / max |want|
    assert abs(rep["size"]["max_norm_rel"] - 0.01) < 1e-6 and abs(rep["size"]["elem_rel_max"] - 0.01) < 1e-5
    # the refinement (centre minus the static grid) is what the network computes: its scale is 1, not 20
    assert abs(rep["center_refinement"]["max_abs_want"] - 1.0) < 1e-6
    assert abs(rep["center_refinement"]["max_norm_rel"] - 0.1) < 1e-6
    # per-element error uses a floor of 1e-3 of the output's scale, so exact zeros do not divide by zero
    assert rep["center"]["elem_rel_max"] <= 0.02 / (1e-3 * 20.0) + 1e-6


def test_reference_copy_is_installed_and_importable_from_baseline_ref():
    """__graft_entry__.build() installs the unmodified reference into baseline/_ref (it travels to the GPU box); bench.py and
    the GPU tests import it from there."""
    marker = os.path.join(ROOT, "baseline", "_ref", "INSTALL.json")
    if not os.path.exists(marker):
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import install_reference
        info = install_reference.install()
        assert info.get("status") == "installed", info
    assert os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "dprt", "models", "dprt.py"))
    assert os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "dprt", "models", "layers", "ms_deform_attn.py"))
    assert reference_shim.available()
    # nothing of it is tracked by git (the copy detector and the history stay free of reference sources)
    with open(os.path.join(ROOT, ".gitignore")) as f:
        assert "baseline/_ref/" in f.read()


def test_cta_budget_nests_and_restores():
    from dpft_b200 import conv
    assert conv._MAX_CTAS == 0
    with conv.cta_budget(48):
        assert conv._MAX_CTAS == 48
        with conv.cta_budget(0):
            assert conv._MAX_CTAS == 0
        assert conv._MAX_CTAS == 48
    assert conv._MAX_CTAS == 0


def test_dropin_hook_registers_dpft_models():
    import dpft_b200.dropin
    import dpft_b200.models as ours
    saved = {k: sys.modules.get(k) for k in ("dprt.models", "MultiScaleDeformableAttention")}
    try:
        dpft_b200.dropin.install(msda_plugin=False)
        assert sys.modules["dprt.models"] is ours
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
