import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a machine without CUDA, and fail loudly — never skip — on a CUDA machine whose
    in-tree library is missing, so that a green GPU run always means the native code ran."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run on the B200 box with `-m gpu`)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def reference_models():
    """The unmodified reference package, importable only in the build container."""
    import reference_shim
    if not reference_shim.available():
        pytest.skip("/root/reference is not present on this machine")
    return reference_shim.import_reference_models()


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
