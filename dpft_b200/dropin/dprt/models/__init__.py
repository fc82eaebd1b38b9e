"""``dprt.models`` served by dpft_b200 (reference src/dprt/models/__init__.py:10-18: ``build(model, config)``,
``load(checkpoint)``)."""
from dpft_b200.models import DPRT, build, build_dprt, load  # noqa: F401
