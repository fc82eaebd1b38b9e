"""Drop-in name for the reference's model package: ``python -m dprt.train`` / ``dprt.evaluate`` pick up ``dpft_b200.models``
with NO edit to the reference (its import sites: src/dprt/train.py:7-8, src/dprt/evaluation/evaluator.py:14).

Two ways, both tested in tests/test_dropin.py:

* path:  put ``<site>/dpft_b200/dropin`` in front of the reference's ``src`` on ``PYTHONPATH``.  The ``dprt`` package found
  there is an overlay: its ``__path__`` continues into every other ``dprt`` directory on ``sys.path``, so ``dprt.train``,
  ``dprt.datasets``, ``dprt.training``, ``dprt.evaluation``, ``dprt.utils`` still come from the reference, and only
  ``dprt.models`` resolves here;
* hook:  ``import dpft_b200.dropin; dpft_b200.dropin.install()`` before anything imports ``dprt.models`` (registers
  ``dpft_b200.models`` as ``sys.modules['dprt.models']`` and the deformable-attention kernels under the name the
  reference's own layers import, ``MultiScaleDeformableAttention``).
"""
import os
import sys

OVERLAY = os.path.dirname(os.path.abspath(__file__))       # the directory to put on PYTHONPATH


def install(msda_plugin: bool = True) -> None:
    import dpft_b200.models as ours
    sys.modules["dprt.models"] = ours
    pkg = sys.modules.get("dprt")
    if pkg is not None:
        setattr(pkg, "models", ours)
    if msda_plugin and "MultiScaleDeformableAttention" not in sys.modules:
        from dpft_b200 import msda
        msda.install_plugin()
