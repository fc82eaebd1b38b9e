"""Imports the UNMODIFIED reference package from /root/reference in the build container.

The reference's one compiled dependency (``MultiScaleDeformableAttention``) is absent, so a stand-in module
backed by the CPU oracle (oracle/msda.py) is registered under that name first.  Used only by
tools/make_golden.py and the in-container parity tests; /root/reference does not exist on the GPU box.
"""
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "dprt"))


def import_reference_models():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import msda as O

    shim = types.ModuleType("MultiScaleDeformableAttention")

    def fwd(value, shapes, lsi, loc, attn, im2col_step):
        return O.msda_forward_torch(value, shapes, loc, attn)

    def bwd(value, shapes, lsi, loc, attn, grad_out, im2col_step):
        return O.msda_backward_torch(value, shapes, loc, attn, grad_out)

    shim.ms_deform_attn_forward = fwd
    shim.ms_deform_attn_backward = bwd
    sys.modules["MultiScaleDeformableAttention"] = shim
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import dprt.models as ref_models  # noqa: E402
    return ref_models
