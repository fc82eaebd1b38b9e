"""Imports the UNMODIFIED reference package: from /root/reference in the build container, or from the copy that
``__graft_entry__.build()`` installs into the git-ignored ``baseline/_ref`` (it travels to the GPU box with gpurun).

The reference's one compiled dependency (``MultiScaleDeformableAttention``) is absent from both, so a stand-in module is
registered under that name first: CUDA tensors go to this repo's kernels (``dpft_b200.msda``, the plugin boundary of
INTEGRATION.md §1), CPU tensors to the CPU oracle (oracle/msda.py).  Used by tools/make_golden*.py, the parity tests and
bench.py's reference / library-baseline legs — never by the product path.
"""
import os
import sys
import types

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = ("/root/reference/src", os.path.join(_ROOT, "baseline", "_ref"))


def _find_src():
    override = os.environ.get("DPFT_REFERENCE_SRC")          # tests: force the installed copy in the build container
    for c in ((override,) if override else _CANDIDATES):
        if os.path.isfile(os.path.join(c, "dprt", "models", "__init__.py")):
            return c
    return _CANDIDATES[0]


REFERENCE_SRC = _find_src()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "dprt", "models", "__init__.py"))


def source() -> str:
    """Which copy of the reference package gets imported (for bench.py's `cpu_baseline.sample`)."""
    return REFERENCE_SRC


def import_reference_models():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import msda as O

    shim = types.ModuleType("MultiScaleDeformableAttention")

    def fwd(value, shapes, lsi, loc, attn, im2col_step):
        if value.is_cuda:                    # the native-op plugin boundary: this repo's kernels under the unmodified reference
            from dpft_b200 import msda
            return msda.ms_deform_attn_forward(value, shapes, lsi, loc, attn, im2col_step)
        return O.msda_forward_torch(value, shapes, loc, attn)

    def bwd(value, shapes, lsi, loc, attn, grad_out, im2col_step):
        if value.is_cuda:
            from dpft_b200 import msda
            return msda.ms_deform_attn_backward(value, shapes, lsi, loc, attn, grad_out, im2col_step)
        return O.msda_backward_torch(value, shapes, loc, attn, grad_out)

    shim.ms_deform_attn_forward = fwd
    shim.ms_deform_attn_backward = bwd
    sys.modules["MultiScaleDeformableAttention"] = shim
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import dprt.models as ref_models  # noqa: E402
    return ref_models


def import_reference_dataset():
    """``dprt.datasets.kradar.dataset`` of the unmodified reference.  Its package __init__ also imports the offline
    pre-processing module, whose point-cloud dependency (pypcd) is not installed here: an empty stand-in module satisfies
    that import; nothing of it is used."""
    import_reference_models()
    if "pypcd" not in sys.modules:
        try:
            import pypcd  # noqa: F401
        except ImportError:
            stub = types.ModuleType("pypcd")
            stub.pypcd = types.ModuleType("pypcd.pypcd")
            sys.modules["pypcd"] = stub
            sys.modules["pypcd.pypcd"] = stub.pypcd
    import dprt.datasets.kradar.dataset as ds
    return ds


def reference_sample_pipeline(ds_module, raw_sample, image_size, scale=True):
    """The per-sample steps of KRadarDataset.__getitem__ (dataset.py:143-169) that follow load_sample_data, on one decoded
    sample, through the reference's own methods (the instance is made without a dataset directory)."""
    d = object.__new__(ds_module.KRadarDataset)
    d.camera = "M" if "camera_mono" in raw_sample else ""
    d.radar = ("B" if "radar_bev" in raw_sample else "") + ("F" if "radar_front" in raw_sample else "")
    d.scale, d.image_size, d.dtype = scale, image_size, "float32"
    sample = {k: v.clone() for k, v in raw_sample.items()}
    if d.scale:
        sample = d.scale_radar_data(sample)
    sample = d._add_transformations(sample)
    sample = d._add_projections(sample)
    sample = d._add_shape(sample)
    if d.image_size is not None:
        sample = d.resize_image(sample)
    return sample


def import_reference_loss(box3d_overlap):
    """``dprt.training.loss`` (+ assigner, utils.iou / bbox) of the unmodified reference, without running the package
    __init__ of dprt.training (it pulls in the trainer -> tensorboard / deepspeed, absent here).  ``pytorch3d`` is not
    installed: ``box3d_overlap`` (dpft_b200.criterion.box3d_overlap) is registered as ``pytorch3d.ops.box3d_overlap`` — the
    reference's own loss / assigner / GIoU code then runs unmodified around that one function."""
    import importlib
    import_reference_models()
    p3d, ops = types.ModuleType("pytorch3d"), types.ModuleType("pytorch3d.ops")
    ops.box3d_overlap = box3d_overlap
    p3d.ops = ops
    sys.modules["pytorch3d"], sys.modules["pytorch3d.ops"] = p3d, ops
    if "dprt.training" not in sys.modules:
        pkg = types.ModuleType("dprt.training")
        pkg.__path__ = [os.path.join(REFERENCE_SRC, "dprt", "training")]
        sys.modules["dprt.training"] = pkg
    for name in ("dprt.utils.iou", "dprt.training.assigner", "dprt.training.loss"):
        sys.modules.pop(name, None)
    return importlib.import_module("dprt.training.loss")


def import_reference_metric(box3d_overlap):
    """``dprt.evaluation.metric`` of the unmodified reference, without running the package __init__ of dprt.evaluation (the
    evaluator needs deepspeed, absent here) and with ``box3d_overlap`` standing in for the absent pytorch3d op (as in
    import_reference_loss)."""
    import importlib
    import_reference_loss(box3d_overlap)
    if "dprt.evaluation" not in sys.modules:
        pkg = types.ModuleType("dprt.evaluation")
        pkg.__path__ = [os.path.join(REFERENCE_SRC, "dprt", "evaluation")]
        sys.modules["dprt.evaluation"] = pkg
    sys.modules.pop("dprt.evaluation.metric", None)
    return importlib.import_module("dprt.evaluation.metric")
