"""Checkpoint interchange with DPFT.

The reference stores a checkpoint as the whole pickled module (``torch.save(model, path)``, src/dprt/training/trainer.py:258)
and reads it back with ``torch.load(checkpoint)`` (src/dprt/models/__init__.py:15-18): a file that only loads where the
``dprt`` package (and its compiled deformable-attention dependency) is importable, and that executes arbitrary pickled
code.  This module adds

  * ``save_state`` / ``load_state``: state dict + config + epoch in a ``weights_only=True``-loadable file, names identical
    to the reference's (so either model class can consume it);
  * ``reference_state_dict``: the tensors of a REFERENCE whole-module checkpoint (e.g. the published Zenodo files) without
    importing ``dprt``: classes that cannot be resolved are replaced by inert stand-ins while unpickling and the module
    tree is walked for ``_parameters`` / ``_buffers`` exactly like ``nn.Module.state_dict`` does;
  * ``convert_reference_checkpoint``: reference pickle + its JSON config -> a ``dpft_b200`` model / a state file.
"""
from __future__ import annotations

import pickle
from collections import OrderedDict
from typing import Any, Dict, Iterable, Optional, Tuple

import torch

FORMAT = "dpft_b200.state/1"


# ---- weights-only state files ------------------------------------------------------------------------------------------
def save_state(model: torch.nn.Module, path: str, config: Optional[Dict[str, Any]] = None, epoch: Optional[int] = None,
               timestamp: Optional[str] = None) -> None:
    """State dict (reference parameter names) + the JSON config it was built from; loadable with ``weights_only=True``."""
    payload = {"format": FORMAT, "state_dict": OrderedDict((k, v.detach().cpu()) for k, v in model.state_dict().items()),
               "config": config, "epoch": epoch, "timestamp": timestamp}
    torch.save(payload, path)


def load_state(path: str, config: Optional[Dict[str, Any]] = None, strict: bool = True) -> Tuple[torch.nn.Module, Optional[int], Optional[str]]:
    """-> (model, epoch, timestamp).  ``config`` overrides the one stored in the file."""
    from . import models
    payload = torch.load(path, map_location="cpu", weights_only=True)
    if not isinstance(payload, dict) or payload.get("format") != FORMAT:
        raise ValueError(f"{path} is not a {FORMAT} file")
    cfg = config if config is not None else payload.get("config")
    if cfg is None:
        raise ValueError(f"{path} carries no config; pass config=")
    model = models.build("dprt", cfg)
    model.load_state_dict(payload["state_dict"], strict=strict)
    return model, payload.get("epoch"), payload.get("timestamp")


# ---- reference whole-module pickles without the reference package ---------------------------------------------------------
class _Standin:
    """Inert replacement for a class that cannot be imported: keeps whatever state the pickle hands it."""

    def __init__(self, *args, **kwargs):
        self._standin_args = (args, kwargs)

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self._standin_state = state

    def __call__(self, *args, **kwargs):          # a pickled reference to a function that is gone
        raise RuntimeError("stand-in for an unavailable class/function of the reference checkpoint")


def _standin_class(module: str, name: str):
    return type(name, (_Standin,), {"__module__": module, "_standin_origin": f"{module}.{name}"})


# Globals a pickled reference model legitimately needs to rebuild its TENSORS and containers.  Everything else — the model
# classes themselves (dprt.*, torchvision.*, torch.nn.*) and anything unknown — becomes an inert stand-in whose constructor
# and __reduce__ hooks never run, so a crafted "checkpoint" cannot execute code through this loader.
_SAFE_GLOBALS = {
    ("collections", "OrderedDict"), ("collections", "defaultdict"), ("builtins", "set"), ("builtins", "frozenset"),
    ("builtins", "list"), ("builtins", "dict"), ("builtins", "tuple"), ("builtins", "int"), ("builtins", "float"),
    ("builtins", "bool"), ("builtins", "str"), ("builtins", "bytes"), ("builtins", "complex"), ("builtins", "slice"),
    ("builtins", "range"), ("__builtin__", "set"), ("__builtin__", "frozenset"),
    ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_tensor"), ("torch._utils", "_rebuild_parameter"),
    ("torch._utils", "_rebuild_parameter_with_state"), ("torch._utils", "_rebuild_qtensor"),
    ("torch._tensor", "_rebuild_from_type_v2"), ("torch", "Size"), ("torch", "device"), ("torch", "dtype"),
    ("torch", "Tensor"), ("torch.nn.parameter", "Parameter"), ("torch.nn.parameter", "Buffer"),
    ("torch.serialization", "_get_layout"), ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
}
_SAFE_TORCH_STORAGES = ("FloatStorage", "DoubleStorage", "HalfStorage", "BFloat16Storage", "LongStorage", "IntStorage",
                        "ShortStorage", "CharStorage", "ByteStorage", "BoolStorage", "UntypedStorage")


class _TolerantUnpickler(pickle.Unpickler):
    force_standin: Tuple[str, ...] = ()

    def find_class(self, module: str, name: str):
        forced = any(module == p or module.startswith(p + ".") for p in self.force_standin)
        allowed = (module, name) in _SAFE_GLOBALS or (module in ("torch", "torch.storage") and name in _SAFE_TORCH_STORAGES)
        if allowed and not forced:
            try:
                return super().find_class(module, name)     # also applies pickle's Python-2 name mapping (__builtin__.set ...)
            except (ImportError, AttributeError):
                pass
        return _standin_class(module, name)


def _pickle_module(force_standin: Iterable[str]):
    class Unpickler(_TolerantUnpickler):
        pass
    Unpickler.force_standin = tuple(force_standin)

    class Mod:                                     # the minimal ``pickle_module`` protocol torch.load uses
        __name__ = "dpft_b200.checkpoint.tolerant_pickle"
    Mod.Unpickler = Unpickler
    Mod.load = staticmethod(lambda f, **kw: Unpickler(f, **kw).load())
    Mod.loads = staticmethod(pickle.loads)
    Mod.dump, Mod.dumps, Mod.Pickler = staticmethod(pickle.dump), staticmethod(pickle.dumps), pickle.Pickler
    return Mod


def _walk_state(obj, prefix: str, out: "OrderedDict[str, torch.Tensor]") -> None:
    """nn.Module.state_dict() on a tree whose nodes may be real modules or stand-ins."""
    d = obj.__dict__
    skip = d.get("_non_persistent_buffers_set", set())
    for name, p in (d.get("_parameters") or {}).items():
        if p is not None:
            out[prefix + name] = p.detach() if isinstance(p, torch.Tensor) else p
    for name, b in (d.get("_buffers") or {}).items():
        if b is not None and name not in skip:
            out[prefix + name] = b.detach()
    for name, m in (d.get("_modules") or {}).items():
        if m is not None:
            _walk_state(m, prefix + name + ".", out)


def reference_state_dict(path: str, force_standin: Iterable[str] = ()) -> "OrderedDict[str, torch.Tensor]":
    """Tensors of a reference ``torch.save(model)`` checkpoint, keyed like ``model.state_dict()``.  Works without the
    ``dprt`` package (or torchvision): unresolvable classes become stand-ins.  ``force_standin``: module prefixes to
    replace even when importable (tests)."""
    obj = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_pickle_module(force_standin))
    if isinstance(obj, dict) and "state_dict" in obj:
        return OrderedDict(obj["state_dict"])
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    _walk_state(obj, "", out)
    if not out:
        raise ValueError(f"{path}: no parameters found (not a pickled module?)")
    return out


def convert_reference_checkpoint(path: str, config: Dict[str, Any], out_path: Optional[str] = None, strict: bool = True):
    """Reference whole-module checkpoint + the JSON config it was trained with -> dpft_b200 model (and, with ``out_path``, a
    weights-only state file).  ``backbones.*.weights`` is cleared so that building the model does not download ImageNet
    weights the checkpoint overrides anyway."""
    import copy
    from . import models
    cfg = copy.deepcopy(config)
    for bb in (cfg.get("model", {}).get("backbones") or {}).values():
        if isinstance(bb, dict):
            bb["weights"] = ""
    model = models.build("dprt", cfg)
    model.load_state_dict(reference_state_dict(path), strict=strict)
    if out_path is not None:
        import os
        name = os.path.splitext(os.path.basename(path))[0].split("_")
        epoch = int(name[2]) if len(name) == 3 and name[2].isdigit() else None
        save_state(model, out_path, cfg, epoch=epoch, timestamp=name[0] if len(name) == 3 else None)
    return model
