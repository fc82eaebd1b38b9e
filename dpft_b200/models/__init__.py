"""Drop-in for ``dprt.models`` (reference src/dprt/models/__init__.py): ``build(model, config)`` and
``load(checkpoint)`` with the same signatures, return values and checkpoint-name convention."""
import os
from typing import Tuple

import torch

from .dprt import DPRT, build_dprt  # noqa: F401


def build(model: str, *args, **kwargs):
    """Returns the model for ``model == 'dprt'`` and ``None`` for any other name (reference __init__.py:10-12)."""
    if model == "dprt":
        return build_dprt(*args, **kwargs)


def load(checkpoint: str, *args, config=None, **kwargs) -> Tuple[torch.nn.Module, int, str]:
    """``<timestamp>_checkpoint_<epoch>.pt`` -> (model, epoch, timestamp) (reference __init__.py:15-18).
    The file is a whole pickled module (reference trainer.py:258), hence ``weights_only=False`` — or a weights-only state
    file written by ``dpft_b200.checkpoint.save_state`` (state dict + config), which is tried first."""
    from ..checkpoint import FORMAT, load_state
    filename = os.path.splitext(os.path.basename(checkpoint))[0]
    timestamp, _, epoch = filename.split("_")
    try:
        payload = torch.load(checkpoint, map_location="cpu", weights_only=True)
    except Exception:                                      # a pickled module: not loadable under weights_only
        payload = None
    if isinstance(payload, dict) and payload.get("format") == FORMAT:
        model, _, _ = load_state(checkpoint, config=config)
        return model, int(epoch), timestamp
    kwargs.setdefault("map_location", "cpu")
    return torch.load(checkpoint, weights_only=False, **kwargs), int(epoch), timestamp
