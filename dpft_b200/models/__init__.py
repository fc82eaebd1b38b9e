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


def load(checkpoint: str, *args, **kwargs) -> Tuple[torch.nn.Module, int, str]:
    """``<timestamp>_checkpoint_<epoch>.pt`` -> (model, epoch, timestamp) (reference __init__.py:15-18).
    The file is a whole pickled module (reference trainer.py:258), hence ``weights_only=False``."""
    filename = os.path.splitext(os.path.basename(checkpoint))[0]
    timestamp, _, epoch = filename.split("_")
    kwargs.setdefault("map_location", "cpu")
    return torch.load(checkpoint, weights_only=False, **kwargs), int(epoch), timestamp
