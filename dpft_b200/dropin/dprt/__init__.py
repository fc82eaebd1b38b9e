"""Overlay of the reference's ``dprt`` package (see dpft_b200/dropin/__init__.py): every sub-module except ``dprt.models``
is looked up in the reference's own ``dprt`` directory further down ``sys.path``."""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
__path__ = [_here]
for _p in _sys.path:
    _cand = _os.path.abspath(_os.path.join(_p or ".", "dprt"))
    if _cand != _here and _cand not in __path__ and _os.path.isfile(_os.path.join(_cand, "__init__.py")):
        __path__.append(_cand)
