"""ctypes binding of libdpft_b200.so (the C ABI in include/dpft_b200.h).

There is NO fallback: if the library is missing this module raises, and every op in the package goes
through it.  Prototypes are generated from the header so the binding cannot drift from the ABI.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

import torch

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libdpft_b200.so")
HEADER_PATH = os.path.join(PKG_DIR, "..", "include", "dpft_b200.h")

DPFT_F32, DPFT_F64, DPFT_F16, DPFT_BF16 = 0, 1, 2, 3
_DTYPE_CODE = {torch.float32: DPFT_F32, torch.float64: DPFT_F64, torch.float16: DPFT_F16,
               torch.bfloat16: DPFT_BF16}

_CTYPES = {
    "int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double,
    "int64_t": ctypes.c_int64, "long long": ctypes.c_longlong, "size_t": ctypes.c_size_t,
    "unsigned": ctypes.c_uint, "unsigned int": ctypes.c_uint,
}


class NativeLibraryError(RuntimeError):
    pass


def declared_symbols(header: str = HEADER_PATH) -> Dict[str, Tuple[str, List[str]]]:
    """Parses ``DPFT_API <ret> name(args);`` declarations out of the header: name -> (ret, [arg types])."""
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"DPFT_API\s+([\w\s\*]+?)\s*\b(dpft_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        arg_types = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    arg_types.append("ptr")
                else:
                    arg_types.append(" ".join(a.replace("const", "").split()[:-1]))
        out[name] = (ret, arg_types)
    return out


def _to_ctype(t: str):
    if t == "ptr":
        return ctypes.c_void_p
    return _CTYPES[t]


_LIB = None


def load_library(path: str = LIB_PATH):
    """Loads the shared library and applies the header's prototypes.  Raises if it is absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(path):
        raise NativeLibraryError(
            f"{path} is missing: build it with `python -m dpft_b200.build` (or __graft_entry__.build()). "
            "dpft_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(path)
    for name, (ret, args) in declared_symbols().items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = ctypes.c_char_p if "char" in ret else ctypes.c_int
        fn.argtypes = [_to_ctype(a) for a in args]
    if lib.dpft_abi_version() != 1:
        raise NativeLibraryError("libdpft_b200.so ABI version mismatch")
    _LIB = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load_library().dpft_last_error().decode()
        raise RuntimeError(f"{what} failed with status {status}: {msg}")


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPE_CODE[t.dtype]
    except KeyError:
        raise RuntimeError(f"dpft_b200: unsupported dtype {t.dtype}") from None


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if not t.is_cuda:
            # same contract as the op DPFT links against: no CPU implementation
            raise RuntimeError("dpft_b200: not implemented on the CPU (tensor is not a CUDA tensor)")


def ptr(t) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr(device=None) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


# launch accounting for bench.py's "gpu_launches" (host-side count of our own kernel launches)
_launches = 0


def count_launch(n: int = 1) -> None:
    global _launches
    _launches += n


def launches() -> int:
    return _launches
