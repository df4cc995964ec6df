"""Locate and load the native libraries.  There is NO CPU or eager fallback: if the
CUDA libraries are missing every op raises with the build hint."""
from __future__ import annotations

import ctypes
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libkvc_b200.so")
EXT_PATH = os.path.join(PKG, "_C.so")

_lib = None
_ext_loaded = False


class NativeLibraryMissing(RuntimeError):
    pass


def _missing(path: str) -> NativeLibraryMissing:
    return NativeLibraryMissing(
        f"{path} not found. Build the sm_100a kernels first: "
        f"python -c 'import __graft_entry__ as g; g.build()' (repo root). "
        f"This package has no CPU fallback by design.")


def load_lib() -> ctypes.CDLL:
    """libkvc_b200.so (C ABI)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise _missing(LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        _lib.kvc_last_error.restype = ctypes.c_char_p
        _lib.kvc_arch.restype = ctypes.c_char_p
        _lib.kvc_schedule_cache_evictions_workspace_bytes.restype = ctypes.c_size_t
    return _lib


def load_torch_ops() -> None:
    """_C.so: registers torch.ops._C / _C_cache_ops / _C_kvc_ops (reference schemas)."""
    global _ext_loaded
    if _ext_loaded:
        return
    if not os.path.exists(EXT_PATH):
        raise _missing(EXT_PATH)
    import torch
    load_lib()
    torch.ops.load_library(EXT_PATH)
    _ext_loaded = True
