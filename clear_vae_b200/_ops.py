"""Loader for the native extension.  There is NO Python/CPU fallback: if the
in-tree `_C.so` / `_lib/libclearvae_b200.so` are missing or fail to load, every
op raises.  Build them with `python -m clear_vae_b200.build` (or
`__graft_entry__.build()`)."""
from __future__ import annotations

import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
_EXT = os.path.join(_PKG, "_C.so")
_LIB = os.path.join(_PKG, "_lib", "libclearvae_b200.so")
_loaded = False


class NativeExtensionMissing(RuntimeError):
    pass


def load():
    """Load the torch custom ops (idempotent)."""
    global _loaded
    if _loaded:
        return torch.ops.clearvae
    if not (os.path.exists(_EXT) and os.path.exists(_LIB)):
        raise NativeExtensionMissing(
            f"clear_vae_b200 native extension not built ({_EXT}); run `python -m clear_vae_b200.build`. "
            "There is no CPU / PyTorch fallback for the hot path.")
    torch.ops.load_library(_EXT)
    _loaded = True
    return torch.ops.clearvae


def ops():
    return load()


def lib_paths():
    return _LIB, _EXT
