"""CPU: the C-ABI library builds, loads and exports every symbol include/*.h declares."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from clear_vae_b200 import build
    path = build.build_lib()
    return ctypes.CDLL(path)


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(clearvae_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_every_declared_symbol_is_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 8
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ but not exported"


def test_version_and_workspace_queries(lib):
    assert lib.clearvae_version() >= 100
    lib.clearvae_latent_workspace_bytes.restype = ctypes.c_size_t
    lib.clearvae_latent_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32]
    assert lib.clearvae_latent_workspace_bytes(1024, 1024, 8, 2) >= 256
    lib.clearvae_recon_workspace_bytes.restype = ctypes.c_size_t
    assert lib.clearvae_recon_workspace_bytes() > 0


def test_bad_arguments_return_error_codes_without_touching_a_gpu(lib):
    # null pointers are rejected before any CUDA call
    lib.clearvae_snn_finalize.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
    assert lib.clearvae_snn_finalize(None, 8, 0, None, None) == -1
    lib.clearvae_recon_fwd.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int64] * 2 + [ctypes.c_void_p] * 2 + [ctypes.c_size_t, ctypes.c_void_p]
    assert lib.clearvae_recon_fwd(None, None, 4, 4, None, None, 0, None) == -1


def test_product_never_imports_the_oracle():
    for py in glob.glob(os.path.join(ROOT, "clear_vae_b200", "**", "*.py"), recursive=True):
        src = open(py).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), py


def test_ops_fail_loudly_on_cpu_tensors():
    import torch
    from clear_vae_b200 import build, losses
    build.build_all()
    with pytest.raises((NotImplementedError, RuntimeError)):
        losses.contrastive_loss(torch.randn(4, 8), torch.randn(4, 8), torch.tensor([0, 0, 1, 1]), "cosine", 0.1)
