"""CPU: host-side pieces of the step that need no GPU — annealer, factor shuffling, the input iterator's ordering, and the
argument validation of the C-ABI entry points added for the data-parallel / multi-draw paths (no compute calls)."""
import ctypes
import math

import numpy as np
import pytest
import torch

from oracle import latent_oracle as lo


def test_logistic_annealer_matches_the_oracle_restatement():
    from clear_vae_b200.trainer import LogisticAnnealer
    a = LogisticAnnealer(loc=5, scale=2, beta=1 / 8)
    for step in range(12):
        want = lo.logistic_anneal(step, 5, 2, 1 / 8)
        assert abs(a.slope() - want) < 1e-15
        assert abs(float(a(torch.tensor(3.0))) - 3.0 * want) < 1e-6
        a.step()
    assert a.current_step == 12
    assert abs(LogisticAnnealer(0, 1, 0.5).slope() - 0.25) < 1e-15      # beta / 2 at step 0 (trainer_utils.py:110-111)


def test_factor_shuffling_rolls_the_style_half_up_by_one_row():
    from clear_vae_b200.trainer import factor_shuffling
    z = torch.arange(5 * 6, dtype=torch.float32).view(5, 6)
    got = factor_shuffling(z)
    assert torch.equal(got, torch.tensor(lo.roll_style_half(z.numpy())))
    assert torch.equal(got[:, :3], z[:, :3]) and torch.equal(got[:-1, 3:], z[1:, 3:]) and torch.equal(got[-1, 3:], z[0, 3:])
    with pytest.raises(TypeError):
        factor_shuffling(z, "full")        # the reference's 'full' branch calls a tensor (trainer.py:581)
    with pytest.raises(ValueError):
        factor_shuffling(z, "other")


def test_prefetcher_keeps_batches_and_order_on_the_host_path():
    from clear_vae_b200.trainer import DevicePrefetcher
    gen = torch.Generator().manual_seed(0)
    src = [(torch.rand(b, 1, 4, 4, generator=gen), torch.randint(0, 9, (b, 1), generator=gen)) for b in (8, 8, 8, 3)]
    out = list(DevicePrefetcher(src, torch.device("cpu")))
    assert len(out) == 4
    for (X, y), (xs, ys) in zip(out, src):
        assert torch.equal(X, xs) and torch.equal(y, ys.reshape(-1)) and y.dtype == torch.int64
    doubled = list(DevicePrefetcher(src, torch.device("cpu"), transform=lambda t: 2 * t))
    assert all(torch.equal(X, 2 * xs) for (X, _), (xs, _) in zip(doubled, src))
    assert list(DevicePrefetcher([], torch.device("cpu"))) == []


@pytest.fixture(scope="module")
def lib():
    from clear_vae_b200 import build
    return ctypes.CDLL(build.build_lib())


def test_peer_and_reparam_entry_points_reject_bad_arguments_before_any_cuda_call(lib):
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    lib.clearvae_peer_gather.argtypes = [vp, i32, i32, i64, i32, vp, vp, vp, vp]
    lib.clearvae_peer_allreduce.argtypes = [vp, i32, i32, i64, i32, vp, vp, vp]
    lib.clearvae_reparam_multi.argtypes = [i32, i32, vp, vp, vp, vp, i64, i32, vp]
    lib.clearvae_peer_alloc.argtypes = [i64, vp]
    lib.clearvae_latent_bwd_workspace_bytes.restype = ctypes.c_size_t
    lib.clearvae_latent_bwd_workspace_bytes.argtypes = [i64, i64, i32, i32]
    assert lib.clearvae_peer_gather(None, 2, 0, 1 << 20, 1, None, None, None, None) == -1      # no buffer table
    bases = (vp * 2)(0x1000, 0x2000)
    assert lib.clearvae_peer_gather(bases, 9, 0, 1 << 20, 1, None, None, None, None) == -1     # more than 8 ranks
    assert lib.clearvae_peer_gather(bases, 2, 2, 1 << 20, 1, None, None, None, None) == -1     # rank out of range
    assert lib.clearvae_peer_gather(bases, 2, 0, 1 << 20, 9, None, None, None, None) == -1     # more than 8 pieces
    src, dst, nb = (vp * 1)(0x3000), (vp * 1)(0x4000), (i64 * 1)(6)
    assert lib.clearvae_peer_gather(bases, 2, 0, 1 << 20, 1, src, dst, nb, None) == -1         # bytes not a multiple of 4
    nb[0] = 4 << 20
    assert lib.clearvae_peer_gather(bases, 2, 0, 1 << 20, 1, src, dst, nb, None) == -3         # does not fit a slot
    assert lib.clearvae_peer_allreduce(bases, 2, 0, 1 << 20, 0, None, None, None) == 0         # nothing to do
    assert lib.clearvae_peer_allreduce(bases, 2, 0, 1 << 20, 1, None, None, None) == -1
    assert lib.clearvae_peer_alloc(16, None) == -1
    assert lib.clearvae_reparam_multi(3, 1, None, None, None, None, 8, 8, None) == -1          # at most two heads
    assert lib.clearvae_reparam_multi(2, 9, None, None, None, None, 8, 8, None) == -1          # at most eight draws
    # the split backward needs room for tickets + 8 partial sets of (D_padded + 1) floats per row and term
    assert lib.clearvae_latent_bwd_workspace_bytes(1024, 8192, 8, 2) >= 2 * 8 * 1024 * 9 * 4
    assert lib.clearvae_latent_bwd_workspace_bytes(1 << 16, 1 << 16, 8, 2) <= 4096             # large batches never split
