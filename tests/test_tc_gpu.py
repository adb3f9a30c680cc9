"""GPU parity: fused density-ratio TC kernels (CLEAR-TC-VAE, trainer.py:573-587, 654-699) against the reference goldens
(`tests/golden/heads.npz`, produced by the unmodified reference) and an fp64 torch restatement at other sizes."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def close(a, b, rel=1e-5, ab=2e-6):
    return abs(float(a) - float(b)) <= rel * abs(float(b)) + ab


def grad_close(a, b, rel=1e-4):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max()) <= rel * float(b.abs().max()) + 1e-9


def make_fc(Z):
    return nn.Sequential(nn.Linear(Z, Z), nn.ReLU(), nn.Linear(Z, 1), nn.Sigmoid())


def test_tc_terms_match_reference_goldens(golden_dir):
    from clear_vae_b200 import tc
    g = np.load(os.path.join(golden_dir, "heads.npz"))
    fc = make_fc(16)
    fc.load_state_dict({k: torch.tensor(g[f"tc/state/{k}"]) for k in ("0.weight", "0.bias", "2.weight", "2.bias")})
    fc = fc.to(DEV)
    fp = tc.fused_params(fc)
    assert fp is not None
    z = torch.tensor(g["tc/z"], device=DEV, requires_grad=True)
    mi = tc.tc_bound(z, fp)
    (2.5 * mi).backward()
    assert close(mi, g["tc/mi"]), (float(mi), float(g["tc/mi"]))
    assert grad_close(z.grad, 2.5 * torch.tensor(g["tc/dz"]))
    fl = tc.disc_grads(z.detach(), fp)
    assert close(fl, g["tc/factor_loss"]), (float(fl), float(g["tc/factor_loss"]))
    # the kernel's shuffle is the reference's factor_shuffling(z, "permute_1")
    from clear_vae_b200.trainer import factor_shuffling
    assert np.array_equal(factor_shuffling(z.detach()).cpu().numpy(), g["tc/shuffled"])


@pytest.mark.parametrize("B,Z", [(512, 64), (1000, 16), (33, 64), (2, 8), (130, 22)])
def test_tc_terms_match_fp64_restatement(B, Z):
    from clear_vae_b200 import tc
    from clear_vae_b200.trainer import factor_shuffling
    torch.manual_seed(B + Z)
    fc = make_fc(Z).to(DEV)
    fp = tc.fused_params(fc)
    z = (torch.randn(B, Z) * 1.5)
    ref = make_fc(Z).double()
    ref.load_state_dict({k: v.detach().cpu().double() for k, v in fc.state_dict().items()})
    zr = z.double().requires_grad_(True)
    d = ref(zr)
    mi_w = F.relu(torch.log(d / (1 - d))).mean()
    mi_w.backward()
    zd = z.to(DEV).requires_grad_(True)
    mi = tc.tc_bound(zd, fp)
    mi.backward()
    assert close(mi, mi_w, ab=5e-6), (float(mi), float(mi_w))
    assert grad_close(zd.grad, zr.grad)
    # discriminator loss + parameter gradients
    ref.zero_grad()
    zz = z.double()
    dj, dm = ref(zz), ref(factor_shuffling(zz))
    fl_w = F.binary_cross_entropy(torch.cat([dj, dm], 0), torch.cat([torch.ones_like(dj), torch.zeros_like(dm)], 0))
    fl_w.backward()
    fl = tc.disc_grads(z.to(DEV), fp)
    assert close(fl, fl_w), (float(fl), float(fl_w))
    for (k, p), (_, q) in zip(fc.named_parameters(), ref.named_parameters()):
        assert grad_close(p.grad, q.grad), k
    fl2 = tc.disc_grads(z.to(DEV), fp)
    assert float(fl2) == float(fl)       # fixed-order reduction: bit-reproducible


def test_other_discriminators_keep_the_module_path():
    from clear_vae_b200 import tc
    assert tc.fused_params(nn.Sequential(nn.Linear(8, 8), nn.Tanh(), nn.Linear(8, 1), nn.Sigmoid()).to(DEV)) is None
    assert tc.fused_params(make_fc(128).to(DEV)) is None
    assert tc.fused_params(make_fc(16)) is None   # CPU parameters
