"""Golden vectors of the comparison baselines from the UNMODIFIED reference (CPU, fp32):
group evidence (ML-VAE / GVAE), group-wise reparameterisation, one HierarchicalVAETrainer step per mode, one SimpleCNN and
one LAM-CNN step.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_baselines.py        ->  tests/golden/baselines.npz

Random draws are made injectable exactly as in make_golden.py: `torch.randn_like` returns recorded tensors; the group noise
of `groupwise_reparam_each` (`torch.randn(n, D)` on the CPU generator, vae.py:205) is reproduced by seeding the generator,
which the product consumes identically.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/code")
from src.losses import lam_loss, vae_loss  # noqa: E402
from src.models.vae import VAE, accumulate_group_evidence, groupwise_reparam_each  # noqa: E402
from src.utils.trainer_utils import get_cnn_trainer, get_hierarchical_vae_trainer, get_lamcnn_trainer  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)
npy = lambda t: t.detach().cpu().numpy().copy()


def sample_of(t, n):
    f = t.detach().flatten()
    step = max(1, f.numel() // n)
    return f[::step][:n].numpy().copy()


def main():
    out = {}
    g = torch.Generator().manual_seed(2024)
    # ---- group evidence: 24 rows, labels with a singleton group and non-contiguous label values
    B, D = 24, 8
    mu = torch.randn(B, D, generator=g).requires_grad_(True)
    lv = (torch.randn(B, D, generator=g) * 0.5).requires_grad_(True)
    lab = torch.tensor([3, 3, 7, 1, 1, 1, 9, 3, 7, 7, 1, 42, 9, 9, 3, 1, 7, 3, 9, 1, 7, 3, 9, 1])
    w1, w2, w3 = torch.randn(5, D, generator=g), torch.randn(5, D, generator=g), torch.randn(B, D, generator=g)
    out["ge/mu"], out["ge/lv"], out["ge/label"], out["ge/w1"], out["ge/w2"], out["ge/w3"] = npy(mu), npy(lv), npy(lab), npy(w1), npy(w2), npy(w3)
    for mode in ("MLVAE", "GVAE"):
        mg, lg, gd = accumulate_group_evidence(mu, lv, lab, mode)
        torch.manual_seed(77)
        z, idx, sizes = groupwise_reparam_each(mg, lg, gd)
        f = (mg * w1).sum() + (lg * w2).sum() + (z * w3).sum()
        gm, gl = torch.autograd.grad(f, [mu, lv])
        out[f"ge/{mode}/mu_g"], out[f"ge/{mode}/lv_g"], out[f"ge/{mode}/z"] = npy(mg), npy(lg), npy(z)
        out[f"ge/{mode}/indices"], out[f"ge/{mode}/sizes"] = npy(idx), npy(sizes)
        out[f"ge/{mode}/dmu"], out[f"ge/{mode}/dlv"], out[f"ge/{mode}/f"] = npy(gm), npy(gl), npy(f)
        out[f"ge/{mode}/keys"] = np.array(list(gd.keys()))
    # ---- one HierarchicalVAETrainer step per mode (VAE(16, 3), B = 16)
    Bh = 16
    X = torch.rand(Bh, 3, 28, 28, generator=g)
    y = torch.randint(0, 4, (Bh,), generator=g)
    eps_s = torch.randn(Bh, 8, generator=g)
    out["hv/X"], out["hv/label"], out["hv/eps_s"] = npy(X), npy(y), npy(eps_s)
    for mode in ("MLVAE", "GVAE"):
        torch.manual_seed(5)
        tr = get_hierarchical_vae_trainer(1 / 8, 5e-4, 16, mode, "cpu", "VAE", 3)
        vae = tr.model
        vae.train()
        o_randn = torch.randn_like
        torch.randn_like = lambda t, *a, **k: eps_s.clone()
        try:
            torch.manual_seed(99)        # the CPU generator state the group noise is drawn from
            xhat, lp = vae(X, label=y)
        finally:
            torch.randn_like = o_randn
        rec, kc, ks = vae_loss(xhat, X, **lp)
        n_groups = len(y.unique())
        rec_a, ks_a = rec * Bh / n_groups, ks * Bh / n_groups
        slope = tr.annealer.slope()
        loss = rec_a + slope * kc + slope * ks_a
        loss.backward()
        out[f"hv/{mode}/recon"], out[f"hv/{mode}/kl_c"], out[f"hv/{mode}/kl_s"], out[f"hv/{mode}/loss"] = npy(rec_a), npy(kc), npy(ks_a), npy(loss)
        out[f"hv/{mode}/mu_c"], out[f"hv/{mode}/logvar_c"] = npy(lp["mu_c"]), npy(lp["logvar_c"])
        for k, p in vae.named_parameters():
            out[f"hv/{mode}/grad_sample/{k}"] = sample_of(p.grad, 256)
    # ---- CNN baselines (3 x 28 x 28, 10 classes, B = 16)
    yc = torch.randint(0, 10, (Bh,), generator=g)
    out["cnn/label"] = npy(yc)
    torch.manual_seed(6)
    tr = get_cnn_trainer(10, "cpu", "SimpleCNNClassifier", 3)
    tr.model.train()
    logits = tr.model(X)
    loss = tr.criterion(logits, yc)
    loss.backward()
    out["cnn/logits"], out["cnn/loss"] = npy(logits), npy(loss)
    for k, p in tr.model.named_parameters():
        out[f"cnn/grad_sample/{k}"] = sample_of(p.grad, 256)
    torch.manual_seed(6)
    tr = get_lamcnn_trainer(10, "cpu", 0.5, "LAMCNNClassifier", 3)
    cnn = tr.model
    cnn.train()
    torch.manual_seed(123)
    Xt = tr.ss_pairing(X, yc)
    logits = cnn(X)
    l_ce = tr.criterion(logits, yc)
    l_lam = lam_loss(cnn.net(X), cnn.net(Xt), yc, cnn.cls_head.weight)
    (l_ce + 0.5 * l_lam).backward()
    out["lam/X_tilde"], out["lam/logits"], out["lam/ce"], out["lam/lam"] = npy(Xt), npy(logits), npy(l_ce), npy(l_lam)
    for k, p in cnn.named_parameters():
        out[f"lam/grad_sample/{k}"] = sample_of(p.grad, 256)
    np.savez_compressed(os.path.join(HERE, "baselines.npz"), **out)
    print("baselines.npz written:", len(out), "arrays;", {m: float(out[f"hv/{m}/loss"]) for m in ("MLVAE", "GVAE")},
          float(out["cnn/loss"]), float(out["lam/lam"]))


if __name__ == "__main__":
    main()
