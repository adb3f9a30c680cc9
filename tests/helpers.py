"""Shared test helpers (CPU + GPU)."""
import os

import numpy as np
import torch

from oracle import model_oracle as mo
from oracle.engine_emulator import Emulator

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_step(tag):
    g = np.load(os.path.join(GOLDEN, f"step_{tag}.npz"))
    kind, arch, zdim, cin, B, hw, ncls, seed, est = (str(v) for v in g["meta"])
    hyper = {str(k): eval(str(v)) for k, v in zip(g["hyper_keys"], g["hyper_vals"])}
    meta = dict(kind=kind, arch=arch, zdim=int(zdim), cin=int(cin), B=int(B), hw=int(hw), ncls=int(ncls), seed=int(seed),
                est=None if est == "None" else est)
    return g, meta, hyper


def seeded_model(meta, device="cpu"):
    """The product module built under the golden's seed == the reference's initial weights
    (same constructors in the same order; checked against the recorded digests)."""
    from clear_vae_b200.models.vae import VAE, VAE64
    torch.manual_seed(meta["seed"])
    m = (VAE if meta["arch"] == "VAE" else VAE64)(total_z_dim=meta["zdim"], in_channel=meta["cin"])
    return m.to(device)


def digest(t):
    t = t.detach().double().flatten().cpu()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()])


def state_of(model):
    return {k: v.detach().clone().cpu() for k, v in model.state_dict().items()}


def sample_of(t, n):
    f = t.detach().flatten()
    step = max(1, f.numel() // n)
    return f[::step][:n].cpu().numpy()


def emulated_step(st, meta, hyper, g, round_bf16, slope, device="cpu"):
    """Engine-emulator forward/backward of the trainer's VAE loss (aux MI/TC terms excluded)."""
    st = {k: v.to(device) for k, v in st.items()}
    em = Emulator(st, meta["arch"], meta["cin"], round_bf16=round_bf16)
    X = torch.tensor(g["X"], device=device)
    label = torch.tensor(g["label"], device=device)
    e_c, e_s = torch.tensor(g["eps/0"], device=device), torch.tensor(g["eps/1"], device=device)
    out = em.forward(X, e_c, e_s, target=X)
    dgrads, dz = em.backward_decoder(out["tape"], X, 1.0)
    lat = out["lat"].detach().requires_grad_(True)
    D = lat.shape[1] // 4
    mu_c, lv_c, mu_s, lv_s = (lat[:, j * D:(j + 1) * D] for j in range(4))
    z = torch.cat([mu_c + e_c * torch.exp(0.5 * lv_c), mu_s + e_s * torch.exp(0.5 * lv_s)], 1)
    kl = lambda m, l: -0.5 * (1 + l - m * m - l.exp()).sum(1).mean()
    kl_c, kl_s = kl(mu_c, lv_c), kl(mu_s, lv_s)
    c = mo.contrastive(mu_c, lv_c, label, "cosine", hyper["temperature"])
    tot = (z * dz).sum() + slope * kl_c + slope * kl_s + hyper["alpha"] * c
    s = None
    if meta["kind"] == "clear":
        ps = hyper["ps"]
        s = mo.contrastive(mu_s, lv_s, label, "cosine", hyper["temperature"], ps=ps)
        if not ps:
            s = -s
        tot = tot + hyper["alpha"] * s
    (dlat,) = torch.autograd.grad(tot, lat)
    egrads = em.backward_encoder(out["tape"], dlat)
    grads = {**dgrads, **egrads}
    return dict(xhat=out["xhat"], recon=out["recon"], lat=out["lat"], z=out["z"], kl_c=kl_c, kl_s=kl_s, c=c, s=s, grads=grads)


def supcon_torch(mu, label, sim, tau, name, ps):
    """fp64 torch restatement of the two SupCon row losses + finite-row mean (oracle/latent_oracle.py:row_losses, which
    follows losses.py:140-170), written on the valid rows only so that autograd gives clean gradients.  Pinned on the CPU
    against the reference goldens (tests/test_oracle_golden.py) and used as the gradient reference of the GPU tests."""
    mu = mu.double()
    if sim == "cosine":
        nrm = mu / mu.norm(dim=1, keepdim=True).clamp_min(1e-8)
        S = nrm @ nrm.T
    else:
        S = -((mu[:, None, :] - mu[None, :, :]) ** 2).sum(-1)
    B = mu.shape[0]
    eye = torch.eye(B, dtype=torch.bool)
    m = (label[None, :] == label[:, None]) != bool(ps)
    pm = m & ~eye
    cnt = pm.sum(1)
    if name == "supcon_out_loss":
        keep = cnt > 0
        S2 = S.masked_fill(eye, -999.0)
        val = -(S2 * pm).sum(1)[keep] / cnt[keep] + torch.logsumexp(S2[keep] / tau, dim=1)
        return val[torch.isfinite(val)].mean()
    n_k = m.sum(1) - 1
    keep = (cnt > 0) & (n_k > 0)
    S2 = S.masked_fill(eye, -float("inf"))[keep]
    pos = S2.masked_fill(~m[keep], -float("inf"))
    val = n_k[keep].double().log() - torch.logsumexp(pos / tau, dim=1) + torch.logsumexp(S2 / tau, dim=1)
    return val[torch.isfinite(val)].mean()


def snn_fp64_chunked(mu, label, tau, ps=False, device="cpu", chunk=4096):
    """fp64 torch restatement of `contrastive_loss(..., "cosine", tau, "snn_loss", ps)` (losses.py:54-55, 98-137) that never
    holds more than a [chunk, B] block: returns (loss, dloss/dmu) with the gradient from autograd, chunk by chunk (row side
    and column side both accumulate into the same leaf).  Pinned against oracle/latent_oracle.py (numpy closed forms, themselves
    pinned to the reference goldens) by tests/test_latent_gpu.py::test_fp64_chunked_restatement_is_pinned before it is
    trusted at sizes the numpy oracle cannot reach in test time (B = 65536)."""
    mu = mu.detach().to(device=device, dtype=torch.float64).requires_grad_(True)
    label = label.to(device)
    B = mu.shape[0]
    total, count = torch.zeros((), dtype=torch.float64, device=device), 0
    for r0 in range(0, B, chunk):
        r1 = min(B, r0 + chunk)
        n = mu / mu.norm(dim=1, keepdim=True).clamp_min(1e-8)
        S = (n[r0:r1] @ n.T) / tau
        idx = torch.arange(r0, r1, device=device)
        eye = torch.zeros_like(S, dtype=torch.bool)
        eye[idx - r0, idx] = True
        m = (label[r0:r1, None] == label[None, :]) != bool(ps)
        S = S.masked_fill(eye, -float("inf"))
        pos = S.masked_fill(~m, -float("inf"))
        has_pos = (m & ~eye).any(1)
        # rows without positives are +inf in the reference and dropped by its finite-row mean
        row = torch.logsumexp(S[has_pos], 1) - torch.logsumexp(pos[has_pos], 1)
        fin = torch.isfinite(row)
        part = row[fin].sum()
        count += int(fin.sum())
        total = total + part.detach()
        if part.requires_grad:
            part.backward()
    if count == 0:
        return float("nan"), torch.zeros_like(mu)
    return float(total / count), (mu.grad / count)
