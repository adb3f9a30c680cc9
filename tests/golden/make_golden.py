"""Generate golden vectors by running the UNMODIFIED reference (CPU, fp32).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz.  The GPU box never runs this file; tests read the
committed .npz fixtures.

Only two things are patched around the reference while it runs, both to make
its random draws injectable and its CUDA-only line executable on CPU:
  * `torch.randn_like` / `torch.randperm` return prepared tensors (the
    reparameterisation noise and the CLUB-S permutation), recorded in the
    fixture;
  * `torch.Tensor.cuda` is the identity while `L1OutUB.forward` runs
    (mi_estimator.py:185 hard-codes `.cuda()`).
"""
import copy
import contextlib
import os
import sys

import numpy as np
import torch

REF = "/root/reference/code"
sys.path.insert(0, REF)
from src.losses import contrastive_loss, vae_loss  # noqa: E402
from src.models.vae import VAE, VAE64  # noqa: E402
from src.models.mi_estimator import CLUBSample, L1OutUB  # noqa: E402
from src.trainer import factor_shuffling, LogisticAnnealer  # noqa: E402
from src.utils.trainer_utils import (  # noqa: E402
    get_clearvae_trainer, get_cleartcvae_trainer, get_clearmimvae_trainer)

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


@contextlib.contextmanager
def injected(randn_queue=None, perm_queue=None, cuda_identity=False):
    """Feed prepared noise / permutations to the reference."""
    o_randn, o_perm, o_cuda = torch.randn_like, torch.randperm, torch.Tensor.cuda
    rq = list(randn_queue or [])
    pq = list(perm_queue or [])

    def randn_like(t, *a, **k):
        e = rq.pop(0)
        assert e.shape == t.shape, (e.shape, t.shape)
        return e.clone()

    def randperm(n, *a, **k):
        p = pq.pop(0)
        assert p.numel() == n
        return p.clone()

    if randn_queue is not None:
        torch.randn_like = randn_like
    if perm_queue is not None:
        torch.randperm = randperm
    if cuda_identity:
        torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.randn_like, torch.randperm, torch.Tensor.cuda = o_randn, o_perm, o_cuda


def npy(t):
    # copy: .numpy() aliases the tensor and parameters are updated in place later
    return t.detach().cpu().numpy().copy() if isinstance(t, torch.Tensor) else np.asarray(t)


# --------------------------------------------------------------------------
# 1. contrastive-loss known answers
# --------------------------------------------------------------------------
def contrastive_cases():
    out = {}
    cases = []
    # SURVEY.md §4 known-answer inputs
    torch.manual_seed(0)
    mu = torch.randn(8, 4)
    lv = torch.randn(8, 4)
    lab = torch.tensor([0, 0, 1, 1, 2, 2, 2, 3])
    for sim in ("cosine", "l2", "jeffrey", "mahalanobis", "modified_l2"):
        for ps in (False, True):
            cases.append((f"kat_{sim}_ps{int(ps)}", mu, lv, lab, sim, 0.1, "snn_loss", ps))
    g = torch.Generator().manual_seed(101)

    def rnd(B, D, ncls, scale=1.0):
        return (torch.randn(B, D, generator=g) * scale, torch.randn(B, D, generator=g) * 0.5,
                torch.randint(0, ncls, (B,), generator=g))

    for (B, D, ncls, tau) in [(128, 8, 10, 0.1), (256, 32, 4, 0.1), (192, 8, 7, 0.5), (100, 32, 10, 2.0),
                              (64, 16, 3, 0.05), (130, 5, 10, 0.1)]:
        m_, l_, y_ = rnd(B, D, ncls)
        for ps in (False, True, None):
            cases.append((f"cos_B{B}_D{D}_t{tau}_ps{ps}", m_, l_, y_, "cosine", tau, "snn_loss", ps))
    # near-singleton labels (many rows dropped from the mean)
    m_, l_, y_ = rnd(96, 8, 48)
    for ps in (False, True):
        cases.append((f"cos_sparse_ps{int(ps)}", m_, l_, y_, "cosine", 0.1, "snn_loss", ps))
    # degenerate: every label equal
    m_, l_, _ = rnd(16, 8, 2)
    y_ = torch.zeros(16, dtype=torch.long)
    for ps in (False, True):
        cases.append((f"cos_allsame_ps{int(ps)}", m_, l_, y_, "cosine", 0.1, "snn_loss", ps))
    # every label distinct
    y_ = torch.arange(16)
    for ps in (False, True):
        cases.append((f"cos_alldiff_ps{int(ps)}", m_, l_, y_, "cosine", 0.1, "snn_loss", ps))
    # B = 2, and a zero row (norm clamp)
    m2 = torch.tensor([[1.0, 2.0, 0.5], [0.3, -1.0, 2.0]])
    cases.append(("cos_B2_same", m2, torch.zeros(2, 3), torch.tensor([1, 1]), "cosine", 0.1, "snn_loss", False))
    cases.append(("cos_B2_diff_ps", m2, torch.zeros(2, 3), torch.tensor([0, 1]), "cosine", 0.1, "snn_loss", True))
    m_, l_, y_ = rnd(32, 8, 4)
    m_[5] = 0.0
    cases.append(("cos_zero_row", m_, l_, y_, "cosine", 0.1, "snn_loss", False))
    # other similarity functions / loss names (the "next" rows)
    m_, l_, y_ = rnd(64, 8, 5, 0.7)
    for sim in ("l2", "jeffrey", "mahalanobis", "modified_l2"):
        for ps in (False, True):
            cases.append((f"{sim}_B64_ps{int(ps)}", m_, l_, y_, sim, 0.5, "snn_loss", ps))
    for ln in ("supcon_in_loss", "supcon_out_loss"):
        for ps in (False, True):
            cases.append((f"{ln}_B64_ps{int(ps)}", m_, l_, y_, "cosine", 0.1, ln, ps))
    names = []
    for (name, mu, lv, lab, sim, tau, ln, ps) in cases:
        mu_ = mu.clone().requires_grad_(True)
        lv_ = lv.clone().requires_grad_(True)
        loss = contrastive_loss(mu_, lv_, lab, sim, tau, ln, ps)
        if torch.isfinite(loss):
            loss.backward()
        out[f"{name}/mu"] = npy(mu)
        out[f"{name}/logvar"] = npy(lv)
        out[f"{name}/label"] = npy(lab)
        out[f"{name}/loss"] = npy(loss)
        out[f"{name}/dmu"] = npy(mu_.grad) if mu_.grad is not None else np.zeros(0, np.float32)
        out[f"{name}/dlogvar"] = npy(lv_.grad) if lv_.grad is not None else np.zeros(0, np.float32)
        out[f"{name}/meta"] = np.array([sim, repr(tau), ln, repr(ps)])
        names.append(name)
        # fp64 run of the same reference code: the accuracy floor of the fp32 oracle
        l64 = contrastive_loss(mu.double(), lv.double(), lab, sim, tau, ln, ps)
        out[f"{name}/loss64"] = npy(l64)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "contrastive.npz"), **out)
    print("contrastive:", len(names), "cases")


# --------------------------------------------------------------------------
# 2. ELBO terms, reparam, MI/TC heads
# --------------------------------------------------------------------------
def head_cases():
    out = {}
    g = torch.Generator().manual_seed(7)
    B, D = 48, 8
    mu_c, lv_c, mu_s, lv_s = (torch.randn(B, D, generator=g) * s for s in (1, 0.4, 1, 0.4))
    x = torch.rand(B, 3, 28, 28, generator=g)
    xh = torch.rand(B, 3, 28, 28, generator=g)
    rec, kc, ks = vae_loss(xh, x, mu_c=mu_c, mu_s=mu_s, logvar_c=lv_c, logvar_s=lv_s)
    out.update({"elbo/x": npy(x), "elbo/xhat": npy(xh), "elbo/mu_c": npy(mu_c), "elbo/logvar_c": npy(lv_c),
                "elbo/mu_s": npy(mu_s), "elbo/logvar_s": npy(lv_s), "elbo/recon": npy(rec), "elbo/kl_c": npy(kc),
                "elbo/kl_s": npy(ks)})
    # reparam through VAE.sample
    vae = VAE(16, 3)
    eps = torch.randn(B, D, generator=g)
    with injected(randn_queue=[eps]):
        z = vae.sample(mu_c, lv_c)
    out.update({"sample/eps": npy(eps), "sample/z": npy(z)})
    # estimators
    for cls in (CLUBSample, L1OutUB):
        torch.manual_seed(11)
        est = cls(D, D, 2 * D)
        zc = torch.randn(B, D, generator=g).requires_grad_(True)
        zs = torch.randn(B, D, generator=g).requires_grad_(True)
        perm = torch.randperm(B, generator=g)
        with injected(perm_queue=[perm], cuda_identity=True):
            ub = est(zc, zs)
        ub.backward()
        ll = est.learning_loss(zc.detach(), zs.detach())
        est.zero_grad()
        ll2 = est.learning_loss(zc.detach(), zs.detach())
        ll2.backward()
        n = cls.__name__
        out.update({f"{n}/x": npy(zc), f"{n}/y": npy(zs), f"{n}/perm": npy(perm), f"{n}/forward": npy(ub),
                    f"{n}/dx": npy(zc.grad), f"{n}/dy": npy(zs.grad), f"{n}/learning_loss": npy(ll)})
        for k, v in est.state_dict().items():
            out[f"{n}/state/{k}"] = npy(v)
        for k, p in est.named_parameters():
            out[f"{n}/lgrad/{k}"] = npy(p.grad)
    # TC term + shuffle + BCE
    torch.manual_seed(13)
    fc = torch.nn.Sequential(torch.nn.Linear(16, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1), torch.nn.Sigmoid())
    z = torch.randn(B, 16, generator=g).requires_grad_(True)
    d = fc(z)
    mi = torch.nn.functional.relu(torch.log(d / (1 - d))).mean()
    mi.backward()
    zsh = factor_shuffling(z.detach())
    fl = torch.nn.BCELoss()(torch.cat([fc(z.detach()), fc(zsh)], 0),
                            torch.cat([torch.ones(B, 1), torch.zeros(B, 1)], 0))
    out.update({"tc/z": npy(z), "tc/mi": npy(mi), "tc/dz": npy(z.grad), "tc/shuffled": npy(zsh), "tc/factor_loss": npy(fl)})
    for k, v in fc.state_dict().items():
        out[f"tc/state/{k}"] = npy(v)
    ann = LogisticAnnealer(0, 1, 1 / 8)
    sl = []
    for _ in range(12):
        sl.append(ann.slope())
        ann.step()
    out["annealer/slopes"] = np.array(sl)
    np.savez_compressed(os.path.join(HERE, "heads.npz"), **out)
    print("heads done")


# --------------------------------------------------------------------------
# 3. model forward / backward / one training step from a seeded state
# --------------------------------------------------------------------------
def tensor_digest(t):
    t = t.detach().double().flatten()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()])


def sample_of(t, n=64):
    f = t.detach().flatten()
    step = max(1, f.numel() // n)
    return npy(f[::step][:n])


def model_case(tag, kind, arch, zdim, cin, B, hw, ncls, hyper, est_name=None):
    out = {}
    g = torch.Generator().manual_seed(101)
    X = torch.rand(B, cin, hw, hw, generator=g)
    label = torch.randint(0, ncls, (B,), generator=g)
    style = torch.randint(0, 4, (B,), generator=g)
    D = zdim // 2
    n_fwd = {"clear": 1, "tc": 2, "mim": 6}[kind]
    eps = [torch.randn(B, D, generator=g) for _ in range(2 * n_fwd)]
    perm = torch.randperm(B, generator=g)
    seed = 1234
    torch.manual_seed(seed)
    if kind == "clear":
        tr = get_clearvae_trainer(hyper["beta"], hyper["ps"], hyper["lr"], zdim, hyper["alpha"], hyper["temperature"],
                                  "cpu", arch, cin)
    elif kind == "tc":
        tr = get_cleartcvae_trainer(hyper["beta"], hyper["lambda"], hyper["lr"], hyper["aux_lr"], zdim, hyper["alpha"],
                                    hyper["temperature"], "cpu", arch, cin)
    else:
        tr = get_clearmimvae_trainer(hyper["beta"], est_name, hyper["lambda"], hyper["lr"], hyper["aux_lr"], zdim,
                                     hyper["alpha"], hyper["temperature"], "cpu", arch, cin)
    vae = tr.model
    out["meta"] = np.array([kind, arch, str(zdim), str(cin), str(B), str(hw), str(ncls), str(seed), str(est_name)])
    out["hyper_keys"] = np.array(list(hyper.keys()))
    out["hyper_vals"] = np.array([repr(v) for v in hyper.values()])
    out["X"], out["label"], out["style"] = npy(X), npy(label), npy(style)
    for i, e in enumerate(eps):
        out[f"eps/{i}"] = npy(e)
    out["perm"] = npy(perm)
    for k, v in vae.state_dict().items():
        out[f"init_digest/{k}"] = tensor_digest(v)
    aux = getattr(tr, "factor_cls", None) or getattr(tr, "mi_estimator", None)
    if aux is not None:
        for k, v in aux.state_dict().items():
            out[f"aux_init/{k}"] = npy(v)
    # ---- (a) forward + loss + grads on a deep copy (values the trainer only prints)
    tr2 = copy.deepcopy(tr)
    v2 = tr2.model
    v2.train()
    with injected(randn_queue=eps[:2], perm_queue=[perm], cuda_identity=True):
        xhat, lp, z = v2(X, explicit=True)
        rec, kc, ks = vae_loss(xhat, X, **lp)
        c = contrastive_loss(lp["mu_c"], lp["logvar_c"], label, "cosine", hyper["temperature"])
        slope = tr2.annealer.slope()
        loss = rec + slope * kc + slope * ks + hyper["alpha"] * c
        if kind == "clear":
            s = contrastive_loss(lp["mu_s"], lp["logvar_s"], label, "cosine", hyper["temperature"], ps=hyper["ps"])
            if not hyper["ps"]:
                s = -s
            loss = loss + hyper["alpha"] * s
            out["s_loss"] = npy(s)
        elif kind == "tc":
            d = tr2.factor_cls(z)
            mi = torch.nn.functional.relu(torch.log(d / (1 - d))).mean()
            loss = loss + hyper["lambda"] * mi
            out["mi_loss"] = npy(mi)
        else:
            mi = tr2.mi_estimator(z[:, :D], z[:, D:])
            loss = loss + hyper["lambda"] * mi
            out["mi_loss"] = npy(mi)
    loss.backward()
    out.update({"xhat": npy(xhat) if xhat.numel() <= 200000 else sample_of(xhat, 4096),
                "xhat_digest": tensor_digest(xhat), "z": npy(z), "recon": npy(rec), "kl_c": npy(kc), "kl_s": npy(ks),
                "c_loss": npy(c), "loss": npy(loss), "slope": np.array(slope)})
    for k in ("mu_c", "logvar_c", "mu_s", "logvar_s"):
        out[f"latent/{k}"] = npy(lp[k])
    for k, p in v2.named_parameters():
        gr = p.grad
        out[f"grad_digest/{k}"] = tensor_digest(gr)
        out[f"grad_sample/{k}"] = sample_of(gr, 256)
    for k, b in v2.named_buffers():
        out[f"buf_after_fwd/{k}"] = npy(b)
    # ---- (b) the real trainer loop body on one batch
    logs1, logs2 = [], []
    with injected(randn_queue=eps, perm_queue=[perm], cuda_identity=True):
        if kind == "clear":
            tr._train([(X, label, style)], False, 1)
        elif kind == "tc":
            tr._train([(X, label, style)], False, 1, logs1)
        else:
            tr._train([(X, label, style)], False, 1, logs1, logs2)
    out["train_logs1"] = np.array(logs1, dtype=np.float64)
    out["train_logs2"] = np.array(logs2, dtype=np.float64)
    for k, v in tr.model.state_dict().items():
        out[f"after_digest/{k}"] = tensor_digest(v)
        out[f"after_sample/{k}"] = sample_of(v, 256)
    if aux is not None:
        aux = getattr(tr, "factor_cls", None) or getattr(tr, "mi_estimator", None)
        for k, v in aux.state_dict().items():
            out[f"aux_after/{k}"] = npy(v)
    np.savez_compressed(os.path.join(HERE, f"step_{tag}.npz"), **out)
    print("step", tag, "done; loss", float(loss))


def main():
    contrastive_cases()
    head_cases()
    mn = dict(beta=1 / 8, lr=5e-4, alpha=1e2, temperature=0.1)
    model_case("clear_vae28_ps", "clear", "VAE", 16, 3, 32, 28, 10, dict(mn, ps=True))
    model_case("clear_vae28_nops", "clear", "VAE", 16, 1, 24, 28, 10, dict(mn, ps=False))
    model_case("tc_vae28", "tc", "VAE", 16, 3, 32, 28, 10, dict(mn, **{"lambda": 1, "aux_lr": 1e-4}))
    model_case("mim_club_vae28", "mim", "VAE", 16, 3, 32, 28, 10, dict(mn, **{"lambda": 3, "aux_lr": 2e-3}), "CLUBSample")
    model_case("mim_l1out_vae28", "mim", "VAE", 16, 3, 32, 28, 10, dict(mn, **{"lambda": 3, "aux_lr": 2e-3}), "L1OutUB")
    m64 = dict(beta=1 / 32, lr=3e-5, alpha=1e2, temperature=0.1)
    model_case("clear_vae64_ps", "clear", "VAE64", 64, 3, 8, 64, 7, dict(m64, ps=True))
    model_case("tc_vae64", "tc", "VAE64", 64, 3, 8, 64, 4, dict(m64, **{"lambda": 1, "aux_lr": 1e-4}))


if __name__ == "__main__":
    main()
