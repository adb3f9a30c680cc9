"""CPU oracle for the CLEAR-VAE model + training step (TEST INFRASTRUCTURE ONLY).

A functional torch-on-CPU restatement of what the reference executes per
training step: `code/src/models/vae.py:7-156` (VAE / VAE64 layer stacks),
`code/src/losses.py:36-137`, `code/src/models/mi_estimator.py:108-198`,
`code/src/trainer.py:22-38, 435-493, 573-587, 629-709, 820-897` and the default
`torch.optim.Adam` the factories build (`code/src/utils/trainer_utils.py:100,
139-140,178-181`).  The model is expressed as a flat `state` dict keyed exactly
like the reference's `state_dict()` so weights move both ways.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module; the product never does.
Pinned against the unmodified reference through `tests/golden/*.npz`
(`tests/golden/make_golden.py`); see `tests/test_oracle_golden.py`.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------
# architecture tables (vae.py:15-46 and vae.py:113-156)
# --------------------------------------------------------------------------
def arch_spec(arch: str, in_channel: int):
    """(encoder convs, unflatten shape, decoder convTs); conv = (idx, cin, cout, k, s, p[, op])."""
    if arch == "VAE":
        enc = [(0, in_channel, 32, 3, 2, 1), (3, 32, 64, 3, 2, 1), (6, 64, 128, 3, 2, 1)]
        unflat = (128, 4, 4)
        dec = [(4, 128, 64, 3, 2, 1, 0), (7, 64, 32, 3, 2, 1, 1), (10, 32, in_channel, 3, 2, 1, 1)]
    elif arch == "VAE64":
        chans = [in_channel, 32, 64, 128, 256, 512]
        enc = [(3 * i, chans[i], chans[i + 1], 4, 2, 1) for i in range(5)]
        unflat = (512, 2, 2)
        rc = chans[::-1]
        dec = [(4 + 3 * i, rc[i], rc[i + 1], 4, 2, 1, 0) for i in range(5)]
    else:
        raise ValueError(arch)
    return enc, unflat, dec


def init_state(arch: str, total_z_dim: int, in_channel: int, seed: int = 0, dtype=torch.float32):
    """Random state with the reference's key names / shapes (NOT its init RNG
    stream; parity tests load the reference's own weights instead)."""
    g = torch.Generator().manual_seed(seed)
    enc, unflat, dec = arch_spec(arch, in_channel)
    D = int(total_z_dim / 2)
    st = {}

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * b).to(dtype)

    def bn(prefix, c):
        st[f"{prefix}.weight"] = torch.ones(c, dtype=dtype)
        st[f"{prefix}.bias"] = torch.zeros(c, dtype=dtype)
        st[f"{prefix}.running_mean"] = torch.zeros(c, dtype=dtype)
        st[f"{prefix}.running_var"] = torch.ones(c, dtype=dtype)
        st[f"{prefix}.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)

    for (i, ci, co, k, s, p) in enc:
        st[f"encoder.{i}.weight"] = uni((co, ci, k, k), ci * k * k)
        st[f"encoder.{i}.bias"] = uni((co,), ci * k * k)
        bn(f"encoder.{i + 1}", co)
    for h in ("mu_c", "logvar_c", "mu_s", "logvar_s"):
        st[f"{h}.weight"] = uni((D, 2048), 2048)
        st[f"{h}.bias"] = uni((D,), 2048)
    st["decoder.0.weight"] = uni((2048, 2 * D), 2 * D)
    st["decoder.0.bias"] = uni((2048,), 2 * D)
    bn("decoder.1", 2048)
    for (i, ci, co, k, s, p, op) in dec:
        st[f"decoder.{i}.weight"] = uni((ci, co, k, k), co * k * k)
        st[f"decoder.{i}.bias"] = uni((co,), co * k * k)
        bn(f"decoder.{i + 1}", co)
    return st


def is_param(key: str) -> bool:
    return not (key.endswith("running_mean") or key.endswith("running_var") or key.endswith("num_batches_tracked"))


# --------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------
def _bn(st, prefix, h, train):
    """BatchNorm{1,2}d: biased batch variance for normalisation, unbiased for the
    running estimate, momentum 0.1, eps 1e-5 (SURVEY.md §8a' layer semantics)."""
    if train:
        out = F.batch_norm(h, st[f"{prefix}.running_mean"], st[f"{prefix}.running_var"],
                           st[f"{prefix}.weight"], st[f"{prefix}.bias"], True, BN_MOMENTUM, BN_EPS)
        st[f"{prefix}.num_batches_tracked"] += 1
        return out
    return F.batch_norm(h, st[f"{prefix}.running_mean"], st[f"{prefix}.running_var"],
                        st[f"{prefix}.weight"], st[f"{prefix}.bias"], False, BN_MOMENTUM, BN_EPS)


def encode(st, x, arch, in_channel, train=True):
    enc, _, _ = arch_spec(arch, in_channel)
    h = x
    for (i, ci, co, k, s, p) in enc:
        h = F.conv2d(h, st[f"encoder.{i}.weight"], st[f"encoder.{i}.bias"], stride=s, padding=p)
        h = torch.relu(_bn(st, f"encoder.{i + 1}", h, train))
    h = h.flatten(1)
    return tuple(F.linear(h, st[f"{n}.weight"], st[f"{n}.bias"]) for n in ("mu_c", "logvar_c", "mu_s", "logvar_s"))


def decode(st, z, arch, in_channel, train=True):
    _, unflat, dec = arch_spec(arch, in_channel)
    h = F.linear(z, st["decoder.0.weight"], st["decoder.0.bias"])
    h = torch.relu(_bn(st, "decoder.1", h, train)).unflatten(1, unflat)
    for n, (i, ci, co, k, s, p, op) in enumerate(dec):
        h = F.conv_transpose2d(h, st[f"decoder.{i}.weight"], st[f"decoder.{i}.bias"], stride=s, padding=p,
                               output_padding=op)
        h = _bn(st, f"decoder.{i + 1}", h, train)
        h = torch.sigmoid(h) if n == len(dec) - 1 else torch.relu(h)
    return h


def forward(st, x, eps_c, eps_s, arch, in_channel, train=True):
    """VAE.forward with label=None (vae.py:81-102); noise injected (c first, then s)."""
    mu_c, lv_c, mu_s, lv_s = encode(st, x, arch, in_channel, train)
    z_c = mu_c + eps_c * torch.exp(0.5 * lv_c)
    z_s = mu_s + eps_s * torch.exp(0.5 * lv_s)
    z = torch.cat([z_c, z_s], dim=-1)
    xhat = decode(st, z, arch, in_channel, train)
    return xhat, dict(mu_c=mu_c, logvar_c=lv_c, mu_s=mu_s, logvar_s=lv_s), z


# --------------------------------------------------------------------------
# losses (differentiable restatement)
# --------------------------------------------------------------------------
def elbo_terms(xhat, x, mu_c, mu_s, logvar_c, logvar_s):
    rec = ((xhat - x) ** 2).flatten(1).sum(1).mean()
    kl = lambda m, lv: -0.5 * (1 + lv - m * m - lv.exp()).sum(1).mean()
    return rec, kl(mu_c, logvar_c), kl(mu_s, logvar_s)


def similarity(sim_fn, mu, lv):
    if sim_fn == "cosine":
        nrm = mu.norm(dim=1, keepdim=True)
        n = mu / torch.where(nrm < 1e-8, torch.full_like(nrm, 1e-8), nrm)  # clamp_min(eps), no grad when clamped
        return n @ n.T
    d2 = (mu[None, :, :] - mu[:, None, :]) ** 2
    if sim_fn == "l2":
        return -d2.sum(-1)
    if sim_fn == "modified_l2":
        return -(d2 / (0.5 * (lv[None] + lv[:, None])).exp()).sum(-1)
    if sim_fn == "mahalanobis":
        return -(d2 / (0.5 * (lv.exp()[None] + lv.exp()[:, None]))).sum(-1)
    if sim_fn == "jeffrey":
        var = lv.exp()
        kl = 0.5 * ((lv.sum(1)[None, :] - lv.sum(1)[:, None] - mu.shape[1])
                    + (d2 / var[None]).sum(-1) + (var[None] / (var[:, None] + 1e-8)).sum(-1))
        return -0.5 * (kl + kl.T)
    raise ValueError("unimplemented similarity measure.")


def contrastive(mu, lv, label, sim_fn, temperature, loss_name="snn_loss", ps=False):
    ninf = float("-inf")
    B = mu.shape[0]
    same = label[None, :] == label[:, None]
    m = ~same if ps else same
    eye = torch.eye(B, dtype=torch.bool, device=mu.device)
    s = similarity(sim_fn, mu, lv)
    if loss_name == "supcon_out_loss":
        s = s.masked_fill(eye, -999.0)
        pm = m & ~eye
        nk = pm.sum(1)
        rows = -(s * pm).sum(1) / nk + torch.logsumexp(s / temperature, dim=1)
        rows = rows[nk > 0]
    else:
        s = s.masked_fill(eye, ninf)
        rows = -torch.logsumexp(s.masked_fill(~m, ninf) / temperature, dim=1) + torch.logsumexp(s / temperature, dim=1)
        if loss_name == "supcon_in_loss":
            rows = (m.sum(1).to(rows.dtype) - 1).log() + rows
        elif loss_name != "snn_loss":
            raise NameError(loss_name)
    rows = rows[torch.isfinite(rows)]
    return rows.mean()


def mlp2(st, prefix, x, tanh=False):
    h = torch.relu(F.linear(x, st[f"{prefix}.0.weight"], st[f"{prefix}.0.bias"]))
    h = F.linear(h, st[f"{prefix}.2.weight"], st[f"{prefix}.2.bias"])
    return torch.tanh(h) if tanh else h


def init_estimator_state(x_dim, y_dim, hidden, seed=0, dtype=torch.float32):
    """CLUBSample / L1OutUB parameter shapes (mi_estimator.py:108-124,149-165)."""
    g = torch.Generator().manual_seed(seed)
    st = {}
    for pre in ("p_mu", "p_logvar"):
        for idx, (o, i) in ((0, (hidden // 2, x_dim)), (2, (y_dim, hidden // 2))):
            b = 1.0 / math.sqrt(i)
            st[f"{pre}.{idx}.weight"] = ((torch.rand((o, i), generator=g, dtype=torch.float64) * 2 - 1) * b).to(dtype)
            st[f"{pre}.{idx}.bias"] = ((torch.rand((o,), generator=g, dtype=torch.float64) * 2 - 1) * b).to(dtype)
    return st


def estimator_heads(est, x):
    return mlp2(est, "p_mu", x), mlp2(est, "p_logvar", x, tanh=True)


def club_sample(est, x, y, perm):
    mu, lv = estimator_heads(est, x)
    iv = (-lv).exp()
    return ((-(mu - y) ** 2 * iv).sum(-1) - (-(mu - y[perm]) ** 2 * iv).sum(-1)).mean() / 2.0


def l1out(est, x, y):
    """L1OutUB.forward as executed (closed form, see latent_oracle.l1out_bound_as_executed)."""
    mu, lv = estimator_heads(est, x)
    B = y.shape[0]
    iv = (-lv).exp()
    ap = (-((y[None, :, :] - mu[:, None, :]) ** 2) * 0.5 * iv[:, None, :] - 0.5 * lv[:, None, :]).sum(-1)
    return ap.diagonal().mean() - ap.mean() - math.log1p(math.exp(-20.0) / (B - 1.0))


def estimator_learning_loss(est, x, y):
    mu, lv = estimator_heads(est, x)
    return -((-(mu - y) ** 2 / lv.exp() - lv).sum(1).mean())


def init_factor_state(total_z_dim, seed=0, dtype=torch.float32):
    """factor_cls = Linear(Z,Z)-ReLU-Linear(Z,1)-Sigmoid (trainer_utils.py:133-138)."""
    g = torch.Generator().manual_seed(seed)
    st = {}
    for idx, (o, i) in ((0, (total_z_dim, total_z_dim)), (2, (1, total_z_dim))):
        b = 1.0 / math.sqrt(i)
        st[f"{idx}.weight"] = ((torch.rand((o, i), generator=g, dtype=torch.float64) * 2 - 1) * b).to(dtype)
        st[f"{idx}.bias"] = ((torch.rand((o,), generator=g, dtype=torch.float64) * 2 - 1) * b).to(dtype)
    return st


def factor_score(fst, z):
    h = torch.relu(F.linear(z, fst["0.weight"], fst["0.bias"]))
    return torch.sigmoid(F.linear(h, fst["2.weight"], fst["2.bias"]))


def roll_style(z):
    d = z.shape[1] // 2
    return torch.cat([z[:, :d], torch.roll(z[:, d:], -1, 0)], 1)


# --------------------------------------------------------------------------
# Adam (torch.optim.Adam defaults: betas (0.9, 0.999), eps 1e-8, no decay)
# --------------------------------------------------------------------------
class Adam:
    def __init__(self, keys, lr, betas=(0.9, 0.999), eps=1e-8):
        self.keys, self.lr, self.b1, self.b2, self.eps = list(keys), lr, betas[0], betas[1], eps
        self.t = 0
        self.m, self.v = {}, {}

    @torch.no_grad()
    def step(self, st, grads):
        self.t += 1
        c1 = 1.0 - self.b1 ** self.t
        c2 = 1.0 - self.b2 ** self.t
        for k in self.keys:
            g = grads.get(k)
            if g is None:
                continue
            if k not in self.m:
                self.m[k] = torch.zeros_like(g)
                self.v[k] = torch.zeros_like(g)
            self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            denom = (self.v[k].sqrt() / math.sqrt(c2)).add_(self.eps)
            st[k].addcdiv_(self.m[k], denom, value=-self.lr / c1)


def _grads(loss, st, keys):
    ps = [st[k] for k in keys]
    gs = torch.autograd.grad(loss, ps, allow_unused=True)
    return {k: g for k, g in zip(keys, gs)}


# --------------------------------------------------------------------------
# training steps
# --------------------------------------------------------------------------
class StepOracle:
    """One object per trainer flavour; `step()` performs exactly the work of one
    iteration of the matching `_train` loop body and returns the logged scalars.

    kind: 'clear' (trainer.py:446-492), 'tc' (:646-709), 'mim' (:841-897).
    Noise (`eps`) and CLUB-S permutations are injected by the caller so the
    CUDA path can be compared on identical draws; when omitted they are drawn
    from `gen`.
    """

    def __init__(self, kind, st, arch, in_channel, hyper, vae_lr, *, aux=None, aux_lr=None, estimator="CLUBSample",
                 sim_fn="cosine", gen=None):
        self.kind, self.st, self.arch, self.cin, self.h = kind, st, arch, in_channel, dict(hyper)
        self.sim_fn = sim_fn
        self.pkeys = [k for k in st if is_param(k)]
        for k in self.pkeys:
            st[k].requires_grad_(True)
        self.opt = Adam(self.pkeys, vae_lr)
        self.aux = aux
        self.estimator = estimator
        if aux is not None:
            for k in aux:
                aux[k].requires_grad_(True)
            self.aux_opt = Adam(list(aux), aux_lr)
        self.t = 0
        self.gen = gen or torch.Generator().manual_seed(0)

    def slope(self):
        return self.h["beta"] / (1.0 + math.exp(-(self.t - self.h.get("loc", 0)) / self.h.get("scale", 1)))

    def _noise(self, B, D, dtype, eps):
        if eps is not None:
            return eps
        return (torch.randn(B, D, generator=self.gen, dtype=dtype), torch.randn(B, D, generator=self.gen, dtype=dtype))

    def step(self, x, label, eps=None, extra_eps=None, perm=None):
        st, h = self.st, self.h
        D = st["mu_c.weight"].shape[0]
        B = x.shape[0]
        e_c, e_s = self._noise(B, D, x.dtype, eps)
        xhat, lp, z = forward(st, x, e_c, e_s, self.arch, self.cin, True)
        rec, kl_c, kl_s = elbo_terms(xhat, x, **lp)
        c = contrastive(lp["mu_c"], lp["logvar_c"], label, self.sim_fn, h["temperature"])
        beta_t = self.slope()
        logs = dict(recon=float(rec), kl_c=float(kl_c), kl_s=float(kl_s), c_loss=float(c))
        loss = rec + beta_t * kl_c + beta_t * kl_s + h["alpha"] * c
        if self.kind == "clear":
            ps = h.get("ps")
            s = contrastive(lp["mu_s"], lp["logvar_s"], label, self.sim_fn, h["temperature"], ps=ps)
            if not ps:
                s = -s
            loss = loss + h["alpha"] * s
            logs["s_loss"] = float(s)
        elif self.kind == "tc":
            d = factor_score(self.aux, z)
            mi = torch.relu(torch.log(d / (1 - d))).mean()
            loss = loss + h["lambda"] * mi
            logs["mi_loss"] = float(mi)
        elif self.kind == "mim":
            if self.estimator == "CLUBSample":
                if perm is None:
                    perm = torch.randperm(B, generator=self.gen)
                mi = club_sample(self.aux, z[:, :D], z[:, D:], perm)
            else:
                mi = l1out(self.aux, z[:, :D], z[:, D:])
            loss = loss + h["lambda"] * mi
            logs["mi_loss"] = float(mi)
        grads = _grads(loss, st, self.pkeys)
        self.last_grads = grads
        self.opt.step(st, grads)
        self.t += 1
        logs["loss"] = float(loss)
        # auxiliary-network phases
        if self.kind == "tc":
            ee = self._noise(B, D, x.dtype, extra_eps[0] if extra_eps else None)
            with torch.no_grad():
                _, _, z2 = forward(st, x, ee[0], ee[1], self.arch, self.cin, True)
            dj = factor_score(self.aux, z2)
            dm = factor_score(self.aux, roll_style(z2))
            fl = F.binary_cross_entropy(torch.cat([dj, dm], 0), torch.cat([torch.ones_like(dj), torch.zeros_like(dm)], 0))
            self.aux_opt.step(self.aux, _grads(fl, self.aux, list(self.aux)))
            logs["factor_loss"] = float(fl)
        elif self.kind == "mim":
            ll = []
            for j in range(5):
                ee = self._noise(B, D, x.dtype, extra_eps[j] if extra_eps else None)
                with torch.no_grad():
                    _, _, z2 = forward(st, x, ee[0], ee[1], self.arch, self.cin, True)
                l = estimator_learning_loss(self.aux, z2[:, :D], z2[:, D:])
                self.aux_opt.step(self.aux, _grads(l, self.aux, list(self.aux)))
                ll.append(float(l))
            logs["mi_learning"] = ll
        return logs
