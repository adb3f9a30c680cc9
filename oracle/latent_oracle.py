"""CPU oracle for the CLEAR-VAE latent-head loss block (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the arithmetic the reference performs in
`code/src/losses.py`, `code/src/models/vae.py:56-60` and
`code/src/models/mi_estimator.py:108-198`.  It is the *checker* for the CUDA
path: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import it.  The product package
(`clear_vae_b200/`) never imports anything under `oracle/`.

Parity pin: the reference ships no tests, so the oracle is
pinned against (i) outputs of the unmodified reference imported in the build
container (`tests/golden/make_golden.py` -> `tests/golden/*.npz`) and (ii) the
known-answer table recorded in SURVEY.md §4.  `tests/test_oracle_golden.py`
checks both.

Everything is computed in float64 from the given inputs unless `dtype` says
otherwise, row-chunked so that B_g = 65536 fits in host memory (the reference
materialises [B,B,D] tensors and cannot run there).
"""
from __future__ import annotations

import math

import numpy as np

SIM_FNS = ("cosine", "l2", "modified_l2", "jeffrey", "mahalanobis")
LOSS_NAMES = ("snn_loss", "supcon_in_loss", "supcon_out_loss")
_COS_EPS = 1e-8  # F.cosine_similarity default eps (losses.py:55)


# --------------------------------------------------------------------------
# masks / pair indexing (bit-exact part of the contract)
# --------------------------------------------------------------------------
def pair_mask(label_rows, label_cols, ps=False):
    """uint8 [R, C] mask restating losses.py:107-110.

    `ps` follows python truthiness exactly like the reference (`None` and
    `False` select the same-label mask, anything truthy the flipped mask).
    """
    lr = np.asarray(label_rows).reshape(-1, 1)
    lc = np.asarray(label_cols).reshape(1, -1)
    return ((lr != lc) if ps else (lr == lc)).astype(np.uint8)


def positive_sets(label, ps=False, row_offset=0, label_cols=None):
    """(candidate, positive) uint8 masks with the diagonal removed
    (losses.py:131: the diagonal of `sim` is overwritten with -inf *before*
    the mask is applied, so j == i is in neither set)."""
    label = np.asarray(label)
    cols = label if label_cols is None else np.asarray(label_cols)
    m = pair_mask(label, cols, ps)
    cand = np.ones_like(m)
    r = np.arange(label.shape[0])
    cand[r, r + row_offset] = 0
    return cand, (m & cand)


# --------------------------------------------------------------------------
# similarities (losses.py:54-84)
# --------------------------------------------------------------------------
def unit_rows(mu):
    """n_i = mu_i / max(||mu_i||, eps): ATen cosine_similarity semantics."""
    nrm = np.sqrt((mu * mu).sum(-1, keepdims=True))
    return mu / np.maximum(nrm, _COS_EPS), nrm


def similarity_block(sim_fn, mu_r, lv_r, mu_c, lv_c):
    """sim[i, j] for rows i of (mu_r, lv_r) against columns j of (mu_c, lv_c).

    Index convention of the reference: `x[None]` varies along j, `x[:, None]`
    along i (losses.py:54-84).
    """
    if sim_fn == "cosine":
        nr, _ = unit_rows(mu_r)
        nc, _ = unit_rows(mu_c)
        return nr @ nc.T
    diff2 = (mu_c[None, :, :] - mu_r[:, None, :]) ** 2  # [R, C, D]
    if sim_fn == "l2":
        return -diff2.sum(-1)
    if sim_fn == "modified_l2":
        var = np.exp(0.5 * (lv_c[None, :, :] + lv_r[:, None, :]))
        return -(diff2 / var).sum(-1)
    if sim_fn == "mahalanobis":
        var = 0.5 * (np.exp(lv_c)[None, :, :] + np.exp(lv_r)[:, None, :])
        return -(diff2 / var).sum(-1)
    if sim_fn == "jeffrey":
        k = mu_r.shape[1]
        var_r, var_c = np.exp(lv_r), np.exp(lv_c)
        ls_r, ls_c = lv_r.sum(-1), lv_c.sum(-1)
        # kl[i, j] as coded at losses.py:62-69 (divisor / numerator on the j side)
        kl_ij = 0.5 * (
            (ls_c[None, :] - ls_r[:, None] - k)
            + (diff2 / var_c[None, :, :]).sum(-1)
            + (var_c[None, :, :] / (var_r[:, None, :] + 1e-8)).sum(-1)
        )
        # kl[j, i]: same expression with roles swapped
        kl_ji = 0.5 * (
            (ls_r[:, None] - ls_c[None, :] - k)
            + (diff2 / var_r[:, None, :]).sum(-1)
            + (var_r[:, None, :] / (var_c[None, :, :] + 1e-8)).sum(-1)
        )
        return -0.5 * (kl_ij + kl_ji)
    raise ValueError("unimplemented similarity measure.")  # losses.py:122-123


# --------------------------------------------------------------------------
# masked log-sum-exp rows (losses.py:87-95, 129-170)
# --------------------------------------------------------------------------
def _lse_rows(x):
    """Row LSE where an all -inf row gives -inf (losses.py:87-95)."""
    m = x.max(axis=1)
    dead = np.isneginf(m)
    m0 = np.where(dead, 0.0, m)
    with np.errstate(divide="ignore"):
        s = np.exp(x - m0[:, None]).sum(axis=1)
        s = np.where(dead, 1.0, s)
        return np.log(s) + np.where(dead, -np.inf, m0)


def row_losses(
    mu,
    logvar,
    label,
    sim_fn,
    temperature,
    loss_name="snn_loss",
    ps=False,
    *,
    mu_cols=None,
    logvar_cols=None,
    label_cols=None,
    row_offset=0,
    dtype=np.float64,
    chunk=1024,
    return_stats=False,
):
    """Per-row loss l_i (inf / nan where the reference produces them).

    Rows are (mu, logvar, label); columns default to the same arrays but may be
    a larger *global* batch (`*_cols`) in which case row i is global row
    `row_offset + i` (data-parallel restatement).
    """
    if sim_fn not in SIM_FNS:
        raise ValueError("unimplemented similarity measure.")
    if loss_name not in LOSS_NAMES:
        raise NameError(loss_name)
    mu = np.asarray(mu, dtype=dtype)
    logvar = np.asarray(logvar, dtype=dtype)
    label = np.asarray(label).reshape(-1)
    mu_c = mu if mu_cols is None else np.asarray(mu_cols, dtype=dtype)
    lv_c = logvar if logvar_cols is None else np.asarray(logvar_cols, dtype=dtype)
    lab_c = label if label_cols is None else np.asarray(label_cols).reshape(-1)
    R = mu.shape[0]
    out = np.empty(R, dtype=dtype)
    lse_all = np.empty(R, dtype=dtype)
    lse_pos = np.empty(R, dtype=dtype)
    t = dtype(temperature)
    if sim_fn != "cosine":
        chunk = max(1, min(chunk, (1 << 27) // max(1, mu_c.shape[0] * mu.shape[1])))
    for r0 in range(0, R, chunk):
        r1 = min(R, r0 + chunk)
        sim = similarity_block(sim_fn, mu[r0:r1], logvar[r0:r1], mu_c, lv_c)
        m = pair_mask(label[r0:r1], lab_c, ps).astype(bool)
        rr = np.arange(r1 - r0)
        diag_cols = rr + r0 + row_offset
        if loss_name == "supcon_out_loss":
            # losses.py:156-170: diag <- -999, mean of positive sims (unscaled)
            sim[rr, diag_cols] = -999.0
            pm = m.copy()
            pm[rr, diag_cols] = False
            n_k = pm.sum(1).astype(dtype)
            la = _lse_rows(sim / t)
            with np.errstate(divide="ignore", invalid="ignore"):
                val = -(np.where(pm, sim, 0.0).sum(1)) / n_k + la
            # rows with n_k == 0 are *selected out* (not merely non-finite)
            val = np.where(n_k > 0, val, np.inf)
            out[r0:r1] = val
            lse_all[r0:r1] = la
            lse_pos[r0:r1] = np.nan
            continue
        sim[rr, diag_cols] = -np.inf
        pos = np.where(m, sim, -np.inf)
        la = _lse_rows(sim / t)
        lp = _lse_rows(pos / t)
        with np.errstate(invalid="ignore", divide="ignore"):
            val = -lp + la
            if loss_name == "supcon_in_loss":
                n_k = m.sum(1).astype(dtype) - 1.0  # losses.py:141 (uses undiagonalised mask)
                val = np.log(n_k) + val
        out[r0:r1] = val
        lse_all[r0:r1] = la
        lse_pos[r0:r1] = lp
    if return_stats:
        return out, lse_all, lse_pos
    return out


def contrastive(
    mu, logvar, label, sim_fn, temperature, loss_name="snn_loss", ps=False, **kw
):
    """`contrastive_loss` of losses.py:98-126: mean over the finite rows."""
    rows = row_losses(mu, logvar, label, sim_fn, temperature, loss_name, ps, **kw)
    fin = np.isfinite(rows)
    if not fin.any():
        return float("nan")
    return float(rows[fin].sum() / fin.sum())


def contrastive_partial(mu, logvar, label, sim_fn, temperature, loss_name="snn_loss", ps=False, **kw):
    """(sum of finite rows, number of finite rows) — what one data-parallel rank
    contributes before the 2-scalar all-reduce."""
    rows = row_losses(mu, logvar, label, sim_fn, temperature, loss_name, ps, **kw)
    fin = np.isfinite(rows)
    return float(rows[fin].sum()), int(fin.sum())


# --------------------------------------------------------------------------
# closed-form gradient of the SNN loss (cosine / l2)  — SURVEY.md §8a'
# --------------------------------------------------------------------------
def snn_grad(mu, label, sim_fn, temperature, ps=False, *, upstream=1.0, dtype=np.float64, chunk=1024):
    """d(upstream * contrastive)/d mu for sim_fn in {cosine, l2}, single shard.

    G_ij = [i in F]/(tau |F|) (a_ij - p_ij);  grad wrt the similarity operand
    collects both the row side (G) and the column side (G^T).
    """
    assert sim_fn in ("cosine", "l2")
    mu = np.asarray(mu, dtype=dtype)
    label = np.asarray(label).reshape(-1)
    B, D = mu.shape
    rows, la, lp = row_losses(mu, np.zeros_like(mu), label, sim_fn, temperature, "snn_loss", ps,
                              dtype=dtype, chunk=chunk, return_stats=True)
    fin = np.isfinite(rows)
    nF = int(fin.sum())
    if nF == 0:
        return np.full_like(mu, np.nan)
    w = np.where(fin, upstream / (temperature * nF), 0.0)
    if sim_fn == "cosine":
        X, nrm = unit_rows(mu)
    else:
        X = mu
    dX = np.zeros_like(mu)
    t = dtype(temperature)
    for r0 in range(0, B, chunk):
        r1 = min(B, r0 + chunk)
        sim = similarity_block(sim_fn, mu[r0:r1], None, mu, None)
        rr = np.arange(r1 - r0)
        sim[rr, rr + r0] = -np.inf
        m = pair_mask(label[r0:r1], label, ps).astype(bool)
        with np.errstate(invalid="ignore", over="ignore"):
            a = np.exp(sim / t - la[r0:r1, None])
            p = np.where(m, np.exp(sim / t - lp[r0:r1, None]), 0.0)
        a[rr, rr + r0] = 0.0
        p[rr, rr + r0] = 0.0
        p[~fin[r0:r1]] = 0.0
        G = w[r0:r1, None] * (a - p)  # [r, B]
        if sim_fn == "cosine":
            dX[r0:r1] += G @ X  # row side
            dX += G.T @ X[r0:r1]  # column side
        else:  # s_ij = -|x_i - x_j|^2
            gs = G.sum(1)
            dX[r0:r1] += -2.0 * (gs[:, None] * X[r0:r1] - G @ X)
            dX += -2.0 * (G.sum(0)[:, None] * X - G.T @ X[r0:r1])
    if sim_fn == "cosine":
        # chain through n = mu / max(|mu|, eps) (clamp carries no gradient)
        den = np.maximum(nrm, _COS_EPS)
        live = (nrm >= _COS_EPS).astype(dtype)
        return (dX - live * X * (X * dX).sum(-1, keepdims=True)) / den
    return dX


# --------------------------------------------------------------------------
# reparameterisation / ELBO terms (vae.py:56-60, losses.py:36-50)
# --------------------------------------------------------------------------
def reparam(mu, logvar, eps):
    return mu + eps * np.exp(0.5 * logvar)


def gaussian_kl(mu, logvar):
    """-1/2 mean_b sum_d (1 + lv - mu^2 - e^lv)   (losses.py:48-49)."""
    mu = np.asarray(mu, dtype=np.float64)
    logvar = np.asarray(logvar, dtype=np.float64)
    return float(-0.5 * (1.0 + logvar - mu * mu - np.exp(logvar)).sum(-1).mean())


def recon_sse(xhat, x):
    """mean_b sum_{chw} (xhat - x)^2   (losses.py:45-47)."""
    d = np.asarray(xhat, dtype=np.float64) - np.asarray(x, dtype=np.float64)
    return float((d * d).reshape(d.shape[0], -1).sum(-1).mean())


# --------------------------------------------------------------------------
# MI / TC heads, given the estimator-network outputs (mu_q, lv_q)
# --------------------------------------------------------------------------
def club_sample_bound(mu_q, lv_q, y, perm):
    """CLUBSample.forward (mi_estimator.py:133-143) with the permutation injected."""
    mu_q, lv_q, y = (np.asarray(a, dtype=np.float64) for a in (mu_q, lv_q, y))
    iv = np.exp(-lv_q)
    pos = -((mu_q - y) ** 2) * iv
    neg = -((mu_q - y[np.asarray(perm)]) ** 2) * iv
    return float((pos.sum(-1) - neg.sum(-1)).mean() / 2.0)


def gaussian_learning_loss(mu_q, lv_q, y):
    """learning_loss = -loglikeli (mi_estimator.py:129-131,145-146,193-198)."""
    mu_q, lv_q, y = (np.asarray(a, dtype=np.float64) for a in (mu_q, lv_q, y))
    return float(-((-((mu_q - y) ** 2) / np.exp(lv_q) - lv_q).sum(-1).mean()))


def l1out_bound_as_executed(mu_q, lv_q, y):
    """L1OutUB.forward *as the reference executes it* (mi_estimator.py:170-191).

    `all_probs [B,B] + diag_mask [B,B,1]` broadcasts to [B,B,B]; the logsumexp
    over dim 0 therefore only adds log((B-1)+e^-20) to every all_probs entry,
    and `positive [B] - negative [B,B]` broadcasts before the mean.  Net:
      mean_c ap_cc - mean_{b,c} ap_bc - log1p(e^-20/(B-1)).
    (checked against the patched reference in make_golden.py.)
    """
    mu_q, lv_q, y = (np.asarray(a, dtype=np.float64) for a in (mu_q, lv_q, y))
    B = y.shape[0]
    iv = np.exp(-lv_q)
    ap = (-((y[None, :, :] - mu_q[:, None, :]) ** 2) * 0.5 * iv[:, None, :]
          - 0.5 * lv_q[:, None, :]).sum(-1)  # ap[b, c]
    positive = np.diagonal(ap)
    return float(positive.mean() - ap.mean() - math.log1p(math.exp(-20.0) / (B - 1.0)))


def l1out_bound_bruteforce(mu_q, lv_q, y):
    """Literal [B,B,B] evaluation (tiny B only) used to pin the closed form."""
    mu_q, lv_q, y = (np.asarray(a, dtype=np.float64) for a in (mu_q, lv_q, y))
    B = y.shape[0]
    iv = np.exp(-lv_q)
    positive = (-((mu_q - y) ** 2) * 0.5 * iv - 0.5 * lv_q).sum(-1)
    ap = (-((y[None, :, :] - mu_q[:, None, :]) ** 2) * 0.5 * iv[:, None, :]
          - 0.5 * lv_q[:, None, :]).sum(-1)
    cube = ap[None, :, :] + (np.eye(B) * -20.0)[:, :, None]  # [a, b, c]
    mx = cube.max(0)
    negative = np.log(np.exp(cube - mx[None]).sum(0)) + mx - math.log(B - 1.0)  # [b, c]
    return float((positive[None, :] - negative).mean())


def tc_penalty(d_score):
    """relu(log(d / (1 - d))).mean()  (trainer.py:664-665)."""
    d = np.asarray(d_score, dtype=np.float64)
    with np.errstate(divide="ignore"):
        return float(np.maximum(np.log(d / (1.0 - d)), 0.0).mean())


def roll_style_half(z):
    """factor_shuffling(z, 'permute_1') (trainer.py:573-587): rows of the style
    half move up by one, row 0 goes last."""
    z = np.asarray(z)
    d = z.shape[1] // 2
    return np.concatenate([z[:, :d], np.roll(z[:, d:], -1, axis=0)], axis=1)


def logistic_anneal(step, loc, scale, beta):
    """LogisticAnnealer.slope (trainer.py:29-34)."""
    return beta / (1.0 + math.exp(-(step - loc) / scale))


# --------------------------------------------------------------------------
# ML-VAE / GVAE group evidence (models/vae.py:159-223), numpy fp64
# --------------------------------------------------------------------------
def group_evidence(mu, logvar, label, mode):
    """accumulate_group_evidence: per sorted unique label g, MLVAE  mu_g = sum mu e^{-lv} / sum e^{-lv}, lv_g = -LSE(-lv)
    (vae.py:174-180);  GVAE  mu_g = mean mu, lv_g = LSE(lv) - log n (vae.py:181-186).  Returns (mu_g, lv_g, groups, gid)."""
    mu, logvar = np.asarray(mu, np.float64), np.asarray(logvar, np.float64)
    groups, gid = np.unique(np.asarray(label), return_inverse=True)
    mg, lg = np.zeros((len(groups), mu.shape[1])), np.zeros((len(groups), mu.shape[1]))
    for g in range(len(groups)):
        m, l = mu[gid == g], logvar[gid == g]
        if mode == "MLVAE":
            a = -l
            L = np.log(np.exp(a - a.max(0)).sum(0)) + a.max(0)
            mg[g], lg[g] = (m * np.exp(a - L)).sum(0), -L
        elif mode == "GVAE":
            mg[g] = m.mean(0)
            lg[g] = np.log(np.exp(l - l.max(0)).sum(0)) + l.max(0) - np.log(len(m))
        else:
            raise NotImplementedError("only support using MLVAE or GVAE")
    return mg, lg, groups, gid


def group_evidence_grad(mu, logvar, label, mode, dmu_g, dlv_g):
    """gradient of sum(mu_g * dmu_g) + sum(lv_g * dlv_g) w.r.t. (mu, logvar) — closed form used by the CUDA backward"""
    mu, logvar = np.asarray(mu, np.float64), np.asarray(logvar, np.float64)
    mg, lg, groups, gid = group_evidence(mu, logvar, label, mode)
    n = np.bincount(gid).astype(np.float64)
    if mode == "MLVAE":
        p = np.exp(lg[gid] - logvar)
        return p * dmu_g[gid], p * (dlv_g[gid] - dmu_g[gid] * (mu - mg[gid]))
    return dmu_g[gid] / n[gid, None], dlv_g[gid] * np.exp(logvar - lg[gid]) / n[gid, None]


def group_reparam(mu_g, lv_g, gid, eps):
    """groupwise_reparam_each in the original row order (vae.py:196-221): z_i = mu_g(i) + eps_i exp(lv_g(i) / 2)"""
    return mu_g[gid] + eps * np.exp(0.5 * lv_g[gid])
