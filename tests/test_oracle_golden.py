"""CPU: pin the oracle (oracle/latent_oracle.py) against vectors produced by the
unmodified reference (tests/golden/make_golden.py) and SURVEY.md §4's table."""
import os

import numpy as np
import pytest

from oracle import latent_oracle as lo

KAT = {  # SURVEY.md §4, recorded from the unmodified reference (CPU, fp32)
    ("cosine", False): 6.940875053405762, ("cosine", True): 0.1634926199913025,
    ("l2", False): 45.183048248291016, ("l2", True): 6.747245788574219e-05,
    ("jeffrey", False): 22.141494750976562, ("jeffrey", True): 5.956801414489746,
    ("mahalanobis", False): 33.96194839477539, ("mahalanobis", True): 3.0064749717712402,
    ("modified_l2", False): 35.615291595458984, ("modified_l2", True): 2.4577407836914062,
}


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "contrastive.npz"))
    for name in g["names"]:
        sim, tau, ln, ps = g[f"{name}/meta"]
        yield name, g, sim, float(tau), ln, eval(ps)


def _close(a, b, rel=1e-5, ab=2e-6):
    if np.isnan(b):
        return np.isnan(a)
    return abs(a - b) <= rel * abs(b) + ab


def test_survey_known_answers(golden_dir):
    g = np.load(os.path.join(golden_dir, "contrastive.npz"))
    for (sim, ps), want in KAT.items():
        name = f"kat_{sim}_ps{int(ps)}"
        got = lo.contrastive(g[f"{name}/mu"], g[f"{name}/logvar"], g[f"{name}/label"], sim, 0.1, ps=ps)
        assert _close(got, want), (name, got, want)
        assert _close(float(g[f"{name}/loss"]), want, rel=1e-6), name


def test_contrastive_values_match_reference(golden_dir):
    n = 0
    for name, g, sim, tau, ln, ps in _cases(golden_dir):
        got = lo.contrastive(g[f"{name}/mu"], g[f"{name}/logvar"], g[f"{name}/label"], sim, tau, ln, ps)
        want64 = float(g[f"{name}/loss64"])
        want32 = float(g[f"{name}/loss"])
        # fp64 oracle == fp64 reference (supcon_in keeps an fp32 log(n_k) in the reference: pair_mat is .float())
        tol = 1e-7 if ln == "supcon_in_loss" else 1e-9
        assert _close(got, want64, rel=tol, ab=tol), (name, got, want64)
        assert _close(got, want32, rel=1e-5, ab=5e-6), (name, got, want32)  # and inside the fp32 gate
        n += 1
    assert n >= 40


def test_contrastive_gradients_match_reference(golden_dir):
    n = 0
    for name, g, sim, tau, ln, ps in _cases(golden_dir):
        if sim not in ("cosine", "l2") or ln != "snn_loss":
            continue
        want = g[f"{name}/dmu"]
        if want.size == 0 or not np.isfinite(float(g[f"{name}/loss"])):
            continue
        got = lo.snn_grad(g[f"{name}/mu"], g[f"{name}/label"], sim, tau, ps)
        scale = np.abs(want).max() + 1e-12
        assert np.abs(got - want).max() <= 1e-4 * scale + 1e-7, (name, np.abs(got - want).max(), scale)
        n += 1
    assert n >= 20


def test_supcon_gradient_restatement_matches_reference(golden_dir):
    """The fp64 autograd restatement the GPU tests use for the SupCon gradients reproduces the reference's own
    values and `dmu` on the golden cases (the numpy oracle has closed-form gradients for snn_loss only)."""
    import torch
    from tests.helpers import supcon_torch
    n = 0
    for name, g, sim, tau, ln, ps in _cases(golden_dir):
        if ln == "snn_loss":
            continue
        mu = torch.tensor(g[f"{name}/mu"]).double().requires_grad_(True)
        val = supcon_torch(mu, torch.tensor(g[f"{name}/label"]), str(sim), tau, str(ln), ps)
        assert _close(float(val), float(g[f"{name}/loss"]), rel=1e-5, ab=5e-6), name
        assert _close(float(val), lo.contrastive(g[f"{name}/mu"], g[f"{name}/logvar"], g[f"{name}/label"], sim, tau, ln, ps), rel=1e-9, ab=1e-9)
        val.backward()
        want = g[f"{name}/dmu"]
        assert np.abs(mu.grad.numpy() - want).max() <= 1e-5 * np.abs(want).max() + 1e-8, name
        n += 1
    assert n == 4


def test_row_drop_semantics(golden_dir):
    g = np.load(os.path.join(golden_dir, "contrastive.npz"))
    # singleton label: +inf row excluded; all rows excluded -> nan
    rows = lo.row_losses(g["kat_cosine_ps0/mu"], g["kat_cosine_ps0/logvar"], g["kat_cosine_ps0/label"], "cosine", 0.1)
    want = [9.3442, 8.2578, 13.4202, 13.3455, 2.6642, 0.8273, 0.7269, np.inf]
    assert np.isinf(rows[7]) and np.allclose(rows[:7], want[:7], atol=1e-4)
    assert np.isnan(lo.contrastive(g["cos_allsame_ps1/mu"], g["cos_allsame_ps1/logvar"], g["cos_allsame_ps1/label"],
                                   "cosine", 0.1, ps=True))
    assert np.isnan(float(g["cos_allsame_ps1/loss"]))
    assert np.isnan(float(g["cos_alldiff_ps0/loss"]))
    with pytest.raises(ValueError):
        lo.contrastive(np.zeros((4, 2)), np.zeros((4, 2)), np.arange(4), "nope", 0.1)


def test_pair_mask_bits():
    lab = np.array([3, 1, 3, 2, 1])
    cand, pos = lo.positive_sets(lab, ps=False)
    assert cand.sum() == 20 and pos.tolist() == [[0, 0, 1, 0, 0], [0, 0, 0, 0, 1], [1, 0, 0, 0, 0], [0] * 5, [0, 1, 0, 0, 0]]
    cand, pos = lo.positive_sets(lab, ps=True)
    assert (pos + lo.positive_sets(lab, ps=False)[1] == cand).all()
    # ps follows python truthiness (None == False)
    assert (lo.pair_mask(lab, lab, None) == lo.pair_mask(lab, lab, False)).all()


def test_sharded_rows_equal_global(golden_dir):
    g = np.load(os.path.join(golden_dir, "contrastive.npz"))
    name = "cos_B128_D8_t0.1_psFalse"
    mu, lv, lab = g[f"{name}/mu"], g[f"{name}/logvar"], g[f"{name}/label"]
    full = lo.row_losses(mu, lv, lab, "cosine", 0.1)
    parts = [lo.row_losses(mu[r * 32:(r + 1) * 32], lv[r * 32:(r + 1) * 32], lab[r * 32:(r + 1) * 32], "cosine", 0.1,
                           mu_cols=mu, logvar_cols=lv, label_cols=lab, row_offset=r * 32) for r in range(4)]
    assert np.allclose(np.concatenate(parts), full, rtol=0, atol=0)
    s = sum(lo.contrastive_partial(mu[r * 32:(r + 1) * 32], lv[r * 32:(r + 1) * 32], lab[r * 32:(r + 1) * 32], "cosine",
                                   0.1, mu_cols=mu, logvar_cols=lv, label_cols=lab, row_offset=r * 32)[0] for r in range(4))
    c = sum(lo.contrastive_partial(mu[r * 32:(r + 1) * 32], lv[r * 32:(r + 1) * 32], lab[r * 32:(r + 1) * 32], "cosine",
                                   0.1, mu_cols=mu, logvar_cols=lv, label_cols=lab, row_offset=r * 32)[1] for r in range(4))
    assert abs(s / c - float(g[f"{name}/loss64"])) < 1e-12


def test_heads_match_reference(golden_dir):
    h = np.load(os.path.join(golden_dir, "heads.npz"))
    assert _close(lo.recon_sse(h["elbo/xhat"], h["elbo/x"]), float(h["elbo/recon"]))
    assert _close(lo.gaussian_kl(h["elbo/mu_c"], h["elbo/logvar_c"]), float(h["elbo/kl_c"]))
    assert _close(lo.gaussian_kl(h["elbo/mu_s"], h["elbo/logvar_s"]), float(h["elbo/kl_s"]))
    z = lo.reparam(h["elbo/mu_c"].astype(np.float64), h["elbo/logvar_c"].astype(np.float64), h["sample/eps"])
    assert np.abs(z - h["sample/z"]).max() < 1e-6
    slopes = [lo.logistic_anneal(t, 0, 1, 1 / 8) for t in range(12)]
    assert np.allclose(slopes, h["annealer/slopes"], rtol=1e-12)
    assert np.array_equal(lo.roll_style_half(h["tc/z"]), h["tc/shuffled"])


def test_l1out_closed_form_is_the_executed_computation():
    rng = np.random.default_rng(0)
    B, D = 12, 5
    mu, lv, y = rng.normal(size=(B, D)), np.tanh(rng.normal(size=(B, D))), rng.normal(size=(B, D))
    assert abs(lo.l1out_bound_as_executed(mu, lv, y) - lo.l1out_bound_bruteforce(mu, lv, y)) < 1e-12


def test_fp64_chunked_restatement_is_pinned():
    """tests/helpers.py::snn_fp64_chunked (the oracle of the 65536-latent GPU test) == the numpy closed forms, values and
    gradients, both mask polarities, with near-singleton labels (rows without positives are dropped)."""
    import torch
    from tests.helpers import snn_fp64_chunked
    g = torch.Generator().manual_seed(1)
    B, D = 700, 8
    mu = torch.randn(B, D, generator=g)
    lab = torch.randint(0, 300, (B,), generator=g)
    for ps in (False, True):
        val, grad = snn_fp64_chunked(mu, lab, 0.1, ps, chunk=256)
        want = lo.contrastive(mu.numpy(), np.zeros((B, D)), lab.numpy(), "cosine", 0.1, ps=ps)
        wg = lo.snn_grad(mu.numpy(), lab.numpy(), "cosine", 0.1, ps)
        assert abs(val - want) <= 1e-12 * abs(want)
        assert np.abs(grad.numpy() - wg).max() <= 1e-12 * np.abs(wg).max()


def test_group_evidence_oracle_matches_reference(golden_dir):
    """ML-VAE / GVAE evidence accumulation + its closed-form gradient + the group-wise reparameterisation (row f-4) against the
    reference's autograd values recorded by tests/golden/make_golden_baselines.py."""
    import torch
    g = np.load(os.path.join(golden_dir, "baselines.npz"))
    mu, lv, lab = g["ge/mu"], g["ge/lv"], g["ge/label"]
    for mode in ("MLVAE", "GVAE"):
        mg, lg, groups, gid = lo.group_evidence(mu, lv, lab, mode)
        assert np.array_equal(groups, g[f"ge/{mode}/keys"])
        assert np.allclose(mg, g[f"ge/{mode}/mu_g"], rtol=1e-5, atol=1e-6) and np.allclose(lg, g[f"ge/{mode}/lv_g"], rtol=1e-5, atol=1e-6)
        # noise as the reference draws it: torch.randn(n, D) per group (sorted labels) on the seeded CPU generator
        torch.manual_seed(77)
        idx = [np.nonzero(gid == k)[0] for k in range(len(groups))]
        e = torch.cat([torch.randn(len(i), mu.shape[1]) for i in idx]).numpy()
        eps = np.zeros_like(e)
        eps[np.concatenate(idx)] = e
        z = lo.group_reparam(mg, lg, gid, eps)
        assert np.allclose(z, g[f"ge/{mode}/z"], rtol=1e-5, atol=1e-5)
        assert np.array_equal(np.concatenate(idx), g[f"ge/{mode}/indices"])
        # gradient of f = sum(mu_g w1) + sum(lv_g w2) + sum(z w3)
        w1, w2, w3 = g["ge/w1"], g["ge/w2"], g["ge/w3"]
        dmg = w1 + np.stack([w3[gid == k].sum(0) for k in range(len(groups))])
        dlg = w2 + np.stack([(w3 * eps)[gid == k].sum(0) for k in range(len(groups))]) * 0.5 * np.exp(0.5 * lg)
        dmu, dlv = lo.group_evidence_grad(mu, lv, lab, mode, dmg, dlg)
        assert np.allclose(dmu, g[f"ge/{mode}/dmu"], rtol=1e-4, atol=1e-6), mode
        assert np.allclose(dlv, g[f"ge/{mode}/dlv"], rtol=1e-4, atol=1e-6), mode
