"""CPU: the engine emulator without rounding == autograd of the fp32 oracle == the reference goldens.
This pins the hand-derived backward the CUDA engine implements."""
import numpy as np
import pytest
import torch

from oracle import model_oracle as mo
from tests.helpers import digest, emulated_step, load_step, seeded_model, state_of


@pytest.mark.parametrize("tag", ["clear_vae28_ps", "clear_vae28_nops", "clear_vae64_ps"])
def test_manual_backward_equals_reference_autograd(tag):
    g, meta, hyper = load_step(tag)
    model = seeded_model(meta)
    for k, v in model.state_dict().items():  # weights are the reference's
        assert np.allclose(digest(v), g[f"init_digest/{k}"], rtol=1e-12, atol=1e-12), k
    slope = float(g["slope"])
    res = emulated_step(state_of(model), meta, hyper, g, round_bf16=False, slope=slope)
    rel = lambda a, b: abs(float(a) - float(b)) / (abs(float(b)) + 1e-12)
    assert rel(res["recon"], g["recon"]) < 1e-5 and rel(res["kl_c"], g["kl_c"]) < 1e-5 and rel(res["kl_s"], g["kl_s"]) < 1e-5
    assert rel(res["c"], g["c_loss"]) < 2e-5 and rel(res["s"], g["s_loss"]) < 2e-4
    if g["xhat"].shape == tuple(res["xhat"].shape):
        assert np.abs(res["xhat"].numpy() - g["xhat"]).max() < 2e-5
    worst = 0.0
    for k, p in model.named_parameters():
        ref = g[f"grad_digest/{k}"]
        if k not in res["grads"]:
            # conv / linear biases feeding a train-mode BatchNorm: true gradient is zero; the
            # reference's autograd returns rounding noise (SURVEY.md §7 "Zero-gradient biases")
            assert k.endswith(".bias") and np.sqrt(ref[2]) < 1e-3, (k, ref)
            continue
        got = digest(res["grads"][k])
        scale = np.sqrt(ref[2]) + 1e-12  # L2 norm of the reference gradient
        err = abs(np.sqrt(got[2]) - np.sqrt(ref[2])) / scale
        err = max(err, abs(got[0] - ref[0]) / (ref[1] + 1e-12))
        worst = max(worst, err)
        assert err < 2e-3, (k, err, got, ref)  # fp32 noise through 10 layers at alpha=100, tau=0.1
    assert worst > 0  # something was compared


def test_oracle_step_matches_reference_training_step():
    """StepOracle (autograd + own Adam) reproduces the reference's one-batch `_train` on identical draws."""
    for tag in ["clear_vae28_ps", "tc_vae28", "mim_club_vae28", "mim_l1out_vae28"]:
        g, meta, hyper = load_step(tag)
        model = seeded_model(meta)
        st = state_of(model)
        aux = None
        if meta["kind"] != "clear":
            aux = {k[len("aux_init/"):]: torch.tensor(g[k]) for k in g.files if k.startswith("aux_init/")}
        h = dict(hyper)
        so = mo.StepOracle(meta["kind"], st, meta["arch"], meta["cin"], h, hyper["lr"], aux=aux, aux_lr=hyper.get("aux_lr"),
                           estimator=meta["est"] or "CLUBSample")
        X, label = torch.tensor(g["X"]), torch.tensor(g["label"])
        n_eps = sum(1 for k in g.files if k.startswith("eps/"))
        eps = [torch.tensor(g[f"eps/{i}"]) for i in range(n_eps)]
        extra = [(eps[2 * i], eps[2 * i + 1]) for i in range(1, n_eps // 2)]
        logs = so.step(X, label, eps=(eps[0], eps[1]), extra_eps=extra or None, perm=torch.tensor(g["perm"]))
        rel = lambda a, b: abs(float(a) - float(b)) / (abs(float(b)) + 1e-12)
        assert rel(logs["recon"], g["recon"]) < 1e-5 and rel(logs["c_loss"], g["c_loss"]) < 2e-5, tag
        assert rel(logs["loss"], g["loss"]) < 1e-5, (tag, logs["loss"], float(g["loss"]))
        if "mi_loss" in logs:
            assert abs(logs["mi_loss"] - float(g["mi_loss"])) < 1e-5 * abs(float(g["mi_loss"])) + 2e-6, tag
        if meta["kind"] == "tc":
            assert rel(logs["factor_loss"], g["train_logs1"][0]) < 1e-5
        if meta["kind"] == "mim":
            assert np.allclose(logs["mi_learning"], g["train_logs2"], rtol=2e-5, atol=1e-6), (tag, logs["mi_learning"], g["train_logs2"])
        # gradients of the VAE loss (per tensor, relative to that tensor's norm)
        for k, gr in so.last_grads.items():
            ref = g[f"grad_digest/{k}"]
            if gr is None:
                continue
            got = digest(gr)
            if np.sqrt(ref[2]) < 1e-3:
                continue  # zero-gradient biases: noise in both
            assert abs(np.sqrt(got[2]) - np.sqrt(ref[2])) / np.sqrt(ref[2]) < 2e-3, (tag, k)
