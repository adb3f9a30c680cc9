"""GPU: comparison baselines of SURVEY.md §8 row f-4 — ML-VAE / GVAE group evidence (segmented-reduction kernels), the
HierarchicalVAETrainer step and the CNN / LAM classifiers on the conv trunk — against goldens recorded from the unmodified
reference (tests/golden/make_golden_baselines.py) and the fp64 oracle (oracle/latent_oracle.py::group_evidence*)."""
import os

import numpy as np
import pytest
import torch

from oracle import latent_oracle as lo
from tests.helpers import sample_of

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF16 = 1e-2


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "baselines.npz"))


def rel(a, b):
    return abs(float(a) - float(b)) / (abs(float(b)) + 1e-12)


@pytest.mark.parametrize("mode", ["MLVAE", "GVAE"])
def test_group_evidence_and_reparam_match_reference(gold, mode):
    from clear_vae_b200.group import accumulate_group_evidence, groupwise_reparam_each
    t = lambda k: torch.tensor(gold[k]).to(DEV)
    mu, lv = t("ge/mu").requires_grad_(True), t("ge/lv").requires_grad_(True)
    mg, lg, gd = accumulate_group_evidence(mu, lv, t("ge/label"), mode)
    assert list(gd.keys()) == gold[f"ge/{mode}/keys"].tolist()                      # sorted unique labels, same dict order
    assert torch.allclose(mg.cpu(), torch.tensor(gold[f"ge/{mode}/mu_g"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(lg.cpu(), torch.tensor(gold[f"ge/{mode}/lv_g"]), rtol=1e-5, atol=1e-6)
    torch.manual_seed(77)                                                            # same CPU-generator state as the reference run
    z, idx, sizes = groupwise_reparam_each(mg, lg, gd)
    assert torch.equal(idx.cpu(), torch.tensor(gold[f"ge/{mode}/indices"])) and torch.equal(sizes.cpu(), torch.tensor(gold[f"ge/{mode}/sizes"]))
    assert torch.allclose(z.cpu(), torch.tensor(gold[f"ge/{mode}/z"]), rtol=1e-5, atol=1e-5)
    f = (mg * t("ge/w1")).sum() + (lg * t("ge/w2")).sum() + (z * t("ge/w3")).sum()
    f.backward()
    assert rel(f, gold[f"ge/{mode}/f"]) < 1e-5
    for got, want in ((mu.grad, gold[f"ge/{mode}/dmu"]), (lv.grad, gold[f"ge/{mode}/dlv"])):
        assert float((got.cpu() - torch.tensor(want)).abs().max()) <= 1e-4 * float(np.abs(want).max())


@pytest.mark.parametrize("mode", ["MLVAE", "GVAE"])
@pytest.mark.parametrize("B,D,ncls", [(1024, 8, 10), (777, 32, 300), (128, 32, 1)])
def test_group_evidence_against_fp64_oracle(mode, B, D, ncls):
    """benchmark-sized and ragged cases: many singleton groups (ncls = 300), one group holding the whole batch (ncls = 1)"""
    from clear_vae_b200.group import _Evidence, MODES
    g = torch.Generator().manual_seed(B + D + ncls)
    mu, lv = torch.randn(B, D, generator=g), torch.randn(B, D, generator=g) * 0.7
    lab = torch.randint(0, ncls, (B,), generator=g) * 3 - 5                          # non-contiguous, negative label values
    dmg_full, dlg_full = torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
    mg_o, lg_o, groups, gid_o = lo.group_evidence(mu.numpy(), lv.numpy(), lab.numpy(), mode)
    G = len(groups)
    a, b = mu.to(DEV).requires_grad_(True), lv.to(DEV).requires_grad_(True)
    _, gid = lab.to(DEV).unique(sorted=True, return_inverse=True)
    assert np.array_equal(gid.cpu().numpy(), gid_o)
    mg, lg, cnt = _Evidence.apply(MODES[mode], a, b, gid.contiguous(), G)
    assert np.array_equal(cnt.cpu().numpy(), np.bincount(gid_o).astype(np.float32))
    assert np.abs(mg.detach().cpu().numpy() - mg_o).max() <= 1e-5 * max(1.0, np.abs(mg_o).max())
    assert np.abs(lg.detach().cpu().numpy() - lg_o).max() <= 1e-5 * max(1.0, np.abs(lg_o).max())
    dmg, dlg = dmg_full[:G].contiguous(), dlg_full[:G].contiguous()
    torch.autograd.backward([mg, lg], [dmg.to(DEV), dlg.to(DEV)])
    dmu_o, dlv_o = lo.group_evidence_grad(mu.numpy(), lv.numpy(), lab.numpy(), mode, dmg.numpy().astype(np.float64), dlg.numpy().astype(np.float64))
    assert np.abs(a.grad.cpu().numpy() - dmu_o).max() <= 1e-4 * np.abs(dmu_o).max()
    assert np.abs(b.grad.cpu().numpy() - dlv_o).max() <= 1e-4 * np.abs(dlv_o).max()


@pytest.mark.parametrize("mode", ["MLVAE", "GVAE"])
def test_hierarchical_vae_step_matches_reference(gold, mode):
    """one HierarchicalVAETrainer iteration (trainer.py:338-362) on the reference's batch, weights (seed 5), style noise and
    CPU-generator state: logged losses within the bf16 envelope, group evidence of the content head, gradient direction"""
    from clear_vae_b200.utils.trainer_utils import get_hierarchical_vae_trainer
    torch.manual_seed(5)
    tr = get_hierarchical_vae_trainer(1 / 8, 5e-4, 16, mode, DEV, "VAE", 3)
    tr.model.train()
    X, y = torch.tensor(gold["hv/X"]).to(DEV), torch.tensor(gold["hv/label"]).to(DEV)
    eps_s = torch.tensor(gold["hv/eps_s"]).to(DEV)
    before = {k: v.detach().clone() for k, v in tr.model.named_parameters()}
    o_randn = torch.randn_like
    torch.randn_like = lambda t, *a, **k: eps_s.clone()
    try:
        torch.manual_seed(99)
        rec, kl_c, kl_s = tr.train_step(X, y)
    finally:
        torch.randn_like = o_randn
    torch.cuda.synchronize()
    assert rel(rec, gold[f"hv/{mode}/recon"]) < BF16 and rel(kl_c, gold[f"hv/{mode}/kl_c"]) < BF16 and rel(kl_s, gold[f"hv/{mode}/kl_s"]) < BF16
    agree = total = 0
    for k, p in tr.model.named_parameters():
        if p.dim() < 2 or p.grad is None:
            continue
        ref = gold[f"hv/{mode}/grad_sample/{k}"]
        got = sample_of(p.grad, 256)
        cos = float(np.dot(got, ref) / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30))
        assert cos > 0.97, (k, cos)
        moved = sample_of(p - before[k], 256)
        big = np.abs(ref) > 0.05 * np.abs(ref).max()
        agree += int((np.sign(moved[big]) == -np.sign(ref[big])).sum())       # Adam's first step moves against the gradient sign
        total += int(big.sum())
    assert total > 100 and agree / total > 0.97


def test_cnn_and_lam_baselines_match_reference(gold):
    from clear_vae_b200.losses import lam_loss
    from clear_vae_b200.utils.trainer_utils import get_cnn_trainer, get_lamcnn_trainer
    X, y = torch.tensor(gold["hv/X"]).to(DEV), torch.tensor(gold["cnn/label"]).to(DEV)
    torch.manual_seed(6)
    tr = get_cnn_trainer(10, DEV, "SimpleCNNClassifier", 3)
    tr.model.train()
    logits = tr.model(X)
    loss = tr.criterion(logits, y)
    loss.backward()
    assert float((logits.detach().cpu() - torch.tensor(gold["cnn/logits"])).abs().max()) < 3e-2   # bf16 trunk, logits O(1)
    assert rel(loss, gold["cnn/loss"]) < BF16
    for k, p in tr.model.named_parameters():
        if p.dim() < 2 or p.grad is None:
            continue
        ref, got = gold[f"cnn/grad_sample/{k}"], sample_of(p.grad, 256)
        if np.linalg.norm(ref) == 0:      # the strided sample can land on a dead-ReLU feature column: zero on both sides
            assert np.linalg.norm(got) == 0, k
            continue
        assert float(np.dot(got, ref) / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30)) > 0.97, k
    # state_dict keys equal the reference's module layout
    assert {"net.0.weight", "net.1.running_mean", "net.7.bias", "cls_head.0.weight", "cls_head.1.running_var", "cls_head.3.bias"} <= set(tr.model.state_dict())
    # LAM: CE + lam_coef * lam_loss on the trunk features of (x, same-class partner)
    torch.manual_seed(6)
    tl = get_lamcnn_trainer(10, DEV, 0.5, "LAMCNNClassifier", 3)
    tl.model.train()
    ce, lam = tl.train_step(X, y, torch.tensor(gold["lam/X_tilde"]).to(DEV))
    assert rel(ce, gold["lam/ce"]) < BF16 and rel(lam, gold["lam/lam"]) < 5 * BF16      # squared differences of bf16 features
    for k, p in tl.model.named_parameters():
        if p.dim() < 2 or p.grad is None:
            continue
        ref, got = gold[f"lam/grad_sample/{k}"], sample_of(p.grad, 256)
        if np.linalg.norm(ref) == 0:
            continue
        assert float(np.dot(got, ref) / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30)) > 0.95, k
    # ss_pairing keeps every sample inside its label stratum
    Xt = tl.ss_pairing(X, y)
    for c in y.unique():
        a = X[y == c].flatten(1).sum(1).sort().values
        b = Xt[y == c].flatten(1).sum(1).sort().values
        assert torch.allclose(a, b)
