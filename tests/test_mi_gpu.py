"""GPU parity: fused MI-estimator kernels (CLUB-S / L1OutUB) and the fused Adam step.

Checked against (i) the goldens produced by the unmodified reference (`tests/golden/heads.npz`,
mi_estimator.py:108-198) and (ii) an fp64 torch restatement of the same formulas at sizes that exercise
ragged row tiles, non-power-of-two widths and the 16/32-wide template instances.
"""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
VAL_REL, VAL_ABS = 1e-5, 2e-6   # north_star: loss components within 1e-5 relative in fp32
GRAD_REL = 1e-4                 # north_star: gradients within 1e-4 (of the tensor's largest entry)

NAMES = ["p_mu.0.weight", "p_mu.0.bias", "p_mu.2.weight", "p_mu.2.bias",
         "p_logvar.0.weight", "p_logvar.0.bias", "p_logvar.2.weight", "p_logvar.2.bias"]


def close(a, b, rel=VAL_REL, ab=VAL_ABS):
    return abs(float(a) - float(b)) <= rel * abs(float(b)) + ab


def grad_close(a, b, rel=GRAD_REL):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max()) <= rel * float(b.abs().max()) + 1e-9


def build(cls_name, x_dim, y_dim, hidden):
    from clear_vae_b200.models import mi_estimator as me
    return getattr(me, cls_name)(x_dim, y_dim, hidden).to(DEV)


def ref_heads(sd, x):
    """fp64 restatement of get_mu_logvar (mi_estimator.py:118-127)."""
    def mlp(pfx, tanh):
        h = torch.relu(x @ sd[f"{pfx}.0.weight"].T + sd[f"{pfx}.0.bias"])
        o = h @ sd[f"{pfx}.2.weight"].T + sd[f"{pfx}.2.bias"]
        return torch.tanh(o) if tanh else o
    return mlp("p_mu", False), mlp("p_logvar", True)


def ref_values(kind, sd, x, y, perm):
    mu, lv = ref_heads(sd, x)
    if kind == "learn":
        return -((-((mu - y) ** 2) / lv.exp() - lv).sum(1).mean(0))
    if kind == "CLUBSample":
        pos = -((mu - y) ** 2) / lv.exp()
        neg = -((mu - y[perm]) ** 2) / lv.exp()
        return (pos.sum(-1) - neg.sum(-1)).mean() / 2.0
    B = y.shape[0]
    ap = (-((y[None, :, :] - mu[:, None, :]) ** 2) / 2.0 / lv.exp()[:, None, :] - lv[:, None, :] / 2.0).sum(-1)  # [b, c]
    return torch.diagonal(ap).mean() - ap.mean() - math.log1p(math.exp(-20.0) / (B - 1.0))


@pytest.mark.parametrize("cls_name", ["CLUBSample", "L1OutUB"])
def test_estimators_match_reference_goldens(golden_dir, cls_name):
    g = np.load(os.path.join(golden_dir, "heads.npz"))
    est = build(cls_name, 8, 8, 16)
    est.load_state_dict({k: torch.tensor(g[f"{cls_name}/state/{k}"]) for k in NAMES})
    x = torch.tensor(g[f"{cls_name}/x"], device=DEV, requires_grad=True)
    y = torch.tensor(g[f"{cls_name}/y"], device=DEV, requires_grad=True)
    if cls_name == "CLUBSample":
        val = est(x, y, torch.tensor(g[f"{cls_name}/perm"]))
    else:
        val = est(x, y)
    val.backward()
    assert close(val, g[f"{cls_name}/forward"]), (float(val), float(g[f"{cls_name}/forward"]))
    assert grad_close(x.grad, torch.tensor(g[f"{cls_name}/dx"]))
    assert grad_close(y.grad, torch.tensor(g[f"{cls_name}/dy"]))
    # learning_loss + its parameter gradients, through autograd and through the hot-path entry
    ll = est.learning_loss(x.detach(), y.detach())
    est.zero_grad()
    ll.backward()
    assert close(ll, g[f"{cls_name}/learning_loss"])
    for k, p in est.named_parameters():
        assert grad_close(p.grad, torch.tensor(g[f"{cls_name}/lgrad/{k}"])), k
    auto = {k: p.grad.clone() for k, p in est.named_parameters()}
    ll2 = est.learning_grads(x.detach(), y.detach())
    assert float(ll2) == float(ll)
    for k, p in est.named_parameters():
        assert torch.equal(p.grad, auto[k]), k   # same kernel, deterministic reduction order


@pytest.mark.parametrize("B,dx,dy,hidden", [(1000, 8, 8, 16), (130, 5, 7, 12), (64, 16, 16, 32), (333, 32, 32, 64),
                                            (2, 8, 8, 16), (4096, 8, 8, 16)])
@pytest.mark.parametrize("cls_name", ["CLUBSample", "L1OutUB"])
def test_estimators_match_fp64_restatement(cls_name, B, dx, dy, hidden):
    torch.manual_seed(B + dx)
    est = build(cls_name, dx, dy, hidden)
    sd = {k: v.detach().double().cpu() for k, v in est.state_dict().items()}
    x = torch.randn(B, dx)
    y = torch.randn(B, dy) * 0.7 + 0.3
    perm = torch.randperm(B)
    xr, yr = x.double().requires_grad_(True), y.double().requires_grad_(True)
    want = ref_values(cls_name, sd, xr, yr, perm)
    want.backward()
    xd, yd = x.to(DEV).requires_grad_(True), y.to(DEV).requires_grad_(True)
    got = est(xd, yd, perm) if cls_name == "CLUBSample" else est(xd, yd)
    (3.0 * got).backward()   # the trainers scale the bound by lambda (trainer.py:861-867)
    assert close(got, want, ab=5e-6), (float(got), float(want))
    assert grad_close(xd.grad, 3.0 * xr.grad)
    assert grad_close(yd.grad, 3.0 * yr.grad)
    # learning loss
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    wl = ref_values("learn", sdg, x.double(), y.double(), None)
    wl.backward()
    ll = est.learning_grads(x.to(DEV), y.to(DEV))
    assert close(ll, wl)
    for k, p in est.named_parameters():
        assert grad_close(p.grad, sdg[k].grad), k
    # bit-reproducible (fixed-order reduction, no float atomics on results)
    ll_b = est.learning_grads(x.to(DEV), y.to(DEV))
    assert float(ll_b) == float(ll)


def test_estimators_reject_cpu_and_wide_layers():
    from clear_vae_b200.models.mi_estimator import CLUBSample
    with pytest.raises(RuntimeError):
        CLUBSample(8, 8, 16)(torch.randn(4, 8), torch.randn(4, 8))
    with pytest.raises(NotImplementedError):
        CLUBSample(8, 8, 130).to(DEV)(torch.randn(4, 8, device=DEV), torch.randn(4, 8, device=DEV))


def test_fused_adam_matches_torch_adam():
    from clear_vae_b200.optim import fused_adam_step
    torch.manual_seed(5)
    shapes = [(7,), (33, 5), (2048, 16), (1,), (128, 64, 3, 3), (4099,)] + [(3, 3)] * 70   # > 64 tensors: two launches
    a = [torch.randn(s, device=DEV).requires_grad_(True) for s in shapes]
    b = [t.detach().clone().requires_grad_(True) for t in a]
    oa = torch.optim.Adam(a, lr=2e-3, capturable=True, foreach=True)
    ob = torch.optim.Adam(b, lr=2e-3)
    for it in range(4):
        gs = [torch.randn(s, device=DEV) * (10.0 ** (it - 2)) for s in shapes]
        for t, u, g in zip(a, b, gs):
            t.grad, u.grad = g.clone(), g.clone()
        if it == 2:           # a parameter without gradient is skipped by both
            a[1].grad = b[1].grad = None
        ver = a[0]._version
        fused_adam_step(oa)
        assert a[0]._version > ver      # packed-weight caches key on the version counter
        ob.step()
        if it == 2:
            break
    for t, u in zip(a, b):
        assert float((t - u).abs().max()) <= 2e-6 * max(1.0, float(u.abs().max()))
    assert float(oa.state[a[0]]["step"]) == 3.0
    sd = oa.state_dict()
    assert len(sd["state"]) == len(shapes)
