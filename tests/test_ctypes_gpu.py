"""GPU: kernel launches through RAW ctypes, exactly as INTEGRATION.md shows a reference maintainer would bind them — the
Python code blocks are extracted from the document and executed, so the documented stubs cannot rot.  No torch custom op,
no `clear_vae_b200` Python on the path: only `libclearvae_b200.so` and device pointers."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import latent_oracle as lo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "clear_vae_b200", "_lib", "libclearvae_b200.so")


@pytest.fixture(scope="module")
def stubs():
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", doc, flags=re.S)
    assert len(blocks) >= 2
    src = "\n".join(blocks[:2]).replace('ctypes.CDLL("libclearvae_b200.so")', f'ctypes.CDLL({LIB!r})')
    ns = {}
    exec(compile(src, "INTEGRATION.md", "exec"), ns)
    return ns


@pytest.mark.parametrize("B,D,ncls", [(257, 8, 10), (1024, 8, 10), (4096, 32, 4)])
def test_snn_forward_stub_matches_oracle(stubs, B, D, ncls):
    g = torch.Generator().manual_seed(B + D)
    mu = torch.randn(B, D, generator=g)
    lab = torch.randint(0, ncls, (B,), generator=g)
    for ps in (False, True):
        val, stats = stubs["snn_forward"](mu.cuda(), lab.cuda(), 0.1, ps)
        torch.cuda.synchronize()
        want = lo.contrastive(mu.numpy(), np.zeros((B, D)), lab.numpy(), "cosine", 0.1, ps=ps)
        assert abs(float(val) - want) <= 1e-5 * abs(want) + 2e-6, (float(val), want)
        assert stats.shape == (B, 2) and bool(torch.isfinite(stats[:, 0]).all())


def test_learning_loss_stub_matches_oracle(stubs):
    from oracle import model_oracle as mo
    B, D = 512, 8
    est = mo.init_estimator_state(D, D, 2 * D, seed=3)

    class _MLP(torch.nn.Module):
        def __init__(self, pre):
            super().__init__()
            self.w0 = torch.nn.Parameter(est[f"{pre}.0.weight"].cuda())
            self.b0 = torch.nn.Parameter(est[f"{pre}.0.bias"].cuda())
            self.w2 = torch.nn.Parameter(est[f"{pre}.2.weight"].cuda())
            self.b2 = torch.nn.Parameter(est[f"{pre}.2.bias"].cuda())

    class _Est:
        p_mu, p_logvar = _MLP("p_mu"), _MLP("p_logvar")

    g = torch.Generator().manual_seed(5)
    x, y = torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
    loss, grads = stubs["learning_loss_and_grads"](_Est, x.cuda().contiguous(), y.cuda().contiguous())
    torch.cuda.synchronize()
    for k in est:
        est[k].requires_grad_(True)
    want = mo.estimator_learning_loss(est, x, y)
    gs = torch.autograd.grad(want, [est[f"{p}.{i}.{t}"] for p in ("p_mu", "p_logvar") for i in (0, 2) for t in ("weight", "bias")])
    flat = torch.cat([t.reshape(-1) for t in gs])
    assert abs(float(loss) - float(want)) <= 1e-5 * abs(float(want))
    assert float((grads.cpu() - flat).abs().max()) <= 1e-4 * float(flat.abs().max())
