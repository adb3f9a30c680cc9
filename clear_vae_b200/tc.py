"""Fused density-ratio total-correlation term of CLEAR-TC-VAE (`csrc/tc_factor.cu`).

The reference builds `factor_cls = Sequential(Linear(Z,Z), ReLU(), Linear(Z,1), Sigmoid())`
(`code/src/utils/trainer_utils.py:133-138`) and uses it twice per step (`code/src/trainer.py:654-699`):
the TC penalty `relu(log(d/(1-d))).mean()` that back-propagates into the VAE, and the discriminator's own
BCE update on `z` vs `factor_shuffling(z)`.  Each use is one launch here; the `nn.Sequential` only holds the
parameters (same `state_dict()` keys).  A discriminator of any other shape keeps the plain module path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _ops

TC_BOUND, TC_DISC = 0, 1
_ws: dict = {}


def _workspace(device, nbytes, shape=()):
    # one zeroed scratch per (device, stream, problem shape): ticket / partial-sum layouts depend on the shape, and launches on
    # different streams may overlap
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream if device.type == "cuda" else 0, shape)
    ws = _ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)  # word 0 = self-resetting ticket
        _ws[key] = ws
    return ws


def fused_params(factor_cls):
    """(w1, b1, w2, b2) when `factor_cls` is the reference's 2-layer discriminator on CUDA with Z <= 64, else None."""
    if not (isinstance(factor_cls, nn.Sequential) and len(factor_cls) == 4 and isinstance(factor_cls[0], nn.Linear)
            and isinstance(factor_cls[1], nn.ReLU) and isinstance(factor_cls[2], nn.Linear) and isinstance(factor_cls[3], nn.Sigmoid)):
        return None
    l1, l2 = factor_cls[0], factor_cls[2]
    Z = l1.in_features
    if not (l1.out_features == Z and l2.in_features == Z and l2.out_features == 1 and Z % 2 == 0 and Z <= 64 and l1.bias is not None
            and l2.bias is not None and l1.weight.is_cuda and l1.weight.dtype == torch.float32):
        return None
    return l1.weight, l1.bias, l2.weight, l2.bias


def _rows(t):
    return t if (t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]) else t.contiguous()


class _Bound(torch.autograd.Function):
    """mi = relu(log(d / (1 - d))).mean(), d = factor_cls(z); gradient flows to z only — the discriminator's own
    gradients from this term are discarded by the reference trainer (`factor_optimizer.zero_grad()`, trainer.py:682)."""

    @staticmethod
    def forward(ctx, z, w1, b1, w2, b2):
        ops = _ops.ops()
        z = _rows(z)
        ws = _workspace(z.device, ops.tc_workspace_bytes(TC_BOUND, z.shape[0], z.shape[1]), (TC_BOUND,) + tuple(z.shape))
        out, dz = ops.tc_factor(TC_BOUND, z, w1.detach(), b1.detach(), w2.detach().reshape(-1), b2.detach(), ws)
        ctx.save_for_backward(dz)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return _ops.ops().scale_by(g.contiguous(), dz), None, None, None, None


def tc_bound(z, params):
    return _Bound.apply(z, *params)


def disc_grads(z, params):
    """BCE discriminator loss on (z, factor_shuffling(z)) and its parameter gradients in one launch; the gradients are left
    in `.grad` of the four parameters (views of one flat buffer).  Returns the loss (0-dim)."""
    ops = _ops.ops()
    z = _rows(z.detach())
    w1, b1, w2, b2 = params
    ws = _workspace(z.device, ops.tc_workspace_bytes(TC_DISC, z.shape[0], z.shape[1]), (TC_DISC,) + tuple(z.shape))
    out, _ = ops.tc_factor(TC_DISC, z, w1.detach(), b1.detach(), w2.detach().reshape(-1), b2.detach(), ws)
    o = 1
    for p in params:
        n = p.numel()
        p.grad = out[o:o + n].view(p.shape)
        o += n
    return out[0]
