"""Variational MI upper bounds used by CLEAR-MIM-VAE, with the reference's class names,
constructor signatures, sub-module names (state_dict keys) and methods
(`code/src/models/mi_estimator.py:108-198`).

On CUDA every method is ONE launch of the fused estimator kernel (`csrc/mi_estimator.cu`):
both MLP forwards, the bound / likelihood, and the whole backward (to the latents for
`forward`, to the eight parameters for `learning_loss`).  The `nn.Sequential` sub-modules only
hold the parameters (same `state_dict()` keys and construction-time RNG consumption as the
reference); only the accessor `get_mu_logvar` calls them.  There is no CPU path: CPU tensors raise,
layer widths above 32 raise `NotImplementedError`.

`L1OutUB.forward` reproduces the value the reference *executes* (its `[B,B] + [B,B,1]`
broadcast makes a [B,B,B] tensor, mi_estimator.py:181-189) through the closed form
    mean_c ap_cc - mean_{b,c} ap_bc - log1p(e^-20 / (B-1))
evaluated from the column moments sum_c y_c and sum_c y_c^2, i.e. O(B*D) work and memory
instead of three 4 GiB temporaries at B = 1024.

Difference from the reference worth knowing: `forward` (the bound that feeds the VAE loss)
back-propagates into its *inputs* only.  The reference also deposits gradients on the
estimator's parameters there, but both trainers discard them (`mi_estimator_optimizer.zero_grad()`
runs before the estimator's own update, trainer.py:885).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import _ops

__all__ = ["CLUBSample", "L1OutUB"]

MI_LEARN, MI_CLUB, MI_L1OUT = 0, 1, 2
_MAX_WIDTH = 32
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


def _rows(t):
    """[B, D] view with unit column stride: column slices of one latent tensor are passed in place (no copy kernel)."""
    return t if (t.dim() == 2 and (t.shape[1] == 1 or t.stride(1) == 1) and t.stride(0) >= t.shape[1]) else t.contiguous()


def _run(mode, x, y, perm, params):
    ops = _ops.ops()
    B, Dx = x.shape
    H, Dy = params[0].shape[0], y.shape[1]
    ws = _workspace(x.device, ops.mi_workspace_bytes(mode, B, Dx, H, Dy), (mode, B, Dx, H, Dy))
    return ops.mi_estimator(mode, x, y, perm, params, ws)


class _Bound(torch.autograd.Function):
    """value = bound(x, y); gradients flow to x and y only (see module docstring)."""

    @staticmethod
    def forward(ctx, mode, x, y, perm, *params):
        x, y = _rows(x), _rows(y)
        out, dx, dy = _run(mode, x, y, perm, [p.detach() for p in params])
        ctx.mode = mode
        ctx.save_for_backward(dx, dy, y, out)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        dx, dy, y, out = ctx.saved_tensors
        gx, gy = _ops.ops().mi_bound_bwd(ctx.mode, g.contiguous(), dx, dy, y, out)
        return (None, gx, gy, None) + (None,) * 8


class _Learn(torch.autograd.Function):
    """learning_loss(x, y) with gradients to the eight parameters (x, y are detached samples on the hot path)."""

    @staticmethod
    def forward(ctx, x, y, *params):
        out, _, _ = _run(MI_LEARN, _rows(x), _rows(y), None, [p.detach() for p in params])
        ctx.shapes = [p.shape for p in params]
        ctx.save_for_backward(out)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        flat = out[1:] * g
        grads, o = [], 0
        for s in ctx.shapes:
            n = math.prod(s)
            grads.append(flat[o:o + n].view(s))
            o += n
        return (None, None, *grads)


class _GaussianHeads(nn.Module):
    def __init__(self, x_dim, y_dim, hidden_size):
        super().__init__()
        self.p_mu = nn.Sequential(nn.Linear(x_dim, hidden_size // 2), nn.ReLU(), nn.Linear(hidden_size // 2, y_dim))
        self.p_logvar = nn.Sequential(nn.Linear(x_dim, hidden_size // 2), nn.ReLU(), nn.Linear(hidden_size // 2, y_dim),
                                      nn.Tanh())

    # ---- fused path -------------------------------------------------------------------
    def _params8(self):
        return [self.p_mu[0].weight, self.p_mu[0].bias, self.p_mu[2].weight, self.p_mu[2].bias,
                self.p_logvar[0].weight, self.p_logvar[0].bias, self.p_logvar[2].weight, self.p_logvar[2].bias]

    def _check(self, x, y):
        w = self.p_mu[0].weight
        if not (x.is_cuda and y.is_cuda and w.is_cuda):
            raise RuntimeError("clear_vae_b200: the MI estimators run on CUDA only (there is no CPU fallback)")
        if x.dtype != torch.float32 or y.dtype != torch.float32 or x.dim() != 2 or y.dim() != 2:
            raise TypeError("clear_vae_b200: estimator inputs must be float32 [B, D] matrices")
        if max(w.shape[0], w.shape[1], y.shape[1]) > _MAX_WIDTH:
            raise NotImplementedError(f"clear_vae_b200: estimator layer widths above {_MAX_WIDTH} are not built")

    def learning_grads(self, x_samples, y_samples):
        """Hot-path form of `learning_loss(...).backward()`: one launch; returns the loss (0-dim) and leaves the
        gradients in `.grad` of the eight parameters (views of one flat buffer)."""
        self._check(x_samples, y_samples)
        params = self._params8()
        out, _, _ = _run(MI_LEARN, _rows(x_samples.detach()), _rows(y_samples.detach()), None,
                         [p.detach() for p in params])
        o = 1
        for p in params:
            n = p.numel()
            p.grad = out[o:o + n].view(p.shape)
            o += n
        return out[0]

    # ---- reference API ----------------------------------------------------------------
    def get_mu_logvar(self, x_samples):
        return self.p_mu(x_samples), self.p_logvar(x_samples)

    def loglikeli(self, x_samples, y_samples):
        return -self.learning_loss(x_samples, y_samples)

    def learning_loss(self, x_samples, y_samples):
        self._check(x_samples, y_samples)
        return _Learn.apply(x_samples, y_samples, *self._params8())


class CLUBSample(_GaussianHeads):
    """Sampled CLUB bound (mi_estimator.py:108-146).  The permutation is drawn with
    `torch.randperm` on the CPU generator exactly like the reference (one draw per call)."""

    def forward(self, x_samples, y_samples, random_index=None):
        if random_index is None:
            random_index = torch.randperm(x_samples.shape[0]).long()
        self._check(x_samples, y_samples)
        random_index = random_index.to(y_samples.device, non_blocking=True)
        return _Bound.apply(MI_CLUB, x_samples, y_samples, random_index.contiguous(), *self._params8())


class L1OutUB(_GaussianHeads):
    """Leave-one-out bound as the reference executes it (mi_estimator.py:149-198)."""

    def forward(self, x_samples, y_samples):
        self._check(x_samples, y_samples)
        return _Bound.apply(MI_L1OUT, x_samples, y_samples, None, *self._params8())
