"""Variational MI upper bounds used by CLEAR-MIM-VAE, with the reference's class names,
constructor signatures, sub-module names (state_dict keys) and methods
(`code/src/models/mi_estimator.py:108-198`).

`L1OutUB.forward` reproduces the value the reference *executes* (its `[B,B] + [B,B,1]`
broadcast makes a [B,B,B] tensor, mi_estimator.py:181-189) through the closed form
    mean_c ap_cc - mean_{b,c} ap_bc - log1p(e^-20 / (B-1))
evaluated from the column moments sum_c y_c and sum_c y_c^2, i.e. O(B*D) work and memory
instead of three 4 GiB temporaries at B = 1024.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

__all__ = ["CLUBSample", "L1OutUB"]


class _GaussianHeads(nn.Module):
    def __init__(self, x_dim, y_dim, hidden_size):
        super().__init__()
        self.p_mu = nn.Sequential(nn.Linear(x_dim, hidden_size // 2), nn.ReLU(), nn.Linear(hidden_size // 2, y_dim))
        self.p_logvar = nn.Sequential(nn.Linear(x_dim, hidden_size // 2), nn.ReLU(), nn.Linear(hidden_size // 2, y_dim),
                                      nn.Tanh())

    def get_mu_logvar(self, x_samples):
        return self.p_mu(x_samples), self.p_logvar(x_samples)

    def loglikeli(self, x_samples, y_samples):
        mu, logvar = self.get_mu_logvar(x_samples)
        return (-((mu - y_samples) ** 2) / logvar.exp() - logvar).sum(dim=1).mean(dim=0)

    def learning_loss(self, x_samples, y_samples):
        return -self.loglikeli(x_samples, y_samples)


class CLUBSample(_GaussianHeads):
    """Sampled CLUB bound (mi_estimator.py:108-146).  The permutation is drawn with
    `torch.randperm` on the CPU generator exactly like the reference (one draw per call)."""

    def forward(self, x_samples, y_samples, random_index=None):
        mu, logvar = self.get_mu_logvar(x_samples)
        if random_index is None:
            random_index = torch.randperm(x_samples.shape[0]).long()
        random_index = random_index.to(y_samples.device, non_blocking=True)
        inv = (-logvar).exp()
        positive = -((mu - y_samples) ** 2) * inv
        negative = -((mu - y_samples[random_index]) ** 2) * inv
        return (positive.sum(dim=-1) - negative.sum(dim=-1)).mean() / 2.0


class L1OutUB(_GaussianHeads):
    """Leave-one-out bound as the reference executes it (mi_estimator.py:149-198)."""

    def forward(self, x_samples, y_samples):
        B = y_samples.shape[0]
        mu, logvar = self.get_mu_logvar(x_samples)
        inv = (-logvar).exp()
        positive = (-0.5 * (mu - y_samples) ** 2 * inv - 0.5 * logvar).sum(dim=-1)
        # mean over (b, c) of ap[b, c] = sum_d -(y_cd - mu_bd)^2 / (2 var_bd) - logvar_bd / 2
        s1 = y_samples.sum(dim=0, keepdim=True)           # sum_c y_cd
        s2 = (y_samples * y_samples).sum(dim=0, keepdim=True)
        sq = s2 - 2.0 * mu * s1 + B * mu * mu              # sum_c (y_cd - mu_bd)^2
        all_mean = ((-0.5 * sq * inv).sum(dim=-1) / B - 0.5 * logvar.sum(dim=-1)).mean()
        return positive.mean() - all_mean - math.log1p(math.exp(-20.0) / (B - 1.0))
