"""Group evidence of the ML-VAE / GVAE baselines with the reference's function names and return values
(`code/src/models/vae.py:159-223`), on the segmented-reduction kernels of `csrc/group_evidence.cu`.

  accumulate_group_evidence(mu_c, logvar_c, label_batch, mode) -> (mu_acc_grp [G, D], logvar_acc_grp [G, D], group_idx)
  groupwise_reparam_each(mu_acc_grp, logvar_acc_grp, g_idx)    -> (z [B, D], indices [B], sizes [B])

`group_idx` maps every sorted unique label to the row indices of its group, like the reference's dict (same keys, same
order, same index tensors); it additionally carries the per-row group id so the kernels need no Python loop.  The noise
of the group-wise reparameterisation is drawn exactly like the reference — `torch.randn(n, D)` on the CPU generator, group
by group in sorted-label order (vae.py:205) — and uploaded once.
"""
from __future__ import annotations

import torch

from . import _ops

MODES = {"MLVAE": 0, "GVAE": 1}


class GroupIndex(dict):
    """label -> row indices (the reference's `group_idx`), plus `gid` (int64 [B], group rank of every row) and `G`."""
    gid: torch.Tensor
    G: int


class _Evidence(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mode, mu, logvar, gid, G):
        mu, logvar = mu.contiguous(), logvar.contiguous()
        mg, lg, cnt = _ops.ops().group_evidence_fwd(mode, mu, logvar, gid, G)
        ctx.mode = mode
        ctx.save_for_backward(mu, logvar, gid, mg, lg, cnt)
        ctx.mark_non_differentiable(cnt)
        return mg, lg, cnt

    @staticmethod
    def backward(ctx, dmg, dlg, _dcnt):
        mu, logvar, gid, mg, lg, cnt = ctx.saved_tensors
        dmg = torch.zeros_like(mg) if dmg is None else dmg.contiguous()
        dlg = torch.zeros_like(lg) if dlg is None else dlg.contiguous()
        dmu, dlv = _ops.ops().group_evidence_bwd(ctx.mode, mu, logvar, gid, mg, lg, cnt, dmg, dlg)
        return None, dmu, dlv, None, None


class _GroupReparam(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mg, lg, eps, gid):
        mg, lg, eps = mg.contiguous(), lg.contiguous(), eps.contiguous()
        z = _ops.ops().group_reparam_fwd(mg, lg, eps, gid)
        ctx.save_for_backward(lg, eps, gid)
        return z

    @staticmethod
    def backward(ctx, dz):
        lg, eps, gid = ctx.saved_tensors
        dmg, dze = _ops.ops().group_reparam_bwd(dz.contiguous(), eps, gid, lg.shape[0])
        return dmg, dze * torch.exp(0.5 * lg) * 0.5, None, None


def accumulate_group_evidence(mu_c, logvar_c, label_batch, mode: str):
    if mode not in MODES:
        raise NotImplementedError("only support using MLVAE or GVAE")   # vae.py:187-188
    if not mu_c.is_cuda:
        raise RuntimeError("clear_vae_b200: group evidence runs on CUDA only (there is no CPU fallback)")
    groups, gid = label_batch.reshape(-1).unique(sorted=True, return_inverse=True)
    G = int(groups.numel())
    mg, lg, _ = _Evidence.apply(MODES[mode], mu_c, logvar_c, gid.contiguous(), G)
    order = torch.argsort(gid, stable=True)                      # rows of group 0 in index order, then group 1, ...
    counts = torch.bincount(gid, minlength=G).tolist()
    idx = GroupIndex()
    o = 0
    for lab, n in zip(groups.tolist(), counts):
        idx[lab] = order[o:o + n]
        o += n
    idx.gid, idx.G = gid.contiguous(), G
    return mg, lg, idx


def groupwise_reparam_each(mu_acc_grp, logvar_acc_grp, g_idx: dict, eps=None):
    device = mu_acc_grp.device
    D = mu_acc_grp.shape[1]
    index_list = list(g_idx.values())
    indices = torch.cat(index_list, dim=0)
    sizes = torch.cat([torch.ones_like(i) * len(i) for i in index_list], dim=0)
    gid = getattr(g_idx, "gid", None)
    if gid is None:   # a plain dict built elsewhere: recover the per-row group id from the index lists
        gid = torch.empty(indices.numel(), dtype=torch.int64, device=device)
        for g, i in enumerate(index_list):
            gid[i] = g
    if eps is None:
        # noise exactly as the reference draws it: one CPU randn per group in dict order (vae.py:205), rows in group order;
        # z is returned in the ORIGINAL row order (vae.py:218-221), so the noise is scattered back to those rows
        e = torch.cat([torch.randn(len(i), D) for i in index_list], dim=0).to(device)
        eps = torch.empty_like(e)
        eps[indices] = e
    z = _GroupReparam.apply(mu_acc_grp, logvar_acc_grp, eps, gid)
    return z, indices, sizes
