"""Loss functions with the reference's names and call signatures
(`code/src/losses.py`), executed by the sm_100a kernels in `csrc/`.

  vae_loss(x_reconstr, x, mu_c, mu_s, logvar_c, logvar_s) -> (recon, kl_c, kl_s)   losses.py:41-50
  contrastive_loss(mu, logvar, label, sim_fn, temperature, loss_name="snn_loss", ps=False)  losses.py:98-126

There is no CPU path: CPU tensors raise (the ops are registered for CUDA only).
"""
from __future__ import annotations

import torch

from . import _ops
from .latent import latent_block, S_KL0, S_KL1, S_LOSS0, _workspace

__all__ = ["vae_loss", "contrastive_loss", "reconstruction_loss", "gaussian_kl", "lam_loss", "accurary", "auc", "mutual_info_gap"]


class _Recon(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xhat, x):
        ops = _ops.ops()
        xhat = xhat.contiguous()
        x = x.contiguous()
        ws = _workspace(xhat.device, ops.recon_workspace_bytes())
        out = ops.recon_fwd(xhat, x, ws)
        ctx.save_for_backward(xhat, x)
        return out

    @staticmethod
    def backward(ctx, g):
        xhat, x = ctx.saved_tensors
        dx = _ops.ops().recon_bwd(xhat, x, g.contiguous())
        return dx, None


def reconstruction_loss(x_reconstr: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """mean_b sum_{chw} (x_reconstr - x)^2  (losses.py:36-47)."""
    if x_reconstr.shape != x.shape:
        raise RuntimeError(f"shape mismatch {tuple(x_reconstr.shape)} vs {tuple(x.shape)}")
    return _Recon.apply(x_reconstr, x)


def gaussian_kl(mu_c, logvar_c, mu_s, logvar_s):
    """(kl_c, kl_s) of losses.py:48-49 in one launch (two when the heads have different row counts: ML-VAE / GVAE pass the
    [G, D] group evidence as the content parameters)."""
    if mu_c.shape != mu_s.shape:
        out = []
        for m, l in ((mu_c, logvar_c), (mu_s, logvar_s)):
            dummy = torch.zeros(m.shape[0], dtype=torch.int64, device=m.device)
            out.append(latent_block([m], [l], [None], dummy, snn=[0], ps=[0], want_z=False)[1][S_KL0])
        return out[0], out[1]
    dummy = torch.zeros(mu_c.shape[0], dtype=torch.int64, device=mu_c.device)
    _, sc = latent_block([mu_c, mu_s], [logvar_c, logvar_s], [None, None], dummy, snn=[0, 0], ps=[0, 0], want_z=False)
    return sc[S_KL0], sc[S_KL1]


def vae_loss(x_reconstr, x, mu_c, mu_s, logvar_c, logvar_s):
    """VAE loss with separating factors — same argument order as the reference
    (note: mu_c, mu_s, logvar_c, logvar_s; called as vae_loss(X_hat, X, **latent_params))."""
    rec = reconstruction_loss(x_reconstr, x)
    kl_c, kl_s = gaussian_kl(mu_c, logvar_c, mu_s, logvar_s)
    return rec, kl_c, kl_s


def contrastive_loss(mu: torch.Tensor, logvar: torch.Tensor, label: torch.Tensor, sim_fn: str, temperature: float,
                     loss_name: str = "snn_loss", ps: bool = False):
    """Temperature-scaled pairwise contrastive loss; mean over rows whose loss is finite
    (rows without a partner under the mask are dropped; `nan` if none is left)."""
    from .latent import SIM_IDS, _LOGVAR_SIMS
    # jeffrey / mahalanobis / modified_l2 read logvar and give it a gradient (losses.py:62-84); cosine / l2 ignore it
    lv = logvar if SIM_IDS.get(sim_fn) in _LOGVAR_SIMS else None
    _, sc = latent_block([mu], [lv], [None], label.reshape(-1).long(), snn=[1], ps=[ps], sim_fn=sim_fn,
                         temperature=temperature, loss_name=loss_name, want_z=False)
    return sc[S_LOSS0]


def lam_loss(feature_x: torch.Tensor, feature_x_tilde: torch.Tensor, y: torch.Tensor, linear_w: torch.nn.Parameter):
    """Labelled LAM penalty of the CNN baseline (losses.py:173-187): squared difference of the class-weighted feature
    contributions of a sample and its same-class partner.  A [B, 2048] elementwise expression on the trunk features; the trunk
    itself (forward and backward) runs on the conv kernels (models/cnn.py)."""
    w_y = linear_w[y]
    return (((feature_x - feature_x_tilde) * w_y) ** 2).sum(dim=1).mean()


def accurary(logit: torch.Tensor, y: torch.Tensor):
    """(sic) losses.py:18-20"""
    yh = logit.argmax(dim=1).cpu()
    return (yh.view(-1) == y.view(-1).cpu()).float().mean()


def auc(logit: torch.Tensor, y: torch.Tensor):
    """per-class average precision / ROC AUC through sklearn, like the reference (losses.py:23-33); CPU-side evaluation glue"""
    from sklearn.metrics import average_precision_score, roc_auc_score
    num_classes = int(y.max() + 1)
    ph = logit.softmax(dim=1).detach().cpu()
    y = y.cpu()
    y_binarized = torch.eye(num_classes)[y]
    aupr_scores, auroc_scores = dict(), dict()
    for i in range(num_classes):
        aupr_scores[i] = round(average_precision_score(y_binarized[:, i], ph[:, i]), 3)
        auroc_scores[i] = round(roc_auc_score(y_binarized[:, i], ph[:, i]), 3)
    return aupr_scores, auroc_scores


def mutual_info_gap(label, latent_c, latent_s):
    from .metrics import mutual_info_gap as _mig
    return _mig(label, latent_c, latent_s)
