"""Factories with the reference's names and argument lists
(`code/src/utils/trainer_utils.py:87-201`): config -> model + Adam optimiser(s) + trainer.
Modules are constructed in the reference's order (VAE first, then the auxiliary network), so a
given `torch.manual_seed` yields the reference's initial weights."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..models.cnn import LAMCNN64Classifier, LAMCNNClassifier, SimpleCNN64Classifier, SimpleCNNClassifier
from ..models.mi_estimator import CLUBSample, L1OutUB
from ..models.vae import VAE, VAE64
from ..trainer import (CLEARVAETrainer, ClearMIMVAETrainer, ClearTCVAETrainer, HierarchicalVAETrainer, LAMCNNTrainer,
                       SimpleCNNTrainer)

_ARCHS = {"VAE": VAE, "VAE64": VAE64}
_CNN_ARCHS = {c.__name__: c for c in (SimpleCNNClassifier, SimpleCNN64Classifier, LAMCNNClassifier, LAMCNN64Classifier)}
_ESTIMATORS = {"CLUBSample": CLUBSample, "L1OutUB": L1OutUB}


def _adam(params, lr, device):
    """`torch.optim.Adam(params, lr=lr)` like the reference factories; on CUDA the graph-capturable variant
    (device-side step counters) so that a whole training step can be replayed as one CUDA graph."""
    if torch.device(device).type == "cuda":
        return torch.optim.Adam(params, lr=lr, capturable=True, foreach=True)
    return torch.optim.Adam(params, lr=lr)


def _arch(name):
    if name not in _ARCHS:
        raise NameError(f"name '{name}' is not defined")  # the reference resolves the string with eval()
    return _ARCHS[name]


def get_clearvae_trainer(beta, ps, vae_lr, z_dim, alpha, temperature, device, vae_arch: str = "VAE", in_channel: int = 1,
                         verbose_period: int = 5):
    vae = _arch(vae_arch)(total_z_dim=z_dim, in_channel=in_channel).to(device)
    optimizer = _adam(vae.parameters(), vae_lr, device)
    return CLEARVAETrainer(vae, optimizer, sim_fn="cosine",
                           hyperparameter={"temperature": temperature, "alpha": alpha, "beta": beta, "ps": ps, "loc": 0,
                                           "scale": 1},
                           verbose_period=verbose_period, device=device)


def get_cleartcvae_trainer(beta, la, vae_lr, factor_cls_lr, z_dim, alpha, temperature, device, vae_arch: str = "VAE",
                           in_channel: int = 1, verbose_period: int = 5):
    vae = _arch(vae_arch)(total_z_dim=z_dim, in_channel=in_channel).to(device)
    factor_cls = nn.Sequential(nn.Linear(z_dim, z_dim), nn.ReLU(), nn.Linear(z_dim, 1), nn.Sigmoid()).to(device)
    vae_optimizer = _adam(vae.parameters(), vae_lr, device)
    factor_optimizer = _adam(factor_cls.parameters(), factor_cls_lr, device)
    return ClearTCVAETrainer(vae, factor_cls, optimizers={"vae_optim": vae_optimizer, "factor_optim": factor_optimizer},
                             sim_fn="cosine",
                             hyperparameter={"temperature": temperature, "alpha": alpha, "beta": beta, "loc": 0, "scale": 1,
                                             "lambda": la},
                             verbose_period=verbose_period, device=device)


def get_clearmimvae_trainer(beta, mi_estimator: str, la, vae_lr, mi_estimator_lr, z_dim, alpha, temperature, device,
                            vae_arch: str = "VAE", in_channel: int = 1, verbose_period: int = 5):
    vae = _arch(vae_arch)(total_z_dim=z_dim, in_channel=in_channel).to(device)
    if mi_estimator not in _ESTIMATORS:
        raise NameError(f"name '{mi_estimator}' is not defined")
    est = _ESTIMATORS[mi_estimator](x_dim=z_dim // 2, y_dim=z_dim // 2, hidden_size=z_dim).to(device)
    vae_optimizer = _adam(vae.parameters(), vae_lr, device)
    est_optimizer = _adam(est.parameters(), mi_estimator_lr, device)
    return ClearMIMVAETrainer(vae, est, optimizers={"vae_optim": vae_optimizer, "mi_estimator_optim": est_optimizer},
                              sim_fn="cosine",
                              hyperparameter={"temperature": temperature, "beta": beta, "loc": 0, "scale": 1, "alpha": alpha,
                                              "lambda": la},
                              verbose_period=verbose_period, device=device)


# ---- comparison baselines (trainer_utils.py:20-84) ----------------------------------------------------------------------
def _cnn_arch(name):
    if name not in _CNN_ARCHS:
        raise NameError(f"name '{name}' is not defined")
    return _CNN_ARCHS[name]


def get_cnn_trainer(n_class, device, cnn_arch: str = "SimpleCNNClassifier", in_channel: int = 1, verbose_period: int = 5):
    cnn = _cnn_arch(cnn_arch)(n_class=n_class, in_channel=in_channel).to(device)
    optimizer = _adam(cnn.parameters(), 1e-4, device)
    return SimpleCNNTrainer(cnn, optimizer, torch.nn.CrossEntropyLoss(), verbose_period=verbose_period, device=device)


def get_lamcnn_trainer(n_class, device, lam_coef, cnn_arch: str = "LAMCNNClassifier", in_channel: int = 1, verbose_period: int = 5):
    cnn = _cnn_arch(cnn_arch)(n_class=n_class, in_channel=in_channel).to(device)
    optimizer = _adam(cnn.parameters(), 1e-4, device)
    return LAMCNNTrainer(cnn, optimizer, torch.nn.CrossEntropyLoss(), {"lam_coef": lam_coef}, verbose_period=verbose_period,
                         device=device)


def get_hierarchical_vae_trainer(beta, vae_lr, z_dim, group_mode, device, vae_arch: str = "VAE", in_channel: int = 1,
                                 verbose_period: int = 5):
    vae = _arch(vae_arch)(total_z_dim=z_dim, in_channel=in_channel, group_mode=group_mode).to(device)
    optimizer = _adam(vae.parameters(), vae_lr, device)
    return HierarchicalVAETrainer(vae, optimizer, hyperparameter={"beta": beta, "scale": 1, "loc": 0}, verbose_period=verbose_period,
                                  device=device)
