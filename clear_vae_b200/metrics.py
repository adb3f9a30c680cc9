"""Evaluation metric of the reference's `evaluate()` paths (losses.py:10-16): generalised mutual
information gap from sklearn's kNN MI estimator.  CPU-side by design (it is sklearn in the
reference too) and not part of the training-step hot path."""
import torch


def mutual_info_gap(label, latent_c, latent_s):
    from sklearn.feature_selection import mutual_info_classif
    label, latent_c, latent_s = label.cpu(), latent_c.cpu(), latent_s.cpu()
    p = torch.bincount(label) / len(label)
    p = p[p > 0]
    H = float(-(p * torch.log(p)).sum())
    mi_c = mutual_info_classif(latent_c, label, discrete_features=False)
    mi_s = mutual_info_classif(latent_s, label, discrete_features=False)
    return (mi_c.mean() - mi_s.mean()) / H
