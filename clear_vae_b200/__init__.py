"""clear_vae_b200 — B200-native (sm_100a) implementation of the CLEAR-VAE
training-step hot path behind the reference's Python API (`code/src`).

Sub-modules mirror the reference layout:
  losses            <- code/src/losses.py
  models.vae        <- code/src/models/vae.py
  models.mi_estimator <- code/src/models/mi_estimator.py
  trainer           <- code/src/trainer.py
  utils.trainer_utils <- code/src/utils/trainer_utils.py
"""
__version__ = "0.1.0"
