"""CNN baselines sharing the conv trunk of the VAEs (reference `code/src/models/cnn.py:7-66`): same class names, constructor
signatures, attribute names (`net`, `cls_head`) and `state_dict()` keys.

`net` is the encoder stack of `VAE` / `VAE64` — Conv2d(k, s=2, p=1) -> BatchNorm2d -> ReLU blocks + Flatten — and runs on the
same sm_100a kernels, forward and backward (`engine.EncoderFn`: tcgen05 implicit GEMMs, fused BatchNorm statistics, hand-derived
backward).  `EncoderFn` ends in a linear map on the flattened features; the trunk uses the identity there, so `net(x)` returns
the [B, 2048] features themselves (exactly: the features are bf16 values and the product with 1.0 is exact in the fp32
accumulator).  The small classification heads (2048 -> 256 -> n_class, or 2048 -> n_class for LAM) are ordinary torch modules.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..engine import EncoderFn, Engine, LayerSpec

__all__ = ["SimpleCNNClassifier", "SimpleCNN64Classifier", "LAMCNNClassifier", "LAMCNN64Classifier"]


class _Trunk(nn.Sequential):
    """Conv/BN/ReLU blocks + Flatten; the sub-modules only hold parameters / buffers (same indices as the reference Sequential)."""

    def __init__(self, in_channel, chans, k, img):
        mods = []
        c = in_channel
        for co in chans:
            mods += [nn.Conv2d(c, co, k, 2, 1), nn.BatchNorm2d(co), nn.ReLU()]
            c = co
        mods.append(nn.Flatten())
        super().__init__(*mods)
        self._geom = (in_channel, tuple(chans), k, img)
        self._engine = None
        self._eye = None

    def _eng(self):
        if self._engine is None:
            cin, chans, k, h = self._geom
            specs, c = [], cin
            for i, co in enumerate(chans):
                sp = LayerSpec(False, k, 2, 1, 0, c, co, h, 3 * i, 3 * i + 1)
                specs.append(sp)
                c, h = co, sp.hout
            self._engine = Engine(specs, [])
        e = self._engine
        e.training = self.training
        e.enc_buffers = [(self[s.bn].running_mean, self[s.bn].running_var) for s in e.enc_specs]
        return e

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("clear_vae_b200: the conv trunk runs on CUDA only (there is no CPU fallback)")
        eng = self._eng()
        params = []
        for s in eng.enc_specs:
            c, b = self[s.conv], self[s.bn]
            params += [c.weight, c.bias, b.weight, b.bias]
        K = eng.enc_specs[-1].cout * eng.enc_specs[-1].hout ** 2
        if self._eye is None or self._eye.device != x.device:
            self._eye = torch.eye(K, device=x.device)
            self._zero = torch.zeros(K, device=x.device)
        h = EncoderFn.apply(eng, x, self._eye, self._zero, *params)
        if self.training:
            torch._foreach_add_([self[s.bn].num_batches_tracked for s in eng.enc_specs], 1)
        return h


class SimpleCNNClassifier(nn.Module):
    def __init__(self, n_class: int = 10, in_channel: int = 1) -> None:
        super().__init__()
        self.net = _Trunk(in_channel, (32, 64, 128), 3, 28)
        self.cls_head = nn.Sequential(nn.Linear(2048, 256), torch.nn.BatchNorm1d(256), torch.nn.ReLU(), torch.nn.Linear(256, n_class))

    def forward(self, x):
        return self.cls_head(self.net(x))


class SimpleCNN64Classifier(SimpleCNNClassifier):
    def __init__(self, n_class: int = 4, in_channel: int = 3) -> None:
        super().__init__(n_class, in_channel)   # like the reference: the 28-px trunk is built first (same RNG consumption), then replaced
        self.net = _Trunk(in_channel, (32, 64, 128, 256, 512), 4, 64)


class LAMCNNClassifier(SimpleCNNClassifier):
    def __init__(self, n_class: int = 10, in_channel: int = 1) -> None:
        super().__init__(n_class, in_channel)
        self.cls_head = nn.Linear(2048, n_class)


class LAMCNN64Classifier(SimpleCNN64Classifier):
    def __init__(self, n_class: int = 4, in_channel: int = 3) -> None:
        super().__init__(n_class, in_channel)
        self.cls_head = nn.Linear(2048, n_class)
