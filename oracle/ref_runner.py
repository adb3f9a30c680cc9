"""Drive the UNMODIFIED reference (staged by oracle/stage_reference.py into oracle/_ref/src) as a same-box comparator.

Test / measurement infrastructure only: `bench.py --impl reference` (host CPU), `bench.py`'s `eager_gpu_baseline` leg
(the same code on cuda:0 through cuDNN / cuBLAS / ATen) and tests import this; the product never does.

What is timed is the reference's own loop, not a restatement: `get_*_trainer(...)` (code/src/utils/trainer_utils.py:87-201)
followed by `trainer._train(dataloader, verbose=False, epoch_id=1, ...)` (code/src/trainer.py:435-493, 629-709, 820-897) over a
`DataLoader(TensorDataset(X, label, style), batch_size=B)` of synthetic batches.  One `_train` call over W
batches is the warm-up, a second over K batches is the timed region (wall clock around the call, device synchronised on
both sides — the reference synchronises every step anyway through its `float(loss)` logging).
"""
from __future__ import annotations

import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF, "src", "trainer.py"))


def modules():
    """(trainer_utils, trainer, losses, vae, mi_estimator) modules of the staged reference."""
    if not available():
        raise RuntimeError("reference not staged: run `python oracle/stage_reference.py` where /root/reference exists")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import src.losses as losses
    import src.models.mi_estimator as mi
    import src.models.vae as vae
    import src.trainer as trainer
    import src.utils.trainer_utils as tu
    return tu, trainer, losses, vae, mi


def build_trainer(cfg, device, seed=101):
    """cfg: a bench.py CONFIGS entry.  Same seed / construction order as the GPU arm => same initial weights."""
    import torch
    tu = modules()[0]
    hp = cfg["hp"]
    torch.manual_seed(seed)
    if cfg["kind"] == "clear":
        return tu.get_clearvae_trainer(hp["beta"], hp["ps"], hp["lr"], cfg["z"], hp["alpha"], hp["temperature"], device, cfg["arch"], cfg["cin"])
    if cfg["kind"] == "tc":
        return tu.get_cleartcvae_trainer(hp["beta"], hp["la"], hp["lr"], hp["aux_lr"], cfg["z"], hp["alpha"], hp["temperature"], device,
                                         cfg["arch"], cfg["cin"])
    return tu.get_clearmimvae_trainer(hp["beta"], cfg["est"], hp["la"], hp["lr"], hp["aux_lr"], cfg["z"], hp["alpha"], hp["temperature"],
                                      device, cfg["arch"], cfg["cin"])


def loader(cfg, pool, n, pin=False):
    """DataLoader over n batches drawn round-robin from `pool` (list of (X, label) CPU tensors)."""
    import torch
    from torch.utils.data import DataLoader, TensorDataset
    X = torch.cat([pool[i % len(pool)][0] for i in range(n)])
    y = torch.cat([pool[i % len(pool)][1] for i in range(n)])
    style = torch.zeros_like(y)   # third column carried by the reference datasets, unused by the loops
    return DataLoader(TensorDataset(X, y, style), batch_size=cfg["B"], shuffle=False, pin_memory=pin)


def extra_args(cfg):
    return {"clear": [], "tc": [[]], "mim": [[], []]}[cfg["kind"]]


def time_train(cfg, device, pool, steps, warmup, seed=101):
    """samples/s of the unmodified reference `_train` on `device` ('cpu' or 'cuda:0'); returns (sps, ms_per_step, logs)."""
    import torch
    dev = torch.device(device)
    tr = build_trainer(cfg, dev, seed)
    cuda = dev.type == "cuda"
    sync = (lambda: torch.cuda.synchronize(dev)) if cuda else (lambda: None)
    if warmup > 0:
        tr._train(loader(cfg, pool, warmup, cuda), False, 1, *extra_args(cfg))
    sync()
    dl = loader(cfg, pool, steps, cuda)
    extra = extra_args(cfg)
    t0 = time.perf_counter()
    tr._train(dl, False, 1, *extra)
    sync()
    dt = (time.perf_counter() - t0) / steps
    return cfg["B"] / dt, dt * 1e3, extra
