"""Whole-step parity at a benchmarked size: the CUDA trainer and the CPU step oracle on the SAME batch, the same injected
reparameterisation noise and the same CLUB-S permutation (reference loop bodies: code/src/trainer.py:446-484, 646-699,
841-888).  Test / measurement infrastructure: used by tests/test_parity_sizes_gpu.py and by bench.py's CPU-baseline leg
(which builds both sides anyway); never imported by the product."""
from __future__ import annotations

import torch

from . import model_oracle as mo


def oracle_for(tr, cfg, dtype=torch.float32):
    """StepOracle holding a CPU copy of the trainer's current weights (VAE + auxiliary network)."""
    hp = cfg["hp"]
    cvt = lambda v: v.detach().clone().cpu().to(dtype) if v.is_floating_point() else v.detach().clone().cpu()
    st = {k: cvt(v) for k, v in tr.model.state_dict().items()}
    hyper = dict(temperature=hp["temperature"], alpha=hp["alpha"], beta=hp["beta"], loc=0, scale=1, ps=hp.get("ps"))
    aux = aux_lr = None
    if cfg["kind"] == "tc":
        aux, aux_lr = {k: cvt(v) for k, v in tr.factor_cls.state_dict().items()}, hp["aux_lr"]
        hyper["lambda"] = hp["la"]
    elif cfg["kind"] == "mim":
        aux, aux_lr = {k: cvt(v) for k, v in tr.mi_estimator.state_dict().items()}, hp["aux_lr"]
        hyper["lambda"] = hp["la"]
    so = mo.StepOracle(cfg["kind"], st, cfg["arch"], cfg["cin"], hyper, hp["lr"], aux=aux, aux_lr=aux_lr,
                       estimator=cfg.get("est", "CLUBSample"))
    so.t = tr.annealer.current_step
    return so


def compare_step(tr, cfg, X, label, seed=7, oracle_dtype=torch.float32, return_oracle=False):
    """Runs ONE training step on both sides (both are updated).  Returns {name: (cuda, oracle, rel_err)} for every logged
    scalar, plus 'latent/<k>' relative-L2 errors of the four latent parameter tensors and 'grad/<param>' relative-L2 errors
    of every VAE parameter gradient."""
    dev = next(tr.model.parameters()).device
    B, D = X.shape[0], tr.model.z_dim
    g = torch.Generator().manual_seed(seed)
    rn = lambda: torch.randn(B, D, generator=g)
    eps = (rn(), rn())
    so = oracle_for(tr, cfg, oracle_dtype)
    Xc, yc = X.detach().cpu().to(oracle_dtype), label.detach().cpu()
    oeps = tuple(t.to(oracle_dtype) for t in eps)
    kw, okw = {}, {}
    if cfg["kind"] == "tc":
        e2 = (rn(), rn())
        kw["eps2"], okw["extra_eps"] = tuple(t.to(dev) for t in e2), [tuple(t.to(oracle_dtype) for t in e2)]
    elif cfg["kind"] == "mim":
        inner = [(rn(), rn()) for _ in range(5)]
        kw["inner_eps"], okw["extra_eps"] = [tuple(t.to(dev) for t in p) for p in inner], [tuple(t.to(oracle_dtype) for t in p) for p in inner]
        if cfg.get("est", "CLUBSample") == "CLUBSample":
            perm = torch.randperm(B, generator=g)
            kw["perm"], okw["perm"] = perm, perm
    with torch.no_grad():
        _, lp_ref, _ = mo.forward({k: v.detach().clone() for k, v in so.st.items()}, Xc, oeps[0], oeps[1], cfg["arch"], cfg["cin"], True)
    logs = so.step(Xc, yc, eps=oeps, **okw)
    tr.model.train()
    with torch.no_grad():   # latent parameters of the same forward (train-mode BatchNorm: running statistics do not enter)
        lp = tr.model.fused_step_forward(X.to(dev), label.to(dev), temperature=cfg["hp"]["temperature"], snn=[0, 0], ps=[False, False],
                                         eps=tuple(t.to(dev) for t in eps))[4]
    graph, tr.use_cuda_graph = tr.use_cuda_graph, False
    try:
        out = tr.train_step(X.to(dev), label.to(dev), eps=tuple(t.to(dev) for t in eps), **kw)
    finally:
        tr.use_cuda_graph = graph
    torch.cuda.synchronize(dev)
    sc = out[1].detach().cpu()
    got = dict(recon=float(out[0]), kl_c=float(sc[0]), kl_s=float(sc[1]), c_loss=float(sc[2]))
    if cfg["kind"] == "clear":
        got["s_loss"] = float(sc[3]) if cfg["hp"].get("ps") else -float(sc[3])
    else:
        got["mi_loss"] = float(out[2])
    res = {k: (v, logs[k], abs(v - logs[k]) / (abs(logs[k]) + 1e-12)) for k, v in got.items()}
    if cfg["kind"] == "tc":
        res["factor_loss"] = (float(out[3]), logs["factor_loss"], abs(float(out[3]) - logs["factor_loss"]) / abs(logs["factor_loss"]))
    elif cfg["kind"] == "mim":
        a, b = out[3].detach().cpu().double(), torch.tensor(logs["mi_learning"], dtype=torch.float64)
        res["mi_learning"] = (a.tolist(), b.tolist(), float(((a - b).abs() / (b.abs() + 1e-12)).max()))
    l2 = lambda a, b: float((a.double().cpu() - b.double()).norm() / (b.double().norm() + 1e-30))
    for k, p in tr.model.named_parameters():
        ref = so.last_grads.get(k)
        if p.grad is not None and ref is not None:
            res[f"grad/{k}"] = (None, None, l2(p.grad, ref))
    for k in ("mu_c", "logvar_c", "mu_s", "logvar_s"):
        res[f"latent/{k}"] = (None, None, l2(lp[k], lp_ref[k]))
    return (res, so) if return_oracle else res
