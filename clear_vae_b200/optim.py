"""One-launch optimiser step for the `torch.optim.Adam` objects the reference factories build
(`code/src/utils/trainer_utils.py:100,139-140,178-181`; `optimizer.step()` at
`code/src/trainer.py:483, 698-699, 870, 886`).

The optimiser object stays a regular `torch.optim.Adam` (same `state_dict()`, same
`exp_avg` / `exp_avg_sq` / `step` entries); only the arithmetic of `step()` is replaced by
`clearvae::adam_step`, which updates every parameter of a group in one grid and keeps the step
counters on the device so the call can sit inside a captured CUDA graph.  Anything the fused
kernel does not cover (weight decay, amsgrad, maximize, tensor learning rates, CPU parameters)
goes through the stock `optimizer.step()`.
"""
from __future__ import annotations

import weakref

import torch

from . import _ops

_cache: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _plain_adam(opt, group) -> bool:
    return (type(opt) is torch.optim.Adam and not group.get("amsgrad", False) and not group.get("maximize", False)
            and group.get("weight_decay", 0) == 0 and isinstance(group["lr"], float)
            and not group.get("differentiable", False))


def _build(opt, gi, group, params):
    dev = params[0].device
    steps = torch.zeros(len(params), dtype=torch.float32, device=dev)
    old = [float(opt.state[p]["step"]) if len(opt.state[p]) else 0.0 for p in params]
    if len(set(old)) != 1:
        return None  # parameters with different histories: leave this group to torch (state untouched)
    for p in params:
        st = opt.state[p]
        if len(st) == 0:
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
    steps.fill_(old[0])
    for i, p in enumerate(params):
        opt.state[p]["step"] = steps[i]  # 0-dim views of one buffer: the kernel advances all of them together
    return dict(key=tuple(id(p) for p in params), steps=steps, counter=torch.zeros(1, dtype=torch.int32, device=dev))


@torch.no_grad()
def fused_adam_step(opt, grad_scale: float = 1.0) -> None:
    """`opt.step()` for a default-configured Adam over CUDA fp32 parameters, as one kernel launch per group."""
    groups = opt.param_groups
    if not all(_plain_adam(opt, g) for g in groups):
        if grad_scale != 1.0:
            for g in groups:
                torch._foreach_mul_([p.grad for p in g["params"] if p.grad is not None], grad_scale)
        opt.step()
        return
    per_opt = _cache.setdefault(opt, {})
    for gi, group in enumerate(groups):
        params = [p for p in group["params"] if p.grad is not None]
        if not params:
            continue
        if any((not p.is_cuda) or p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous()
               for p in params):
            raise RuntimeError("clear_vae_b200.fused_adam_step: contiguous CUDA fp32 parameters required (no CPU path)")
        ent = per_opt.get(gi)
        if ent is not None and not torch.cuda.is_current_stream_capturing():
            # `load_state_dict` replaces the per-parameter `step` tensors: an entry whose counters are no longer the ones the
            # optimiser holds is stale (bias correction would keep using the pre-load count) -> rebuild from the loaded state
            lo, hi = ent["steps"].data_ptr(), ent["steps"].data_ptr() + 4 * ent["steps"].numel()
            for p in params:
                stp = opt.state[p].get("step") if len(opt.state[p]) else None
                if not (torch.is_tensor(stp) and stp.is_cuda and lo <= stp.data_ptr() < hi):
                    ent = None
                    break
        if ent is None or ent["key"] != tuple(id(p) for p in params):
            ent = _build(opt, gi, group, params)
            if ent is None:
                opt.step()
                return
            per_opt[gi] = ent
        b1, b2 = group["betas"]
        st = [opt.state[p] for p in params]
        _ops.ops().adam_step(params, [p.grad for p in params], [s["exp_avg"] for s in st], [s["exp_avg_sq"] for s in st],
                             ent["steps"], ent["counter"], float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                             float(grad_scale))
