"""Loader for the native extension.  There is NO Python/CPU fallback: if the
in-tree `_C.so` / `_lib/libclearvae_b200.so` are missing or fail to load, every
op raises.  Build them with `python -m clear_vae_b200.build` (or
`__graft_entry__.build()`).

`ops()` returns a thin proxy over `torch.ops.clearvae` that counts kernel launches per op
(one launch per call for every op except the host-side size queries) and can bracket the
calls of selected ops with CUDA events on the launching stream — this is what `bench.py`
uses for `gpu_launches` and for the live per-kernel roofline timing.
"""
from __future__ import annotations

import os
from collections import defaultdict

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
_EXT = os.path.join(_PKG, "_C.so")
_LIB = os.path.join(_PKG, "_lib", "libclearvae_b200.so")
_loaded = False
_HOST_ONLY = {"latent_workspace_bytes", "latent_bwd_workspace_bytes", "recon_workspace_bytes", "bn_act_workspace_bytes", "mi_workspace_bytes", "tc_workspace_bytes"}


class NativeExtensionMissing(RuntimeError):
    pass


def load():
    """Load the torch custom ops (idempotent)."""
    global _loaded
    if _loaded:
        return torch.ops.clearvae
    if not (os.path.exists(_EXT) and os.path.exists(_LIB)):
        raise NativeExtensionMissing(
            f"clear_vae_b200 native extension not built ({_EXT}); run `python -m clear_vae_b200.build`. "
            "There is no CPU / PyTorch fallback for the hot path.")
    torch.ops.load_library(_EXT)
    _loaded = True
    return torch.ops.clearvae


class _Meter:
    def __init__(self):
        self.counts = defaultdict(int)
        self.timed = set()        # op names to bracket with events
        self.events = defaultdict(list)
        self.meta = defaultdict(list)
        self.bytes = defaultdict(int)   # per timed op: bytes of every tensor operand + result, each counted once per call
        self.enabled = True

    def reset(self):
        self.counts.clear()
        self.events.clear()
        self.meta.clear()
        self.bytes.clear()

    def launches(self):
        return sum(v for k, v in self.counts.items() if k not in _HOST_ONLY)

    def elapsed_ms(self):
        """{op: (n, total_ms)} for the timed ops; call after a synchronize.  A pair whose launch the host delivered late (the
        device idles between the two events: first-use module load, a descheduled autograd thread) is not kernel time: values
        beyond 8x the op's median are replaced by the median."""
        out = {}
        for k, v in self.events.items():
            ts = [a.elapsed_time(b) for a, b in v]
            if len(ts) >= 3:
                med = sorted(ts)[len(ts) // 2]
                ts = [t if t <= 8 * med else med for t in ts]
            out[k] = (len(ts), sum(ts))
        return out


meter = _Meter()


def _tensor_bytes(x) -> int:
    """operand footprint of a call: every tensor argument / result once (what the kernel must at least read or write)."""
    if isinstance(x, torch.Tensor):
        return x.numel() * x.element_size()
    if isinstance(x, (list, tuple)):
        return sum(_tensor_bytes(t) for t in x)
    return 0


class _Proxy:
    def __init__(self, ns):
        self._ns = ns
        self._cache = {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            raw = getattr(self._ns, name)

            def fn(*a, _raw=raw, _name=name, **k):
                if not meter.enabled:
                    return _raw(*a, **k)
                meter.counts[_name] += 1
                if _name in meter.timed:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    out = _raw(*a, **k)
                    e1.record()
                    meter.events[_name].append((e0, e1))
                    meter.bytes[_name] += _tensor_bytes(a) + _tensor_bytes(out)
                    return out
                return _raw(*a, **k)

            self._cache[name] = fn
        return fn


_proxy = None


def ops():
    global _proxy
    if _proxy is None:
        _proxy = _Proxy(load())
    return _proxy


def lib_paths():
    return _LIB, _EXT
