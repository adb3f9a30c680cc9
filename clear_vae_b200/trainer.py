"""Trainer classes with the reference's names, constructor signatures, `fit` return values and
hyper-parameter dictionaries (`code/src/trainer.py:22-75, 415-493, 573-709, 781-897`).

The per-batch bodies run the fused sm_100a path:
  encode (tcgen05 conv stack) -> one latent kernel (reparam + KL + SNN terms) -> decode with the
  reconstruction error fused into the last elementwise pass -> hand-written backward -> optimiser.
Scalars are read back once per step and only when a progress bar is shown; the per-step lists
that `fit` returns are filled from device buffers at the end of the epoch (same values, no
per-step synchronisation).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
from torch.optim.optimizer import Optimizer
from torch.utils.data import DataLoader
from tqdm import tqdm

from .latent import S_KL0, S_KL1, S_LOSS0, S_LOSS1, DistSpec, latent_block
from .losses import contrastive_loss, vae_loss  # noqa: F401  (re-exported like the reference module)
from .models.mi_estimator import CLUBSample
from .models.vae import VAE
from .optim import fused_adam_step


class LogisticAnnealer:
    """KL weight beta / (1 + exp(-(step - loc) / scale))  (trainer.py:22-38)."""

    def __init__(self, loc, scale, beta) -> None:
        self.current_step = 0
        self.loc = loc
        self.scale = scale
        self.beta = beta

    def __call__(self, kl_loss) -> torch.Tensor:
        return kl_loss * self.slope()

    def slope(self) -> float:
        exponent = -(self.current_step - self.loc) / self.scale
        return self.beta / (1 + math.exp(exponent))

    def step(self) -> None:
        self.current_step += 1


class Trainer:
    def __init__(self, model: nn.Module, optimizer: Optimizer, verbose_period: int, device: torch.device, transform=None) -> None:
        self.model = model
        self.optimizer = optimizer
        self.verbose_period = verbose_period
        self.device = device
        self.transform = transform

    def fit(self, epochs: int, train_loader: DataLoader, valid_loader: None | DataLoader = None):
        for epoch in range(epochs):
            verbose = (epoch % self.verbose_period) == 0
            self._train(train_loader, verbose, epoch)
            if valid_loader is not None:
                self._valid(valid_loader, verbose, epoch)

    def evaluate(self, **kwarg):
        pass

    def _train(self, **kwarg):
        pass

    def _valid(self, **kwarg):
        pass


class DevicePrefetcher:
    """Iterates `(X, label)` device batches one batch ahead of the consumer: while step i is being enqueued / executed,
    batch i+1 is copied host->device on a copy stream into one of three rotating static device slots per shape (no
    allocator traffic, no `record_stream`), so the PCIe transfer (9.6 MB per 1024 x 3 x 28 x 28 batch, ~0.18 ms) leaves
    the step's critical path.  A slot is reused three batches later; by then the step that read it has normally
    finished, which is checked on the host (event query / wait) so that the copy itself carries no device-side
    dependency and starts at once.  Source tensors should be pinned (`DataLoader(pin_memory=True)`).  Same batches,
    same order as the wrapped iterable; batch i stays valid until the consumer asks for batch i+2."""

    NSLOT = 5   # device input slots: the host may stage up to four batches ahead of the step that is executing

    def __init__(self, batches, device, transform=None):
        self.batches, self.device, self.transform = batches, torch.device(device), transform
        self.enabled = self.device.type == "cuda"
        self.stream = torch.cuda.Stream(device=self.device) if self.enabled else None
        self._slots = {}
        self._consumed = [None] * self.NSLOT

    def _stage(self, batch, i):
        X, label = batch[0], batch[1].reshape(-1).long()
        if not self.enabled:
            X, label = X.to(self.device), label.to(self.device)
            return (self.transform(X) if self.transform else X), label, None
        key = (tuple(X.shape), X.dtype)
        slots = self._slots.get(key)
        if slots is None:
            slots = self._slots[key] = [(torch.empty(X.shape, dtype=X.dtype, device=self.device),
                                         torch.empty(label.shape, dtype=torch.int64, device=self.device))
                                        for _ in range(self.NSLOT)]
            # the caching allocator may hand back blocks that kernels already queued on the consumer's stream still read
            # (e.g. the previous epoch's slots): the first copies into a new slot set are ordered after that stream
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
        k = i % self.NSLOT
        Xd, ld = slots[k]
        if self._consumed[k] is not None:
            self._consumed[k].synchronize()      # the step that read this slot (batch i - NSLOT): normally long finished
        with torch.cuda.stream(self.stream):
            Xd.copy_(X, non_blocking=True)
            ld.copy_(label, non_blocking=True)
            out = Xd
            if self.transform:
                out = self.transform(Xd)
                out.record_stream(torch.cuda.current_stream(self.device))
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return out, ld, ev

    def _hand_over(self, staged):
        X, label, ev = staged
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
        return X, label

    def __iter__(self):
        pending, i = None, 0
        if self.enabled:
            # a new pass over the slots (next epoch): the last two steps of the previous pass never recorded a
            # "consumed" event, so order this pass's copies after everything the consumer has queued so far
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            self._consumed = [None] * self.NSLOT
        for batch in self.batches:
            if self.enabled and i >= 2:   # the consumer has enqueued its step on batch i-2 (the one before `pending`)
                e = torch.cuda.Event(blocking=True)   # the host sleeps instead of spinning while it waits (8 ranks share the host cores)
                e.record(torch.cuda.current_stream(self.device))
                self._consumed[(i - 2) % self.NSLOT] = e
            staged = self._stage(batch, i)
            i += 1
            if pending is not None:
                yield self._hand_over(pending)
            pending = staged
        if pending is not None:
            yield self._hand_over(pending)


class VAETrainer(Trainer):
    def fit(self, epochs: int, train_loader: DataLoader, valid_loader: None | DataLoader = None):
        out = super().fit(epochs, train_loader, valid_loader)
        self._check_collectives()
        return out

    def _check_collectives(self):
        """a peer-memory collective that timed out records the missing rank instead of hanging: surface it (once per fit)"""
        d = getattr(self, "dist", None)
        if d is not None and getattr(d, "peer", None) is not None:
            d.peer.check()

    def prefetch(self, batches):
        """device batches of `batches` with the H2D copy of the next batch overlapped with the current step.  One prefetcher
        (its copy stream and device slots) lives on the trainer across epochs."""
        pf = getattr(self, "_prefetcher", None)
        if pf is None or pf.device != torch.device(self.device) or pf.transform is not self.transform:
            pf = self._prefetcher = DevicePrefetcher(batches, self.device, self.transform)
        pf.batches = batches
        return pf

    def _valid(self, dataloader, verbose, epoch_id):
        if verbose:
            mig, mse = self.evaluate(dataloader, verbose, epoch_id)
            print(f"gMIG: {round(mig, 3)}; mse: {round(float(mse), 3)}")

    # ---- shared pieces of the fused step ---------------------------------------------
    def _batch(self, batch):
        X, label = batch[0], batch[1].reshape(-1).long()
        X, label = X.to(self.device, non_blocking=True), label.to(self.device, non_blocking=True)
        if self.transform:
            X = self.transform(X)
        return X, label

    # A step is split into a host part (annealer / weights, before and after) and a device body that is free of
    # host decisions, so the body can be captured once in a CUDA graph and replayed (no tracing compiler involved).
    # Host-written step inputs (annealer weights, CLUB-S permutation) go through two alternating pinned slots and a
    # stream-ordered copy into a static device tensor, issued OUTSIDE the captured graph.  Nothing synchronises per step
    # any more (the input prefetcher and the asynchronous loss read-back let the host run ahead), so a single pinned
    # buffer read at graph-execution time could already hold a later step's values; a slot is rewritten only after the
    # step that last used it has consumed it (`_slot_done`), which bounds the run-ahead to two steps.
    _step_no = 0
    NSTAGE = 4   # pinned staging slots of the host-written step inputs = how many steps the host may run ahead of the device
                 # (data parallel: a late graph launch on ANY rank stalls every rank at the next collective, so the queue must
                 # be deep enough to absorb host jitter; two slots left one step of slack)

    def _slot(self):
        return self._step_no % self.NSTAGE

    def _begin_step(self):
        self._step_no += 1
        ev = getattr(self, "_slot_done", None)
        if ev is None:
            ev = self._slot_done = [None] * self.NSTAGE
        if ev[self._slot()] is not None:
            ev[self._slot()].synchronize()

    def _end_step(self, device):
        if torch.device(device).type == "cuda":
            e = torch.cuda.Event(blocking=True)   # the host sleeps instead of spinning while it waits (8 ranks share the host cores)
            e.record(torch.cuda.current_stream(device))
            self._slot_done[self._slot()] = e

    def _set_weights(self, slope, alpha_c, alpha_s):
        """grad weights of the packed scalars (kl_c, kl_s, c, s, ...) for autograd.backward, staged in pinned memory."""
        host = getattr(self, "_w_host", None)
        if host is None:
            host = torch.zeros(self.NSTAGE, 8, dtype=torch.float32)
            if torch.cuda.is_available():
                host = host.pin_memory()
            self._w_host = host
        w = host[self._slot()]
        w[S_KL0], w[S_KL1], w[S_LOSS0], w[S_LOSS1] = slope, slope, alpha_c, alpha_s

    def _upload_weights(self, dev):
        wd = getattr(self, "_w_dev", None)
        if wd is None or wd.device != torch.device(dev):
            wd = self._w_dev = torch.zeros(8, dtype=torch.float32, device=dev)
        wd.copy_(self._w_host[self._slot()], non_blocking=True)

    def _weights_dev(self, dev):
        return self._w_dev   # static device tensor, refreshed by `_upload_weights` before the step body / graph replay

    use_cuda_graph = False
    overlap_branches = True   # False: every kernel of the step on one stream (bench.py's per-kernel event timing)
    _graph = None

    # ---- packed bf16 GEMM operands: refreshed in ONE launch right after the optimiser step (csrc/conv_tc.cu:
    # pack_weight_multi_kernel), so the next step's GEMMs find them ready instead of each packing its weight first
    def _engine_of(self):
        m = self.model
        return getattr(m, "_engine", None) or getattr(getattr(m, "net", None), "_engine", None)

    def _refresh_packs(self):
        eng = self._engine_of()
        if eng is not None:
            eng.packs.refresh_all()

    def _weights_signature(self):
        return tuple(p._version for p in self.model.parameters())

    def _needs_perm(self):
        return False

    def train_step(self, X, label, **inject):
        """One iteration of the reference loop body.  With `use_cuda_graph` the device body is captured on first
        use (per input shape) and replayed afterwards; injected noise / permutations force the eager path."""
        self._begin_step()
        self._host_pre()
        self._upload_weights(X.device)
        eng = getattr(self.model, "_engine", None)
        if eng is not None:
            eng.packs.epoch += 1   # packed-weight copies made inside a graph capture are valid for this step only
            # weights written from Python since the last step (load_state_dict, a broadcast, a manual edit): the packed copies
            # the replayed graph would read are stale -> refresh them eagerly (graph replays do not move the version counters)
            if self.use_cuda_graph and getattr(self, "_w_sig", None) is not None and self._w_sig != self._weights_signature():
                self._refresh_packs()
        if self.use_cuda_graph and not inject and X.is_cuda:
            out = self._graph_step(X, label)
        else:
            out = self._device_step(X, label, **inject)
        if eng is not None:
            self._w_sig = self._weights_signature()
        self._end_step(X.device)
        self.annealer.step()
        return out

    def _graph_step(self, X, label):
        g = self._graph
        if g is None or g["X"].shape != X.shape:
            cache = self.__dict__.setdefault("_graphs", {})   # one captured graph per input shape (partial last batches)
            g = cache.get(tuple(X.shape))
            if g is None or self._graph is None:
                g = cache[tuple(X.shape)] = self._capture(X, label)
            self._graph = g
        g["X"].copy_(X, non_blocking=True)
        g["label"].copy_(label, non_blocking=True)
        if g["perm"] is not None:  # CLUB-S: the CPU generator draws the permutation exactly like the reference
            ph = g["perm_host"][self._slot()]
            ph.copy_(self._perm(ph.numel()))
            g["perm"].copy_(ph, non_blocking=True)
        g["graph"].replay()
        # the graph's output buffers are overwritten by the next replay: hand out copies — one launch for all of them
        flat = g.get("flat")
        if flat is None:
            return tuple(t.clone() for t in g["out"])
        c = flat.clone()
        return tuple(c[a:b].view(sh) for (a, b, sh) in g["views"])

    # ---- graph capture must not train: the eager warm-up steps below (allocator / workspace / optimiser-state
    # warm-up that CUDA-graph capture requires) run on the real batch, so everything they touch is put back
    def _stateful(self):
        mods = [m for m in (self.model, getattr(self, "factor_cls", None), getattr(self, "mi_estimator", None)) if m is not None]
        opts = [o for o in (self.optimizer, getattr(self, "factor_optimizer", None), getattr(self, "mi_estimator_optimizer", None))
                if o is not None]
        return mods, opts

    def _snapshot_state(self, device):
        mods, opts = self._stateful()
        tensors = {}
        for m in mods:
            for t in list(m.parameters()) + list(m.buffers()):
                tensors[t] = t.detach().clone()
        known = set()
        for o in opts:
            for st in o.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        tensors[v] = v.detach().clone()
                        known.add(id(v))
        rng = (torch.get_rng_state(), torch.cuda.get_rng_state(device))
        return tensors, known, rng

    def _restore_state(self, snap, device):
        tensors, known, rng = snap
        _, opts = self._stateful()
        with torch.no_grad():
            for t, v in tensors.items():
                t.copy_(v)
            for o in opts:   # optimiser state created by the warm-up itself goes back to "never stepped"
                for st in o.state.values():
                    for v in st.values():
                        if torch.is_tensor(v) and id(v) not in known:
                            v.zero_()
        torch.set_rng_state(rng[0])
        torch.cuda.set_rng_state(rng[1], device)

    def _capture(self, X, label):
        import os, sys
        dbg = (lambda m: print(f"[capture] {m}", file=sys.stderr, flush=True)) if os.environ.get("CLEARVAE_DEBUG") else (lambda m: None)
        sX, sl = X.clone(), label.clone()
        perm = perm_host = None
        kw = {}
        snap = self._snapshot_state(X.device)
        if self._needs_perm():
            nper = X.shape[0] * (self.dist.world if (self.dist is not None and self.dist.world > 1) else 1)   # permutation of the global batch
            perm_host = torch.stack([torch.arange(nper)] * self.NSTAGE).pin_memory()   # one slot per staging step, rewritten before every replay
            perm = perm_host[0].to(X.device)
            kw["perm"] = perm
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        dbg("side-stream warm-up")
        with torch.cuda.stream(side):       # optimiser state / workspaces / caches must exist before capture
            for _ in range(2):
                self._device_step(sX, sl, **kw)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._restore_state(snap, X.device)   # the warm-up steps were real updates: undo them (weights, BN buffers, Adam, RNG)
        self._refresh_packs()                 # ... and re-pack the GEMM operands from the restored weights (one launch, eager)
        torch.cuda.synchronize()
        dbg("begin capture")
        graph = torch.cuda.CUDAGraph()
        from . import _ops
        before = _ops.meter.launches()
        # thread-local error mode: the NCCL watchdog thread may touch CUDA while this thread captures (data parallel)
        eng = getattr(self.model, "_engine", None)
        if eng is not None:
            eng.packs.epoch += 1   # nothing packed during the eager warm-up may be reused by captured kernels
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            out = self._device_step(sX, sl, **kw)
            # the step's (small, fp32) outputs packed into one buffer inside the graph: replays hand out one clone
            flat = views = None
            if all(t.dtype == torch.float32 for t in out):
                flat = torch.cat([t.reshape(-1) for t in out])
                views, o = [], 0
                for t in out:
                    views.append((o, o + t.numel(), tuple(t.shape)))
                    o += t.numel()
        dbg("captured")
        self._graph = dict(graph=graph, X=sX, label=sl, perm=perm, perm_host=perm_host, out=out, flat=flat, views=views,
                           launches=_ops.meter.launches() - before)
        return self._graph

    dist: DistSpec | None = None

    # ---- data-parallel gradient averaging -------------------------------------------------------------------------
    # north_star: "gradients are allreduced overlapped with backward".  The decoder's gradients are complete as soon as the
    # decoder backward returns (autograd accumulates them before it runs the latent block and the encoder), so they travel
    # as a first bucket on a communication stream while the latent + encoder backward still compute; the encoder / heads
    # bucket follows on the same stream.  The optimiser kernel waits for that stream and applies the 1/world.  Buckets are
    # launched from post-accumulate-grad hooks (fork / join nodes inside a captured graph); which parameters receive a
    # gradient is learnt from the first (un-overlapped) step.
    overlap_grad_sync = True

    def _comm_stream(self, device):
        st = getattr(self, "_comm", None)
        if st is None or st.device != torch.device(device):
            st = self._comm = torch.cuda.Stream(device=device)
        return st

    # one-shot peer all-reduce: every rank pulls (world - 1) x bytes — ideal for the latency-bound 1 MB of a 28x28 model, wasteful
    # for VAE64's 12 MB buckets on 4-8 GPUs (7 x 12 MB per rank against NCCL's ring / NVLS 2 x 7/8 x 12 MB): large buckets go to NCCL
    PEER_ALLREDUCE_MAX_BYTES = 4 << 20

    def _allreduce(self, grads):
        d = self.dist
        nbytes = sum(g.numel() for g in grads) * 4
        if d.peer is not None and (nbytes <= self.PEER_ALLREDUCE_MAX_BYTES or d.world <= 2):
            d.peer.allreduce_chunked_(grads)   # pack -> publish -> pull + sum in rank order -> scatter back, <= one slot per call
            return
        import torch.distributed as td
        flat = torch.cat([g.reshape(-1) for g in grads])
        td.all_reduce(flat, group=d.group)
        torch._foreach_copy_(grads, [t.view_as(g) for t, g in zip(flat.split([g.numel() for g in grads]), grads)])

    def _arm_buckets(self, params):
        """after the first synchronous step: bucket 0 = decoder parameters, bucket 1 = everything else (encoder, heads)"""
        names = {id(p): n for n, p in self.model.named_parameters()}
        live = [p for p in params if p.grad is not None]
        b0 = [p for p in live if names.get(id(p), "").startswith("decoder.")]
        b1 = [p for p in live if not names.get(id(p), "").startswith("decoder.")]
        st = self._buckets = dict(params=[b0, b1], left=[len(b0), len(b1)], of={}, fired=[False, False], handles=[])
        for bi, bucket in enumerate(st["params"]):
            for p in bucket:
                st["of"][id(p)] = bi
                st["handles"].append(p.register_post_accumulate_grad_hook(self._grad_ready))

    def _grad_ready(self, p):
        st = getattr(self, "_buckets", None)
        d = self.dist
        if st is None or d is None or d.world == 1 or not getattr(self, "_sync_armed", False):
            return
        bi = st["of"].get(id(p))
        if bi is None:
            return
        st["left"][bi] -= 1
        if st["left"][bi] == 0:
            dev = p.device
            comm, cur = self._comm_stream(dev), torch.cuda.current_stream(dev)
            comm.wait_stream(cur)
            with torch.cuda.stream(comm):
                self._allreduce([q.grad for q in st["params"][bi]])
            st["fired"][bi] = True

    def _begin_grad_sync(self):
        """call before backward: arms the hooks of this step"""
        st = getattr(self, "_buckets", None)
        d = self.dist
        self._sync_armed = bool(st is not None and d is not None and d.world > 1 and self.overlap_grad_sync and d.peer is not None)
        if self._sync_armed:
            st["left"] = [len(b) for b in st["params"]]
            st["fired"] = [False, False]

    def _sync_grads(self, params, model_grads=False):
        """data-parallel gradient averaging; returns the factor the optimiser applies (1/world).  No-op on a single GPU."""
        d = self.dist
        if d is None or d.world == 1:
            return 1.0
        st = getattr(self, "_buckets", None)
        if model_grads and getattr(self, "_sync_armed", False) and st is not None and all(st["fired"]):
            dev = params[0].device
            torch.cuda.current_stream(dev).wait_stream(self._comm_stream(dev))   # both buckets were launched during backward
            self._sync_armed = False
            return 1.0 / d.world
        self._sync_armed = False
        grads = [p.grad for p in params if p.grad is not None]
        self._allreduce(grads)
        if model_grads and st is None and self.overlap_grad_sync and d.peer is not None:
            self._arm_buckets(params)
        return 1.0 / d.world   # the rank average is applied by the optimiser kernel (grad_scale)

    def _perm(self, n):
        """CLUB-S permutation (mi_estimator.py:138): the global CPU generator on one GPU, like the reference; under data
        parallelism a trainer-owned generator seeded identically on every rank, so all ranks draw the same permutation of
        the global batch without a broadcast"""
        d = self.dist
        if d is None or d.world == 1:
            return torch.randperm(n)
        g = getattr(self, "_perm_gen", None)
        if g is None:
            g = self._perm_gen = torch.Generator().manual_seed(20240229)
        return torch.randperm(n, generator=g)


class CLEARVAETrainer(VAETrainer):
    def __init__(self, model: VAE, optimizer: Optimizer, sim_fn: str, hyperparameter: dict[str, float], verbose_period: int,
                 device: torch.device, transform=None) -> None:
        super().__init__(model, optimizer, verbose_period, device, transform)
        self.sim_fn = sim_fn
        self.hyperparameter = hyperparameter
        self.annealer = LogisticAnnealer(loc=hyperparameter["loc"], scale=hyperparameter["scale"], beta=hyperparameter["beta"])

    def _host_pre(self):
        # loss = recon + ann(kl_c) + ann(kl_s) + alpha*c + alpha*s, with s = -s_same when not ps (trainer.py:471-480)
        ps, alpha = self.hyperparameter["ps"], self.hyperparameter["alpha"]
        self._set_weights(self.annealer.slope(), alpha, alpha if ps else -alpha)

    def _device_step(self, X, label, eps=None):
        """Device body of one iteration (trainer.py:446-484); returns device scalars (recon, packed scalars)."""
        vae, hp = self.model, self.hyperparameter
        self.optimizer.zero_grad()
        xhat, recon, z, sc, _ = vae.fused_step_forward(X, label, temperature=hp["temperature"], snn=[1, 1],
                                                        ps=[False, bool(hp["ps"])], sim_fn=self.sim_fn, eps=eps, dist=self.dist)
        self._begin_grad_sync()
        torch.autograd.backward([recon, sc], [torch.ones_like(recon), self._weights_dev(X.device)])
        fused_adam_step(self.optimizer, self._sync_grads(list(vae.parameters()), model_grads=True))
        self._refresh_packs()
        return recon, sc

    def _train(self, dataloader: DataLoader, verbose: bool, epoch_id: int):
        vae = self.model
        vae.train()
        ps = self.hyperparameter["ps"]
        with tqdm(dataloader, unit="batch", mininterval=0, disable=not verbose) as bar:
            bar.set_description(f"Epoch {epoch_id}")
            for X, label in self.prefetch(bar):
                recon, sc = self.train_step(X, label)
                if verbose:
                    v = torch.cat([recon.detach().view(1), sc.detach()[:4]]).tolist()  # one read-back per step
                    bar.set_postfix(recontr_loss=v[0], kl_c=v[1], kl_s=v[2], c_loss=v[3], s_loss=v[4] if ps else -v[4])
        return

    def evaluate(self, dataloader, verbose, epoch_id):
        return _evaluate(self, dataloader, verbose, epoch_id, style_term=True)


def factor_shuffling(z: torch.Tensor, strategy: str = "permute_1"):
    """z = (z_c, z_s); 'permute_1' rolls the style half up by one row (trainer.py:573-587)."""
    z_dim = int(z.shape[1] / 2)
    z_c, z_s = z[:, :z_dim], z[:, z_dim:]
    if strategy == "permute_1":
        return torch.cat([z_c, torch.roll(z_s, -1, 0)], dim=1)
    if strategy == "full":
        raise TypeError("'Tensor' object is not callable")  # the reference's 'full' branch calls a tensor (trainer.py:581)
    raise ValueError("this strategy is not implemented yet")


class ClearTCVAETrainer(VAETrainer):
    def __init__(self, model: VAE, factor_cls: nn.Module, optimizers: dict[str, Optimizer], sim_fn: str,
                 hyperparameter: dict[str, float], verbose_period: int, device: torch.device, transform=None) -> None:
        super().__init__(model, optimizers["vae_optim"], verbose_period, device, transform)
        self.sim_fn = sim_fn
        self.factor_optimizer = optimizers["factor_optim"]
        self.factor_cls = factor_cls
        self.hyperparameter = hyperparameter
        self.annealer = LogisticAnnealer(loc=hyperparameter["loc"], scale=hyperparameter["scale"], beta=hyperparameter["beta"])

    def fit(self, epochs: int, train_loader: DataLoader, valid_loader: None | DataLoader = None):
        factor_d_losses = []
        for epoch in range(epochs):
            verbose = (epoch % self.verbose_period) == 0
            self._train(train_loader, verbose, epoch, factor_d_losses)
            if valid_loader is not None:
                self._valid(valid_loader, verbose, epoch)
        self._check_collectives()
        return factor_d_losses

    def _host_pre(self):
        self._set_weights(self.annealer.slope(), self.hyperparameter["alpha"], 0.0)

    def _device_step(self, X, label, eps=None, eps2=None):
        vae, fc, hp = self.model, self.factor_cls, self.hyperparameter
        from . import tc
        fp = tc.fused_params(fc)   # the reference's 2-layer discriminator -> one launch per use
        if fp is None:
            raise RuntimeError("clear_vae_b200: factor_cls must be the reference's Linear(Z,Z)-ReLU-Linear(Z,1)-Sigmoid on the GPU with "
                               "Z <= 64 (trainer_utils.py:133-138); there is no eager fallback for other discriminators")
        # --- VAE update (trainer.py:654-677)
        self.optimizer.zero_grad()
        xhat, recon, z, sc, _ = vae.fused_step_forward(X, label, temperature=hp["temperature"], snn=[1, 0], ps=[False, False],
                                                        sim_fn=self.sim_fn, eps=eps, dist=self.dist)
        mi = tc.tc_bound(z, fp)   # mean over rows of a per-row quantity: the local mean is exactly this rank's share of the global one
        self._begin_grad_sync()
        torch.autograd.backward([recon, sc, mi], [torch.ones_like(recon), self._weights_dev(X.device), torch.full_like(mi, hp["lambda"])])
        fused_adam_step(self.optimizer, self._sync_grads(list(vae.parameters()), model_grads=True))
        self._refresh_packs()
        # --- density-ratio discriminator update (trainer.py:680-699)
        with torch.no_grad():
            _, _, z2 = vae(X, explicit=True) if eps2 is None else _forward_with_eps(vae, X, eps2)
        d = self.dist
        if d is not None and d.world > 1:
            # `factor_shuffling` rolls the style half across the WHOLE batch (trainer.py:583-585: row i pairs with row i+1, the
            # last row of a shard with the first row of the next rank's): every rank runs the (tiny, deterministic)
            # discriminator step on the gathered global batch — identical gradients on all ranks, no gradient all-reduce
            z2 = d.peer.gather([z2])[0] if (d.peer is not None and d.peer.fits([z2])) else _gather_rows(z2, d)
        factor_loss = tc.disc_grads(z2, fp)
        fused_adam_step(self.factor_optimizer)
        return recon, sc, mi.detach(), factor_loss.detach()

    def _train(self, dataloader: DataLoader, verbose: bool, epoch_id: int, factor_d_losses: list):
        self.model.train()
        self.factor_cls.train()
        pending = []
        with tqdm(dataloader, unit="batch", mininterval=0, disable=not verbose) as bar:
            bar.set_description(f"Epoch {epoch_id}")
            for X, label in self.prefetch(bar):
                recon, sc, mi, fl = self.train_step(X, label)
                pending.append(fl)
                if verbose:
                    v = torch.cat([fl.view(1), recon.detach().view(1), sc.detach()[:3], mi.view(1)]).tolist()
                    bar.set_postfix(factor_cls_loss=v[0], recontr_loss=v[1], kl_c=v[2], kl_s=v[3], c_loss=v[4], mi_loss=v[5])
        if pending:
            factor_d_losses.extend(torch.stack(pending).tolist())

    def evaluate(self, dataloader, verbose, epoch_id):
        return _evaluate(self, dataloader, verbose, epoch_id, style_term=False)


class ClearMIMVAETrainer(VAETrainer):
    def __init__(self, model: VAE, mi_estimator: nn.Module, optimizers: dict[str, Optimizer], sim_fn: str,
                 hyperparameter: dict[str, float], verbose_period: int, device: torch.device, transform=None) -> None:
        super().__init__(model, optimizers["vae_optim"], verbose_period, device, transform)
        self.sim_fn = sim_fn
        self.mi_estimator_optimizer = optimizers["mi_estimator_optim"]
        self.mi_estimator = mi_estimator
        self.hyperparameter = hyperparameter
        self.annealer = LogisticAnnealer(loc=hyperparameter["loc"], scale=hyperparameter["scale"], beta=hyperparameter["beta"])

    def fit(self, epochs: int, train_loader: DataLoader, valid_loader: None | DataLoader = None):
        mi_losses, mi_learning_losses = [], []
        for epoch in range(epochs):
            verbose = (epoch % self.verbose_period) == 0
            self._train(train_loader, verbose, epoch, mi_losses, mi_learning_losses)
            if valid_loader is not None:
                self._valid(valid_loader, verbose, epoch)
        self._check_collectives()
        return mi_losses, mi_learning_losses

    def _host_pre(self):
        self._set_weights(self.annealer.slope(), self.hyperparameter["alpha"], 0.0)

    def _needs_perm(self):
        return isinstance(self.mi_estimator, CLUBSample)

    def _side_stream(self, device):
        st = getattr(self, "_side", None)
        if st is None or st.device != torch.device(device):
            st = torch.cuda.Stream(device=device)
            self._side = st
        return st

    def _device_step(self, X, label, eps=None, inner_eps=None, perm=None):
        vae, est, hp = self.model, self.mi_estimator, self.hyperparameter
        D = vae.z_dim
        # --- VAE update (trainer.py:848-871)
        self.optimizer.zero_grad()
        d = self.dist
        dp = d is not None and d.world > 1
        out = vae.fused_step_forward(X, label, temperature=hp["temperature"], snn=[1, 0], ps=[False, False],
                                     sim_fn=self.sim_fn, eps=eps, dist=self.dist, gather_z=dp)
        xhat, recon, z, sc = out[0], out[1], out[2], out[3]
        # Data parallel: CLUB-S pairs row i with row perm(i) of the WHOLE batch and L1OutUB's `all_probs` spans all y
        # (mi_estimator.py:138-143, 170-191), so the bound is evaluated on the gathered latents by every rank (a few CTAs,
        # deterministic => identical on all ranks).  Each rank back-propagates into its own rows only; with the bound weighted
        # by `world` the later rank-average of the parameter gradients equals the single-process global-batch gradient.
        zb = out[5] if dp else z
        zc, zs = zb[:, :D], zb[:, D:]
        if isinstance(est, CLUBSample):
            mi = est(zc, zs, perm if perm is not None else self._perm(zb.shape[0]))
        else:
            mi = est(zc, zs)
        lam = hp["lambda"] * (d.world if dp else 1)
        self._begin_grad_sync()
        torch.autograd.backward([recon, sc, mi], [torch.ones_like(recon), self._weights_dev(X.device), torch.full_like(mi, lam)])
        fused_adam_step(self.optimizer, self._sync_grads(list(vae.parameters()), model_grads=True))
        self._refresh_packs()
        # --- estimator updates: 5 fresh forwards on detached latents (trainer.py:874-888)
        # The encoder is unchanged across the 5 iterations, so its output is computed once and its BatchNorm
        # running statistics receive 5 momentum updates; each iteration still draws fresh noise (c then s) and
        # runs the decoder for its running-statistic side effects, as the reference's full forwards do.
        learn = []
        with torch.no_grad():
            mu_c, lv_c, mu_s, lv_s = vae.encode(X, bn_repeat=5)
            # noise in the reference's order (c then s, iteration by iteration); the five reparameterisations are one launch
            e = [t for j in range(5) for t in ((torch.randn_like(lv_c), torch.randn_like(lv_s)) if inner_eps is None else inner_eps[j])]
            if X.is_cuda:
                from . import _ops
                zs = list(_ops.ops().reparam_multi([mu_c, mu_s], [lv_c, lv_s], [t.contiguous() for t in e]))
            else:   # CPU tensors: the latent op raises (no CPU path), exactly as before
                dummy = torch.zeros(X.shape[0], dtype=torch.int64, device=X.device)
                zs = [latent_block([mu_c, mu_s], [lv_c, lv_s], e[2 * j:2 * j + 2], dummy, snn=[0, 0], ps=[0, 0])[0] for j in range(5)]
        # The five estimator updates (two ~10 us launches each, a handful of CTAs) depend only on the latents; the five
        # decoder passes only feed BatchNorm running statistics.  They run as two parallel branches — a side stream in
        # eager mode, a fork/join inside the captured graph — so the small estimator kernels fill SMs the decoder leaves idle.
        main = torch.cuda.current_stream(X.device)
        side = self._side_stream(X.device) if self.overlap_branches else main   # serial mode: per-kernel timing passes
        side.wait_stream(main)
        with torch.cuda.stream(side):
            z_est = zs
            if d is not None and d.world > 1:
                # Data parallel: the five detached latent batches are gathered ONCE and every rank runs the (tiny)
                # estimator updates on the global batch.  The fixed-order reduction of the estimator kernel makes the
                # gradients bit-identical on all ranks, so the parameters stay in sync without five gradient all-reduces.
                # The exchange sits on the estimator branch (the decoder passes do not wait for it); the join below
                # orders it before the next step's collectives on every rank.
                if d.peer is not None and d.peer.fits(zs):
                    z_est = d.peer.gather(zs)                                        # five pieces, one kernel, final layout
                else:
                    import torch.distributed as td
                    loc = torch.stack(zs)                                            # [5, B, 2D]
                    allz = torch.empty((d.world * 5,) + tuple(loc.shape[1:]), dtype=loc.dtype, device=loc.device)
                    td.all_gather_into_tensor(allz, loc, group=d.group)               # rank-major concatenation along dim 0
                    allz = allz.view((d.world,) + tuple(loc.shape)).permute(1, 0, 2, 3).reshape(5, d.world * loc.shape[1], loc.shape[2])
                    z_est = [allz[j] for j in range(5)]
            for j in range(5):
                ll = est.learning_grads(z_est[j][:, :D], z_est[j][:, D:])   # loss + all parameter gradients: one launch
                fused_adam_step(self.mi_estimator_optimizer)
                learn.append(ll)
            learn_t = torch.stack(learn)
        with torch.no_grad():
            vae.decode_stats_many(zs)
        main.wait_stream(side)
        return recon, sc, mi.detach(), learn_t

    def _train(self, dataloader: DataLoader, verbose: bool, epoch_id: int, mi_losses: list, mi_learning_losses: list):
        self.model.train()
        self.mi_estimator.train()
        p_mi, p_learn = [], []
        with tqdm(dataloader, unit="batch", mininterval=0, disable=not verbose) as bar:
            bar.set_description(f"Epoch {epoch_id}")
            for X, label in self.prefetch(bar):
                recon, sc, mi, learn = self.train_step(X, label)
                p_mi.append(mi)
                p_learn.append(learn)
                if verbose:
                    v = torch.cat([recon.detach().view(1), sc.detach()[:3], mi.view(1)]).tolist()
                    bar.set_postfix(recontr_loss=v[0], kl_c=v[1], kl_s=v[2], c_loss=v[3], mi_loss=v[4])
        if p_mi:
            mi_losses.extend(torch.stack(p_mi).tolist())
            mi_learning_losses.extend(torch.cat(p_learn).tolist())

    def evaluate(self, dataloader, verbose, epoch_id):
        return _evaluate(self, dataloader, verbose, epoch_id, style_term=False)


def _gather_rows(t, dist):
    import torch.distributed as td
    out = torch.empty((dist.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    td.all_gather_into_tensor(out, t.contiguous(), group=dist.group)
    return out


def _forward_with_eps(vae, X, eps):
    """`vae(X, explicit=True)` with injected reparameterisation noise (tests / parity runs)."""
    xhat, recon, z, sc, lp = vae.fused_step_forward(X, torch.zeros(X.shape[0], dtype=torch.int64, device=X.device),
                                                    temperature=1.0, snn=[0, 0], ps=[False, False], eps=eps)
    return xhat, lp, z


def _evaluate(tr, dataloader, verbose, epoch_id, style_term):
    """Eval-mode pass: running-stat BatchNorm, losses accumulated on device, gMIG through the
    reference's sklearn estimator (trainer.py:495-570, 711-778, 899-965; losses.py:10-16)."""
    vae = tr.model
    vae.eval()
    hp = tr.hyperparameter
    tot = None
    labels, lat_c, lat_s = [], [], []
    n = 0
    with torch.no_grad():
        for batch in tqdm(dataloader, disable=not verbose, desc=f"val-epoch {epoch_id}"):
            X, label = tr._batch(batch)
            ps = bool(hp.get("ps", False)) if style_term else False
            xhat, recon, z, sc, _ = vae.fused_step_forward(X, label, temperature=hp["temperature"],
                                                            snn=[1, 1 if style_term else 0], ps=[False, ps], sim_fn=tr.sim_fn)
            vals = torch.cat([recon.view(1), sc[:4]])
            if style_term and not ps:
                vals[4] = -vals[4]
            tot = vals if tot is None else tot + vals
            labels.append(label)
            lat_c.append(z[:, :vae.z_dim])
            lat_s.append(z[:, vae.z_dim:])
            n += 1
    from .metrics import mutual_info_gap
    mig = mutual_info_gap(torch.cat(labels), torch.cat(lat_c), torch.cat(lat_s))
    t = (tot / n).tolist()
    if verbose:
        print("val_recontr_loss={:.3f}, val_kl_c={:.3f}, val_kl_s={:.3f}, val_c_loss={:.3f}".format(*t[:4])
              + (", val_s_loss={:.3f}".format(t[4]) if style_term else ""))
    return mig, float(t[0])


# ======================================================================================================================
# Comparison baselines of the reference's experiment scripts: supervised CNN / LAM classifiers on the
# same conv trunk, the downstream MLP probe on frozen content latents, and the ML-VAE / GVAE group-evidence VAEs.  Same class
# names, constructor signatures and loop semantics as `code/src/trainer.py:92-412`; the conv stacks (forward and backward) run
# on the sm_100a kernels, the optimiser step is the fused Adam kernel.
# ======================================================================================================================
def _refresh_model_packs(model):
    eng = getattr(model, "_engine", None) or getattr(getattr(model, "net", None), "_engine", None)
    if eng is not None:
        eng.packs.refresh_all()


def _classifier_eval(trainer, forward, dataloader, verbose, epoch_id):
    from .losses import accurary, auc
    all_y, all_logits = [], []
    with torch.no_grad():
        for batch in tqdm(dataloader, disable=not verbose, desc=f"val-epoch {epoch_id}"):
            X_batch, y_batch = batch[0], batch[1].reshape(-1)
            all_logits.append(forward(X_batch.to(trainer.device)))
            all_y.append(y_batch)
    all_y, all_logits = torch.cat(all_y), torch.cat(all_logits)
    return auc(all_logits, all_y), accurary(all_logits, all_y)


class _ClassifierValid:
    def _valid(self, dataloader: DataLoader, verbose: bool, epoch_id: int):
        if verbose:
            import numpy as np
            (aupr_scores, auroc_scores), acc = self.evaluate(dataloader, verbose, epoch_id)
            print("val_aupr:", aupr_scores)
            print(np.mean(list(aupr_scores.values())).round(3))
            print("val_auroc:", auroc_scores)
            print(np.mean(list(auroc_scores.values())).round(3))
            print("val_acc:", acc.numpy().round(3))


class DownstreamMLPTrainer(_ClassifierValid, Trainer):
    """MLP probe on the (frozen) content means `vae.encode(x)[0]` (trainer.py:92-165)."""

    def __init__(self, vae: nn.Module, model: nn.Module, optimizer: Optimizer, criterion: nn.Module, verbose_period: int,
                 device: torch.device, transform=None) -> None:
        super().__init__(model, optimizer, verbose_period, device, transform)
        self.criterion = criterion
        self.vae = vae

    def _train(self, dataloader: DataLoader, verbose: bool, epoch_id: int):
        self.model.train()
        with tqdm(dataloader, unit="batch", disable=not verbose) as bar:
            bar.set_description(f"epoch {epoch_id}")
            for batch in bar:
                X_batch, y_batch = batch[0].to(self.device), batch[1].reshape(-1).long().to(self.device)
                if self.transform:
                    X_batch = self.transform(X_batch)
                self.optimizer.zero_grad()
                mu_c = self.vae.encode(X_batch)[0]
                loss = self.criterion(self.model(mu_c), y_batch)
                loss.backward()
                self.optimizer.step()
                bar.set_postfix(loss=float(loss))

    def evaluate(self, dataloader: DataLoader, verbose: bool, epoch_id: int):
        self.model.eval()
        return _classifier_eval(self, lambda X: self.model(self.vae.encode(X)[0]), dataloader, verbose, epoch_id)


class SimpleCNNTrainer(_ClassifierValid, Trainer):
    """Supervised CNN baseline (trainer.py:168-232): cross-entropy on `cnn(x)`."""

    def __init__(self, model: nn.Module, optimizer: Optimizer, criterion: nn.Module, verbose_period: int, device: torch.device,
                 transform=None) -> None:
        super().__init__(model, optimizer, verbose_period, device, transform)
        self.criterion = criterion

    def train_step(self, X_batch, y_batch):
        self.optimizer.zero_grad()
        loss = self.criterion(self.model(X_batch), y_batch)
        loss.backward()
        fused_adam_step(self.optimizer)
        _refresh_model_packs(self.model)
        return loss.detach()

    def _train(self, dataloader: DataLoader, verbose: bool, epoch_id: int):
        self.model.train()
        with tqdm(dataloader, unit="batch", disable=not verbose) as bar:
            bar.set_description(f"epoch {epoch_id}")
            for batch in bar:
                X_batch, y_batch = batch[0].to(self.device), batch[1].reshape(-1).long().to(self.device)
                if self.transform:
                    X_batch = self.transform(X_batch)
                loss = self.train_step(X_batch, y_batch)
                if verbose:
                    bar.set_postfix(loss=float(loss))

    def evaluate(self, dataloader: DataLoader, verbose: bool, epoch_id: int):
        self.model.eval()
        return _classifier_eval(self, self.model, dataloader, verbose, epoch_id)


class LAMCNNTrainer(SimpleCNNTrainer):
    """CNN + labelled LAM penalty (trainer.py:235-288): every sample is paired with a same-class sample of the batch."""

    def __init__(self, model: nn.Module, optimizer: Optimizer, criterion: nn.Module, hyperparameter: dict[str, float],
                 verbose_period: int, device: torch.device, transform=None):
        super().__init__(model, optimizer, criterion, verbose_period, device, transform)
        self.hyperparameter = hyperparameter

    def ss_pairing(self, x, y):
        """stratified shuffle: within every label stratum the samples are permuted (CPU generator, one `randperm` per
        stratum in sorted-label order, like trainer.py:250-258)"""
        new_x = x.clone()
        for c in torch.unique(y):
            _idx = (y == c).nonzero(as_tuple=True)[0]
            _perm = torch.randperm(_idx.shape[0])
            new_x[_idx] = x[_idx[_perm.to(_idx.device)]]
        return new_x

    def train_step(self, X_batch, y_batch, X_tilde_batch=None):
        from .losses import lam_loss
        cnn = self.model
        if X_tilde_batch is None:
            X_tilde_batch = self.ss_pairing(X_batch, y_batch)
        self.optimizer.zero_grad()
        logits = cnn(X_batch)
        loss_ce = self.criterion(logits, y_batch)
        loss_lam = lam_loss(cnn.net(X_batch), cnn.net(X_tilde_batch), y_batch, cnn.cls_head.weight)
        loss = loss_ce + self.hyperparameter["lam_coef"] * loss_lam
        loss.backward()
        fused_adam_step(self.optimizer)
        _refresh_model_packs(self.model)
        return loss_ce.detach(), loss_lam.detach()

    def _train(self, dataloader: DataLoader, verbose: bool, epoch_id: int):
        self.model.train()
        with tqdm(dataloader, unit="batch", disable=not verbose) as bar:
            bar.set_description(f"epoch {epoch_id}")
            for batch in bar:
                X_batch, y_batch = batch[0].to(self.device), batch[1].reshape(-1).long().to(self.device)
                if self.transform:
                    X_batch = self.transform(X_batch)
                loss_ce, loss_lam = self.train_step(X_batch, y_batch)
                if verbose:
                    bar.set_postfix(ce_loss=float(loss_ce), lam_loss=float(loss_lam))


class HierarchicalVAETrainer(VAETrainer):
    """ML-VAE / GVAE (trainer.py:291-412): the content parameters of a batch are pooled per label group
    (`accumulate_group_evidence`), the content code is sampled group-wise, reconstruction and style KL are rescaled by
    batch size / number of groups."""

    def __init__(self, model: VAE, optimizer: Optimizer, hyperparameter: dict[str, float], verbose_period: int, device: torch.device,
                 transform=None) -> None:
        super().__init__(model, optimizer, verbose_period, device, transform)
        self.hyperparameter = hyperparameter
        self.annealer = LogisticAnnealer(loc=hyperparameter["loc"], scale=hyperparameter["scale"], beta=hyperparameter["beta"])

    def fit(self, epochs: int, train_loader: DataLoader, valid_loader: None | DataLoader = None, eval_evidence_acc: bool = False):
        for epoch in range(epochs):
            verbose = (epoch % self.verbose_period) == 0
            self._train(train_loader, verbose, epoch)
            if valid_loader is not None:
                self._valid(valid_loader, verbose, epoch, eval_evidence_acc)

    def _group_adjust(self, B, m, *losses):
        "B: batch size; m: number of groups"
        return [loss * B / m for loss in losses]

    def train_step(self, X, label):
        """one iteration of trainer.py:338-362; returns (reconstr_loss, kl_c, kl_s) as logged there (after the group adjust)"""
        vae = self.model
        batch_size, n_groups = X.size(0), len(label.unique())
        self.optimizer.zero_grad()
        X_hat, latent_params = vae(X, label=label)
        rec, kl_c, kl_s = vae_loss(X_hat, X, **latent_params)
        rec, kl_s = self._group_adjust(batch_size, n_groups, rec, kl_s)
        loss = rec + self.annealer(kl_c) + self.annealer(kl_s)
        loss.backward()
        fused_adam_step(self.optimizer)
        self._refresh_packs()
        self.annealer.step()
        return rec.detach(), kl_c.detach(), kl_s.detach()

    def _train(self, dataloader: DataLoader, verbose: bool, epoch_id: int):
        self.model.train()
        with tqdm(dataloader, unit="batch", disable=not verbose) as bar:
            bar.set_description(f"epoch {epoch_id}")
            for batch in bar:
                X, label = self._batch(batch)
                rec, kl_c, kl_s = self.train_step(X, label)
                if verbose:
                    bar.set_postfix(reconstr_loss=float(rec), kl_c=float(kl_c), kl_s=float(kl_s))

    def _valid(self, dataloader, verbose, epoch_id, with_evidence_acc=False):
        if verbose:
            mig, mse = self.evaluate(dataloader, verbose, epoch_id, with_evidence_acc)
            print(f"gMIG: {round(mig, 3)}; mse: {round(float(mse), 3)}")

    def evaluate(self, dataloader, verbose, epoch_id, with_evidence_acc=False):
        vae = self.model
        vae.eval()
        tot, n = None, 0
        labels, lat_c, lat_s = [], [], []
        with torch.no_grad():
            for batch in tqdm(dataloader, disable=not verbose, desc=f"val-epoch {epoch_id}"):
                X, label = self._batch(batch)
                X_hat, latent_params, z = vae(X, label, explicit=True) if with_evidence_acc else vae(X, explicit=True)
                vals = torch.stack(vae_loss(X_hat, X, **latent_params))
                tot = vals if tot is None else tot + vals
                n += 1
                labels.append(label)
                lat_c.append(z[:, :vae.z_dim])
                lat_s.append(z[:, vae.z_dim:])
        from .metrics import mutual_info_gap
        mig = mutual_info_gap(torch.cat(labels), torch.cat(lat_c), torch.cat(lat_s))
        t = (tot / n).tolist()
        if verbose:
            print("val_recontr_loss={:.3f}, val_kl_c={:.3f}, val_kl_s={:.3f}".format(*t))
        return mig, float(t[0])
