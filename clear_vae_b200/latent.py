"""Autograd front-end of the fused latent-head loss block.

`latent_block` is one forward launch + one backward launch for everything the
reference does between `encode` and `decode` plus its latent losses:
  reparameterisation (vae.py:56-60), Gaussian KL (losses.py:48-49) and the
  contrastive / anti-contrastive SNN terms (losses.py:98-137),
for one or two latent heads.  Under data parallelism the column side is the
all-gathered global batch (`DistSpec`); the backward needs only the gathered
forward row statistics (no reduce-scatter of dZ — see DESIGN.md §3).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _ops

SIM_IDS = {"cosine": 0, "l2": 1, "modified_l2": 2, "jeffrey": 3, "mahalanobis": 4}
_LOGVAR_SIMS = (2, 3, 4)   # similarities of losses.py:62-84 that also read (and give gradient to) logvar
LOSS_IDS = {"snn_loss": 0, "supcon_in_loss": 1, "supcon_out_loss": 2}

# indices into the packed scalar vector (include/clearvae_b200.h)
S_KL0, S_KL1, S_LOSS0, S_LOSS1, S_SUM0, S_SUM1, S_CNT0, S_CNT1 = range(8)

_workspaces: dict = {}
import os as _os
REDUNDANT_ROWS_MAX = int(_os.environ.get("CLEARVAE_DP_REDUNDANT_ROWS", "4096"))   # data parallel: up to this global batch every rank computes all rows' statistics itself (one exchange per step)


def _stream_key(device):
    return torch.cuda.current_stream(device).cuda_stream if device.type == "cuda" else 0


def _workspace(device, nbytes, tag="fwd"):
    """Zero-initialised scratch the kernels leave zeroed (tickets, partial sums).  One buffer per (device, stream, tag): the
    layout of tickets / partials depends on the problem shape (callers put it in `tag`), and two launches that may run
    concurrently on different streams must not share one."""
    key = (device.type, device.index, _stream_key(device), tag)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


@dataclass
class DistSpec:
    """Data-parallel context: rows [rank*B, (rank+1)*B) of a global batch."""
    group: object
    rank: int
    world: int
    peer: object = None   # clear_vae_b200.peer.PeerComm: one-kernel exchanges over NVLink peer memory instead of NCCL


def _all_gather_rows(t, dist):
    import torch.distributed as td
    out = torch.empty((dist.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    td.all_gather_into_tensor(out, t.contiguous(), group=dist.group)
    return out


class _LatentBlock(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, label, *flat):
        n = cfg["n"]
        mu = [t.contiguous() for t in flat[0:n]]
        logvar = [None if t is None else t.contiguous() for t in flat[n:2 * n]]
        eps = [None if t is None else t.contiguous() for t in flat[2 * n:3 * n]]
        ops = _ops.ops()
        dist = cfg["dist"]
        B, D = mu[0].shape
        label = label.contiguous()
        want_z = cfg["want_z"]
        use_lv = cfg["sim"] in _LOGVAR_SIMS
        if dist is None or dist.world == 1:
            ws = _workspace(mu[0].device, ops.latent_workspace_bytes(B, B, D, n), ("fwd", B, B, D, n))
            z, scalars, stats = ops.latent_fwd(mu, logvar, eps, [None] * n, [None] * n, label, None, cfg["snn"], cfg["ps"], 0,
                                               cfg["sim"], cfg["loss"], cfg["tau"], True, want_z, ws)
            cols, lv_cols, label_cols, stats_all, row_off = [None] * n, [None] * n, None, list(stats), 0
        else:
            snn_terms = [i for i in range(n) if cfg["snn"][i]]
            cols, lv_cols = [None] * n, [None] * n
            row_off = dist.rank * B
            Bg = dist.world * B
            supcon = cfg["loss"] != 0           # SupCon row losses: the op returns [stats..., aux...] (one more per-row statistic)
            # ---- one exchange per step (global batch <= REDUNDANT_ROWS_MAX): reparameterisation + KL of the local rows first, so
            # the sampled z can travel WITH the similarity operands and labels in a single gather; every rank then evaluates the
            # row statistics of ALL global rows itself (Bg x Bg pairs instead of B x Bg) — cheaper than a second exchange of the
            # [B, 2] statistics, whose cost is the synchronisation with the slowest rank, not its bytes, as long as the global
            # batch is small: measured at B = 1024 per GPU, 2 GPUs 1.71 -> 1.61 ms per step; at 8 GPUs (8192 global rows) the
            # all-rows forward costs 0.13 ms more than it saves, hence the 4096-row limit.  Values are those of the single-process
            # global batch, identical on all ranks.
            redundant = bool(snn_terms) and Bg <= REDUNDANT_ROWS_MAX and not use_lv and not supcon
            if redundant:
                ws = _workspace(mu[0].device, ops.latent_workspace_bytes(B, B, D, n), ("fwd", B, B, D, n))
                z, scalars, _ = ops.latent_fwd(mu, logvar, eps, [None] * n, [None] * n, label, None, [0] * n, cfg["ps"], 0,
                                               cfg["sim"], cfg["loss"], cfg["tau"], True, want_z, ws)
                pieces = [mu[i] for i in snn_terms] + [label] + ([z] if cfg["gather_z"] else [])
                if dist.peer is not None and dist.peer.fits(pieces):
                    got = dist.peer.gather(pieces)
                else:
                    got = [_all_gather_rows(t, dist) for t in pieces]
                for k, i in enumerate(snn_terms):
                    cols[i] = got[k]
                label_cols = got[len(snn_terms)]
                z_all = got[-1] if cfg["gather_z"] else None
                gmu = [cols[i] if cols[i] is not None else cols[snn_terms[0]] for i in range(n)]   # terms without an SNN loss do nothing here
                ws = _workspace(mu[0].device, ops.latent_workspace_bytes(Bg, Bg, D, n), ("fwd", Bg, Bg, D, n))
                _, sc_g, stats_g = ops.latent_fwd(gmu, [None] * n, [None] * n, [None] * n, [None] * n, label_cols, None, cfg["snn"],
                                                  cfg["ps"], 0, cfg["sim"], cfg["loss"], cfg["tau"], True, False, ws)
                scalars = torch.cat([scalars[:2], sc_g[2:]])      # (kl of the local rows | loss, sum, count of the global batch)
                stats_all = [stats_g[i] if cfg["snn"][i] else None for i in range(n)]
            else:
                pieces = [mu[i] for i in snn_terms] + ([logvar[i] for i in snn_terms] if use_lv else []) + [label]
                if dist.peer is not None and dist.peer.fits(pieces):
                    # one kernel over NVLink peer memory: every operand lands as its own contiguous [Bg, ...] tensor
                    got = dist.peer.gather(pieces)
                    for k, i in enumerate(snn_terms):
                        cols[i] = got[k]
                        if use_lv:
                            lv_cols[i] = got[len(snn_terms) + k]
                    label_cols = got[-1]
                else:
                    # one packed NCCL all-gather of the similarity operands + labels
                    per = 2 * D if use_lv else D   # logvar travels too for the logvar-dependent similarities
                    packed = torch.cat([torch.cat([mu[i], logvar[i]], 1) if use_lv else mu[i] for i in snn_terms]
                                       + [label.view(B, 1).view(torch.float32)], dim=1)
                    g = _all_gather_rows(packed, dist)
                    for k, i in enumerate(snn_terms):
                        cols[i] = g[:, k * per:k * per + D].contiguous()
                        if use_lv:
                            lv_cols[i] = g[:, k * per + D:(k + 1) * per].contiguous()
                    label_cols = g[:, len(snn_terms) * per:].contiguous().view(torch.int64).view(-1)
                ws = _workspace(mu[0].device, ops.latent_workspace_bytes(B, Bg, D, n), ("fwd", B, Bg, D, n))
                z, scalars, stats = ops.latent_fwd(mu, logvar, eps, cols, lv_cols, label, label_cols, cfg["snn"], cfg["ps"], row_off,
                                                   cfg["sim"], cfg["loss"], cfg["tau"], False, want_z, ws)
                stats_all = [None] * (2 * n if supcon else n)
                mine = [stats[i] for i in snn_terms] + ([stats[n + i] for i in snn_terms] if supcon else [])
                # the sampled latents ride along with the row statistics when the caller needs the global batch of z
                # (CLUB-S / L1OutUB bounds: mi_estimator.py:138-143, 170-191 pair rows across the whole batch)
                z_all = None
                if cfg["gather_z"]:
                    mine = mine + [z]
                if mine and dist.peer is not None and dist.peer.fits(mine):
                    got = dist.peer.gather(mine)
                elif mine:
                    got = [_all_gather_rows(t, dist) for t in mine]
                if cfg["gather_z"]:
                    z_all = got[-1]
                for k, i in enumerate(snn_terms):
                    stats_all[i] = got[k]
                    if supcon:
                        stats_all[n + i] = got[len(snn_terms) + k]
                for i in snn_terms:
                    ops.snn_finalize(stats_all[i], i, scalars)
        ctx.cfg = cfg
        ctx.row_off = row_off
        ctx.n_saved = (len(mu), len(logvar), len(eps))
        ctx.n_stats = len(stats_all)
        ctx.save_for_backward(label, label_cols, scalars, *mu, *logvar, *eps, *cols, *lv_cols, *stats_all)
        if cfg["gather_z"] and dist is not None and dist.world > 1:
            return z, scalars, z_all
        return z, scalars

    @staticmethod
    def backward(ctx, dz, dscal, dz_all=None):
        if dz_all is not None:
            # gradient of whatever consumed the gathered latents: this rank owns rows [row_off, row_off + B) — the other rows'
            # gradients are produced (identically) by their owners, so no reduce-scatter is needed
            own = dz_all[ctx.row_off:ctx.row_off + (dz.shape[0] if dz is not None else ctx.saved_tensors[3].shape[0])]
            dz = own if dz is None else dz + own
        cfg = ctx.cfg
        n = cfg["n"]
        saved = ctx.saved_tensors
        label, label_cols, scalars = saved[0], saved[1], saved[2]
        rest = list(saved[3:])
        mu, logvar, eps, cols, lv_cols = (rest[k * n:(k + 1) * n] for k in range(5))
        stats_all = rest[5 * n:5 * n + ctx.n_stats]
        ops = _ops.ops()
        if dscal is None:
            dscal = torch.zeros(8, dtype=scalars.dtype, device=scalars.device)
        dscal = dscal.contiguous()
        dist = cfg["dist"]
        if dist is not None and dist.world > 1:
            # gradients are averaged over ranks afterwards; the SNN term is a *global* loss
            dscal = dscal.clone()
            dscal[S_LOSS0:S_LOSS1 + 1] *= float(dist.world)
        if dz is not None:
            dz = dz.contiguous()
            if dz.numel() == 0:
                dz = None
        B, D = mu[0].shape
        Bg = label_cols.numel() if label_cols is not None else B
        ws = _workspace(mu[0].device, ops.latent_bwd_workspace_bytes(B, Bg, D, n), ("bwd", B, Bg, D, n))   # column-split partials + tickets
        dmu, dlv = ops.latent_bwd(mu, logvar, eps, cols, lv_cols, stats_all, dz, label, label_cols, cfg["snn"], cfg["ps"],
                                  ctx.row_off, cfg["sim"], cfg["loss"], cfg["tau"], scalars, dscal, ws)
        g_mu = list(dmu)
        g_lv = [dlv[i] if logvar[i] is not None else None for i in range(n)]
        g_eps = [None] * n
        return (None, None, *g_mu, *g_lv, *g_eps)


def latent_block(mu, logvar, eps, label, *, snn, ps, sim_fn="cosine", temperature=0.1, loss_name="snn_loss",
                 want_z=True, dist: DistSpec | None = None, gather_z=False):
    """Fused latent block over `len(mu)` heads.

    mu/logvar/eps: lists of [B, D] tensors (logvar/eps entries may be None);
    snn[i]: whether head i carries a contrastive term; ps[i]: its pair-switch flag.
    Returns (z [B, n*D] or None, scalars[8]) with kl_i = scalars[i],
    loss_i = scalars[2 + i] (global finite-row mean), count_i = scalars[6 + i].
    With `gather_z` under data parallelism a third value is returned: the sampled latents of the global batch
    [world * B, n * D] (differentiable: the gradient of this rank's rows flows back into `z`).
    """
    if sim_fn not in SIM_IDS:
        raise ValueError("unimplemented similarity measure.")  # losses.py:122-123
    if loss_name not in LOSS_IDS:
        raise NameError(f"name '{loss_name}' is not defined")  # reference: eval(loss_name) (losses.py:124)
    n = len(mu)
    dp = dist is not None and dist.world > 1
    cfg = dict(n=n, snn=[int(bool(s)) for s in snn], ps=[1 if p else 0 for p in ps], sim=SIM_IDS[sim_fn],
               loss=LOSS_IDS[loss_name], tau=float(temperature), want_z=bool(want_z), dist=dist, gather_z=bool(gather_z and dp and want_z))
    out = _LatentBlock.apply(cfg, label, *mu, *logvar, *eps)
    if cfg["gather_z"]:
        return out[0], out[1], out[2]      # (z, scalars, z of the global batch [world * B, n * D], rank-major)
    z, scalars = out
    return (z if want_z else None), scalars


def pair_mask(label, ps=False, label_cols=None, row_offset=0):
    """Debug view of the kernels' pair indexing: uint8 [B, Bg], bit0 = candidate
    (j != i), bit1 = positive (losses.py:107-110,131-135)."""
    return _ops.ops().pair_mask(label.contiguous(), label_cols, int(row_offset), 1 if ps else 0)
