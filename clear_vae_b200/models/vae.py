"""`VAE` (28x28) and `VAE64` (64x64) with the reference's constructor signatures, attribute
names and `state_dict()` keys (`code/src/models/vae.py:7-156`), executed by the sm_100a
engine (`clear_vae_b200.engine`).

The sub-modules (`encoder`, `mu_c`, ..., `decoder`) are ordinary `torch.nn` containers that
only *hold* parameters / buffers (so checkpoints move both ways and the RNG consumption at
construction equals the reference's); they are never called.  All arithmetic goes through the
fused conv / BatchNorm / latent kernels.  There is no CPU execution path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..engine import DecoderFn, EncoderFn, Engine, LayerSpec
from ..latent import latent_block

__all__ = ["VAE", "VAE64"]


def _conv_stack(chans, k, transposed, out_pads=None, final_sigmoid=False, head=()):
    """Conv/BN/ReLU triplets in the reference's order (module indices matter for state_dict keys)."""
    mods = list(head)
    n = len(chans) - 1
    for i in range(n):
        if transposed:
            mods.append(nn.ConvTranspose2d(chans[i], chans[i + 1], k, 2, 1, out_pads[i] if out_pads else 0))
        else:
            mods.append(nn.Conv2d(chans[i], chans[i + 1], k, 2, 1))
        mods.append(nn.BatchNorm2d(chans[i + 1]))
        mods.append(nn.Sigmoid() if (final_sigmoid and i == n - 1) else nn.ReLU())
    return mods


class VAE(nn.Module):
    """Weakly-supervised VAE with content / style latent halves (reference vae.py:7-102)."""

    _K, _ENC, _DEC, _UNFLAT, _OUT_PADS, _IMG = 3, (32, 64, 128), (128, 64, 32), (128, 4, 4), (0, 1, 1), 28
    # numerics of the conv / linear stacks: "bf16" (fast path: bf16 operands, fp32 accumulate) or "fp32x3" (fp32 activations,
    # bf16 x 3 split products on the tensor cores — fp32-grade, for parity runs against the reference's fp32 arithmetic)
    conv_precision = "bf16"

    def __init__(self, total_z_dim, in_channel: int = 1, group_mode: str | None = None) -> None:
        super().__init__()
        self.mode = group_mode
        self.z_dim = int(total_z_dim / 2)
        self.in_channel = in_channel
        self._build(self._K, self._ENC, self._DEC, self._UNFLAT, self._OUT_PADS, in_channel)
        self._engine = None

    # -- construction ---------------------------------------------------------------
    def _build(self, k, enc, dec, unflat, out_pads, in_channel):
        self.encoder = nn.Sequential(*_conv_stack((in_channel,) + tuple(enc), k, False), nn.Flatten())
        self.mu_c = nn.Linear(2048, self.z_dim)
        self.logvar_c = nn.Linear(2048, self.z_dim)
        self.mu_s = nn.Linear(2048, self.z_dim)
        self.logvar_s = nn.Linear(2048, self.z_dim)
        head = (nn.Linear(self.z_dim * 2, 2048), nn.BatchNorm1d(2048), nn.ReLU(), nn.Unflatten(1, unflat))
        self.decoder = nn.Sequential(*_conv_stack(tuple(dec) + (in_channel,), k, True, out_pads, True, head))
        self._arch = (k, tuple(enc), tuple(dec), tuple(unflat), tuple(out_pads), self._IMG_FOR(k))

    @staticmethod
    def _IMG_FOR(k):
        return 28 if k == 3 else 64

    # -- engine ----------------------------------------------------------------------
    def _eng(self) -> Engine:
        if self._engine is None:
            k, enc, dec, unflat, out_pads, img = self._arch
            enc_specs, h = [], img
            chans = (self.in_channel,) + enc
            for i in range(len(enc)):
                sp = LayerSpec(False, k, 2, 1, 0, chans[i], chans[i + 1], h, 3 * i, 3 * i + 1)
                enc_specs.append(sp)
                h = sp.hout
            dec_specs, h = [], unflat[1]
            chans = dec + (self.in_channel,)
            for i in range(len(dec)):
                sp = LayerSpec(True, k, 2, 1, out_pads[i], chans[i], chans[i + 1], h, 4 + 3 * i, 5 + 3 * i)
                dec_specs.append(sp)
                h = sp.hout
            self._engine = Engine(enc_specs, dec_specs)
        e = self._engine
        e.set_precision(self.conv_precision)
        e.training = self.training
        e.dist, e.sync_bn = getattr(self, "dist", None), bool(getattr(self, "sync_bn", False))
        e.enc_buffers = [(self.encoder[s.bn].running_mean, self.encoder[s.bn].running_var) for s in e.enc_specs]
        e.dec_buffers = [(self.decoder[s.bn].running_mean, self.decoder[s.bn].running_var) for s in e.dec_specs]
        e.dec_fc_buffers = (self.decoder[1].running_mean, self.decoder[1].running_var)
        return e

    def _tick(self, bns, n=1):
        if self.training:
            torch._foreach_add_([b.num_batches_tracked for b in bns], n)

    def _enc_params(self):
        out = []
        for s in self._eng().enc_specs:
            c, b = self.encoder[s.conv], self.encoder[s.bn]
            out += [c.weight, c.bias, b.weight, b.bias]
        return out

    def _dec_params(self):
        out = []
        for s in self._eng().dec_specs:
            c, b = self.decoder[s.conv], self.decoder[s.bn]
            out += [c.weight, c.bias, b.weight, b.bias]
        return out

    # -- reference API ---------------------------------------------------------------
    def encode(self, x, bn_repeat: int = 1):
        """`bn_repeat` > 1 accounts for that many identical train-mode forwards in one pass: the
        BatchNorm running statistics receive `bn_repeat` momentum updates with the same batch
        statistics (what CLEAR-MIM's 5 inner forwards on an unchanged encoder amount to)."""
        eng = self._eng()
        heads = (self.mu_c, self.logvar_c, self.mu_s, self.logvar_s)
        hw = torch.cat([h.weight for h in heads], 0)
        hb = torch.cat([h.bias for h in heads], 0)
        eng.bn_repeat = int(bn_repeat)
        try:
            lat = EncoderFn.apply(eng, x, hw, hb, *self._enc_params())
        finally:
            eng.bn_repeat = 1
        self._tick([self.encoder[s.bn] for s in eng.enc_specs], int(bn_repeat))
        D = self.z_dim
        return tuple(lat[:, i * D:(i + 1) * D].contiguous() for i in range(4))

    def decode(self, z):
        return self._decode(z, None)[0]

    def _decode(self, z, target, stats_only=False):
        eng = self._eng()
        fc, bn = self.decoder[0], self.decoder[1]
        eng.stats_only = bool(stats_only)
        try:
            xhat, recon = DecoderFn.apply(eng, z, target, fc.weight, fc.bias, bn.weight, bn.bias, *self._dec_params())
        finally:
            eng.stats_only = False
        self._tick([bn] + [self.decoder[s.bn] for s in eng.dec_specs])
        return xhat, recon

    def decode_stats_many(self, zs):
        """Statistics-only train-mode decoder passes over several latent batches (CLEAR-MIM's inner forwards,
        trainer.py:874-888, whose outputs are discarded): only the BatchNorm running statistics change.

        The passes are independent except for those running statistics, so they run as parallel branches (side streams in
        eager mode, fork/join nodes in a captured graph): every branch owns its statistic accumulators and writes its
        batch mean / unbiased variance into its own scratch row (momentum 1); after the join the n momentum updates are
        applied in order in closed form:  r <- (1-m)^n r + sum_j m (1-m)^(n-1-j) s_j."""
        eng = self._eng()
        n = len(zs)
        dp_sync = eng.sync_bn and eng.dist is not None and eng.dist.world > 1   # SyncBN collectives stay on one stream
        if not (eng.parallel_stats and self.training and n > 1 and zs[0].is_cuda) or dp_sync:
            for z in zs:
                self._decode(z, None, stats_only=True)
            return
        dev = zs[0].device
        fc, bn = self.decoder[0], self.decoder[1]
        params = self._dec_params()
        eng.prepack_decoder(fc.weight, params)
        scratch, sizes = eng.branch_running(n, dev)
        main = torch.cuda.current_stream(dev)
        side = eng.branch_streams(n - 1, dev)
        real_fc, real_dec = eng.dec_fc_buffers, eng.dec_buffers
        eng.stats_only, eng.momentum = True, 1.0
        try:
            for j in range(n):
                st = main if j == 0 else side[j - 1]
                if j:
                    st.wait_stream(main)
                rows = list(scratch[j].split([c for c in sizes for _ in (0, 1)]))
                eng.dec_fc_buffers = (rows[0], rows[1])
                eng.dec_buffers = [(rows[2 + 2 * i], rows[3 + 2 * i]) for i in range(len(eng.dec_specs))]
                eng.stat_tag = j
                with torch.cuda.stream(st):
                    DecoderFn.apply(eng, zs[j], None, fc.weight, fc.bias, bn.weight, bn.bias, *params)
        finally:
            eng.stats_only, eng.momentum, eng.stat_tag = False, 0.1, 0
            eng.dec_fc_buffers, eng.dec_buffers = real_fc, real_dec
        for st in side:
            main.wait_stream(st)
        m = 0.1
        coef = getattr(self, "_stat_coef", None)
        if coef is None or coef.numel() != n or coef.device != dev:
            coef = self._stat_coef = torch.tensor([m * (1 - m) ** (n - 1 - j) for j in range(n)], dtype=torch.float32, device=dev)
        comb = torch.mv(scratch.t(), coef)                       # [sum 2C]
        real = [real_fc[0], real_fc[1]] + [b for pair in real_dec for b in pair]
        torch._foreach_mul_(real, (1 - m) ** n)
        torch._foreach_add_(real, list(comb.split([c for c in sizes for _ in (0, 1)])))
        self._tick([bn] + [self.decoder[s.bn] for s in eng.dec_specs], n)

    def sample(self, mu, logvar):
        """Reparameterisation (vae.py:56-60); the noise is drawn exactly like the reference
        (`randn_like` on a tensor of logvar's shape), the arithmetic runs in the latent kernel."""
        eps = torch.randn_like(logvar)
        dummy = torch.zeros(mu.shape[0], dtype=torch.int64, device=mu.device)
        z, _ = latent_block([mu], [logvar], [eps], dummy, snn=[0], ps=[0])
        return z

    def generate(self, mu_c, logvar_c, mu_s, logvar_s, g_dict: dict | None = None, explicit=False):
        if g_dict is not None:
            # ML-VAE / GVAE (vae.py:69-73): the content code is sampled group-wise from the accumulated evidence, the style code
            # per sample; same order of random draws as the reference (groups on the CPU generator first, then `randn_like`)
            from ..group import groupwise_reparam_each
            z_c, _, _ = groupwise_reparam_each(mu_c, logvar_c, g_dict)
            z_s = self.sample(mu_s, logvar_s)
            z = torch.cat([z_c, z_s], dim=-1)
        else:
            eps_c = torch.randn_like(logvar_c)  # c first, then s: same Philox consumption as vae.py:70,73
            eps_s = torch.randn_like(logvar_s)
            dummy = torch.zeros(mu_c.shape[0], dtype=torch.int64, device=mu_c.device)
            z, _ = latent_block([mu_c, mu_s], [logvar_c, logvar_s], [eps_c, eps_s], dummy, snn=[0, 0], ps=[0, 0])
        xhat = self.decode(z)
        return (xhat, z) if explicit else xhat

    def forward(self, x, label=None, explicit=False) -> tuple:
        mu_c, logvar_c, mu_s, logvar_s = self.encode(x)
        g_dict = None
        if label is not None:  # ML-VAE / GVAE: content parameters become the [G, D] group evidence (vae.py:84-88)
            from ..group import accumulate_group_evidence
            mu_c, logvar_c, g_dict = accumulate_group_evidence(mu_c, logvar_c, label, mode=self.mode)
        latent_params = {"mu_c": mu_c, "logvar_c": logvar_c, "mu_s": mu_s, "logvar_s": logvar_s}
        if g_dict is not None:
            if explicit:
                xhat, z = self.generate(mu_c, logvar_c, mu_s, logvar_s, g_dict, True)
                return xhat, latent_params, z
            return self.generate(mu_c, logvar_c, mu_s, logvar_s, g_dict, False), latent_params
        if explicit:
            xhat, z = self.generate(mu_c, logvar_c, mu_s, logvar_s, None, True)
            return xhat, latent_params, z
        return self.generate(mu_c, logvar_c, mu_s, logvar_s, None, False), latent_params

    # -- fused training-step forward (used by the trainers) ----------------------------
    def fused_step_forward(self, x, label, *, temperature, snn, ps, sim_fn="cosine", eps=None, dist=None, gather_z=False):
        """encode -> [reparam + KL + SNN terms] -> decode + reconstruction error.

        Returns (xhat, recon, z, scalars, latent_params); `scalars` as in `latent_block`.
        Equivalent to `forward(x, explicit=True)` followed by `vae_loss` and the
        `contrastive_loss` calls of the trainers (trainer.py:452-470), with the same
        random draws.  `gather_z` (data parallel only) appends the sampled latents of the global batch.
        """
        mu_c, logvar_c, mu_s, logvar_s = self.encode(x)
        if eps is None:
            eps = (torch.randn_like(logvar_c), torch.randn_like(logvar_s))
        out = latent_block([mu_c, mu_s], [logvar_c, logvar_s], list(eps), label, snn=snn, ps=ps, sim_fn=sim_fn,
                           temperature=temperature, dist=dist, gather_z=gather_z)
        z, sc = out[0], out[1]
        xhat, recon = self._decode(z, x)
        lp = {"mu_c": mu_c, "logvar_c": logvar_c, "mu_s": mu_s, "logvar_s": logvar_s}
        if len(out) == 3:       # data parallel + gather_z: the sampled latents of the global batch as a sixth value
            return xhat, recon, z, sc, lp, out[2]
        return xhat, recon, z, sc, lp


class VAE64(VAE):
    """64x64 variant (reference vae.py:105-156).  Like the reference it first builds the
    28x28 layers and then replaces them, which keeps the construction-time RNG stream —
    and therefore seeded initial weights — identical."""

    def __init__(self, total_z_dim, in_channel: int = 3, group_mode: str | None = None) -> None:
        super().__init__(total_z_dim, in_channel, group_mode)
        self.z_dim = int(total_z_dim / 2)
        self._build(4, (32, 64, 128, 256, 512), (512, 256, 128, 64, 32), (512, 2, 2), (0, 0, 0, 0, 0), in_channel)
        self._engine = None
