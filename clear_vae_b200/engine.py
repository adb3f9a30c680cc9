"""Layer engine: runs the conv encoder / decoder stacks of `VAE` / `VAE64`
(reference `code/src/models/vae.py:15-46,113-156`) on the sm_100a kernels and
implements their backward by hand (no autograd graph inside a stack).

Data layout in HBM
  * activations between layers: raw (pre-BatchNorm) conv outputs, NHWC, bf16 by default
    (`act_dtype`); BatchNorm-apply + ReLU is folded into the *consumer's* operand load, so
    every activation is written once and read once per consumer;
  * boundary tensors keep the reference layout: input `x` and `xhat` are NCHW fp32 (read /
    written through strides), the encoder output feeding the linear heads and the decoder
    fc output are stored in the reference's flatten order (channel-major) so the
    `Linear(2048, D)` / `Linear(2D, 2048)` weights are used unpermuted;
  * weights: fp32 masters in the reference layout (state_dict-compatible) + packed bf16
    K-major copies per GEMM role, re-packed when the master's version counter changes.

Per conv block the forward is 2 launches (GEMM with fused bias + batch-statistics epilogue,
BatchNorm finalize) and the backward 4 (coefficients, dy materialisation, weight-gradient
GEMM, data-gradient GEMM with the previous block's ReLU mask + BatchNorm sums fused into
its epilogue).
"""
from __future__ import annotations

import torch

from . import _ops

F32, BF16 = 0, 1
FPROP, DGRAD = 0, 1
EPI_BIAS_STATS, EPI_MASK_STATS = 0, 1
BN_MOMENTUM, BN_EPS = 0.1, 1e-5

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def nhwc_strides(H, W, C):
    return [H * W * C, W * C, C, 1]


def nchw_strides(C, H, W):
    """(n, h, w, c) strides of an NCHW / channel-major-flattened buffer."""
    return [C * H * W, W, 1, H * W]


def out_hw(transposed, k, s, p, op, h):
    return (h - 1) * s - 2 * p + k + op if transposed else (h + 2 * p - k) // s + 1


class _PackCache:
    """bf16 packed copies of a weight, keyed by GEMM role; refreshed when the master changes.

    Eager mode: an entry is reused while the master's (version counter, data pointer) are unchanged.  Inside a CUDA
    graph capture the version counter says nothing about *replays* (the optimiser node changes the weights every
    replay), so an entry is reused only within the training step that packed it (`epoch`, bumped by the trainers at
    the start of every step): each replay then re-packs every weight once, after which the 5 inner forwards of
    CLEAR-MIM reuse the copy."""

    def __init__(self):
        self.entries = {}
        self.epoch = 0
        self.specs = {}      # (key, role) -> (weight, geom): what refresh_all re-packs
        self._table = None   # (device table, entries, blocks, signature)

    def get(self, key, w: torch.Tensor, geom, role, cacheable=True):
        key = (key, role)
        ent = self.entries.get(key)
        ver = w._version
        if cacheable and ent is not None and ent[0] == ver and ent[1] == w.data_ptr():
            if ent[3] == self.epoch or not torch.cuda.is_current_stream_capturing():
                return ent[2]
        if cacheable and ent is not None and ent[1] == w.data_ptr() and not (role & 16):
            # same master, new values: re-pack into the SAME buffer (refresh_all's table keeps pointing at it)
            packed = ent[2]
            tab, ne, nb = self._upload(*_ops.ops().conv_pack_multi_build(list(geom), [role], [w.detach()], [packed]), w.device)
            _ops.ops().conv_pack_multi(tab, ne, nb)
        else:
            packed = _ops.ops().conv_pack_weight(geom, role, w.detach().contiguous())
            if cacheable:
                self._table = None
        self.entries[key] = (ver, w.data_ptr(), packed, self.epoch)
        if cacheable and not (role & 16):
            self.specs[key] = (w, list(geom))
        return packed

    def _upload(self, host, ne, nb, device):
        """pinned host table -> device, stream-ordered (legal inside a graph capture: a memcpy node from static pinned memory)"""
        pinned = host.pin_memory()
        dev = torch.empty(host.shape, dtype=host.dtype, device=device)
        dev.copy_(pinned, non_blocking=True)
        self._alive = getattr(self, "_alive", [])[-63:] + [(pinned, dev)]   # keep both alive while launches / replays read them
        return dev, ne, nb

    def refresh_all(self):
        """Re-pack every cached operand from its master in ONE launch (call right after the optimiser step): the entries are
        then valid for the rest of this step and for the next one, whose GEMMs find them without launching anything."""
        keys = [k for k in self.specs if k in self.entries and self.entries[k][1] == self.specs[k][0].data_ptr()]
        if not keys:
            return
        sig = tuple((k, self.entries[k][2].data_ptr()) for k in keys)
        if self._table is None or self._table[3] != sig:
            geoms = [v for k in keys for v in self.specs[k][1]]
            tab, ne, nb = self._upload(*_ops.ops().conv_pack_multi_build(geoms, [k[1] for k in keys], [self.specs[k][0].detach() for k in keys],
                                                                         [self.entries[k][2] for k in keys]), self.specs[keys[0]][0].device)
            self._table = (tab, ne, nb, sig)
        tab, ne, nb, _ = self._table
        _ops.ops().conv_pack_multi(tab, ne, nb)
        for k in keys:
            w = self.specs[k][0]
            self.entries[k] = (w._version, w.data_ptr(), self.entries[k][2], self.epoch + 1)


class LayerSpec:
    __slots__ = ("transposed", "k", "s", "p", "op", "cin", "cout", "hin", "hout", "geom", "conv", "bn")

    def __init__(self, transposed, k, s, p, op, cin, cout, hin, conv, bn):
        self.transposed, self.k, self.s, self.p, self.op = transposed, k, s, p, op
        self.cin, self.cout, self.hin = cin, cout, hin
        self.hout = out_hw(transposed, k, s, p, op, hin)
        self.geom = [int(transposed), k, s, p, op, cin, cout, hin, hin]
        self.conv, self.bn = conv, bn  # module names inside the Sequential


def linear_geom(k, n):
    return [0, 1, 1, 0, 0, k, n, 1, 1]


def _stats(C, dev):
    # [sum | sum of squares] + one ticket word used by the fused finalize+apply kernel (kept zero between launches)
    return torch.zeros(2 * C + 2, dtype=torch.float64, device=dev)


def _reduce_stats(eng, st):
    """SyncBN: sum the per-channel accumulators over the data-parallel group (SURVEY.md §8e "BN")."""
    d = eng.dist
    if eng.sync_bn and d is not None and d.world > 1:
        import torch.distributed as td
        td.all_reduce(st, group=d.group)
        return d.world
    return 1


def _ws(dev):
    from .latent import _workspace
    return _workspace(dev, max(_ops.ops().bn_act_workspace_bytes(), 1 << 16))


# ======================================================================================
# encoder: x -> latent parameters [B, 4D]
# ======================================================================================
class EncoderFn(torch.autograd.Function):
    """forward(x, heads_w, heads_b, *[conv_w, conv_b, bn_w, bn_b] per layer) -> lat [B, 4D] fp32"""

    @staticmethod
    def forward(ctx, eng, x, heads_w, heads_b, *params):
        ops = _ops.ops()
        specs = eng.enc_specs
        B = x.shape[0]
        dev = x.device
        x = x.contiguous()
        need_grad = any(ctx.needs_input_grad)
        src, src_strides, pre = x, nchw_strides(specs[0].cin, specs[0].hin, specs[0].hin), None
        saved_raw, saved_pre, saved_stats, saved_act = [], [], [], []
        n = len(specs)
        for i, sp in enumerate(specs):
            w, b, gamma, beta = params[4 * i:4 * i + 4]
            rm, rv = eng.enc_buffers[i]
            last = i == n - 1
            H = sp.hout
            # last block: the heads read the reference's Flatten (channel-major) order.  In training the GEMM still
            # writes channels-last (fast epilogue) and the fused finalize kernel emits the channel-major copies.
            tr_last = (last and eng.training and eng.materialize and eng.act_dtype == torch.bfloat16 and i > 0
                       and H * H * (sp.cout + 2) * 2 + 8 * sp.cout <= 48 * 1024)
            if last and not tr_last:
                raw = torch.empty(B, sp.cout * H * H, dtype=eng.act_dtype, device=dev)
                dst_strides = nchw_strides(sp.cout, H, H)
            else:
                raw = torch.empty(B, H, H, sp.cout, dtype=eng.act_dtype, device=dev)
                dst_strides = nhwc_strides(H, H, sp.cout)
            st = eng.stat_buf(("enc", i), sp.cout, dev) if eng.training else None
            # boundary layer (Cin <= 4): direct CUDA-core kernel; everything else: tcgen05 implicit GEMM
            if not (i == 0 and eng.use_direct and ops.conv_direct_fwd(sp.geom, B, src, src_strides, None, None, False,
                                                                      w.detach(), b.detach(), raw, dst_strides, st)):
                pw = eng.packs.get(("enc", i), w, sp.geom, eng.role(FPROP))
                ops.conv_gemm(sp.geom, eng.role(FPROP), B, src, src_strides, pre[0] if pre else None, pre[1] if pre else None,
                              pre is not None, pw, b, raw, dst_strides, EPI_BIAS_STATS, None, [0, 0, 0, 0], None, None, st)
            mat = eng.materialize and raw.dtype == torch.bfloat16
            act = None
            if eng.training:
                nw = _reduce_stats(eng, st)
                if mat:
                    # statistics -> scale/shift, running estimates, and relu(bn(raw)) written once in bf16, in one
                    # launch: every consumer GEMM then cp.async-copies its operand
                    act, scale, shift, mean, invstd, raw_cm = ops.bn_finalize_apply(
                        st, sp.cout, float(nw * B * H * H), gamma, beta, rm, rv, BN_MOMENTUM, BN_EPS, H * H if last else 1,
                        eng.bn_repeat, raw, (3 if tr_last else 1) if last else 0, H * H if last else 1)
                    if tr_last:
                        raw, dst_strides = raw_cm, nchw_strides(sp.cout, H, H)
                else:
                    scale, shift, mean, invstd = ops.bn_finalize(st, sp.cout, 1, float(nw * B * H * H), gamma, beta, rm, rv,
                                                                 BN_MOMENTUM, BN_EPS, H * H if last else 1, eng.bn_repeat)
            else:
                scale, shift, mean, invstd = eng.eval_affine(gamma, beta, rm, rv, H * H if last else 1)
                if mat:
                    act = ops.bn_relu_apply(raw, scale, shift, scale.numel())
            saved_raw.append((raw, dst_strides))
            saved_pre.append((scale, shift))
            saved_stats.append((mean, invstd))
            if mat:
                saved_act.append(act)
                src, src_strides, pre = act, dst_strides, None
            else:
                saved_act.append(None)
                src, src_strides, pre = raw, dst_strides, (scale, shift)
        # linear heads on the flattened (channel-major) encoder output
        K = specs[-1].cout * specs[-1].hout ** 2
        N = heads_w.shape[0]
        lat = torch.empty(B, N, dtype=torch.float32, device=dev)
        hg = linear_geom(K, N)
        pw = eng.packs.get("heads", heads_w, hg, eng.role(FPROP), cacheable=False)
        ops.conv_gemm(hg, eng.role(FPROP), B, src, [K, 0, 0, 1], pre[0] if pre else None, pre[1] if pre else None, pre is not None, pw,
                      heads_b, lat, [N, 0, 0, 1], EPI_BIAS_STATS, None, [0, 0, 0, 0], None, None, None)
        if eng.debug is not None:
            eng.debug["enc_raw"] = [r for r, _ in saved_raw]
        if need_grad:
            ctx.eng = eng
            ctx.saved = (x, saved_raw, saved_pre, saved_stats, saved_act, heads_w, params)
        return lat

    @staticmethod
    def backward(ctx, dlat):
        ops = _ops.ops()
        eng = ctx.eng
        x, saved_raw, saved_pre, saved_stats, saved_act, heads_w, params = ctx.saved
        specs = eng.enc_specs
        if not eng.training:
            raise RuntimeError("clear_vae_b200: backward through eval-mode BatchNorm is not implemented")
        B = x.shape[0]
        dev = x.device
        n = len(specs)
        dlat = dlat.contiguous()
        K = specs[-1].cout * specs[-1].hout ** 2
        N = heads_w.shape[0]
        hg = linear_geom(K, N)
        raw_last, _ = saved_raw[-1]
        sc_last, sh_last = saved_pre[-1]
        # heads: weight / bias gradients
        keep = [dlat]   # operands of side-stream kernels stay referenced until the join
        with eng.wgrad_branch(dev):
            d_heads_w = torch.zeros_like(heads_w)
            if saved_act[-1] is not None:
                ops.conv_wgrad(hg, B, saved_act[-1], [K, 0, 0, 1], None, None, False, dlat, [N, 0, 0, 1], d_heads_w, eng.x3)
            else:
                ops.conv_wgrad(hg, B, raw_last, [K, 0, 0, 1], sc_last, sh_last, True, dlat, [N, 0, 0, 1], d_heads_w, eng.x3)
            d_heads_b = ops.colsum(dlat)
        # heads: data gradient with the last block's ReLU mask + BatchNorm sums in the epilogue
        g = torch.empty(B, K, dtype=eng.grad_dtype, device=dev)
        st = eng.stat_buf(("enc_b", n - 1), K, dev)
        pw = eng.packs.get("heads", heads_w, hg, eng.role(DGRAD), cacheable=False)
        ops.conv_gemm(hg, eng.role(DGRAD), B, dlat, [N, 0, 0, 1], None, None, False, pw, None, g, [K, 0, 0, 1], EPI_MASK_STATS,
                      raw_last, [K, 0, 0, 1], sc_last, sh_last, st)
        grads = [None] * len(params)
        for i in range(n - 1, -1, -1):
            sp = specs[i]
            w, b, gamma, beta = params[4 * i:4 * i + 4]
            raw, raw_strides = saved_raw[i]
            mean, invstd = saved_stats[i]
            H = sp.hout
            last = i == n - 1
            group = H * H if last else 1
            nw = _reduce_stats(eng, st)
            coef, dgamma, dbeta = ops.bn_bwd_coef(st, sp.cout, group, float(nw * B * H * H), gamma, mean, invstd)
            if nw > 1:  # sums were global: undo the later rank-averaging's double count of the affine grads
                dgamma, dbeta = dgamma / nw, dbeta / nw
            # the last block's tensors are channel-major (flatten order of the heads); its dy is written channels-last
            dy = ops.bn_bwd_apply(g, raw, None, None, None, coef, sp.cout, group, last, eng.dy_code())
            if last:
                raw_strides = nhwc_strides(H, H, sp.cout)
            if i == 0:
                src, src_strides, pre = x, nchw_strides(sp.cin, sp.hin, sp.hin), None
            else:
                src, src_strides = saved_raw[i - 1]
                pre = saved_pre[i - 1]
            keep.append(dy)
            with eng.wgrad_branch(dev):
                dw = torch.zeros_like(w)
                if i == 0 and eng.use_direct and ops.conv_direct_wgrad(sp.geom, B, src, src_strides, dy, raw_strides, dw):
                    pass  # Cin <= 4: register-blocked CUDA-core kernel
                elif i > 0 and saved_act[i - 1] is not None:
                    ops.conv_wgrad(sp.geom, B, saved_act[i - 1], src_strides, None, None, False, dy, raw_strides, dw, eng.x3)
                else:
                    ops.conv_wgrad(sp.geom, B, src, src_strides, pre[0] if pre else None, pre[1] if pre else None,
                                   pre is not None, dy, raw_strides, dw, eng.x3)
            grads[4 * i], grads[4 * i + 2], grads[4 * i + 3] = dw, dgamma, dbeta
            # grads[4*i+1] (conv bias) stays None: a bias feeding a train-mode BatchNorm has exactly zero gradient
            if i > 0:
                g = torch.empty(B, sp.hin, sp.hin, sp.cin, dtype=eng.grad_dtype, device=dev)
                st = eng.stat_buf(("enc_b", i - 1), sp.cin, dev)
                pwd = eng.packs.get(("enc", i), w, sp.geom, eng.role(DGRAD))
                ops.conv_gemm(sp.geom, eng.role(DGRAD), B, dy, raw_strides, None, None, False, pwd, None, g, nhwc_strides(sp.hin, sp.hin, sp.cin),
                              EPI_MASK_STATS, src, src_strides, pre[0], pre[1], st)
        eng.wgrad_join(dev)
        del keep
        return (None, None, d_heads_w, d_heads_b, *grads)


# ======================================================================================
# decoder: z -> xhat (NCHW fp32) [+ fused reconstruction term]
# ======================================================================================
class DecoderFn(torch.autograd.Function):
    """forward(z, target|None, fc_w, fc_b, fc_bn_w, fc_bn_b, *[w, b, bn_w, bn_b] per convT) -> (xhat, recon)"""

    @staticmethod
    def forward(ctx, eng, z, target, fc_w, fc_b, fc_g, fc_beta, *params):
        ops = _ops.ops()
        specs = eng.dec_specs
        B = z.shape[0]
        dev = z.device
        z = z.contiguous()
        need_grad = any(ctx.needs_input_grad)
        K0, N0 = fc_w.shape[1], fc_w.shape[0]
        fg = linear_geom(K0, N0)
        raw_fc = torch.empty(B, N0, dtype=torch.float32, device=dev)
        st = eng.stat_buf(("dec_fc",), N0, dev) if eng.training else None
        if not (eng.use_direct and ops.fc_fwd(z, fc_w.detach(), fc_b.detach(), raw_fc, st)):   # K = 2D <= 64: fp32 CUDA-core kernel
            ops.conv_gemm(fg, eng.role(FPROP), B, z, [K0, 0, 0, 1], None, None, False, eng.packs.get("fc", fc_w, fg, eng.role(FPROP)), fc_b, raw_fc,
                          [N0, 0, 0, 1], EPI_BIAS_STATS, None, [0, 0, 0, 0], None, None, st)
        rm, rv = eng.dec_fc_buffers
        C0, H0 = specs[0].cin, specs[0].hin
        # BatchNorm1d + ReLU, written channels-last so the first transposed conv copies 16-byte channel runs
        if eng.training and eng.materialize and eng.act_dtype == torch.bfloat16 and C0 % 8 == 0:
            nw = _reduce_stats(eng, st)
            a_fc, sc, sh, mean_fc, inv_fc, _ = ops.bn_finalize_apply(st, N0, float(nw * B), fc_g, fc_beta, rm, rv, eng.momentum,
                                                                      BN_EPS, 1, 1, raw_fc, 2, H0 * H0)
        else:
            if eng.training:
                nw = _reduce_stats(eng, st)
                sc, sh, mean_fc, inv_fc = ops.bn_finalize(st, N0, 1, float(nw * B), fc_g, fc_beta, rm, rv, eng.momentum, BN_EPS, 1, 1)
            else:
                sc, sh, mean_fc, inv_fc = eng.eval_affine(fc_g, fc_beta, rm, rv, 1)
            a_fc, _ = ops.bn_act_fwd(raw_fc, sc, sh, N0, 1, 1, C0, H0 * H0, _DT[eng.act_dtype], None, B, _ws(dev))
        src, src_strides, pre = a_fc, nhwc_strides(H0, H0, C0), None
        saved_raw, saved_pre, saved_stats, saved_act = [], [], [], []
        n = len(specs)
        for j, sp in enumerate(specs):
            w, b, gamma, beta = params[4 * j:4 * j + 4]
            rm, rv = eng.dec_buffers[j]
            last = j == n - 1
            H = sp.hout
            if last:
                # statistics-only forwards (CLEAR-MIM inner loop) need the last layer's batch moments, not its output
                raw = torch.empty((0,) if (eng.stats_only and eng.use_direct and j > 0) else (B, sp.cout, H, H),
                                  dtype=torch.float32, device=dev)
                dst_strides = nchw_strides(sp.cout, H, H)
            else:
                raw = torch.empty(B, H, H, sp.cout, dtype=eng.act_dtype, device=dev)
                dst_strides = nhwc_strides(H, H, sp.cout)
            st = eng.stat_buf(("dec", j), sp.cout, dev) if eng.training else None
            if not (last and eng.use_direct and j > 0 and
                    ops.conv_direct_fwd(sp.geom, B, src, src_strides, pre[0] if pre else None, pre[1] if pre else None,
                                        pre is not None, w.detach(), b.detach(), raw, dst_strides, st)):
                if raw.numel() == 0:
                    raw = torch.empty(B, sp.cout, H, H, dtype=torch.float32, device=dev)
                ops.conv_gemm(sp.geom, eng.role(FPROP), B, src, src_strides, pre[0] if pre else None, pre[1] if pre else None,
                              pre is not None, eng.packs.get(("dec", j), w, sp.geom, eng.role(FPROP)), b, raw, dst_strides, EPI_BIAS_STATS,
                              None, [0, 0, 0, 0], None, None, st)
            mat = eng.materialize and not last and raw.dtype == torch.bfloat16
            act = None
            if eng.training:
                nw = _reduce_stats(eng, st)
                if mat:
                    act, scale, shift, mean, invstd, _ = ops.bn_finalize_apply(st, sp.cout, float(nw * B * H * H), gamma, beta, rm,
                                                                               rv, eng.momentum, BN_EPS, 1, 1, raw, 0, 1)
                else:
                    scale, shift, mean, invstd = ops.bn_finalize(st, sp.cout, 1, float(nw * B * H * H), gamma, beta, rm, rv,
                                                                 eng.momentum, BN_EPS, 1, 1)
            else:
                scale, shift, mean, invstd = eng.eval_affine(gamma, beta, rm, rv, 1)
                if mat:
                    act = ops.bn_relu_apply(raw, scale, shift, sp.cout)
            saved_raw.append((raw, dst_strides))
            saved_pre.append((scale, shift))
            saved_stats.append((mean, invstd))
            if mat:
                saved_act.append(act)
                src, src_strides, pre = act, dst_strides, None
            else:
                saved_act.append(None)
                src, src_strides, pre = raw, dst_strides, (scale, shift)
        sp = specs[-1]
        H = sp.hout
        if eng.stats_only:  # caller only wants the BatchNorm running-statistic side effects (CLEAR-MIM inner loop)
            e = torch.empty(0, device=dev)
            return e, e
        xhat, recon = ops.bn_act_fwd(src, pre[0], pre[1], sp.cout, H * H, 2, 0, 0, F32, target, B, _ws(dev))
        if eng.debug is not None:
            eng.debug["dec_raw"] = [r for r, _ in saved_raw]
            eng.debug["fc_raw"], eng.debug["fc_act"] = raw_fc, a_fc
        if need_grad:
            ctx.eng = eng
            ctx.has_target = target is not None
            ctx.saved = (z, target, fc_w, fc_g, raw_fc, a_fc, (sc, sh), mean_fc, inv_fc, saved_raw, saved_pre, saved_stats,
                         saved_act, params)
            ctx.save_for_backward(xhat)
        return xhat, recon

    @staticmethod
    def backward(ctx, d_xhat, d_recon):
        ops = _ops.ops()
        eng = ctx.eng
        (z, target, fc_w, fc_g, raw_fc, a_fc, fc_aff, mean_fc, inv_fc, saved_raw, saved_pre, saved_stats, saved_act,
         params) = ctx.saved
        (xhat,) = ctx.saved_tensors
        if not eng.training:
            raise RuntimeError("clear_vae_b200: backward through eval-mode BatchNorm is not implemented")
        specs = eng.dec_specs
        B = z.shape[0]
        dev = z.device
        n = len(specs)
        sp = specs[-1]
        H = sp.hout
        raw_last, last_strides = saved_raw[-1]
        st = eng.stat_buf(("dec_b", n - 1), sp.cout, dev)
        gr = d_recon.contiguous() if (ctx.has_target and d_recon is not None) else None
        ge = d_xhat.contiguous() if d_xhat is not None else None
        tgt = target if ctx.has_target else xhat  # without a target the MSE term is absent (gr is None)
        g = ops.sigmoid_mse_bwd(xhat, tgt, gr, ge, raw_last, sp.cout, H * H, B, st)
        grads = [None] * len(params)
        keep = []   # operands of side-stream kernels stay referenced until the join
        for j in range(n - 1, -1, -1):
            sp = specs[j]
            w, b, gamma, beta = params[4 * j:4 * j + 4]
            raw, raw_strides = saved_raw[j]
            mean, invstd = saved_stats[j]
            H = sp.hout
            last = j == n - 1
            inner = H * H if last else 1
            nw = _reduce_stats(eng, st)
            coef, dgamma, dbeta = ops.bn_bwd_coef(st, sp.cout, 1, float(nw * B * H * H), gamma, mean, invstd)
            if nw > 1:
                dgamma, dbeta = dgamma / nw, dbeta / nw
            dy = ops.bn_bwd_apply(g, raw, None, None, None, coef, sp.cout, inner, False, eng.dy_code())
            if j == 0:
                src, src_strides, pre = a_fc, nhwc_strides(sp.hin, sp.hin, sp.cin), None
            else:
                src, src_strides = saved_raw[j - 1]
                pre = saved_pre[j - 1]
            keep.append(dy)
            with eng.wgrad_branch(dev):
                dw = torch.zeros_like(w)
                if (last and j > 0 and eng.use_direct and saved_act[j - 1] is not None and
                        ops.conv_direct_wgrad(sp.geom, B, saved_act[j - 1], src_strides, dy, raw_strides, dw)):
                    pass  # Cout <= 4: register-blocked CUDA-core kernel
                elif j > 0 and saved_act[j - 1] is not None:
                    ops.conv_wgrad(sp.geom, B, saved_act[j - 1], src_strides, None, None, False, dy, raw_strides, dw, eng.x3)
                else:
                    ops.conv_wgrad(sp.geom, B, src, src_strides, pre[0] if pre else None, pre[1] if pre else None,
                                   pre is not None, dy, raw_strides, dw, eng.x3)
            grads[4 * j], grads[4 * j + 2], grads[4 * j + 3] = dw, dgamma, dbeta
            if j > 0:
                g = torch.empty(B, sp.hin, sp.hin, sp.cin, dtype=eng.grad_dtype, device=dev)
                st = eng.stat_buf(("dec_b", j - 1), sp.cin, dev)
                # Cout <= 4: the data gradient is a Conv2d(Cout -> 32) of dy — direct CUDA-core kernel
                if not (last and eng.use_direct and ops.conv_direct_dgrad(sp.geom, B, dy, raw_strides, w.detach(), g,
                                                                          nhwc_strides(sp.hin, sp.hin, sp.cin), src, src_strides,
                                                                          pre[0], pre[1], st)):
                    ops.conv_gemm(sp.geom, eng.role(DGRAD), B, dy, raw_strides, None, None, False, eng.packs.get(("dec", j), w, sp.geom, eng.role(DGRAD)),
                                  None, g, nhwc_strides(sp.hin, sp.hin, sp.cin), EPI_MASK_STATS, src, src_strides, pre[0], pre[1], st)
            else:
                # gradient w.r.t. the activated fc output, channel-major like a_fc
                N0 = fc_w.shape[0]
                g_a = torch.empty(B, N0, dtype=torch.float32, device=dev)
                ops.conv_gemm(sp.geom, eng.role(DGRAD), B, dy, raw_strides, None, None, False, eng.packs.get(("dec", j), w, sp.geom, eng.role(DGRAD)), None, g_a,
                              nchw_strides(sp.cin, sp.hin, sp.hin), EPI_BIAS_STATS, None, [0, 0, 0, 0], None, None, None)
        # fc block: BatchNorm1d + ReLU backward, then Linear
        K0, N0 = fc_w.shape[1], fc_w.shape[0]
        fg = linear_geom(K0, N0)
        st = eng.stat_buf(("dec_fc_b",), N0, dev)
        ops.bn_reduce(raw_fc, g_a, None, fc_aff[0], fc_aff[1], N0, 1, 1, st)  # ReLU mask recomputed from the raw fc output
        nw = _reduce_stats(eng, st)
        coef, d_fc_g, d_fc_beta = ops.bn_bwd_coef(st, N0, 1, float(nw * B), fc_g, mean_fc, inv_fc)
        if nw > 1:
            d_fc_g, d_fc_beta = d_fc_g / nw, d_fc_beta / nw
        dy_fc = ops.bn_bwd_apply(g_a, raw_fc, None, fc_aff[0], fc_aff[1], coef, N0, 1, False, F32)
        keep.append(dy_fc)
        with eng.wgrad_branch(dev):
            d_fc_w = torch.zeros_like(fc_w)
            ops.conv_wgrad(fg, B, z, [K0, 0, 0, 1], None, None, False, dy_fc, [N0, 0, 0, 1], d_fc_w, eng.x3)
        dz = torch.empty(B, K0, dtype=torch.float32, device=dev)
        ops.conv_gemm(fg, eng.role(DGRAD), B, dy_fc, [N0, 0, 0, 1], None, None, False, eng.packs.get("fc", fc_w, fg, eng.role(DGRAD)), None, dz,
                      [K0, 0, 0, 1], EPI_BIAS_STATS, None, [0, 0, 0, 0], None, None, None)
        eng.wgrad_join(dev)
        del keep
        return (None, dz, None, d_fc_w, None, d_fc_g, d_fc_beta, *grads)


class Engine:
    """Per-model execution state: layer specs, packed-weight cache, statistic accumulators."""

    def __init__(self, enc_specs, dec_specs, act_dtype=torch.bfloat16):
        self.enc_specs, self.dec_specs = enc_specs, dec_specs
        self.act_dtype = act_dtype
        self.grad_dtype = torch.float32
        self.packs = _PackCache()
        self.training = True
        self.enc_buffers, self.dec_buffers, self.dec_fc_buffers = [], [], None
        self._stat = {}
        self.debug = None  # set to a dict to capture raw activations (tests / tools only)
        self.bn_repeat = 1    # momentum updates per encoder forward (see VAE.encode)
        self.stats_only = False
        self.dist = None      # DistSpec for SyncBN
        self.sync_bn = False
        self.use_direct = True   # direct kernels for the Cin<=4 / Cout<=4 boundary layers
        self.materialize = True  # write relu(bn(raw)) once in bf16 so the GEMM operand loads are pure cp.async copies
        self.overlap_wgrad = True  # weight gradients on a side stream, overlapping the data-gradient chain
        self._side = {}
        self.x3 = False          # fp32-grade convolutions: bf16 x 3 split products on fp32 activations (set_precision)
        self.momentum = BN_MOMENTUM   # running-statistic momentum of the decoder forwards (1.0 inside parallel stat branches)
        self.stat_tag = 0             # statistic-accumulator set: concurrent decoder passes must not share accumulators
        self.parallel_stats = True    # CLEAR-MIM's statistics-only decoder passes as parallel graph branches
        self._branch = {}

    # ---- numerics ------------------------------------------------------------------------------------------
    def set_precision(self, mode: str):
        """'bf16' (default): bf16 activations / operands, fp32 accumulation — the fast path.
        'fp32x3': fp32 activations and gradients in HBM, every tensor-core product as hi*hi + lo*hi + hi*lo of the bf16 split
        of both operands (CLEARVAE_ROLE_SPLIT3) — the accuracy class of the reference's fp32 run, ~3x the MMA work."""
        if mode not in ("bf16", "fp32x3"):
            raise ValueError("precision must be 'bf16' or 'fp32x3'")
        x3 = mode == "fp32x3"
        if x3 != self.x3:
            self.x3 = x3
            self.act_dtype = torch.float32 if x3 else torch.bfloat16
            self.materialize = not x3
            self.packs = _PackCache()

    def role(self, r):
        return r | 16 if self.x3 else r

    def dy_code(self):
        return F32 if self.x3 else BF16

    # ---- weight-gradient branch -----------------------------------------------------------------------------
    # dW of a block needs only that block's dy and input activation, while the data-gradient chain continues to the
    # previous block: the weight-gradient kernels are issued on a side stream (a fork/join inside a captured graph),
    # so two latency-bound kernels share the SMs instead of running back to back.
    class _Branch:
        def __init__(self, side, main):
            self.side, self.main, self.ctx = side, main, None

        def __enter__(self):
            self.side.wait_stream(self.main)
            self.ctx = torch.cuda.stream(self.side)
            self.ctx.__enter__()
            return self

        def __exit__(self, *exc):
            return self.ctx.__exit__(*exc)

    N_WGRAD_STREAMS = 3

    def wgrad_branch(self, dev):
        """the weight-gradient kernels of different layers are independent of each other too: they rotate over a few side
        streams, so the backward's critical path is the data-gradient chain, not the sum of the weight-gradient kernels
        (0.46 ms serial on one side stream vs 0.35 ms of data-gradient chain at configs[1])"""
        if not self.overlap_wgrad:
            import contextlib
            return contextlib.nullcontext()
        pool = self._side.get(dev)
        if pool is None:
            pool = self._side[dev] = [torch.cuda.Stream(device=dev) for _ in range(self.N_WGRAD_STREAMS)]
        self._side_rr = (getattr(self, "_side_rr", -1) + 1) % len(pool)
        return Engine._Branch(pool[self._side_rr], torch.cuda.current_stream(dev))

    def wgrad_join(self, dev):
        pool = self._side.get(dev)
        if self.overlap_wgrad and pool is not None:
            cur = torch.cuda.current_stream(dev)
            for st in pool:
                cur.wait_stream(st)

    def branch_streams(self, n, dev):
        st = self._branch.setdefault(("streams", dev), [])
        while len(st) < n:
            st.append(torch.cuda.Stream(device=dev))
        return st[:n]

    def branch_running(self, n, dev):
        """[n, sum of 2C over the decoder's BatchNorm layers] zero-initialised scratch: row j receives pass j's batch mean /
        unbiased variance of every layer (momentum 1), in the order (fc, convT 0, convT 1, ...) x (mean, var)."""
        sizes = [self.dec_specs[0].cin * self.dec_specs[0].hin ** 2] + [sp.cout for sp in self.dec_specs]
        key = ("running", n, dev)
        t = self._branch.get(key)
        if t is None:
            t = self._branch[key] = torch.zeros(n, 2 * sum(sizes), dtype=torch.float32, device=dev)
        return t, sizes

    def prepack_decoder(self, fc_w, params):
        """pack every decoder weight the forward GEMMs will ask for on the CURRENT stream (before branches fork)."""
        specs = self.dec_specs
        for j, sp in enumerate(specs):
            last = j == len(specs) - 1
            if last and self.use_direct and j > 0:
                continue   # direct CUDA-core kernel reads the fp32 master
            self.packs.get(("dec", j), params[4 * j], sp.geom, self.role(FPROP))

    def stat_buf(self, key, C, dev):
        k = (key, C, dev, self.stat_tag)
        t = self._stat.get(k)
        if t is None:
            t = _stats(C, dev)
            self._stat[k] = t
        return t

    @staticmethod
    def eval_affine(gamma, beta, rm, rv, expand):
        invstd = torch.rsqrt(rv + BN_EPS)
        scale = gamma * invstd
        shift = beta - rm * scale
        if expand > 1:
            scale = scale.repeat_interleave(expand)
            shift = shift.repeat_interleave(expand)
        return scale.contiguous(), shift.contiguous(), rm, invstd
