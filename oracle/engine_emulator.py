"""Torch restatement of the *engine's* arithmetic (TEST INFRASTRUCTURE ONLY).

The CUDA engine (clear_vae_b200/engine.py) does not use autograd inside the conv stacks: it
runs a hand-derived backward (BatchNorm-backward as an affine map, ReLU masks from the stored
raw activations, bias gradients of BN-fed convs dropped) and rounds MMA operands / stored
activations to bf16.  This module restates exactly that procedure in plain torch so that

  (1) with `round_bf16=False` it can be checked against autograd of the fp32 oracle
      (oracle/model_oracle.py, itself pinned to the reference goldens) — proving the manual
      backward is the true gradient, to 1e-5;
  (2) with `round_bf16=True` it predicts what the tensor-core path should produce, rounding
      included, so the CUDA result can be held to a tight tolerance instead of the loose
      "bf16 convs within 1e-2" envelope.

Layer semantics: reference vae.py:15-46,113-156 (SURVEY.md §8a').
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .model_oracle import BN_EPS, BN_MOMENTUM, arch_spec


class Emulator:
    def __init__(self, st, arch, in_channel, round_bf16=True, update_running=False, direct_boundary=True):
        self.st, self.arch, self.cin = st, arch, in_channel
        self.round = round_bf16
        self.update_running = update_running
        # the engine runs the first conv / last conv-transpose *forward* on fp32 CUDA-core kernels (no operand
        # rounding there); their data- and weight-gradients still go through the bf16 tensor-core GEMMs
        self.direct = direct_boundary

    def r(self, x):
        return x.to(torch.bfloat16).to(x.dtype) if self.round else x

    # ---- BatchNorm pieces -----------------------------------------------------------
    def _bn_fwd(self, acc, prefix, dims):
        st = self.st
        accd = acc.double()
        mean = accd.mean(dims)
        var = (accd * accd).mean(dims) - mean * mean
        var = var.clamp_min(0)
        invstd = (1.0 / torch.sqrt(var + BN_EPS)).to(acc.dtype)
        mean = mean.to(acc.dtype)
        g, b = st[f"{prefix}.weight"], st[f"{prefix}.bias"]
        scale = g * invstd
        shift = b - mean * scale
        if self.update_running:
            n = acc.numel() / acc.shape[1]
            st[f"{prefix}.running_mean"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean)
            st[f"{prefix}.running_var"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var.float() * n / (n - 1))
            st[f"{prefix}.num_batches_tracked"] += 1
        return scale, shift, mean, invstd

    @staticmethod
    def _bc(v, ndim):
        return v.view(1, -1, *([1] * (ndim - 2)))

    def _bn_bwd(self, g, y, prefix, mean, invstd, dims):
        """dy = a*g + b*y + c (per channel) and BatchNorm parameter gradients."""
        gd, yd = g.double(), y.double()
        s1 = gd.sum(dims)
        s2 = (gd * yd).sum(dims)
        count = g.numel() / g.shape[1]
        mu, rr, gam = mean.double(), invstd.double(), self.st[f"{prefix}.weight"].double()
        sgx = rr * (s2 - mu * s1)
        a = (gam * rr).to(g.dtype)
        b = (-gam * rr * rr * sgx / count).to(g.dtype)
        c = (gam * rr * (mu * rr * sgx - s1) / count).to(g.dtype)
        nd = g.dim()
        dy = self._bc(a, nd) * g + (self._bc(b, nd) * y + self._bc(c, nd))
        return dy, sgx.to(g.dtype), s1.to(g.dtype)

    # ---- forward ----------------------------------------------------------------------
    def forward(self, x, eps_c, eps_s, target=None):
        st, r = self.st, self.r
        enc, unflat, dec = arch_spec(self.arch, self.cin)
        tape = {"x": x}
        a = r(x)
        tape["enc"] = []
        for li, (i, ci, co, k, s, p) in enumerate(enc):
            w = r(st[f"encoder.{i}.weight"])
            if li == 0 and self.direct:
                acc = F.conv2d(x, st[f"encoder.{i}.weight"], st[f"encoder.{i}.bias"], stride=s, padding=p)
            else:
                acc = F.conv2d(a, w, st[f"encoder.{i}.bias"], stride=s, padding=p)
            scale, shift, mean, invstd = self._bn_fwd(acc, f"encoder.{i + 1}", (0, 2, 3))
            y = r(acc)
            a_in = a
            a = r(torch.relu(y * self._bc(scale, 4) + self._bc(shift, 4)))
            tape["enc"].append(dict(a_in=a_in, y=y, scale=scale, shift=shift, mean=mean, invstd=invstd, idx=i, s=s, p=p))
        h = a.flatten(1)
        hw = r(torch.cat([st[f"{n}.weight"] for n in ("mu_c", "logvar_c", "mu_s", "logvar_s")], 0))
        hb = torch.cat([st[f"{n}.bias"] for n in ("mu_c", "logvar_c", "mu_s", "logvar_s")], 0)
        lat = F.linear(h, hw, hb)
        tape["h"], tape["hw"] = h, hw
        D = lat.shape[1] // 4
        mu_c, lv_c, mu_s, lv_s = (lat[:, j * D:(j + 1) * D] for j in range(4))
        z = torch.cat([mu_c + eps_c * torch.exp(0.5 * lv_c), mu_s + eps_s * torch.exp(0.5 * lv_s)], 1)
        out = self.decode(z, target, tape)
        out.update(lat=lat, z=z, tape=tape)
        return out

    def decode(self, z, target, tape):
        st, r = self.st, self.r
        _, unflat, dec = arch_spec(self.arch, self.cin)
        zr = r(z)
        wfc = r(st["decoder.0.weight"])
        if self.direct:   # the engine's fc forward is an fp32 CUDA-core kernel (its gradients stay bf16 tensor-core GEMMs)
            raw_fc = F.linear(z, st["decoder.0.weight"], st["decoder.0.bias"])
        else:
            raw_fc = F.linear(zr, wfc, st["decoder.0.bias"])
        sc, sh, mean_fc, inv_fc = self._bn_fwd(raw_fc, "decoder.1", (0,))
        a_fc = r(torch.relu(raw_fc * sc + sh))
        tape["fc"] = dict(zr=zr, wfc=wfc, raw=raw_fc, a=a_fc, mean=mean_fc, invstd=inv_fc)
        a = a_fc.unflatten(1, unflat)
        tape["dec"] = []
        n = len(dec)
        for j, (i, ci, co, k, s, p, op) in enumerate(dec):
            w = r(st[f"decoder.{i}.weight"])
            last = j == n - 1
            if last and self.direct and j > 0:
                # direct fp32 kernel: the materialised (bf16-rounded) activation times the fp32 master weights
                acc = F.conv_transpose2d(a, st[f"decoder.{i}.weight"], st[f"decoder.{i}.bias"], stride=s, padding=p,
                                         output_padding=op)
            else:
                acc = F.conv_transpose2d(a, w, st[f"decoder.{i}.bias"], stride=s, padding=p, output_padding=op)
            scale, shift, mean, invstd = self._bn_fwd(acc, f"decoder.{i + 1}", (0, 2, 3))
            y = acc if last else r(acc)
            a_in = a
            if last:
                a = torch.sigmoid(y * self._bc(scale, 4) + self._bc(shift, 4))
            else:
                a = r(torch.relu(y * self._bc(scale, 4) + self._bc(shift, 4)))
            tape["dec"].append(dict(a_in=a_in, y=y, scale=scale, shift=shift, mean=mean, invstd=invstd, idx=i, s=s, p=p, op=op))
        xhat = a
        recon = None
        if target is not None:
            recon = ((xhat - target) ** 2).flatten(1).sum(1).mean()
        return dict(xhat=xhat, recon=recon)

    # ---- backward ---------------------------------------------------------------------
    @staticmethod
    def _lin_grads(fn, inputs, dy):
        ins = [t.detach().requires_grad_(True) for t in inputs]
        with torch.enable_grad():
            out = fn(*ins)
        return torch.autograd.grad(out, ins, dy)

    def backward_decoder(self, tape, target, g_recon=1.0, g_xhat=None):
        """returns (grads dict, dz)."""
        st, r = self.st, self.r
        grads = {}
        dec = tape["dec"]
        last = dec[-1]
        xhat = torch.sigmoid(last["y"] * self._bc(last["scale"], 4) + self._bc(last["shift"], 4))
        B = xhat.shape[0]
        g = torch.zeros_like(xhat)
        if target is not None:
            g = g + g_recon * 2.0 / B * (xhat - target)
        if g_xhat is not None:
            g = g + g_xhat
        g = g * xhat * (1 - xhat)
        for j in range(len(dec) - 1, -1, -1):
            L = dec[j]
            i = L["idx"]
            dy, dgam, dbeta = self._bn_bwd(g, L["y"], f"decoder.{i + 1}", L["mean"], L["invstd"], (0, 2, 3))
            dy = r(dy)
            w = r(st[f"decoder.{i}.weight"])
            da, dw = self._lin_grads(lambda a_, w_: F.conv_transpose2d(a_, w_, None, stride=L["s"], padding=L["p"],
                                                                       output_padding=L["op"]), (L["a_in"], w), dy)
            grads[f"decoder.{i}.weight"], grads[f"decoder.{i + 1}.weight"], grads[f"decoder.{i + 1}.bias"] = dw, dgam, dbeta
            if self.direct and j == len(dec) - 1 and j > 0:
                # the last layer's data gradient runs on the direct fp32 kernel: unrounded master weights
                (da,) = self._lin_grads(lambda a_: F.conv_transpose2d(a_, st[f"decoder.{i}.weight"], None, stride=L["s"],
                                                                      padding=L["p"], output_padding=L["op"]), (L["a_in"],), dy)
            if j > 0:
                P = dec[j - 1]
                mask = (P["y"] * self._bc(P["scale"], 4) + self._bc(P["shift"], 4)) > 0
                g = da * mask
            else:
                g_a = da.flatten(1)
        fc = tape["fc"]
        g_m = g_a * (fc["a"] > 0)
        dy_fc, dgam, dbeta = self._bn_bwd(g_m, fc["raw"], "decoder.1", fc["mean"], fc["invstd"], (0,))
        grads["decoder.1.weight"], grads["decoder.1.bias"] = dgam, dbeta
        dyr = r(dy_fc)
        dz, dwfc = self._lin_grads(lambda z_, w_: F.linear(z_, w_), (fc["zr"], fc["wfc"]), dyr)
        grads["decoder.0.weight"] = dwfc
        return grads, dz

    def backward_encoder(self, tape, dlat):
        st, r = self.st, self.r
        grads = {}
        enc = tape["enc"]
        dl = r(dlat)
        dh, dhw = self._lin_grads(lambda h_, w_: F.linear(h_, w_), (tape["h"], tape["hw"]), dl)
        D = dlat.shape[1] // 4
        for j, n in enumerate(("mu_c", "logvar_c", "mu_s", "logvar_s")):
            grads[f"{n}.weight"] = dhw[j * D:(j + 1) * D]
            grads[f"{n}.bias"] = dlat[:, j * D:(j + 1) * D].sum(0)
        last = enc[-1]
        mask = (last["y"] * self._bc(last["scale"], 4) + self._bc(last["shift"], 4)) > 0
        g = dh.view_as(last["y"]) * mask
        for k in range(len(enc) - 1, -1, -1):
            L = enc[k]
            i = L["idx"]
            dy, dgam, dbeta = self._bn_bwd(g, L["y"], f"encoder.{i + 1}", L["mean"], L["invstd"], (0, 2, 3))
            dy = r(dy)
            w = r(st[f"encoder.{i}.weight"])
            # first layer with the direct kernels: its weight gradient multiplies the unrounded fp32 image
            a_in = tape["x"] if (k == 0 and self.direct) else L["a_in"]
            da, dw = self._lin_grads(lambda a_, w_: F.conv2d(a_, w_, None, stride=L["s"], padding=L["p"]), (a_in, w), dy)
            grads[f"encoder.{i}.weight"], grads[f"encoder.{i + 1}.weight"], grads[f"encoder.{i + 1}.bias"] = dw, dgam, dbeta
            if k > 0:
                P = enc[k - 1]
                mask = (P["y"] * self._bc(P["scale"], 4) + self._bc(P["shift"], 4)) > 0
                g = da * mask
        return grads
