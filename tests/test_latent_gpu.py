"""GPU parity: fused latent-loss kernels vs the reference goldens and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import latent_oracle as lo
from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu
DEV = "cuda"

LOSS_REL, LOSS_ABS = 1e-5, 2e-6   # north_star: loss components within 1e-5 relative in fp32 (+ fp32 LSE floor)
GRAD_REL = 1e-4                   # north_star: gradients within 1e-4


def close(a, b, rel=LOSS_REL, ab=LOSS_ABS):
    if np.isnan(b):
        return np.isnan(a)
    return abs(a - b) <= rel * abs(b) + ab


def cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "contrastive.npz"))
    for name in g["names"]:
        sim, tau, ln, ps = g[f"{name}/meta"]
        yield name, g, str(sim), float(tau), str(ln), eval(str(ps))


def test_contrastive_matches_reference_goldens(golden_dir):
    from clear_vae_b200.losses import contrastive_loss
    n = 0
    for name, g, sim, tau, ln, ps in cases(golden_dir):
        lv_sim = sim in ("jeffrey", "mahalanobis", "modified_l2")
        mu = torch.tensor(g[f"{name}/mu"], device=DEV, requires_grad=True)
        lv = torch.tensor(g[f"{name}/logvar"], device=DEV, requires_grad=lv_sim)
        lab = torch.tensor(g[f"{name}/label"], device=DEV)
        loss = contrastive_loss(mu, lv, lab, sim, tau, ln, ps)
        want = float(g[f"{name}/loss"])
        assert loss.dim() == 0
        # the distance-type similarities reach |s|/tau ~ 1e3: the reference's own fp32 value is only good to ~1e-5 there
        assert close(float(loss), want, rel=LOSS_REL if not lv_sim else 3e-5), (name, float(loss), want)
        if np.isfinite(want):
            loss.backward()
            wg = g[f"{name}/dmu"]
            err = np.abs(mu.grad.cpu().numpy() - wg).max()
            # l2 at tau<=0.5 is sharply peaked: |grad| ~ 1e-3 is itself an fp32 cancellation residue -> absolute floor
            floor = 1e-6 if sim != "cosine" else 1e-7
            assert err <= GRAD_REL * np.abs(wg).max() + floor, (name, err, np.abs(wg).max())
            if lv_sim:   # these similarities also back-propagate into logvar (losses.py:62-84)
                wl = g[f"{name}/dlogvar"]
                errl = np.abs(lv.grad.cpu().numpy() - wl).max()
                assert errl <= GRAD_REL * np.abs(wl).max() + floor, (name, errl, np.abs(wl).max())
            else:
                assert lv.grad is None
        n += 1
    assert n >= 44   # incl. the four SupCon row-loss cases (losses.py:140-170)


def test_pair_mask_bit_exact():
    from clear_vae_b200.latent import pair_mask
    gen = torch.Generator().manual_seed(5)
    for B, ncls in [(1, 2), (7, 3), (257, 10), (1000, 500)]:
        lab = torch.randint(0, ncls, (B,), generator=gen)
        for ps in (False, True, None):
            got = pair_mask(lab.to(DEV), ps).cpu().numpy()
            cand, pos = lo.positive_sets(lab.numpy(), ps)
            assert np.array_equal(got & 1, cand) and np.array_equal(got >> 1, pos), (B, ps)
    # sharded rows against a global column set
    lab = torch.randint(0, 6, (96,), generator=gen)
    for r in range(3):
        rows = lab[r * 32:(r + 1) * 32]
        got = pair_mask(rows.to(DEV), True, label_cols=lab.to(DEV), row_offset=r * 32).cpu().numpy()
        cand, pos = lo.positive_sets(rows.numpy(), True, row_offset=r * 32, label_cols=lab.numpy())
        assert np.array_equal(got & 1, cand) and np.array_equal(got >> 1, pos)


@pytest.mark.parametrize("B,D,ncls,tau,ps", [(128, 8, 10, 0.1, True), (1024, 8, 10, 0.1, False), (512, 32, 4, 0.1, True),
                                              (300, 32, 7, 0.5, None), (64, 16, 32, 0.02, False), (2500, 8, 10, 0.1, True)])
def test_fused_block_matches_oracle(B, D, ncls, tau, ps):
    """reparam + KL + content SNN + style (anti-)SNN in one fwd/bwd pair vs the fp64 oracle."""
    from clear_vae_b200.latent import latent_block
    gen = torch.Generator().manual_seed(B + D)
    mu_c, lv_c, mu_s, lv_s, e_c, e_s = (torch.randn(B, D, generator=gen) * s for s in (1, .3, 1, .3, 1, 1))
    lab = torch.randint(0, ncls, (B,), generator=gen)
    dev = [t.to(DEV).requires_grad_(True) for t in (mu_c, lv_c, mu_s, lv_s)]
    z, sc = latent_block([dev[0], dev[2]], [dev[1], dev[3]], [e_c.to(DEV), e_s.to(DEV)], lab.to(DEV), snn=[1, 1],
                         ps=[False, ps], temperature=tau)
    # oracle in float64 with autograd
    o = [t.double().requires_grad_(True) for t in (mu_c, lv_c, mu_s, lv_s)]
    zc = o[0] + e_c.double() * torch.exp(0.5 * o[1])
    zs = o[2] + e_s.double() * torch.exp(0.5 * o[3])
    kl = lambda m, l: -0.5 * (1 + l - m * m - l.exp()).sum(1).mean()
    c = mo.contrastive(o[0], o[1], lab, "cosine", tau)
    s = mo.contrastive(o[2], o[3], lab, "cosine", tau, ps=ps)
    assert np.abs(z.detach().cpu().numpy() - torch.cat([zc, zs], 1).detach().numpy()).max() < 2e-6
    got = sc.detach().cpu().numpy()
    for g_, w_ in ((got[0], kl(o[0], o[1])), (got[1], kl(o[2], o[3])), (got[2], c), (got[3], s)):
        assert close(float(g_), float(w_)), (float(g_), float(w_))
    # also against the numpy closed form (independent restatement)
    assert close(float(got[2]), lo.contrastive(mu_c.numpy(), lv_c.numpy(), lab.numpy(), "cosine", tau))
    # gradients of a trainer-like combination
    wz = torch.randn(B, 2 * D, generator=gen)
    coef = (0.07, 0.05, 100.0, -100.0 if not ps else 100.0)
    tot = (z * wz.to(DEV)).sum() + coef[0] * sc[0] + coef[1] * sc[1] + coef[2] * sc[2] + coef[3] * sc[3]
    tot.backward()
    tot_o = (torch.cat([zc, zs], 1) * wz.double()).sum() + coef[0] * kl(o[0], o[1]) + coef[1] * kl(o[2], o[3]) + coef[2] * c + coef[3] * s
    tot_o.backward()
    for a, b in zip(dev, o):
        w = b.grad.numpy()
        err = np.abs(a.grad.cpu().numpy() - w).max()
        assert err <= GRAD_REL * np.abs(w).max() + 1e-7, (err, np.abs(w).max())


def test_large_batch_against_chunked_oracle():
    from clear_vae_b200.losses import contrastive_loss
    gen = torch.Generator().manual_seed(9)
    B, D = 8192, 32
    mu = torch.randn(B, D, generator=gen)
    lab = torch.randint(0, 10, (B,), generator=gen)
    for ps in (False, True):
        m = mu.to(DEV).requires_grad_(True)
        got = contrastive_loss(m, torch.zeros_like(m), lab.to(DEV), "cosine", 0.1, ps=ps)
        want = lo.contrastive(mu.numpy(), np.zeros((B, D)), lab.numpy(), "cosine", 0.1, ps=ps)
        assert close(float(got), want), (float(got), want)
        got.backward()
        wg = lo.snn_grad(mu.numpy(), lab.numpy(), "cosine", 0.1, ps)
        err = np.abs(m.grad.cpu().numpy() - wg).max()
        assert err <= GRAD_REL * np.abs(wg).max() + 1e-9


def test_sharded_rows_sum_to_global_loss():
    """size-independent property: per-shard (sum, count) against gathered columns add up to the global loss."""
    from clear_vae_b200 import _ops
    from clear_vae_b200.latent import _workspace
    ops = _ops.ops()
    gen = torch.Generator().manual_seed(3)
    Bg, D, W = 4096, 8, 4
    mu = torch.randn(Bg, D, generator=gen).to(DEV)
    lab = torch.randint(0, 10, (Bg,), generator=gen).to(DEV)
    ws = _workspace(mu.device, ops.latent_workspace_bytes(Bg, Bg, D, 1))
    _, sc_full, st_full = ops.latent_fwd([mu], [None], [None], [None], [None], lab, None, [1], [0], 0, 0, 0, 0.1, True, False, ws)
    B = Bg // W
    parts = []
    for r in range(W):
        rows = mu[r * B:(r + 1) * B].contiguous()
        _, _, st = ops.latent_fwd([rows], [None], [None], [mu], [None], lab[r * B:(r + 1) * B].contiguous(), lab, [1], [0], r * B,
                                  0, 0, 0.1, False, False, ws)
        parts.append(st[0])
    st_all = torch.cat(parts)
    # row statistics do not depend on how rows are sharded (full batch may take the tensor-core path)
    assert torch.allclose(st_all, st_full[0], rtol=0, atol=2e-6)
    sc = torch.zeros(8, device=DEV)
    ops.snn_finalize(st_all, 0, sc)
    assert abs(float(sc[2]) - float(sc_full[2])) <= 1e-6 * abs(float(sc_full[2])) + 1e-7


def test_error_behaviour():
    from clear_vae_b200.losses import contrastive_loss
    mu = torch.randn(8, 4, device=DEV)
    lab = torch.zeros(8, dtype=torch.long, device=DEV)
    with pytest.raises(ValueError, match="unimplemented similarity measure"):
        contrastive_loss(mu, mu, lab, "euclid", 0.1)
    # every row dropped -> nan, not an exception (losses.py:125-126)
    assert torch.isnan(contrastive_loss(mu, mu, lab, "cosine", 0.1, ps=True))


def test_vae_loss_matches_golden(golden_dir):
    from clear_vae_b200.losses import vae_loss
    h = np.load(os.path.join(golden_dir, "heads.npz"))
    t = lambda k: torch.tensor(h[k], device=DEV)
    xh = t("elbo/xhat").requires_grad_(True)
    rec, kc, ks = vae_loss(xh, t("elbo/x"), mu_c=t("elbo/mu_c"), mu_s=t("elbo/mu_s"), logvar_c=t("elbo/logvar_c"),
                           logvar_s=t("elbo/logvar_s"))
    assert close(float(rec), float(h["elbo/recon"])) and close(float(kc), float(h["elbo/kl_c"])) and close(float(ks), float(h["elbo/kl_s"]))
    (3.0 * rec).backward()
    want = 3.0 * 2.0 * (h["elbo/xhat"] - h["elbo/x"]) / h["elbo/x"].shape[0]
    assert np.abs(xh.grad.cpu().numpy() - want).max() < 1e-6


def _set_tc_min_rows(n):
    import ctypes
    from clear_vae_b200 import _ops
    _ops.load()
    lib = ctypes.CDLL(_ops.lib_paths()[0])
    lib.clearvae_set_latent_tc_min_rows(ctypes.c_int32(n))


def test_tensor_core_forward_matches_goldens_and_ffma(golden_dir):
    """The tcgen05 (3xTF32) forward, forced on for every size, against the reference goldens, and
    bit-compatible enough with the FFMA path that the shared backward stays within 1e-4."""
    from clear_vae_b200.losses import contrastive_loss
    try:
        _set_tc_min_rows(1)
        n = 0
        for name, g, sim, tau, ln, ps in cases(golden_dir):
            if sim != "cosine" or ln != "snn_loss" or tau < 0.03:
                continue
            mu = torch.tensor(g[f"{name}/mu"], device=DEV, requires_grad=True)
            lab = torch.tensor(g[f"{name}/label"], device=DEV)
            loss = contrastive_loss(mu, torch.zeros_like(mu), lab, sim, tau, ln, ps)
            want = float(g[f"{name}/loss"])
            assert close(float(loss), want), (name, float(loss), want)
            if np.isfinite(want):
                loss.backward()
                wg = g[f"{name}/dmu"]
                assert np.abs(mu.grad.cpu().numpy() - wg).max() <= GRAD_REL * np.abs(wg).max() + 1e-7, name
            n += 1
        assert n >= 20
        # labels whose high 32 bits differ must still compare as full int64
        gen = torch.Generator().manual_seed(1)
        mu = torch.randn(300, 8, generator=gen)
        lab = torch.randint(0, 3, (300,), generator=gen) + (torch.randint(0, 2, (300,), generator=gen) << 33)
        got = contrastive_loss(mu.to(DEV), mu.to(DEV), lab.to(DEV), "cosine", 0.1)
        want = lo.contrastive(mu.numpy(), mu.numpy(), lab.numpy(), "cosine", 0.1)
        assert close(float(got), want)
    finally:
        _set_tc_min_rows(4096)


@pytest.mark.parametrize("sim", ["modified_l2", "jeffrey", "mahalanobis"])
def test_logvar_similarities_sharded_rows_and_oracle(sim):
    """losses.py:62-84 on the CUDA path: value + both gradients against the fp64 oracle at a size the goldens do not cover,
    and row shards against a gathered column side (mu AND logvar columns) reproduce the full-batch row statistics."""
    from clear_vae_b200 import _ops
    from clear_vae_b200.latent import SIM_IDS, _workspace
    from clear_vae_b200.losses import contrastive_loss
    ops = _ops.ops()
    gen = torch.Generator().manual_seed(11)
    Bg, D, W, tau = 768, 16, 3, 0.5
    mu = (torch.randn(Bg, D, generator=gen) * 0.5)
    lv = (torch.randn(Bg, D, generator=gen) * 0.3)
    lab = torch.randint(0, 6, (Bg,), generator=gen)
    for ps in (False, True):
        m = mu.clone().double().requires_grad_(True)
        l = lv.clone().double().requires_grad_(True)
        want = mo.contrastive(m, l, lab, sim, tau, ps=ps)
        want.backward()
        md, ld = mu.to(DEV).requires_grad_(True), lv.to(DEV).requires_grad_(True)
        got = contrastive_loss(md, ld, lab.to(DEV), sim, tau, ps=ps)
        got.backward()
        assert close(float(got), float(want), rel=3e-5), (sim, ps, float(got), float(want))
        for a, b in ((md.grad, m.grad), (ld.grad, l.grad)):
            assert float((a.cpu().double() - b).abs().max()) <= GRAD_REL * float(b.abs().max()) + 1e-6
    mud, lvd, labd = mu.to(DEV), lv.to(DEV), lab.to(DEV)
    ws = _workspace(mud.device, ops.latent_workspace_bytes(Bg, Bg, D, 1))
    sid = SIM_IDS[sim]
    _, _, st_full = ops.latent_fwd([mud], [lvd], [None], [None], [None], labd, None, [1], [0], 0, sid, 0, tau, True, False, ws)
    B = Bg // W
    parts = []
    for r in range(W):
        sl = slice(r * B, (r + 1) * B)
        _, _, st = ops.latent_fwd([mud[sl].contiguous()], [lvd[sl].contiguous()], [None], [mud], [lvd], labd[sl].contiguous(), labd,
                                  [1], [0], r * B, sid, 0, tau, False, False, ws)
        parts.append(st[0])
    assert torch.allclose(torch.cat(parts), st_full[0], rtol=0, atol=1e-5)


@pytest.mark.parametrize("sim,tau,D,ps", [("cosine", 0.1, 8, False), ("cosine", 0.1, 32, True), ("cosine", 0.02, 8, False),
                                          ("l2", 0.5, 16, False)])
def test_column_split_shard_against_oracle(sim, tau, D, ps):
    """A data-parallel shard (1024 local rows) against the gathered global batch (8192 columns): the FFMA kernels split the
    column sweep over several CTAs per row block (partials merged in a fixed order).  Row statistics, the global loss and
    the shard's gradient rows must match the fp64 oracle, and two runs must agree bit for bit."""
    from clear_vae_b200 import _ops
    from clear_vae_b200.latent import SIM_IDS, _workspace
    ops = _ops.ops()
    gen = torch.Generator().manual_seed(17)
    Bg, W = 8192, 8
    B = Bg // W
    mu_h = torch.randn(Bg, D, generator=gen) * (0.3 if sim == "l2" else 1.0)
    lab_h = torch.randint(0, 10, (Bg,), generator=gen)
    mu, lab = mu_h.to(DEV), lab_h.to(DEV)
    sid = SIM_IDS[sim]
    ws = _workspace(mu.device, ops.latent_workspace_bytes(B, Bg, D, 1))
    stats = []
    for r in range(W):
        rows = mu[r * B:(r + 1) * B].contiguous()
        _, _, st = ops.latent_fwd([rows], [None], [None], [mu], [None], lab[r * B:(r + 1) * B].contiguous(), lab, [1], [int(ps)],
                                  r * B, sid, 0, tau, False, False, ws)
        stats.append(st[0])
    st_all = torch.cat(stats)
    sc = torch.zeros(8, device=DEV)
    ops.snn_finalize(st_all, 0, sc)
    want = lo.contrastive(mu_h.numpy(), np.zeros((Bg, D)), lab_h.numpy(), sim, tau, ps=ps)
    assert close(float(sc[2]), want), (float(sc[2]), want)
    wg = lo.snn_grad(mu_h.numpy(), lab_h.numpy(), sim, tau, ps)
    gscal = torch.tensor([0.0, 0.0, 1.0, 0.0], device=DEV)
    wsb = _workspace(mu.device, ops.latent_bwd_workspace_bytes(B, Bg, D, 1), "bwd")
    for r in (0, 3, W - 1):
        rows = mu[r * B:(r + 1) * B].contiguous()
        args = ([rows], [None], [None], [mu], [None], [st_all], None, lab[r * B:(r + 1) * B].contiguous(), lab, [1], [int(ps)],
                r * B, sid, 0, tau, sc, gscal)
        dmu, _ = ops.latent_bwd(*args, wsb)
        got = dmu[0].cpu().numpy()
        ref = wg[r * B:(r + 1) * B]
        assert np.abs(got - ref).max() <= GRAD_REL * np.abs(wg).max() + 1e-9, (r, np.abs(got - ref).max(), np.abs(wg).max())
        dmu2, _ = ops.latent_bwd(*args, wsb)
        assert torch.equal(dmu[0], dmu2[0])                  # fixed-order merge: bit-reproducible
        dmu1, _ = ops.latent_bwd(*args, None)                 # no workspace -> unsplit sweep, same gradient
        assert np.abs(dmu1[0].cpu().numpy() - got).max() <= 1e-6 * np.abs(wg).max() + 1e-9


from tests.helpers import supcon_torch as _supcon_torch  # noqa: E402  (fp64 restatement, pinned to the goldens on the CPU)


@pytest.mark.parametrize("name", ["supcon_in_loss", "supcon_out_loss"])
@pytest.mark.parametrize("sim,tau,D", [("cosine", 0.1, 8), ("cosine", 0.02, 32), ("l2", 0.5, 16)])
@pytest.mark.parametrize("ps", [False, True])
def test_supcon_row_losses_against_oracle_and_fp64_autograd(name, sim, tau, D, ps):
    """f-1: the SupCon row losses behind the same `contrastive_loss` signature (losses.py:124,140-170): value vs the numpy
    oracle (itself pinned to the reference goldens), gradient vs an fp64 autograd restatement.  The batch holds singleton
    classes (rows without positives are selected out) and, under ps, the n_k = #different - 1 quirk of losses.py:141."""
    from clear_vae_b200.losses import contrastive_loss
    gen = torch.Generator().manual_seed(23)
    B = 333
    mu_h = torch.randn(B, D, generator=gen) * (0.4 if sim == "l2" else 1.0)
    lab_h = torch.randint(0, 7, (B,), generator=gen)
    lab_h[:5] = torch.arange(100, 105)                      # singleton classes
    mu = mu_h.to(DEV).requires_grad_(True)
    got = contrastive_loss(mu, torch.zeros_like(mu), lab_h.to(DEV), sim, tau, name, ps)
    want = lo.contrastive(mu_h.numpy(), np.zeros((B, D)), lab_h.numpy(), sim, tau, loss_name=name, ps=ps)
    assert close(float(got), want, rel=3e-5 if sim == "l2" else LOSS_REL), (float(got), want)
    ref_in = mu_h.clone().double().requires_grad_(True)
    ref = _supcon_torch(ref_in, lab_h, sim, tau, name, ps)
    assert abs(float(ref) - want) <= 1e-9 * abs(want) + 1e-12
    ref.backward()
    got.backward()
    wg = ref_in.grad.numpy()
    err = np.abs(mu.grad.cpu().numpy() - wg).max()
    assert err <= GRAD_REL * np.abs(wg).max() + 1e-7, (err, np.abs(wg).max())


@pytest.mark.parametrize("D", [8, 32])
def test_65536_latents_tensor_core_path_against_fp64(D):
    """The configuration `latent_roofline` is quoted on (BASELINE configs[4] top: 65536 latents, tcgen05 3xTF32 path):
    forward value and the full gradient, both mask polarities in one launch, against the chunked fp64 restatement
    (tests/helpers.py::snn_fp64_chunked, pinned on the CPU to the numpy oracle and through it to the reference goldens).
    Gates: loss 1e-5 relative (+ fp32 LSE floor), gradient 1e-4 of its max — north_star's fp32 tolerances."""
    from clear_vae_b200.latent import latent_block
    from tests.helpers import snn_fp64_chunked
    gen = torch.Generator().manual_seed(65536 + D)
    B = 65536
    mu_c, mu_s = torch.randn(B, D, generator=gen), torch.randn(B, D, generator=gen)
    lab = torch.randint(0, 10, (B,), generator=gen)
    a = mu_c.to(DEV).requires_grad_(True)
    b = mu_s.to(DEV).requires_grad_(True)
    _, sc = latent_block([a, b], [None, None], [None, None], lab.to(DEV), snn=[1, 1], ps=[False, True], temperature=0.1, want_z=False)
    w = torch.zeros(8, device=DEV)
    w[2], w[3] = 100.0, 100.0          # the trainers' alpha
    torch.autograd.backward([sc], [w])
    torch.cuda.synchronize()
    for t, ps, idx in ((a, False, 2), (b, True, 3)):
        want, wg = snn_fp64_chunked(t.detach(), lab, 0.1, ps, device=DEV, chunk=4096)
        assert close(float(sc[idx]), want), (D, ps, float(sc[idx]), want)
        wg = wg * 100.0
        err = float((t.grad.double() - wg).abs().max())
        assert err <= GRAD_REL * float(wg.abs().max()) + 1e-12, (D, ps, err, float(wg.abs().max()))


@pytest.mark.parametrize("B,D", [(4500, 16), (5000, 4), (4100, 32)])
def test_tensor_core_path_ragged_sizes_against_fp64(B, D):
    """Tensor-core path at sizes that are multiples of neither the 128-row nor the 96-column tile, on the operand widths the
    65536 test does not reach (D = 16: two coefficient buffers, five ring stages; D = 4: dims padded to 8; D = 32: split rings,
    two producer sets), with more than 32 column tiles so the accumulator drain runs.  Same gates as at 65536."""
    from clear_vae_b200.latent import latent_block
    from tests.helpers import snn_fp64_chunked
    gen = torch.Generator().manual_seed(B + D)
    mu_c, mu_s = torch.randn(B, D, generator=gen), torch.randn(B, D, generator=gen)
    lab = torch.randint(0, 7, (B,), generator=gen)
    a = mu_c.to(DEV).requires_grad_(True)
    b = mu_s.to(DEV).requires_grad_(True)
    _, sc = latent_block([a, b], [None, None], [None, None], lab.to(DEV), snn=[1, 1], ps=[False, True], temperature=0.1, want_z=False)
    w = torch.zeros(8, device=DEV)
    w[2], w[3] = 100.0, 100.0
    torch.autograd.backward([sc], [w])
    torch.cuda.synchronize()
    for t, ps, idx in ((a, False, 2), (b, True, 3)):
        want, wg = snn_fp64_chunked(t.detach(), lab, 0.1, ps, device=DEV, chunk=2048)
        assert close(float(sc[idx]), want), (D, ps, float(sc[idx]), want)
        wg = wg * 100.0
        err = float((t.grad.double() - wg).abs().max())
        assert err <= GRAD_REL * float(wg.abs().max()) + 1e-12, (B, D, ps, err, float(wg.abs().max()))


def test_data_parallel_shards_alternating_with_global_batches_on_one_gpu(monkeypatch):
    """The data-parallel latent path of every rank, emulated on one GPU (the exchange is replaced by the known global tensors), run
    ALTERNATELY with single-process global-batch calls in the same process: row gradients x 1/world and the loss must equal the
    global run.  Regression test of the scratch-buffer bug found at 4 / 8 GPUs in round 2: the column-split backward's tickets /
    partials have a shape-dependent layout, and one cached buffer per device was corrupted by alternating shapes."""
    import clear_vae_b200.latent as L
    from clear_vae_b200.latent import DistSpec, latent_block
    world, Bl, D = 4, 96, 8
    Bg = world * Bl
    g = torch.Generator().manual_seed(44)
    mu_c, lv_c, mu_s, lv_s, e_c, e_s = ((torch.randn(Bg, D, generator=g) * s).to(DEV) for s in (1, .3, 1, .3, 1, 1))
    lab = torch.randint(0, 10, (Bg,), generator=g).to(DEV)
    w = torch.tensor([0.1, 0.1, 100.0, 0.0, 0, 0, 0, 0], device=DEV)
    table = {}

    def fake_gather(t, dist):
        return table[(tuple(t.shape[1:]), t.dtype)]

    monkeypatch.setattr(L, "_all_gather_rows", fake_gather)
    for redundant in (4096, 0):          # single-exchange path (all-rows statistics) and the two-exchange path
        monkeypatch.setattr(L, "REDUNDANT_ROWS_MAX", redundant)
        for rep in range(2):
            full = [t.clone().requires_grad_(True) for t in (mu_c, lv_c, mu_s, lv_s)]
            z, sc = latent_block([full[0], full[2]], [full[1], full[3]], [e_c, e_s], lab, snn=[1, 0], ps=[False, False], temperature=0.1)
            torch.autograd.backward([sc], [w])
            # what the exchanges would deliver: the global operands, labels, and (two-exchange path) the global row statistics
            ops = __import__("clear_vae_b200._ops", fromlist=["ops"]).ops()
            ws = L._workspace(lab.device, ops.latent_workspace_bytes(Bg, Bg, D, 2), ("test", Bg))
            _, _, st = ops.latent_fwd([mu_c, mu_s], [None, None], [None, None], [None, None], [None, None], lab, None, [1, 0], [0, 0], 0, 0, 0,
                                      0.1, True, False, ws)
            table.clear()
            table[((D,), torch.float32)] = mu_c
            table[((), torch.int64)] = lab
            table[((2,), torch.float32)] = st[0]
            table[((D + 2,), torch.float32)] = torch.cat([mu_c, lab.view(Bg, 1).view(torch.float32)], dim=1)   # the packed NCCL exchange
            for r in range(world):
                sl = slice(r * Bl, (r + 1) * Bl)
                loc = [t[sl].clone().requires_grad_(True) for t in (mu_c, lv_c, mu_s, lv_s)]
                dist = DistSpec(None, r, world, None)
                z2, sc2 = latent_block([loc[0], loc[2]], [loc[1], loc[3]], [e_c[sl].contiguous(), e_s[sl].contiguous()], lab[sl].contiguous(),
                                       snn=[1, 0], ps=[False, False], temperature=0.1, dist=dist)
                torch.autograd.backward([sc2], [w])
                assert abs(float(sc2[2]) - float(sc[2])) <= 2e-6 * abs(float(sc[2])), (redundant, rep, r)
                assert torch.allclose(z2, z[sl], rtol=0, atol=1e-6)
                # SNN gradients come back scaled by `world`; the KL part is a local mean over Bl rows
                got = loc[0].grad / world - 0.1 * loc[0].detach() / Bl / world
                want = full[0].grad[sl] - 0.1 * full[0].detach()[sl] / Bg
                assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max()), (redundant, rep, r)
