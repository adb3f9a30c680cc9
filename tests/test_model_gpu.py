"""GPU parity of the conv engine + trainers.

Chain of evidence:
  reference goldens == fp32 oracle autograd == engine emulator without rounding   (CPU tests)
  CUDA engine ~= engine emulator WITH bf16 rounding                               (here, tight on the shallow net)
  CUDA engine ~= fp32 reference goldens within the bf16 envelope                  (here, forward 1e-2)
"""
import numpy as np
import pytest
import torch

from tests.helpers import digest, emulated_step, load_step, sample_of, seeded_model, state_of

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF16_FWD = 1e-2   # north_star: bf16 convs within a stated 1e-2 (loss components / activations, relative)


@pytest.fixture(autouse=True)
def _fp32_reference_math():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


def rel(a, b):
    return abs(float(a) - float(b)) / (abs(float(b)) + 1e-12)


def l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def run_cuda_step(tag):
    g, meta, hyper = load_step(tag)
    m = seeded_model(meta, DEV)
    m.train()
    for k, v in m.state_dict().items():
        assert np.allclose(digest(v), g[f"init_digest/{k}"], rtol=1e-6, atol=1e-6), k
    st0 = state_of(m)
    X, label = torch.tensor(g["X"]).to(DEV), torch.tensor(g["label"]).to(DEV)
    eps = (torch.tensor(g["eps/0"]).to(DEV), torch.tensor(g["eps/1"]).to(DEV))
    ps = hyper.get("ps", False)
    snn = [1, 1] if meta["kind"] == "clear" else [1, 0]
    xhat, recon, z, sc, lp = m.fused_step_forward(X, label, temperature=hyper["temperature"], snn=snn, ps=[False, bool(ps)], eps=eps)
    slope = float(g["slope"])
    w = torch.zeros(8, device=DEV)
    w[0] = w[1] = slope
    w[2] = hyper["alpha"]
    if meta["kind"] == "clear":
        w[3] = hyper["alpha"] if ps else -hyper["alpha"]
    torch.autograd.backward([recon, sc], [torch.ones_like(recon), w])
    torch.cuda.synchronize()
    return g, meta, hyper, m, st0, dict(xhat=xhat, recon=recon, z=z, sc=sc, lp=lp), slope


@pytest.mark.parametrize("tag", ["clear_vae28_ps", "clear_vae28_nops", "clear_vae64_ps"])
def test_forward_within_bf16_envelope_of_reference(tag):
    g, meta, hyper, m, st0, out, slope = run_cuda_step(tag)
    sc = out["sc"]
    assert rel(out["recon"], g["recon"]) < BF16_FWD
    assert rel(sc[0], g["kl_c"]) < BF16_FWD and rel(sc[1], g["kl_s"]) < BF16_FWD
    assert rel(sc[2], g["c_loss"]) < 2 * BF16_FWD          # tau = 0.1 multiplies latent error by 10 before exp
    for k in ("mu_c", "logvar_c", "mu_s", "logvar_s"):
        ref = torch.tensor(g[f"latent/{k}"])
        assert l2(out["lp"][k], ref) < 2 * BF16_FWD, k
    if g["xhat"].shape == tuple(out["xhat"].shape):
        assert float((out["xhat"].cpu() - torch.tensor(g["xhat"])).abs().max()) < 2e-2
    # BatchNorm running statistics (momentum update from the batch stats)
    for k, b in m.named_buffers():
        ref = torch.tensor(g[f"buf_after_fwd/{k}"])
        if k.endswith("num_batches_tracked"):
            assert int(b) == int(ref)
        else:
            assert l2(b, ref) < BF16_FWD, k


def test_engine_matches_bf16_emulator_on_shallow_net():
    """VAE(28x28), B=32: rounding flips are rare enough that the CUDA path must reproduce the
    emulated bf16 arithmetic almost exactly — any indexing / BatchNorm-backward / mask bug shows here."""
    g, meta, hyper, m, st0, out, slope = run_cuda_step("clear_vae28_ps")
    em = emulated_step(st0, meta, hyper, g, True, slope, device=DEV)
    lat = torch.cat([out["lp"][k] for k in ("mu_c", "logvar_c", "mu_s", "logvar_s")], 1)
    assert l2(lat, em["lat"]) < 1e-3
    assert l2(out["xhat"], em["xhat"]) < 5e-3
    assert rel(out["recon"], em["recon"]) < 1e-4 and rel(out["sc"][2], em["c"]) < 1e-3
    # per-tensor gradient error: bf16 rounding flips (1 ulp = 4e-3) propagate through the latents; the heads of the
    # log-variances carry the smallest gradients and are the noisiest (3-5e-2 with either first-layer kernel,
    # tools/debug_emu.py) — a wrong tap / mask / BatchNorm coefficient gives O(1) errors
    errs = []
    for k, p in m.named_parameters():
        if p.grad is None:
            assert k.endswith(".bias") and k not in em["grads"], k  # BN-fed biases: exact zero gradient
            continue
        e = l2(p.grad, em["grads"][k])
        errs.append(e)
        assert e < 6e-2, (k, e)
    assert float(np.median(errs)) < 1.5e-2, errs


def test_gradients_track_fp32_reference_within_bf16_noise():
    """vs the fp32 goldens the bf16 pipeline differs by rounding + ReLU-mask flips:
    a few % in L2 per tensor on the 28x28 net; the direction must agree."""
    g, meta, hyper, m, st0, out, slope = run_cuda_step("clear_vae28_ps")
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        ref = g[f"grad_sample/{k}"]
        got = sample_of(p.grad, 256)
        cos = float(np.dot(got, ref) / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30))
        assert cos > 0.97, (k, cos)


@pytest.mark.parametrize("tag", ["clear_vae28_ps", "tc_vae28", "mim_club_vae28", "mim_l1out_vae28", "tc_vae64"])
def test_trainer_step_matches_reference_logs(tag):
    from clear_vae_b200.utils.trainer_utils import get_clearmimvae_trainer, get_cleartcvae_trainer, get_clearvae_trainer
    g, meta, hyper = load_step(tag)
    torch.manual_seed(meta["seed"])
    if meta["kind"] == "clear":
        tr = get_clearvae_trainer(hyper["beta"], hyper["ps"], hyper["lr"], meta["zdim"], hyper["alpha"], hyper["temperature"],
                                  DEV, meta["arch"], meta["cin"])
        aux = None
    elif meta["kind"] == "tc":
        tr = get_cleartcvae_trainer(hyper["beta"], hyper["lambda"], hyper["lr"], hyper["aux_lr"], meta["zdim"], hyper["alpha"],
                                    hyper["temperature"], DEV, meta["arch"], meta["cin"])
        aux = tr.factor_cls
    else:
        tr = get_clearmimvae_trainer(hyper["beta"], meta["est"], hyper["lambda"], hyper["lr"], hyper["aux_lr"], meta["zdim"],
                                     hyper["alpha"], hyper["temperature"], DEV, meta["arch"], meta["cin"])
        aux = tr.mi_estimator
    if aux is not None:  # same construction order as the reference => same seeded weights
        for k, v in aux.state_dict().items():
            assert np.allclose(v.cpu().numpy(), g[f"aux_init/{k}"], rtol=1e-6, atol=1e-7), k
    X, label = torch.tensor(g["X"]).to(DEV), torch.tensor(g["label"]).to(DEV)
    n_eps = sum(1 for k in g.files if k.startswith("eps/"))
    eps = [torch.tensor(g[f"eps/{i}"]).to(DEV) for i in range(n_eps)]
    tr.model.train()
    before = {k: v.detach().clone() for k, v in tr.model.named_parameters()}
    if meta["kind"] == "clear":
        recon, sc = tr.train_step(X, label, eps=(eps[0], eps[1]))
    elif meta["kind"] == "tc":
        recon, sc, mi, fl = tr.train_step(X, label, eps=(eps[0], eps[1]), eps2=(eps[2], eps[3]))
        assert rel(mi, g["mi_loss"]) < 5e-2 and rel(fl, g["train_logs1"][0]) < BF16_FWD
    else:
        inner = [(eps[2 + 2 * j], eps[3 + 2 * j]) for j in range(5)]
        recon, sc, mi, learn = tr.train_step(X, label, eps=(eps[0], eps[1]), inner_eps=inner, perm=torch.tensor(g["perm"]))
        assert abs(float(mi) - float(g["mi_loss"])) < 5e-2 * abs(float(g["mi_loss"])) + 2e-2
        assert np.allclose(learn.cpu().numpy(), g["train_logs2"], rtol=3e-2, atol=1e-2)
    assert rel(recon, g["recon"]) < BF16_FWD and rel(sc[2], g["c_loss"]) < 2 * BF16_FWD
    # Adam's first step moves every weight by ~lr * sign(grad): the update direction must agree with the
    # reference wherever the reference gradient is not rounding noise
    agree, total = 0, 0
    for k, p in tr.model.named_parameters():
        if p.dim() < 2:
            continue
        ref_after, ref_grad = g[f"after_sample/{k}"], g[f"grad_sample/{k}"]
        mine = sample_of(p - before[k], 256)
        refd = ref_after - sample_of(before[k], 256)
        big = np.abs(ref_grad) > 0.05 * np.abs(ref_grad).max()
        agree += int((np.sign(mine[big]) == np.sign(refd[big])).sum())
        total += int(big.sum())
    assert total > 100 and agree / total > 0.97, (agree, total)


@pytest.mark.parametrize("arch", ["VAE", "VAE64"])
def test_parallel_statistics_only_decoder_passes_equal_sequential_passes(arch):
    """CLEAR-MIM's five discarded inner forwards (trainer.py:874-888) only move the decoder's BatchNorm running
    statistics; running them as parallel branches + the closed-form momentum fold must give what five sequential
    train-mode passes give (same kernels, same batch statistics; only the fp32 order of the momentum update differs)."""
    from clear_vae_b200.models.vae import VAE, VAE64
    torch.manual_seed(5)
    cls, zdim = (VAE, 16) if arch == "VAE" else (VAE64, 64)
    a = cls(zdim, 3).to(DEV)
    b = cls(zdim, 3).to(DEV)
    b.load_state_dict(a.state_dict())
    a.train(), b.train()
    g = torch.Generator().manual_seed(3)
    zs = [torch.randn(256, zdim, generator=g).to(DEV) * (1 + j) for j in range(5)]
    with torch.no_grad():
        a._eng().parallel_stats = False
        a.decode_stats_many(zs)
        b.decode_stats_many(zs)
        b.decode_stats_many(zs[:3])     # a different branch count reuses nothing stale
        a.decode_stats_many(zs[:3])
    torch.cuda.synchronize()
    sa, sb = a.state_dict(), b.state_dict()
    moved = 0
    for k in sa:
        if "running" in k:
            assert torch.allclose(sa[k], sb[k], rtol=2e-6, atol=1e-7), (k, float((sa[k] - sb[k]).abs().max()))
            moved += int(k.startswith("decoder") and not torch.equal(sa[k], torch.zeros_like(sa[k])) and not torch.equal(sa[k], torch.ones_like(sa[k])))
        if "num_batches_tracked" in k:
            assert int(sa[k]) == int(sb[k]) == (8 if k.startswith("decoder") else 0), k
    assert moved >= 8


def test_reparam_multi_matches_the_latent_kernel_bitwise():
    from clear_vae_b200 import _ops
    from clear_vae_b200.latent import latent_block
    g = torch.Generator().manual_seed(11)
    B, D = 1000, 8
    mu = [torch.randn(B, D, generator=g).to(DEV) for _ in range(2)]
    lv = [(torch.randn(B, D, generator=g) * 0.5).to(DEV) for _ in range(2)]
    eps = [torch.randn(B, D, generator=g).to(DEV) for _ in range(10)]
    zs = _ops.ops().reparam_multi(mu, lv, eps)
    dummy = torch.zeros(B, dtype=torch.int64, device=DEV)
    for j in range(5):
        want, _ = latent_block(mu, lv, eps[2 * j:2 * j + 2], dummy, snn=[0, 0], ps=[0, 0])
        assert torch.equal(zs[j], want)


def test_device_prefetcher_hands_out_every_batch_intact_under_a_slow_consumer():
    """`VAETrainer.prefetch` (the input path of `fit()`): three rotating device slots, one batch of look-ahead.  A consumer
    that is far behind on the GPU (a 1 ms spin kernel in front of every use) must still read every batch unmodified —
    a slot may only be refilled after the work enqueued on it has executed."""
    from clear_vae_b200.trainer import DevicePrefetcher
    gen = torch.Generator().manual_seed(4)
    sizes = [64] * 9 + [17]          # ragged last batch, like a DataLoader without drop_last
    src = [(torch.rand(b, 3, 28, 28, generator=gen).pin_memory(), torch.randint(0, 10, (b, 1), generator=gen).pin_memory()) for b in sizes]
    kept = []
    for X, y in DevicePrefetcher(src, torch.device(DEV)):
        assert X.is_cuda and y.dtype == torch.int64 and y.dim() == 1
        torch.cuda._sleep(2_000_000)
        kept.append((X.clone(), y.clone()))       # executes after the spin, long after the next batches were staged
    torch.cuda.synchronize()
    assert len(kept) == len(src)
    for (X, y), (xs, ys) in zip(kept, src):
        assert torch.equal(X.cpu(), xs) and torch.equal(y.cpu(), ys.reshape(-1))


def test_graph_replayed_fit_over_a_dataloader_and_step_input_staging():
    """`fit()` = DataLoader -> prefetcher -> CUDA-graph replay of the step, with the host-written step inputs (annealer
    weight) staged through alternating pinned slots and uploaded outside the graph.  Nothing synchronises per step, so
    the host runs ahead of the GPU: every step must still see ITS OWN annealer weight (read back here from the device
    tensor the captured kernels use, in stream order), and an epoch over a DataLoader must train (finite, moved) weights."""
    from torch.utils.data import DataLoader, TensorDataset
    from clear_vae_b200.utils.trainer_utils import get_clearvae_trainer
    gen = torch.Generator().manual_seed(8)
    X = torch.rand(64 * 10, 3, 28, 28, generator=gen)
    y = torch.randint(0, 10, (64 * 10,), generator=gen)
    loader = DataLoader(TensorDataset(X, y), batch_size=64, shuffle=False, pin_memory=True)
    torch.manual_seed(21)
    tr = get_clearvae_trainer(1 / 8, True, 5e-4, 16, 1e2, 0.1, torch.device(DEV), "VAE", 3)
    tr.annealer.loc, tr.annealer.scale = 5, 2      # the KL weight changes every step: a stale slot would be visible
    tr.use_cuda_graph = True
    tr.verbose_period = 10 ** 9                    # quiet epochs: no per-step read-back, the host runs ahead
    tr.model.train()
    init = {k: v.clone() for k, v in tr.model.state_dict().items()}
    seen, want = [], []
    for xb, yb in tr.prefetch(loader):
        want.append(tr.annealer.slope())
        torch.cuda._sleep(3_000_000)               # keep the GPU ~1.5 ms behind the host
        tr.train_step(xb, yb)
        seen.append(tr._w_dev.clone())             # stream-ordered: the weights this step's kernels read
    torch.cuda.synchronize()
    got = torch.stack(seen).cpu()
    assert torch.allclose(got[:, 0], torch.tensor(want, dtype=torch.float32), rtol=1e-6, atol=0)
    assert torch.equal(got[:, 0], got[:, 1]) and bool((got[:, 2] == 100.0).all())
    assert len(set(round(w, 9) for w in want)) == 10
    tr.fit(2, loader)                              # epoch 0 is verbose (per-step read-back), epoch 1 quiet
    torch.cuda.synchronize()
    assert tr.annealer.current_step == 30
    for k, v in tr.model.state_dict().items():
        assert bool(torch.isfinite(v.float()).all()), k
    assert l2(tr.model.state_dict()["encoder.0.weight"], init["encoder.0.weight"]) > 1e-3


def test_first_graph_step_is_exactly_one_update():
    """Capturing the step needs eager warm-up runs; they must leave no trace: after the first graphed `train_step` the
    model has seen ONE update (Adam step counter, BatchNorm `num_batches_tracked`), the CPU generator has drawn what one
    reference iteration draws, and a second trainer stepping eagerly from the same state lands on the same weights."""
    from clear_vae_b200.utils.trainer_utils import get_clearmimvae_trainer
    gen = torch.Generator().manual_seed(12)
    X = torch.rand(128, 3, 28, 28, generator=gen).to(DEV)
    y = torch.randint(0, 10, (128,), generator=gen).to(DEV)

    def make():
        torch.manual_seed(33)
        tr = get_clearmimvae_trainer(1 / 8, "CLUBSample", 3, 5e-4, 2e-3, 16, 1e2, 0.1, torch.device(DEV), "VAE", 3)
        tr.model.train()
        return tr

    a = make()
    a.use_cuda_graph = True
    torch.manual_seed(5)
    a.train_step(X, y)
    torch.cuda.synchronize()
    cpu_state_after_graph = torch.get_rng_state()
    b = make()
    torch.manual_seed(5)
    b.train_step(X, y)
    torch.cuda.synchronize()
    assert torch.equal(cpu_state_after_graph, torch.get_rng_state())        # one randperm each (CLUB-S), nothing else
    for tr in (a, b):
        assert int(tr.model.encoder[1].num_batches_tracked) == 6            # 1 full forward + 5 inner forwards (trainer.py:874-888)
        assert int(tr.model.decoder[1].num_batches_tracked) == 6
        st = tr.optimizer.state[next(iter(tr.model.parameters()))]
        assert float(st["step"]) == 1.0
        est_p = next(iter(tr.mi_estimator.parameters()))
        assert float(tr.mi_estimator_optimizer.state[est_p]["step"]) == 5.0
    # one Adam step moves every weight by ~lr; the graphed path must have moved them by one step, not by three (the two
    # warm-up runs).  The two paths draw different reparameterisation noise (graph-safe Philox offsets), so the updates
    # agree in size and broadly in direction, not element for element.
    sa, sb = a.model.state_dict(), b.model.state_dict()
    init = make().model.state_dict()
    for k in ("encoder.0.weight", "decoder.0.weight", "mu_c.weight"):
        da, db = (sa[k] - init[k].to(DEV)).double().flatten(), (sb[k] - init[k].to(DEV)).double().flatten()
        assert 0.8 < float(da.norm() / db.norm()) < 1.25, (k, float(da.norm()), float(db.norm()))
        assert float(da @ db / (da.norm() * db.norm())) > 0.3, k


@pytest.mark.parametrize("arch", ["VAE", "VAE64"])
def test_eval_mode_forward_matches_the_torch_modules_and_evaluate_runs(arch):
    """f-3: `evaluate()` runs the engine in eval mode (running-statistic BatchNorm).  The parameter containers are
    ordinary torch modules, so calling them directly (cuDNN / cuBLAS, fp32) is an independent reference for the eval-mode
    forward: latent parameters and reconstruction must agree within the bf16 envelope."""
    from torch.utils.data import DataLoader, TensorDataset
    from clear_vae_b200.models.vae import VAE, VAE64
    from clear_vae_b200.utils.trainer_utils import get_clearvae_trainer
    torch.manual_seed(9)
    hw, zdim = (28, 16) if arch == "VAE" else (64, 64)
    tr = get_clearvae_trainer(1 / 8, True, 5e-4, zdim, 1e2, 0.1, torch.device(DEV), arch, 3)
    vae = tr.model
    gen = torch.Generator().manual_seed(2)
    X = torch.rand(96, 3, hw, hw, generator=gen)
    y = torch.randint(0, 4, (96,), generator=gen)
    vae.train()
    for i in range(3):                      # move the running statistics away from (0, 1)
        tr.train_step(X[:64].to(DEV), y[:64].to(DEV))
    vae.eval()
    with torch.no_grad():
        xd = X.to(DEV)
        mu_c, lv_c, mu_s, lv_s = vae.encode(xd)
        h = vae.encoder(xd)
        for got, head in ((mu_c, vae.mu_c), (lv_c, vae.logvar_c), (mu_s, vae.mu_s), (lv_s, vae.logvar_s)):
            want = head(h)
            assert float((got - want).abs().max()) <= BF16_FWD * float(want.abs().max()) + 2e-3
        z = torch.cat([mu_c, mu_s], 1)
        xhat = vae.decode(z)
        want = vae.decoder(z)
        assert float((xhat - want).abs().max()) <= 2e-2     # sigmoid output in [0, 1]
    loader = DataLoader(TensorDataset(X, y), batch_size=32)
    mig, mse = tr.evaluate(loader, False, 0)
    assert np.isfinite(mig) and np.isfinite(mse) and mse > 0
    with torch.no_grad():
        ref = torch.stack([((vae.decoder(torch.cat([vae.mu_c(vae.encoder(b)), vae.mu_s(vae.encoder(b))], 1)) - b) ** 2).flatten(1).sum(1).mean()
                           for b in xd.split(32)]).mean()
    # evaluate() reconstructs from sampled z (eval-mode forward still draws noise, vae.py:81-102), the line above from the
    # means: same order of magnitude, not equal
    assert 0.5 < mse / float(ref) < 2.0
    assert not vae.training
