#!/usr/bin/env python
"""bench.py — CLEAR-VAE training-step throughput on B200 (metric of BASELINE.json).

    python bench.py --gpus N --steps K --warmup W [--config NAME] [--impl reference]

A "step" = one iteration of the reference trainer's loop body (`_train`, trainer.py:446-492 /
646-709 / 841-897) on one synthetic batch per GPU: forward, backward, optimiser update(s),
including the 5 estimator iterations of CLEAR-MIM / the discriminator update of CLEAR-TC.
Default workload = BASELINE.json configs[1]: CLEAR-MIM-VAE (CLUB-S), Styled-MNIST-shaped
3x28x28, batch 1024 per GPU (weak scaling).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: kind, arch, total_z, in_ch, image, per-GPU batch, n_classes, hyper-parameters (reference scripts)
    "clear28": dict(kind="clear", arch="VAE", z=16, cin=3, hw=28, B=128, ncls=10,
                    hp=dict(beta=1 / 8, ps=True, lr=5e-4, alpha=1e2, temperature=0.1)),
    "mim_club": dict(kind="mim", est="CLUBSample", arch="VAE", z=16, cin=3, hw=28, B=1024, ncls=10,
                     hp=dict(beta=1 / 8, lr=5e-4, alpha=1e2, temperature=0.1, la=3, aux_lr=2e-3)),
    "mim_l1out": dict(kind="mim", est="L1OutUB", arch="VAE", z=16, cin=3, hw=28, B=1024, ncls=10,
                      hp=dict(beta=1 / 8, lr=5e-4, alpha=1e2, temperature=0.1, la=3, aux_lr=2e-3)),
    "tc64": dict(kind="tc", arch="VAE64", z=64, cin=3, hw=64, B=512, ncls=4,
                 hp=dict(beta=1 / 32, lr=3e-5, alpha=1e2, temperature=0.1, la=1, aux_lr=1e-4)),
    "clear64": dict(kind="clear", arch="VAE64", z=64, cin=3, hw=64, B=128, ncls=7,
                    hp=dict(beta=1 / 32, ps=True, lr=3e-5, alpha=1e2, temperature=0.1)),
}
WORKLOAD_NAMES = {
    "clear28": "CLEAR-VAE, synthetic Styled-MNIST-shaped 3x28x28, batch 128/GPU (BASELINE configs[0])",
    "mim_club": "CLEAR-MIM-VAE (CLUB-S), synthetic Styled-MNIST-shaped 3x28x28, batch 1024/GPU (BASELINE configs[1])",
    "mim_l1out": "CLEAR-MIM-VAE (L1OutUB), synthetic Styled-MNIST-shaped 3x28x28, batch 1024/GPU (BASELINE configs[1])",
    "tc64": "CLEAR-TC-VAE, synthetic CelebA-shaped 3x64x64, batch 512/GPU (BASELINE configs[2])",
    "clear64": "CLEAR-VAE, synthetic PACS-shaped 3x64x64, batch 128/GPU (BASELINE configs[3])",
}
N_POOL = 16  # distinct input batches rotated through the timed region
SHAPES = {"clear28": "configs[0]", "mim_club": "configs[1]", "mim_l1out": "configs[1]", "tc64": "configs[2]", "clear64": "configs[3]"}


def bench_config(name):
    """`config` of the JSON line — identical on the b200 and the reference arm (same workload, same synthetic batches)."""
    c = CONFIGS[name]
    return dict(workload=WORKLOAD_NAMES[name], baseline_config=SHAPES[name], trainer=c["kind"], estimator=c.get("est"), model=c["arch"],
                total_z_dim=c["z"], image=[c["cin"], c["hw"], c["hw"]], per_gpu_batch=c["B"], n_classes=c["ncls"],
                hyperparameters=c["hp"], seed=101,
                l2=f"inputs rotate through {N_POOL} distinct batches ({N_POOL * c['B'] * c['cin'] * c['hw'] ** 2 * 4 / 1e6:.0f} MB"
                   f"{' > 126 MB L2' if N_POOL * c['B'] * c['cin'] * c['hw'] ** 2 * 4 > 126e6 else ''}); activations and gradients are rewritten every step")


def input_pool(cfg, rank=0, pin=False):
    import torch
    g = torch.Generator().manual_seed(101 + rank)
    pool = [(torch.rand(cfg["B"], cfg["cin"], cfg["hw"], cfg["hw"], generator=g), torch.randint(0, cfg["ncls"], (cfg["B"],), generator=g))
            for _ in range(N_POOL)]
    return [(x.pin_memory(), y.pin_memory()) for x, y in pool] if pin else pool


def ncu_traffic(kernel, config):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/r2_ncu_traffic.json, written by
    tools/ncu_summary.py from the raw CSV: mean of dram__bytes_read.sum + dram__bytes_write.sum over the captured launches)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
        return d.get(config, {}).get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops_sustained"], tf_burst=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, src="fallback")


def layer_flops(arch, cin, zdim, B):
    D = zdim // 2
    if arch == "VAE":
        enc = [(cin, 32, 3, 14), (32, 64, 3, 7), (64, 128, 3, 4)]          # (cin, cout, k, hout)
        dec = [(128, 64, 3, 4), (64, 32, 3, 7), (32, cin, 3, 14)]          # (cin, cout, k, hin)
    else:
        ch = [cin, 32, 64, 128, 256, 512]
        enc = [(ch[i], ch[i + 1], 4, 64 >> (i + 1)) for i in range(5)]
        rc = ch[::-1]
        dec = [(rc[i], rc[i + 1], 4, 2 << i) for i in range(5)]
    f_enc = [2.0 * B * h * h * co * ci * k * k for (ci, co, k, h) in enc]
    f_dec = [2.0 * B * h * h * ci * co * k * k for (ci, co, k, h) in dec]
    f_heads, f_fc = 2.0 * B * 2048 * 4 * D, 2.0 * B * 2 * D * 2048
    fwd = sum(f_enc) + sum(f_dec) + f_heads + f_fc
    # what actually runs on the tensor-core GEMM kernel (`conv_gemm`): the first conv, the last conv-transpose (forward and
    # data gradient) and the fc forward are CUDA-core kernels (conv_direct_* / fc_fwd)
    tc_enc_fwd = sum(f_enc[1:]) + f_heads
    tc_dec_fwd = sum(f_dec[:-1])
    tc_dgrad = sum(f_enc[1:]) + f_heads + f_fc + sum(f_dec[:-1])
    return dict(fwd=fwd, dgrad=fwd - f_enc[0], wgrad=fwd, tc_enc_fwd=tc_enc_fwd, tc_dec_fwd=tc_dec_fwd, tc_dgrad=tc_dgrad,
                tc_wgrad=fwd - f_enc[0] - f_dec[-1])


# --------------------------------------------------------------------------------------
# reference arm: the reference's algorithm on the host CPU (oracle port), same config
# --------------------------------------------------------------------------------------
def oracle_stepper(cfg, seed=101):
    import torch
    from oracle import model_oracle as mo
    hp = cfg["hp"]
    B = cfg["B"]
    st = mo.init_state(cfg["arch"], cfg["z"], cfg["cin"], seed=seed)
    hyper = dict(temperature=hp["temperature"], alpha=hp["alpha"], beta=hp["beta"], loc=0, scale=1, ps=hp.get("ps"))
    aux, aux_lr = None, None
    if cfg["kind"] == "tc":
        aux, aux_lr = mo.init_factor_state(cfg["z"], seed), hp["aux_lr"]
        hyper["lambda"] = hp["la"]
    elif cfg["kind"] == "mim":
        aux, aux_lr = mo.init_estimator_state(cfg["z"] // 2, cfg["z"] // 2, cfg["z"], seed), hp["aux_lr"]
        hyper["lambda"] = hp["la"]
    so = mo.StepOracle(cfg["kind"], st, cfg["arch"], cfg["cin"], hyper, hp["lr"], aux=aux, aux_lr=aux_lr,
                       estimator=cfg.get("est", "CLUBSample"))
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(B, cfg["cin"], cfg["hw"], cfg["hw"], generator=g)
    label = torch.randint(0, cfg["ncls"], (B,), generator=g)
    return lambda: so.step(X, label)


def time_cpu(cfg, steps, warmup, pool=None):
    """The reference's own CPU implementation of the step on the host cores: the UNMODIFIED reference staged under oracle/_ref
    (`kind` "reference") when it travelled with the snapshot, else the oracle port.  Returns (samples/s, ms/step, threads, kind, note)."""
    import torch
    from oracle import ref_runner as rr
    if rr.available() and cfg.get("est") != "L1OutUB":
        pool = pool or input_pool(cfg)
        sps, ms, _ = rr.time_train(cfg, "cpu", pool, steps, warmup)
        return sps, ms, torch.get_num_threads(), "reference", ("unmodified reference `get_*_trainer(..., device='cpu')._train(loader, False, 1, ...)` "
                                                               "(oracle/_ref, staged by oracle/stage_reference.py)")
    if cfg.get("est") == "L1OutUB":
        note = "oracle port with the closed-form O(B*D) L1OutUB (the reference's own forward is CUDA-only: mi_estimator.py:185)"
    else:
        note = "oracle port (oracle/model_oracle.py::StepOracle): reference not staged"
    step = oracle_stepper(cfg)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return cfg["B"] / dt, dt * 1e3, torch.get_num_threads(), "port", note


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    steps, warm = args.steps, max(3, args.warmup)
    try:    # torchrun exports OMP_NUM_THREADS=1; the reference arm is meant to use every host core it can
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    sps, ms, cores, kind, note = time_cpu(cfg, steps, warm)
    sample = f"{steps} full steps of batch {cfg['B']} after {warm} warm-up steps; {note}"
    line = dict(impl="reference", metric="train samples/sec", value=sps, unit="samples/s", n_gpus=args.gpus, steps=steps,
                warmup=warm, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=bench_config(args.config), impl_config=dict(device="host CPU", threads=cores),
                cpu_baseline=dict(value=sps, unit="samples/s", cores=cores, kind=kind, sample=sample),
                e2e=dict(value=sps, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def eager_gpu_baseline(cfg, dev, pool, steps, warmup):
    """Second same-box comparator: the UNMODIFIED reference in eager PyTorch on the B200 (cuDNN / cuBLAS / ATen),
    host batches through its own DataLoader -> `.to(device)` path, per-step `float(loss)` synchronisation as written."""
    import torch
    from oracle import ref_runner as rr
    if not rr.available():
        return dict(unavailable="reference not staged (oracle/_ref missing)")
    try:
        sps, ms, _ = rr.time_train(cfg, str(dev), pool, steps, warmup)
    except Exception as e:
        return dict(unavailable=repr(e)[:300])
    return dict(value=sps, unit="samples/s", ms_per_step=ms, steps=steps, warmup=warmup, kind="reference",
                cudnn_allow_tf32=bool(torch.backends.cudnn.allow_tf32), matmul_allow_tf32=bool(torch.backends.cuda.matmul.allow_tf32),
                path="unmodified reference get_*_trainer(..., device='cuda:0')._train over a pinned DataLoader (H2D + per-step float() inside)")


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], None, set()
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx = float(c[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def build_trainer(cfg, device):
    import torch
    from clear_vae_b200.utils.trainer_utils import get_clearmimvae_trainer, get_cleartcvae_trainer, get_clearvae_trainer
    hp = cfg["hp"]
    torch.manual_seed(101)  # the reference scripts' default seed (run_mig_expr_mnist.py:35)
    if cfg["kind"] == "clear":
        return get_clearvae_trainer(hp["beta"], hp["ps"], hp["lr"], cfg["z"], hp["alpha"], hp["temperature"], device, cfg["arch"], cfg["cin"])
    if cfg["kind"] == "tc":
        return get_cleartcvae_trainer(hp["beta"], hp["la"], hp["lr"], hp["aux_lr"], cfg["z"], hp["alpha"], hp["temperature"], device,
                                      cfg["arch"], cfg["cin"])
    return get_clearmimvae_trainer(hp["beta"], cfg["est"], hp["la"], hp["lr"], hp["aux_lr"], cfg["z"], hp["alpha"], hp["temperature"],
                                   device, cfg["arch"], cfg["cin"])


def latent_roofline(dev, pk):
    """Fused latent-loss kernel pair at the top of BASELINE configs[4] (65536 all-gathered latents, D = 8 and D = 32):
    pairs/s against the measured MUFU ex2 rate — the pipe that bounds it (1 exp per (row, column) pair)."""
    import torch
    from clear_vae_b200.latent import latent_block
    probe = os.path.join(ROOT, "tools", "mufu_probe")
    peak, src = 4594.0, "recorded (tools/mufu_probe on this pool: 148 SMs x 16 ex2/clk)"
    try:
        out = subprocess.run([probe], capture_output=True, text=True, timeout=60).stdout
        vals = [json.loads(l)["gexp_per_s"] for l in out.splitlines() if '"ex2"' in l]
        if vals:
            peak, src = max(vals), "measured now (tools/mufu_probe)"
    except Exception:
        pass

    def point(B, D):
        g = torch.Generator().manual_seed(0)
        mu_c = torch.randn(B, D, generator=g).to(dev).requires_grad_(True)
        mu_s = torch.randn(B, D, generator=g).to(dev).requires_grad_(True)
        lv = (torch.randn(B, D, generator=g) * .3).to(dev).requires_grad_(True)
        eps = torch.randn(B, D, generator=g).to(dev)
        lab = torch.randint(0, 10, (B,), generator=g).to(dev)
        w = torch.tensor([0.1, 0.1, 100.0, 100.0, 0, 0, 0, 0], device=dev)

        def run(n):
            tf = tb = 0.0
            for _ in range(n):
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
                z, sc = latent_block([mu_c, mu_s], [lv, lv], [eps, eps], lab, snn=[1, 1], ps=[False, True], temperature=0.1)
                e1.record()
                torch.autograd.backward([sc], [w])
                e2.record()
                torch.cuda.synchronize()
                tf += e0.elapsed_time(e1)
                tb += e1.elapsed_time(e2)
                mu_c.grad = mu_s.grad = lv.grad = None
            return tf / n, tb / n

        run(3)
        tf, tb = run(5)
        pairs = 2.0 * B * B  # two terms (content, style)
        f, b = pairs / tf * 1e-6, pairs / tb * 1e-6
        return dict(forward=dict(ms=tf, achieved=f, frac=f / peak), backward=dict(ms=tb, achieved=b, frac=b / peak))

    d8, d32 = point(65536, 8), point(65536, 32)
    return dict(workload="contrastive + anti-contrastive terms, 65536 x 65536 pairs each, fp32 (3xTF32 on tcgen05); headline D=8",
                bound="mufu_ex2", unit="Gpair/s", peak=peak, peak_source=src, forward=d8["forward"], backward=d8["backward"],
                d32=d32, algorithmic="1 ex2 + 2*D (fwd) / 6*D (bwd) flop per pair; HBM bytes O(B*D), negligible")


def other_configs(args, K, W, budget_s=420.0):
    """One summary per remaining BASELINE config (configs[0], [2], [3]) measured by a sub-run of this script on the same GPU:
    value / e2e / dominant-kernel roofline / CPU and eager-B200 reference comparators, same timing rules as the headline."""
    out, t0 = [], time.time()
    for name in ("clear28", "tc64", "clear64"):
        if name == args.config:
            continue
        if time.time() - t0 > budget_s:
            out.append(dict(config=name, skipped="time budget of the default run"))
            continue
        cmd = [sys.executable, os.path.abspath(__file__), "--config", name, "--steps", str(min(K, 20)), "--warmup", str(W), "--sub",
               "--no-latent", "--precision", args.precision]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if not lines:
                out.append(dict(config=name, error=f"sub-run exited {r.returncode}: " + r.stderr.strip().splitlines()[-1][:300] if r.stderr.strip() else "no output"))
                continue
            d = json.loads(lines[-1])
            keep = ("value", "unit", "ms_per_step", "steps", "warmup", "dtype", "config", "e2e", "gpu_launches_per_step", "roofline",
                    "cpu_baseline", "eager_gpu_baseline", "parity", "clocks")
            out.append({k: d.get(k) for k in keep})
        except Exception as e:
            out.append(dict(config=name, error=repr(e)[:300]))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="mim_club", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-latent", action="store_true", help="skip the latent-loss roofline point")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the eager-PyTorch-on-B200 reference comparator")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config summary lines (configs[0], [2], [3]) of the `configs` array")
    ap.add_argument("--no-parity", action="store_true", help="skip the first-step parity check against the CPU step oracle")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32x3"],
                    help="conv / linear numerics: bf16 operands (default) or the fp32-grade bf16 x 3 split")
    ap.add_argument("--sub", action="store_true", help="(internal) a per-config sub-run of the `configs` array")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as td
    from clear_vae_b200 import _ops
    from clear_vae_b200.latent import DistSpec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep NCCL's banner / debug lines off stdout: one JSON line only
        td.init_process_group("nccl", device_id=dev)
    _ops.load()
    cfg = CONFIGS[args.config]
    B, K, W = cfg["B"], args.steps, max(3, args.warmup)
    tr = build_trainer(cfg, dev)
    tr.model.conv_precision = args.precision
    sampler = ClockSampler(local)     # started before warm-up so the record never comes back empty; covers every timed region
    if rank == 0:                     # one sampler per job: N nvidia-smi loops polling the driver at 50 Hz perturb N ranks' launches
        sampler.start()
    if world > 1:
        from clear_vae_b200.peer import PeerComm
        peer = PeerComm.create(td.group.WORLD, rank, world, dev)   # None -> NCCL collectives (e.g. IPC mapping unavailable)
        tr.dist = DistSpec(td.group.WORLD, rank, world, peer)
        for p in list(tr.model.parameters()):  # identical replicas
            td.broadcast(p.data, 0)
    tr.model.train()

    pool_h = input_pool(cfg, rank, pin=True)
    pool_d = [(x.to(dev), y.to(dev)) for x, y in pool_h]

    # ---- parity gate on the very first step: the CUDA step and the CPU step oracle on the same batch, noise and permutation
    parity = None
    if world == 1 and not args.no_parity:
        from oracle import parity as op
        res = op.compare_step(tr, cfg, pool_h[0][0], pool_h[0][1])
        tol = 1e-2 if args.precision == "bf16" else 1e-4
        checked = {k: v[2] for k, v in res.items() if k in ("recon", "kl_c", "kl_s", "c_loss", "s_loss")}
        lat_tol = 2 * tol   # latent parameter TENSORS (relative L2): see tests/test_parity_sizes_gpu.py (measured 0.75e-2 VAE / 1.3e-2 VAE64 in bf16)
        lat = {k: v[2] for k, v in res.items() if k.startswith("latent/")}
        gerr = sorted(v[2] for k, v in res.items() if k.startswith("grad/"))
        parity = dict(oracle="oracle/model_oracle.py::StepOracle (CPU fp32), same batch / eps / permutation", tolerance=tol,
                      rel_err=checked, latent_tolerance=lat_tol, latent_rel_l2=lat, mi_loss=res.get("mi_loss", (None, None, None))[:2],
                      grad_rel_l2=dict(median=gerr[len(gerr) // 2], max=gerr[-1]) if gerr else None,
                      ok=all(v < tol for v in checked.values()) and all(v < lat_tol for v in lat.values()))
        if not parity["ok"]:
            print(json.dumps(dict(error="first-step parity check failed", parity=parity)), file=sys.stderr, flush=True)
            sys.exit(3)

    def step_dev(i):
        x, y = pool_d[i % N_POOL]
        return tr.train_step(x, y)

    def step_e2e(i):
        xh, yh = pool_h[i % N_POOL]
        x, y = xh.to(dev, non_blocking=True), yh.to(dev, non_blocking=True)
        out = tr.train_step(x, y)
        vals = torch.cat([out[0].detach().view(1), out[1].detach().view(-1)]).to("cpu")  # the step's scalars, one read-back
        return vals

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also fills the packed-weight caches / cuBLAS handles)
    for i in range(W):
        step_dev(i)
    barrier()
    graphed = False
    if not args.no_graph:
        try:
            tr.use_cuda_graph = True
            for i in range(3):      # first call captures the device body of the step, the others replay it
                step_dev(i)
                if os.environ.get("CLEARVAE_DEBUG"):
                    torch.cuda.synchronize()
                    print(f"[bench] rank {rank} graph step {i} done", file=sys.stderr, flush=True)
            barrier()
            graphed = True
        except Exception as e:      # capture is an optimisation, never a requirement
            tr.use_cuda_graph = False
            tr._graph = None
            torch.cuda.synchronize()
            print(f"[bench] CUDA-graph capture unavailable, running eagerly: {e!r}", file=sys.stderr, flush=True)

    def dbg(msg):
        if os.environ.get("CLEARVAE_DEBUG"):
            torch.cuda.synchronize()
            print(f"[bench] rank {rank}: {msg}", file=sys.stderr, flush=True)

    dbg("graphs ready")
    # ---- one profiling pass: which of OUR kernels dominates the step?
    meter = _ops.meter
    meter.reset()
    tr.use_cuda_graph = False   # per-kernel event timing needs the eager path (same kernels, same order)
    # ... and one stream: with the weight-gradient / estimator branches on side streams an event pair would also count
    # whatever shares the SMs with the bracketed kernel
    eng = getattr(tr.model, "_engine", None)
    tr.overlap_branches = False
    if eng is not None:
        eng.overlap_wgrad = False
        eng.parallel_stats = False
    meter.timed = {"conv_gemm", "conv_direct_fwd", "conv_direct_dgrad", "conv_direct_wgrad", "fc_fwd", "conv_wgrad", "latent_fwd", "latent_bwd",
                   "bn_finalize", "bn_finalize_apply", "bn_relu_apply", "bn_bwd_coef", "bn_bwd_apply", "bn_act_fwd", "bn_reduce",
                   "sigmoid_mse_bwd", "conv_pack_weight", "colsum", "snn_finalize", "mi_estimator", "mi_bound_bwd", "adam_step",
                   "peer_gather", "peer_allreduce"}
    def queued_step(i):
        # eager launches are CPU-bound: without work queued ahead, an event pair around one op would also time the host
        # side of the op (output allocation, argument checks).  A spin kernel in front lets the host run ahead, so the
        # events bracket back-to-back device execution only.
        torch.cuda._sleep(int(1.5e8))   # ~75 ms: longer than the host needs to enqueue one eager step
        step_dev(i)
        torch.cuda.synchronize()

    for i in range(2):
        queued_step(i)
    dbg("profiling pass done")
    prof = meter.elapsed_ms()
    launches_per_step = meter.launches() // 2
    dominant = max(prof, key=lambda k: prof[k][1]) if prof else None
    breakdown = {k: round(v[1] / 2, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    meter.reset()
    meter.timed = {dominant} if dominant else set()
    for i in range(K):          # eager pass over the same K steps: live CUDA-event timing of the dominant kernel
        queued_step(i)
    dom = meter.elapsed_ms().get(dominant, (0, 0.0))
    dom_bytes = meter.bytes.get(dominant, 0)
    eager_launches = meter.launches()
    meter.reset()
    meter.timed = set()
    tr.use_cuda_graph = graphed
    tr.overlap_branches = True
    if eng is not None:
        eng.overlap_wgrad = True
        eng.parallel_stats = True
    dbg("eager kernel-timing pass done")

    # ---- timed region: device-resident inputs
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step_dev(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / K
    dbg("timed region done")
    launches = tr._graph["launches"] * K if graphed else meter.launches()

    # ---- end to end through the trainers' own input path (`VAETrainer.prefetch`, what `fit()` uses): pinned host batch ->
    # H2D on the copy stream (batch i+1 while step i runs) -> step -> D2H of the step's scalars into pinned memory, every
    # step; the region ends when the last read-back has landed
    def e2e_pass(n, sink):
        i = 0
        for X, y in tr.prefetch(pool_h[j % N_POOL] for j in range(n)):
            out = tr.train_step(X, y)
            vals = torch.cat([out[0].detach().view(1), out[1].detach().view(-1)])   # the step's scalars, one read-back
            sink[i].copy_(vals, non_blocking=True)
            i += 1
        return vals

    probe = step_e2e(0)
    sink = torch.zeros(max(K, 4), probe.numel(), dtype=torch.float32).pin_memory()
    e2e_pass(4, sink)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    t0.record()
    vals = e2e_pass(K, sink)
    t1.record()
    barrier()
    ms_e2e_ev, ms_e2e_wall = t0.elapsed_time(t1) / K, (time.perf_counter() - wall0) * 1e3 / K
    ms_e2e = max(ms_e2e_ev, ms_e2e_wall)
    e2e_finite = bool(torch.isfinite(sink[:K, 0]).all())   # the read-back really holds every step's loss
    dbg("e2e region done")
    clocks = sampler.stop()   # sampled across both timed regions (device-resident and end-to-end)
    h2d = pool_h[0][0].numel() * 4 + pool_h[0][1].numel() * 8
    d2h = vals.numel() * 4

    peer_phases = None
    if world > 1 and tr.dist.peer is not None:
        # where the collectives' time goes: CTA-0 %globaltimer stamps of the last 64 peer calls (stage own pieces | publish + wait
        # for every peer's flag = flag round trip + waiting for the slowest rank | pull)
        tl = tr.dist.peer.timeline()
        tl = tl[(tl > 0).all(1)]
        if tl.numel():
            ph = ((tl[:, 1:] - tl[:, :-1]) / 1e3)
            peer_phases = dict(calls=int(tl.shape[0]), stage_us=float(ph[:, 0].mean()), publish_wait_us=float(ph[:, 1].mean()),
                               publish_wait_us_max=float(ph[:, 1].max()), pull_us=float(ph[:, 2].mean()))
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    # ---- roofline of the dominant kernel (algorithmic flops from the layer geometry)
    pk = peaks()
    fl = layer_flops(cfg["arch"], cfg["cin"], cfg["z"], B)
    # forwards per step: CLEAR 1; TC 2 (second full forward for the discriminator); MIM 2 encoder + 6 decoder passes
    n_enc, n_dec = {"clear": (1, 1), "tc": (2, 2), "mim": (2, 6)}[cfg["kind"]]
    alg = {"conv_gemm": n_enc * fl["tc_enc_fwd"] + n_dec * fl["tc_dec_fwd"] + fl["tc_dgrad"], "conv_wgrad": fl["tc_wgrad"]}
    roof = None
    if dominant is not None and dom[0] > 0:
        # Two floors per launch set: operand bytes / HBM peak and (GEMM kernels) algorithmic flops / tensor peak; the larger
        # floor is the roof that binds.  Operand bytes = every tensor argument and result of the op counted once per call
        # (clear_vae_b200/_ops.py:_tensor_bytes) — an im2col re-read or an L2 miss does not add to it.
        per_step_ms = dom[1] / K
        bytes_step = dom_bytes / K
        gbs = bytes_step / (per_step_ms * 1e-3) / 1e9
        floor_hbm = bytes_step / (pk["hbm"] * 1e9)
        flops = alg.get(dominant)
        floor_tc = flops / (pk["tf"] * 1e12) if flops else 0.0
        common = dict(kernel=dominant, launches_per_step=dom[0] // K, ms_per_step=per_step_ms, peak_source=pk["src"],
                      timing="CUDA events around every launch of this kernel, eager single-stream pass over the same K steps "
                             "(host run-ahead behind a spin kernel, so the pairs bracket device time only)",
                      algorithmic_bytes_per_launch=bytes_step / max(1, dom[0] // K),
                      hbm=dict(achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"]))
        # DRAM bytes per launch from the committed `ncu --set full` capture of this config (profiles/r2_ncu_traffic.json); null if
        # this kernel / config was not captured
        traffic = ncu_traffic(dominant, args.config)
        if flops:
            tfs = flops / (per_step_ms * 1e-3) / 1e12
            common["tensor"] = dict(achieved=tfs, peak=pk["tf"], unit="TFLOP/s", frac=tfs / pk["tf"],
                                    algorithmic_gflop_per_step=flops / 1e9)
        if floor_tc > floor_hbm:
            roof = dict(bound="tensor", achieved=tfs, peak=pk["tf"], unit="TFLOP/s", frac=tfs / pk["tf"], traffic=traffic, **common)
        else:
            roof = dict(bound="hbm", achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"], traffic=traffic, **common)

    if not e2e_finite:
        print(json.dumps(dict(error="non-finite loss read back during the end-to-end region")), file=sys.stderr, flush=True)
        sys.exit(4)
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            slow = cfg["arch"] == "VAE64"
            csteps, cwarm = (3, 1) if slow else (10, 2)
            sps, cms, cores, kind, note = time_cpu(cfg, csteps, cwarm, pool_h)
            cpu = dict(value=sps, unit="samples/s", cores=cores, kind=kind,
                       sample=f"{csteps} full steps of batch {B} after {cwarm} warm-up on the host; {note}; {cms:.0f} ms/step")
        eager = None
        if world == 1 and not args.no_eager_baseline:
            eager = eager_gpu_baseline(cfg, dev, pool_h, K, W)
        lat = None
        if world == 1 and not args.no_latent:
            try:
                lat = latent_roofline(dev, pk)
            except Exception as e:  # never lose the headline line to the side measurement
                lat = dict(error=repr(e))
        line = dict(metric="train samples/sec", value=world * B / (ms * 1e-3), unit="samples/s", n_gpus=world, steps=K, warmup=W,
                    ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="bf16" if args.precision == "bf16" else "f32", data="synthetic", config=bench_config(args.config),
                    impl_config=dict(global_batch=world * B, parallelism=f"dp{world}", cuda_graph=graphed,
                                     bn="per-GPU batch statistics",
                                     conv_math=("bf16 operands, fp32 accumulate (tcgen05)" if args.precision == "bf16" else
                                                "fp32 activations, bf16 x 3 split products, fp32 accumulate (tcgen05)"),
                                     precision=args.precision, latent_math="fp32",
                                     collectives=("none (1 GPU)" if world == 1 else "one-shot peer-memory kernels over NVLink (csrc/peer_comm.cu)"
                                                  if tr.dist.peer is not None else "NCCL"),
                                     fallbacks="none: missing extension / unsupported discriminator / CPU tensors raise",
                                     peer_error=(tr.dist.peer.error() if (world > 1 and tr.dist.peer is not None) else None),
                                     peer_phases_rank0=peer_phases),
                    clocks=clocks,
                    e2e=dict(value=world * B / (ms_e2e * 1e-3), unit="samples/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             ms_per_step=ms_e2e, ms_per_step_events=ms_e2e_ev, ms_per_step_wall=ms_e2e_wall, losses_finite=e2e_finite,
                             path="VAETrainer.prefetch (H2D of batch i+1 on a copy stream) -> train_step -> async D2H of the "
                                  "step's scalars into pinned memory; region ends when the last read-back has landed"),
                    gpu_launches=launches, gpu_launches_per_step=launches_per_step, kernel_ms_per_step=breakdown,
                    roofline=roof, latent_roofline=lat, cpu_baseline=cpu, eager_gpu_baseline=eager, parity=parity)
        if world == 1 and not args.sub and not args.no_configs:
            line["configs"] = other_configs(args, K, W)
        print(json.dumps(line), flush=True)
    if world > 1:
        # CUDA graphs that captured NCCL kernels keep the communicator busy: tearing the process group down with
        # them alive can hang, so release the graphs first and leave without the (optional) NCCL teardown
        tr._graph = None
        sys.stdout.flush()
        sys.stderr.flush()
        torch.cuda.synchronize()
        td.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
