"""fp32x3 mode vs the fp64 CPU oracle, next to the reference's OWN fp32 GPU run (cuDNN/cuBLAS, TF32 off) vs the same oracle:
the second number is the cross-implementation floor of fp32 arithmetic at that size (ReLU-mask flips of pre-activations
within ~1e-6 of zero make per-tensor gradient differences of 1e-4..1e-3 at B >= 256 even between two fp32 implementations)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from oracle import parity, ref_runner as rr


def ref_gpu_grads(cfg, state, X, y, eps):
    """gradients of the unmodified reference modules on cuda:0 in fp32 (TF32 off) for the CLEAR loss at annealer step 0"""
    tu, trainer, losses, vae, mi = rr.modules()
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    m = getattr(vae, cfg["arch"])(total_z_dim=cfg["z"], in_channel=cfg["cin"]).cuda()
    m.load_state_dict(state)
    m.train()
    seq = iter([e.cuda() for e in eps])
    orig = torch.randn_like
    torch.randn_like = lambda t, **k: next(seq)
    try:
        xhat, lp = m(X.cuda())
    finally:
        torch.randn_like = orig
    hp = cfg["hp"]
    rec, kc, ks = losses.vae_loss(xhat, X.cuda(), **lp)
    c = losses.contrastive_loss(lp["mu_c"], lp["logvar_c"], y.cuda(), "cosine", hp["temperature"])
    s = losses.contrastive_loss(lp["mu_s"], lp["logvar_s"], y.cuda(), "cosine", hp["temperature"], ps=hp["ps"])
    if not hp["ps"]:
        s = -s
    bt = hp["beta"] / 2.0
    (rec + bt * kc + bt * ks + hp["alpha"] * c + hp["alpha"] * s).backward()
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return {k: p.grad.detach().cpu() for k, p in m.named_parameters() if p.grad is not None}


def run(name, B, floor=False, **eng_flags):
    cfg = dict(bench.CONFIGS[name]); cfg["B"] = B
    tr = bench.build_trainer(cfg, torch.device("cuda"))
    tr.model.conv_precision = "fp32x3"
    eng = tr.model._eng()
    for k, v in eng_flags.items():
        setattr(eng, k, v)
    g = torch.Generator().manual_seed(101)
    X = torch.rand(B, cfg["cin"], cfg["hw"], cfg["hw"], generator=g)
    y = torch.randint(0, cfg["ncls"], (B,), generator=g)
    state = {k: v.detach().clone() for k, v in tr.model.state_dict().items()}
    res, so = parity.compare_step(tr, cfg, X, y, oracle_dtype=torch.float64, return_oracle=True)
    gr = {k[5:]: v[2] for k, v in res.items() if k.startswith("grad/")}
    srt = sorted(gr.values())
    keys = [k for k in gr if k.startswith("decoder")][::3]
    line = (f"{name} B={B} {eng_flags}: OURS median {srt[len(srt)//2]:.2e} max {srt[-1]:.2e} lat "
            f"{max(v[2] for k, v in res.items() if k.startswith('latent/')):.2e} | " + " ".join(f"{k}={gr[k]:.1e}" for k in keys))
    if floor and cfg["kind"] == "clear":
        ge = torch.Generator().manual_seed(7)
        D = cfg["z"] // 2
        eps = (torch.randn(B, D, generator=ge), torch.randn(B, D, generator=ge))
        rg = ref_gpu_grads(cfg, state, X, y, eps)
        l2 = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
        fl = {k: l2(rg[k], so.last_grads[k]) for k in gr if k in rg}
        fs = sorted(fl.values())
        line += f"\n      REFERENCE fp32 on GPU vs the same oracle: median {fs[len(fs)//2]:.2e} max {fs[-1]:.2e} | " + " ".join(f"{k}={fl[k]:.1e}" for k in keys)
    print(line, flush=True)


if __name__ == "__main__":
    for B in (128, 256, 1024):
        run("clear28", B, floor=True)
    run("mim_club", 1024)
    for B in (32, 128):
        run("clear64", B, floor=True)
