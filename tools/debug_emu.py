"""CUDA engine vs bf16 emulator on random inputs at a chosen batch (no goldens needed)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from clear_vae_b200.models.vae import VAE, VAE64
from oracle.engine_emulator import Emulator
from oracle import model_oracle as mo

def run(arch, B, zdim=64, cin=3):
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(5)
    hw = 28 if arch == "VAE" else 64
    m = (VAE if arch == "VAE" else VAE64)(zdim, cin).cuda(); m.train()
    st = {k: v.detach().clone() for k, v in m.state_dict().items()}
    X = torch.rand(B, cin, hw, hw, device="cuda"); label = torch.randint(0, 7, (B,), device="cuda")
    D = zdim // 2
    eps = (torch.randn(B, D, device="cuda"), torch.randn(B, D, device="cuda"))
    m._eng().debug = {}
    direct = not os.environ.get("NO_DIRECT")
    m._eng().use_direct = direct
    xhat, recon, z, sc, lp = m.fused_step_forward(X, label, temperature=0.1, snn=[1, 1], ps=[False, True], eps=eps)
    loss = recon + 0.03 * sc[0] + 0.03 * sc[1] + 100 * sc[2] + 100 * sc[3]
    loss.backward(); torch.cuda.synchronize()
    edev, edt = (os.environ.get("EMU_DEV", "cuda"), torch.float64 if os.environ.get("EMU_F64") else torch.float32)
    st = {k: (v.to(edev).to(edt) if v.is_floating_point() else v.to(edev)) for k, v in st.items()}
    X, label = X.to(edev).to(edt), label.to(edev)
    eps = (eps[0].to(edev).to(edt), eps[1].to(edev).to(edt))
    em = Emulator(st, arch, cin, round_bf16=True, update_running=True, direct_boundary=direct)
    out = em.forward(X, eps[0], eps[1], target=X)
    dgr, dz = em.backward_decoder(out["tape"], X, 1.0)
    lat = out["lat"].detach().requires_grad_(True)
    mu_c, lv_c, mu_s, lv_s = (lat[:, j * D:(j + 1) * D] for j in range(4))
    zz = torch.cat([mu_c + eps[0] * torch.exp(0.5 * lv_c), mu_s + eps[1] * torch.exp(0.5 * lv_s)], 1)
    kl = lambda a, l: -0.5 * (1 + l - a * a - l.exp()).sum(1).mean()
    tot = (zz * dz).sum() + 0.03 * kl(mu_c, lv_c) + 0.03 * kl(mu_s, lv_s) + 100 * mo.contrastive(mu_c, lv_c, label, "cosine", 0.1) \
        + 100 * mo.contrastive(mu_s, lv_s, label, "cosine", 0.1, ps=True)
    (dlat,) = torch.autograd.grad(tot, lat)
    egr = em.backward_encoder(out["tape"], dlat)
    grads = {**dgr, **egr}
    l2 = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30))
    latc = torch.cat([lp[k] for k in ("mu_c", "logvar_c", "mu_s", "logvar_s")], 1)
    print(f"== {arch} B={B}: lat l2 {l2(latc, out['lat']):.2e} xhat l2 {l2(xhat, out['xhat']):.2e} recon {float(recon):.4f}/{float(out['recon']):.4f}")
    dbg = m._eng().debug
    for i, (a, L) in enumerate(zip(dbg["enc_raw"], out["tape"]["enc"])):
        y = L["y"]
        a = a.float().view(y.shape[0], y.shape[1], -1) if i == len(dbg["enc_raw"]) - 1 else a.float().permute(0, 3, 1, 2).reshape(y.shape[0], y.shape[1], -1)
        yy = y.reshape(y.shape[0], y.shape[1], -1)
        nflip = float((a.cpu().double() != yy.cpu().double()).double().mean())
        print(f"  enc raw[{i}] l2 {l2(a, yy):.2e} differing-elements {nflip:.2e}")
    print(f"  fc raw l2 {l2(dbg['fc_raw'], out['tape']['fc']['raw']):.2e}  fc act l2 {l2(dbg['fc_act'].float(), out['tape']['fc']['a']):.2e}")
    for i, (a, L) in enumerate(zip(dbg["dec_raw"], out["tape"]["dec"])):
        y = L["y"]
        a = a.float() if i == len(dbg["dec_raw"]) - 1 else a.float().permute(0, 3, 1, 2)
        nflip = float((a.cpu().double() != y.cpu().double()).double().mean())
        print(f"  dec raw[{i}] l2 {l2(a, y):.2e} differing-elements {nflip:.2e}")
    sd = m.state_dict()
    for k in sd:
        if k.endswith("running_var"):
            print(f"  {k:32s} l2 {l2(sd[k], st[k]):.2e}")
    for k, p in m.named_parameters():
        if p.grad is not None:
            print(f"  grad {k:24s} l2 {l2(p.grad, grads[k]):.2e}")

if __name__ == "__main__":
    run(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 64)
