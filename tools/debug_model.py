"""Prints per-tensor deviations of the CUDA model from (a) the bf16-rounding engine emulator and
(b) the fp32 reference golden."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import load_step, seeded_model, state_of, emulated_step, sample_of

def run(tag):
    torch.backends.cudnn.allow_tf32 = False
    g, meta, hyper = load_step(tag)
    m = seeded_model(meta, "cuda"); m.train()
    st0 = state_of(m)
    X = torch.tensor(g["X"]).cuda(); label = torch.tensor(g["label"]).cuda()
    eps = (torch.tensor(g["eps/0"]).cuda(), torch.tensor(g["eps/1"]).cuda())
    ps = hyper.get("ps", False); kind = meta["kind"]
    snn = [1, 1] if kind == "clear" else [1, 0]
    xhat, recon, z, sc, lp = m.fused_step_forward(X, label, temperature=hyper["temperature"], snn=snn, ps=[False, bool(ps)], eps=eps)
    slope = float(g["slope"])
    loss = recon + slope * sc[0] + slope * sc[1] + hyper["alpha"] * sc[2]
    if kind == "clear":
        s = sc[3] if ps else -sc[3]
        loss = loss + hyper["alpha"] * s
    loss.backward(); torch.cuda.synchronize()
    em = emulated_step(st0, meta, hyper, g, True, slope, device="cuda")
    def rel(a, b): return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
    def l2(a, b): return float(np.linalg.norm((a - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-30))
    n = lambda t: t.detach().float().cpu().numpy()
    print(f"== {tag}: recon {float(recon):.5f} emu {float(em['recon']):.5f} ref {float(g['recon']):.5f} | c {float(sc[2]):.5f} emu {float(em['c']):.5f} ref {float(g['c_loss']):.5f}"
          f" | kl_c {float(sc[0]):.5f} emu {float(em['kl_c']):.5f} ref {float(g['kl_c']):.5f}")
    lat = torch.cat([lp[k] for k in ("mu_c", "logvar_c", "mu_s", "logvar_s")], 1)
    print(f"  lat vs emu: max-rel {rel(n(lat), n(em['lat'])):.2e} l2 {l2(n(lat), n(em['lat'])):.2e};  xhat vs emu {rel(n(xhat), n(em['xhat'])):.2e}")
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        e = em["grads"][k]
        print(f"  grad {k:24s} vs emu: max-rel {rel(n(p.grad), n(e)):.2e} l2 {l2(n(p.grad), n(e)):.2e} | vs fp32 ref sample l2 {l2(sample_of(p.grad,256), g['grad_sample/'+k]):.2e}")

if __name__ == "__main__":
    for t in sys.argv[1:] or ["clear_vae28_ps", "clear_vae64_ps"]:
        run(t)
