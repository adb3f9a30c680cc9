import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from clear_vae_b200 import _ops
from clear_vae_b200.losses import contrastive_loss
from oracle import latent_oracle as lo
_ops.load()
lib = ctypes.CDLL(_ops.lib_paths()[0])
for B, D, ps in [(1024, 8, False), (4096, 8, False), (4096, 32, True), (8192, 32, False), (16384, 8, True)]:
    g = torch.Generator().manual_seed(9)
    mu = torch.randn(B, D, generator=g); lab = torch.randint(0, 10, (B,), generator=g)
    wg = lo.snn_grad(mu.numpy(), lab.numpy(), "cosine", 0.1, ps)
    for force in (1, 1 << 30):
        lib.clearvae_set_latent_tc_min_rows(ctypes.c_int32(force))
        m = mu.cuda().requires_grad_(True)
        l = contrastive_loss(m, torch.zeros_like(m), lab.cuda(), "cosine", 0.1, ps=ps)
        l.backward()
        gr = m.grad.cpu().numpy()
        print(B, D, ps, "TC" if force == 1 else "FFMA", "loss", float(l), "grad max-relerr", np.abs(gr - wg).max() / np.abs(wg).max(),
              "l2", np.linalg.norm(gr - wg) / np.linalg.norm(wg))
