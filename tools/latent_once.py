"""One forward + backward of the fused latent block at a chosen size (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clear_vae_b200 import _ops
from clear_vae_b200.latent import latent_block
_ops.load()
B, D = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator().manual_seed(0)
mu_c = torch.randn(B, D, generator=g).cuda().requires_grad_(True)
mu_s = torch.randn(B, D, generator=g).cuda().requires_grad_(True)
lv = (torch.randn(B, D, generator=g) * .3).cuda().requires_grad_(True)
eps = torch.randn(B, D, generator=g).cuda()
lab = torch.randint(0, 10, (B,), generator=g).cuda()
for _ in range(2):
    z, sc = latent_block([mu_c, mu_s], [lv, lv], [eps, eps], lab, snn=[1, 1], ps=[False, True], temperature=0.1)
    (z.sum() + sc[:4].sum()).backward()
torch.cuda.synchronize()
print("ok")
