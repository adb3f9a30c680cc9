"""One forward + one backward of the FFMA latent kernels on a data-parallel shard shape (1024 local rows x 8192 gathered
columns, D = 8) — the launch pair `ncu --set full -k regex:snn_` captures for profiles/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clear_vae_b200 import _ops
from clear_vae_b200.latent import _workspace

ops = _ops.ops()
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
B, Bg, D = 1024, 8192, 8
mu = torch.randn(Bg, D, generator=g).to(dev)
lab = torch.randint(0, 10, (Bg,), generator=g).to(dev)
rows, lr = mu[:B].contiguous(), lab[:B].contiguous()
ws = _workspace(dev, ops.latent_workspace_bytes(B, Bg, D, 1))
wsb = _workspace(dev, ops.latent_bwd_workspace_bytes(B, Bg, D, 1), "bwd")
for _ in range(2):
    _, _, st = ops.latent_fwd([rows], [None], [None], [mu], [None], lr, lab, [1], [0], 0, 0, 0, 0.1, False, False, ws)
    st_all = st[0].repeat(Bg // B, 1)
    sc = torch.zeros(8, device=dev)
    ops.snn_finalize(st_all, 0, sc)
    gs = torch.tensor([0.0, 0.0, 1.0, 0.0], device=dev)
    ops.latent_bwd([rows], [None], [None], [mu], [None], [st_all], None, lr, lab, [1], [0], 0, 0, 0, 0.1, sc, gs, wsb)
torch.cuda.synchronize()
print("ok")
