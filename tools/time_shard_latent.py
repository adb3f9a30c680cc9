"""Latent kernels on a data-parallel shard shape (B local rows x Bg gathered columns), CUDA-event timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clear_vae_b200 import _ops
from clear_vae_b200.latent import _workspace

ops = _ops.ops()
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
for B, Bg, D in [(1024, 1024, 8), (1024, 2048, 8), (1024, 8192, 8), (128, 1024, 32), (2048, 2048, 8)]:
    mu = torch.randn(Bg, D, generator=g).to(dev)
    lab = torch.randint(0, 10, (Bg,), generator=g).to(dev)
    rows, lr = mu[:B].contiguous(), lab[:B].contiguous()
    ws = _workspace(dev, ops.latent_workspace_bytes(B, Bg, D, 1))
    wsb = _workspace(dev, ops.latent_bwd_workspace_bytes(B, Bg, D, 1), "bwd")
    fwd = lambda: ops.latent_fwd([rows], [None], [None], [mu], [None], lr, lab, [1], [0], 0, 0, 0, 0.1, False, False, ws)
    _, _, st = fwd()
    st_all = torch.zeros(Bg, 2, device=dev); st_all[:B] = st[0]; st_all[B:] = st[0][:1]
    sc = torch.zeros(8, device=dev); ops.snn_finalize(st_all, 0, sc)
    gs = torch.tensor([0.0, 0.0, 1.0, 0.0], device=dev)
    bwd = lambda w: ops.latent_bwd([rows], [None], [None], [mu], [None], [st_all], None, lr, lab, [1], [0], 0, 0, 0, 0.1, sc, gs, w)

    def t(fn, n=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(10):
                fn()
        gr.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            gr.replay()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (n * 10) * 1e3

    print(f"rows {B} x cols {Bg}, D={D}: forward {t(fwd):.1f} us; backward split {t(lambda: bwd(wsb)):.1f} us, unsplit {t(lambda: bwd(None)):.1f} us", flush=True)
