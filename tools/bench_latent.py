"""Times the fused latent-loss kernels (CUDA events on the current stream)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clear_vae_b200 import _ops
from clear_vae_b200.latent import latent_block


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    _ops.load()
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    for D in (8, 32):
        for B in (128, 1024, 4096, 16384, 65536):
            mu_c = torch.randn(B, D, generator=g).to(dev).requires_grad_(True)
            mu_s = torch.randn(B, D, generator=g).to(dev).requires_grad_(True)
            lv = (torch.randn(B, D, generator=g) * .3).to(dev).requires_grad_(True)
            eps = torch.randn(B, D, generator=g).to(dev)
            lab = torch.randint(0, 10, (B,), generator=g).to(dev)
            it = 3 if B >= 16384 else 20

            def fwd():
                return latent_block([mu_c, mu_s], [lv, lv], [eps, eps], lab, snn=[1, 1], ps=[False, True], temperature=0.1)
            t_f = timeit(lambda: fwd(), it)
            z, sc = fwd()
            loss = z.sum() + sc[:4].sum()

            def bwd():
                torch.autograd.grad(loss, [mu_c, mu_s, lv], retain_graph=True)
            t_b = timeit(bwd, it)
            pairs = 2.0 * B * B
            print(json.dumps(dict(B=B, D=D, fwd_ms=round(t_f, 4), bwd_ms=round(t_b, 4),
                                  fwd_gpairs_s=round(pairs / t_f * 1e-6, 1), bwd_gpairs_s=round(pairs / t_b * 1e-6, 1))), flush=True)


if __name__ == "__main__":
    main()
