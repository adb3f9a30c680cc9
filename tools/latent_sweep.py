"""BASELINE configs[4]: contrastive-loss global-batch sweep 1k-64k latents, all-gathered across the ranks.

    python tools/latent_sweep.py                                  # 1 GPU
    torchrun --nproc-per-node N tools/latent_sweep.py             # N GPUs: B_g rows sharded, columns gathered

Per global batch B_g and latent width D: forward and backward time of the fused latent block (content + style term)
on this rank's B_g / N rows against all B_g columns — including the exchanges of operands / labels / row statistics —
CUDA events, max over ranks; pairs/s over the WHOLE job against the MUFU ex2 roof (1 exp per (row, column, term)).
One JSON line per point on rank 0.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as td

from clear_vae_b200 import _ops
from clear_vae_b200.latent import DistSpec, latent_block

MUFU_PEAK_GPAIRS = 4590.0   # 148 SMs x 16 ex2/clk x 1.965 GHz minus measurement loss (tools/mufu_probe on this pool)


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        td.init_process_group("nccl", device_id=dev)
        from clear_vae_b200.peer import PeerComm
        dist = DistSpec(td.group.WORLD, rank, world, PeerComm.create(td.group.WORLD, rank, world, dev, nbytes=64 << 20))
    _ops.load()
    sizes = [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192, 16384, 32768, 65536]
    w = torch.tensor([0.1, 0.1, 100.0, 100.0, 0, 0, 0, 0], device=dev)
    for D in (8, 32):
        for Bg in sizes:
            if Bg % world:
                continue
            B = Bg // world
            g = torch.Generator().manual_seed(Bg + D)          # same global tensors on every rank, each takes its rows
            full = [torch.randn(Bg, D, generator=g) * s for s in (1, 1, .3)]
            lab = torch.randint(0, 10, (Bg,), generator=g)
            eps = torch.randn(Bg, D, generator=g)
            sl = slice(rank * B, (rank + 1) * B)
            mu_c, mu_s, lv = (t[sl].to(dev).requires_grad_(True) for t in full)
            e, lb = eps[sl].to(dev), lab[sl].to(dev)

            def once():
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
                z, sc = latent_block([mu_c, mu_s], [lv, lv], [e, e], lb, snn=[1, 1], ps=[False, True], temperature=0.1, dist=dist)
                e1.record()
                torch.autograd.backward([sc], [w])
                e2.record()
                torch.cuda.synchronize()
                mu_c.grad = mu_s.grad = lv.grad = None
                return e0.elapsed_time(e1), e1.elapsed_time(e2), float(sc[2])

            n = 3 if Bg >= 16384 else 10
            for _ in range(2):
                once()
            tf = tb = 0.0
            for _ in range(n):
                if world > 1:
                    td.barrier()
                a, b, loss = once()
                tf, tb = tf + a / n, tb + b / n
            t = torch.tensor([tf, tb], device=dev)
            if world > 1:
                td.all_reduce(t, op=td.ReduceOp.MAX)
            tf, tb = float(t[0]), float(t[1])
            pairs = 2.0 * Bg * Bg
            if rank == 0:
                print(json.dumps(dict(Bg=Bg, D=D, n_gpus=world, rows_per_gpu=B, fwd_ms=round(tf, 4), bwd_ms=round(tb, 4), c_loss=round(loss, 6),
                                      fwd_gpairs_s=round(pairs / tf * 1e-6, 1), bwd_gpairs_s=round(pairs / tb * 1e-6, 1),
                                      fwd_frac_of_mufu_roof=round(pairs / tf * 1e-6 / (MUFU_PEAK_GPAIRS * world), 4),
                                      bwd_frac_of_mufu_roof=round(pairs / tb * 1e-6 / (MUFU_PEAK_GPAIRS * world), 4),
                                      collectives=("none" if world == 1 else "peer" if dist.peer is not None else "nccl"))), flush=True)
    if world > 1:
        td.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
