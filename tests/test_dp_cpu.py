"""CPU, world_size 2 over gloo: host-side data-parallel logic (the CUDA kernels themselves need a GPU)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from oracle import latent_oracle as lo


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from clear_vae_b200.latent import DistSpec, _all_gather_rows
        from clear_vae_b200.trainer import VAETrainer
        dist = DistSpec(td.group.WORLD, rank, world)
        g = torch.Generator().manual_seed(0)
        B, D = 6, 4
        mu_all = torch.randn(world * B, D, generator=g)
        lab_all = torch.randint(-5, 5, (world * B,), generator=g) + (1 << 40) * (torch.arange(world * B) % 2)
        mu, lab = mu_all[rank * B:(rank + 1) * B], lab_all[rank * B:(rank + 1) * B]
        # (1) packed all-gather: int64 labels travel bit-exactly inside the fp32 payload
        packed = torch.cat([mu, lab.view(B, 1).view(torch.float32)], dim=1)
        got = _all_gather_rows(packed, dist)
        assert torch.equal(got[:, :D], mu_all)
        assert torch.equal(got[:, D:].contiguous().view(torch.int64).view(-1), lab_all)
        # (2) per-rank partial (sum, count) of the row losses add up to the single-process loss
        s, c = lo.contrastive_partial(mu.numpy(), mu.numpy(), lab.numpy(), "cosine", 0.1, mu_cols=mu_all.numpy(),
                                      logvar_cols=mu_all.numpy(), label_cols=lab_all.numpy(), row_offset=rank * B)
        t = torch.tensor([s, float(c)], dtype=torch.float64)
        td.all_reduce(t)
        want = lo.contrastive(mu_all.numpy(), mu_all.numpy(), lab_all.numpy(), "cosine", 0.1)
        assert abs(float(t[0] / t[1]) - want) < 1e-12
        # (3) row-side gradient of the global loss * world, averaged over ranks, equals the global gradient rows
        full = lo.snn_grad(mu_all.numpy(), lab_all.numpy(), "cosine", 0.1)
        mine = torch.tensor(full[rank * B:(rank + 1) * B]) * world     # what latent_block's backward returns per rank
        avg = torch.zeros(world * B, D, dtype=torch.float64)
        avg[rank * B:(rank + 1) * B] = mine / world                      # "DDP average" of disjoint row blocks
        td.all_reduce(avg)
        assert np.allclose(avg.numpy(), full, rtol=0, atol=0)
        # (4) flat gradient averaging used by the trainers
        tr = VAETrainer.__new__(VAETrainer)
        tr.dist = dist
        ps = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(1))]
        ps[0].grad = torch.full((3, 2), float(rank + 1))
        ps[1].grad = torch.arange(5.0) * (rank + 1)
        scale = tr._sync_grads(ps)                                       # third parameter has no grad: skipped
        assert scale == 1.0 / world                                      # the average is applied by the optimiser kernel
        assert torch.allclose(ps[0].grad * scale, torch.full((3, 2), 1.5))
        assert torch.allclose(ps[1].grad * scale, torch.arange(5.0) * 1.5)
        # (5) the CLEAR-MIM estimator batches are gathered once as [world, 5, B, 2D] -> [5, world*B, 2D]
        loc = torch.arange(5 * 3 * 4, dtype=torch.float32).view(5, 3, 4) + 1000.0 * rank
        allz = torch.empty((world * 5,) + tuple(loc.shape[1:]))
        td.all_gather_into_tensor(allz, loc)
        glob = allz.view((world,) + tuple(loc.shape)).permute(1, 0, 2, 3).reshape(5, world * 3, 4)
        for j in range(5):
            for r in range(world):
                assert torch.equal(glob[j, r * 3:(r + 1) * 3], torch.arange(5 * 3 * 4, dtype=torch.float32).view(5, 3, 4)[j] + 1000.0 * r)
        assert ps[2].grad is None
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        td.destroy_process_group()


def test_data_parallel_host_logic_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
