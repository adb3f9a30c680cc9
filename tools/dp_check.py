"""torchrun --nproc-per-node N tools/dp_check.py : data-parallel parity on real GPUs.
(1) latent block: sharded rows + gathered columns == single-GPU result on the global batch (loss, row grads)
(2) full CLEAR-VAE step with SyncBN: rank-averaged gradients == single-GPU gradients on the global batch (bf16 noise)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as td
from clear_vae_b200.latent import DistSpec, latent_block
from clear_vae_b200.models.vae import VAE

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
td.init_process_group("nccl", device_id=dev)
from clear_vae_b200.peer import PeerComm
peer = PeerComm.create(td.group.WORLD, rank, world, dev)
dist_nccl = DistSpec(td.group.WORLD, rank, world)
dist = DistSpec(td.group.WORLD, rank, world, peer)
if rank == 0:
    print("collectives:", "peer-memory kernels (NVLink, cudaIpc)" if peer is not None else "NCCL (peer memory unavailable)", flush=True)
g = torch.Generator().manual_seed(1)
if peer is not None and not os.environ.get("DP_CHECK_ONLY"):
    # peer kernels vs NCCL, eager and replayed from a CUDA graph (device-side call counter), odd sizes included
    gg = torch.Generator().manual_seed(100 + rank)
    a = [torch.randn(n, d, generator=gg).to(dev) for n, d in ((1024, 8), (1024, 2), (333, 3))]
    lab0 = torch.randint(0, 1 << 40, (1024,), generator=gg).to(dev)
    grads = [torch.randn(n, generator=gg).to(dev) for n in (864, 32, 18432, 64, 73728, 128, 262144, 5, 1)]
    def once():
        out = peer.gather(a + [lab0])
        gs = [t.clone() for t in grads]
        peer.allreduce_(gs)
        return out, gs
    out, gs = once()
    ok = True
    for t, o in zip(a + [lab0], out):
        r = torch.empty_like(o); td.all_gather_into_tensor(r, t); ok &= torch.equal(r, o)
    for t, o in zip(grads, gs):
        r = t.clone(); td.all_reduce(r); ok &= torch.allclose(r, o, rtol=1e-5, atol=1e-5)
    allg = [torch.empty(world * t.numel(), device=dev) for t in gs[:1]]
    td.all_gather_into_tensor(allg[0], gs[0].reshape(-1).contiguous())
    same_bits = all(torch.equal(allg[0][r * gs[0].numel():(r + 1) * gs[0].numel()], gs[0].reshape(-1)) for r in range(world))
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        once()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize(); td.barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
        out_g, gs_g = once()
    for _ in range(50):
        graph.replay()
    torch.cuda.synchronize()
    for o, og in zip(out + gs, out_g + gs_g):
        ok &= torch.equal(o, og)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_graph(fn, reps=10, iters=100):
        fn(); torch.cuda.synchronize(); td.barrier()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, capture_error_mode="thread_local"):
            for _ in range(reps):
                fn()
        for _ in range(5):
            gr.replay()
        td.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            gr.replay()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (iters * reps) * 1e3

    def phases(tag):
        tl = peer.timeline()
        tl = tl[(tl > 0).all(1)]
        ph = ((tl[:, 1:] - tl[:, :-1]) / 1e3).mean(0).tolist()
        print(f"   rank {rank} {tag}: CTA-0 stage {ph[0]:.2f} us, publish+wait {ph[1]:.2f} us, pull {ph[2]:.2f} us", flush=True)

    zg = [torch.zeros_like(t) for t in grads]
    lat_pieces = [a[0], lab0]                                  # the latent exchange of configs[1]: mu [1024, 8] + labels
    t_g = timed_graph(lambda: peer.gather(lat_pieces)); phases("gather(latents 40 KB)")
    z5 = [torch.zeros(1024, 16, device=dev) for _ in range(5)]
    t_g5 = timed_graph(lambda: peer.gather(z5)); phases("gather(5 x 64 KB)")
    t_a = timed_graph(lambda: peer.allreduce_(zg)); phases("all-reduce(1.4 MB, 9 tensors)")
    flat = torch.cat([t.reshape(-1) for t in zg]); packed = torch.cat([t.reshape(-1) for t in a[:1]] + [lab0.view(torch.float32)])
    gout = torch.empty(world * packed.numel(), device=dev)
    t_ng = timed_graph(lambda: td.all_gather_into_tensor(gout, packed))
    t_na = timed_graph(lambda: td.all_reduce(flat))
    t_peer, t_nccl = t_g + t_a, t_ng + t_na
    if rank == 0:
        print(f"   per call, back to back in a graph: peer gather {t_g:.1f} us (5 pieces {t_g5:.1f} us) vs NCCL all-gather {t_ng:.1f} us; "
              f"peer all-reduce {t_a:.1f} us vs NCCL all-reduce {t_na:.1f} us", flush=True)
    if rank == 0:
        print(f"peer kernels world={world}: match NCCL {bool(ok)}, all-reduce bit-identical across ranks {bool(same_bits)}, error flag {peer.error()}; "
              f"gather + all-reduce {t_peer:.1f} us vs NCCL {t_nccl:.1f} us", flush=True)
for Bl, D in ([] if os.environ.get("DP_CHECK_ONLY") else [(64, 8), (512, 32), (4096, 8)]):
    Bg = Bl * world
    mu_c, lv_c, mu_s, lv_s, e_c, e_s = ((torch.randn(Bg, D, generator=g) * s).to(dev) for s in (1, .3, 1, .3, 1, 1))
    lab = torch.randint(0, 10, (Bg,), generator=g).to(dev)
    sl = slice(rank * Bl, (rank + 1) * Bl)
    full = [t.clone().requires_grad_(True) for t in (mu_c, lv_c, mu_s, lv_s)]
    z, sc = latent_block([full[0], full[2]], [full[1], full[3]], [e_c, e_s], lab, snn=[1, 1], ps=[False, True], temperature=0.1)
    w = torch.tensor([0.1, 0.1, 100.0, 100.0, 0, 0, 0, 0], device=dev)
    torch.autograd.backward([sc], [w])
    loc = [t[sl].clone().requires_grad_(True) for t in (mu_c, lv_c, mu_s, lv_s)]
    z2, sc2 = latent_block([loc[0], loc[2]], [loc[1], loc[3]], [e_c[sl].contiguous(), e_s[sl].contiguous()], lab[sl].contiguous(),
                           snn=[1, 1], ps=[False, True], temperature=0.1, dist=dist)
    torch.autograd.backward([sc2], [w])
    ok_loss = torch.allclose(sc2[2:4], sc[2:4], rtol=2e-6, atol=1e-7)
    # SNN grads come back scaled by `world` (the trainers average over ranks afterwards); KL is a local mean over Bl rows
    gm = loc[0].grad / world
    ref = full[0].grad[sl]
    kl_part_ref = 0.1 * full[0].detach()[sl] / Bg
    kl_part_loc = 0.1 * loc[0].detach() / Bl / world
    err = ((gm - kl_part_loc) - (ref - kl_part_ref)).abs().max() / ref.abs().max()
    if rank == 0:
        print(f"latent DP Bl={Bl} D={D} world={world}: loss match {bool(ok_loss)} ({sc2[2].item():.6f} vs {sc[2].item():.6f}), row-grad relerr {err.item():.2e}", flush=True)

# ---- full model step with SyncBN
torch.manual_seed(7)
m1 = VAE(16, 3).to(dev); m1.train()
m2 = VAE(16, 3).to(dev); m2.load_state_dict(m1.state_dict()); m2.train()
m2.dist, m2.sync_bn = dist_nccl, True
Bl = 64; Bg = Bl * world
X = torch.rand(Bg, 3, 28, 28, generator=g).to(dev); lab = torch.randint(0, 10, (Bg,), generator=g).to(dev)
eps = (torch.randn(Bg, 8, generator=g).to(dev), torch.randn(Bg, 8, generator=g).to(dev))
w = torch.tensor([0.06, 0.06, 100.0, 100.0, 0, 0, 0, 0], device=dev)
xh, rec, z, sc, _ = m1.fused_step_forward(X, lab, temperature=0.1, snn=[1, 1], ps=[False, True], eps=eps)
torch.autograd.backward([rec, sc], [torch.ones_like(rec), w])
sl = slice(rank * Bl, (rank + 1) * Bl)
xh2, rec2, z2, sc2, _ = m2.fused_step_forward(X[sl].contiguous(), lab[sl].contiguous(), temperature=0.1, snn=[1, 1], ps=[False, True],
                                              eps=(eps[0][sl].contiguous(), eps[1][sl].contiguous()), dist=dist)
torch.autograd.backward([rec2, sc2], [torch.ones_like(rec2), w])
worst = 0.0
for (k, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
    if p1.grad is None:
        continue
    g2 = p2.grad.clone(); td.all_reduce(g2); g2 /= world
    e = ((g2 - p1.grad).norm() / (p1.grad.norm() + 1e-30)).item()
    worst = max(worst, e)
    if rank == 0 and e > 5e-2:
        print("   large deviation:", k, e, flush=True)
rec_g = rec2.clone(); td.all_reduce(rec_g); rec_g /= world
if rank == 0:
    print(f"model DP+SyncBN world={world}: recon {rec_g.item():.4f} vs {rec.item():.4f}; c_loss {sc2[2].item():.5f} vs {sc[2].item():.5f}; worst grad l2 relerr {worst:.2e}", flush=True)

# ---- (3) whole trainer steps: global-batch semantics of the MI / TC terms and the overlapped gradient buckets
from clear_vae_b200.utils.trainer_utils import get_clearmimvae_trainer, get_cleartcvae_trainer
from clear_vae_b200.models.mi_estimator import CLUBSample


def trainer_pair(kind, est=None):
    def make():
        torch.manual_seed(11)
        if kind == "mim":
            return get_clearmimvae_trainer(1 / 8, est, 3, 5e-4, 2e-3, 16, 1e2, 0.1, dev, "VAE", 3)
        return get_cleartcvae_trainer(1 / 8, 1, 5e-4, 1e-4, 16, 1e2, 0.1, dev, "VAE", 3)
    a, b = make(), make()
    b.dist = dist
    b.overlap_grad_sync = os.environ.get("DP_OVERLAP", "1") == "1"
    b.model.dist, b.model.sync_bn = dist_nccl, True
    b._local = {}
    for p_ in b.model.parameters():   # local gradients as autograd delivers them (registered before any bucket hook)
        p_.register_post_accumulate_grad_hook(lambda q, store=b._local: store.__setitem__(id(q), q.grad.detach().clone()) if q.grad is not None else None)
    return a, b


Bl = 96; Bg = Bl * world
gg = torch.Generator().manual_seed(5)
X = torch.rand(Bg, 3, 28, 28, generator=gg).to(dev); lab = torch.randint(0, 10, (Bg,), generator=gg).to(dev)
rn = lambda: torch.randn(Bg, 8, generator=gg).to(dev)
eps = (rn(), rn()); inner = [(rn(), rn()) for _ in range(5)]; perm = torch.randperm(Bg, generator=gg)
sl = slice(rank * Bl, (rank + 1) * Bl)
cut = lambda pair: tuple(t[sl].contiguous() for t in pair)
for kind, est in (("mim", "CLUBSample"), ("mim", "L1OutUB"), ("tc", None)):
    one, dp = trainer_pair(kind, est)
    for step in range(2):      # step 0 is the synchronous all-reduce that arms the buckets, step 1 runs them overlapped
        if kind == "mim":
            kw1 = dict(eps=eps, inner_eps=inner); kw2 = dict(eps=cut(eps), inner_eps=[cut(p_) for p_ in inner])
            if est == "CLUBSample":
                kw1["perm"] = perm; kw2["perm"] = perm
        else:
            kw1 = dict(eps=eps, eps2=inner[0]); kw2 = dict(eps=cut(eps), eps2=cut(inner[0]))
        o1 = one.train_step(X, lab, **kw1)
        o2 = dp.train_step(X[sl].contiguous(), lab[sl].contiguous(), **kw2)
        torch.cuda.synchronize()
        worst, errs, red_ok = 0.0, [], True
        for (k, p1), (_, p2) in zip(one.model.named_parameters(), dp.model.named_parameters()):
            if p1.grad is None:
                continue
            e = ((p2.grad / world - p1.grad).norm() / (p1.grad.norm() + 1e-30)).item()   # dp grads hold the rank SUM after the all-reduce
            worst = max(worst, e)
            errs.append((e, k))
            want = dp._local[id(p2)].clone(); td.all_reduce(want)
            red_ok &= ((p2.grad - want).norm() / (want.norm() + 1e-30)).item() < 1e-5
        if rank == 0:
            print("      reduction == NCCL sum of local gradients:", red_ok, "| largest deviations:", [(round(e_, 4), k_) for e_, k_ in sorted(errs, reverse=True)[:5]], flush=True)
        aux1 = one.mi_estimator if kind == "mim" else one.factor_cls
        aux2 = dp.mi_estimator if kind == "mim" else dp.factor_cls
        aux_err = max(((a_ - b_).abs().max() / (a_.abs().max() + 1e-30)).item() for a_, b_ in zip(aux1.parameters(), aux2.parameters()))
        rec = o2[0].clone(); td.all_reduce(rec); rec /= world
        mi2 = o2[2].clone()
        if kind == "tc":
            td.all_reduce(mi2); mi2 /= world      # row-separable bound: the global value is the rank average
        if rank == 0:
            print(f"trainer DP {kind}/{est} step {step} world={world}: recon {rec.item():.4f} vs {o1[0].item():.4f}; bound {mi2.item():.6f} vs {o1[2].item():.6f}; "
                  f"worst VAE grad l2 relerr {worst:.2e} (bf16 noise ~1e-2); aux-net params after update max relerr {aux_err:.2e}; "
                  f"buckets overlapped: {bool(getattr(dp, '_buckets', None)) and step > 0}", flush=True)
# ---- (4) overlapped gradient buckets == one synchronous all-reduce, bit for bit (same fixed rank order per element)
if peer is not None:
    from clear_vae_b200.utils.trainer_utils import get_clearvae_trainer
    def mk(overlap):
        torch.manual_seed(13)
        t = get_clearvae_trainer(1 / 32, True, 3e-5, 64, 1e2, 0.1, dev, "VAE64", 3)   # 23.9 MB of gradients: two buckets, chunked
        t.dist, t.overlap_grad_sync = dist, overlap
        return t
    ta, tb = mk(True), mk(False)
    # local gradients of `ta` as autograd delivers them, cloned by a hook registered BEFORE the bucket hooks (hooks run in
    # registration order): the overlapped all-reduce must return exactly their rank sum
    local = {}
    for p_ in ta.model.parameters():
        p_.register_post_accumulate_grad_hook(lambda q: local.__setitem__(id(q), q.grad.detach().clone()) if q.grad is not None else None)
    X64 = torch.rand(16, 3, 64, 64, generator=gg).to(dev); y64 = torch.randint(0, 7, (16,), generator=gg).to(dev)
    e64 = (torch.randn(16, 32, generator=gg).to(dev), torch.randn(16, 32, generator=gg).to(dev))
    worst = 0.0
    for step in range(3):
        tb.model.load_state_dict(ta.model.state_dict())      # same weights every step: only the reduction schedule differs
        ta.train_step(X64, y64, eps=e64); tb.train_step(X64, y64, eps=e64)
        torch.cuda.synchronize()
        exact = True
        for pa in ta.model.parameters():
            if pa.grad is not None:
                want = local[id(pa)].clone(); td.all_reduce(want)       # NCCL sum of the local gradients
                exact &= ((pa.grad - want).norm() / (want.norm() + 1e-30)).item() < 1e-5     # rank-order sum vs NCCL's order
        if rank == 0:
            print(f"   step {step}: overlapped buckets == NCCL sum of the local gradients: {exact}", flush=True)
        for pa, pb in zip(ta.model.parameters(), tb.model.parameters()):
            if pa.grad is not None:   # the weight-gradient kernels add with fp32 atomics (order varies run to run): compare to 1e-5, not bitwise
                worst = max(worst, ((pa.grad - pb.grad).norm() / (pb.grad.norm() + 1e-30)).item())
    if rank == 0:
        print(f"overlapped buckets vs synchronous all-reduce (VAE64, 3 steps, world={world}): worst gradient rel-L2 difference {worst:.2e} "
              f"(fp32 atomic-order noise ~1e-6); buckets armed {getattr(ta, '_buckets', None) is not None}", flush=True)
    peer.check()
td.destroy_process_group()
