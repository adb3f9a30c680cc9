"""Where does the end-to-end step time go?  Variants of the input path around the same graphed train_step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from clear_vae_b200 import _ops

_ops.load()
dev = torch.device("cuda", 0)
cfg = bench.CONFIGS["mim_club"]
tr = bench.build_trainer(cfg, dev)
tr.model.train()
B, K, NP = cfg["B"], 40, 16
g = torch.Generator().manual_seed(0)
pool_h = [(torch.rand(B, 3, 28, 28, generator=g).pin_memory(), torch.randint(0, 10, (B,), generator=g).pin_memory()) for _ in range(NP)]
pool_d = [(x.to(dev), y.to(dev)) for x, y in pool_h]
for i in range(3):
    tr.train_step(*pool_d[i])
tr.use_cuda_graph = True
for i in range(3):
    tr.train_step(*pool_d[i])
torch.cuda.synchronize()
sink = torch.zeros(K, 9).pin_memory()


def timed(name, body):
    body(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    body(K)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:50s} events {e0.elapsed_time(e1) / K:.3f} ms/step, wall {(time.perf_counter() - w0) * 1e3 / K:.3f} ms/step", flush=True)


def scal(out):
    return torch.cat([out[0].detach().view(1), out[1].detach().view(-1)])


def resident(n):
    for i in range(n):
        tr.train_step(*pool_d[i % NP])


def serial_sync(n):
    for i in range(n):
        x, y = pool_h[i % NP]
        out = tr.train_step(x.to(dev, non_blocking=True), y.to(dev, non_blocking=True))
        scal(out).to("cpu")


def serial_async(n):
    for i in range(n):
        x, y = pool_h[i % NP]
        out = tr.train_step(x.to(dev, non_blocking=True), y.to(dev, non_blocking=True))
        sink[i].copy_(scal(out), non_blocking=True)


def prefetch_async(n):
    i = 0
    for X, y in tr.prefetch(pool_h[j % NP] for j in range(n)):
        out = tr.train_step(X, y)
        sink[i].copy_(scal(out), non_blocking=True)
        i += 1


def prefetch_sync(n):
    for X, y in tr.prefetch(pool_h[j % NP] for j in range(n)):
        out = tr.train_step(X, y)
        scal(out).to("cpu")


def host_only(n):   # host cost of one step's enqueue with the GPU drained each step
    t = 0.0
    for i in range(n):
        torch.cuda.synchronize()
        w = time.perf_counter()
        tr.train_step(*pool_d[i % NP])
        t += time.perf_counter() - w
    print(f"   host enqueue per step: {t / n * 1e3:.3f} ms", flush=True)


copy_stream = torch.cuda.Stream()
scratch = [torch.empty_like(pool_d[0][0]) for _ in range(2)]


def background_copy(n, nbytes_frac=1.0):
    """resident inputs for the step; an unrelated H2D of a pinned batch runs on a copy stream every step"""
    rows = max(1, int(B * nbytes_frac))
    for i in range(n):
        with torch.cuda.stream(copy_stream):
            scratch[i & 1][:rows].copy_(pool_h[i % NP][0][:rows], non_blocking=True)
        tr.train_step(*pool_d[i % NP])


def copy_only(n):
    for i in range(n):
        with torch.cuda.stream(copy_stream):
            scratch[i & 1].copy_(pool_h[i % NP][0], non_blocking=True)
    copy_stream.synchronize()


timed("H2D copies alone (copy stream)", copy_only)
timed("device-resident inputs", resident)
timed("resident + unrelated 9.6 MB H2D per step", background_copy)
timed("resident + unrelated 1 MB H2D per step", lambda n: background_copy(n, 0.1))
timed("H2D on the step's stream + blocking D2H", serial_sync)
timed("H2D on the step's stream + async D2H", serial_async)
timed("prefetch stream + blocking D2H", prefetch_sync)
timed("prefetch stream + async D2H", prefetch_async)
host_only(10)
