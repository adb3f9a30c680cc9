"""Phase timeline of one CLEAR-MIM step on the main stream (side-stream branches active): CUDA events between the phases of
`ClearMIMVAETrainer._device_step`, eager launches queued behind a spin kernel so the device runs back to back like a graph replay."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from clear_vae_b200 import _ops
from clear_vae_b200.optim import fused_adam_step
from clear_vae_b200.models.mi_estimator import CLUBSample

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "mim_club"]
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:   # under torchrun: the data-parallel step (rank 0 prints)
    import torch.distributed as td
    td.init_process_group("nccl", device_id=dev)
tr = bench.build_trainer(cfg, dev)
if world > 1:
    from clear_vae_b200.latent import DistSpec
    from clear_vae_b200.peer import PeerComm
    tr.dist = DistSpec(td.group.WORLD, rank, world, PeerComm.create(td.group.WORLD, rank, world, dev))
    for p_ in list(tr.model.parameters()):
        td.broadcast(p_.data, 0)
tr.model.train()
g = torch.Generator().manual_seed(101 + rank)
X = torch.rand(cfg["B"], cfg["cin"], cfg["hw"], cfg["hw"], generator=g).to(dev)
y = torch.randint(0, cfg["ncls"], (cfg["B"],), generator=g).to(dev)
for _ in range(4):
    tr.train_step(X, y)
torch.cuda.synchronize()
marks = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))

vae, est, hp = tr.model, getattr(tr, "mi_estimator", None), tr.hyperparameter
D = vae.z_dim
import clear_vae_b200.engine as E
orig_dec_bwd, orig_enc_bwd = E.DecoderFn.backward, E.EncoderFn.backward
def dec_bwd(ctx, *a):
    mark("bwd: loss/estimator grads -> decoder backward start")
    r = orig_dec_bwd(ctx, *a); mark("bwd: decoder done (incl. join of its weight gradients)"); return r
def enc_bwd(ctx, *a):
    mark("bwd: latent block backward done")
    r = orig_enc_bwd(ctx, *a); mark("bwd: encoder done (incl. join)"); return r
E.DecoderFn.backward, E.EncoderFn.backward = staticmethod(dec_bwd), staticmethod(enc_bwd)
orig_encode, orig_decode = vae.encode, vae._decode
def encode(*a, **k):
    r = orig_encode(*a, **k); mark("encoder forward done"); return r
def decode(*a, **k):
    r = orig_decode(*a, **k); mark("decoder forward done"); return r
vae.encode, vae._decode = encode, decode

for rep in range(3):
    marks.clear()
    torch.cuda._sleep(int(2e8))
    tr._begin_step(); tr._host_pre(); tr._upload_weights(dev)
    vae._engine.packs.epoch += 1
    mark("start")
    out = tr._device_step(X, y)
    mark("step end (all branches joined)")
    tr._end_step(dev); tr.annealer.step()
    torch.cuda.synchronize()
if rank != 0:
    sys.exit(0)
t0 = marks[0][1]
prev = 0.0
print(f"world {world} | {cfg['kind']} {cfg['arch']} B={cfg['B']}: phase boundaries on the main stream (ms since start, delta)")
for name, e in marks[1:]:
    t = t0.elapsed_time(e)
    print(f"  {t:8.3f}  +{t - prev:6.3f}  {name}")
    prev = t
