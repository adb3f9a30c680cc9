"""Runs W warm-up steps then ONE training step inside cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="mim_club")
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
cfg = bench.CONFIGS[a.config]
dev = torch.device("cuda", 0)
tr = bench.build_trainer(cfg, dev)
tr.model.train()
g = torch.Generator().manual_seed(101)
X = torch.rand(cfg["B"], cfg["cin"], cfg["hw"], cfg["hw"], generator=g).to(dev)
y = torch.randint(0, cfg["ncls"], (cfg["B"],), generator=g).to(dev)
for _ in range(a.warmup):
    tr.train_step(X, y)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.steps):
    tr.train_step(X, y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
