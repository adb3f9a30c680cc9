"""Whole-step parity numbers at the benchmarked sizes, both conv precisions, fp32 and fp64 CPU oracle (no asserts):
    python tools/parity_report.py [out.json] [configs...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from oracle import parity

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/parity_report.json"
names = sys.argv[2:] or ["clear28", "mim_club", "mim_l1out", "clear64", "tc64"]
rep = {}
for name in names:
    for prec in ("bf16", "fp32x3"):
        for od in (torch.float32, torch.float64):
            if prec == "bf16" and od == torch.float64:
                continue
            cfg = dict(bench.CONFIGS[name])
            if od == torch.float64 and cfg["arch"] == "VAE64":
                cfg["B"] = min(cfg["B"], 128)
            t0 = time.time()
            tr = bench.build_trainer(cfg, torch.device("cuda"))
            tr.model.conv_precision = prec
            g = torch.Generator().manual_seed(101)
            X = torch.rand(cfg["B"], cfg["cin"], cfg["hw"], cfg["hw"], generator=g)
            y = torch.randint(0, cfg["ncls"], (cfg["B"],), generator=g)
            try:
                res = parity.compare_step(tr, cfg, X, y, oracle_dtype=od)
            except Exception as e:
                rep[f"{name}/{prec}/{od}"] = dict(error=repr(e))
                print(name, prec, od, "ERROR", repr(e), flush=True)
                continue
            sc = {k: v for k, v in res.items() if "/" not in k}
            gr = {k: v[2] for k, v in res.items() if k.startswith("grad/")}
            lt = {k: v[2] for k, v in res.items() if k.startswith("latent/")}
            worst = sorted(gr.items(), key=lambda kv: -kv[1])[:5]
            rep[f"{name}/{prec}/{od}"] = dict(B=cfg["B"], scalars=sc, latents=lt, grads=gr, seconds=time.time() - t0)
            print(name, prec, od, "B", cfg["B"], {k: (v[2] if not isinstance(v[0], list) else v[2]) for k, v in sc.items()},
                  "lat", max(lt.values()), "grad worst", worst[:3], "median", sorted(gr.values())[len(gr) // 2], flush=True)
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
json.dump(rep, open(out, "w"), indent=1)
