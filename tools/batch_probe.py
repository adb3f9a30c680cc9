import sys, os
sys.path.insert(0, os.getcwd())
import torch, bench
from oracle import parity
for name in ("mim_club",):
    for B in (64, 96, 192, 512, 768):
        cfg = dict(bench.CONFIGS[name]); cfg["B"] = B
        tr = bench.build_trainer(cfg, torch.device("cuda"))
        g = torch.Generator().manual_seed(101)
        X = torch.rand(B, 3, 28, 28, generator=g); y = torch.randint(0, 10, (B,), generator=g)
        res = parity.compare_step(tr, cfg, X, y)
        gr = sorted(((v[2], k) for k, v in res.items() if k.startswith("grad/")), reverse=True)
        print(name, B, "worst", [(round(a, 3), k[5:]) for a, k in gr[:4]], "median", round(gr[len(gr)//2][0], 4), flush=True)
