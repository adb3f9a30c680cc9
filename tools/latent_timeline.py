"""Per-role timeline of the tensor-core latent backward (clearvae_debug_latent_timeline hook): for column tiles of CTA (0, 0),
SM-clock stamps of the producer, the MMA issuer and the two epilogue groups -- shows which hand-off bounds the column sweep."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clear_vae_b200 import _ops
from clear_vae_b200.latent import latent_block

_ops.load()
lib = ctypes.CDLL(_ops.lib_paths()[0])
lib.clearvae_debug_latent_timeline.argtypes = [ctypes.c_void_p]
B, D = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, int(sys.argv[2]) if len(sys.argv) > 2 else 8
g = torch.Generator().manual_seed(0)
mu_c = torch.randn(B, D, generator=g).cuda().requires_grad_(True)
mu_s = torch.randn(B, D, generator=g).cuda().requires_grad_(True)
lv = (torch.randn(B, D, generator=g) * .3).cuda().requires_grad_(True)
eps = torch.randn(B, D, generator=g).cuda()
lab = torch.randint(0, 10, (B,), generator=g).cuda()


def once():
    z, sc = latent_block([mu_c, mu_s], [lv, lv], [eps, eps], lab, snn=[1, 1], ps=[False, True], temperature=0.1)
    (z.sum() + sc[:4].sum()).backward()


once()
torch.cuda.synchronize()
buf = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
lib.clearvae_debug_latent_timeline(ctypes.c_void_p(buf.data_ptr()))
once()
torch.cuda.synchronize()
lib.clearvae_debug_latent_timeline(None)
t = buf.view(64, 16).cpu()
t0 = int(t[16, 2])
names = {0: "prod free", 1: "prod staged", 2: "mma B seen", 3: "mma S issued", 4: "mma P seen", 5: "mma2 issued", 8: "epi S seen",
         9: "ld0", 10: "ld1", 11: "ld2", 12: "cmp0", 13: "cmp1", 14: "cmp2", 15: "P published"}
print(f"B={B} D={D}: cycles relative to tile 16's 'mma B seen'")
print("tile " + " ".join(f"{names[s]:>12s}" for s in sorted(names)))
for jt in range(16, 40):
    print(f"{jt:4d} " + " ".join(f"{int(t[jt, s]) - t0:12d}" if int(t[jt, s]) else f"{'-':>12s}" for s in sorted(names)))
d = (t[39, 15] - t[17, 15]).item() / 22.0
print(f"steady state: {d:.0f} cycles per column tile (MUFU floor: 128 x BN / 16 per tile)")
for a, b, what in ((8, 9, "S seen -> chunk 0 loaded"), (9, 12, "chunk 0 compute"), (12, 10, "chunk 0 st issue + chunk 1 load"), (10, 13, "chunk 1 compute"),
                   (13, 11, "st + chunk 2 load"), (11, 14, "chunk 2 compute"), (14, 15, "st + wait + publish"), (8, 15, "epilogue total")):
    print(f"  {what:34s} {(t[16:40, b] - t[16:40, a]).double().mean().item():8.0f}")
x = t[16:38]
print(f"  P published -> mma P seen          {(x[:, 4] - x[:, 15]).double().mean().item():8.0f}")
print(f"  mma P seen -> mma2 issued          {(x[:, 5] - x[:, 4]).double().mean().item():8.0f}")
print(f"  P published(jt) -> S seen(jt+2)    {(t[18:40, 8] - t[16:38, 15]).double().mean().item():8.0f}")
print(f"  mma S issued(jt) -> S seen(jt)     {(x[:, 8] - x[:, 3]).double().mean().item():8.0f}")
