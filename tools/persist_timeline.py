"""Per-role timeline of the persistent conv GEMM (clearvae_debug_conv_timeline hook): for each CTA's first tiles, when the
producers started / finished issuing the tile, when the MMA warp committed it, when the epilogue finished it."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clear_vae_b200 import _ops
from clear_vae_b200.engine import nhwc_strides, FPROP, EPI_BIAS_STATS, out_hw

ops = _ops.ops()
lib = ctypes.CDLL(_ops.lib_paths()[0])
lib.clearvae_debug_conv_timeline.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024


def run(name, tr, k, op, cin, cout, hin):
    hout = out_hw(tr, k, 2, 1, op, hin)
    geom = [int(tr), k, 2, 1, op, cin, cout, hin, hin]
    src = torch.randn(B, hin, hin, cin, device=dev).to(torch.bfloat16)
    w = torch.randn((cin, cout, k, k) if tr else (cout, cin, k, k), device=dev)
    pw = ops.conv_pack_weight(geom, FPROP, w)
    bias = torch.zeros(cout, device=dev)
    dst = torch.empty(B, hout, hout, cout, device=dev, dtype=torch.bfloat16)
    st = torch.zeros(2 * cout + 2, dtype=torch.float64, device=dev)
    go = lambda: ops.conv_gemm(geom, FPROP, B, src, nhwc_strides(hin, hin, cin), None, None, False, pw, bias, dst,
                               nhwc_strides(hout, hout, cout), EPI_BIAS_STATS, None, [0, 0, 0, 0], None, None, st)
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    buf = torch.zeros(64 * 512, dtype=torch.int64, device=dev)
    lib.clearvae_debug_conv_timeline(ctypes.c_void_p(buf.data_ptr()))
    go()
    torch.cuda.synchronize()
    lib.clearvae_debug_conv_timeline(None)
    t = buf.view(-1, 64).cpu().double()
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    print(f"== {name}: {t.shape[0]} CTAs, kernel span {float(t[:, 1].max() - t0) / 1e3:.1f} us, CTA life mean {float((t[:, 1] - t[:, 0]).mean()) / 1e3:.1f} us")
    ph = t[:, 56:62]
    ok = ph[:, 5] > 0
    if ok.any():
        m = ph[ok].mean(0)
        print(f"   epilogue thread, cycles per tile (mean over {int(ok.sum())} CTAs, {m[5]:.1f} tiles each): accumulator wait {m[0] / m[5]:.0f} | TMEM load + bias/mask {m[1] / m[5]:.0f}"
              f" | global stores {m[2] / m[5]:.0f} | statistics {m[3] / m[5]:.0f} | whole loop {m[4] / m[5]:.0f}")
    for tile in range(10):
        s = t[:, 4 + tile * 4:8 + tile * 4]
        ok = (s > 0).all(1)
        if ok.sum() == 0:
            break
        s = s[ok] - t[ok, 0:1]
        m = s.mean(0) / 1e3
        print(f"   tile {tile}: producer start {m[0]:6.2f}  issued {m[1]:6.2f}  mma committed {m[2]:6.2f}  epilogue done {m[3]:6.2f}  (us after CTA start, mean over {int(ok.sum())} CTAs)")


run("convT 64->32 7->14", True, 3, 1, 64, 32, 7)
run("convT 128->64 4->7", True, 3, 0, 128, 64, 4)
run("conv 32->64 14->7", False, 3, 0, 32, 64, 14)
