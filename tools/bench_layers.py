"""Average launch time of the conv GEMMs of the 28x28 / 64x64 stacks (CUDA events, 20 launches after warm-up)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clear_vae_b200 import _ops
from clear_vae_b200.engine import nhwc_strides, FPROP, DGRAD, EPI_BIAS_STATS, EPI_MASK_STATS, out_hw

ops = _ops.ops()
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
arch = sys.argv[2] if len(sys.argv) > 2 else "VAE"


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


def layer(name, tr, k, op, cin, cout, hin):
    hout = out_hw(tr, k, 2, 1, op, hin)
    geom = [int(tr), k, 2, 1, op, cin, cout, hin, hin]
    src = torch.randn(B, hin, hin, cin, device=dev).to(torch.bfloat16)
    w = torch.randn((cin, cout, k, k) if tr else (cout, cin, k, k), device=dev)
    bias = torch.zeros(cout, device=dev)
    dst = torch.empty(B, hout, hout, cout, device=dev, dtype=torch.bfloat16)
    st = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
    pw = ops.conv_pack_weight(geom, FPROP, w)
    f = lambda: ops.conv_gemm(geom, FPROP, B, src, nhwc_strides(hin, hin, cin), None, None, False, pw, bias, dst,
                              nhwc_strides(hout, hout, cout), EPI_BIAS_STATS, None, [0, 0, 0, 0], None, None, st)
    t_f = timeit(f)
    # data gradient with the previous block's mask + BN sums
    dy = torch.randn(B, hout, hout, cout, device=dev).to(torch.bfloat16)
    g = torch.empty(B, hin, hin, cin, device=dev)
    sc, sh = torch.ones(cin, device=dev), torch.zeros(cin, device=dev)
    st2 = torch.zeros(2 * cin, dtype=torch.float64, device=dev)
    pwd = ops.conv_pack_weight(geom, DGRAD, w)
    d = lambda: ops.conv_gemm(geom, DGRAD, B, dy, nhwc_strides(hout, hout, cout), None, None, False, pwd, None, g,
                              nhwc_strides(hin, hin, cin), EPI_MASK_STATS, src, nhwc_strides(hin, hin, cin), sc, sh, st2)
    t_d = timeit(d)
    dw = torch.zeros_like(w)
    wg = lambda: ops.conv_wgrad(geom, B, src, nhwc_strides(hin, hin, cin), None, None, False, dy, nhwc_strides(hout, hout, cout), dw)
    t_w = timeit(wg)
    fl = 2.0 * B * (hin * hin if tr else hout * hout) * cin * cout * k * k
    print(f"{name:22s} fprop {t_f:7.1f} us ({fl/t_f/1e6:6.1f} TF/s)  dgrad {t_d:7.1f} us  wgrad {t_w:7.1f} us")


if arch == "VAE":
    layer("conv 32->64 14->7", False, 3, 0, 32, 64, 14)
    layer("conv 64->128 7->4", False, 3, 0, 64, 128, 7)
    layer("convT 128->64 4->7", True, 3, 0, 128, 64, 4)
    layer("convT 64->32 7->14", True, 3, 1, 64, 32, 7)
else:
    ch = [32, 64, 128, 256, 512]
    for i in range(4):
        layer(f"conv {ch[i]}->{ch[i+1]} {32 >> i}", False, 4, 0, ch[i], ch[i + 1], 32 >> i)
    for i in range(4):
        layer(f"convT {ch[4-i]}->{ch[3-i]} {2 << i}", True, 4, 0, ch[4 - i], ch[3 - i], 2 << i)
