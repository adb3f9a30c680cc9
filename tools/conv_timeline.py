"""Per-CTA phase timeline of one conv GEMM launch (uses the clearvae_debug_conv_timeline hook)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clear_vae_b200 import _ops
from clear_vae_b200.engine import nhwc_strides, FPROP, DGRAD, EPI_BIAS_STATS

ops = _ops.ops()
lib = ctypes.CDLL(_ops.lib_paths()[0])
lib.clearvae_debug_conv_timeline.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024


def run(name, geom, role, hin, cin, cout, hout, transposed):
    src = torch.randn(B, hin, hin, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(cin, cout, geom[1], geom[1], device=dev) if transposed else torch.randn(cout, cin, geom[1], geom[1], device=dev)
    pw = ops.conv_pack_weight(geom, role, w)
    bias = torch.zeros(cout, device=dev)
    dst = torch.empty(B, hout, hout, cout, device=dev, dtype=torch.bfloat16)
    st = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
    def go():
        ops.conv_gemm(geom, role, B, src, nhwc_strides(hin, hin, cin), None, None, False, pw, bias, dst, nhwc_strides(hout, hout, cout),
                      EPI_BIAS_STATS, None, [0, 0, 0, 0], None, None, st)
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); go(); e1.record(); torch.cuda.synchronize()
    buf = torch.zeros(8 * 65536, dtype=torch.int64, device=dev)
    lib.clearvae_debug_conv_timeline(ctypes.c_void_p(buf.data_ptr()))
    go()
    torch.cuda.synchronize()
    lib.clearvae_debug_conv_timeline(None)
    t = buf.view(-1, 8).cpu()
    t = t[t[:, 0] > 0].double()
    t0 = t[:, 0].min()
    names = ["prologue", "issue loads", "loads land", "mma done", "epilogue", "teardown"]
    print(f"== {name}: {t.shape[0]} CTAs, kernel {e0.elapsed_time(e1)*1e3:.1f} us (events), span {float(t[:, 6].max() - t0)/1e3:.1f} us")
    for i, n in enumerate(names):
        d = t[:, i + 1] - t[:, i]
        print(f"   {n:12s} mean {float(d.mean())/1e3:7.2f} us  p50 {float(d.median())/1e3:7.2f}  max {float(d.max())/1e3:7.2f}")
    life = t[:, 6] - t[:, 0]
    print(f"   CTA life     mean {float(life.mean())/1e3:7.2f} us; start-time quantiles (us): " +
          " ".join(f"{float(q)/1e3:.1f}" for q in torch.quantile(t[:, 0] - t0, torch.tensor([0.1, 0.25, 0.5, 0.75, 0.9, 1.0], dtype=torch.double))))


# decoder convT 64->32, 7x7 -> 14x14 (k3 s2 p1 op1): the (392,1,4) launch
run("convT 64->32 7->14", [1, 3, 2, 1, 1, 64, 32, 7, 7], FPROP, 7, 64, 32, 14, True)
run("convT 128->64 4->7", [1, 3, 2, 1, 0, 128, 64, 4, 4], FPROP, 4, 128, 64, 7, True)
run("conv 32->64 14->7", [0, 3, 2, 1, 0, 32, 64, 14, 14], FPROP, 14, 32, 64, 7, False)
run("conv 64->128 7->4", [0, 3, 2, 1, 0, 64, 128, 7, 7], FPROP, 7, 64, 128, 4, False)
