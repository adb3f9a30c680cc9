"""Phase totals of the weight-gradient GEMM (clearvae_debug_conv_timeline hook): per CTA, where producer thread 0, the MMA
thread and epilogue thread 0 spend their SM clocks, plus launch time by CUDA events."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clear_vae_b200 import _ops
from clear_vae_b200.engine import nhwc_strides, out_hw

ops = _ops.ops()
lib = ctypes.CDLL(_ops.lib_paths()[0])
lib.clearvae_debug_conv_timeline.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda")


def run(name, B, tr, k, op, cin, cout, hin):
    hout = out_hw(tr, k, 2, 1, op, hin)
    geom = [int(tr), k, 2, 1, op, cin, cout, hin, hin]
    src = torch.randn(B, hin, hin, cin, device=dev).to(torch.bfloat16)
    dy = torch.randn(B, hout, hout, cout, device=dev).to(torch.bfloat16)
    dw = torch.zeros((cin, cout, k, k) if tr else (cout, cin, k, k), device=dev)
    go = lambda: ops.conv_wgrad(geom, B, src, nhwc_strides(hin, hin, cin), None, None, False, dy, nhwc_strides(hout, hout, cout), dw)
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        go()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    buf = torch.zeros(64 * 4096, dtype=torch.int64, device=dev)
    lib.clearvae_debug_conv_timeline(ctypes.c_void_p(buf.data_ptr()))
    go()
    torch.cuda.synchronize()
    lib.clearvae_debug_conv_timeline(None)
    t = buf.view(-1, 64).cpu().double()
    t = t[t[:, 13] > 0]
    flops = 2.0 * B * (hin * hin if tr else hout * hout) * cin * cout * k * k
    if t.shape[0] == 0:   # the TMA kernel carries no stamps: launch time only
        print(f"== {name} B={B}: {us:.1f} us per launch ({flops / us / 1e6:.1f} TF/s)  [TMA-operand kernel; CLEARVAE_NO_TMA_WGRAD=1 for the cp.async kernel and its phase totals]")
        return
    life = (t[:, 1] - t[:, 0]) / 1e3
    span = (t[:, 1].max() - t[:, 0].min()) / 1e3
    kb = t[:, 13]
    m = lambda i: float((t[:, i] / kb).mean())
    print(f"== {name} B={B}: {us:.1f} us per launch ({flops / us / 1e6:.1f} TF/s), {t.shape[0]} working CTAs, span {span:.1f} us, CTA life {life.mean():.1f} us (max {life.max():.1f}), {kb.mean():.1f} k-blocks of 64 pixels per CTA")
    print(f"   cycles per k-block: producer math + cp.async issue {m(8):.0f} | free-stage wait {m(9):.0f} | own-copy wait {m(10):.0f} || MMA operand wait {m(11):.0f} | MMA issue {m(12):.0f}")
    print(f"   epilogue thread: accumulator wait {float(t[:, 14].mean()):.0f} cycles, scatter-add {float(t[:, 15].mean()):.0f} cycles")


if len(sys.argv) > 1 and sys.argv[1] == "64":
    run("conv 32->64 32->16", 512, False, 4, 0, 32, 64, 32)
    run("conv 128->256 8->4", 512, False, 4, 0, 128, 256, 8)
    run("convT 256->128 4->8", 512, True, 4, 0, 256, 128, 4)
    run("convT 64->32 16->32", 512, True, 4, 0, 64, 32, 16)
else:
    run("conv 32->64 14->7", 1024, False, 3, 0, 32, 64, 14)
    run("convT 128->64 4->7", 1024, True, 3, 0, 128, 64, 4)
    run("convT 64->32 7->14", 1024, True, 3, 1, 64, 32, 7)
