"""GPU: tcgen05 implicit-GEMM conv kernels vs torch convolutions on bf16-rounded operands."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def nhwc_strides(t):  # t is [N,H,W,C] contiguous
    return [t.stride(0), t.stride(1), t.stride(2), t.stride(3)]


def run_gemm(geom, role, batch, src, src_strides, w, bias, dst, dst_strides, pre=None, relu=False, epi=0, mask=None,
             mask_strides=(0, 0, 0, 0), mscale=None, mshift=None, stats=None):
    from clear_vae_b200 import _ops
    ops = _ops.ops()
    pw = ops.conv_pack_weight(geom, role, w.contiguous())
    ops.conv_gemm(geom, role, batch, src, list(src_strides), None if pre is None else pre[0], None if pre is None else pre[1],
                  relu, pw, bias, dst, list(dst_strides), epi, mask, list(mask_strides), mscale, mshift, stats)
    return dst


@pytest.fixture(autouse=True)
def _no_tf32():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


def rel_err(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.mark.parametrize("B,K,N", [(300, 2048, 32), (256, 64, 2048), (128, 16, 2048), (1000, 2048, 128), (77, 40, 20)])
def test_linear_as_1x1(B, K, N):
    g = torch.Generator().manual_seed(B + K + N)
    x = bf(torch.randn(B, K, generator=g)).to(DEV)
    w = bf(torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    geom = [0, 1, 1, 0, 0, K, N, 1, 1]
    dst = torch.empty(B, N, device=DEV)
    stats = torch.zeros(2 * N, dtype=torch.float64, device=DEV)
    run_gemm(geom, 0, B, x, (K, 0, 0, 1), w, b, dst, (N, 0, 0, 1), stats=stats)
    want = F.linear(x, w, b)
    assert rel_err(dst, want) < 2e-5, rel_err(dst, want)
    assert torch.allclose(stats[:N].float(), want.sum(0), rtol=1e-4, atol=1e-3)
    assert torch.allclose(stats[N:].float(), (want * want).sum(0), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("B,K,N,role,f32", [(300, 2048, 32, 0, False), (1024, 2048, 32, 0, False), (77, 4096, 32, 0, True),
                                               (1000, 2048, 16, 1, True), (5, 512, 16, 1, False)])
def test_skinny_linear_without_statistics(B, K, N, role, f32):
    """The latent heads (N = 4D = 32) and the data gradient of the decoder's first Linear (N = 2D = 16) take the warp-MMA
    kernel (no split-K atomics): same contract as the general GEMM, bias optional, ragged batch."""
    g = torch.Generator().manual_seed(B + K + N)
    x = bf(torch.randn(B, K, generator=g)).to(DEV)
    if not f32:     # fp32 rows (bf16-representable here) are rounded to bf16 on load, like the general kernel's gather
        x = x.to(torch.bfloat16)
    if role == 0:   # y = x W^T + b, W [N, K]
        w = bf(torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
        b = torch.randn(N, generator=g).to(DEV)
        geom = [0, 1, 1, 0, 0, K, N, 1, 1]
        want = F.linear(x.float(), w, b)
    else:           # dz = dy W, W [K, N] (forward Linear N -> K)
        w = bf(torch.randn(K, N, generator=g) / K ** 0.5).to(DEV)
        b = None
        geom = [0, 1, 1, 0, 0, N, K, 1, 1]
        want = x.float() @ w
    dst = torch.full((B, N), float("nan"), device=DEV)
    run_gemm(geom, role, B, x, (K, 0, 0, 1), w, b, dst, (N, 0, 0, 1))
    assert rel_err(dst, want) < 2e-5, rel_err(dst, want)


@pytest.mark.parametrize("k,cin,cout,H,B", [(4, 32, 64, 32, 8), (3, 32, 64, 14, 16), (4, 3, 32, 64, 4), (3, 1, 32, 28, 8),
                                             (4, 256, 512, 4, 32), (3, 64, 128, 7, 33)])
def test_conv2d_fprop_and_dgrad(k, cin, cout, H, B):
    g = torch.Generator().manual_seed(k * 1000 + cin + cout)
    x = bf(torch.randn(B, cin, H, H, generator=g)).to(DEV)        # NCHW like the reference
    w = bf(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    want = F.conv2d(x, w, b, stride=2, padding=1)
    Ho = want.shape[-1]
    geom = [0, k, 2, 1, 0, cin, cout, H, H]
    # NCHW source read through strides (n, h, w, c)
    dst = torch.empty(B, Ho, Ho, cout, device=DEV)
    run_gemm(geom, 0, B, x, (x.stride(0), x.stride(2), x.stride(3), x.stride(1)), w, b, dst, nhwc_strides(dst))
    assert rel_err(dst.permute(0, 3, 1, 2), want) < 2e-5
    # NHWC source (vector gather path when cin % 8 == 0)
    xn = x.permute(0, 2, 3, 1).contiguous()
    dst2 = torch.empty(B, Ho, Ho, cout, device=DEV)
    run_gemm(geom, 0, B, xn, nhwc_strides(xn), w, b, dst2, nhwc_strides(dst2))
    assert rel_err(dst2.permute(0, 3, 1, 2), want) < 2e-5
    # data gradient: dy [B,Ho,Ho,cout] -> dx [B,H,H,cin]
    dy = bf(torch.randn(B, cout, Ho, Ho, generator=g)).to(DEV)
    dx_want = torch.nn.grad.conv2d_input(x.shape, w, dy, stride=2, padding=1)
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    dx = torch.empty(B, H, H, cin, device=DEV)
    run_gemm(geom, 1, B, dyn, nhwc_strides(dyn), w, None, dx, nhwc_strides(dx))
    assert rel_err(dx.permute(0, 3, 1, 2), dx_want) < 2e-5


@pytest.mark.parametrize("k,op,cin,cout,H,B", [(4, 0, 64, 32, 16, 8), (3, 0, 128, 64, 4, 16), (3, 1, 64, 32, 7, 8),
                                                (3, 1, 32, 3, 14, 8), (4, 0, 32, 3, 32, 4), (4, 0, 512, 256, 2, 32)])
def test_conv_transpose_fprop_and_dgrad(k, op, cin, cout, H, B):
    g = torch.Generator().manual_seed(k * 100 + op + cin + cout)
    x = bf(torch.randn(B, cin, H, H, generator=g)).to(DEV)
    w = bf(torch.randn(cin, cout, k, k, generator=g) / (cin * k * k) ** 0.5).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    want = F.conv_transpose2d(x, w, b, stride=2, padding=1, output_padding=op)
    Ho = want.shape[-1]
    geom = [1, k, 2, 1, op, cin, cout, H, H]
    xn = x.permute(0, 2, 3, 1).contiguous()
    dst = torch.empty(B, Ho, Ho, cout, device=DEV)
    run_gemm(geom, 0, B, xn, nhwc_strides(xn), w, b, dst, nhwc_strides(dst))
    assert rel_err(dst.permute(0, 3, 1, 2), want) < 2e-5
    # NCHW destination written through strides (what the last decoder layer does)
    dst2 = torch.empty(B, cout, Ho, Ho, device=DEV)
    run_gemm(geom, 0, B, xn, nhwc_strides(xn), w, b, dst2, (dst2.stride(0), dst2.stride(2), dst2.stride(3), dst2.stride(1)))
    assert rel_err(dst2, want) < 2e-5
    # data gradient of the transposed conv = strided gather
    dy = bf(torch.randn(B, cout, Ho, Ho, generator=g)).to(DEV)
    dx_want = F.conv2d(dy, w, None, stride=2, padding=1)
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    dx = torch.empty(B, H, H, cin, device=DEV)
    run_gemm(geom, 1, B, dyn, nhwc_strides(dyn), w, None, dx, nhwc_strides(dx))
    assert rel_err(dx.permute(0, 3, 1, 2), dx_want) < 2e-5


def test_preop_bf16_io_and_mask_epilogue():
    g = torch.Generator().manual_seed(42)
    B, cin, cout, H, k = 8, 64, 128, 16, 4
    raw = torch.randn(B, H, H, cin, generator=g).to(DEV)
    scale = (torch.rand(cin, generator=g) + 0.5).to(DEV)
    shift = (torch.randn(cin, generator=g) * 0.3).to(DEV)
    w = bf(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    raw16 = raw.to(torch.bfloat16)
    act = bf(torch.relu(raw16.float() * scale + shift))
    want = F.conv2d(act.permute(0, 3, 1, 2), w, b, stride=2, padding=1)
    Ho = want.shape[-1]
    geom = [0, k, 2, 1, 0, cin, cout, H, H]
    dst = torch.empty(B, Ho, Ho, cout, device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    run_gemm(geom, 0, B, raw16, nhwc_strides(raw16), w, b, dst, nhwc_strides(dst), pre=(scale, shift), relu=True, stats=stats)
    assert rel_err(dst.float().permute(0, 3, 1, 2), want) < 6e-3            # bf16 output rounding
    assert torch.allclose(stats[:cout].float(), want.sum((0, 2, 3)), rtol=1e-3, atol=5e-2)  # stats from fp32 accumulators
    # dgrad with the ReLU mask of the previous layer + BN-backward sums in the epilogue
    dy = bf(torch.randn(B, Ho, Ho, cout, generator=g)).to(DEV)
    dact_want = torch.nn.grad.conv2d_input((B, cin, H, H), w, dy.permute(0, 3, 1, 2), stride=2, padding=1).permute(0, 2, 3, 1)
    mask = (raw16.float() * scale + shift) > 0
    g_want = dact_want * mask
    gout = torch.empty(B, H, H, cin, device=DEV)
    st2 = torch.zeros(2 * cin, dtype=torch.float64, device=DEV)
    run_gemm(geom, 1, B, dy, nhwc_strides(dy), w, None, gout, nhwc_strides(gout), epi=1, mask=raw16,
             mask_strides=nhwc_strides(raw16), mscale=scale, mshift=shift, stats=st2)
    assert rel_err(gout, g_want) < 2e-5
    assert torch.allclose(st2[:cin].float(), g_want.sum((0, 1, 2)), rtol=1e-3, atol=1e-3)
    assert torch.allclose(st2[cin:].float(), (g_want * raw16.float()).sum((0, 1, 2)), rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("transposed,k,op,cin,cout,H,B", [(0, 4, 0, 32, 64, 32, 8), (0, 3, 0, 3, 32, 28, 16), (0, 4, 0, 256, 512, 4, 32),
                                                           (1, 4, 0, 512, 256, 2, 32), (1, 3, 1, 32, 3, 14, 8), (1, 4, 0, 64, 32, 16, 8),
                                                           (0, 4, 0, 3, 32, 64, 8), (1, 4, 0, 32, 3, 32, 8),
                                                           # the 28x28 stack: odd class grids (7, 3) take one zero-filled box column
                                                           (0, 3, 0, 32, 64, 14, 16), (0, 3, 0, 64, 128, 7, 16), (1, 3, 0, 128, 64, 4, 16),
                                                           (1, 3, 1, 64, 32, 7, 16), (1, 3, 1, 64, 32, 7, 5), (0, 4, 0, 128, 256, 8, 3)])
def test_weight_gradient(transposed, k, op, cin, cout, H, B):
    from clear_vae_b200 import _ops
    ops = _ops.ops()
    g = torch.Generator().manual_seed(7 * k + cin + cout + transposed)
    x = bf(torch.randn(B, cin, H, H, generator=g)).to(DEV)
    wshape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    w = torch.randn(*wshape, generator=g).to(DEV).requires_grad_(True)
    if transposed:
        y = F.conv_transpose2d(x, w, None, stride=2, padding=1, output_padding=op)
    else:
        y = F.conv2d(x, w, None, stride=2, padding=1)
    Ho = y.shape[-1]
    dy = bf(torch.randn(B, cout, Ho, Ho, generator=g)).to(DEV)
    (dw_want,) = torch.autograd.grad(y, w, dy)
    geom = [transposed, k, 2, 1, op, cin, cout, H, H]
    xn = x.permute(0, 2, 3, 1).contiguous()
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    dw = torch.zeros(*wshape, device=DEV)
    ops.conv_wgrad(geom, B, xn, nhwc_strides(xn), None, None, False, dyn, nhwc_strides(dyn), dw)
    assert rel_err(dw, dw_want) < 2e-5, rel_err(dw, dw_want)
    # strided (NCHW) operands + accumulate-into semantics
    dw2 = torch.ones(*wshape, device=DEV)
    ops.conv_wgrad(geom, B, x, [x.stride(0), x.stride(2), x.stride(3), x.stride(1)], None, None, False, dy,
                   [dy.stride(0), dy.stride(2), dy.stride(3), dy.stride(1)], dw2)
    assert rel_err(dw2 - 1.0, dw_want) < 5e-5


@pytest.mark.parametrize("k,cin,H,B", [(3, 3, 28, 16), (3, 1, 28, 8), (4, 3, 64, 4), (3, 3, 26, 5), (4, 2, 18, 3)])   # last two: odd output width
def test_direct_first_conv(k, cin, H, B):
    from clear_vae_b200 import _ops
    ops = _ops.ops()
    g = torch.Generator().manual_seed(k + cin)
    x = torch.rand(B, cin, H, H, generator=g).to(DEV)
    w = (torch.randn(32, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(DEV)
    b = torch.randn(32, generator=g).to(DEV)
    want = F.conv2d(x, w, b, stride=2, padding=1)
    Ho = want.shape[-1]
    for dt in (torch.float32, torch.bfloat16):
        dst = torch.empty(B, Ho, Ho, 32, device=DEV, dtype=dt)
        stats = torch.zeros(64, dtype=torch.float64, device=DEV)
        ok = ops.conv_direct_fwd([0, k, 2, 1, 0, cin, 32, H, H], B, x, [x.stride(0), x.stride(2), x.stride(3), x.stride(1)], None, None,
                                 False, w, b, dst, nhwc_strides(dst), stats)
        assert ok
        assert rel_err(dst.float().permute(0, 3, 1, 2), want) < (1e-5 if dt == torch.float32 else 5e-3)
        assert torch.allclose(stats[:32].float(), want.sum((0, 2, 3)), rtol=1e-4, atol=1e-2)
        assert torch.allclose(stats[32:].float(), (want * want).sum((0, 2, 3)), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("k,op,cout,H,B", [(3, 1, 3, 14, 8), (3, 1, 1, 14, 8), (4, 0, 3, 32, 4)])
def test_direct_last_conv_transpose(k, op, cout, H, B):
    from clear_vae_b200 import _ops
    ops = _ops.ops()
    g = torch.Generator().manual_seed(k + cout)
    raw = torch.randn(B, H, H, 32, generator=g).to(DEV)
    scale = (torch.rand(32, generator=g) + 0.5).to(DEV)
    shift = (torch.randn(32, generator=g) * 0.3).to(DEV)
    w = (torch.randn(32, cout, k, k, generator=g) / (32 * k * k) ** 0.5).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    for dt in (torch.float32, torch.bfloat16):
        r = raw.to(dt)
        act = torch.relu(r.float() * scale + shift)
        want = F.conv_transpose2d(act.permute(0, 3, 1, 2), w, b, stride=2, padding=1, output_padding=op)
        Ho = want.shape[-1]
        dst = torch.empty(B, cout, Ho, Ho, device=DEV)
        stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
        ok = ops.conv_direct_fwd([1, k, 2, 1, op, 32, cout, H, H], B, r, nhwc_strides(r), scale, shift, True, w, b, dst,
                                 [dst.stride(0), dst.stride(2), dst.stride(3), dst.stride(1)], stats)
        assert ok
        assert rel_err(dst, want) < 1e-5
        assert torch.allclose(stats[:cout].float(), want.sum((0, 2, 3)), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("k,op,cout,H,B", [(3, 1, 3, 14, 8), (3, 0, 3, 7, 5), (3, 1, 1, 14, 9), (4, 0, 3, 32, 3), (4, 0, 3, 16, 2), (3, 1, 4, 9, 3)])
def test_direct_last_conv_transpose_materialised_input(k, op, cout, H, B):
    """The training form of the last conv-transpose (`convt_last_x2_kernel`: bf16 activation already normalised, rows staged through
    shared memory, two pixel pairs per thread): even / odd output sizes, image borders, 1-4 output channels, ragged batch, and the
    statistics-only call of CLEAR-MIM's inner forwards (empty destination)."""
    from clear_vae_b200 import _ops
    ops = _ops.ops()
    g = torch.Generator().manual_seed(11 * k + cout + H)
    act = torch.relu(torch.randn(B, H, H, 32, generator=g)).to(DEV).to(torch.bfloat16)
    w = (torch.randn(32, cout, k, k, generator=g) / (32 * k * k) ** 0.5).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    want = F.conv_transpose2d(act.float().permute(0, 3, 1, 2), w, b, stride=2, padding=1, output_padding=op)
    Ho = want.shape[-1]
    geom = [1, k, 2, 1, op, 32, cout, H, H]
    dst = torch.full((B, cout, Ho, Ho), float("nan"), device=DEV)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    assert ops.conv_direct_fwd(geom, B, act, nhwc_strides(act), None, None, False, w, b, dst,
                               [dst.stride(0), dst.stride(2), dst.stride(3), dst.stride(1)], stats)
    assert rel_err(dst, want) < 1e-5, rel_err(dst, want)
    assert torch.allclose(stats[:cout].float(), want.sum((0, 2, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(stats[cout:].float(), (want * want).sum((0, 2, 3)), rtol=1e-4, atol=1e-2)
    stats2 = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    empty = torch.empty((0,), dtype=torch.float32, device=DEV)
    assert ops.conv_direct_fwd(geom, B, act, nhwc_strides(act), None, None, False, w, b, empty,
                               [cout * Ho * Ho, Ho, 1, Ho * Ho], stats2)
    assert torch.allclose(stats2, stats, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("transposed,k,op,c,H,B,img_bf16", [(0, 3, 0, 3, 28, 16, False), (0, 3, 0, 1, 28, 5, False), (0, 4, 0, 3, 64, 3, False),
                                                            (1, 3, 1, 3, 14, 16, True), (1, 3, 1, 1, 14, 7, True), (1, 4, 0, 3, 32, 3, True),
                                                            (1, 3, 1, 4, 14, 4, False), (0, 4, 0, 2, 16, 9, True)])
def test_direct_boundary_weight_gradient(transposed, k, op, c, H, B, img_bf16):
    """CUDA-core weight gradient of the two boundary layers: 32-channel side channels-last bf16, <=4-channel side NCHW."""
    from clear_vae_b200 import _ops
    ops = _ops.ops()
    g = torch.Generator().manual_seed(3 * k + c + transposed)
    cin, cout = (32, c) if transposed else (c, 32)
    x = torch.randn(B, cin, H, H, generator=g).to(DEV)
    wshape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    w = torch.randn(*wshape, generator=g).to(DEV).requires_grad_(True)
    fwd = (lambda a: F.conv_transpose2d(a, w, None, stride=2, padding=1, output_padding=op)) if transposed else \
          (lambda a: F.conv2d(a, w, None, stride=2, padding=1))
    Ho = fwd(x).shape[-1]
    dy = torch.randn(B, cout, Ho, Ho, generator=g).to(DEV)
    # the 32-channel side is bf16; the image side is fp32 or bf16
    if transposed:
        x = bf(x)
        dy = bf(dy) if img_bf16 else dy
    else:
        dy = bf(dy)
        x = bf(x) if img_bf16 else x
    (dw_want,) = torch.autograd.grad(fwd(x), w, dy)
    geom = [transposed, k, 2, 1, op, cin, cout, H, H]
    if transposed:
        src = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        src_strides = nhwc_strides(src)
        dyt = dy.to(torch.bfloat16) if img_bf16 else dy
        dy_strides = [dyt.stride(0), dyt.stride(2), dyt.stride(3), dyt.stride(1)]
    else:
        src = x.to(torch.bfloat16) if img_bf16 else x
        src_strides = [src.stride(0), src.stride(2), src.stride(3), src.stride(1)]
        dyt = dy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        dy_strides = nhwc_strides(dyt)
    dw = torch.zeros(*wshape, device=DEV)
    assert ops.conv_direct_wgrad(geom, B, src, src_strides, dyt, dy_strides, dw)
    assert rel_err(dw, dw_want) < 2e-5, rel_err(dw, dw_want)
    # a layer that is not a boundary layer is declined (the caller then uses the tensor-core kernel)
    other = torch.zeros(64, 32, 3, 3, device=DEV)
    a = torch.zeros(2, 14, 14, 32, device=DEV, dtype=torch.bfloat16)
    b2 = torch.zeros(2, 7, 7, 64, device=DEV, dtype=torch.bfloat16)
    assert not ops.conv_direct_wgrad([0, 3, 2, 1, 0, 32, 64, 14, 14], 2, a, nhwc_strides(a), b2, nhwc_strides(b2), other)


@pytest.mark.parametrize("B,K,N", [(1024, 16, 2048), (37, 64, 2048), (5, 8, 130)])
def test_direct_fc_forward(B, K, N):
    """Linear(2D -> 2048) on the fp32 CUDA-core kernel + BatchNorm1d batch moments per column (vae.py:33-34)."""
    from clear_vae_b200 import _ops
    ops = _ops.ops()
    g = torch.Generator().manual_seed(B + K)
    z = torch.randn(B, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    out = torch.empty(B, N, device=DEV)
    stats = torch.zeros(2 * N + 2, dtype=torch.float64, device=DEV)
    assert ops.fc_fwd(z, w, b, out, stats)
    want = (z.double() @ w.double().T + b.double())
    assert rel_err(out.double(), want) < 2e-6
    assert torch.allclose(stats[:N], want.sum(0), rtol=1e-5, atol=1e-4)
    assert torch.allclose(stats[N:2 * N], (want * want).sum(0), rtol=1e-5, atol=1e-4)
    # K > 64 is declined (tensor-core GEMM takes over)
    assert not ops.fc_fwd(torch.zeros(4, 128, device=DEV), torch.zeros(8, 128, device=DEV), None, torch.empty(4, 8, device=DEV), None)


@pytest.mark.parametrize("layout,C,HW,B", [(0, 64, 49, 8), (0, 32, 196, 3), (1, 128, 16, 6), (1, 512, 4, 5), (2, 2048, 16, 64), (2, 2048, 4, 7),
                                           (3, 128, 16, 9), (3, 512, 4, 4)])
def test_bn_finalize_apply_layouts(layout, C, HW, B):
    """statistics -> (scale, shift, mean, invstd, running estimates) + relu(bn(raw)) in one launch, all four layouts,
    against nn.BatchNorm in training mode (vae.py:17-18,34-35)."""
    from clear_vae_b200 import _ops
    ops = _ops.ops()
    g = torch.Generator().manual_seed(layout * 100 + C + HW)
    if layout == 0:      # channels-last [B, HW, C]
        raw = (torch.randn(B, HW, C, generator=g) * 2 + 0.5).to(DEV).to(torch.bfloat16)
        x_ncl = raw.float().permute(0, 2, 1)                     # [B, C, HW]
    elif layout == 1:    # channel-major [B, C*HW]
        raw = (torch.randn(B, C * HW, generator=g) * 2 + 0.5).to(DEV).to(torch.bfloat16)
        x_ncl = raw.float().view(B, C, HW)
    elif layout == 2:    # fc block: fp32 [B, C] (BatchNorm1d over C features), output permuted to [B, HW, C0]
        raw = (torch.randn(B, C, generator=g) * 2 + 0.5).to(DEV)
        x_ncl = raw.view(B, C, 1)
    else:                # channels-last in, channel-major copies out
        raw = (torch.randn(B, HW, C, generator=g) * 2 + 0.5).to(DEV).to(torch.bfloat16)
        x_ncl = raw.float().permute(0, 2, 1)
    gamma = (torch.rand(C, generator=g) + 0.5).to(DEV)
    beta = (torch.randn(C, generator=g) * 0.3).to(DEV)
    bn = torch.nn.BatchNorm1d(C).to(DEV).train()
    with torch.no_grad():
        bn.weight.copy_(gamma); bn.bias.copy_(beta)
        bn.running_mean.normal_(generator=None); bn.running_var.uniform_(0.5, 2.0)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    want = torch.relu(bn(x_ncl))                                  # also updates bn.running_*
    count = x_ncl.numel() // C
    stats = torch.zeros(2 * C + 2, dtype=torch.float64, device=DEV)
    stats[:C] = x_ncl.double().sum((0, 2))
    stats[C:2 * C] = (x_ncl.double() ** 2).sum((0, 2))
    expand = HW if layout in (1, 3) else 1
    act, scale, shift, mean, invstd, raw_cm = ops.bn_finalize_apply(stats, C, float(count), gamma, beta, rm, rv, 0.1, 1e-5, expand, 1, raw,
                                                                    layout, HW if layout != 0 else 1)
    if layout == 0:
        got = act.float().permute(0, 2, 1)
    elif layout == 1:
        got = act.float().view(B, C, HW)
    elif layout == 2:
        C0 = C // HW
        got = act.float().view(B, HW, C0).permute(0, 2, 1).reshape(B, C, 1)   # back to the (c0, hw) feature order
    else:
        got = act.float().view(B, C, HW)
        assert torch.equal(raw_cm.view(B, C, HW), raw.permute(0, 2, 1))
    assert float((got - want).abs().max()) <= 8e-3 * float(want.abs().max()) + 1e-3          # bf16 output
    assert torch.allclose(mean, x_ncl.mean((0, 2)), rtol=1e-5, atol=1e-5)
    assert torch.allclose(invstd, torch.rsqrt(x_ncl.var((0, 2), unbiased=False) + 1e-5), rtol=1e-5)
    assert torch.allclose(rm, bn.running_mean, rtol=1e-5, atol=1e-6) and torch.allclose(rv, bn.running_var, rtol=1e-5, atol=1e-6)
    assert scale.numel() == C * expand and torch.allclose(scale.view(C, expand)[:, 0], gamma * invstd, rtol=1e-6)
    assert float(stats.abs().max()) == 0.0                       # accumulator and ticket cleared for the next step


@pytest.mark.parametrize("k,op,cout,H,B,dy_bf16", [(3, 1, 3, 14, 8, True), (3, 1, 1, 14, 5, False), (4, 0, 3, 32, 3, True)])
def test_direct_last_layer_data_gradient(k, op, cout, H, B, dy_bf16):
    """dgrad of ConvTranspose2d(32 -> C<=4) == Conv2d(C -> 32) of dy, with the ReLU mask and BN-backward sums fused."""
    from clear_vae_b200 import _ops
    ops = _ops.ops()
    g = torch.Generator().manual_seed(k + cout + H)
    w = (torch.randn(32, cout, k, k, generator=g) / (cout * k * k) ** 0.5).to(DEV)
    x = torch.randn(B, 32, H, H, generator=g).to(DEV).requires_grad_(True)
    y = F.conv_transpose2d(x, w, None, stride=2, padding=1, output_padding=op)
    Ho = y.shape[-1]
    dy = torch.randn(B, cout, Ho, Ho, generator=g).to(DEV)
    dy = bf(dy) if dy_bf16 else dy
    (dx_want,) = torch.autograd.grad(y, x, dy)
    raw = torch.randn(B, H, H, 32, generator=g).to(DEV).to(torch.bfloat16)     # previous block's raw output (mask source)
    sc = (torch.rand(32, generator=g) + 0.5).to(DEV)
    sh = (torch.randn(32, generator=g) * 0.3).to(DEV)
    mask = (raw.float() * sc + sh) > 0
    want = dx_want.permute(0, 2, 3, 1) * mask
    dyt = dy.to(torch.bfloat16) if dy_bf16 else dy
    dst = torch.empty(B, H, H, 32, device=DEV)
    stats = torch.zeros(66, dtype=torch.float64, device=DEV)
    assert ops.conv_direct_dgrad([1, k, 2, 1, op, 32, cout, H, H], B, dyt, [dyt.stride(0), dyt.stride(2), dyt.stride(3), dyt.stride(1)], w, dst,
                                 nhwc_strides(dst), raw, nhwc_strides(raw), sc, sh, stats)
    assert rel_err(dst, want) < 1e-5
    assert torch.allclose(stats[:32].float(), want.sum((0, 1, 2)), rtol=1e-4, atol=1e-3)
    assert torch.allclose(stats[32:64].float(), (want * raw.float()).sum((0, 1, 2)), rtol=1e-4, atol=1e-3)


# ------------------------------------------------------------------------------------------------------------------
# fp32-grade mode (CLEARVAE_ROLE_SPLIT3): fp32 operands, hi*hi + lo*hi + hi*lo on the bf16 tensor cores.
# Reference = torch fp32 convolutions with TF32 off on the UNROUNDED operands; gate 2e-5 of the output scale (the plain
# bf16 path needs bf16-rounded operands to meet the same gate — on unrounded operands it sits at ~3e-3).
# ------------------------------------------------------------------------------------------------------------------
SPLIT3 = 16


@pytest.mark.parametrize("transposed,k,op,cin,cout,H,B", [(0, 3, 0, 32, 64, 14, 16), (0, 4, 0, 64, 128, 16, 8), (0, 4, 0, 3, 32, 16, 4),
                                                           (1, 3, 1, 64, 32, 7, 8), (1, 4, 0, 256, 128, 4, 16), (1, 3, 1, 32, 3, 14, 8)])
def test_split3_gemm_and_wgrad_are_fp32_grade(transposed, k, op, cin, cout, H, B):
    from clear_vae_b200 import _ops
    ops = _ops.ops()
    g = torch.Generator().manual_seed(7 * k + cin + cout + transposed)
    x = torch.randn(B, cin, H, H, generator=g).to(DEV)
    scale = (torch.rand(cin, generator=g) + 0.5).to(DEV)
    shift = (torch.randn(cin, generator=g) * 0.3).to(DEV)
    if transposed:
        w = (torch.randn(cin, cout, k, k, generator=g) / (cin * k * k) ** 0.5).to(DEV)
    else:
        w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    act = torch.relu(x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    conv = (lambda a: F.conv_transpose2d(a, w, b, stride=2, padding=1, output_padding=op)) if transposed else \
        (lambda a: F.conv2d(a, w, b, stride=2, padding=1))
    act_r = act.detach().requires_grad_(True)
    w.requires_grad_(True)
    want = conv(act_r)
    Ho = want.shape[-1]
    geom = [transposed, k, 2, 1, op, cin, cout, H, H]
    xn = x.permute(0, 2, 3, 1).contiguous()
    # forward with the BatchNorm-apply + ReLU pre-op on the fp32 operand, fp32 channels-last destination + statistics
    dst = torch.empty(B, Ho, Ho, cout, device=DEV)
    stats = torch.zeros(2 * cout + 2, dtype=torch.float64, device=DEV)
    run_gemm(geom, 0 | SPLIT3, B, xn, nhwc_strides(xn), w.detach(), b, dst, nhwc_strides(dst), pre=(scale, shift), relu=True, stats=stats)
    e = rel_err(dst.permute(0, 3, 1, 2), want.detach())
    assert e < 2e-5, e
    assert torch.allclose(stats[:cout], want.detach().double().sum((0, 2, 3)), rtol=1e-5, atol=2e-5 * float(want.detach().abs().sum((0, 2, 3)).max()))
    # the plain bf16 path on the same unrounded operands is two orders of magnitude away: the split is what buys the accuracy
    dst_b = torch.empty(B, Ho, Ho, cout, device=DEV)
    run_gemm(geom, 0, B, xn, nhwc_strides(xn), w.detach(), b, dst_b, nhwc_strides(dst_b), pre=(scale, shift), relu=True)
    assert rel_err(dst_b.permute(0, 3, 1, 2), want.detach()) > 20 * e
    # data gradient and weight gradient
    dy = torch.randn(B, cout, Ho, Ho, generator=g).to(DEV)
    dx_want, dw_want = torch.autograd.grad(want, [act_r, w], dy)
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    dx = torch.empty(B, H, H, cin, device=DEV)
    run_gemm(geom, 1 | SPLIT3, B, dyn, nhwc_strides(dyn), w.detach(), None, dx, nhwc_strides(dx))
    assert rel_err(dx.permute(0, 3, 1, 2), dx_want) < 2e-5
    dw = torch.zeros_like(w)
    ops.conv_wgrad(geom, B, xn, nhwc_strides(xn), scale, shift, True, dyn, nhwc_strides(dyn), dw, True)
    assert rel_err(dw, dw_want) < 2e-5, rel_err(dw, dw_want)


@pytest.mark.parametrize("B,K,N", [(300, 2048, 32), (128, 16, 2048), (1000, 2048, 128)])
def test_split3_linear(B, K, N):
    g = torch.Generator().manual_seed(B + K + N)
    x = torch.randn(B, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    geom = [0, 1, 1, 0, 0, K, N, 1, 1]
    dst = torch.empty(B, N, device=DEV)
    run_gemm(geom, 0 | SPLIT3, B, x, (K, 0, 0, 1), w, b, dst, (N, 0, 0, 1))
    want = (x.double() @ w.double().T + b.double()).float()
    assert rel_err(dst, want) < 2e-5, rel_err(dst, want)


# ------------------------------------------------------------------------------------------------------------------
# TMA-operand persistent kernel (bf16 channels-last operand, channels-last destination): every interior layer geometry of
# VAE / VAE64 in both roles, batch sizes that leave partial image groups / partial row tiles, with the statistics and the
# masked (ReLU + BatchNorm-backward sums) epilogues.  Reference: torch fp32 convolutions on the same bf16-valued operands.
# ------------------------------------------------------------------------------------------------------------------
_INTERIOR = [  # transposed, k, op, cin, cout, Hin
    (0, 3, 0, 32, 64, 14), (0, 3, 0, 64, 128, 7), (1, 3, 0, 128, 64, 4), (1, 3, 1, 64, 32, 7),
    (0, 4, 0, 32, 64, 32), (0, 4, 0, 64, 128, 16), (0, 4, 0, 128, 256, 8), (0, 4, 0, 256, 512, 4),
    (1, 4, 0, 512, 256, 2), (1, 4, 0, 256, 128, 4), (1, 4, 0, 128, 64, 8), (1, 4, 0, 64, 32, 16),
]


@pytest.mark.parametrize("transposed,k,op,cin,cout,H", _INTERIOR)
@pytest.mark.parametrize("B", [5, 37])
def test_tma_operand_kernel_matches_torch(transposed, k, op, cin, cout, H, B):
    g = torch.Generator().manual_seed(11 * k + cin + cout + H + B)
    x = bf(torch.randn(B, cin, H, H, generator=g)).to(DEV)
    if transposed:
        w = bf(torch.randn(cin, cout, k, k, generator=g) / (cin * k * k) ** 0.5).to(DEV)
        fwd = lambda a: F.conv_transpose2d(a, w, b, stride=2, padding=1, output_padding=op)
    else:
        w = bf(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(DEV)
        fwd = lambda a: F.conv2d(a, w, b, stride=2, padding=1)
    b = torch.randn(cout, generator=g).to(DEV)
    xr = x.clone().requires_grad_(True)
    want = fwd(xr)
    Ho = want.shape[-1]
    geom = [transposed, k, 2, 1, op, cin, cout, H, H]
    xn = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)          # exact: values are bf16 already
    dst = torch.empty(B, Ho, Ho, cout, device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros(2 * cout + 2, dtype=torch.float64, device=DEV)
    run_gemm(geom, 0, B, xn, nhwc_strides(xn), w, b, dst, nhwc_strides(dst), stats=stats)
    wn = want.detach()
    assert rel_err(dst.float().permute(0, 3, 1, 2), wn) < 6e-3                       # bf16 output rounding
    assert torch.allclose(stats[:cout].float(), wn.sum((0, 2, 3)), rtol=1e-4, atol=2e-5 * float(wn.abs().sum((0, 2, 3)).max()))
    assert torch.allclose(stats[cout:2 * cout].float(), (wn * wn).sum((0, 2, 3)), rtol=1e-4, atol=1e-2)
    # fp32 destination: the accumulators themselves
    dst32 = torch.empty(B, Ho, Ho, cout, device=DEV)
    run_gemm(geom, 0, B, xn, nhwc_strides(xn), w, b, dst32, nhwc_strides(dst32))
    assert rel_err(dst32.permute(0, 3, 1, 2), wn) < 2e-5
    # data gradient with the previous block's ReLU mask + BatchNorm-backward sums in the epilogue
    dy = bf(torch.randn(B, cout, Ho, Ho, generator=g)).to(DEV)
    (dx_want,) = torch.autograd.grad(want, xr, dy)
    dyn = dy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    raw_prev = bf(torch.randn(B, H, H, cin, generator=g)).to(DEV).to(torch.bfloat16)   # previous block's raw output (mask source)
    ms = (torch.rand(cin, generator=g) + 0.5).to(DEV)
    mh = (torch.randn(cin, generator=g) * 0.3).to(DEV)
    keep = (raw_prev.float() * ms + mh) > 0
    gdst = torch.empty(B, H, H, cin, device=DEV)
    st2 = torch.zeros(2 * cin + 2, dtype=torch.float64, device=DEV)
    run_gemm(geom, 1, B, dyn, nhwc_strides(dyn), w, None, gdst, nhwc_strides(gdst), epi=1, mask=raw_prev,
             mask_strides=nhwc_strides(raw_prev), mscale=ms, mshift=mh, stats=st2)
    gw = dx_want.permute(0, 2, 3, 1) * keep
    assert rel_err(gdst, gw) < 2e-5
    assert torch.allclose(st2[:cin].float(), gw.sum((0, 1, 2)), rtol=1e-4, atol=2e-5 * float(gw.abs().sum((0, 1, 2)).max()))
    assert torch.allclose(st2[cin:2 * cin].float(), (gw * raw_prev.float()).sum((0, 1, 2)), rtol=1e-4,
                          atol=2e-5 * float((gw * raw_prev.float()).abs().sum((0, 1, 2)).max()))
