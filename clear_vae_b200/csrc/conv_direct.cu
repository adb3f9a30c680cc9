// Direct (CUDA-core) forward kernels for the two boundary layers of the VAE stacks, whose GEMM shapes do not
// suit 128-row tensor-core tiles ("Layer-1 convs have K=27/48 ... the last convT has N*k^2=27/48"):
//
//   conv_first : Conv2d(Cin<=4 -> 32, k in {3,4}, stride 2, pad 1)   x NCHW fp32  ->  raw NHWC (bf16 | fp32)
//                (vae.py:16,114)   one thread per output pixel, 32 accumulators, weights broadcast from smem
//   convt_last : ConvTranspose2d(32 -> Cout<=4, k in {3,4}, stride 2, pad 1, out_pad)  raw NHWC + BN/ReLU pre-op
//                (vae.py:43,153)   -> raw NCHW fp32; one thread per output pixel, only the taps of its parity
//
// Both are HBM/L2-streaming kernels (FLOPs are negligible): 16-byte accesses, grid = a multiple of the SM
// count with a grid-stride loop, per-thread BatchNorm statistics in registers, one warp-shuffle + shared
// reduction per CTA and one fp64 atomic per channel per CTA.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int kNT = 128;

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---------------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(kNT) conv_first_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, void* __restrict__ out, int out_bf16,
                                                         double* stats, int B, int Cin, int H, int W, int Ho, int Wo) {
  constexpr int CO = 32;
  __shared__ __align__(16) float sW[4 * K * K * CO];  // [ci*K*K + kh*K + kw][co]
  __shared__ float sRed[2][kNT / 32][CO];
  const int KK = Cin * K * K;
  for (int i = threadIdx.x; i < KK * CO; i += kNT) {
    const int k = i / CO, co = i % CO;  // reference layout w[co][ci][kh][kw] = w[co * KK + k]
    sW[i] = w[co * KK + k];
  }
  __syncthreads();
  float s[CO], q[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) s[c] = q[c] = 0.f;
  const long long npix = (long long)B * Ho * Wo;
  for (long long pix = (long long)blockIdx.x * kNT + threadIdx.x; pix < npix; pix += (long long)gridDim.x * kNT) {
    const int ow = (int)(pix % Wo), oh = (int)((pix / Wo) % Ho);
    const long long n = pix / ((long long)Wo * Ho);
    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = bias ? __ldg(bias + c) : 0.f;
    for (int ci = 0; ci < Cin; ++ci) {
      const float* xp = x + (n * Cin + ci) * (long long)H * W;
#pragma unroll
      for (int kh = 0; kh < K; ++kh) {
        const int ih = oh * 2 - 1 + kh;
#pragma unroll
        for (int kw = 0; kw < K; ++kw) {
          const int iw = ow * 2 - 1 + kw;
          const bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
          const float xv = ok ? __ldg(xp + (long long)ih * W + iw) : 0.f;
          const float4* wr = reinterpret_cast<const float4*>(sW + ((ci * K + kh) * K + kw) * CO);
#pragma unroll
          for (int c4 = 0; c4 < CO / 4; ++c4) {
            const float4 wv = wr[c4];
            acc[4 * c4] = fmaf(xv, wv.x, acc[4 * c4]);
            acc[4 * c4 + 1] = fmaf(xv, wv.y, acc[4 * c4 + 1]);
            acc[4 * c4 + 2] = fmaf(xv, wv.z, acc[4 * c4 + 2]);
            acc[4 * c4 + 3] = fmaf(xv, wv.w, acc[4 * c4 + 3]);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CO; ++c) { s[c] += acc[c]; q[c] = fmaf(acc[c], acc[c], q[c]); }
    if (out_bf16) {
      uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + pix * CO);
#pragma unroll
      for (int c = 0; c < CO; c += 8)
        o[c / 8] = make_uint4(pack2(acc[c], acc[c + 1]), pack2(acc[c + 2], acc[c + 3]), pack2(acc[c + 4], acc[c + 5]),
                              pack2(acc[c + 6], acc[c + 7]));
    } else {
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + pix * CO);
#pragma unroll
      for (int c = 0; c < CO; c += 4) o[c / 4] = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
    }
  }
  if (stats == nullptr) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < CO; ++c) {
    const float a = cv::warp_sum(s[c]), b = cv::warp_sum(q[c]);
    if (lane == 0) { sRed[0][warp][c] = a; sRed[1][warp][c] = b; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * CO) {
    const int which = threadIdx.x / CO, c = threadIdx.x % CO;
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < kNT / 32; ++wv) t += sRed[which][wv][c];
    atomicAdd(stats + which * CO + c, (double)t);
  }
}

// ---------------------------------------------------------------------------------------------------------
// One thread = two horizontally adjacent output pixels (ow = 2j, 2j+1) of one output row; threads are ordered so that
// a warp stays inside one row parity (uniform set of valid kh) — no divergence, weights are warp-broadcast from
// shared memory, every needed input pixel is loaded (and BatchNorm+ReLU'ed) exactly once per thread.
template <int K, bool SRC_BF16>
__global__ void __launch_bounds__(kNT) convt_last_kernel(const void* __restrict__ src, const float* __restrict__ pre_scale,
                                                         const float* __restrict__ pre_shift, int pre_relu,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ out, double* stats, int B, int Cout, int Hi, int Wi,
                                                         int Ho, int Wo) {
  constexpr int CI = 32;
  __shared__ __align__(16) float sW[K * K * CI * 4];  // [kh*K+kw][ci][co padded to 4]
  __shared__ float sScale[CI], sShift[CI];
  __shared__ float sRed[2][kNT / 32][4];
  for (int i = threadIdx.x; i < K * K * CI * 4; i += kNT) {
    const int co = i & 3, ci = (i >> 2) % CI, t = i / (4 * CI);
    sW[i] = co < Cout ? w[(ci * Cout + co) * K * K + t] : 0.f;  // reference layout w[ci][co][kh][kw]
  }
  if (threadIdx.x < CI) {
    sScale[threadIdx.x] = pre_scale ? pre_scale[threadIdx.x] : 1.f;
    sShift[threadIdx.x] = pre_shift ? pre_shift[threadIdx.x] : 0.f;
  }
  __syncthreads();
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  const int W2 = (Wo + 1) / 2, H2 = (Ho + 1) / 2;
  const long long nwork = (long long)B * 2 * H2 * W2;
  for (long long p = (long long)blockIdx.x * kNT + threadIdx.x; p < nwork; p += (long long)gridDim.x * kNT) {
    const int j = (int)(p % W2);
    long long t = p / W2;
    const int oh2 = (int)(t % H2);
    t /= H2;
    const int a = (int)(t & 1);
    const long long n = t >> 1;
    const int oh = 2 * oh2 + a;
    if (oh >= Ho) continue;
    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kh = 0; kh < K; ++kh) {
      const int th = oh + 1 - kh;  // = 2 * ih
      if (th < 0 || (th & 1) || (th >> 1) >= Hi) continue;   // warp-uniform (same row parity)
      const int ih = th >> 1;
#pragma unroll
      for (int d = -1; d <= 1; ++d) {
        // input column iw = j + d feeds ow = 2j through kw0 = 1 - 2d and ow = 2j + 1 through kw1 = 2 - 2d
        constexpr int kDummy = 0;
        (void)kDummy;
        const int kw0 = 1 - 2 * d, kw1 = 2 - 2 * d;
        const bool use0 = kw0 >= 0 && kw0 < K, use1 = kw1 >= 0 && kw1 < K;
        if (!use0 && !use1) continue;                         // compile-time
        const int iw = j + d;
        if (iw < 0 || iw >= Wi) continue;
        const long long off = ((n * Hi + ih) * Wi + iw) * CI;
        float v[CI];
        if (SRC_BF16) {
          const uint4* ptr = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(src) + off);
#pragma unroll
          for (int c8 = 0; c8 < CI / 8; ++c8) {
            const uint4 u = __ldg(ptr + c8);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h2[i]); v[c8 * 8 + 2 * i] = f.x; v[c8 * 8 + 2 * i + 1] = f.y; }
          }
        } else {
          const float4* ptr = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + off);
#pragma unroll
          for (int c4 = 0; c4 < CI / 4; ++c4) { const float4 u = __ldg(ptr + c4); v[4 * c4] = u.x; v[4 * c4 + 1] = u.y; v[4 * c4 + 2] = u.z; v[4 * c4 + 3] = u.w; }
        }
        const float4* w0 = reinterpret_cast<const float4*>(sW + (kh * K + (use0 ? kw0 : 0)) * CI * 4);
        const float4* w1 = reinterpret_cast<const float4*>(sW + (kh * K + (use1 ? kw1 : 0)) * CI * 4);
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) {
          float x = fmaf(v[ci], sScale[ci], sShift[ci]);
          if (pre_relu) x = fmaxf(x, 0.f);
          if (use0) {
            const float4 wv = w0[ci];
            acc0[0] = fmaf(x, wv.x, acc0[0]); acc0[1] = fmaf(x, wv.y, acc0[1]); acc0[2] = fmaf(x, wv.z, acc0[2]); acc0[3] = fmaf(x, wv.w, acc0[3]);
          }
          if (use1) {
            const float4 wv = w1[ci];
            acc1[0] = fmaf(x, wv.x, acc1[0]); acc1[1] = fmaf(x, wv.y, acc1[1]); acc1[2] = fmaf(x, wv.z, acc1[2]); acc1[3] = fmaf(x, wv.w, acc1[3]);
          }
        }
      }
    }
    const int ow = 2 * j;
    const bool two = ow + 1 < Wo;
#pragma unroll
    for (int co = 0; co < 4; ++co) {
      if (co < Cout) {
        const float bsv = bias ? __ldg(bias + co) : 0.f;
        const float y0 = acc0[co] + bsv, y1 = acc1[co] + bsv;
        float* o = out + ((n * Cout + co) * Ho + oh) * (long long)Wo + ow;
        if (two && ((reinterpret_cast<uintptr_t>(o) & 7) == 0)) {
          *reinterpret_cast<float2*>(o) = make_float2(y0, y1);
        } else {
          o[0] = y0;
          if (two) o[1] = y1;
        }
        s[co] += y0; q[co] = fmaf(y0, y0, q[co]);
        if (two) { s[co] += y1; q[co] = fmaf(y1, y1, q[co]); }
      }
    }
  }
  if (stats == nullptr) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float a = cv::warp_sum(s[c]), b = cv::warp_sum(q[c]);
    if (lane == 0) { sRed[0][warp][c] = a; sRed[1][warp][c] = b; }
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int which = threadIdx.x >> 2, c = threadIdx.x & 3;
    if (c < Cout) {
      float t = 0.f;
#pragma unroll
      for (int wv = 0; wv < kNT / 32; ++wv) t += sRed[which][wv][c];
      atomicAdd(stats + which * Cout + c, (double)t);
    }
  }
}

inline int grid_for(long long npix) {
  long long g = (npix + kNT - 1) / kNT;
  const long long cap = 148 * 8;
  return (int)(g < 1 ? 1 : g > cap ? cap : g);
}

}  // namespace

extern "C" {

int clearvae_conv_direct_supported(const clearvae_conv_geom* g, const clearvae_tensor4* src, const clearvae_tensor4* dst) {
  if (!g || !src || !dst || g->stride != 2 || g->pad != 1 || (g->k != 3 && g->k != 4) || g->Hin != g->Win) return 0;
  if (!g->transposed) {
    // x: NCHW fp32 contiguous; dst: NHWC contiguous
    const int Ho = (g->Hin + 2 - g->k) / 2 + 1;
    return g->Cin <= 4 && g->Cout == 32 && src->dtype == CLEARVAE_F32 && src->sw == 1 && src->sh == g->Win &&
           src->sc == (int64_t)g->Hin * g->Win && src->sn == (int64_t)g->Cin * g->Hin * g->Win && dst->sc == 1 && dst->sw == 32 &&
           dst->sh == (int64_t)Ho * 32 && dst->sn == (int64_t)Ho * Ho * 32;
  }
  const int Ho = (g->Hin - 1) * 2 - 2 + g->k + g->out_pad;
  return g->Cin == 32 && g->Cout <= 4 && src->sc == 1 && src->sw == 32 && src->sh == (int64_t)g->Win * 32 &&
         src->sn == (int64_t)g->Hin * g->Win * 32 && dst->dtype == CLEARVAE_F32 && dst->sw == 1 && dst->sh == Ho &&
         dst->sc == (int64_t)Ho * Ho && dst->sn == (int64_t)g->Cout * Ho * Ho;
}

int clearvae_conv_direct_fwd(const clearvae_conv_geom* g, int64_t batch, const clearvae_tensor4* src, const float* pre_scale,
                             const float* pre_shift, int32_t pre_relu, const float* weight, const float* bias,
                             const clearvae_tensor4* dst, double* stats, void* stream) {
  if (!g || !src || !src->ptr || !dst || !dst->ptr || !weight || batch <= 0) return CLEARVAE_EINVAL;
  if (!clearvae_conv_direct_supported(g, src, dst)) return CLEARVAE_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (!g->transposed) {
    if (pre_scale || pre_relu) return CLEARVAE_EUNSUPPORTED;
    const int Ho = (g->Hin + 2 - g->k) / 2 + 1;
    const long long npix = batch * Ho * Ho;
    if (g->k == 3)
      conv_first_kernel<3><<<grid_for(npix), kNT, 0, st>>>((const float*)src->ptr, weight, bias, dst->ptr, dst->dtype == CLEARVAE_BF16,
                                                          stats, (int)batch, g->Cin, g->Hin, g->Win, Ho, Ho);
    else
      conv_first_kernel<4><<<grid_for(npix), kNT, 0, st>>>((const float*)src->ptr, weight, bias, dst->ptr, dst->dtype == CLEARVAE_BF16,
                                                          stats, (int)batch, g->Cin, g->Hin, g->Win, Ho, Ho);
  } else {
    const int Ho = (g->Hin - 1) * 2 - 2 + g->k + g->out_pad;
    const long long npix = batch * 2 * ((Ho + 1) / 2) * ((Ho + 1) / 2);  // one thread per output-pixel pair
    const bool bf = src->dtype == CLEARVAE_BF16;
#define CV_LAUNCH_T(KK, BF)                                                                                                   \
  convt_last_kernel<KK, BF><<<grid_for(npix), kNT, 0, st>>>(src->ptr, pre_scale, pre_shift, pre_relu, weight, bias,           \
                                                            (float*)dst->ptr, stats, (int)batch, g->Cout, g->Hin, g->Win, Ho, Ho)
    if (g->k == 3) { if (bf) CV_LAUNCH_T(3, true); else CV_LAUNCH_T(3, false); }
    else { if (bf) CV_LAUNCH_T(4, true); else CV_LAUNCH_T(4, false); }
#undef CV_LAUNCH_T
  }
  CV_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
