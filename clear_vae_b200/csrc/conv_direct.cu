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
// MASKED = data-gradient use (see clearvae_conv_direct_dgrad): no bias; the epilogue applies the previous block's
// ReLU mask and accumulates the BatchNorm-backward sums (sum g, sum g*y) instead of (sum y, sum y^2).
template <int K, bool X_BF16, bool MASKED>
__global__ void __launch_bounds__(kNT) conv_first_kernel(const void* __restrict__ xv_, const float* __restrict__ w,
                                                         const float* __restrict__ bias, void* __restrict__ out, int out_bf16,
                                                         double* stats, int B, int Cin, int H, int W, int Ho, int Wo,
                                                         const void* __restrict__ msk, int msk_bf16,
                                                         const float* __restrict__ msk_scale, const float* __restrict__ msk_shift) {
  constexpr int CO = 32;
  __shared__ __align__(16) float sW[4 * K * K * CO];  // [ci*K*K + kh*K + kw][co]
  __shared__ float sRed[2][kNT / 32][CO];
  __shared__ float sMs[2][CO];
  // per-channel statistics: each thread parks its 32 outputs in a padded tile and thread (column, row quarter) adds 32
  // rows — two running registers per thread instead of 64, which is what lets 4-5 CTAs share an SM
  __shared__ float sT[MASKED ? 2 : 1][kNT][CO + 1];
  const int KK = Cin * K * K;
  for (int i = threadIdx.x; i < KK * CO; i += kNT) {
    const int k = i / CO, co = i % CO;  // reference layout w[co][ci][kh][kw] = w[co * KK + k]
    sW[i] = w[co * KK + k];
  }
  if (MASKED && threadIdx.x < CO) {
    sMs[0][threadIdx.x] = msk_scale ? msk_scale[threadIdx.x] : 1.f;
    sMs[1][threadIdx.x] = msk_shift ? msk_shift[threadIdx.x] : 0.f;
  }
  __syncthreads();
  float s_run = 0.f, q_run = 0.f;   // running sums of column (threadIdx.x & 31) over the row quarter (threadIdx.x >> 5)
  const long long npix = (long long)B * Ho * Wo;
  const long long npix_pad = (npix + kNT - 1) / kNT * kNT;   // whole CTA iterates together (barriers inside the loop)
  for (long long pix0 = (long long)blockIdx.x * kNT + threadIdx.x; pix0 < npix_pad; pix0 += (long long)gridDim.x * kNT) {
    const bool pvalid = pix0 < npix;
    const long long pix = pvalid ? pix0 : 0;
    const int ow = (int)(pix % Wo), oh = (int)((pix / Wo) % Ho);
    const long long n = pix / ((long long)Wo * Ho);
    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = (!MASKED && bias) ? __ldg(bias + c) : 0.f;
    for (int ci = 0; ci < Cin; ++ci) {
      const long long xoff = (n * Cin + ci) * (long long)H * W;
#pragma unroll
      for (int kh = 0; kh < K; ++kh) {
        const int ih = oh * 2 - 1 + kh;
#pragma unroll
        for (int kw = 0; kw < K; ++kw) {
          const int iw = ow * 2 - 1 + kw;
          const bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
          float xv = 0.f;
          if (ok) {
            const long long o = xoff + (long long)ih * W + iw;
            xv = X_BF16 ? __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(xv_) + o))
                        : __ldg(reinterpret_cast<const float*>(xv_) + o);
          }
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
    if (MASKED) {
      float y[CO];
      if (msk_bf16) {
        const uint4* mp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(msk) + pix * CO);
#pragma unroll
        for (int i = 0; i < CO / 8; ++i) {
          const uint4 u = __ldg(mp + i);
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h2[k]); y[8 * i + 2 * k] = f.x; y[8 * i + 2 * k + 1] = f.y; }
        }
      } else {
        const float4* mp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(msk) + pix * CO);
#pragma unroll
        for (int i = 0; i < CO / 4; ++i) { const float4 u = __ldg(mp + i); y[4 * i] = u.x; y[4 * i + 1] = u.y; y[4 * i + 2] = u.z; y[4 * i + 3] = u.w; }
      }
#pragma unroll
      for (int c = 0; c < CO; ++c) {
        const float a = (pvalid && fmaf(y[c], sMs[0][c], sMs[1][c]) > 0.f) ? acc[c] : 0.f;
        acc[c] = a;
        if (stats != nullptr) { sT[0][threadIdx.x][c] = a; sT[MASKED ? 1 : 0][threadIdx.x][c] = a * y[c]; }
      }
    } else if (stats != nullptr) {
#pragma unroll
      for (int c = 0; c < CO; ++c) sT[0][threadIdx.x][c] = pvalid ? acc[c] : 0.f;
    }
    if (stats != nullptr) {
      __syncthreads();
      const int col = threadIdx.x & 31, r0 = (threadIdx.x >> 5) * 32;
      float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
      for (int rr = 0; rr < 32; rr += 2) {
        const float v0 = sT[0][r0 + rr][col], v1 = sT[0][r0 + rr + 1][col];
        a0 += v0; a1 += v1;
        if (MASKED) { b0 += sT[MASKED ? 1 : 0][r0 + rr][col]; b1 += sT[MASKED ? 1 : 0][r0 + rr + 1][col]; }
        else { b0 = fmaf(v0, v0, b0); b1 = fmaf(v1, v1, b1); }
      }
      s_run += a0 + a1;
      q_run += b0 + b1;
      __syncthreads();
    }
    if (!pvalid) continue;
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
  sRed[0][threadIdx.x >> 5][threadIdx.x & 31] = s_run;
  sRed[1][threadIdx.x >> 5][threadIdx.x & 31] = q_run;
  __syncthreads();
  if (threadIdx.x < 2 * CO) {
    const int which = threadIdx.x / CO, c = threadIdx.x % CO;
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < kNT / 32; ++wv) t += sRed[which][wv][c];
    atomicAdd(stats + which * CO + c, (double)t);
  }
}

// Same layer, TWO horizontally adjacent output pixels per thread: the eight weight vectors of a tap are loaded once for both
// (LDS : FMA 1:8 instead of 1:4), the FMAs are packed fp32x2, the two pixels share K - 2 of their K + 2 input columns, the thread
// stores 128 contiguous bytes, and the statistics transposition (two tiles: sums and squares / mask products) runs once per pair.
template <int K, bool X_BF16, bool MASKED>
__global__ void __launch_bounds__(kNT) conv_first2_kernel(const void* __restrict__ xv_, const float* __restrict__ w,
                                                          const float* __restrict__ bias, void* __restrict__ out, int out_bf16,
                                                          double* stats, int B, int Cin, int H, int W, int Ho, int Wo,
                                                          const void* __restrict__ msk, int msk_bf16,
                                                          const float* __restrict__ msk_scale, const float* __restrict__ msk_shift) {
  constexpr int CO = 32;
  __shared__ __align__(16) float sW[4 * K * K * CO];  // [ci*K*K + kh*K + kw][co]
  __shared__ float sRed[2][kNT / 32][CO];
  __shared__ float sMs[2][CO];
  __shared__ float sT[2][kNT][CO + 1];
  const int KK = Cin * K * K;
  for (int i = threadIdx.x; i < KK * CO; i += kNT) {
    const int k = i / CO, co = i % CO;  // reference layout w[co][ci][kh][kw] = w[co * KK + k]
    sW[i] = w[co * KK + k];
  }
  if (MASKED && threadIdx.x < CO) {
    sMs[0][threadIdx.x] = msk_scale ? msk_scale[threadIdx.x] : 1.f;
    sMs[1][threadIdx.x] = msk_shift ? msk_shift[threadIdx.x] : 0.f;
  }
  __syncthreads();
  float s_run = 0.f, q_run = 0.f;   // running sums of column (threadIdx.x & 31) over the row quarter (threadIdx.x >> 5)
  const int Wp = (Wo + 1) / 2;
  const long long npair = (long long)B * Ho * Wp;
  const long long npair_pad = (npair + kNT - 1) / kNT * kNT;   // whole CTA iterates together (barriers inside the loop)
  for (long long pr0 = (long long)blockIdx.x * kNT + threadIdx.x; pr0 < npair_pad; pr0 += (long long)gridDim.x * kNT) {
    const bool pvalid = pr0 < npair;
    const long long pr = pvalid ? pr0 : 0;
    const int ow = 2 * (int)(pr % Wp), oh = (int)((pr / Wp) % Ho);
    const long long n = pr / ((long long)Wp * Ho);
    const bool v1ok = pvalid && ow + 1 < Wo;
    const long long pix = (n * Ho + oh) * (long long)Wo + ow;
    float2 acc0[CO / 2], acc1[CO / 2];
#pragma unroll
    for (int c = 0; c < CO / 2; ++c) {
      const float2 b2 = (!MASKED && bias) ? make_float2(__ldg(bias + 2 * c), __ldg(bias + 2 * c + 1)) : make_float2(0.f, 0.f);
      acc0[c] = b2; acc1[c] = b2;
    }
    for (int ci = 0; ci < Cin; ++ci) {
      const long long xoff = (n * Cin + ci) * (long long)H * W;
#pragma unroll
      for (int kh = 0; kh < K; ++kh) {
        const int ih = oh * 2 - 1 + kh;
        const bool rok = ih >= 0 && ih < H;
        float xr[K + 2];
#pragma unroll
        for (int j = 0; j < K + 2; ++j) {
          const int iw = ow * 2 - 1 + j;
          float xv = 0.f;
          if (rok && iw >= 0 && iw < W) {
            const long long o = xoff + (long long)ih * W + iw;
            xv = X_BF16 ? __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(xv_) + o))
                        : __ldg(reinterpret_cast<const float*>(xv_) + o);
          }
          xr[j] = xv;
        }
#pragma unroll
        for (int kw = 0; kw < K; ++kw) {
          const float4* wr = reinterpret_cast<const float4*>(sW + ((ci * K + kh) * K + kw) * CO);
          const float2 x0 = make_float2(xr[kw], xr[kw]), x1 = make_float2(xr[kw + 2], xr[kw + 2]);
#pragma unroll
          for (int c4 = 0; c4 < CO / 4; ++c4) {
            const float4 wv = wr[c4];
            const float2 wa = make_float2(wv.x, wv.y), wb = make_float2(wv.z, wv.w);
            acc0[2 * c4] = __ffma2_rn(x0, wa, acc0[2 * c4]);
            acc0[2 * c4 + 1] = __ffma2_rn(x0, wb, acc0[2 * c4 + 1]);
            acc1[2 * c4] = __ffma2_rn(x1, wa, acc1[2 * c4]);
            acc1[2 * c4 + 1] = __ffma2_rn(x1, wb, acc1[2 * c4 + 1]);
          }
        }
      }
    }
    float* a0 = reinterpret_cast<float*>(acc0);
    float* a1 = reinterpret_cast<float*>(acc1);
    if (MASKED) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float* a = half ? a1 : a0;
        const bool ok = half ? v1ok : pvalid;
        float y[CO];
        const long long mp0 = ok ? (pix + half) * CO : 0;
        if (msk_bf16) {
          const uint4* mp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(msk) + mp0);
#pragma unroll
          for (int i = 0; i < CO / 8; ++i) {
            const uint4 u = __ldg(mp + i);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h2[k]); y[8 * i + 2 * k] = f.x; y[8 * i + 2 * k + 1] = f.y; }
          }
        } else {
          const float4* mp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(msk) + mp0);
#pragma unroll
          for (int i = 0; i < CO / 4; ++i) { const float4 u = __ldg(mp + i); y[4 * i] = u.x; y[4 * i + 1] = u.y; y[4 * i + 2] = u.z; y[4 * i + 3] = u.w; }
        }
#pragma unroll
        for (int c = 0; c < CO; ++c) {
          const float v = (ok && fmaf(y[c], sMs[0][c], sMs[1][c]) > 0.f) ? a[c] : 0.f;
          a[c] = v;
          if (stats != nullptr) {
            if (half == 0) { sT[0][threadIdx.x][c] = v; sT[1][threadIdx.x][c] = v * y[c]; }
            else { sT[0][threadIdx.x][c] += v; sT[1][threadIdx.x][c] += v * y[c]; }
          }
        }
      }
    } else if (stats != nullptr) {
#pragma unroll
      for (int c = 0; c < CO; ++c) {
        const float v0 = pvalid ? a0[c] : 0.f, v1 = v1ok ? a1[c] : 0.f;
        sT[0][threadIdx.x][c] = v0 + v1;
        sT[1][threadIdx.x][c] = fmaf(v0, v0, v1 * v1);
      }
    }
    if (stats != nullptr) {
      __syncthreads();
      const int col = threadIdx.x & 31, r0 = (threadIdx.x >> 5) * 32;
      float sa0 = 0.f, sa1 = 0.f, sb0 = 0.f, sb1 = 0.f;
#pragma unroll
      for (int rr = 0; rr < 32; rr += 2) {
        sa0 += sT[0][r0 + rr][col]; sa1 += sT[0][r0 + rr + 1][col];
        sb0 += sT[1][r0 + rr][col]; sb1 += sT[1][r0 + rr + 1][col];
      }
      s_run += sa0 + sa1;
      q_run += sb0 + sb1;
      __syncthreads();
    }
    if (!pvalid) continue;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if (half && !v1ok) break;
      const float* a = half ? a1 : a0;
      if (out_bf16) {
        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + (pix + half) * CO);
#pragma unroll
        for (int c = 0; c < CO; c += 8)
          o[c / 8] = make_uint4(pack2(a[c], a[c + 1]), pack2(a[c + 2], a[c + 3]), pack2(a[c + 4], a[c + 5]), pack2(a[c + 6], a[c + 7]));
      } else {
        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (pix + half) * CO);
#pragma unroll
        for (int c = 0; c < CO; c += 4) o[c / 4] = make_float4(a[c], a[c + 1], a[c + 2], a[c + 3]);
      }
    }
  }
  if (stats == nullptr) return;
  sRed[0][threadIdx.x >> 5][threadIdx.x & 31] = s_run;
  sRed[1][threadIdx.x >> 5][threadIdx.x & 31] = q_run;
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
template <int K, bool SRC_BF16, bool PRE>
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
          float x = v[ci];
          if (PRE) {  // raw input: BatchNorm-apply (+ ReLU) on the fly; a materialised activation is used as is
            x = fmaf(x, sScale[ci], sShift[ci]);
            if (pre_relu) x = fmaxf(x, 0.f);
          }
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
        if (out != nullptr) {  // statistics-only callers (CLEAR-MIM inner forwards) pass no destination
          float* o = out + ((n * Cout + co) * Ho + oh) * (long long)Wo + ow;
          if (two && ((reinterpret_cast<uintptr_t>(o) & 7) == 0)) {
            *reinterpret_cast<float2*>(o) = make_float2(y0, y1);
          } else {
            o[0] = y0;
            if (two) o[1] = y1;
          }
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

// Same layer, materialised bf16 channels-last input (the training path).  CTA = one image x kRT input rows (+ one halo row on
// either side), staged through shared memory with coalesced 16-byte loads: the previous form (every thread loading its own
// 64-byte pixels from global memory, lanes 128 bytes apart) spent its time in L1 tag look-ups (32 lines per load instruction)
// and ran at 32 us whatever its instruction count.  One thread = TWO adjacent output-pixel pairs (ow = 4jj .. 4jj+3) of one
// output row: the weight vectors of a (kh, ci) are loaded once and feed both pairs (3 LDS.128 per 24 FMAs), the FMAs are
// packed fp32x2 (co padded to 4 = two FFMA2 per tap), the input pixels of a kernel row stay packed as bf16 in registers.
// Shared-memory pixels have an 80-byte pitch and their four 16-byte channel chunks are XOR-swizzled with bit 3 of the pixel
// column, so the lanes of a quarter warp (pixel columns 2 apart) hit eight different 16-byte bank groups.
constexpr int kPixPitch = 80;
template <int K, int kRT>
__global__ void __launch_bounds__(kNT) convt_last_x2_kernel(const __nv_bfloat16* __restrict__ src, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ out, double* stats,
                                                            int B, int Cout, int Hi, int Wi, int Ho, int Wo) {
  constexpr int CI = 32;
  extern __shared__ __align__(16) unsigned char sIn[];   // [(kRT + 2) rows][Wi pixels][80 bytes]
  __shared__ __align__(16) float sW[K * K * CI * 4];     // [kh*K+kw][ci][co padded to 4]
  __shared__ float sRed[2][kNT / 32][4];
  for (int i = threadIdx.x; i < K * K * CI * 4; i += kNT) {
    const int co = i & 3, ci = (i >> 2) % CI, t = i / (4 * CI);
    sW[i] = co < Cout ? w[(ci * Cout + co) * K * K + t] : 0.f;  // reference layout w[ci][co][kh][kw]
  }
  float bs[4];
#pragma unroll
  for (int co = 0; co < 4; ++co) bs[co] = (bias != nullptr && co < Cout) ? __ldg(bias + co) : 0.f;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  const int W2 = (Wo + 1) / 2, W4 = (W2 + 1) / 2;
  const int tiles_per_img = (Hi + kRT - 1) / kRT;
  const long long ntiles = (long long)B * tiles_per_img;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long n = tile / tiles_per_img;
    const int r0 = (int)(tile % tiles_per_img) * kRT;
    const int row_lo = r0 - 1 < 0 ? 0 : r0 - 1;                    // first staged input row
    const int row_hi = r0 + kRT + 1 > Hi ? Hi : r0 + kRT + 1;      // one past the last staged row
    __syncthreads();                                               // previous tile's readers are done (also orders the sW fill)
    {
      const int nchunk = (row_hi - row_lo) * Wi * (CI / 8);
      const uint4* g4 = reinterpret_cast<const uint4*>(src + ((n * Hi + row_lo) * Wi) * (long long)CI);
      for (int c = threadIdx.x; c < nchunk; c += kNT) {
        const int pix = c >> 2, c8 = c & 3, col = pix % Wi;
        *reinterpret_cast<uint4*>(sIn + (size_t)pix * kPixPitch + ((c8 ^ ((col >> 3) & 1)) << 4)) = __ldg(g4 + c);
      }
    }
    __syncthreads();
    // input column j0 + ii (ii = -1 .. 2) feeds pair pp (0 / 1) with d = ii - pp: ow = 2j through kw0 = 1 - 2d, ow = 2j + 1 through kw1 = 2 - 2d
    // all warps sweep the even output rows (one kernel row each), then the odd ones (two): no warp waits for another's parity
    for (int it = threadIdx.x; it < 2 * ((kRT * W4 + kNT - 1) / kNT) * kNT; it += kNT) {
      const int per = ((kRT * W4 + kNT - 1) / kNT) * kNT;
      const int a = it >= per ? 1 : 0;
      const int li = it - a * per;
      if (li >= kRT * W4) continue;
      const int rl = li / W4, jj = li - rl * W4;
      const int oh = 2 * (r0 + rl) + a;
      if (r0 + rl >= Hi || oh >= Ho) continue;
      const int j0 = 2 * jj;
      float2 acc[2][2][2];   // [pair][left / right pixel of the pair][co 01 / co 23]
#pragma unroll
      for (int i = 0; i < 8; ++i) (&acc[0][0][0])[i] = make_float2(0.f, 0.f);
#pragma unroll
      for (int kh = 0; kh < K; ++kh) {
        const int th = oh + 1 - kh;  // = 2 * ih
        if (th < 0 || (th & 1) || (th >> 1) >= Hi) continue;   // warp-uniform (same row parity) except at the image border
        const int ih = th >> 1;
        const unsigned char* rowp = sIn + (size_t)(ih - row_lo) * Wi * kPixPitch;
        uint4 v[4][CI / 8];
#pragma unroll
        for (int ii = -1; ii <= 2; ++ii) {
          bool used = false;
#pragma unroll
          for (int pp = 0; pp < 2; ++pp) {
            const int d = ii - pp;
            if (d >= -1 && d <= 1 && ((1 - 2 * d >= 0 && 1 - 2 * d < K) || (2 - 2 * d >= 0 && 2 - 2 * d < K))) used = true;
          }
          if (!used) continue;                                    // compile-time
          const int iw = j0 + ii;
          const bool ok = iw >= 0 && iw < Wi;
          const unsigned char* pp8 = rowp + (ok ? iw : 0) * kPixPitch;
          const int sw = (iw >> 3) & 1;
#pragma unroll
          for (int c8 = 0; c8 < CI / 8; ++c8)
            v[ii + 1][c8] = ok ? *reinterpret_cast<const uint4*>(pp8 + ((c8 ^ sw) << 4)) : make_uint4(0u, 0u, 0u, 0u);
        }
        const float4* wrow = reinterpret_cast<const float4*>(sW + kh * K * CI * 4);
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) {
          float4 wv[K];
#pragma unroll
          for (int kw = 0; kw < K; ++kw) wv[kw] = wrow[kw * CI + ci];
#pragma unroll
          for (int ii = -1; ii <= 2; ++ii) {
            float x = 0.f;
            bool have = false;
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
              const int d = ii - pp;
              if (d < -1 || d > 1) continue;
              const int kw0 = 1 - 2 * d, kw1 = 2 - 2 * d;
              const bool use0 = kw0 >= 0 && kw0 < K, use1 = kw1 >= 0 && kw1 < K;
              if (!use0 && !use1) continue;
              if (!have) {
                const uint4 u4 = v[ii + 1][ci / 8];
                const uint32_t word = ((ci % 8) / 2 == 0) ? u4.x : ((ci % 8) / 2 == 1) ? u4.y : ((ci % 8) / 2 == 2) ? u4.z : u4.w;
                x = __uint_as_float((ci & 1) ? (word & 0xffff0000u) : (word << 16));
                have = true;
              }
              const float2 xx = make_float2(x, x);
              if (use0) {
                acc[pp][0][0] = __ffma2_rn(xx, make_float2(wv[kw0].x, wv[kw0].y), acc[pp][0][0]);
                acc[pp][0][1] = __ffma2_rn(xx, make_float2(wv[kw0].z, wv[kw0].w), acc[pp][0][1]);
              }
              if (use1) {
                acc[pp][1][0] = __ffma2_rn(xx, make_float2(wv[kw1].x, wv[kw1].y), acc[pp][1][0]);
                acc[pp][1][1] = __ffma2_rn(xx, make_float2(wv[kw1].z, wv[kw1].w), acc[pp][1][1]);
              }
            }
          }
        }
      }
      const int ow0 = 4 * jj;
#pragma unroll
      for (int co = 0; co < 4; ++co) {
        if (co < Cout) {
          float y[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 a2 = acc[k >> 1][k & 1][co >> 1];
            y[k] = ((co & 1) ? a2.y : a2.x) + bs[co];
          }
          if (out != nullptr) {  // statistics-only callers (CLEAR-MIM inner forwards) pass no destination
            float* o = out + ((n * Cout + co) * Ho + oh) * (long long)Wo + ow0;
            if (ow0 + 3 < Wo && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
              *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (ow0 + k < Wo) o[k] = y[k];
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (ow0 + k < Wo) { s[co] += y[k]; q[co] = fmaf(y[k], y[k], q[co]); }
        }
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

// ---------------------------------------------------------------------------------------------------------
// Weight gradient of the same two boundary layers (Conv2d(C<=4 -> 32) and ConvTranspose2d(32 -> C<=4), stride 2, pad 1):
// both are  dW[f][c][kh][kw] = sum_{n,h,w} feat[n,h,w,f] * img[n,c,2h-1+kh,2w-1+kw]  with a 32-channel channels-last bf16
// "feature" side (dy of the first conv / the materialised input activation of the last conv-transpose) and a
// <=4-channel NCHW "image" side (x / dy of the last layer) — vae.py:16,43,114,153 through autograd.
// Register-blocked outer product on CUDA cores: lane = (feature group, pixel slot); each lane keeps C*K*K*FPT
// accumulators, one image value feeds FPT FMAs, image loads are 8/16-lane broadcasts served by L1.  As a 128-row
// tensor-core GEMM this shape (M = C*K*K = 27..48) ran the scalar gather path and took ~250 us per layer.
template <int K, int C, int FPT, bool IMG_BF16>
__global__ void __launch_bounds__(kNT, (C * K * K * FPT <= 112 ? 3 : 2)) boundary_wgrad_kernel(const __nv_bfloat16* __restrict__ feat, const void* __restrict__ img,
                                                             float* __restrict__ dw, int B, int Hf, int Wf, int Hi, int Wi) {
  constexpr int FG = 32 / FPT;   // lanes per pixel slot
  constexpr int PS = 32 / FG;    // pixel slots per warp
  constexpr int KK = K * K;
  __shared__ float sAcc[C * KK * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fg = lane % FG, pg = lane / FG;
  float acc[C * KK][FPT];
#pragma unroll
  for (int i = 0; i < C * KK; ++i)
#pragma unroll
    for (int j = 0; j < FPT; ++j) acc[i][j] = 0.f;
  const int npix = B * Hf * Wf;                      // host guarantees < 2^31
  // each warp walks a contiguous range of work items (PS pixels each): the pixel coordinates advance incrementally (no
  // integer division in the loop) and consecutive items re-use the image rows they share through L1
  const int nitems = (npix + PS - 1) / PS;
  const int nwarps = gridDim.x * (kNT / 32), gw = blockIdx.x * (kNT / 32) + warp;
  const int per = (nitems + nwarps - 1) / nwarps;
  const int q_begin = min(gw * per, nitems), q_end = min(q_begin + per, nitems);
  int pix = q_begin * PS + pg;
  int w = pix % Wf, h = (pix / Wf) % Hf, n = pix / (Wf * Hf);
  const int HiWi = Hi * Wi;
  for (int q = q_begin; q < q_end; ++q) {
    const bool pv = pix < npix;
    float f[FPT];
    {
      const __nv_bfloat16* fp = feat + (long long)(pv ? pix : 0) * 32 + fg * FPT;
      if (FPT == 4) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(fp));
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        f[0] = a.x; f[1] = a.y; f[FPT - 2] = b.x; f[FPT - 1] = b.y;
      } else {
        const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(fp));
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
        f[0] = a.x; f[1] = a.y;
      }
#pragma unroll
      for (int j = 0; j < FPT; ++j) f[j] = pv ? f[j] : 0.f;
    }
    const int h0 = 2 * h - 1, w0 = 2 * w - 1;
    bool rok[K], cok[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { rok[k] = pv && h0 + k >= 0 && h0 + k < Hi; cok[k] = w0 + k >= 0 && w0 + k < Wi; }
    // one row pointer per (channel, kh); the kw offsets are immediates
    const long long ibase = ((long long)n * C * Hi + h0) * Wi + w0;
    float xv[C * KK];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int kh = 0; kh < K; ++kh) {
        const long long roff = ibase + c * HiWi + kh * Wi;
        if (IMG_BF16) {
          const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(img) + roff;
#pragma unroll
          for (int kw = 0; kw < K; ++kw) xv[c * KK + kh * K + kw] = (rok[kh] && cok[kw]) ? __bfloat162float(__ldg(rp + kw)) : 0.f;
        } else {
          const float* rp = reinterpret_cast<const float*>(img) + roff;
#pragma unroll
          for (int kw = 0; kw < K; ++kw) xv[c * KK + kh * K + kw] = (rok[kh] && cok[kw]) ? __ldg(rp + kw) : 0.f;
        }
      }
#pragma unroll
    for (int i = 0; i < C * KK; ++i)
#pragma unroll
      for (int j = 0; j < FPT; ++j) acc[i][j] = fmaf(xv[i], f[j], acc[i][j]);
    // next item of this lane: PS pixels further along the row-major pixel order (PS <= Wf)
    pix += PS;
    w += PS;
    if (w >= Wf) {
      w -= Wf;
      if (++h == Hf) { h = 0; ++n; }
    }
  }
  // pixel slots -> one value per (tap, feature) per warp; warps -> one per CTA (shared memory); CTAs -> fp32 atomics
#pragma unroll
  for (int i = 0; i < C * KK; ++i)
#pragma unroll
    for (int j = 0; j < FPT; ++j) {
      float v = acc[i][j];
#pragma unroll
      for (int o = FG; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[i][j] = v;
    }
  for (int wv = 0; wv < kNT / 32; ++wv) {
    if (warp == wv && pg == 0) {
#pragma unroll
      for (int i = 0; i < C * KK; ++i)
#pragma unroll
        for (int j = 0; j < FPT; ++j) {
          float* d = sAcc + i * 32 + fg * FPT + j;
          *d = wv == 0 ? acc[i][j] : *d + acc[i][j];
        }
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < C * KK * 32; e += kNT) {
    const int i = e >> 5, fch = e & 31;            // i = c * KK + tap
    atomicAdd(dw + (long long)fch * (C * KK) + i, sAcc[e]);   // reference layout [f][c][kh][kw]
  }
}

template <int K, int C, int FPT>
int launch_bwg(bool img_bf16, const __nv_bfloat16* feat, const void* img, float* dw, int B, int Hf, int Wf, int Hi, int Wi,
               cudaStream_t st) {
  const long long items = ((long long)B * Hf * Wf + FPT - 1) / FPT;   // warp iterations (FPT pixel slots per warp)
  const long long g = (items + kNT / 32 - 1) / (kNT / 32);
  const long long cap = 148 * (C * K * K * FPT <= 112 ? 3 : 2);        // CTAs per SM allowed by the accumulator count
  const int grid = (int)(g < 1 ? 1 : g > cap ? cap : g);
  if (img_bf16) boundary_wgrad_kernel<K, C, FPT, true><<<grid, kNT, 0, st>>>(feat, img, dw, B, Hf, Wf, Hi, Wi);
  else boundary_wgrad_kernel<K, C, FPT, false><<<grid, kNT, 0, st>>>(feat, img, dw, B, Hf, Wf, Hi, Wi);
  CV_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Decoder fc layer Linear(2D -> 2048) (vae.py:33,137): K = 2D <= 64 is far too shallow for a tensor-core tile
// (one k-block), so it runs on CUDA cores in fp32: thread = output column (its K weights in registers), a CTA
// streams 32 rows of z through shared memory (broadcast reads), stores are 512-byte coalesced rows, and the
// BatchNorm1d batch moments of the column are thread-local sums (two fp64 atomics per thread at the end).
template <int MAXK>
__global__ void __launch_bounds__(128) fc_fwd_kernel(const float* __restrict__ z, const float* __restrict__ w,
                                                     const float* __restrict__ bias, float* __restrict__ out, double* stats,
                                                     int B, int K, int N) {
  constexpr int R = 32;
  __shared__ __align__(16) float sZ[R * MAXK];
  const int n = blockIdx.x * 128 + threadIdx.x;
  const int row0 = blockIdx.y * R;
  const int nrows = min(R, B - row0);
  for (int i = threadIdx.x; i < nrows * K; i += 128) sZ[(i / K) * MAXK + (i % K)] = __ldg(z + (long long)row0 * K + i);
  float wr[MAXK];
#pragma unroll
  for (int k = 0; k < MAXK; ++k) wr[k] = (n < N && k < K) ? __ldg(w + (long long)n * K + k) : 0.f;
  const float bv = (n < N && bias) ? __ldg(bias + n) : 0.f;
  __syncthreads();
  float s = 0.f, q = 0.f;
  if (n < N) {
    for (int r = 0; r < nrows; ++r) {
      const float4* zr = reinterpret_cast<const float4*>(sZ + r * MAXK);
      float a0 = bv, a1 = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < MAXK / 4; ++k4) {
        if (k4 * 4 < K) {
          const float4 zz = zr[k4];
          a0 = fmaf(zz.x, wr[4 * k4], a0); a1 = fmaf(zz.y, wr[4 * k4 + 1], a1);
          a0 = fmaf(zz.z, wr[4 * k4 + 2], a0); a1 = fmaf(zz.w, wr[4 * k4 + 3], a1);
        }
      }
      const float a = a0 + a1;
      out[(long long)(row0 + r) * N + n] = a;
      s += a;
      q = fmaf(a, a, q);
    }
    if (stats != nullptr) {
      atomicAdd(stats + n, (double)s);
      atomicAdd(stats + N + n, (double)q);
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
  if (!g || !src || !src->ptr || !dst || !weight || batch <= 0) return CLEARVAE_EINVAL;
  if (!dst->ptr && !(g->transposed && stats)) return CLEARVAE_EINVAL;  // no destination = statistics only (last conv-transpose)
  if (!clearvae_conv_direct_supported(g, src, dst)) return CLEARVAE_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (!g->transposed) {
    if (pre_scale || pre_relu) return CLEARVAE_EUNSUPPORTED;
    const int Ho = (g->Hin + 2 - g->k) / 2 + 1;
    const long long npix = batch * Ho * Ho;
    static const bool v1 = getenv("CLEARVAE_CONV_FIRST_V1") != nullptr;
    const long long npair = batch * Ho * ((Ho + 1) / 2);
    if (v1 && g->k == 3)
      conv_first_kernel<3, false, false><<<grid_for(npix), kNT, 0, st>>>(src->ptr, weight, bias, dst->ptr, dst->dtype == CLEARVAE_BF16,
                                                                        stats, (int)batch, g->Cin, g->Hin, g->Win, Ho, Ho, nullptr, 0, nullptr, nullptr);
    else if (v1)
      conv_first_kernel<4, false, false><<<grid_for(npix), kNT, 0, st>>>(src->ptr, weight, bias, dst->ptr, dst->dtype == CLEARVAE_BF16,
                                                                        stats, (int)batch, g->Cin, g->Hin, g->Win, Ho, Ho, nullptr, 0, nullptr, nullptr);
    else if (g->k == 3)
      conv_first2_kernel<3, false, false><<<grid_for(npair), kNT, 0, st>>>(src->ptr, weight, bias, dst->ptr, dst->dtype == CLEARVAE_BF16,
                                                                         stats, (int)batch, g->Cin, g->Hin, g->Win, Ho, Ho, nullptr, 0, nullptr, nullptr);
    else
      conv_first2_kernel<4, false, false><<<grid_for(npair), kNT, 0, st>>>(src->ptr, weight, bias, dst->ptr, dst->dtype == CLEARVAE_BF16,
                                                                         stats, (int)batch, g->Cin, g->Hin, g->Win, Ho, Ho, nullptr, 0, nullptr, nullptr);
  } else {
    const int Ho = (g->Hin - 1) * 2 - 2 + g->k + g->out_pad;
    const long long npix = batch * 2 * ((Ho + 1) / 2) * ((Ho + 1) / 2);  // one thread per output-pixel pair
    const bool bf = src->dtype == CLEARVAE_BF16;
    const bool pre = pre_scale != nullptr || pre_relu;
#define CV_LAUNCH_T(KK, BF, PR)                                                                                               \
  convt_last_kernel<KK, BF, PR><<<grid_for(npix), kNT, 0, st>>>(src->ptr, pre_scale, pre_shift, pre_relu, weight, bias,       \
                                                                (float*)dst->ptr, stats, (int)batch, g->Cout, g->Hin, g->Win, Ho, Ho)
#define CV_LAUNCH_P(KK, BF) do { if (pre) CV_LAUNCH_T(KK, BF, true); else CV_LAUNCH_T(KK, BF, false); } while (0)
    // rows per CTA: the whole image when it is small (all warps sweep one parity at a time), 8 input rows otherwise
    const int rt = g->Hin <= 16 ? 16 : 8;
    const size_t stage_bytes = (size_t)(rt + 2) * g->Win * kPixPitch;
    if (bf && !pre && stage_bytes <= 64 * 1024 && getenv("CLEARVAE_CONVT_LAST_V1") == nullptr) {
      long long nt = batch * ((g->Hin + rt - 1) / rt);              // persistent past 8 CTAs per SM
      if (nt > 148 * 8) nt = 148 * 8;
      const __nv_bfloat16* sp = reinterpret_cast<const __nv_bfloat16*>(src->ptr);
#define CV_LAUNCH_X2(KK, RT)                                                                                                  \
  do {                                                                                                                        \
    static bool attr_done = false;                                                                                            \
    if (!attr_done) {                                                                                                         \
      cudaFuncSetAttribute(convt_last_x2_kernel<KK, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);             \
      attr_done = true;                                                                                                       \
    }                                                                                                                         \
    convt_last_x2_kernel<KK, RT><<<(unsigned)nt, kNT, stage_bytes, st>>>(sp, weight, bias, (float*)dst->ptr, stats, (int)batch, \
                                                                         g->Cout, g->Hin, g->Win, Ho, Ho);                    \
  } while (0)
      if (g->k == 3) { if (rt == 16) CV_LAUNCH_X2(3, 16); else CV_LAUNCH_X2(3, 8); }
      else { if (rt == 16) CV_LAUNCH_X2(4, 16); else CV_LAUNCH_X2(4, 8); }
#undef CV_LAUNCH_X2
    } else if (g->k == 3) { if (bf) CV_LAUNCH_P(3, true); else CV_LAUNCH_P(3, false); }
    else { if (bf) CV_LAUNCH_P(4, true); else CV_LAUNCH_P(4, false); }
#undef CV_LAUNCH_P
#undef CV_LAUNCH_T
  }
  CV_LAUNCH_CHECK();
  return 0;
}

int clearvae_conv_direct_wgrad(const clearvae_conv_geom* g, int64_t batch, const clearvae_tensor4* src, const clearvae_tensor4* dy,
                               float* dweight, void* stream) {
  if (!g || !src || !src->ptr || !dy || !dy->ptr || !dweight || batch <= 0) return CLEARVAE_EINVAL;
  if (g->stride != 2 || g->pad != 1 || (g->k != 3 && g->k != 4) || g->Hin != g->Win || batch * (int64_t)g->Hin * g->Win * 4 >= (1LL << 31))
    return CLEARVAE_EUNSUPPORTED;
  // feature side: 32 channels, channels-last bf16; image side: C <= 4 channels, NCHW, fp32 or bf16
  const clearvae_tensor4 *feat, *img;
  int C, Hf, Hi;
  if (!g->transposed) {
    if (g->Cout != 32 || g->Cin < 1 || g->Cin > 4) return CLEARVAE_EUNSUPPORTED;
    feat = dy; img = src; C = g->Cin; Hi = g->Hin; Hf = (g->Hin + 2 - g->k) / 2 + 1;
  } else {
    if (g->Cin != 32 || g->Cout < 1 || g->Cout > 4) return CLEARVAE_EUNSUPPORTED;
    feat = src; img = dy; C = g->Cout; Hf = g->Hin; Hi = (g->Hin - 1) * 2 - 2 + g->k + g->out_pad;
  }
  if (Hf < 4) return CLEARVAE_EUNSUPPORTED;   // the kernel advances up to 4 pixels along a row per step
  if (feat->dtype != CLEARVAE_BF16 || feat->sc != 1 || feat->sw != 32 || feat->sh != (int64_t)Hf * 32 ||
      feat->sn != (int64_t)Hf * Hf * 32 || ((uintptr_t)feat->ptr & 7))
    return CLEARVAE_EUNSUPPORTED;
  if (img->sw != 1 || img->sh != Hi || img->sc != (int64_t)Hi * Hi || img->sn != (int64_t)C * Hi * Hi) return CLEARVAE_EUNSUPPORTED;
  const bool ibf = img->dtype == CLEARVAE_BF16;
  const __nv_bfloat16* fp = reinterpret_cast<const __nv_bfloat16*>(feat->ptr);
  cudaStream_t st = (cudaStream_t)stream;
  const int B = (int)batch;
#define CV_BWG(KK, CC, FF) return launch_bwg<KK, CC, FF>(ibf, fp, img->ptr, dweight, B, Hf, Hf, Hi, Hi, st)
  if (g->k == 3) {
    switch (C) { case 1: CV_BWG(3, 1, 4); case 2: CV_BWG(3, 2, 4); case 3: CV_BWG(3, 3, 4); default: CV_BWG(3, 4, 2); }
  } else {
    switch (C) { case 1: CV_BWG(4, 1, 4); case 2: CV_BWG(4, 2, 2); case 3: CV_BWG(4, 3, 2); default: CV_BWG(4, 4, 2); }
  }
#undef CV_BWG
}

int clearvae_conv_direct_dgrad(const clearvae_conv_geom* g, int64_t batch, const clearvae_tensor4* dy, const float* weight,
                               const clearvae_tensor4* dst, const clearvae_tensor4* mask_src, const float* mask_scale,
                               const float* mask_shift, double* stats, void* stream) {
  if (!g || !dy || !dy->ptr || !weight || !dst || !dst->ptr || !mask_src || !mask_src->ptr || batch <= 0) return CLEARVAE_EINVAL;
  if (!g->transposed || g->stride != 2 || g->pad != 1 || (g->k != 3 && g->k != 4) || g->Hin != g->Win || g->Cin != 32 || g->Cout < 1 ||
      g->Cout > 4)
    return CLEARVAE_EUNSUPPORTED;
  const int Ho = (g->Hin - 1) * 2 - 2 + g->k + g->out_pad, Hi = g->Hin;
  // dy: NCHW [B, Cout, Ho, Ho]; dst / mask: channels-last [B, Hin, Hin, 32]
  if (dy->sw != 1 || dy->sh != Ho || dy->sc != (int64_t)Ho * Ho || dy->sn != (int64_t)g->Cout * Ho * Ho) return CLEARVAE_EUNSUPPORTED;
  auto cl32 = [&](const clearvae_tensor4* t) {
    return t->sc == 1 && t->sw == 32 && t->sh == (int64_t)Hi * 32 && t->sn == (int64_t)Hi * Hi * 32 && !((uintptr_t)t->ptr & 15);
  };
  if (!cl32(dst) || !cl32(mask_src)) return CLEARVAE_EUNSUPPORTED;

  const bool xb = dy->dtype == CLEARVAE_BF16;
  cudaStream_t st = (cudaStream_t)stream;
#define CV_DG(KK, XB)                                                                                                          \
  conv_first2_kernel<KK, XB, true><<<grid_for(batch * Hi * ((Hi + 1) / 2)), kNT, 0, st>>>(                                      \
      dy->ptr, weight, nullptr, dst->ptr, dst->dtype == CLEARVAE_BF16, stats, (int)batch, g->Cout, Ho, Ho, Hi, Hi, mask_src->ptr, \
      mask_src->dtype == CLEARVAE_BF16, mask_scale, mask_shift)
  if (g->k == 3) { if (xb) CV_DG(3, true); else CV_DG(3, false); }
  else { if (xb) CV_DG(4, true); else CV_DG(4, false); }
#undef CV_DG
  CV_LAUNCH_CHECK();
  return 0;
}

int clearvae_fc_fwd(const float* z, const float* weight, const float* bias, float* out, double* stats, int64_t B, int32_t K,
                    int32_t N, void* stream) {
  if (!z || !weight || !out || B <= 0 || K <= 0 || N <= 0) return CLEARVAE_EINVAL;
  if (K > 64 || K % 4 != 0 || B > (1 << 24) || ((uintptr_t)z & 15)) return CLEARVAE_EUNSUPPORTED;
  dim3 grid((unsigned)((N + 127) / 128), (unsigned)((B + 31) / 32));
  cudaStream_t st = (cudaStream_t)stream;
  if (K <= 16) fc_fwd_kernel<16><<<grid, 128, 0, st>>>(z, weight, bias, out, stats, (int)B, K, N);
  else fc_fwd_kernel<64><<<grid, 128, 0, st>>>(z, weight, bias, out, stats, (int)B, K, N);
  CV_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
