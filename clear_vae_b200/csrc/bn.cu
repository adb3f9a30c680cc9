// Train-mode BatchNorm pieces around the conv GEMMs (nn.BatchNorm1d / nn.BatchNorm2d as
// stacked at vae.py:17-44,115-154; semantics in SURVEY.md §8a': biased batch variance for
// normalisation, unbiased for the running estimate, momentum 0.1, eps 1e-5).
//
// The batch statistics themselves are accumulated by the GEMM epilogues (double
// atomics into a [2][C] buffer); the kernels here turn them into per-channel
// scale/shift (forward) or into the affine coefficients of the BatchNorm backward,
//     dy = a_c * g + b_c * y + c_c,
// and stream the few elementwise passes that cannot ride on a GEMM (fc block, final
// BN + sigmoid + reconstruction error).  All elementwise kernels are HBM-bound:
// 16-byte accesses, grid = multiple of the SM count.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {
constexpr int kNT = 256;

__device__ __forceinline__ float ldf(const void* p, long long i, int bf) {
  return bf ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void stf(void* p, long long i, int bf, float v) {
  if (bf) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16(v);
  else reinterpret_cast<float*>(p)[i] = v;
}
inline int grid_for(long long n, int per_thread = 4) {
  long long g = (n / per_thread + kNT - 1) / kNT;
  const long long cap = 148 * 8;
  return (int)(g < 1 ? 1 : g > cap ? cap : g);
}

// ---------------------------------------------------------------------------
// stats[0..Cs) = sum, stats[Cs..2Cs) = sum of squares, Cs = C * group entries;
// channel c owns entries [c*group, (c+1)*group).  Writes scale/shift (optionally
// expanded `expand` times), saves mean / invstd, updates the running estimates and
// clears the accumulator for the next step.
// ---------------------------------------------------------------------------
__global__ void bn_finalize_kernel(double* stats, int C, int group, double count, const float* gamma, const float* beta,
                                   float* running_mean, float* running_var, float momentum, float eps, float* scale,
                                   float* shift, int expand, float* save_mean, float* save_invstd, int repeat) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int Cs = C * group;
  double s = 0.0, q = 0.0;
  for (int i = 0; i < group; ++i) {
    s += stats[c * group + i];
    q += stats[Cs + c * group + i];
    stats[c * group + i] = 0.0;
    stats[Cs + c * group + i] = 0.0;
  }
  const double mean = s / count;
  double var = q / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float sc = g * invstd, sh = b - (float)mean * g * invstd;
  for (int i = 0; i < expand; ++i) { scale[c * expand + i] = sc; shift[c * expand + i] = sh; }
  save_mean[c] = (float)mean;
  save_invstd[c] = invstd;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    float rm = running_mean[c], rv = running_var[c];
    for (int i = 0; i < repeat; ++i) {  // `repeat` identical forwards (CLEAR-MIM inner loop) = repeated momentum updates
      rm = (1.f - momentum) * rm + momentum * (float)mean;
      rv = (1.f - momentum) * rv + momentum * (float)unbiased;
    }
    running_mean[c] = rm;
    running_var[c] = rv;
  }
}

// ---------------------------------------------------------------------------
// generic two-moment reduction over a flat tensor, channel(idx) = (idx / inner) % C.
//   mode 0: (sum y, sum y^2)
//   mode 1: g' = g * [act > 0] (act optional), (sum g', sum g'*y)
// ---------------------------------------------------------------------------
// position owned by this thread inside one sample's [C * inner] period; `blocked` mapping gives every
// thread of a CTA the same channel (one atomic pair per CTA instead of one per thread)
__device__ __forceinline__ long long reduce_pos(int C, long long inner, int blocked, int* chan) {
  if (blocked) {
    const int chunks = (int)((inner + kNT - 1) / kNT);
    const int c = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
    const long long i = (long long)chunk * kNT + threadIdx.x;
    *chan = c;
    return i < inner ? c * inner + i : -1;
  }
  const long long pos = (long long)blockIdx.x * kNT + threadIdx.x;
  *chan = (int)(pos / inner);
  return pos < (long long)C * inner ? pos : -1;
}
__device__ __forceinline__ void reduce_commit(double d0, double d1, int c, int C, int blocked, double* stats) {
  if (blocked) {
    __shared__ double sh[2][kNT / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { d0 += __shfl_xor_sync(0xffffffffu, d0, o); d1 += __shfl_xor_sync(0xffffffffu, d1, o); }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = d0; sh[1][threadIdx.x >> 5] = d1; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int i = 0; i < kNT / 32; ++i) { a += sh[0][i]; b += sh[1][i]; }
      atomicAdd(stats + c, a);
      atomicAdd(stats + C + c, b);
    }
  } else if (c < C) {
    atomicAdd(stats + c, d0);
    atomicAdd(stats + C + c, d1);
  }
}

__global__ void __launch_bounds__(kNT) bn_reduce_kernel(const void* y, int y_bf, const void* g, int g_bf, const void* act,
                                                        int act_bf, const float* __restrict__ mscale,
                                                        const float* __restrict__ mshift, long long total, int C,
                                                        long long inner, int mode, int blocked, double* stats) {
  const long long period = (long long)C * inner;
  const long long nper = total / period;  // samples
  int c;
  const long long pos = reduce_pos(C, inner, blocked, &c);
  double d0 = 0.0, d1 = 0.0;
  if (pos >= 0) {
    float s0 = 0.f, s1 = 0.f;
    int cnt = 0;
    for (long long smp = blockIdx.y; smp < nper; smp += gridDim.y) {
      const long long i = smp * period + pos;
      const float yv = ldf(y, i, y_bf);
      if (mode == 0) {
        s0 += yv; s1 = fmaf(yv, yv, s1);
      } else {
        float gv = ldf(g, i, g_bf);
        if (act != nullptr && !(ldf(act, i, act_bf) > 0.f)) gv = 0.f;
        if (mscale != nullptr && !(fmaf(yv, __ldg(mscale + c), __ldg(mshift + c)) > 0.f)) gv = 0.f;
        s0 += gv; s1 = fmaf(gv, yv, s1);
      }
      if (++cnt == 64) { d0 += s0; d1 += s1; s0 = s1 = 0.f; cnt = 0; }
    }
    d0 += s0; d1 += s1;
  }
  reduce_commit(d0, d1, c, C, blocked, stats);
}

// ---------------------------------------------------------------------------
// out = act(raw * scale[ch] + shift[ch]); act: 0 none, 1 relu, 2 sigmoid.
// optional fused reconstruction error vs `target` (sum of squares, * invB) like losses.py:45-47.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kNT) bn_act_fwd_kernel(const void* raw, int raw_bf, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, long long total, int C,
                                                         long long inner, int act, int tC, int tHW, void* out, int out_bf,
                                                         const float* __restrict__ target, float invB, float* sse_out,
                                                         float* partial, unsigned* ticket) {
  __shared__ float sred[kNT / 32 + 1];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * kNT + threadIdx.x; i < total; i += (long long)gridDim.x * kNT) {
    const int ch = (int)((i / inner) % C);
    float v = fmaf(ldf(raw, i, raw_bf), __ldg(scale + ch), __ldg(shift + ch));
    if (act == 1) v = fmaxf(v, 0.f);
    else if (act == 2) v = 1.f / (1.f + __expf(-v));
    long long o = i;
    if (tC > 0) {  // channel-major [b][c][hw] -> channels-last [b][hw][c]
      const long long per = (long long)tC * tHW, b = i / per, r = i - b * per;
      const int c2 = (int)(r / tHW), hw = (int)(r - (long long)c2 * tHW);
      o = b * per + (long long)hw * tC + c2;
    }
    stf(out, o, out_bf, v);
    if (target != nullptr) { const float d = v - __ldg(target + i); acc = fmaf(d, d, acc); }
  }
  if (target == nullptr) return;
  acc = cv::block_sum<kNT>(acc, sred);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = acc;
    __threadfence();
    sred[kNT / 32] = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1.f : 0.f;
  }
  __syncthreads();
  if (sred[kNT / 32] != 0.f) {
    __threadfence();
    float s = 0.f;
    for (int c = threadIdx.x; c < (int)gridDim.x; c += kNT) s += __ldcg(partial + c);
    s = cv::block_sum<kNT>(s, sred);
    if (threadIdx.x == 0) { *sse_out = s * invB; *ticket = 0u; }
  }
}

// ---------------------------------------------------------------------------
// final layer backward: xhat = sigmoid(bn(raw)); recon = invB * sum (xhat - x)^2
//   g_pre = (grad_xhat_ext + grad_recon * 2 invB (xhat - x)) * xhat (1 - xhat)
// writes g_pre and accumulates (sum g_pre, sum g_pre * raw) per channel.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kNT) sigmoid_mse_bwd_kernel(const float* __restrict__ xhat, const float* __restrict__ x,
                                                              const float* __restrict__ grad_recon,
                                                              const float* __restrict__ grad_ext, const void* raw, int raw_bf,
                                                              long long total, int C, long long inner, float twoInvB,
                                                              int blocked, float* __restrict__ g_pre, double* stats) {
  const float gr = grad_recon ? __ldg(grad_recon) * twoInvB : 0.f;
  const long long period = (long long)C * inner;
  const long long nper = total / period;
  int c;
  const long long pos = reduce_pos(C, inner, blocked, &c);
  double d0 = 0.0, d1 = 0.0;
  if (pos >= 0) {
    float s0 = 0.f, s1 = 0.f;
    int cnt = 0;
    for (long long smp = blockIdx.y; smp < nper; smp += gridDim.y) {
      const long long i = smp * period + pos;
      const float xh = __ldg(xhat + i);
      float g = gr * (xh - __ldg(x + i));
      if (grad_ext) g += __ldg(grad_ext + i);
      g *= xh * (1.f - xh);
      g_pre[i] = g;
      s0 += g; s1 = fmaf(g, ldf(raw, i, raw_bf), s1);
      if (++cnt == 64) { d0 += s0; d1 += s1; s0 = s1 = 0.f; cnt = 0; }
    }
    d0 += s0; d1 += s1;
  }
  reduce_commit(d0, d1, c, C, blocked, stats);
}

// float4 form (fp32 raw, inner % 4 == 0, 16-byte aligned): thread = four consecutive positions of one channel, CTAs along y
// stride over the samples; the scalar kernel ran at ~1.1 TB/s on the 3 x 28 x 28 output (35 us of a 1.5 ms step)
__global__ void __launch_bounds__(kNT) sigmoid_mse_bwd_vec_kernel(const float4* __restrict__ xhat, const float4* __restrict__ x,
                                                                  const float* __restrict__ grad_recon,
                                                                  const float4* __restrict__ grad_ext, const float4* __restrict__ raw,
                                                                  long long nper, int C, long long inner4, float twoInvB,
                                                                  float4* __restrict__ g_pre, double* stats) {
  const float gr = grad_recon ? __ldg(grad_recon) * twoInvB : 0.f;
  const int chunks = (int)((inner4 + kNT - 1) / kNT);
  const int c = blockIdx.x / chunks;
  const long long i4 = (long long)(blockIdx.x % chunks) * kNT + threadIdx.x;
  const long long period4 = (long long)C * inner4;
  double d0 = 0.0, d1 = 0.0;
  if (i4 < inner4) {
    float s0 = 0.f, s1 = 0.f;
    int cnt = 0;
    const long long pos = c * inner4 + i4;
    auto one = [&](const float4 xh, const float4 xv, const float4 rv, const float4 ge, long long i) {
      float4 g;
      g.x = (gr * (xh.x - xv.x) + ge.x) * (xh.x * (1.f - xh.x));
      g.y = (gr * (xh.y - xv.y) + ge.y) * (xh.y * (1.f - xh.y));
      g.z = (gr * (xh.z - xv.z) + ge.z) * (xh.z * (1.f - xh.z));
      g.w = (gr * (xh.w - xv.w) + ge.w) * (xh.w * (1.f - xh.w));
      g_pre[i] = g;
      s0 += (g.x + g.y) + (g.z + g.w);
      s1 = fmaf(g.x, rv.x, fmaf(g.y, rv.y, fmaf(g.z, rv.z, fmaf(g.w, rv.w, s1))));
    };
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    long long smp = blockIdx.y;
    for (; smp + gridDim.y < nper; smp += 2 * (long long)gridDim.y) {   // two samples in flight
      const long long i = smp * period4 + pos, j = i + (long long)gridDim.y * period4;
      const float4 a0 = __ldg(xhat + i), a1 = __ldg(xhat + j), b0 = __ldg(x + i), b1 = __ldg(x + j);
      const float4 r0 = __ldg(raw + i), r1 = __ldg(raw + j);
      const float4 e0 = grad_ext ? __ldg(grad_ext + i) : z4, e1 = grad_ext ? __ldg(grad_ext + j) : z4;
      one(a0, b0, r0, e0, i);
      one(a1, b1, r1, e1, j);
      if ((cnt += 2) >= 16) { d0 += s0; d1 += s1; s0 = s1 = 0.f; cnt = 0; }
    }
    if (smp < nper) {
      const long long i = smp * period4 + pos;
      one(__ldg(xhat + i), __ldg(x + i), __ldg(raw + i), grad_ext ? __ldg(grad_ext + i) : z4, i);
    }
    d0 += s0; d1 += s1;
  }
  reduce_commit(d0, d1, c, C, 1, stats);
}

// ---------------------------------------------------------------------------
// BatchNorm backward coefficients from (S1 = sum g, S2 = sum g*y):
//   sum g*xhat = r (S2 - mu S1);  dgamma = that;  dbeta = S1
//   dy = gamma r g  -  gamma r^2 (sum g xhat)/M * y  +  gamma r (mu r (sum g xhat) - S1)/M
// coef layout [3][C] (a, b, c); clears the accumulator.
// ---------------------------------------------------------------------------
__global__ void bn_bwd_coef_kernel(double* stats, int C, int group, double count, const float* gamma, const float* save_mean,
                                   const float* save_invstd, float* coef, float* dgamma, float* dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int Cs = C * group;
  double s1 = 0.0, s2 = 0.0;
  for (int i = 0; i < group; ++i) {
    s1 += stats[c * group + i];
    s2 += stats[Cs + c * group + i];
    stats[c * group + i] = 0.0;
    stats[Cs + c * group + i] = 0.0;
  }
  const double mu = save_mean[c], r = save_invstd[c], g = gamma ? gamma[c] : 1.0;
  const double sgx = r * (s2 - mu * s1);
  coef[c] = (float)(g * r);
  coef[C + c] = (float)(-g * r * r * sgx / count);
  coef[2 * C + c] = (float)(g * r * (mu * r * sgx - s1) / count);
  if (dgamma) dgamma[c] = (float)sgx;
  if (dbeta) dbeta[c] = (float)s1;
}

// dy = a[ch] * g' + b[ch] * y + c[ch],  g' = g * [act > 0] when act is given
__device__ __forceinline__ uint32_t bf2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// fp32 gradient and fp32 pre-activation, four consecutive elements per thread: either one channel per element run
// (inner % 4 == 0: the NCHW last decoder block) or four consecutive channels (inner == 1: the decoder's fc block with its
// recomputed ReLU mask).  The scalar kernel below paid an integer division and five scalar coefficient loads per element.
template <bool SAME_CH>
__global__ void __launch_bounds__(kNT) bn_bwd_apply_vec4_kernel(const float4* __restrict__ g, const float4* __restrict__ y,
                                                                const float* __restrict__ mscale, const float* __restrict__ mshift,
                                                                const float* __restrict__ coef, long long n4, int C, long long inner4,
                                                                void* dy, int dy_bf) {
  for (long long i = (long long)blockIdx.x * kNT + threadIdx.x; i < n4; i += (long long)gridDim.x * kNT) {
    float4 gv = __ldg(g + i);
    const float4 yv = __ldg(y + i);
    float4 a, b, c;
    if (SAME_CH) {
      const int ch = (int)((i / inner4) % C);
      const float a1 = __ldg(coef + ch), b1 = __ldg(coef + C + ch), c1 = __ldg(coef + 2 * C + ch);
      a = make_float4(a1, a1, a1, a1); b = make_float4(b1, b1, b1, b1); c = make_float4(c1, c1, c1, c1);
      if (mscale != nullptr) {
        const float ms = __ldg(mscale + ch), mh = __ldg(mshift + ch);
        if (!(fmaf(yv.x, ms, mh) > 0.f)) gv.x = 0.f;
        if (!(fmaf(yv.y, ms, mh) > 0.f)) gv.y = 0.f;
        if (!(fmaf(yv.z, ms, mh) > 0.f)) gv.z = 0.f;
        if (!(fmaf(yv.w, ms, mh) > 0.f)) gv.w = 0.f;
      }
    } else {
      const int ch = (int)((i * 4) % C);
      a = __ldg(reinterpret_cast<const float4*>(coef + ch));
      b = __ldg(reinterpret_cast<const float4*>(coef + C + ch));
      c = __ldg(reinterpret_cast<const float4*>(coef + 2 * C + ch));
      if (mscale != nullptr) {
        const float4 ms = __ldg(reinterpret_cast<const float4*>(mscale + ch)), mh = __ldg(reinterpret_cast<const float4*>(mshift + ch));
        if (!(fmaf(yv.x, ms.x, mh.x) > 0.f)) gv.x = 0.f;
        if (!(fmaf(yv.y, ms.y, mh.y) > 0.f)) gv.y = 0.f;
        if (!(fmaf(yv.z, ms.z, mh.z) > 0.f)) gv.z = 0.f;
        if (!(fmaf(yv.w, ms.w, mh.w) > 0.f)) gv.w = 0.f;
      }
    }
    const float4 v = make_float4(fmaf(a.x, gv.x, fmaf(b.x, yv.x, c.x)), fmaf(a.y, gv.y, fmaf(b.y, yv.y, c.y)),
                                 fmaf(a.z, gv.z, fmaf(b.z, yv.z, c.z)), fmaf(a.w, gv.w, fmaf(b.w, yv.w, c.w)));
    if (dy_bf) reinterpret_cast<uint2*>(dy)[i] = make_uint2(bf2(v.x, v.y), bf2(v.z, v.w));
    else reinterpret_cast<float4*>(dy)[i] = v;
  }
}

__global__ void __launch_bounds__(kNT) bn_bwd_apply_kernel(const void* g, int g_bf, const void* y, int y_bf, const void* act,
                                                           int act_bf, const float* __restrict__ mscale,
                                                           const float* __restrict__ mshift, const float* __restrict__ coef,
                                                           long long total, int C, long long inner, int to_nhwc, void* dy,
                                                           int dy_bf) {
  for (long long i = (long long)blockIdx.x * kNT + threadIdx.x; i < total; i += (long long)gridDim.x * kNT) {
    const int ch = (int)((i / inner) % C);
    float gv = ldf(g, i, g_bf);
    const float yv = ldf(y, i, y_bf);
    if (act != nullptr && !(ldf(act, i, act_bf) > 0.f)) gv = 0.f;
    if (mscale != nullptr && !(fmaf(yv, __ldg(mscale + ch), __ldg(mshift + ch)) > 0.f)) gv = 0.f;
    const float v = fmaf(__ldg(coef + ch), gv, fmaf(__ldg(coef + C + ch), yv, __ldg(coef + 2 * C + ch)));
    long long o = i;
    if (to_nhwc) {  // channel-major [b][c][hw] -> channels-last [b][hw][c] so the consumers gather 16-byte channel runs
      const long long per = (long long)C * inner, b = i / per;
      o = b * per + (i % inner) * C + ch;
    }
    stf(dy, o, dy_bf, v);
  }
}

// channels-last fast path of bn_bwd_apply (g fp32, y bf16 -> dy bf16, C % 8 == 0): 8 elements per thread, 16-byte
// accesses, the three coefficient rows staged in shared memory
__global__ void __launch_bounds__(kNT) bn_bwd_apply_vec_kernel(const float4* __restrict__ g, const uint4* __restrict__ y,
                                                               const float* __restrict__ coef, long long n8, int C,
                                                               uint4* __restrict__ dy) {
  extern __shared__ float sCoef[];  // [3][C]
  for (int i = threadIdx.x; i < 3 * C; i += kNT) sCoef[i] = __ldg(coef + i);
  __syncthreads();
  const int C8 = C >> 3;
  for (long long i = (long long)blockIdx.x * kNT + threadIdx.x; i < n8; i += (long long)gridDim.x * kNT) {
    const int c0 = (int)(i % C8) << 3;
    const float4 g0 = __ldg(g + 2 * i), g1 = __ldg(g + 2 * i + 1);
    const uint4 q = __ldg(y + i);
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float v[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 yy = __bfloat1622float2(h2[k]);
      const int c = c0 + 2 * k;
      v[2 * k] = fmaf(sCoef[c], gv[2 * k], fmaf(sCoef[C + c], yy.x, sCoef[2 * C + c]));
      v[2 * k + 1] = fmaf(sCoef[c + 1], gv[2 * k + 1], fmaf(sCoef[C + c + 1], yy.y, sCoef[2 * C + c + 1]));
    }
    uint4 o;
    __nv_bfloat162 t;
    t = __floats2bfloat162_rn(v[0], v[1]); o.x = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(v[2], v[3]); o.y = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(v[4], v[5]); o.z = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(v[6], v[7]); o.w = *reinterpret_cast<uint32_t*>(&t);
    dy[i] = o;
  }
}

// ---------------------------------------------------------------------------
// BatchNorm finalize + apply + ReLU in ONE launch (replaces bn_finalize followed by an elementwise pass): every CTA
// turns the fp64 batch moments into per-channel scale/shift in shared memory (C <= 4096 channels: a few hundred
// flops), CTA 0 also publishes scale/shift/mean/invstd for the backward and updates the running estimates; the last
// CTA to have read the moments (ticket word stored behind them: stats[2C]) clears the accumulator for the next step.
//   layout 0: raw bf16 channels-last [.., C]            -> act bf16, same layout          (conv blocks)
//   layout 1: raw bf16 channel-major [B, C*HW]          -> act bf16, same layout          (last encoder block -> heads)
//   layout 2: raw fp32 [B, C] with C = C0*HW (c0 major) -> act bf16 [B, HW, C0]           (decoder fc block, BatchNorm1d)
// ---------------------------------------------------------------------------
struct BnFaParams {
  double* stats; int C; double count; const float *gamma, *beta; float *rm, *rv; float momentum, eps; int repeat;
  float *scale, *shift; int expand; float *mean, *invstd;
  const void* raw; void* act; long long total; int HW, C0;
};


// per-channel statistics -> (scale, shift); `publish` threads also store the backward state and the running estimates.
// The mean / variance are formed in fp64 (cancellation), everything after that in fp32.
__device__ __forceinline__ void bn_channel_affine(const BnFaParams& p, int c, bool publish, float* sc_out, float* sh_out) {
  const double inv_count = 1.0 / p.count;
  const double mean = p.stats[c] * inv_count;
  double var = p.stats[p.C + c] * inv_count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = 1.f / sqrtf((float)var + p.eps);
  const float g = p.gamma ? p.gamma[c] : 1.f, b = p.beta ? p.beta[c] : 0.f;
  const float sc = g * invstd, sh = b - (float)mean * sc;
  *sc_out = sc;
  *sh_out = sh;
  if (publish) {
    for (int i = 0; i < p.expand; ++i) { p.scale[c * p.expand + i] = sc; p.shift[c * p.expand + i] = sh; }
    p.mean[c] = (float)mean;
    p.invstd[c] = invstd;
    if (p.rm) {
      const float unbiased = (float)(p.count > 1.0 ? var * p.count / (p.count - 1.0) : var);
      float rm = p.rm[c], rv = p.rv[c];
      for (int i = 0; i < p.repeat; ++i) {
        rm = (1.f - p.momentum) * rm + p.momentum * (float)mean;
        rv = (1.f - p.momentum) * rv + p.momentum * unbiased;
      }
      p.rm[c] = rm;
      p.rv[c] = rv;
    }
  }
}
// ticket: the last CTA that has read its moments clears the accumulator (and the ticket) for the next step
__device__ __forceinline__ void bn_ticket_clear(const BnFaParams& p, int* s_last) {
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y;
    *s_last = atomicAdd(reinterpret_cast<unsigned int*>(p.stats + 2 * p.C), 1u) == total - 1 ? 1 : 0;
  }
  __syncthreads();
  if (*s_last) {
    __threadfence();
    for (int i = threadIdx.x; i < 2 * p.C; i += kNT) p.stats[i] = 0.0;
    if (threadIdx.x == 0) *reinterpret_cast<unsigned int*>(p.stats + 2 * p.C) = 0u;
  }
}

template <int LAYOUT>  // 0: channels-last, 1: channel-major
__global__ void __launch_bounds__(kNT) bn_finalize_apply_kernel(const BnFaParams p) {
  extern __shared__ float sAff[];  // [2][C]
  __shared__ int s_last;
  const int C = p.C;
  const uint4* raw = reinterpret_cast<const uint4*>(p.raw);
  uint4* out = reinterpret_cast<uint4*>(p.act);
  const long long n8 = p.total >> 3;
  const int C8 = C >> 3;
  const long long stride = (long long)gridDim.x * kNT;
  long long i = (long long)blockIdx.x * kNT + threadIdx.x;
  // the first two chunks are in flight while the moments are read and turned into (scale, shift); the ticket's atomic round trip
  // overlaps the streaming loop (its result is only needed for the clear at the end)
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
  uint4 q0 = i < n8 ? __ldg(raw + i) : z4, q1 = i + stride < n8 ? __ldg(raw + i + stride) : z4;
  for (int c = threadIdx.x; c < C; c += kNT) bn_channel_affine(p, c, blockIdx.x == 0, &sAff[c], &sAff[C + c]);
  __syncthreads();               // all of this CTA's reads of the moments are done
  unsigned ticket = 0u;
  if (threadIdx.x == 0) ticket = atomicAdd(reinterpret_cast<unsigned int*>(p.stats + 2 * p.C), 1u);
  auto apply = [&](const uint4 q, long long idx) {
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
    float v[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h2[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
    if (LAYOUT == 0) {
      const int c0 = (int)(idx % C8) << 3;
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaxf(fmaf(v[k], sAff[c0 + k], sAff[C + c0 + k]), 0.f);
    } else {
      const long long e0 = idx << 3;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = (int)(((e0 + k) / p.HW) % C);
        v[k] = fmaxf(fmaf(v[k], sAff[c], sAff[C + c]), 0.f);
      }
    }
    out[idx] = make_uint4(bf2(v[0], v[1]), bf2(v[2], v[3]), bf2(v[4], v[5]), bf2(v[6], v[7]));
  };
  for (; i < n8; i += 2 * stride) {
    const long long j = i + stride, i2 = i + 2 * stride, j2 = j + 2 * stride;
    const uint4 n0 = i2 < n8 ? __ldg(raw + i2) : z4, n1 = j2 < n8 ? __ldg(raw + j2) : z4;   // next pair in flight
    apply(q0, i);
    if (j < n8) apply(q1, j);
    q0 = n0; q1 = n1;
  }
  if (threadIdx.x == 0) s_last = ticket == gridDim.x * gridDim.y - 1 ? 1 : 0;
  __syncthreads();
  if (s_last) {   // the last CTA that has read its moments clears the accumulator (and the ticket) for the next step
    __threadfence();
    for (int k = threadIdx.x; k < 2 * p.C; k += kNT) p.stats[k] = 0.0;
    if (threadIdx.x == 0) *reinterpret_cast<unsigned int*>(p.stats + 2 * p.C) = 0u;
  }
}

// layout 2 (decoder fc block): raw fp32 [B, C], C = C0*HW with column n = c0*HW + hw  ->  act bf16 [B, HW, C0].
// CTA (gx, gy) owns the 8*HW contiguous columns of channel group gx (so it forms only those 8*HW scale/shift pairs)
// and the rows gy, gy + gridDim.y, ...; thread (row slot, hw) reads 8 values HW apart — for fixed k the 16 threads of
// a row read 64 contiguous bytes — and writes one 16-byte run of the channels-last output.
__global__ void __launch_bounds__(kNT) bn_finalize_apply_fc_kernel(const BnFaParams p, int rows) {
  extern __shared__ float sAff[];  // [2][8*HW]
  __shared__ int s_last;
  const int HW = p.HW, C0 = p.C0, W = 8 * HW, n_base = blockIdx.x * W;
  for (int j = threadIdx.x; j < W; j += kNT) bn_channel_affine(p, n_base + j, blockIdx.y == 0, &sAff[j], &sAff[W + j]);
  __syncthreads();
  bn_ticket_clear(p, &s_last);
  const float* raw = reinterpret_cast<const float*>(p.raw);
  uint4* out = reinterpret_cast<uint4*>(p.act);
  const int hw = threadIdx.x % HW, slot = threadIdx.x / HW, slots = kNT / HW;
  for (int b = blockIdx.y * slots + slot; b < rows; b += gridDim.y * slots) {
    const float* rp = raw + (long long)b * p.C + n_base + hw;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = fmaxf(fmaf(__ldg(rp + k * HW), sAff[k * HW + hw], sAff[W + k * HW + hw]), 0.f);
    out[((long long)b * HW + hw) * (C0 >> 3) + blockIdx.x] = make_uint4(bf2(v[0], v[1]), bf2(v[2], v[3]), bf2(v[4], v[5]), bf2(v[6], v[7]));
  }
}

// layout 3 (last encoder block): raw bf16 channels-last [B, HW, C] -> channel-major copies [B, C*HW] of the raw tensor
// (what the heads' data-gradient epilogue and the BatchNorm backward read) and of relu(bn(raw)) (the heads' operand,
// in the reference's Flatten order, vae.py:26).  One CTA per image, transposition through shared memory.
__global__ void __launch_bounds__(kNT) bn_finalize_apply_tr_kernel(const BnFaParams p, __nv_bfloat16* __restrict__ raw_cm) {
  extern __shared__ float sAff[];  // [2][C] then the [HW][C+2] bf16 tile
  __shared__ int s_last;
  const int C = p.C, HW = p.HW, LDT = C + 2;
  __nv_bfloat16* sT = reinterpret_cast<__nv_bfloat16*>(sAff + 2 * C);
  for (int c = threadIdx.x; c < C; c += kNT) bn_channel_affine(p, c, blockIdx.x == 0, &sAff[c], &sAff[C + c]);
  __syncthreads();
  bn_ticket_clear(p, &s_last);
  const __nv_bfloat16* raw = reinterpret_cast<const __nv_bfloat16*>(p.raw);
  __nv_bfloat16* act = reinterpret_cast<__nv_bfloat16*>(p.act);
  const long long per = (long long)C * HW, nimg = p.total / per;
  for (long long b = blockIdx.x; b < nimg; b += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < (int)per / 2; i += kNT) {   // coalesced 4-byte reads of [hw][c]
      const int e = 2 * i, hw = e / C, c = e - hw * C;
      *reinterpret_cast<uint32_t*>(sT + hw * LDT + c) = __ldg(reinterpret_cast<const uint32_t*>(raw + b * per + e));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (int)per; i += kNT) {       // coalesced 2-byte writes of [c][hw]
      const int c = i / HW, hw = i - c * HW;
      const __nv_bfloat16 y = sT[hw * LDT + c];
      raw_cm[b * per + i] = y;
      act[b * per + i] = __float2bfloat16(fmaxf(fmaf(__bfloat162float(y), sAff[c], sAff[C + c]), 0.f));
    }
  }
}

// out[c] = sum over rows of x[r][c]  (bias gradients of the linear heads): 32 columns per CTA, 32 row groups
// per column (coalesced 128-byte row segments), fixed-order shared-memory reduction
__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ x, long long rows, int cols, float* out) {
  __shared__ float sp[32][33];
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (c < cols)
    for (long long r = rg; r < rows; r += 32) s += x[r * cols + c];
  sp[rg][lane] = s;
  __syncthreads();
  if (rg == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += sp[i][lane];
    out[c] = t;
  }
}

// act = relu(raw * scale[c] + shift[c]) on a channels-last bf16 tensor (C % 8 == 0): the BatchNorm-apply + ReLU
// between two tensor-core GEMMs, materialised once so that the consumers' operand loads are plain cp.async copies.
// HBM/L2 streaming: 16-byte loads and stores, scale / shift staged in shared memory.
__global__ void __launch_bounds__(kNT) bn_relu_bf16_kernel(const uint4* __restrict__ raw, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, long long n8, int C,
                                                           uint4* __restrict__ out) {
  extern __shared__ float sAff[];  // [2][C]
  for (int i = threadIdx.x; i < C; i += kNT) { sAff[i] = __ldg(scale + i); sAff[C + i] = __ldg(shift + i); }
  __syncthreads();
  const int C8 = C >> 3;
  for (long long i = (long long)blockIdx.x * kNT + threadIdx.x; i < n8; i += (long long)gridDim.x * kNT) {
    const int c0 = (int)(i % C8) << 3;
    const uint4 q = __ldg(raw + i);
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
    const float4 s0 = *reinterpret_cast<const float4*>(sAff + c0), s1 = *reinterpret_cast<const float4*>(sAff + c0 + 4);
    const float4 h0 = *reinterpret_cast<const float4*>(sAff + C + c0), h1 = *reinterpret_cast<const float4*>(sAff + C + c0 + 4);
    const float2 a = __bfloat1622float2(h2[0]), b = __bfloat1622float2(h2[1]), c = __bfloat1622float2(h2[2]), d = __bfloat1622float2(h2[3]);
    uint4 o;
    __nv_bfloat162 t;
    t = __floats2bfloat162_rn(fmaxf(fmaf(a.x, s0.x, h0.x), 0.f), fmaxf(fmaf(a.y, s0.y, h0.y), 0.f)); o.x = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(fmaxf(fmaf(b.x, s0.z, h0.z), 0.f), fmaxf(fmaf(b.y, s0.w, h0.w), 0.f)); o.y = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(fmaxf(fmaf(c.x, s1.x, h1.x), 0.f), fmaxf(fmaf(c.y, s1.y, h1.y), 0.f)); o.z = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(fmaxf(fmaf(d.x, s1.z, h1.z), 0.f), fmaxf(fmaf(d.y, s1.w, h1.w), 0.f)); o.w = *reinterpret_cast<uint32_t*>(&t);
    out[i] = o;
  }
}

// grid for the two-moment reductions: channel-blocked mapping when a channel spans >= 64 positions
inline void reduce_grid(int C, long long inner, long long nper, dim3* grid, int* blocked) {
  const long long period = (long long)C * inner;
  long long gx;
  if (inner >= 64) { *blocked = 1; gx = (long long)C * ((inner + kNT - 1) / kNT); }
  else { *blocked = 0; gx = (period + kNT - 1) / kNT; }
  long long gy = (148 * 4 + gx - 1) / gx;
  if (gy > nper) gy = nper;
  if (gy < 1) gy = 1;
  *grid = dim3((unsigned)gx, (unsigned)gy);
}
}  // namespace

extern "C" {

int clearvae_bn_finalize(double* stats, int32_t C, int32_t group, double count, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, float momentum, float eps, float* scale, float* shift,
                         int32_t expand, float* save_mean, float* save_invstd, int32_t repeat, void* stream) {
  if (!stats || !scale || !shift || !save_mean || !save_invstd || C <= 0 || group <= 0 || expand <= 0 || count <= 0 || repeat < 1) return CLEARVAE_EINVAL;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, C, group, count, gamma, beta, running_mean,
                                                                         running_var, momentum, eps, scale, shift, expand,
                                                                         save_mean, save_invstd, repeat);
  CV_LAUNCH_CHECK();
  return 0;
}

int clearvae_bn_reduce(const void* y, int32_t y_dtype, const void* g, int32_t g_dtype, const void* act, int32_t act_dtype,
                       const float* mask_scale, const float* mask_shift, int64_t total, int32_t C, int64_t inner, int32_t mode,
                       double* stats, void* stream) {
  if (!y || !stats || total <= 0 || C <= 0 || inner <= 0 || (mode == 1 && !g)) return CLEARVAE_EINVAL;
  const long long period = (long long)C * inner;
  if (total % period) return CLEARVAE_EINVAL;
  dim3 grid;
  int blocked;
  reduce_grid(C, inner, total / period, &grid, &blocked);
  bn_reduce_kernel<<<grid, kNT, 0, (cudaStream_t)stream>>>(y, y_dtype, g, g_dtype, act, act_dtype, mask_scale, mask_shift, total, C,
                                                           inner, mode, blocked, stats);
  CV_LAUNCH_CHECK();
  return 0;
}

size_t clearvae_bn_act_workspace_bytes(void) { return 256 + 148 * 8 * sizeof(float); }

int clearvae_bn_act_fwd(const void* raw, int32_t raw_dtype, const float* scale, const float* shift, int64_t total, int32_t C,
                        int64_t inner, int32_t act, int32_t to_nhwc_C, int32_t to_nhwc_HW, void* out, int32_t out_dtype,
                        const float* target, int64_t batch,
                        float* sse_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!raw || !scale || !shift || !out || total <= 0 || C <= 0 || inner <= 0) return CLEARVAE_EINVAL;
  unsigned* ticket = nullptr;
  float* partial = nullptr;
  if (target) {
    if (!sse_out || !workspace || batch <= 0) return CLEARVAE_EINVAL;
    if (workspace_bytes < clearvae_bn_act_workspace_bytes()) return CLEARVAE_EWORKSPACE;
    ticket = reinterpret_cast<unsigned*>(workspace);
    partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
  }
  bn_act_fwd_kernel<<<grid_for(total), kNT, 0, (cudaStream_t)stream>>>(raw, raw_dtype, scale, shift, total, C, inner, act, to_nhwc_C,
                                                                       to_nhwc_HW, out, out_dtype, target, target ? 1.f / (float)batch : 0.f,
                                                                       sse_out, partial, ticket);
  CV_LAUNCH_CHECK();
  return 0;
}

int clearvae_sigmoid_mse_bwd(const float* xhat, const float* x, const float* grad_recon, const float* grad_ext, const void* raw,
                             int32_t raw_dtype, int64_t total, int32_t C, int64_t inner, int64_t batch, float* g_pre,
                             double* stats, void* stream) {
  if (!xhat || !x || !raw || !g_pre || !stats || total <= 0 || C <= 0 || inner <= 0 || batch <= 0) return CLEARVAE_EINVAL;
  const long long period = (long long)C * inner;
  if (total % period) return CLEARVAE_EINVAL;
  dim3 grid;
  int blocked;
  if (raw_dtype == CLEARVAE_F32 && inner % 4 == 0 && inner >= 64 &&
      !(((uintptr_t)xhat | (uintptr_t)x | (uintptr_t)raw | (uintptr_t)g_pre | (uintptr_t)grad_ext) & 15)) {
    const long long inner4 = inner / 4, nper = total / period;
    const long long gx = (long long)C * ((inner4 + kNT - 1) / kNT);
    long long gy = (148 * 8 + gx - 1) / gx;
    gy = gy > nper ? nper : gy < 1 ? 1 : gy;
    sigmoid_mse_bwd_vec_kernel<<<dim3((unsigned)gx, (unsigned)gy), kNT, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(xhat), reinterpret_cast<const float4*>(x), grad_recon, reinterpret_cast<const float4*>(grad_ext),
        reinterpret_cast<const float4*>(raw), nper, C, inner4, 2.f / (float)batch, reinterpret_cast<float4*>(g_pre), stats);
    CV_LAUNCH_CHECK();
    return 0;
  }
  reduce_grid(C, inner, total / period, &grid, &blocked);
  sigmoid_mse_bwd_kernel<<<grid, kNT, 0, (cudaStream_t)stream>>>(xhat, x, grad_recon, grad_ext, raw, raw_dtype, total, C, inner,
                                                                 2.f / (float)batch, blocked, g_pre, stats);
  CV_LAUNCH_CHECK();
  return 0;
}

int clearvae_bn_bwd_coef(double* stats, int32_t C, int32_t group, double count, const float* gamma, const float* save_mean,
                         const float* save_invstd, float* coef, float* dgamma, float* dbeta, void* stream) {
  if (!stats || !save_mean || !save_invstd || !coef || C <= 0 || group <= 0 || count <= 0) return CLEARVAE_EINVAL;
  bn_bwd_coef_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, C, group, count, gamma, save_mean, save_invstd, coef,
                                                                         dgamma, dbeta);
  CV_LAUNCH_CHECK();
  return 0;
}

int clearvae_bn_bwd_apply(const void* g, int32_t g_dtype, const void* y, int32_t y_dtype, const void* act, int32_t act_dtype,
                          const float* mask_scale, const float* mask_shift, const float* coef, int64_t total, int32_t C,
                          int64_t inner, int32_t to_nhwc, void* dy, int32_t dy_dtype, void* stream) {
  if (!g || !y || !coef || !dy || total <= 0 || C <= 0 || inner <= 0) return CLEARVAE_EINVAL;
  if (inner == 1 && !to_nhwc && !act && !mask_scale && g_dtype == CLEARVAE_F32 && y_dtype == CLEARVAE_BF16 &&
      dy_dtype == CLEARVAE_BF16 && C % 8 == 0 && C <= 4096 && total % C == 0 &&
      !(((uintptr_t)g | (uintptr_t)y | (uintptr_t)dy) & 15)) {
    const long long n8 = total / 8;
    long long gr = (n8 + kNT - 1) / kNT;
    if (gr > 148 * 8) gr = 148 * 8;
    bn_bwd_apply_vec_kernel<<<(unsigned)gr, kNT, 3 * C * sizeof(float), (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(g), reinterpret_cast<const uint4*>(y), coef, n8, C, reinterpret_cast<uint4*>(dy));
    CV_LAUNCH_CHECK();
    return 0;
  }
  if (!act && !to_nhwc && g_dtype == CLEARVAE_F32 && y_dtype == CLEARVAE_F32 && total % 4 == 0 &&
      ((inner == 1 && C % 4 == 0) || inner % 4 == 0) &&
      !(((uintptr_t)g | (uintptr_t)y | (uintptr_t)dy | (uintptr_t)coef | (uintptr_t)mask_scale | (uintptr_t)mask_shift) & 15)) {
    const long long n4 = total / 4;
    long long gr = (n4 + kNT - 1) / kNT;
    if (gr > 148 * 8) gr = 148 * 8;
    if (inner == 1)
      bn_bwd_apply_vec4_kernel<false><<<(unsigned)gr, kNT, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(y), mask_scale, mask_shift, coef, n4, C, 1, dy, dy_dtype == CLEARVAE_BF16);
    else
      bn_bwd_apply_vec4_kernel<true><<<(unsigned)gr, kNT, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(y), mask_scale, mask_shift, coef, n4, C, inner / 4, dy, dy_dtype == CLEARVAE_BF16);
    CV_LAUNCH_CHECK();
    return 0;
  }
  bn_bwd_apply_kernel<<<grid_for(total), kNT, 0, (cudaStream_t)stream>>>(g, g_dtype, y, y_dtype, act, act_dtype, mask_scale,
                                                                         mask_shift, coef, total, C, inner, to_nhwc, dy, dy_dtype);
  CV_LAUNCH_CHECK();
  return 0;
}

int clearvae_bn_finalize_apply(double* stats, int32_t C, double count, const float* gamma, const float* beta, float* running_mean,
                               float* running_var, float momentum, float eps, int32_t repeat, float* scale, float* shift,
                               int32_t expand, float* save_mean, float* save_invstd, const void* raw, int32_t layout, int64_t total,
                               int32_t HW, void* act_bf16, void* raw_cm_bf16, void* stream) {
  if (!stats || !scale || !shift || !save_mean || !save_invstd || !raw || !act_bf16 || C <= 0 || count <= 0 || repeat < 1 ||
      expand < 1 || total <= 0 || HW < 1)
    return CLEARVAE_EINVAL;
  if (C > 4096 || layout < 0 || layout > 3 || total % 8 != 0 || (((uintptr_t)raw | (uintptr_t)act_bf16) & 15)) return CLEARVAE_EUNSUPPORTED;
  BnFaParams p{stats, C, count, gamma, beta, running_mean, running_var, momentum, eps, repeat, scale, shift, expand, save_mean,
               save_invstd, raw, act_bf16, total, HW, 0};
  cudaStream_t st = (cudaStream_t)stream;
  if (layout == 2) {
    if (C % HW != 0 || (C / HW) % 8 != 0 || total % C != 0 || kNT % HW != 0 || HW > kNT) return CLEARVAE_EUNSUPPORTED;
    p.C0 = C / HW;
    const int rows = (int)(total / C), slots = kNT / HW;
    int gy = (rows + slots * 4 - 1) / (slots * 4);   // ~4 rows per thread
    gy = gy < 1 ? 1 : gy > 64 ? 64 : gy;
    bn_finalize_apply_fc_kernel<<<dim3((unsigned)(p.C0 / 8), (unsigned)gy), kNT, 2 * 8 * HW * sizeof(float), st>>>(p, rows);
    CV_LAUNCH_CHECK();
    return 0;
  }
  if (layout == 3) {
    if (!raw_cm_bf16 || C % 2 != 0 || total % ((long long)C * HW) != 0) return CLEARVAE_EUNSUPPORTED;
    const size_t smem = 2 * (size_t)C * sizeof(float) + (size_t)HW * (C + 2) * 2;
    if (smem > 48 * 1024) return CLEARVAE_EUNSUPPORTED;
    long long g = total / ((long long)C * HW);
    if (g > 148 * 8) g = 148 * 8;
    bn_finalize_apply_tr_kernel<<<(unsigned)g, kNT, smem, st>>>(p, reinterpret_cast<__nv_bfloat16*>(raw_cm_bf16));
    CV_LAUNCH_CHECK();
    return 0;
  }
  if (layout == 0) {
    if (C % 8 != 0 || total % C != 0) return CLEARVAE_EUNSUPPORTED;
  } else {
    if (total % ((long long)C * HW) != 0) return CLEARVAE_EUNSUPPORTED;
  }
  long long g = (total / 8 + kNT * 4 - 1) / (kNT * 4);   // >= 4 vectors per thread: the per-CTA prologue is amortised
  if (g > 148 * 4) g = 148 * 4;
  if (g < 1) g = 1;
  const size_t smem = 2 * (size_t)C * sizeof(float);
  if (layout == 0) bn_finalize_apply_kernel<0><<<(unsigned)g, kNT, smem, st>>>(p);
  else bn_finalize_apply_kernel<1><<<(unsigned)g, kNT, smem, st>>>(p);
  CV_LAUNCH_CHECK();
  return 0;
}

int clearvae_colsum(const float* x, int64_t rows, int32_t cols, float* out, void* stream) {
  if (!x || !out || rows <= 0 || cols <= 0) return CLEARVAE_EINVAL;
  colsum_kernel<<<(cols + 31) / 32, 1024, 0, (cudaStream_t)stream>>>(x, rows, cols, out);
  CV_LAUNCH_CHECK();
  return 0;
}

int clearvae_bn_relu_apply(const void* raw_bf16, const float* scale, const float* shift, int64_t total, int32_t C, void* act_bf16,
                           void* stream) {
  if (!raw_bf16 || !scale || !shift || !act_bf16 || total <= 0 || C <= 0) return CLEARVAE_EINVAL;
  if (C % 8 != 0 || total % C != 0 || C > 4096 || ((uintptr_t)raw_bf16 & 15) || ((uintptr_t)act_bf16 & 15)) return CLEARVAE_EUNSUPPORTED;
  const long long n8 = total / 8;
  long long g = (n8 + kNT - 1) / kNT;
  if (g > 148 * 8) g = 148 * 8;
  bn_relu_bf16_kernel<<<(unsigned)g, kNT, 2 * C * sizeof(float), (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(raw_bf16), scale, shift, n8, C, reinterpret_cast<uint4*>(act_bf16));
  CV_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
