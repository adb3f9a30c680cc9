// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   dst[m, n] = sum_k A[m, k] * W[n, k]           m = output pixel, n = output channel,
//                                                  k = (tap, input channel)
//   * A (im2col rows) is never materialised: 128 producer threads (one per tile row)
//     gather 16-byte channel runs straight from the NHWC activation, apply the
//     previous layer's BatchNorm scale/shift + ReLU on the fly, convert to bf16 and
//     store them in the UMMA K-major "interleave" (no-swizzle) layout
//     [k-chunk][row][8 x bf16]: conflict-free 16-byte shared stores.
//   * W comes from a packed bf16 [N][Kp] copy through TMA (128-byte swizzle).
//   * one elected thread issues tcgen05.mma (M = 128, N = BN, K = 16) into a TMEM
//     accumulator; tcgen05.commit releases shared-memory stages / signals the epilogue.
//   * the epilogue reads TMEM (tcgen05.ld 32x32b), adds bias, writes the raw output and
//     reduces the BatchNorm batch statistics (or the ReLU/BN-backward sums) with a
//     transposing warp-shuffle reduction + one double atomicAdd per column per warp.
// 4-stage mbarrier pipeline, warp roles: 0-3 gather + epilogue, 4 TMA, 5 MMA.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "conv_plan.h"
#include "sm100_ptx.cuh"

namespace {

using namespace sm100;
using cvplan::Cls;
using cvplan::Plan;

constexpr int BM = 128, BK = 64, NS = 4;
constexpr int kAStage = BM * BK * 2;  // 16 KiB
constexpr int kThreads = 192;
constexpr int kTabMax = 256;    // entries of the k -> (tap, channel) table used by the scalar gather
constexpr int kMaxPreC = 2048;  // source channels whose BatchNorm scale/shift are staged in shared memory

// optional per-CTA phase timeline (tools/conv_timeline.py): 8 x int64 per CTA, written by one thread per role
__device__ long long* g_conv_timeline = nullptr;
__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define CV_TL(slot)                                                                                          \
  do {                                                                                                       \
    if (tl_buf) tl_buf[(((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (slot)] = gtimer(); \
  } while (0)

struct TmapPack {
  CUtensorMap t[cvplan::kMaxClasses];
};

struct GemmParams {
  Plan plan;
  long long batch;
  const void* src; long long s_n, s_h, s_w, s_c; int src_bf16;
  const float *pre_scale, *pre_shift; int pre_relu;
  const float* bias;
  void* dst; long long d_n, d_h, d_w, d_c; int dst_bf16;
  int epi;
  const void* msk; long long m_n, m_h, m_w, m_c; int msk_bf16;
  const float *msk_scale, *msk_shift;
  double* stats;
  int splits;  // split-K over grid.z (fp32 atomic epilogue); 1 = off
  int x3;      // CLEARVAE_ROLE_SPLIT3: every k-block runs kSplitPasses times over the bf16 parts of A and W — fp32-grade products
  // persistent kernels: flattened (class, n-tile, m-tile) work list
  int tile_start[cvplan::kMaxClasses + 1];
  int n_tiles;
  // TMA-operand kernel: an m-tile of class c is a box of tn images x th rows x tw (= Wd) columns of the class-local output grid
  // (rows of the GEMM tile in that order, tn * th * tw <= 128); tiles_h = m-tiles per image group along H, mtiles = m-tiles per n-tile
  int t_tw[cvplan::kMaxClasses], t_th[cvplan::kMaxClasses], t_tn[cvplan::kMaxClasses], t_tiles_h[cvplan::kMaxClasses],
      t_mtiles[cvplan::kMaxClasses];
  int t_cb;        // channels per pipeline stage: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B)
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// bf16 split of an fp32 value into three parts, x = p0 + p1 + p2 (+ <= 2^-25 |x|): p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1).
// split_part(x, k) returns the fp32 value whose bf16 rounding is part k.
__device__ __forceinline__ float split_part(float x, int k) {
  if (k == 0) return x;
  const float r1 = x - __bfloat162float(__float2bfloat16(x));
  if (k == 1) return r1;
  return r1 - __bfloat162float(__float2bfloat16(r1));
}
// Split mode (CLEARVAE_ROLE_SPLIT): every k-block runs six times, accumulating the products of operand parts
// (a, w) = (0,0) (0,1) (1,0) (1,1) (0,2) (2,0) — everything down to 2^-24 of |a||w|, i.e. fp32-grade products on bf16 tensor cores.
constexpr int kSplitPasses = 6;
// The tensor core adds into its fp32 TMEM accumulator with truncation at the accumulator's exponent.  Split mode therefore keeps
// three accumulators by magnitude class — (0,0) | (0,1),(1,0) | (1,1),(0,2),(2,0) — kAccStride columns apart, so the small
// products are neither truncated at the big sum's ulp nor lengthen its chain; the epilogue adds the three in fp32.
__device__ __forceinline__ int split_acc(int sub) { return sub == 0 ? 0 : sub <= 2 ? 1 : 2; }
__device__ __forceinline__ void split_merge32(uint32_t (&raw)[32], uint32_t taddr, uint32_t acc_stride) {
  uint32_t r1[32], r2[32];
  sm100::tmem_ld32(taddr + acc_stride, r1);
  sm100::tmem_ld32(taddr + 2 * acc_stride, r2);
  sm100::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) + (__uint_as_float(r1[i]) + __uint_as_float(r2[i])));
}
__device__ __forceinline__ void split_merge16(uint32_t (&raw)[16], uint32_t taddr, uint32_t acc_stride) {
  uint32_t r1[16], r2[16];
  sm100::tmem_ld16(taddr + acc_stride, r1);
  sm100::tmem_ld16(taddr + 2 * acc_stride, r2);
  sm100::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) + (__uint_as_float(r1[i]) + __uint_as_float(r2[i])));
}
__device__ __forceinline__ int split_a_part(int sub) { return (0x201100 >> (4 * sub)) & 3; }   // 0,0,1,1,0,2
__device__ __forceinline__ int split_b_part(int sub) { return (0x021010 >> (4 * sub)) & 3; }   // 0,1,0,1,2,0
__device__ __forceinline__ float ld_elem(const void* base, long long off, int is_bf16) {
  return is_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[off]) : reinterpret_cast<const float*>(base)[off];
}

// column sums of a 32-row x 32-column register tile held one row per lane:
// after the call v[0] of lane j is sum over lanes of their v[j]
__device__ __forceinline__ void transpose_reduce32(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kThreads) conv_tc_kernel(const __grid_constant__ TmapPack tm, const __grid_constant__ TmapPack tm_p1, const __grid_constant__ TmapPack tm_p2, const GemmParams p) {
  constexpr int kBStage = BN * BK * 2;
  constexpr uint32_t kAccStride = BN < 32 ? 32 : BN;
  const uint32_t kTmemCols = p.x3 ? 4 * kAccStride : kAccStride;   // split mode: three accumulators (allocation is a power of two)
  extern __shared__ unsigned char smem_raw[];
  // 128-byte-swizzle atoms need 1024-byte alignment (the launch reserves the slack)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;
  unsigned char* sB = smem + NS * kAStage;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + NS * kBStage);
  uint64_t* empty = full + NS;
  uint64_t* tmem_full = empty + NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  int4* sTab = reinterpret_cast<int4*>(reinterpret_cast<unsigned char*>(full) + 96);  // 16-byte aligned; k -> (tap, channel)
  int* sTapOff = reinterpret_cast<int*>(sTab + kTabMax);                // per-tap element offset (16 entries)
  float* sScale = reinterpret_cast<float*>(sTapOff + cvplan::kMaxTaps); // pre-op scale / shift of the source channels
  float* sShift = sScale + kMaxPreC;

  long long* const tl_buf = g_conv_timeline;
  if (threadIdx.x == 0) CV_TL(0);
  const int cls_id = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const Cls& c = p.plan.cls[cls_id];
  const long long Mc = p.batch * c.Hd * c.Wd;
  const long long m0 = (long long)blockIdx.x * BM;
  if (m0 >= Mc) return;
  const int n0 = blockIdx.y * BN;
  const int nkb_all = c.Kp / BK;
  const int kb0 = nkb_all * split / p.splits;
  const int nkb = nkb_all * (split + 1) / p.splits - kb0;
  if (nkb <= 0) return;
  const int nsub = p.x3 ? kSplitPasses : 1;   // passes per k-block (split mode: six part products)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], BM + 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    if (lane == 0) { prefetch_tmap(&tm.t[cls_id]); if (p.x3) { prefetch_tmap(&tm_p1.t[cls_id]); prefetch_tmap(&tm_p2.t[cls_id]); } }
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  {
    const int Cs = p.plan.Cs;
    if (p.pre_scale != nullptr)
      for (int i = threadIdx.x; i < Cs; i += kThreads) { sScale[i] = __ldg(p.pre_scale + i); sShift[i] = __ldg(p.pre_shift + i); }
    const int Kreal = c.ntaps * Cs;
    // k -> (tap, channel, element offset of (tap, channel) relative to the row's base pixel); thread independent
    for (int k = threadIdx.x; k < kTabMax && k < Kreal; k += kThreads) {
      const int t = k / Cs, ch = k - t * Cs;
      sTab[k] = make_int4(t, ch, (int)(c.dh[t] * p.s_h + c.dw[t] * p.s_w + ch * p.s_c), 0);
    }
    if (threadIdx.x < cvplan::kMaxTaps)
      sTapOff[threadIdx.x] = threadIdx.x < c.ntaps ? (int)(c.dh[threadIdx.x] * p.s_h + c.dw[threadIdx.x] * p.s_w) : 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) CV_TL(1);

  if (warp < 4) {
    // ================= A producer: one thread per tile row =================
    const int r = threadIdx.x;
    const long long m = m0 + r;
    const bool mvalid = m < Mc;
    const long long mm = mvalid ? m : 0;
    const int wd = (int)(mm % c.Wd);
    const int hd = (int)((mm / c.Wd) % c.Hd);
    const long long img = mm / ((long long)c.Wd * c.Hd);
    const int hbase = hd * p.plan.sh, wbase = wd * p.plan.sh;
    const int Cs = p.plan.Cs, Kreal = c.ntaps * Cs;
    const bool vec = (Cs % 8 == 0) && p.s_c == 1;
    const long long img_off = img * p.s_n;
    const long long base = img_off + hbase * p.s_h + wbase * p.s_w;  // element offset of this row's (dh = dw = 0) pixel
    unsigned tapmask = 0;                                            // taps whose source pixel is inside the image
    if (mvalid) {
      for (int t = 0; t < c.ntaps; ++t) {
        const int hs = hbase + c.dh[t], ws = wbase + c.dw[t];
        tapmask |= (hs >= 0 && hs < p.plan.Hs && ws >= 0 && ws < p.plan.Ws) ? (1u << t) : 0u;
      }
    }
    // incremental (tap, channel) of the next 8-chunk (vector path)
    int t_cur = vec ? (kb0 * BK) / Cs : 0, c_cur = vec ? (kb0 * BK) % Cs : 0;
    const bool has_pre = p.pre_scale != nullptr;
    // Pure-copy operand (already activated bf16, channels contiguous): cp.async straight into the interleave layout,
    // zero-fill for padding taps / K padding; a k-block is announced two iterations after it was issued, so up to
    // three k-blocks of loads are in flight per thread and nothing is staged through registers.
    const bool cpa = vec && p.src_bf16 && !has_pre && !p.pre_relu && !p.x3;
    if (cpa) {
      const __nv_bfloat16* srcb = reinterpret_cast<const __nv_bfloat16*>(p.src);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % NS;
        mbar_wait(&empty[s], ((kb / NS) & 1) ^ 1);
        const uint32_t dst = smem_u32(sA + s * kAStage) + r * 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int kg = (kb0 + kb) * BK + j * 8;
          const bool v = kg < Kreal && ((tapmask >> t_cur) & 1u);
          const long long off = v ? base + sTapOff[t_cur & (cvplan::kMaxTaps - 1)] + c_cur : 0;
          cp_async16(dst + j * (BM * 16), srcb + off, v ? 16u : 0u);
          c_cur += 8;
          if (c_cur >= Cs) { c_cur = 0; ++t_cur; }
        }
        cp_async_commit();
        if (kb >= 2) {
          cp_async_wait<2>();
          fence_proxy_async();
          mbar_arrive(&full[(kb - 2) % NS]);
        }
      }
      if (nkb >= 2) {
        cp_async_wait<1>();
        fence_proxy_async();
        mbar_arrive(&full[(nkb - 2) % NS]);
      }
      if (threadIdx.x == 0) CV_TL(2);
      cp_async_wait<0>();
      fence_proxy_async();
      mbar_arrive(&full[(nkb - 1) % NS]);
      if (threadIdx.x == 0) CV_TL(3);
    }
    for (int it = 0; it < (cpa ? 0 : nkb * nsub); ++it) {
      const int kb = it / nsub, sub = it - kb * nsub;
      const int a_part = split_a_part(sub);  // split mode: which bf16 part of A this pass carries
      const int s = it % NS;
      unsigned char* a_st = sA + s * kAStage;
      const int t_save = t_cur, c_save = c_cur;
      if (vec) {
        // ---- phase 1: addresses + ALL loads of the k-block in flight (clamped address, masked afterwards)
        uint4 q0[8], q1[8];
        int cch[8];
        bool ok[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int kg = (kb0 + kb) * BK + j * 8;
          const bool v = kg < Kreal && ((tapmask >> t_cur) & 1u);
          const long long off = v ? base + sTapOff[t_cur & (cvplan::kMaxTaps - 1)] + c_cur : 0;
          ok[j] = v;
          cch[j] = c_cur;
          if (p.src_bf16) {
            q0[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.src) + off));
          } else {
            q0[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.src) + off));
            q1[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.src) + off + 4));
          }
          c_cur += 8;
          if (c_cur >= Cs) { c_cur = 0; ++t_cur; }
        }
        mbar_wait(&empty[s], ((it / NS) & 1) ^ 1);
        if (sub + 1 < nsub) { t_cur = t_save; c_cur = c_save; }   // the next pass walks the same k-block again
        // ---- phase 2: BatchNorm-apply + ReLU (scale/shift from shared memory), bf16 pack, 16-byte stores
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v[8];
          if (p.src_bf16) {
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q0[j]);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h2[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
          } else {
            v[0] = __uint_as_float(q0[j].x); v[1] = __uint_as_float(q0[j].y); v[2] = __uint_as_float(q0[j].z); v[3] = __uint_as_float(q0[j].w);
            v[4] = __uint_as_float(q1[j].x); v[5] = __uint_as_float(q1[j].y); v[6] = __uint_as_float(q1[j].z); v[7] = __uint_as_float(q1[j].w);
          }
          if (has_pre) {
            const float4 s0 = *reinterpret_cast<const float4*>(sScale + cch[j]);
            const float4 s1 = *reinterpret_cast<const float4*>(sScale + cch[j] + 4);
            const float4 h0 = *reinterpret_cast<const float4*>(sShift + cch[j]);
            const float4 h1 = *reinterpret_cast<const float4*>(sShift + cch[j] + 4);
            v[0] = fmaf(v[0], s0.x, h0.x); v[1] = fmaf(v[1], s0.y, h0.y); v[2] = fmaf(v[2], s0.z, h0.z); v[3] = fmaf(v[3], s0.w, h0.w);
            v[4] = fmaf(v[4], s1.x, h1.x); v[5] = fmaf(v[5], s1.y, h1.y); v[6] = fmaf(v[6], s1.z, h1.z); v[7] = fmaf(v[7], s1.w, h1.w);
          }
          if (p.pre_relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (a_part) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = split_part(v[i], a_part);
          }
          uint4 out = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          if (!ok[j]) out = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(a_st + j * (BM * 16) + r * 16) = out;
        }
      } else {
        mbar_wait(&empty[s], ((it / NS) & 1) ^ 1);
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
          const int kg = (kb0 + kb) * BK + j * 8;
          if (kg >= Kreal || tapmask == 0u) {  // zero padding of K (or a row outside the problem)
            *reinterpret_cast<uint4*>(a_st + j * (BM * 16) + r * 16) = make_uint4(0u, 0u, 0u, 0u);
            continue;
          }
          float v[8];
          bool ok[8];
          int chn[8];
          // phase 1: eight independent loads (k -> (tap, channel, offset) from the shared-memory table)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int k = kg + i;
            bool vld = false;
            long long off = 0;
            int ch = 0;
            if (k < Kreal) {
              int t, koff;
              if (k < kTabMax) { const int4 e = sTab[k]; t = e.x; ch = e.y; koff = e.z; }
              else { t = k / Cs; ch = k - t * Cs; koff = (int)(c.dh[t] * p.s_h + c.dw[t] * p.s_w + ch * p.s_c); }
              vld = (tapmask >> t) & 1u;
              off = vld ? base + koff : 0;
            }
            ok[i] = vld;
            chn[i] = ch;
            v[i] = ld_elem(p.src, off, p.src_bf16);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float x = v[i];
            if (has_pre) x = fmaf(x, sScale[chn[i]], sShift[chn[i]]);
            if (p.pre_relu) x = fmaxf(x, 0.f);
            if (a_part) x = split_part(x, a_part);
            v[i] = ok[i] ? x : 0.f;
          }
          *reinterpret_cast<uint4*>(a_st + j * (BM * 16) + r * 16) =
              make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        }
      }
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }

    // ================= epilogue: TMEM -> registers -> global =================
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    if (threadIdx.x == 0) CV_TL(4);
    const long long dst_off = img * p.d_n + (long long)(hd * p.plan.os + c.oa) * p.d_h + (long long)(wd * p.plan.os + c.ob) * p.d_w;
    const long long msk_off = img * p.m_n + (long long)(hd * p.plan.os + c.oa) * p.m_h + (long long)(wd * p.plan.os + c.ob) * p.m_w;
    const int Nn = p.plan.Nn;
    constexpr int CH = BN < 32 ? BN : 32;
    // Fast epilogue (channels-last destination, full column tile): vector stores straight from the TMEM registers,
    // and the per-channel sums through a padded shared-memory transposition — thread (column, row-quarter) adds 32
    // rows of its column, squares formed on the fly — ~170 instructions per 32-column chunk instead of the ~1100
    // of the generic path below (two warp-butterflies + per-element predicates), which bounded the whole kernel.
    const bool fast = p.splits == 1 && p.d_c == 1 && n0 + BN <= Nn && ((p.d_n | p.d_h | p.d_w) & 7) == 0 &&
                      (p.epi == CLEARVAE_EPI_BIAS_STATS || (p.m_c == 1 && ((p.m_n | p.m_h | p.m_w) & 7) == 0));
    if (fast) {
      constexpr int LD = 36;                               // padded row stride (floats): conflict-free both ways
      float* sEp = reinterpret_cast<float*>(sA);           // the operand stages are dead once tmem_full has fired
      float* sEp2 = sEp + BM * LD;
      float* sRed = sEp2 + BM * LD;                        // [2][4][32]
      const bool masked = p.epi != CLEARVAE_EPI_BIAS_STATS;
#pragma unroll 1
      for (int ch0 = 0; ch0 < BN; ch0 += CH) {
        uint32_t raw[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)ch0;
        if (CH == 32) {
          tmem_ld32(taddr, raw);
          if (p.x3) { tmem_ld_wait(); split_merge32(raw, taddr, kAccStride); }
        } else {
          uint32_t r16[16];
          tmem_ld16(taddr, r16);
          if (p.x3) { tmem_ld_wait(); split_merge16(r16, taddr, kAccStride); }
#pragma unroll
          for (int i = 0; i < 16; ++i) { raw[i] = r16[i]; raw[i + 16] = 0u; }
        }
        const int nb = n0 + ch0;
        float v[CH], u[CH];
        if (!masked) {
          float bsv[CH];
#pragma unroll
          for (int i = 0; i < CH; ++i) bsv[i] = p.bias != nullptr ? __ldg(p.bias + nb + i) : 0.f;
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < CH; ++i) v[i] = mvalid ? __uint_as_float(raw[i]) + bsv[i] : 0.f;
        } else {
          float y[CH];
          if (p.msk_bf16) {
            const uint4* mp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.msk) + (mvalid ? msk_off + nb : 0));
#pragma unroll
            for (int i = 0; i < CH / 8; ++i) {
              const uint4 q = __ldg(mp + i);
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
              for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h2[k]); y[8 * i + 2 * k] = f.x; y[8 * i + 2 * k + 1] = f.y; }
            }
          } else {
            const float4* mp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.msk) + (mvalid ? msk_off + nb : 0));
#pragma unroll
            for (int i = 0; i < CH / 4; ++i) { const float4 q = __ldg(mp + i); y[4 * i] = q.x; y[4 * i + 1] = q.y; y[4 * i + 2] = q.z; y[4 * i + 3] = q.w; }
          }
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const float act = p.msk_scale ? fmaf(y[i], __ldg(p.msk_scale + nb + i), __ldg(p.msk_shift + nb + i)) : y[i];
            const float a = (mvalid && act > 0.f) ? __uint_as_float(raw[i]) : 0.f;
            v[i] = a;
            u[i] = a * y[i];
          }
        }
        if (mvalid) {
          if (p.dst_bf16) {
            uint4* d = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.dst) + dst_off + nb);
#pragma unroll
            for (int i = 0; i < CH / 8; ++i)
              d[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                                pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
          } else {
            float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.dst) + dst_off + nb);
#pragma unroll
            for (int i = 0; i < CH / 4; ++i) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
        if (p.stats != nullptr) {
          float4* row = reinterpret_cast<float4*>(sEp + r * LD);
#pragma unroll
          for (int i = 0; i < CH / 4; ++i) row[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          if (masked) {
            float4* row2 = reinterpret_cast<float4*>(sEp2 + r * LD);
#pragma unroll
            for (int i = 0; i < CH / 4; ++i) row2[i] = make_float4(u[4 * i], u[4 * i + 1], u[4 * i + 2], u[4 * i + 3]);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          float s1 = 0.f, s2 = 0.f;
          if (lane < CH) {
            const float* col = sEp + (warp * 32) * LD + lane;
            if (!masked) {
#pragma unroll 8
              for (int rr = 0; rr < 32; ++rr) { const float a = col[rr * LD]; s1 += a; s2 = fmaf(a, a, s2); }
            } else {
              const float* col2 = sEp2 + (warp * 32) * LD + lane;
#pragma unroll 8
              for (int rr = 0; rr < 32; ++rr) { s1 += col[rr * LD]; s2 += col2[rr * LD]; }
            }
          }
          sRed[warp * 32 + lane] = s1;
          sRed[128 + warp * 32 + lane] = s2;
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (warp == 0 && lane < CH) {
            const float a = (sRed[lane] + sRed[32 + lane]) + (sRed[64 + lane] + sRed[96 + lane]);
            const float b = (sRed[128 + lane] + sRed[160 + lane]) + (sRed[192 + lane] + sRed[224 + lane]);
            atomicAdd(p.stats + nb + lane, (double)a);
            atomicAdd(p.stats + Nn + nb + lane, (double)b);
          }
        }
      }
    }
#pragma unroll 1
    for (int ch0 = fast ? BN : 0; ch0 < BN; ch0 += CH) {
      uint32_t raw[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)ch0;
      if (CH == 32) {
        tmem_ld32(taddr, raw);
        if (p.x3) { tmem_ld_wait(); split_merge32(raw, taddr, kAccStride); }
      } else {
        uint32_t r16[16];
        tmem_ld16(taddr, r16);
        if (p.x3) { tmem_ld_wait(); split_merge16(r16, taddr, kAccStride); }
#pragma unroll
        for (int i = 0; i < 16; ++i) { raw[i] = r16[i]; raw[i + 16] = 0u; }
      }
      tmem_ld_wait();
      float v[32], u[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int n = n0 + ch0 + i;
        float a = __uint_as_float(raw[i]);
        float second = 0.f;
        const bool live = mvalid && i < CH && n < Nn;
        if (p.epi == CLEARVAE_EPI_BIAS_STATS) {
          if (live && p.bias != nullptr && split == 0) a += __ldg(p.bias + n);
          if (!live) a = 0.f;
          second = a * a;
        } else {
          if (live) {
            const float y = ld_elem(p.msk, msk_off + n * p.m_c, p.msk_bf16);
            const float act = p.msk_scale ? fmaf(y, __ldg(p.msk_scale + n), __ldg(p.msk_shift + n)) : y;
            a = act > 0.f ? a : 0.f;
            second = a * y;
          } else {
            a = 0.f;
          }
        }
        v[i] = a;
        u[i] = second;
      }
      // ---- store (vectorised when channels are the innermost dst dimension)
      if (mvalid && p.splits > 1) {
        float* d = reinterpret_cast<float*>(p.dst) + dst_off;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int n = n0 + ch0 + i;
          if (n < Nn) atomicAdd(d + n * p.d_c, v[i]);
        }
      } else if (mvalid) {
        if (p.d_c == 1 && n0 + ch0 + CH <= Nn && (CH % 8) == 0) {
          if (p.dst_bf16) {
            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.dst) + dst_off + n0 + ch0;
#pragma unroll
            for (int i = 0; i < CH; i += 8)
              *reinterpret_cast<uint4*>(d + i) = make_uint4(pack_bf16(v[i], v[i + 1]), pack_bf16(v[i + 2], v[i + 3]),
                                                            pack_bf16(v[i + 4], v[i + 5]), pack_bf16(v[i + 6], v[i + 7]));
          } else {
            float* d = reinterpret_cast<float*>(p.dst) + dst_off + n0 + ch0;
#pragma unroll
            for (int i = 0; i < CH; i += 4) *reinterpret_cast<float4*>(d + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const int n = n0 + ch0 + i;
            if (n < Nn) {
              if (p.dst_bf16) reinterpret_cast<__nv_bfloat16*>(p.dst)[dst_off + n * p.d_c] = __float2bfloat16(v[i]);
              else reinterpret_cast<float*>(p.dst)[dst_off + n * p.d_c] = v[i];
            }
          }
        }
      }
      // ---- per-channel statistics: warp transpose-reduce, then across the 4 epilogue warps in shared
      //      memory, so one double atomic per column per CTA reaches L2
      if (p.stats != nullptr) {
        transpose_reduce32(v);
        transpose_reduce32(u);
        float* sSt = sScale;  // the pre-op staging area is dead once the producers are done with the main loop
        asm volatile("bar.sync 1, 128;" ::: "memory");
        sSt[(warp * 2 + 0) * 32 + lane] = v[0];
        sSt[(warp * 2 + 1) * 32 + lane] = u[0];
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 0) {
          const int n = n0 + ch0 + lane;
          if (lane < CH && n < Nn) {
            const float a = (sSt[lane] + sSt[64 + lane]) + (sSt[128 + lane] + sSt[192 + lane]);
            const float b = (sSt[32 + lane] + sSt[96 + lane]) + (sSt[160 + lane] + sSt[224 + lane]);
            atomicAdd(p.stats + n, (double)a);
            atomicAdd(p.stats + Nn + n, (double)b);
          }
        }
      }
    }
  } else if (warp == 4) {
    // ================= B producer: TMA =================
    if (lane == 0) {
      for (int it = 0; it < nkb * nsub; ++it) {
        const int kb = it / nsub, sub = it - kb * nsub;
        const int s = it % NS;
        mbar_wait(&empty[s], ((it / NS) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], kBStage);
        const int bp = p.x3 ? split_b_part(sub) : 0;
        tma_load_2d(sB + s * kBStage, bp == 0 ? &tm.t[cls_id] : bp == 1 ? &tm_p1.t[cls_id] : &tm_p2.t[cls_id], &full[s], (kb0 + kb) * BK, n0);
      }
    }
  } else {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc(kFmtBF16, BM, BN, 0, 0);
      for (int kb = 0; kb < nkb * nsub; ++kb) {   // split mode: six passes per k-block into three accumulators
        const int s = kb % NS;
        mbar_wait(&full[s], (kb / NS) & 1);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + s * kAStage), b_base = smem_u32(sB + s * kBStage);
        const int sub = kb % nsub;
        const uint32_t acc = p.x3 ? (uint32_t)split_acc(sub) * kAccStride : 0u;
        const bool first = kb < nsub && (sub == 0 || sub == 1 || sub == 3);   // first pass that writes this accumulator
#pragma unroll
        for (int k4 = 0; k4 < BK / 16; ++k4) {
          // A: interleave layout, 16-byte k-chunks BM*16 bytes apart (LBO), 8-row groups 128 bytes apart (SBO)
          const uint64_t ad = smem_desc(a_base + k4 * 2 * (BM * 16), BM * 16, 128, kLayoutNone);
          // B: 128-byte swizzled rows, 8-row groups 1024 bytes apart; K advance = +32 bytes inside the atom
          const uint64_t bd = smem_desc(b_base + k4 * 32, 16, 1024, kLayoutSw128);
          umma_f16(tmem_base + acc, ad, bd, idesc, (first && k4 == 0) ? 0u : 1u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
  }
  if (threadIdx.x == 0) CV_TL(5);
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, kTmemCols);
  if (threadIdx.x == 128) CV_TL(6);
}



// ---------------------------------------------------------------------------------------------------------
// Persistent, warp-specialised variant for the common case (bf16 channels-last operand without pre-op, channels-last
// destination, full column tiles): each CTA walks a round-robin list of (class, n-tile, m-tile) work items;
//   warps 0-3  cp.async the A operand (one tile row per thread), announcing k-blocks LAG iterations after issue,
//   warp  4    TMA for the packed weights,          warp 5   tcgen05.mma issuer,
//   warps 6-9  epilogue (TMEM -> registers -> vector stores + per-channel sums),
// coupled only through mbarriers: a shared-memory stage ring (full/empty) and a double-buffered TMEM accumulator
// (acc_full/acc_empty), so the epilogue of tile i overlaps the loads + MMAs of tile i+1, the ~1.2 us of per-CTA
// set-up/tear-down is paid once per SM instead of once per tile, and the batch statistics leave the SM as one
// double atomic per column per CTA instead of one per tile.
constexpr int kPThreads = 320;
constexpr int kEpLd = 36;   // padded row stride (floats) of the epilogue staging tile

template <int BN>
constexpr int persist_stages() { return BN >= 128 ? 4 : 3; }  // BN = 128 runs one CTA per SM (registers), so it can afford the deeper ring
template <int BN, bool MASKED>
constexpr size_t persist_smem() {
  return (size_t)persist_stages<BN>() * (kAStage + BN * BK * 2) + (MASKED ? 2 : 1) * BM * kEpLd * 4 + 2 * 4 * BN * 4 +
         cvplan::kMaxClasses * cvplan::kMaxTaps * 4 + 256 + 1024;
}

struct PTile {
  int cls, n0, nkb;
  long long m0;
};
__device__ __forceinline__ bool p_get_tile(const GemmParams& p, int id, int BN, PTile* t) {
  if (id >= p.n_tiles) return false;
  int cls = 0;
#pragma unroll
  for (int i = 1; i < cvplan::kMaxClasses; ++i)
    if (i < p.plan.n_classes && id >= p.tile_start[i]) cls = i;
  const Cls& c = p.plan.cls[cls];
  const int local = id - p.tile_start[cls];
  const int mtiles = (int)((p.batch * c.Hd * c.Wd + BM - 1) / BM);
  const int nt = local / mtiles;
  t->cls = cls;
  t->n0 = nt * BN;
  t->m0 = (long long)(local - nt * mtiles) * BM;
  t->nkb = c.Kp / BK;
  return true;
}

template <int BN, bool MASKED>
__global__ void __launch_bounds__(kPThreads, (BN <= 64 ? 2 : 1)) conv_tc_persist_kernel(const __grid_constant__ TmapPack tm, const GemmParams p) {
  constexpr int NSP = persist_stages<BN>();
  constexpr int LAG = NSP - 1;
  constexpr int kBStage = BN * BK * 2;
  constexpr uint32_t kAccCols = BN < 32 ? 32 : BN;
  constexpr uint32_t kTmemCols = 2 * kAccCols;
  constexpr int CH = BN < 32 ? BN : 32;
  constexpr int NCH = BN / CH;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;
  unsigned char* sB = sA + NSP * kAStage;
  float* sEp = reinterpret_cast<float*>(sB + NSP * kBStage);
  float* sEp2 = sEp + BM * kEpLd;
  float* sRed = sEp + (MASKED ? 2 : 1) * BM * kEpLd;                 // [2][4][BN]
  int* sTapOff = reinterpret_cast<int*>(sRed + 2 * 4 * BN);          // [classes][16]
  uint64_t* full = reinterpret_cast<uint64_t*>(sTapOff + cvplan::kMaxClasses * cvplan::kMaxTaps);
  uint64_t* empty = full + NSP;
  uint64_t* acc_full = empty + NSP;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Cs = p.plan.Cs;
  long long* const tl_buf = g_conv_timeline ? g_conv_timeline + (long long)blockIdx.x * 64 : nullptr;   // debug timeline
#define CV_PTL(tile_no, k) do { if (tl_buf && (tile_no) < 15) tl_buf[4 + (tile_no) * 4 + (k)] = gtimer(); } while (0)
  if (tl_buf && threadIdx.x == 0) tl_buf[0] = gtimer();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NSP; ++s) { mbar_init(&full[s], BM + 1); mbar_init(&empty[s], 1); }
#pragma unroll
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
    fence_barrier_init();
  }
  if (warp == 4) {
    if (lane == 0)
      for (int i = 0; i < p.plan.n_classes; ++i) prefetch_tmap(&tm.t[i]);
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < cvplan::kMaxClasses * cvplan::kMaxTaps; i += kPThreads) {
    const int ci = i / cvplan::kMaxTaps, t = i % cvplan::kMaxTaps;
    const Cls& c = p.plan.cls[ci];
    sTapOff[i] = (ci < p.plan.n_classes && t < c.ntaps) ? (int)(c.dh[t] * p.s_h + c.dw[t] * p.s_w) : 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ================= A producers =================
    const int r = threadIdx.x;
    const __nv_bfloat16* srcb = reinterpret_cast<const __nv_bfloat16*>(p.src);
    uint32_t kbg = 0, kbs = 0;  // k-blocks issued / announced
    PTile tl;
    int tno = 0;
    for (int tile = blockIdx.x; p_get_tile(p, tile, BN, &tl); tile += gridDim.x, ++tno) {
      if (threadIdx.x == 0) CV_PTL(tno, 0);
      const Cls& c = p.plan.cls[tl.cls];
      const int Kreal = c.ntaps * Cs;
      const long long Mc = p.batch * c.Hd * c.Wd;
      const long long m = tl.m0 + r;
      const bool mvalid = m < Mc;
      const long long mm = mvalid ? m : 0;
      const int wd = (int)(mm % c.Wd), hd = (int)((mm / c.Wd) % c.Hd);
      const long long img = mm / ((long long)c.Wd * c.Hd);
      const int hbase = hd * p.plan.sh, wbase = wd * p.plan.sh;
      const long long base = img * p.s_n + hbase * p.s_h + wbase * p.s_w;
      unsigned tapmask = 0;
      if (mvalid) {
#pragma unroll 1
        for (int t = 0; t < c.ntaps; ++t) {
          const int hs = hbase + c.dh[t], ws = wbase + c.dw[t];
          tapmask |= (hs >= 0 && hs < p.plan.Hs && ws >= 0 && ws < p.plan.Ws) ? (1u << t) : 0u;
        }
      }
      const int* tapoff = sTapOff + tl.cls * cvplan::kMaxTaps;
      int t_cur = 0, c_cur = 0;
      for (int kb = 0; kb < tl.nkb; ++kb) {
        const int s = kbg % NSP;
        mbar_wait(&empty[s], ((kbg / NSP) & 1) ^ 1);
        const uint32_t dst = smem_u32(sA + s * kAStage) + r * 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int kg = kb * BK + j * 8;
          const bool v = kg < Kreal && ((tapmask >> t_cur) & 1u);
          const long long off = v ? base + tapoff[t_cur & (cvplan::kMaxTaps - 1)] + c_cur : 0;
          cp_async16(dst + j * (BM * 16), srcb + off, v ? 16u : 0u);
          c_cur += 8;
          if (c_cur >= Cs) { c_cur = 0; ++t_cur; }
        }
        cp_async_commit();
        ++kbg;
        if (kbg - kbs > (uint32_t)LAG) {
          cp_async_wait<LAG>();
          fence_proxy_async();
          mbar_arrive(&full[kbs % NSP]);
          ++kbs;
        }
      }
      if (threadIdx.x == 0) CV_PTL(tno, 1);
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (; kbs < kbg; ++kbs) mbar_arrive(&full[kbs % NSP]);
  } else if (warp == 4) {
    // ================= B producer: TMA =================
    if (lane == 0) {
      uint32_t kbg = 0;
      PTile tl;
      for (int tile = blockIdx.x; p_get_tile(p, tile, BN, &tl); tile += gridDim.x) {
        for (int kb = 0; kb < tl.nkb; ++kb, ++kbg) {
          const int s = kbg % NSP;
          mbar_wait(&empty[s], ((kbg / NSP) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[s], kBStage);
          tma_load_2d(sB + s * kBStage, &tm.t[tl.cls], &full[s], kb * BK, tl.n0);
        }
      }
    }
  } else if (warp == 5) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc(kFmtBF16, BM, BN, 0, 0);
      uint32_t kbg = 0, tcount = 0;
      PTile tl;
      for (int tile = blockIdx.x; p_get_tile(p, tile, BN, &tl); tile += gridDim.x, ++tcount) {
        const uint32_t b = tcount & 1;
        mbar_wait(&acc_empty[b], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < tl.nkb; ++kb, ++kbg) {
          const int s = kbg % NSP;
          mbar_wait(&full[s], (kbg / NSP) & 1);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + s * kAStage), b_base = smem_u32(sB + s * kBStage);
#pragma unroll
          for (int k4 = 0; k4 < BK / 16; ++k4) {
            const uint64_t ad = smem_desc(a_base + k4 * 2 * (BM * 16), BM * 16, 128, kLayoutNone);
            const uint64_t bd = smem_desc(b_base + k4 * 32, 16, 1024, kLayoutSw128);
            umma_f16(tmem_base + b * kAccCols, ad, bd, idesc, (kb | k4) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[b]);
        CV_PTL((int)tcount, 2);
      }
    }
  } else {
    // ================= epilogue warps =================
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;
    const int Nn = p.plan.Nn;
    float run1[NCH], run2[NCH];             // running column sums of (quarter q, column ch*CH + lane)
#pragma unroll
    for (int i = 0; i < NCH; ++i) run1[i] = run2[i] = 0.f;
    int acc_n0 = -1;
    auto flush = [&]() {
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        if (lane < CH) { sRed[q * BN + i * CH + lane] = run1[i]; sRed[4 * BN + q * BN + i * CH + lane] = run2[i]; }
        run1[i] = run2[i] = 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int col = (int)threadIdx.x - 192; col < BN; col += 128) {
        const float a = (sRed[col] + sRed[BN + col]) + (sRed[2 * BN + col] + sRed[3 * BN + col]);
        const float b2 = (sRed[4 * BN + col] + sRed[5 * BN + col]) + (sRed[6 * BN + col] + sRed[7 * BN + col]);
        atomicAdd(p.stats + acc_n0 + col, (double)a);
        atomicAdd(p.stats + Nn + acc_n0 + col, (double)b2);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    };
    uint32_t tcount = 0;
    PTile tl;
    for (int tile = blockIdx.x; p_get_tile(p, tile, BN, &tl); tile += gridDim.x, ++tcount) {
      const Cls& c = p.plan.cls[tl.cls];
      const long long Mc = p.batch * c.Hd * c.Wd;
      const long long m = tl.m0 + r;
      const bool mvalid = m < Mc;
      const long long mm = mvalid ? m : 0;
      const int wd = (int)(mm % c.Wd), hd = (int)((mm / c.Wd) % c.Hd);
      const long long img = mm / ((long long)c.Wd * c.Hd);
      const long long dst_off = img * p.d_n + (long long)(hd * p.plan.os + c.oa) * p.d_h + (long long)(wd * p.plan.os + c.ob) * p.d_w;
      const long long msk_off = img * p.m_n + (long long)(hd * p.plan.os + c.oa) * p.m_h + (long long)(wd * p.plan.os + c.ob) * p.m_w;
      if (p.stats != nullptr && tl.n0 != acc_n0) {
        if (acc_n0 >= 0) flush();
        acc_n0 = tl.n0;
      }
      const uint32_t b = tcount & 1;
      mbar_wait(&acc_full[b], (tcount >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const int ch0 = ci * CH;
        uint32_t raw[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * kAccCols + ch0);
        if (CH == 32) {
          tmem_ld32(taddr, raw);
        } else {
          uint32_t r16[16];
          tmem_ld16(taddr, r16);
#pragma unroll
          for (int i = 0; i < 16; ++i) { raw[i] = r16[i]; raw[i + 16] = 0u; }
        }
        const int nb = tl.n0 + ch0;
        float v[CH], u[CH];
        if (!MASKED) {
          float bsv[CH];
#pragma unroll
          for (int i = 0; i < CH; ++i) bsv[i] = p.bias != nullptr ? __ldg(p.bias + nb + i) : 0.f;
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < CH; ++i) v[i] = mvalid ? __uint_as_float(raw[i]) + bsv[i] : 0.f;
        } else {
          float y[CH];
          if (p.msk_bf16) {
            const uint4* mp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.msk) + (mvalid ? msk_off + nb : 0));
#pragma unroll
            for (int i = 0; i < CH / 8; ++i) {
              const uint4 qv = __ldg(mp + i);
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&qv);
#pragma unroll
              for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h2[k]); y[8 * i + 2 * k] = f.x; y[8 * i + 2 * k + 1] = f.y; }
            }
          } else {
            const float4* mp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.msk) + (mvalid ? msk_off + nb : 0));
#pragma unroll
            for (int i = 0; i < CH / 4; ++i) { const float4 qv = __ldg(mp + i); y[4 * i] = qv.x; y[4 * i + 1] = qv.y; y[4 * i + 2] = qv.z; y[4 * i + 3] = qv.w; }
          }
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const float act = p.msk_scale ? fmaf(y[i], __ldg(p.msk_scale + nb + i), __ldg(p.msk_shift + nb + i)) : y[i];
            const float a = (mvalid && act > 0.f) ? __uint_as_float(raw[i]) : 0.f;
            v[i] = a;
            u[i] = a * y[i];
          }
        }
        if (ci == NCH - 1) {  // the accumulator has been read completely: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          mbar_arrive(&acc_empty[b]);
        }
        if (mvalid) {
          if (p.dst_bf16) {
            uint4* d = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.dst) + dst_off + nb);
#pragma unroll
            for (int i = 0; i < CH / 8; ++i)
              d[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                                pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
          } else {
            float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.dst) + dst_off + nb);
#pragma unroll
            for (int i = 0; i < CH / 4; ++i) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
        if (p.stats != nullptr) {
          float4* row = reinterpret_cast<float4*>(sEp + r * kEpLd);
#pragma unroll
          for (int i = 0; i < CH / 4; ++i) row[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          if (MASKED) {
            float4* row2 = reinterpret_cast<float4*>(sEp2 + r * kEpLd);
#pragma unroll
            for (int i = 0; i < CH / 4; ++i) row2[i] = make_float4(u[4 * i], u[4 * i + 1], u[4 * i + 2], u[4 * i + 3]);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (lane < CH) {
            const float* col = sEp + (q * 32) * kEpLd + lane;
            float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
            if (!MASKED) {
#pragma unroll
              for (int rr = 0; rr < 32; rr += 2) {
                const float a0 = col[rr * kEpLd], a1 = col[(rr + 1) * kEpLd];
                s1a += a0; s1b += a1; s2a = fmaf(a0, a0, s2a); s2b = fmaf(a1, a1, s2b);
              }
            } else {
              const float* col2 = sEp2 + (q * 32) * kEpLd + lane;
#pragma unroll
              for (int rr = 0; rr < 32; rr += 2) {
                s1a += col[rr * kEpLd]; s1b += col[(rr + 1) * kEpLd];
                s2a += col2[rr * kEpLd]; s2b += col2[(rr + 1) * kEpLd];
              }
            }
            run1[ci] += s1a + s1b;
            run2[ci] += s2a + s2b;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      }
      if (threadIdx.x == 192) CV_PTL((int)tcount, 3);
    }
    if (p.stats != nullptr && acc_n0 >= 0) flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, kTmemCols);
  if (tl_buf && threadIdx.x == 0) tl_buf[1] = gtimer();
#undef CV_PTL
}

// ---------------------------------------------------------------------------------------------------------
// TMA-operand variant of the persistent kernel: the im2col A operand is no longer gathered by 128 threads with one cp.async per
// 16 bytes (1024 per tap per tile), but by ONE tiled 4-D TMA instruction per (tap, <= 64-channel block): an m-tile is a box of
// tn images x th rows x Wd columns of the (class-local) output grid, whose source pixels for tap (dh, dw) form the box
// (c0, w = dw, h = h0 * sh + dh, n = img0) walked with traversal stride sh (elementStrides) — zero padding is the TMA's
// out-of-bounds fill, and the tile lands densely as [row][channels] in the 64 / 128-byte swizzled K-major layout tcgen05 reads
// (tools/tma_stride_probe.cu checks these semantics on the hardware).  Every tap re-reads its pixels from L2, but no thread
// touches an operand byte: warp 0 = TMA (A and B), warp 1 = MMA issuer, warps 2-5 = epilogue, same mbarrier rings and
// double-buffered TMEM accumulator as conv_tc_persist_kernel.
// ---------------------------------------------------------------------------------------------------------
constexpr int kTThreads = 192;
// Residency / ring depth of the TMA kernel.  tools/persist_timeline.py shows the epilogue (TMEM load -> bias / mask -> stores ->
// shared-memory statistics) setting the pace of the small-channel layers at ~1.1 tiles per microsecond per SM: MMA commits run
// ahead of "epilogue done" by a growing margin.  Three CTAs per SM with two stages each (twelve epilogue warps instead of eight)
// did NOT raise that rate (14.8 vs 15.4 us on convT 64->32, slower on convT 128->64), so the limit is a per-SM resource of
// the epilogue code, not its dependent latency; two CTAs x three stages stay.
template <int BN, bool MASKED>
constexpr int tma_ctas_per_sm() { return BN > 64 ? 1 : 2; }
template <int BN, bool MASKED>
constexpr int tma_stages() { return BN > 64 ? 4 : 3; }
template <int BN, bool MASKED>
constexpr size_t tma_smem() {
  return (size_t)tma_stages<BN, MASKED>() * (kAStage + BN * BK * 2) + (MASKED ? 2 : 1) * BM * kEpLd * 4 + 2 * 4 * BN * 4 + 256 + 1024;
}

struct TTile {
  int cls, n0, img0, h0;
};
__device__ __forceinline__ bool t_get_tile(const GemmParams& p, int id, int BN, TTile* t) {
  if (id >= p.n_tiles) return false;
  int cls = 0;
#pragma unroll
  for (int i = 1; i < cvplan::kMaxClasses; ++i)
    if (i < p.plan.n_classes && id >= p.tile_start[i]) cls = i;
  const int local = id - p.tile_start[cls];
  const int mt_all = p.t_mtiles[cls];
  const int nt = local / mt_all, mt = local - nt * mt_all;
  const int th_tiles = p.t_tiles_h[cls];
  t->cls = cls;
  t->n0 = nt * BN;
  t->img0 = (mt / th_tiles) * p.t_tn[cls];
  t->h0 = (mt % th_tiles) * p.t_th[cls];
  return true;
}

template <int BN, bool MASKED>
__global__ void __launch_bounds__(kTThreads, (tma_ctas_per_sm<BN, MASKED>())) conv_tma_kernel(const __grid_constant__ TmapPack tmA,
                                                                                              const __grid_constant__ TmapPack tmB,
                                                                                              const GemmParams p) {
  constexpr int NSP = tma_stages<BN, MASKED>();
  constexpr int kBStage = BN * BK * 2;
  constexpr uint32_t kAccCols = BN < 32 ? 32 : BN;
  constexpr uint32_t kTmemCols = 2 * kAccCols;
  constexpr int CH = BN < 32 ? BN : 32;
  constexpr int NCH = BN / CH;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;
  unsigned char* sB = sA + NSP * kAStage;
  float* sEp = reinterpret_cast<float*>(sB + NSP * kBStage);
  float* sEp2 = sEp + BM * kEpLd;
  float* sRed = sEp + (MASKED ? 2 : 1) * BM * kEpLd;                 // [2][4][BN]
  uint64_t* full = reinterpret_cast<uint64_t*>(sRed + 2 * 4 * BN);
  uint64_t* empty = full + NSP;
  uint64_t* acc_full = empty + NSP;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Cs = p.plan.Cs, cb = p.t_cb, kchunks = Cs / cb;
  long long* const tl_buf = g_conv_timeline ? g_conv_timeline + (long long)blockIdx.x * 64 : nullptr;   // debug timeline (tools/persist_timeline.py)
#define CV_TTL(tile_no, k) do { if (tl_buf && (tile_no) < 13) tl_buf[4 + (tile_no) * 4 + (k)] = gtimer(); } while (0)
  // epilogue phase totals of thread 64 (SM clocks): [56] accumulator wait, [57] TMEM load + bias / mask, [58] global stores,
  // [59] statistics (shared-memory transposition, two barriers), [60] whole loop, [61] tiles
  long long ph_wait = 0, ph_ld = 0, ph_st = 0, ph_stat = 0, ph_t = 0, ph_all = 0;
#define CV_PH(acc) do { if (tl_buf) { const long long now_ = clock64(); acc += now_ - ph_t; ph_t = now_; } } while (0)
  if (tl_buf && threadIdx.x == 0) tl_buf[0] = gtimer();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NSP; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
#pragma unroll
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0)
    for (int i = 0; i < p.plan.n_classes; ++i) { prefetch_tmap(&tmA.t[i]); prefetch_tmap(&tmB.t[i]); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: A box + weight slice per (tap, channel block) =================
    if (lane == 0) {
      uint32_t kbg = 0;
      TTile tl;
      const uint32_t b_bytes = (uint32_t)BN * cb * 2;
      int tno = 0;
      for (int tile = blockIdx.x; t_get_tile(p, tile, BN, &tl); tile += gridDim.x, ++tno) {
        const Cls& c = p.plan.cls[tl.cls];
        CV_TTL(tno, 0);
        const uint32_t a_bytes = (uint32_t)(p.t_tn[tl.cls] * p.t_th[tl.cls] * p.t_tw[tl.cls]) * cb * 2;
        const int hbase = tl.h0 * p.plan.sh;
        for (int t = 0; t < c.ntaps; ++t) {
          for (int cc = 0; cc < kchunks; ++cc, ++kbg) {
            const int s = kbg % NSP;
            mbar_wait(&empty[s], ((kbg / NSP) & 1) ^ 1);
            mbar_arrive_expect_tx(&full[s], a_bytes + b_bytes);
            tma_load_4d(sA + s * kAStage, &tmA.t[tl.cls], &full[s], cc * cb, c.dw[t], hbase + c.dh[t], tl.img0);
            tma_load_2d(sB + s * kBStage, &tmB.t[tl.cls], &full[s], t * Cs + cc * cb, tl.n0);
          }
        }
        CV_TTL(tno, 1);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc(kFmtBF16, BM, BN, 0, 0);
      const uint32_t layout = cb == 64 ? kLayoutSw128 : kLayoutSw64;
      const uint32_t sbo = cb == 64 ? 1024u : 512u;      // 8 rows of 128 / 64 bytes
      const int k4n = cb / 16;
      uint32_t kbg = 0, tcount = 0;
      TTile tl;
      for (int tile = blockIdx.x; t_get_tile(p, tile, BN, &tl); tile += gridDim.x, ++tcount) {
        const uint32_t b = tcount & 1;
        mbar_wait(&acc_empty[b], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        const int nst = p.plan.cls[tl.cls].ntaps * kchunks;
        for (int st = 0; st < nst; ++st, ++kbg) {
          const int s = kbg % NSP;
          mbar_wait(&full[s], (kbg / NSP) & 1);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + s * kAStage), b_base = smem_u32(sB + s * kBStage);
          for (int k4 = 0; k4 < k4n; ++k4) {
            // both operands K-major, rows of cb * 2 bytes, hardware swizzle; K advance = +32 bytes inside the swizzle atom
            const uint64_t ad = smem_desc(a_base + k4 * 32, 16, sbo, layout);
            const uint64_t bd = smem_desc(b_base + k4 * 32, 16, sbo, layout);
            umma_f16(tmem_base + b * kAccCols, ad, bd, idesc, (st | k4) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[b]);
        CV_TTL((int)tcount, 2);
      }
    }
  } else {
    // ================= epilogue warps =================
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;
    const int et = (int)threadIdx.x - 64;   // 0..127 among the epilogue threads
    const int Nn = p.plan.Nn;
    float run1[NCH], run2[NCH];             // running column sums of (quarter q, column ch*CH + lane)
#pragma unroll
    for (int i = 0; i < NCH; ++i) run1[i] = run2[i] = 0.f;
    int acc_n0 = -1;
    auto flush = [&]() {
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        if (lane < CH) { sRed[q * BN + i * CH + lane] = run1[i]; sRed[4 * BN + q * BN + i * CH + lane] = run2[i]; }
        run1[i] = run2[i] = 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int col = et; col < BN; col += 128) {
        const float a = (sRed[col] + sRed[BN + col]) + (sRed[2 * BN + col] + sRed[3 * BN + col]);
        const float b2 = (sRed[4 * BN + col] + sRed[5 * BN + col]) + (sRed[6 * BN + col] + sRed[7 * BN + col]);
        atomicAdd(p.stats + acc_n0 + col, (double)a);
        atomicAdd(p.stats + Nn + acc_n0 + col, (double)b2);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    };
    uint32_t tcount = 0;
    TTile tl;
    int map_cls = -1, n_l = 0, h_l = 0, wd = 0;   // this thread's (image, row, column) inside a tile: depends on the class only
    if (tl_buf) ph_all = ph_t = clock64();
    for (int tile = blockIdx.x; t_get_tile(p, tile, BN, &tl); tile += gridDim.x, ++tcount) {
      const Cls& c = p.plan.cls[tl.cls];
      if (tl.cls != map_cls) {
        const int tw = p.t_tw[tl.cls], thw = p.t_th[tl.cls] * tw;
        n_l = r / thw;
        const int rem = r - n_l * thw;
        h_l = rem / tw;
        wd = rem - h_l * tw;
        map_cls = tl.cls;
      }
      const int hd = tl.h0 + h_l;
      const long long img = tl.img0 + n_l;
      const bool mvalid = n_l < p.t_tn[tl.cls] && img < p.batch && hd < c.Hd;
      const long long dst_off = img * p.d_n + (long long)(hd * p.plan.os + c.oa) * p.d_h + (long long)(wd * p.plan.os + c.ob) * p.d_w;
      const long long msk_off = img * p.m_n + (long long)(hd * p.plan.os + c.oa) * p.m_h + (long long)(wd * p.plan.os + c.ob) * p.m_w;
      if (p.stats != nullptr && tl.n0 != acc_n0) {
        if (acc_n0 >= 0) flush();
        acc_n0 = tl.n0;
      }
      const uint32_t b = tcount & 1;
      mbar_wait(&acc_full[b], (tcount >> 1) & 1);
      tc_fence_after();
      CV_PH(ph_wait);
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const int ch0 = ci * CH;
        uint32_t raw[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * kAccCols + ch0);
        if (CH == 32) {
          tmem_ld32(taddr, raw);
        } else {
          uint32_t r16[16];
          tmem_ld16(taddr, r16);
#pragma unroll
          for (int i = 0; i < 16; ++i) { raw[i] = r16[i]; raw[i + 16] = 0u; }
        }
        const int nb = tl.n0 + ch0;
        float v[CH], u[CH];
        if (!MASKED) {
          float bsv[CH];
#pragma unroll
          for (int i = 0; i < CH; ++i) bsv[i] = p.bias != nullptr ? __ldg(p.bias + nb + i) : 0.f;
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < CH; ++i) v[i] = mvalid ? __uint_as_float(raw[i]) + bsv[i] : 0.f;
        } else {
          float y[CH];
          if (p.msk_bf16) {
            const uint4* mp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.msk) + (mvalid ? msk_off + nb : 0));
#pragma unroll
            for (int i = 0; i < CH / 8; ++i) {
              const uint4 qv = __ldg(mp + i);
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&qv);
#pragma unroll
              for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h2[k]); y[8 * i + 2 * k] = f.x; y[8 * i + 2 * k + 1] = f.y; }
            }
          } else {
            const float4* mp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.msk) + (mvalid ? msk_off + nb : 0));
#pragma unroll
            for (int i = 0; i < CH / 4; ++i) { const float4 qv = __ldg(mp + i); y[4 * i] = qv.x; y[4 * i + 1] = qv.y; y[4 * i + 2] = qv.z; y[4 * i + 3] = qv.w; }
          }
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const float act = p.msk_scale ? fmaf(y[i], __ldg(p.msk_scale + nb + i), __ldg(p.msk_shift + nb + i)) : y[i];
            const float a = (mvalid && act > 0.f) ? __uint_as_float(raw[i]) : 0.f;
            v[i] = a;
            u[i] = a * y[i];
          }
        }
        if (ci == NCH - 1) {  // the accumulator has been read completely: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          mbar_arrive(&acc_empty[b]);
        }
        CV_PH(ph_ld);
        if (mvalid) {
          if (p.dst_bf16) {
            uint4* d = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.dst) + dst_off + nb);
#pragma unroll
            for (int i = 0; i < CH / 8; ++i)
              d[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                                pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
          } else {
            float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.dst) + dst_off + nb);
#pragma unroll
            for (int i = 0; i < CH / 4; ++i) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
        CV_PH(ph_st);
        if (p.stats != nullptr) {
          float4* row = reinterpret_cast<float4*>(sEp + r * kEpLd);
#pragma unroll
          for (int i = 0; i < CH / 4; ++i) row[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          if (MASKED) {
            float4* row2 = reinterpret_cast<float4*>(sEp2 + r * kEpLd);
#pragma unroll
            for (int i = 0; i < CH / 4; ++i) row2[i] = make_float4(u[4 * i], u[4 * i + 1], u[4 * i + 2], u[4 * i + 3]);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (lane < CH) {
            const float* col = sEp + (q * 32) * kEpLd + lane;
            float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
            if (!MASKED) {
#pragma unroll
              for (int rr = 0; rr < 32; rr += 2) {
                const float a0 = col[rr * kEpLd], a1 = col[(rr + 1) * kEpLd];
                s1a += a0; s1b += a1; s2a = fmaf(a0, a0, s2a); s2b = fmaf(a1, a1, s2b);
              }
            } else {
              const float* col2 = sEp2 + (q * 32) * kEpLd + lane;
#pragma unroll
              for (int rr = 0; rr < 32; rr += 2) {
                s1a += col[rr * kEpLd]; s1b += col[(rr + 1) * kEpLd];
                s2a += col2[rr * kEpLd]; s2b += col2[(rr + 1) * kEpLd];
              }
            }
            run1[ci] += s1a + s1b;
            run2[ci] += s2a + s2b;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        CV_PH(ph_stat);
      }
      if (et == 0) CV_TTL((int)tcount, 3);
    }
    if (p.stats != nullptr && acc_n0 >= 0) flush();
    if (tl_buf && et == 0) {
      tl_buf[56] = ph_wait; tl_buf[57] = ph_ld; tl_buf[58] = ph_st; tl_buf[59] = ph_stat; tl_buf[60] = clock64() - ph_all; tl_buf[61] = tcount;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
  if (tl_buf && threadIdx.x == 0) tl_buf[1] = gtimer();
#undef CV_TTL
#undef CV_PH
}

// ---------------------------------------------------------------------------
// weight gradient:  dW[kidx, n] += sum_{pixels m} act[m @ tap(kidx), c(kidx)] * dy[m, n]
//   GEMM with M = (tap, channel) rows, N = output channels, K = pixels (split over CTAs).
//   Both operands are pixel-major in HBM (NHWC: channels contiguous), i.e. MN-major for
//   this GEMM, so the producers copy 16-byte channel runs into the UMMA MN-major
//   interleave layout [mn-group][pixel][8 x bf16] unchanged — no transposition anywhere.
//   Partial sums of the split-K CTAs are reduced with fp32 red.global.add into the
//   reference weight layout.
// ---------------------------------------------------------------------------
struct WgradParams {
  Plan plan;   // the FPROP plan of the layer
  long long batch;
  const void* src; long long s_n, s_h, s_w, s_c; int src_bf16;
  const float *pre_scale, *pre_shift; int pre_relu;
  const void* dy; long long y_n, y_h, y_w, y_c; int dy_bf16;
  float* dw;
  int splits;
  int x3;   // split mode: every pixel block runs six times over the bf16 parts of both operands (see kSplitPasses)
};

constexpr int WK = 64;                       // pixels per k-block
constexpr int kWStageA = 128 * WK * 2;       // 16 KiB

template <int BN>
__global__ void __launch_bounds__(kThreads) wgrad_tc_kernel(const WgradParams p) {
  constexpr int kBStage = BN * WK * 2;
  constexpr uint32_t kAccStride = BN < 32 ? 32 : BN;
  const uint32_t kTmemCols = p.x3 ? 4 * kAccStride : kAccStride;   // split mode: three accumulators by magnitude class
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;
  unsigned char* sB = smem + NS * kWStageA;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + NS * kBStage);
  uint64_t* empty = full + NS;
  uint64_t* tmem_full = empty + NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* sScale = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(full) + 96);
  float* sShift = sScale + kMaxPreC;

  const int cls_id = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const Cls& c = p.plan.cls[cls_id];
  const int Cs = p.plan.Cs, Kreal = c.ntaps * Cs;
  const int k0 = blockIdx.x * 128;  // first (tap, channel) row of this tile
  if (k0 >= Kreal) return;
  const int n0 = blockIdx.y * BN;
  const long long Mc = p.batch * c.Hd * c.Wd;
  const long long nkb_total = (Mc + WK - 1) / WK;
  const long long kb_begin = nkb_total * split / p.splits, kb_end = nkb_total * (split + 1) / p.splits;
  const int nkb = (int)(kb_end - kb_begin);
  if (nkb <= 0) return;
  const int nsub = p.x3 ? kSplitPasses : 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // debug timeline (tools/wgrad_timeline.py): per CTA [0] start, [1] end (globaltimer ns); SM-clock totals of producer thread 0
  // [8] index math + cp.async issue, [9] waiting for a free stage, [10] waiting for its own copies; of the MMA thread [11] waiting
  // for operands, [12] issuing; [13] k-blocks; epilogue thread 0 [14] accumulator wait, [15] scatter-add
  long long* const tl_buf = g_conv_timeline
                                ? g_conv_timeline + ((long long)blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z)) * 64
                                : nullptr;
  long long ph_a = 0, ph_b = 0, ph_c = 0, ph_t = 0;
#define WG_PH(acc) do { if (tl_buf) { const long long now_ = clock64(); acc += now_ - ph_t; ph_t = now_; } } while (0)
  if (tl_buf && threadIdx.x == 0) tl_buf[0] = gtimer();

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 128); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  if (p.pre_scale != nullptr)
    for (int i = threadIdx.x; i < Cs; i += kThreads) { sScale[i] = __ldg(p.pre_scale + i); sShift[i] = __ldg(p.pre_shift + i); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // producers: thread = (pixel px of the k-block, half hf); 8 A chunks + BN/16 B chunks each
    const int px = threadIdx.x & 63, hf = threadIdx.x >> 6;
    const bool vec = (Cs % 8 == 0) && p.s_c == 1;
    const int Nn = p.plan.Nn;
    // both operands are plain bf16 channel runs: cp.async them into the MN-major interleave layout (zero-fill for
    // padding taps / rows past the problem), announcing each k-block two iterations after it was issued
    const bool cpa = vec && p.src_bf16 && p.pre_scale == nullptr && !p.pre_relu && p.y_c == 1 && p.dy_bf16 && (Nn % 8 == 0) && !p.x3;
    if (cpa) {
      const __nv_bfloat16* srcb = reinterpret_cast<const __nv_bfloat16*>(p.src);
      const __nv_bfloat16* dyb = reinterpret_cast<const __nv_bfloat16*>(p.dy);
      constexpr int NG = BN / 8, NGH = (NG + 1) / 2;
      if (tl_buf) ph_t = clock64();
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % NS;
        const long long m = (kb_begin + kb) * WK + px;
        const bool mvalid = m < Mc;
        const long long mm = mvalid ? m : 0;
        const int wd = (int)(mm % c.Wd);
        const int hd = (int)((mm / c.Wd) % c.Hd);
        const long long img = mm / ((long long)c.Wd * c.Hd);
        const int hbase = hd * p.plan.sh, wbase = wd * p.plan.sh;
        const long long img_off = img * p.s_n;
        const long long dy_off = img * p.y_n + (long long)(hd * p.plan.os + c.oa) * p.y_h + (long long)(wd * p.plan.os + c.ob) * p.y_w;
        WG_PH(ph_a);
        mbar_wait(&empty[s], ((kb / NS) & 1) ^ 1);
        WG_PH(ph_b);
        const uint32_t a_dst = smem_u32(sA + s * kWStageA) + px * 16, b_dst = smem_u32(sB + s * kBStage) + px * 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int grp = hf * 8 + j;
          const int kidx = k0 + grp * 8;
          bool v = mvalid && kidx < Kreal;
          long long off = 0;
          if (v) {
            const int t = kidx / Cs, ch = kidx - t * Cs;
            const int hs = hbase + c.dh[t], ws = wbase + c.dw[t];
            v = hs >= 0 && hs < p.plan.Hs && ws >= 0 && ws < p.plan.Ws;
            off = v ? img_off + hs * p.s_h + ws * p.s_w + ch : 0;
          }
          cp_async16(a_dst + grp * (WK * 16), srcb + off, v ? 16u : 0u);
        }
#pragma unroll
        for (int j = 0; j < NGH; ++j) {
          const int grp = hf * NGH + j;
          if (grp < NG) {
            const int n = n0 + grp * 8;
            const bool v = mvalid && n + 8 <= Nn;
            cp_async16(b_dst + grp * (WK * 16), dyb + (v ? dy_off + n : 0), v ? 16u : 0u);
          }
        }
        cp_async_commit();
        WG_PH(ph_a);
        if (kb >= 2) {
          cp_async_wait<2>();
          fence_proxy_async();
          mbar_arrive(&full[(kb - 2) % NS]);
        }
        WG_PH(ph_c);
      }
      if (tl_buf && threadIdx.x == 0) { tl_buf[8] = ph_a; tl_buf[9] = ph_b; tl_buf[10] = ph_c; tl_buf[13] = nkb; }
      if (nkb >= 2) {
        cp_async_wait<1>();
        fence_proxy_async();
        mbar_arrive(&full[(nkb - 2) % NS]);
      }
      cp_async_wait<0>();
      fence_proxy_async();
      mbar_arrive(&full[(nkb - 1) % NS]);
    }
    for (int it = 0; it < (cpa ? 0 : nkb * nsub); ++it) {
      const int kb = it / nsub, sub = it - kb * nsub;
      const int a_lo = p.x3 ? split_a_part(sub) : 0, b_lo = p.x3 ? split_b_part(sub) : 0;
      const int s = it % NS;
      const long long m = (kb_begin + kb) * WK + px;
      const bool mvalid = m < Mc;
      const long long mm = mvalid ? m : 0;
      const int wd = (int)(mm % c.Wd);
      const int hd = (int)((mm / c.Wd) % c.Hd);
      const long long img = mm / ((long long)c.Wd * c.Hd);
      const int hbase = hd * p.plan.sh, wbase = wd * p.plan.sh;
      const long long img_off = img * p.s_n;
      const long long dy_off = img * p.y_n + (long long)(hd * p.plan.os + c.oa) * p.y_h + (long long)(wd * p.plan.os + c.ob) * p.y_w;
      unsigned char* a_st = sA + s * kWStageA;
      unsigned char* b_st = sB + s * kBStage;
      constexpr int NG = BN / 8, NGH = (NG + 1) / 2;
      const bool dyvec = p.y_c == 1;
      // ---- phase 1: every load of this k-block in flight before the stage is even free
      uint4 qa0[8], qa1[8], qb0[NGH], qb1[NGH];
      bool oka[8], okb[NGH];
      int cha[8];
      if (vec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int kidx = k0 + (hf * 8 + j) * 8;
          bool v = mvalid && kidx < Kreal;
          long long off = 0;
          int ch = 0;
          if (v) {
            const int t = kidx / Cs;
            ch = kidx - t * Cs;
            const int hs = hbase + c.dh[t], ws = wbase + c.dw[t];
            v = hs >= 0 && hs < p.plan.Hs && ws >= 0 && ws < p.plan.Ws;
            off = v ? img_off + hs * p.s_h + ws * p.s_w + ch : 0;
          }
          oka[j] = v;
          cha[j] = ch;
          if (p.src_bf16) {
            qa0[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.src) + off));
          } else {
            qa0[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.src) + off));
            qa1[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.src) + off + 4));
          }
        }
      }
      if (dyvec) {
#pragma unroll
        for (int j = 0; j < NGH; ++j) {
          const int grp = hf * NGH + j;
          const int n = n0 + grp * 8;
          const bool v = mvalid && grp < NG && n + 8 <= Nn;
          okb[j] = v;
          const long long off = v ? dy_off + n : 0;
          if (p.dy_bf16) {
            qb0[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.dy) + off));
          } else {
            qb0[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.dy) + off));
            qb1[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.dy) + off + 4));
          }
        }
      }
      mbar_wait(&empty[s], ((it / NS) & 1) ^ 1);
      // ---- phase 2a: A = gathered activation (BatchNorm-apply + ReLU), mn-groups hf*8 .. hf*8+7
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int grp = hf * 8 + j;
        const int kidx = k0 + grp * 8;
        uint4 out = make_uint4(0u, 0u, 0u, 0u);
        float v[8];
        if (vec) {
          if (p.src_bf16) {
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&qa0[j]);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h2[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
          } else {
            v[0] = __uint_as_float(qa0[j].x); v[1] = __uint_as_float(qa0[j].y); v[2] = __uint_as_float(qa0[j].z); v[3] = __uint_as_float(qa0[j].w);
            v[4] = __uint_as_float(qa1[j].x); v[5] = __uint_as_float(qa1[j].y); v[6] = __uint_as_float(qa1[j].z); v[7] = __uint_as_float(qa1[j].w);
          }
          if (p.pre_scale != nullptr) {
            const float4 s0 = *reinterpret_cast<const float4*>(sScale + cha[j]), s1 = *reinterpret_cast<const float4*>(sScale + cha[j] + 4);
            const float4 h0 = *reinterpret_cast<const float4*>(sShift + cha[j]), h1 = *reinterpret_cast<const float4*>(sShift + cha[j] + 4);
            v[0] = fmaf(v[0], s0.x, h0.x); v[1] = fmaf(v[1], s0.y, h0.y); v[2] = fmaf(v[2], s0.z, h0.z); v[3] = fmaf(v[3], s0.w, h0.w);
            v[4] = fmaf(v[4], s1.x, h1.x); v[5] = fmaf(v[5], s1.y, h1.y); v[6] = fmaf(v[6], s1.z, h1.z); v[7] = fmaf(v[7], s1.w, h1.w);
          }
          if (p.pre_relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (a_lo) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = split_part(v[i], a_lo);
          }
          if (oka[j]) out = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        } else if (mvalid && kidx < Kreal) {
          bool okk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int kk = kidx + i;
            bool vld = kk < Kreal;
            long long off = 0;
            int ch = 0;
            if (vld) {
              const int t = kk / Cs;
              ch = kk - t * Cs;
              const int hs = hbase + c.dh[t], ws = wbase + c.dw[t];
              vld = hs >= 0 && hs < p.plan.Hs && ws >= 0 && ws < p.plan.Ws;
              off = vld ? img_off + hs * p.s_h + ws * p.s_w + ch * p.s_c : 0;
            }
            okk[i] = vld;
            cha[0] = ch;
            float x = ld_elem(p.src, off, p.src_bf16);
            if (p.pre_scale != nullptr) x = fmaf(x, sScale[ch], sShift[ch]);
            if (p.pre_relu) x = fmaxf(x, 0.f);
            if (a_lo) x = split_part(x, a_lo);
            v[i] = vld ? x : 0.f;
          }
          (void)okk;
          out = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        }
        *reinterpret_cast<uint4*>(a_st + grp * (WK * 16) + px * 16) = out;
      }
      // ---- phase 2b: B = dy rows, n-groups split between the two halves
#pragma unroll
      for (int j = 0; j < NGH; ++j) {
        const int grp = hf * NGH + j;
        if (grp < NG) {
          const int n = n0 + grp * 8;
          uint4 out = make_uint4(0u, 0u, 0u, 0u);
          if (dyvec && okb[j]) {
            if (p.dy_bf16) {
              out = b_lo ? make_uint4(0u, 0u, 0u, 0u) : qb0[j];   // a bf16 gradient has no low half
            } else {
              float v[8] = {__uint_as_float(qb0[j].x), __uint_as_float(qb0[j].y), __uint_as_float(qb0[j].z), __uint_as_float(qb0[j].w),
                            __uint_as_float(qb1[j].x), __uint_as_float(qb1[j].y), __uint_as_float(qb1[j].z), __uint_as_float(qb1[j].w)};
              if (b_lo) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = split_part(v[i], b_lo);
              }
              out = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            }
          } else if (mvalid && n < Nn && !(dyvec && n + 8 <= Nn)) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[i] = (n + i < Nn) ? ld_elem(p.dy, dy_off + (n + i) * p.y_c, p.dy_bf16) : 0.f;
              if (b_lo) v[i] = split_part(v[i], b_lo);
            }
            out = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          }
          *reinterpret_cast<uint4*>(b_st + grp * (WK * 16) + px * 16) = out;
        }
      }
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }

    // ---- epilogue: scatter-add the tile into the reference weight layout
    if (tl_buf) ph_t = clock64();
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    long long ph_e0 = 0, ph_e1 = 0;
    WG_PH(ph_e0);
    const int kidx = k0 + threadIdx.x;
    const bool rvalid = kidx < Kreal;
    const int t = rvalid ? kidx / Cs : 0, ch = rvalid ? kidx - t * Cs : 0;
    const long long row_off = ch * p.plan.ws_c + c.wtap[t];
    constexpr int CH = BN < 32 ? BN : 32;
#pragma unroll 1
    for (int ch0 = 0; ch0 < BN; ch0 += CH) {
      uint32_t raw[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)ch0;
      if (CH == 32) {
        tmem_ld32(taddr, raw);
        if (p.x3) { tmem_ld_wait(); split_merge32(raw, taddr, kAccStride); }
      } else {
        uint32_t r16[16];
        tmem_ld16(taddr, r16);
        if (p.x3) { tmem_ld_wait(); split_merge16(r16, taddr, kAccStride); }
#pragma unroll
        for (int i = 0; i < 16; ++i) { raw[i] = r16[i]; raw[i + 16] = 0u; }
      }
      tmem_ld_wait();
      if (rvalid) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int n = n0 + ch0 + i;
          if (n < Nn) atomicAdd(p.dw + n * p.plan.ws_n + row_off, __uint_as_float(raw[i]));
        }
      }
    }
    WG_PH(ph_e1);
    if (tl_buf && threadIdx.x == 0) { tl_buf[14] = ph_e0; tl_buf[15] = ph_e1; }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc(kFmtBF16, 128, BN, 1, 1);  // both operands MN-major
      if (tl_buf) ph_t = clock64();
      for (int kb = 0; kb < nkb * nsub; ++kb) {
        const int s = kb % NS;
        mbar_wait(&full[s], (kb / NS) & 1);
        tc_fence_after();
        WG_PH(ph_a);
        const uint32_t a_base = smem_u32(sA + s * kWStageA), b_base = smem_u32(sB + s * kBStage);
        const int sub = kb % nsub;
        const uint32_t acc = p.x3 ? (uint32_t)split_acc(sub) * kAccStride : 0u;
        const bool first = kb < nsub && (sub == 0 || sub == 1 || sub == 3);
#pragma unroll
        for (int k4 = 0; k4 < WK / 16; ++k4) {
          // MN-major interleave: 8-pixel k-groups 128 B apart (LBO), 8-channel mn-groups WK*16 B apart (SBO)
          const uint64_t ad = smem_desc(a_base + k4 * 256, 128, WK * 16, kLayoutNone);
          const uint64_t bd = smem_desc(b_base + k4 * 256, 128, WK * 16, kLayoutNone);
          umma_f16(tmem_base + acc, ad, bd, idesc, (first && k4 == 0) ? 0u : 1u);
        }
        umma_commit(&empty[s]);
        WG_PH(ph_b);
      }
      umma_commit(tmem_full);
      if (tl_buf) { tl_buf[11] = ph_a; tl_buf[12] = ph_b; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, kTmemCols);
  if (tl_buf && threadIdx.x == 0) tl_buf[1] = gtimer();
#undef WG_PH
}

// ---------------------------------------------------------------------------
// weight gradient, TMA operands (round 2, third session).  tools/wgrad_timeline.py showed the cp.async kernel above spending
// ~3400 SM cycles per 64-pixel k-block in its producers (ten 16-byte copies per thread at 128-byte lane stride: LSU / L1 tag
// throughput) against ~390 cycles of MMA issue.  Here a k-block is one BOX of pixels (tn images x th rows x tw columns of the
// class-local output grid, tn*th*tw a multiple of 16, <= 128) and both operands arrive as tiled 4-D TMA boxes [pixel][channel]
// with 128-byte (64 channels) or 64-byte (32 channels) swizzled rows -- which is exactly the canonical MN-major SWIZZLE_128B /
// SWIZZLE_64B UMMA layout ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)): 64 (32) MN-contiguous elements per row, 8 pixel rows per atom,
// SBO = 1024 (512) bytes between 8-row groups, LBO = one box between 64- (32-)channel blocks.  Operand A: the activation at
// tap (dh, dw) of the M-tile's (tap, channel-block) pairs -- one box each, zero padding by out-of-bounds fill; operand B: dy on
// the class grid (traversal stride `os`), zero beyond the grid / the batch, so padded box positions contribute nothing.
//   warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue (fp32 red.global.add into the reference weight layout).
// ---------------------------------------------------------------------------
struct WgradTmaParams {
  Plan plan;
  long long batch;
  float* dw;
  int splits;
  int cb, cbn;                       // channels per A / B box (64 or 32)
  int tw[cvplan::kMaxClasses], th[cvplan::kMaxClasses], tn[cvplan::kMaxClasses], tiles_h[cvplan::kMaxClasses], nbox[cvplan::kMaxClasses];
};
constexpr int kWtStages = 3;
template <int BN>
constexpr size_t wgrad_tma_smem() { return (size_t)kWtStages * (128 * 128 * 2 + BN * 128 * 2) + 256 + 1024; }

template <int BN>
__global__ void __launch_bounds__(kTThreads, 1) wgrad_tma_kernel(const __grid_constant__ TmapPack tmA, const __grid_constant__ TmapPack tmB,
                                                                const WgradTmaParams p) {
  constexpr int kAStageW = 128 * 128 * 2, kBStageW = BN * 128 * 2;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;
  unsigned char* sB = sA + kWtStages * kAStageW;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + kWtStages * kBStageW);
  uint64_t* empty = full + kWtStages;
  uint64_t* tmem_full = empty + kWtStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int cls_id = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const Cls& c = p.plan.cls[cls_id];
  const int Cs = p.plan.Cs, Kreal = c.ntaps * Cs, cb = p.cb, cbn = p.cbn;
  const int k0 = blockIdx.x * 128;  // first (tap, channel) row of this tile
  if (k0 >= Kreal) return;
  const int n0 = blockIdx.y * BN;
  const int nbox = p.nbox[cls_id];
  const int bx0 = (int)((long long)nbox * split / p.splits), bx1 = (int)((long long)nbox * (split + 1) / p.splits);
  if (bx1 <= bx0) return;
  const int tw = p.tw[cls_id], th = p.th[cls_id], tn = p.tn[cls_id], tiles_h = p.tiles_h[cls_id];
  const int kpix = tw * th * tn;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kWtStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA.t[cls_id]); prefetch_tmap(&tmB.t[cls_id]); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int na = 128 / cb, nb = BN / cbn;                    // boxes per stage
  const uint32_t a_box = (uint32_t)kpix * cb * 2, b_box = (uint32_t)kpix * cbn * 2;

  if (warp == 0) {
    if (lane == 0) {
      int n_a = 0;
      for (int j = 0; j < na; ++j) n_a += (k0 + j * cb < Kreal) ? 1 : 0;
      const uint32_t bytes = (uint32_t)n_a * a_box + (uint32_t)nb * b_box;
      for (int bx = bx0; bx < bx1; ++bx) {
        const int it = bx - bx0, s = it % kWtStages;
        const int img0 = (bx / tiles_h) * tn, h0 = (bx % tiles_h) * th;
        mbar_wait(&empty[s], ((it / kWtStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], bytes);
        for (int j = 0; j < na; ++j) {
          const int kidx = k0 + j * cb;
          if (kidx >= Kreal) break;
          const int t = kidx / Cs, ch = kidx - t * Cs;
          tma_load_4d(sA + s * kAStageW + j * a_box, &tmA.t[cls_id], &full[s], ch, c.dw[t], h0 * p.plan.sh + c.dh[t], img0);
        }
        for (int j = 0; j < nb; ++j)
          tma_load_4d(sB + s * kBStageW + j * b_box, &tmB.t[cls_id], &full[s], n0 + j * cbn, c.ob, h0 * p.plan.os + c.oa, img0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc(kFmtBF16, 128, BN, 1, 1);  // both operands MN-major
      const uint32_t a_lay = cb == 64 ? kLayoutSw128 : kLayoutSw64, b_lay = cbn == 64 ? kLayoutSw128 : kLayoutSw64;
      const uint32_t a_sbo = cb == 64 ? 1024u : 512u, b_sbo = cbn == 64 ? 1024u : 512u;   // 8 pixel rows of 128 / 64 bytes
      for (int bx = bx0; bx < bx1; ++bx) {
        const int it = bx - bx0, s = it % kWtStages;
        mbar_wait(&full[s], (it / kWtStages) & 1);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + s * kAStageW), b_base = smem_u32(sB + s * kBStageW);
        for (int k16 = 0; k16 < kpix / 16; ++k16) {
          const uint64_t ad = smem_desc(a_base + k16 * 2 * a_sbo, a_box, a_sbo, a_lay);
          const uint64_t bd = smem_desc(b_base + k16 * 2 * b_sbo, b_box, b_sbo, b_lay);
          umma_f16(tmem_base, ad, bd, idesc, (it | k16) != 0 ? 1u : 0u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
  } else {
    // ---- epilogue: scatter-add the tile into the reference weight layout
    const int q = warp & 3;
    const int r = q * 32 + lane;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int kidx = k0 + r;
    const bool rvalid = kidx < Kreal;
    const int t = rvalid ? kidx / Cs : 0, ch = rvalid ? kidx - t * Cs : 0;
    const long long row_off = ch * p.plan.ws_c + c.wtap[t];
    const int Nn = p.plan.Nn;
    constexpr int CH = BN < 32 ? BN : 32;
#pragma unroll 1
    for (int ch0 = 0; ch0 < BN; ch0 += CH) {
      uint32_t raw[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ch0;
      if (CH == 32) {
        tmem_ld32(taddr, raw);
      } else {
        uint32_t r16[16];
        tmem_ld16(taddr, r16);
#pragma unroll
        for (int i = 0; i < 16; ++i) { raw[i] = r16[i]; raw[i + 16] = 0u; }
      }
      tmem_ld_wait();
      if (rvalid) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int n = n0 + ch0 + i;
          if (n < Nn) atomicAdd(p.dw + n * p.plan.ws_n + row_off, __uint_as_float(raw[i]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------
// weight packing: fp32 reference layout -> bf16 [class][n_pad][Kp], zero padded
// ---------------------------------------------------------------------------
__global__ void pack_weight_kernel(const Plan plan, const float* __restrict__ w, __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out_p1, __nv_bfloat16* __restrict__ out_p2) {
  const Cls& c = plan.cls[blockIdx.y];
  const int n_pad = (plan.Nn + 15) / 16 * 16;
  const long long total = (long long)n_pad * c.Kp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / c.Kp), k = (int)(i % c.Kp);
    float v = 0.f;
    if (n < plan.Nn && k < c.ntaps * plan.Cs) {
      const int t = k / plan.Cs, ch = k - t * plan.Cs;
      v = w[n * plan.ws_n + ch * plan.ws_c + c.wtap[t]];
    }
    out[c.w_off + i] = __float2bfloat16(v);
    if (out_p1 != nullptr) {   // split mode: W = W_0 + W_1 + W_2
      out_p1[c.w_off + i] = __float2bfloat16(split_part(v, 1));
      out_p2[c.w_off + i] = __float2bfloat16(split_part(v, 2));
    }
  }
}

// All packed copies of a model in ONE launch (after the optimiser step): a device-resident table of (plan, class, source weight,
// destination, first block); a block finds its entry by binary search over the block offsets.  Replaces one pack_weight_kernel
// launch per (layer, role) per step — ~20 launches of 3-12 us each on the critical path of the next GEMM.
struct MultiPackEntry {
  Plan plan;
  const float* w;
  __nv_bfloat16* out;
  int cls;
  int block0;     // first block of this (entry, class)
  int nblocks;
  int pad;
};
__global__ void pack_weight_multi_kernel(const MultiPackEntry* __restrict__ tab, int n) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const MultiPackEntry& e = tab[lo];
  const Cls& c = e.plan.cls[e.cls];
  const int n_pad = (e.plan.Nn + 15) / 16 * 16;
  const long long total = (long long)n_pad * c.Kp;
  const int Cs = e.plan.Cs, kreal = c.ntaps * Cs;
  for (long long i = (long long)(blockIdx.x - e.block0) * blockDim.x + threadIdx.x; i < total; i += (long long)e.nblocks * blockDim.x) {
    const int nn = (int)(i / c.Kp), k = (int)(i % c.Kp);
    float v = 0.f;
    if (nn < e.plan.Nn && k < kreal) {
      const int t = k / Cs, ch = k - t * Cs;
      v = e.w[nn * e.plan.ws_n + ch * e.plan.ws_c + c.wtap[t]];
    }
    e.out[c.w_off + i] = __float2bfloat16(v);
  }
}

// ---------------------------------------------------------------------------
// Skinny linear layer  out[B, N] = src[B, K] . W[N, K]^T + bias  with N = 16 or 32 and a deep K (the latent heads on the
// 2048-wide flattened encoder output, the data gradient of the decoder's first Linear): 134 MFLOP for which a 128-row tcgen05
// tile needs split-K over 16 CTAs, fp32 atomics on a zeroed destination and a memset (24 us per call on the step's critical path).
// Here: CTA = 16 rows, eight warps each own K/8 of the reduction and run warp-level mma.sync.m16n8k16 (bf16, fp32 accumulate)
// with A / B fragments loaded straight from global memory in the instruction's register layout (every 32-byte sector fully
// used; W is L2 resident), partial tiles meet in shared memory, one pass adds the bias and stores.  No atomics, no memset.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t skinny_a_pair(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
__device__ __forceinline__ uint32_t skinny_a_pair(const float* p) {   // fp32 rows are rounded to bf16 like the general kernel's gather
  const float2 f = __ldg(reinterpret_cast<const float2*>(p));
  return pack_bf16(f.x, f.y);
}
template <int NT, typename TS>   // n-tiles of 8 columns: N = 8 * NT; TS = source element type (bf16 or fp32)
__global__ void __launch_bounds__(256) skinny_linear_kernel(const TS* __restrict__ src, long long s_n, const __nv_bfloat16* __restrict__ w,
                                                           int Kp, const float* __restrict__ bias, float* __restrict__ out, long long d_n,
                                                           int B, int K) {
  constexpr int N = 8 * NT;
  __shared__ float sRed[8][16][N + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * 16;
  const int ra = min(row0 + g, B - 1), rb = min(row0 + g + 8, B - 1);      // clamped: rows past the batch are computed and dropped
  const int kslice = K / 8, kbeg = warp * kslice;
  const TS* pa = src + (long long)ra * s_n + kbeg + t * 2;
  const TS* pb = src + (long long)rb * s_n + kbeg + t * 2;
  const __nv_bfloat16* pw = w + (long long)g * Kp + kbeg + t * 2;
  float acc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll 8
  for (int k = 0; k < kslice; k += 16) {
    const uint32_t a0 = skinny_a_pair(pa + k), a1 = skinny_a_pair(pb + k), a2 = skinny_a_pair(pa + k + 8), a3 = skinny_a_pair(pb + k + 8);
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const uint32_t b0 = __ldg(reinterpret_cast<const uint32_t*>(pw + (long long)j * 8 * Kp + k));
      const uint32_t b1 = __ldg(reinterpret_cast<const uint32_t*>(pw + (long long)j * 8 * Kp + k + 8));
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                   : "+f"(acc[j][0]), "+f"(acc[j][1]), "+f"(acc[j][2]), "+f"(acc[j][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
#pragma unroll
  for (int j = 0; j < NT; ++j) {   // c0, c1: (row g, cols 2t, 2t+1); c2, c3: (row g + 8, same columns)
    sRed[warp][g][j * 8 + 2 * t] = acc[j][0];
    sRed[warp][g][j * 8 + 2 * t + 1] = acc[j][1];
    sRed[warp][g + 8][j * 8 + 2 * t] = acc[j][2];
    sRed[warp][g + 8][j * 8 + 2 * t + 1] = acc[j][3];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 16 * N; o += 256) {
    const int r = o / N, n = o - r * N;
    if (row0 + r >= B) continue;
    float v = bias != nullptr ? __ldg(bias + n) : 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) v += sRed[wv][r][n];
    out[(long long)(row0 + r) * d_n + n] = v;
  }
}

// ---------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

inline int pick_bn(int Nn) {
  const int n_pad = (Nn + 15) / 16 * 16;
  return n_pad >= 128 ? 128 : n_pad > 32 ? 64 : n_pad > 16 ? 32 : 16;
}

template <int BN>
int launch(const TmapPack& tm, const TmapPack& tm_p1, const TmapPack& tm_p2, const GemmParams& p, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = NS * kAStage + NS * BN * BK * 2 + 96 /*barriers + tmem slot*/ + kTabMax * 16 + 64 + 2 * kMaxPreC * 4 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  conv_tc_kernel<BN><<<grid, kThreads, smem, st>>>(tm, tm_p1, tm_p2, p);
  CV_LAUNCH_CHECK();
  return 0;
}

template <int BN, bool MASKED>
int launch_persist(const TmapPack& tm, const GemmParams& p, cudaStream_t st) {
  constexpr size_t smem = persist_smem<BN, MASKED>();
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_persist_kernel<BN, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  // two CTAs per SM when registers (launch bounds) and shared memory (+1 KiB the driver reserves per CTA) allow it
  const int per_sm = (BN <= 64 && 2 * (smem + 1024) <= 227 * 1024) ? 2 : 1;
  int nsm = 148;
  const int grid = std::min(p.n_tiles, per_sm * nsm);
  conv_tc_persist_kernel<BN, MASKED><<<grid, kPThreads, smem, st>>>(tm, p);
  CV_LAUNCH_CHECK();
  return 0;
}
template <int BN, bool MASKED>
int launch_tma(const TmapPack& tmA, const TmapPack& tmB, const GemmParams& p, cudaStream_t st) {
  constexpr size_t smem = tma_smem<BN, MASKED>();
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tma_kernel<BN, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  // resident CTAs per SM: the launch bound (registers) and shared memory (+1 KiB the driver reserves per CTA)
  const int per_sm = std::max(1, std::min<int>(tma_ctas_per_sm<BN, MASKED>(), (int)((227 * 1024) / (smem + 1024))));
  const int grid = std::min(p.n_tiles, per_sm * 148);
  conv_tma_kernel<BN, MASKED><<<grid, kTThreads, smem, st>>>(tmA, tmB, p);
  CV_LAUNCH_CHECK();
  return 0;
}
template <int BN>
int launch_tma_bn(const TmapPack& tmA, const TmapPack& tmB, const GemmParams& p, cudaStream_t st) {
  return p.epi == CLEARVAE_EPI_BIAS_STATS ? launch_tma<BN, false>(tmA, tmB, p, st) : launch_tma<BN, true>(tmA, tmB, p, st);
}

template <int BN>
int launch_persist_bn(const TmapPack& tm, const GemmParams& p, cudaStream_t st) {
  return p.epi == CLEARVAE_EPI_BIAS_STATS ? launch_persist<BN, false>(tm, p, st) : launch_persist<BN, true>(tm, p, st);
}

void fill_t4(const clearvae_tensor4* t, const void*& ptr, long long& sn, long long& sh, long long& sw, long long& sc, int& bf) {
  ptr = t->ptr; sn = t->sn; sh = t->sh; sw = t->sw; sc = t->sc; bf = t->dtype == CLEARVAE_BF16;
}


template <int BN>
int launch_wgrad(const WgradParams& p, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = NS * kWStageA + NS * BN * WK * 2 + 96 + 2 * kMaxPreC * 4 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  wgrad_tc_kernel<BN><<<grid, kThreads, smem, st>>>(p);
  CV_LAUNCH_CHECK();
  return 0;
}

template <int BN>
int launch_wgrad_tma(const TmapPack& tmA, const TmapPack& tmB, const WgradTmaParams& p, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = wgrad_tma_smem<BN>();
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  wgrad_tma_kernel<BN><<<grid, kTThreads, smem, st>>>(tmA, tmB, p);
  CV_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

int clearvae_debug_conv_timeline(long long* device_buffer) {
  cudaError_t e = cudaMemcpyToSymbol(g_conv_timeline, &device_buffer, sizeof(device_buffer));
  return e == cudaSuccess ? 0 : (int)e;
}

static int conv_wgrad_impl(const clearvae_conv_geom* g, int64_t batch, const clearvae_tensor4* src, const float* pre_scale,
                           const float* pre_shift, int32_t pre_relu, const clearvae_tensor4* dy, float* dweight, int x3, void* stream) {
  if (!g || !src || !src->ptr || !dy || !dy->ptr || !dweight || batch <= 0) return CLEARVAE_EINVAL;
  if ((pre_scale == nullptr) != (pre_shift == nullptr)) return CLEARVAE_EINVAL;
  WgradParams p{};
  if (!cvplan::make_plan(*g, cvplan::kFprop, BK, &p.plan)) return CLEARVAE_EUNSUPPORTED;
  if (pre_scale != nullptr && p.plan.Cs > kMaxPreC) return CLEARVAE_EUNSUPPORTED;
  const int BN = pick_bn(p.plan.Nn);
  const int n_pad = (p.plan.Nn + 15) / 16 * 16;
  p.batch = batch;
  fill_t4(src, p.src, p.s_n, p.s_h, p.s_w, p.s_c, p.src_bf16);
  p.pre_scale = pre_scale; p.pre_shift = pre_shift; p.pre_relu = pre_relu;
  fill_t4(dy, p.dy, p.y_n, p.y_h, p.y_w, p.y_c, p.dy_bf16);
  p.dw = dweight;
  p.x3 = x3;
  int max_k = 0;
  long long max_m = 0;
  for (int i = 0; i < p.plan.n_classes; ++i) {
    max_k = std::max(max_k, p.plan.cls[i].ntaps * p.plan.Cs);
    max_m = std::max(max_m, (long long)batch * p.plan.cls[i].Hd * p.plan.cls[i].Wd);
  }
  const int tiles = ((max_k + 127) / 128) * ((n_pad + BN - 1) / BN) * p.plan.n_classes;
  cudaStream_t st = (cudaStream_t)stream;
  // ---- first choice: TMA operands (plain bf16 channels-last activation and gradient, 32 or a multiple of 64 channels)
  static const bool no_tma_w = getenv("CLEARVAE_NO_TMA_WGRAD") != nullptr;
  EncodeTiledFn enc = get_encode();
  const int Cs = p.plan.Cs, Nn = p.plan.Nn;
  if (!no_tma_w && enc && !x3 && p.src_bf16 && p.dy_bf16 && p.s_c == 1 && p.y_c == 1 && pre_scale == nullptr && !pre_relu &&
      (Cs == 32 || Cs % 64 == 0) && (Nn == 32 || Nn % 64 == 0) && BN >= 32 && Nn % BN == 0 && batch < (1 << 24) &&
      !(((uintptr_t)p.src | (uintptr_t)p.dy) & 15) && ((p.s_n | p.s_h | p.s_w | p.y_n | p.y_h | p.y_w) & 7) == 0) {
    WgradTmaParams q{};
    q.plan = p.plan; q.batch = batch; q.dw = dweight;
    q.cb = Cs == 32 ? 32 : 64;
    q.cbn = Nn == 32 ? 32 : 64;
    TmapPack tmA{}, tmB{};
    bool ok = true;
    long long max_box = 0;
    for (int i = 0; i < q.plan.n_classes && ok; ++i) {
      const Cls& c = q.plan.cls[i];
      const int sh = q.plan.sh, os = q.plan.os;
      const int tw = (c.Wd + 1) & ~1;                       // even box width: an odd grid gets one zero-filled column
      if (tw > 128) { ok = false; break; }
      const int th = std::min(c.Hd, 128 / tw);
      int tn = (int)std::min<long long>(batch, std::min(256, 128 / (tw * th)));
      while (tn > 1 && (tw * th * tn) % 16 != 0) --tn;      // the MMA consumes 16 pixels per instruction
      const int kpix = tw * th * tn;
      if (kpix % 16 != 0 || tw * sh > 256 || th * sh > 256 || tw * os > 256 || th * os > 256) { ok = false; break; }
      q.tw[i] = tw; q.th[i] = th; q.tn[i] = tn;
      q.tiles_h[i] = (c.Hd + th - 1) / th;
      const long long nbox = ((batch + tn - 1) / tn) * q.tiles_h[i];
      if (nbox >= (1LL << 30)) { ok = false; break; }
      q.nbox[i] = (int)nbox;
      max_box = std::max(max_box, nbox);
      {
        cuuint64_t dims[4] = {(cuuint64_t)Cs, (cuuint64_t)q.plan.Ws, (cuuint64_t)q.plan.Hs, (cuuint64_t)batch};
        cuuint64_t strides[3] = {(cuuint64_t)p.s_w * 2, (cuuint64_t)p.s_h * 2, (cuuint64_t)p.s_n * 2};
        cuuint32_t box[4] = {(cuuint32_t)q.cb, (cuuint32_t)(tw * sh), (cuuint32_t)(th * sh), (cuuint32_t)tn};
        cuuint32_t estr[4] = {1, (cuuint32_t)sh, (cuuint32_t)sh, 1};
        if (q.plan.Ws == 1) strides[0] = (cuuint64_t)Cs * 2;
        if (q.plan.Hs == 1) strides[1] = strides[0] * (cuuint64_t)q.plan.Ws;
        if ((strides[0] | strides[1] | strides[2]) & 15) { ok = false; break; }
        if (enc(&tmA.t[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.src), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                q.cb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { ok = false; break; }
      }
      {
        cuuint64_t dims[4] = {(cuuint64_t)Nn, (cuuint64_t)q.plan.Wb, (cuuint64_t)q.plan.Hb, (cuuint64_t)batch};
        cuuint64_t strides[3] = {(cuuint64_t)p.y_w * 2, (cuuint64_t)p.y_h * 2, (cuuint64_t)p.y_n * 2};
        cuuint32_t box[4] = {(cuuint32_t)q.cbn, (cuuint32_t)(tw * os), (cuuint32_t)(th * os), (cuuint32_t)tn};
        cuuint32_t estr[4] = {1, (cuuint32_t)os, (cuuint32_t)os, 1};
        if (q.plan.Wb == 1) strides[0] = (cuuint64_t)Nn * 2;
        if (q.plan.Hb == 1) strides[1] = strides[0] * (cuuint64_t)q.plan.Wb;
        if ((strides[0] | strides[1] | strides[2]) & 15) { ok = false; break; }
        if (enc(&tmB.t[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.dy), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                q.cbn == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { ok = false; break; }
      }
    }
    if (ok) {
      // split the pixel boxes so the grid is ONE wave of (one-CTA-per-SM) CTAs, each keeping >= 2 boxes: with TMA operands a box
      // costs ~0.4 us of main loop while the scatter-add epilogue costs 2-9 us per CTA and grows with the number of splits
      long long sp = std::max<long long>(1, 148 / tiles);
      sp = std::min<long long>(sp, std::max<long long>(1, max_box / 2));
      q.splits = (int)sp;
      dim3 grid((unsigned)((max_k + 127) / 128), (unsigned)(Nn / BN), (unsigned)(q.plan.n_classes * q.splits));
      switch (BN) {
        case 32: return launch_wgrad_tma<32>(tmA, tmB, q, grid, st);
        case 64: return launch_wgrad_tma<64>(tmA, tmB, q, grid, st);
        default: return launch_wgrad_tma<128>(tmA, tmB, q, grid, st);
      }
    }
  }
  // split the pixel reduction so the grid is ~2 waves of 148 SMs, each CTA keeping >= 4 k-blocks
  long long splits = std::max<long long>(1, (2 * 148 + tiles - 1) / tiles);
  splits = std::min<long long>(splits, std::max<long long>(1, max_m / (WK * 4)));
  p.splits = (int)splits;
  dim3 grid((unsigned)((max_k + 127) / 128), (unsigned)((n_pad + BN - 1) / BN), (unsigned)(p.plan.n_classes * p.splits));
  switch (BN) {
    case 16: return launch_wgrad<16>(p, grid, st);
    case 32: return launch_wgrad<32>(p, grid, st);
    case 64: return launch_wgrad<64>(p, grid, st);
    default: return launch_wgrad<128>(p, grid, st);
  }
}

int clearvae_conv_wgrad(const clearvae_conv_geom* g, int64_t batch, const clearvae_tensor4* src, const float* pre_scale,
                        const float* pre_shift, int32_t pre_relu, const clearvae_tensor4* dy, float* dweight, void* stream) {
  return conv_wgrad_impl(g, batch, src, pre_scale, pre_shift, pre_relu, dy, dweight, 0, stream);
}
int clearvae_conv_wgrad_split3(const clearvae_conv_geom* g, int64_t batch, const clearvae_tensor4* src, const float* pre_scale,
                               const float* pre_shift, int32_t pre_relu, const clearvae_tensor4* dy, float* dweight, void* stream) {
  return conv_wgrad_impl(g, batch, src, pre_scale, pre_shift, pre_relu, dy, dweight, 1, stream);
}

size_t clearvae_conv_packed_weight_bytes(const clearvae_conv_geom* g, int32_t role) {
  Plan plan;
  const int x3 = (role & CLEARVAE_ROLE_SPLIT3) != 0;
  if (!g || !cvplan::make_plan(*g, role & ~CLEARVAE_ROLE_SPLIT3, BK, &plan)) return 0;
  // split mode: [part 0 | part 1 | part 2], each part starting at a 128-byte boundary
  const size_t one = ((size_t)cvplan::packed_weight_elems(plan) * 2 + 127) / 128 * 128;
  return x3 ? 3 * one : (size_t)cvplan::packed_weight_elems(plan) * 2;
}

int clearvae_conv_pack_weight(const clearvae_conv_geom* g, int32_t role, const float* weight, void* packed, void* stream) {
  if (!g || !weight || !packed) return CLEARVAE_EINVAL;
  Plan plan;
  const int x3 = (role & CLEARVAE_ROLE_SPLIT3) != 0;
  if (!cvplan::make_plan(*g, role & ~CLEARVAE_ROLE_SPLIT3, BK, &plan)) return CLEARVAE_EUNSUPPORTED;
  const int n_pad = (plan.Nn + 15) / 16 * 16;
  long long mx = 0;
  for (int i = 0; i < plan.n_classes; ++i) mx = std::max(mx, (long long)n_pad * plan.cls[i].Kp);
  dim3 grid((unsigned)std::min<long long>((mx + 255) / 256, 148 * 8), (unsigned)plan.n_classes);
  const size_t one = ((size_t)cvplan::packed_weight_elems(plan) * 2 + 127) / 128 * 128;
  __nv_bfloat16* p1 = x3 ? reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(packed) + one) : nullptr;
  __nv_bfloat16* p2 = x3 ? reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(packed) + 2 * one) : nullptr;
  pack_weight_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(plan, weight, reinterpret_cast<__nv_bfloat16*>(packed), p1, p2);
  CV_LAUNCH_CHECK();
  return 0;
}

size_t clearvae_conv_pack_multi_table_bytes(int32_t n_weights) {
  return n_weights > 0 ? (size_t)n_weights * cvplan::kMaxClasses * sizeof(MultiPackEntry) : 0;
}

int clearvae_conv_pack_multi_build(int32_t n_weights, const clearvae_conv_geom* geoms_host, const int32_t* roles_host,
                                   const float* const* weights, void* const* packed, void* table_host, int32_t* n_entries,
                                   int32_t* n_blocks) {
  if (n_weights <= 0 || !geoms_host || !roles_host || !weights || !packed || !table_host || !n_entries || !n_blocks) return CLEARVAE_EINVAL;
  MultiPackEntry* tab = reinterpret_cast<MultiPackEntry*>(table_host);
  int ne = 0, nb = 0;
  for (int i = 0; i < n_weights; ++i) {
    if (!weights[i] || !packed[i] || (roles_host[i] & CLEARVAE_ROLE_SPLIT3)) return CLEARVAE_EINVAL;   // split packs stay per weight
    Plan plan;
    if (!cvplan::make_plan(geoms_host[i], roles_host[i], BK, &plan)) return CLEARVAE_EUNSUPPORTED;
    const int n_pad = (plan.Nn + 15) / 16 * 16;
    for (int c = 0; c < plan.n_classes; ++c) {
      MultiPackEntry& e = tab[ne++];
      e.plan = plan;
      e.w = weights[i];
      e.out = reinterpret_cast<__nv_bfloat16*>(packed[i]);
      e.cls = c;
      e.block0 = nb;
      e.nblocks = (int)std::min<long long>(((long long)n_pad * plan.cls[c].Kp + 255) / 256, 64);
      e.pad = 0;
      nb += e.nblocks;
    }
  }
  *n_entries = ne;
  *n_blocks = nb;
  return 0;
}

int clearvae_conv_pack_multi_launch(const void* table_device, int32_t n_entries, int32_t n_blocks, void* stream) {
  if (!table_device || n_entries <= 0 || n_blocks <= 0) return CLEARVAE_EINVAL;
  pack_weight_multi_kernel<<<n_blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const MultiPackEntry*>(table_device), n_entries);
  CV_LAUNCH_CHECK();
  return 0;
}

int clearvae_conv_gemm(const clearvae_conv_geom* g, int32_t role, int64_t batch, const clearvae_tensor4* src,
                       const float* pre_scale, const float* pre_shift, int32_t pre_relu, const void* packed_weight,
                       const float* bias, const clearvae_tensor4* dst, int32_t epilogue, const clearvae_tensor4* mask_src,
                       const float* mask_scale, const float* mask_shift, double* stats, void* stream) {
  if (!g || !src || !src->ptr || !dst || !dst->ptr || !packed_weight || batch <= 0) return CLEARVAE_EINVAL;
  if (epilogue == CLEARVAE_EPI_MASK_STATS && (!mask_src || !mask_src->ptr)) return CLEARVAE_EINVAL;
  if ((pre_scale == nullptr) != (pre_shift == nullptr)) return CLEARVAE_EINVAL;
  if ((uintptr_t)packed_weight & 127) return CLEARVAE_EINVAL;
  GemmParams p{};
  p.x3 = (role & CLEARVAE_ROLE_SPLIT3) != 0;
  role &= ~CLEARVAE_ROLE_SPLIT3;
  if (!cvplan::make_plan(*g, role, BK, &p.plan)) return CLEARVAE_EUNSUPPORTED;
  if (pre_scale != nullptr && p.plan.Cs > kMaxPreC) return CLEARVAE_EUNSUPPORTED;
  {
    // skinny linear layers (N = 16 / 32, deep K, plain bf16 rows, bias-only epilogue): warp-MMA kernel without split-K atomics
    static const bool no_skinny = getenv("CLEARVAE_NO_SKINNY_LINEAR") != nullptr;
    const Cls& c0 = p.plan.cls[0];
    const int Nn = p.plan.Nn, Cs = p.plan.Cs;
    if (!no_skinny && !p.x3 && g->k == 1 && g->Hin == 1 && g->Win == 1 && p.plan.n_classes == 1 && (Nn == 16 || Nn == 32) && Cs % 128 == 0 &&
        Cs >= 512 && c0.Kp >= Cs && c0.w_off == 0 && src->sc == 1 && (src->sn & 1) == 0 && !((uintptr_t)src->ptr & 7) && pre_scale == nullptr && !pre_relu && epilogue == CLEARVAE_EPI_BIAS_STATS && stats == nullptr &&
        dst->dtype == CLEARVAE_F32 && dst->sc == 1 && batch < (1LL << 27)) {
      const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(packed_weight);
      const unsigned grid = (unsigned)((batch + 15) / 16);
      cudaStream_t st = (cudaStream_t)stream;
#define CV_SKINNY(NT, TS) skinny_linear_kernel<NT, TS><<<grid, 256, 0, st>>>(reinterpret_cast<const TS*>(src->ptr), src->sn, wp, c0.Kp, bias, \
                                                                              (float*)dst->ptr, dst->sn, (int)batch, Cs)
      if (src->dtype == CLEARVAE_BF16) { if (Nn == 32) CV_SKINNY(4, __nv_bfloat16); else CV_SKINNY(2, __nv_bfloat16); }
      else { if (Nn == 32) CV_SKINNY(4, float); else CV_SKINNY(2, float); }
#undef CV_SKINNY
      CV_LAUNCH_CHECK();
      return 0;
    }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) return CLEARVAE_EUNSUPPORTED;
  const int BN = pick_bn(p.plan.Nn);
  const int n_pad = (p.plan.Nn + 15) / 16 * 16;
  TmapPack tm{}, tm_p1{}, tm_p2{};
  long long max_m = 0;
  const size_t part_off = ((size_t)cvplan::packed_weight_elems(p.plan) * 2 + 127) / 128 * 128;   // split mode: [part 0 | 1 | 2]
  for (int i = 0; i < p.plan.n_classes; ++i) {
    const Cls& c = p.plan.cls[i];
    cuuint64_t dims[2] = {(cuuint64_t)c.Kp, (cuuint64_t)n_pad};
    cuuint64_t strides[1] = {(cuuint64_t)c.Kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    for (int h = 0; h < (p.x3 ? 3 : 1); ++h) {
      void* base = (void*)(reinterpret_cast<const char*>(packed_weight) + (size_t)c.w_off * 2 + h * part_off);
      CUresult r = enc(h == 0 ? &tm.t[i] : h == 1 ? &tm_p1.t[i] : &tm_p2.t[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return CLEARVAE_EINVAL;
    }
    max_m = std::max(max_m, (long long)batch * c.Hd * c.Wd);
  }
  p.batch = batch;
  fill_t4(src, p.src, p.s_n, p.s_h, p.s_w, p.s_c, p.src_bf16);
  p.pre_scale = pre_scale; p.pre_shift = pre_shift; p.pre_relu = pre_relu;
  p.bias = bias;
  { const void* q; fill_t4(dst, q, p.d_n, p.d_h, p.d_w, p.d_c, p.dst_bf16); p.dst = const_cast<void*>(q); }
  p.epi = epilogue;
  if (mask_src) fill_t4(mask_src, p.msk, p.m_n, p.m_h, p.m_w, p.m_c, p.msk_bf16);
  p.msk_scale = mask_scale; p.msk_shift = mask_shift;
  p.stats = stats;
  p.splits = 1;
  cudaStream_t st = (cudaStream_t)stream;
  {
    // skinny-M / deep-K linear layers (the latent heads, the fc data-gradient): split K across CTAs so the
    // GPU is filled; partial sums meet in fp32 atomics on a zeroed destination
    const long long ctas = ((max_m + BM - 1) / BM) * ((n_pad + BN - 1) / BN) * p.plan.n_classes;
    const int nkb = p.plan.cls[0].Kp / BK;
    const bool dense_dst = !p.dst_bf16 && p.d_c == 1 && p.d_n == p.plan.Nn && g->k == 1 && g->Hin == 1 && g->Win == 1;
    if (dense_dst && stats == nullptr && epilogue == CLEARVAE_EPI_BIAS_STATS && nkb >= 8 && ctas * 2 <= 148) {
      int sp = (int)std::min<long long>(nkb / 2, 148 / ctas);
      if (sp > 1) {
        p.splits = sp;
        cudaError_t e = cudaMemsetAsync(p.dst, 0, (size_t)batch * p.plan.Nn * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
      }
    }
  }
  // common case -> persistent warp-specialised kernel (see conv_tc_persist_kernel)
  static const bool no_persist = getenv("CLEARVAE_NO_PERSIST") != nullptr;
  const bool masked = epilogue != CLEARVAE_EPI_BIAS_STATS;
  if (!no_persist && !p.x3 && p.splits == 1 && p.src_bf16 && p.s_c == 1 && p.plan.Cs % 8 == 0 && pre_scale == nullptr && !pre_relu &&
      p.d_c == 1 && p.plan.Nn % BN == 0 && ((p.d_n | p.d_h | p.d_w) & 7) == 0 && !((uintptr_t)p.src & 15) &&
      ((p.s_n | p.s_h | p.s_w) & 7) == 0 && !((uintptr_t)p.dst & 15) &&
      (!masked || (p.m_c == 1 && ((p.m_n | p.m_h | p.m_w) & 7) == 0 && !((uintptr_t)p.msk & 15)))) {
    // ---- first choice: the TMA-operand kernel (one bulk-tensor instruction per tap instead of 1024 cp.async)
    static const bool no_tma_a = getenv("CLEARVAE_NO_TMA_A") != nullptr;
    const int Cs = p.plan.Cs;
    if (!no_tma_a && (Cs == 32 || Cs % 64 == 0) && batch < (1 << 24)) {
      GemmParams q = p;
      q.t_cb = Cs == 32 ? 32 : 64;
      bool ok = true;
      long long tiles = 0;
      const int n_ntiles = q.plan.Nn / BN;
      TmapPack tmA{}, tmB = tm;
      for (int i = 0; i < q.plan.n_classes && ok; ++i) {
        const Cls& c = q.plan.cls[i];
        const int sh = q.plan.sh;
        if (c.Wd > 128 || c.Wd * sh > 256) { ok = false; break; }
        int tw = c.Wd, th, tn;
        if (c.Hd * c.Wd <= 128) { th = c.Hd; tn = std::min<long long>(128 / (c.Hd * c.Wd), batch); }
        else { th = 128 / c.Wd; tn = 1; }
        if (th * sh > 256 || tn > 256) { ok = false; break; }
        const int tiles_h = (c.Hd + th - 1) / th;
        const long long mt = ((batch + tn - 1) / tn) * tiles_h;
        q.t_tw[i] = tw; q.t_th[i] = th; q.t_tn[i] = tn; q.t_tiles_h[i] = tiles_h; q.t_mtiles[i] = (int)mt;
        q.tile_start[i] = (int)tiles;
        tiles += mt * n_ntiles;
        cuuint64_t dims[4] = {(cuuint64_t)Cs, (cuuint64_t)q.plan.Ws, (cuuint64_t)q.plan.Hs, (cuuint64_t)batch};
        cuuint64_t strides[3] = {(cuuint64_t)q.s_w * 2, (cuuint64_t)q.s_h * 2, (cuuint64_t)q.s_n * 2};
        cuuint32_t box[4] = {(cuuint32_t)q.t_cb, (cuuint32_t)(tw * sh), (cuuint32_t)(th * sh), (cuuint32_t)tn};
        cuuint32_t estr[4] = {1, (cuuint32_t)sh, (cuuint32_t)sh, 1};
        // a Linear is a 1x1 "image" per sample (Hs = Ws = 1): strides of the unit dimensions are free but must be legal
        if (q.plan.Ws == 1) strides[0] = (cuuint64_t)Cs * 2;
        if (q.plan.Hs == 1) strides[1] = strides[0] * (cuuint64_t)q.plan.Ws;
        if ((strides[0] | strides[1] | strides[2]) & 15) { ok = false; break; }
        CUresult r = enc(&tmA.t[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(q.src), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, q.t_cb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { ok = false; break; }
        if (q.t_cb == 32) {   // 64-byte weight rows: its own map (box 32 x BN, SWIZZLE_64B)
          cuuint64_t wd[2] = {(cuuint64_t)c.Kp, (cuuint64_t)n_pad};
          cuuint64_t ws[1] = {(cuuint64_t)c.Kp * 2};
          cuuint32_t wb[2] = {32u, (cuuint32_t)BN};
          cuuint32_t we[2] = {1, 1};
          void* base = (void*)(reinterpret_cast<const char*>(packed_weight) + (size_t)c.w_off * 2);
          r = enc(&tmB.t[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, wd, ws, wb, we, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r != CUDA_SUCCESS) { ok = false; break; }
        }
      }
      if (ok && tiles > 0 && tiles < (1LL << 30)) {
        q.tile_start[q.plan.n_classes] = (int)tiles;
        q.n_tiles = (int)tiles;
        switch (BN) {
          case 16: return launch_tma_bn<16>(tmA, tmB, q, st);
          case 32: return launch_tma_bn<32>(tmA, tmB, q, st);
          case 64: return launch_tma_bn<64>(tmA, tmB, q, st);
          default: return launch_tma_bn<128>(tmA, tmB, q, st);
        }
      }
    }
    long long tiles = 0;
    const int n_ntiles = p.plan.Nn / BN;
    for (int i = 0; i < p.plan.n_classes; ++i) {
      p.tile_start[i] = (int)tiles;
      tiles += (((long long)batch * p.plan.cls[i].Hd * p.plan.cls[i].Wd + BM - 1) / BM) * n_ntiles;
    }
    p.tile_start[p.plan.n_classes] = (int)tiles;
    if (tiles > 0 && tiles < (1LL << 30)) {
      p.n_tiles = (int)tiles;
      switch (BN) {
        case 16: return launch_persist_bn<16>(tm, p, st);
        case 32: return launch_persist_bn<32>(tm, p, st);
        case 64: return launch_persist_bn<64>(tm, p, st);
        default: return launch_persist_bn<128>(tm, p, st);
      }
    }
  }
  dim3 grid((unsigned)((max_m + BM - 1) / BM), (unsigned)((n_pad + BN - 1) / BN), (unsigned)(p.plan.n_classes * p.splits));
  switch (BN) {
    case 16: return launch<16>(tm, tm_p1, tm_p2, p, grid, st);
    case 32: return launch<32>(tm, tm_p1, tm_p2, p, grid, st);
    case 64: return launch<64>(tm, tm_p1, tm_p2, p, grid, st);
    default: return launch<128>(tm, tm_p1, tm_p2, p, grid, st);
  }
}

}  // extern "C"
