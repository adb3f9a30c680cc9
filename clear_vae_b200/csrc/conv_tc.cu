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
constexpr int kThreads = 448;   // conv GEMM: 2 x 4 producer warps (alternating k-blocks), TMA, MMA, 4 epilogue warps
constexpr int kWThreads = 192;  // wgrad GEMM: 4 producer/epilogue warps, (idle), MMA
constexpr int kMaxPreC = 2048;  // source channels whose BatchNorm scale/shift are staged in shared memory

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
  int splits;  // split-K (fp32 atomic epilogue); 1 = off
  int n_tiles, n_ntiles;                     // persistent scheduler: total work items, n-tiles per m-tile
  int tile_start[cvplan::kMaxClasses];       // first work item of each class
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
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

// ---- persistent tile scheduler -------------------------------------------------------
struct Tile {
  int cls, split, n0, kb0, nkb;
  long long m0;
};
__device__ __forceinline__ bool get_tile(const GemmParams& p, int tile_id, int BN, Tile* t) {
  if (tile_id >= p.n_tiles) return false;
  int cls = 0;
#pragma unroll
  for (int i = 1; i < cvplan::kMaxClasses; ++i)
    if (i < p.plan.n_classes && tile_id >= p.tile_start[i]) cls = i;
  int local = tile_id - p.tile_start[cls];
  const int nt = local % p.n_ntiles;
  local /= p.n_ntiles;
  const int split = local % p.splits;
  const int mt = local / p.splits;
  const int nkb_all = p.plan.cls[cls].Kp / BK;
  t->cls = cls; t->split = split; t->n0 = nt * BN; t->m0 = (long long)mt * BM;
  t->kb0 = nkb_all * split / p.splits;
  t->nkb = nkb_all * (split + 1) / p.splits - t->kb0;
  return true;
}

struct RowCtx {
  bool mvalid;
  int hd, wd;
  long long img, base;
  unsigned tapmask;
};
__device__ __forceinline__ void decode_row(const GemmParams& p, const Cls& c, long long m, RowCtx* r) {
  const long long Mc = p.batch * c.Hd * c.Wd;
  r->mvalid = m < Mc;
  const long long mm = r->mvalid ? m : 0;
  r->wd = (int)(mm % c.Wd);
  r->hd = (int)((mm / c.Wd) % c.Hd);
  r->img = mm / ((long long)c.Wd * c.Hd);
  const int hbase = r->hd * p.plan.sh, wbase = r->wd * p.plan.sh;
  r->base = r->img * p.s_n + hbase * p.s_h + wbase * p.s_w;
  unsigned mask = 0;
  if (r->mvalid) {
#pragma unroll 1
    for (int t = 0; t < c.ntaps; ++t) {
      const int hs = hbase + c.dh[t], ws = wbase + c.dw[t];
      mask |= (hs >= 0 && hs < p.plan.Hs && ws >= 0 && ws < p.plan.Ws) ? (1u << t) : 0u;
    }
  }
  r->tapmask = mask;
}

// rare: scalar-gather layers whose K exceeds the table (kept out of line so the hot loop stays small)
__device__ __noinline__ void slow_k_lookup(const GemmParams& p, const Cls& c, int k, int Cs, int* t, int* ch, int* koff) {
  *t = k / Cs;
  *ch = k - *t * Cs;
  *koff = (int)(c.dh[*t] * p.s_h + c.dw[*t] * p.s_w + *ch * p.s_c);
}

constexpr int kTabPerCls = 64;   // k -> (tap, channel, offset) entries per class for the scalar gather (K <= 64 there)

// Persistent, warp-specialised implicit-GEMM kernel.  Each CTA walks a static round-robin list of
// (class, m-tile, split, n-tile) work items; producers, the TMA warp, the MMA warp and the epilogue
// warps each run that list on their own, coupled only through mbarriers:
//   smem stage ring   full[NS] / empty[NS]          (producers + TMA  ->  MMA)
//   TMEM accumulators acc_full[2] / acc_empty[2]    (MMA -> epilogue), double buffered so the epilogue of
//                                                   tile i overlaps the gather + MMA of tile i+1.
template <int BN, bool SRC_BF16, int EPI>
__global__ void __launch_bounds__(kThreads) conv_tc_kernel(const __grid_constant__ TmapPack tm, const GemmParams p) {
  constexpr int kBStage = BN * BK * 2;
  constexpr uint32_t kAccCols = BN < 32 ? 32 : BN;
  constexpr uint32_t kTmemCols = 2 * kAccCols;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;
  unsigned char* sB = smem + NS * kAStage;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + NS * kBStage);
  uint64_t* empty = full + NS;
  uint64_t* acc_full = empty + NS;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  int4* sTab = reinterpret_cast<int4*>(reinterpret_cast<unsigned char*>(full) + 128);   // [classes][kTabPerCls]
  int* sTapOff = reinterpret_cast<int*>(sTab + cvplan::kMaxClasses * kTabPerCls);       // [classes][16]
  float* sStat = reinterpret_cast<float*>(sTapOff + cvplan::kMaxClasses * cvplan::kMaxTaps);  // [4 warps][2][BN]
  float* sScale = sStat + 4 * 2 * 128;
  float* sShift = sScale + p.plan.Cs;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Cs = p.plan.Cs;
  const bool vec = (Cs % 8 == 0) && p.s_c == 1;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], BM + 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    fence_barrier_init();
  }
  if (warp == 9) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  if (warp == 8 && lane == 0)
    for (int i = 0; i < p.plan.n_classes; ++i) prefetch_tmap(&tm.t[i]);
  if (p.pre_scale != nullptr)
    for (int i = threadIdx.x; i < Cs; i += kThreads) { sScale[i] = __ldg(p.pre_scale + i); sShift[i] = __ldg(p.pre_shift + i); }
  for (int ci = 0; ci < p.plan.n_classes; ++ci) {
    const Cls& c = p.plan.cls[ci];
    const int Kreal = c.ntaps * Cs;
    for (int k = threadIdx.x; k < kTabPerCls && k < Kreal; k += kThreads) {
      const int t = k / Cs, ch = k - t * Cs;
      sTab[ci * kTabPerCls + k] = make_int4(t, ch, (int)(c.dh[t] * p.s_h + c.dw[t] * p.s_w + ch * p.s_c), 0);
    }
    if (threadIdx.x < cvplan::kMaxTaps)
      sTapOff[ci * cvplan::kMaxTaps + threadIdx.x] =
          threadIdx.x < c.ntaps ? (int)(c.dh[threadIdx.x] * p.s_h + c.dw[threadIdx.x] * p.s_w) : 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool has_pre = p.pre_scale != nullptr;

  if (warp < 8) {
    // ================= A producers: one thread per tile row; two groups of 4 warps take alternate k-blocks, so
    // two k-blocks' worth of gathers are in flight per CTA (memory-level parallelism is what bounds this kernel)
    const int r = threadIdx.x & 127;
    const uint32_t grp = threadIdx.x >> 7;
    uint32_t kbg = 0;  // k-blocks produced so far by this CTA (stage ring position)
    Tile tl;
    for (int tile = blockIdx.x; get_tile(p, tile, BN, &tl); tile += gridDim.x) {
      const Cls& c = p.plan.cls[tl.cls];
      const int Kreal = c.ntaps * Cs;
      RowCtx rc;
      decode_row(p, c, tl.m0 + r, &rc);
      const int* tapoff = sTapOff + tl.cls * cvplan::kMaxTaps;
      const int4* tab = sTab + tl.cls * kTabPerCls;
      for (int kb = 0; kb < tl.nkb; ++kb, ++kbg) {
        if ((kbg & 1u) != grp) continue;
        const int s = kbg % NS;
        const uint32_t ph = ((kbg / NS) & 1) ^ 1;
        unsigned char* a_st = sA + s * kAStage;
        if (vec) {
          int t_cur = ((tl.kb0 + kb) * BK) / Cs, c_cur = ((tl.kb0 + kb) * BK) % Cs;
          // phase 1: every load of the k-block in flight (clamped address, masked afterwards)
          uint4 q0[8], q1[8];
          int cch[8];
          bool ok[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int kg = (tl.kb0 + kb) * BK + j * 8;
            const bool v = kg < Kreal && ((rc.tapmask >> t_cur) & 1u);
            const long long off = v ? rc.base + tapoff[t_cur & (cvplan::kMaxTaps - 1)] + c_cur : 0;
            ok[j] = v;
            cch[j] = c_cur;
            if (SRC_BF16) {
              q0[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.src) + off));
            } else {
              q0[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.src) + off));
              q1[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.src) + off + 4));
            }
            c_cur += 8;
            if (c_cur >= Cs) { c_cur = 0; ++t_cur; }
          }
          mbar_wait(&empty[s], ph);
          // phase 2: BatchNorm-apply + ReLU (scale/shift from shared memory), bf16 pack, 16-byte stores
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v[8];
            if (SRC_BF16) {
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
            uint4 out = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            if (!ok[j]) out = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(a_st + j * (BM * 16) + r * 16) = out;
          }
        } else {
          mbar_wait(&empty[s], ph);
#pragma unroll 1
          for (int j = 0; j < 8; ++j) {
            const int kg = (tl.kb0 + kb) * BK + j * 8;
            if (kg >= Kreal || rc.tapmask == 0u) {  // zero padding of K (or a row outside the problem)
              *reinterpret_cast<uint4*>(a_st + j * (BM * 16) + r * 16) = make_uint4(0u, 0u, 0u, 0u);
              continue;
            }
            float v[8];
            bool ok[8];
            int chn[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {  // eight independent loads; (tap, channel, offset) from the shared table
              const int k = kg + i;
              bool vld = false;
              long long off = 0;
              int ch = 0;
              if (k < Kreal) {
                int t, koff;
                if (k < kTabPerCls) { const int4 e = tab[k]; t = e.x; ch = e.y; koff = e.z; }
                else slow_k_lookup(p, c, k, Cs, &t, &ch, &koff);
                vld = (rc.tapmask >> t) & 1u;
                off = vld ? rc.base + koff : 0;
              }
              ok[i] = vld;
              chn[i] = ch;
              v[i] = ld_elem(p.src, off, SRC_BF16);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float x = v[i];
              if (has_pre) x = fmaf(x, sScale[chn[i]], sShift[chn[i]]);
              if (p.pre_relu) x = fmaxf(x, 0.f);
              v[i] = ok[i] ? x : 0.f;
            }
            *reinterpret_cast<uint4*>(a_st + j * (BM * 16) + r * 16) =
                make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          }
        }
        fence_proxy_async();
        mbar_arrive(&full[s]);
      }
    }
  } else if (warp == 8) {
    // ================= B producer: TMA =================
    if (lane == 0) {
      uint32_t kbg = 0;
      Tile tl;
      for (int tile = blockIdx.x; get_tile(p, tile, BN, &tl); tile += gridDim.x) {
        for (int kb = 0; kb < tl.nkb; ++kb, ++kbg) {
          const int s = kbg % NS;
          mbar_wait(&empty[s], ((kbg / NS) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[s], kBStage);
          tma_load_2d(sB + s * kBStage, &tm.t[tl.cls], &full[s], (tl.kb0 + kb) * BK, tl.n0);
        }
      }
    }
  } else if (warp == 9) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc(kFmtBF16, BM, BN, 0, 0);
      uint32_t kbg = 0, tcount = 0;
      Tile tl;
      for (int tile = blockIdx.x; get_tile(p, tile, BN, &tl); tile += gridDim.x, ++tcount) {
        const uint32_t b = tcount & 1;
        mbar_wait(&acc_empty[b], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < tl.nkb; ++kb, ++kbg) {
          const int s = kbg % NS;
          mbar_wait(&full[s], (kbg / NS) & 1);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + s * kAStage), b_base = smem_u32(sB + s * kBStage);
#pragma unroll
          for (int k4 = 0; k4 < BK / 16; ++k4) {
            // A: interleave layout, 16-byte k-chunks BM*16 bytes apart (LBO), 8-row groups 128 bytes apart (SBO)
            const uint64_t ad = smem_desc(a_base + k4 * 2 * (BM * 16), BM * 16, 128, kLayoutNone);
            // B: 128-byte swizzled rows, 8-row groups 1024 bytes apart; K advance = +32 bytes inside the atom
            const uint64_t bd = smem_desc(b_base + k4 * 32, 16, 1024, kLayoutSw128);
            umma_f16(tmem_base + b * kAccCols, ad, bd, idesc, (kb | k4) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[b]);
      }
    }
  } else {
    // ================= epilogue warps: TMEM -> registers -> global =================
    const int ew = warp - 10;
    const int lane_grp = warp & 3;            // TMEM lanes 32*(warp%4) .. +31
    const int r = lane_grp * 32 + lane;
    const int Nn = p.plan.Nn;
    constexpr int CH = BN < 32 ? BN : 32;
    float* myAcc = sStat + ew * 2 * 128;       // running per-column statistics of this warp: [2][BN], lane j owns column ch0 + j
    for (int i = lane; i < 2 * 128; i += 32) myAcc[i] = 0.f;
    __syncwarp();
    int acc_n0 = -1;
    auto flush_stats = [&]() {
      // combine the four epilogue warps in shared memory: one double atomic per column reaches L2
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (acc_n0 >= 0) {
        for (int col = threadIdx.x - 10 * 32; col < BN; col += 128) {
          const int n = acc_n0 + col;
          if (n < Nn) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) { a += sStat[(w * 2 + 0) * 128 + col]; b += sStat[(w * 2 + 1) * 128 + col]; }
            atomicAdd(p.stats + n, (double)a);
            atomicAdd(p.stats + Nn + n, (double)b);
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = lane; i < 2 * 128; i += 32) myAcc[i] = 0.f;
      __syncwarp();
    };
    uint32_t tcount = 0;
    Tile tl;
    for (int tile = blockIdx.x; get_tile(p, tile, BN, &tl); tile += gridDim.x, ++tcount) {
      const Cls& c = p.plan.cls[tl.cls];
      const long long Mc = p.batch * c.Hd * c.Wd;
      const long long m = tl.m0 + r;
      const bool mvalid = m < Mc;
      const long long mm = mvalid ? m : 0;
      const int wd = (int)(mm % c.Wd), hd = (int)((mm / c.Wd) % c.Hd);
      const long long img = mm / ((long long)c.Wd * c.Hd);
      const long long dst_off = img * p.d_n + (long long)(hd * p.plan.os + c.oa) * p.d_h + (long long)(wd * p.plan.os + c.ob) * p.d_w;
      const long long msk_off = img * p.m_n + (long long)(hd * p.plan.os + c.oa) * p.m_h + (long long)(wd * p.plan.os + c.ob) * p.m_w;
      if (p.stats != nullptr && tl.n0 != acc_n0) {  // column block changed: publish what was accumulated so far
        if (acc_n0 >= 0) flush_stats();
        acc_n0 = tl.n0;
      }
      const uint32_t b = tcount & 1;
      mbar_wait(&acc_full[b], (tcount >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int ch0 = 0; ch0 < BN; ch0 += CH) {
        uint32_t raw[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(b * kAccCols + ch0);
        if (CH == 32) {
          tmem_ld32(taddr, raw);
        } else {
          uint32_t r16[16];
          tmem_ld16(taddr, r16);
#pragma unroll
          for (int i = 0; i < 16; ++i) { raw[i] = r16[i]; raw[i + 16] = 0u; }
        }
        tmem_ld_wait();
        const int nbase = tl.n0 + ch0;
        const int nlive = mvalid ? min(CH, Nn - nbase) : 0;  // leading columns of this chunk that exist
        float v[32], u[32];
        if (EPI == CLEARVAE_EPI_BIAS_STATS) {
          const bool add_bias = p.bias != nullptr && tl.split == 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float a = __uint_as_float(raw[i]);
            if (add_bias && i < nlive) a += __ldg(p.bias + nbase + i);
            a = i < nlive ? a : 0.f;
            v[i] = a;
            u[i] = a * a;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float a = 0.f, second = 0.f;
            if (i < nlive) {
              const int n = nbase + i;
              const float y = ld_elem(p.msk, msk_off + n * p.m_c, p.msk_bf16);
              const float act = p.msk_scale ? fmaf(y, __ldg(p.msk_scale + n), __ldg(p.msk_shift + n)) : y;
              a = act > 0.f ? __uint_as_float(raw[i]) : 0.f;
              second = a * y;
            }
            v[i] = a;
            u[i] = second;
          }
        }
        // ---- store (vectorised when channels are the innermost dst dimension)
        if (mvalid && p.splits > 1) {
          float* d = reinterpret_cast<float*>(p.dst) + dst_off + (long long)nbase * p.d_c;
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            if (i < nlive) atomicAdd(d, v[i]);
            d += p.d_c;
          }
        } else if (mvalid) {
          if (p.d_c == 1 && nlive == CH && (CH % 8) == 0) {
            if (p.dst_bf16) {
              __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.dst) + dst_off + nbase;
#pragma unroll
              for (int i = 0; i < CH; i += 8)
                *reinterpret_cast<uint4*>(d + i) = make_uint4(pack_bf16(v[i], v[i + 1]), pack_bf16(v[i + 2], v[i + 3]),
                                                              pack_bf16(v[i + 4], v[i + 5]), pack_bf16(v[i + 6], v[i + 7]));
            } else {
              float* d = reinterpret_cast<float*>(p.dst) + dst_off + nbase;
#pragma unroll
              for (int i = 0; i < CH; i += 4) *reinterpret_cast<float4*>(d + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
          } else if (p.dst_bf16) {
            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.dst) + dst_off + (long long)nbase * p.d_c;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
              if (i < nlive) *d = __float2bfloat16(v[i]);
              d += p.d_c;
            }
          } else {
            float* d = reinterpret_cast<float*>(p.dst) + dst_off + (long long)nbase * p.d_c;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
              if (i < nlive) *d = v[i];
              d += p.d_c;
            }
          }
        }
        // ---- per-channel statistics: transposing warp reduce into this warp's running sums (shared memory)
        if (p.stats != nullptr) {
          transpose_reduce32(v);
          transpose_reduce32(u);
          if (lane < CH) {
            myAcc[ch0 + lane] += v[0];
            myAcc[128 + ch0 + lane] += u[0];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
    }
    if (p.stats != nullptr) flush_stats();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, kTmemCols);
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
};

constexpr int WK = 64;                       // pixels per k-block
constexpr int kWStageA = 128 * WK * 2;       // 16 KiB

template <int BN>
__global__ void __launch_bounds__(kWThreads) wgrad_tc_kernel(const WgradParams p) {
  constexpr int kBStage = BN * WK * 2;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
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
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 128); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  if (p.pre_scale != nullptr)
    for (int i = threadIdx.x; i < Cs; i += kWThreads) { sScale[i] = __ldg(p.pre_scale + i); sShift[i] = __ldg(p.pre_shift + i); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // producers: thread = (pixel px of the k-block, half hf); 8 A chunks + BN/16 B chunks each
    const int px = threadIdx.x & 63, hf = threadIdx.x >> 6;
    const bool vec = (Cs % 8 == 0) && p.s_c == 1;
    const int Nn = p.plan.Nn;
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
      mbar_wait(&empty[s], ((kb / NS) & 1) ^ 1);
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
            if (p.dy_bf16) out = qb0[j];
            else out = make_uint4(pack_bf16(__uint_as_float(qb0[j].x), __uint_as_float(qb0[j].y)), pack_bf16(__uint_as_float(qb0[j].z), __uint_as_float(qb0[j].w)),
                                  pack_bf16(__uint_as_float(qb1[j].x), __uint_as_float(qb1[j].y)), pack_bf16(__uint_as_float(qb1[j].z), __uint_as_float(qb1[j].w)));
          } else if (mvalid && n < Nn && !(dyvec && n + 8 <= Nn)) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (n + i < Nn) ? ld_elem(p.dy, dy_off + (n + i) * p.y_c, p.dy_bf16) : 0.f;
            out = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          }
          *reinterpret_cast<uint4*>(b_st + grp * (WK * 16) + px * 16) = out;
        }
      }
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }

    // ---- epilogue: scatter-add the tile into the reference weight layout
    mbar_wait(tmem_full, 0);
    tc_fence_after();
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
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc(kFmtBF16, 128, BN, 1, 1);  // both operands MN-major
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % NS;
        mbar_wait(&full[s], (kb / NS) & 1);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + s * kWStageA), b_base = smem_u32(sB + s * kBStage);
#pragma unroll
        for (int k4 = 0; k4 < WK / 16; ++k4) {
          // MN-major interleave: 8-pixel k-groups 128 B apart (LBO), 8-channel mn-groups WK*16 B apart (SBO)
          const uint64_t ad = smem_desc(a_base + k4 * 256, 128, WK * 16, kLayoutNone);
          const uint64_t bd = smem_desc(b_base + k4 * 256, 128, WK * 16, kLayoutNone);
          umma_f16(tmem_base, ad, bd, idesc, (kb | k4) != 0 ? 1u : 0u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------
// weight packing: fp32 reference layout -> bf16 [class][n_pad][Kp], zero padded
// ---------------------------------------------------------------------------
__global__ void pack_weight_kernel(const Plan plan, const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
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

inline size_t conv_smem_bytes(int BN, int Cs) {
  return (size_t)NS * kAStage + (size_t)NS * BN * BK * 2 + 128 /*barriers + tmem slot*/ +
         cvplan::kMaxClasses * kTabPerCls * 16 + cvplan::kMaxClasses * cvplan::kMaxTaps * 4 + 4 * 2 * 128 * 4 +
         2 * (size_t)Cs * 4 + 1024;
}

template <int BN, bool SRC_BF16, int EPI>
int launch(const TmapPack& tm, const GemmParams& p, cudaStream_t st) {
  static bool attr_done = false;
  static int num_sms = 148;
  auto kern = conv_tc_kernel<BN, SRC_BF16, EPI>;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem_bytes(BN, kMaxPreC));
    if (e != cudaSuccess) return (int)e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    attr_done = true;
  }
  const size_t smem = conv_smem_bytes(BN, p.plan.Cs);
  const int grid = std::min(p.n_tiles, num_sms);
  kern<<<grid, kThreads, smem, st>>>(tm, p);
  CV_LAUNCH_CHECK();
  return 0;
}
template <int BN>
int launch_bn(const TmapPack& tm, const GemmParams& p, cudaStream_t st) {
  if (p.epi == CLEARVAE_EPI_BIAS_STATS)
    return p.src_bf16 ? launch<BN, true, CLEARVAE_EPI_BIAS_STATS>(tm, p, st) : launch<BN, false, CLEARVAE_EPI_BIAS_STATS>(tm, p, st);
  return p.src_bf16 ? launch<BN, true, CLEARVAE_EPI_MASK_STATS>(tm, p, st) : launch<BN, false, CLEARVAE_EPI_MASK_STATS>(tm, p, st);
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
  wgrad_tc_kernel<BN><<<grid, kWThreads, smem, st>>>(p);
  CV_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

int clearvae_conv_wgrad(const clearvae_conv_geom* g, int64_t batch, const clearvae_tensor4* src, const float* pre_scale,
                        const float* pre_shift, int32_t pre_relu, const clearvae_tensor4* dy, float* dweight, void* stream) {
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
  int max_k = 0;
  long long max_m = 0;
  for (int i = 0; i < p.plan.n_classes; ++i) {
    max_k = std::max(max_k, p.plan.cls[i].ntaps * p.plan.Cs);
    max_m = std::max(max_m, (long long)batch * p.plan.cls[i].Hd * p.plan.cls[i].Wd);
  }
  const int tiles = ((max_k + 127) / 128) * ((n_pad + BN - 1) / BN) * p.plan.n_classes;
  // split the pixel reduction so the grid is ~2 waves of 148 SMs, each CTA keeping >= 4 k-blocks
  long long splits = std::max<long long>(1, (2 * 148 + tiles - 1) / tiles);
  splits = std::min<long long>(splits, std::max<long long>(1, max_m / (WK * 4)));
  p.splits = (int)splits;
  dim3 grid((unsigned)((max_k + 127) / 128), (unsigned)((n_pad + BN - 1) / BN), (unsigned)(p.plan.n_classes * p.splits));
  cudaStream_t st = (cudaStream_t)stream;
  switch (BN) {
    case 16: return launch_wgrad<16>(p, grid, st);
    case 32: return launch_wgrad<32>(p, grid, st);
    case 64: return launch_wgrad<64>(p, grid, st);
    default: return launch_wgrad<128>(p, grid, st);
  }
}

size_t clearvae_conv_packed_weight_bytes(const clearvae_conv_geom* g, int32_t role) {
  Plan plan;
  if (!g || !cvplan::make_plan(*g, role, BK, &plan)) return 0;
  return (size_t)cvplan::packed_weight_elems(plan) * 2;
}

int clearvae_conv_pack_weight(const clearvae_conv_geom* g, int32_t role, const float* weight, void* packed, void* stream) {
  if (!g || !weight || !packed) return CLEARVAE_EINVAL;
  Plan plan;
  if (!cvplan::make_plan(*g, role, BK, &plan)) return CLEARVAE_EUNSUPPORTED;
  const int n_pad = (plan.Nn + 15) / 16 * 16;
  long long mx = 0;
  for (int i = 0; i < plan.n_classes; ++i) mx = std::max(mx, (long long)n_pad * plan.cls[i].Kp);
  dim3 grid((unsigned)std::min<long long>((mx + 255) / 256, 148 * 8), (unsigned)plan.n_classes);
  pack_weight_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(plan, weight, reinterpret_cast<__nv_bfloat16*>(packed));
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
  if (!cvplan::make_plan(*g, role, BK, &p.plan)) return CLEARVAE_EUNSUPPORTED;
  if (pre_scale != nullptr && p.plan.Cs > kMaxPreC) return CLEARVAE_EUNSUPPORTED;
  EncodeTiledFn enc = get_encode();
  if (!enc) return CLEARVAE_EUNSUPPORTED;
  const int BN = pick_bn(p.plan.Nn);
  const int n_pad = (p.plan.Nn + 15) / 16 * 16;
  TmapPack tm{};
  long long max_m = 0;
  for (int i = 0; i < p.plan.n_classes; ++i) {
    const Cls& c = p.plan.cls[i];
    cuuint64_t dims[2] = {(cuuint64_t)c.Kp, (cuuint64_t)n_pad};
    cuuint64_t strides[1] = {(cuuint64_t)c.Kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    void* base = (void*)(reinterpret_cast<const char*>(packed_weight) + (size_t)c.w_off * 2);
    CUresult r = enc(&tm.t[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return CLEARVAE_EINVAL;
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
  // work list of the persistent kernel: class-major, then m-tile, split, n-tile (n fastest: neighbouring CTAs
  // gather the same activation rows at the same time and share them through L2)
  p.n_ntiles = (n_pad + BN - 1) / BN;
  int total = 0;
  for (int i = 0; i < p.plan.n_classes; ++i) {
    p.tile_start[i] = total;
    const long long mc = (long long)batch * p.plan.cls[i].Hd * p.plan.cls[i].Wd;
    total += (int)((mc + BM - 1) / BM) * p.splits * p.n_ntiles;
  }
  p.n_tiles = total;
  switch (BN) {
    case 16: return launch_bn<16>(tm, p, st);
    case 32: return launch_bn<32>(tm, p, st);
    case 64: return launch_bn<64>(tm, p, st);
    default: return launch_bn<128>(tm, p, st);
  }
}

}  // extern "C"
