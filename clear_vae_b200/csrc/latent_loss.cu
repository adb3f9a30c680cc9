// Fused latent-head loss block, exact-fp32 (FFMA) path.
//
// Replaces, per term (content / style), the reference's
//   vae.py:56-60      reparameterisation  z = mu + eps * exp(logvar / 2)
//   losses.py:48-49   Gaussian KL
//   losses.py:98-137  pair mask, pairwise similarity, snn_loss, finite-row mean
// and their autograd backward, with two launches per training step.
//
// Tiling: one warp owns RM rows of the local shard; the 32 lanes sweep the
// (global) columns, which are staged TN at a time in shared memory in [d][j]
// order (conflict-free, one thread normalises one column).  The B x B matrix is
// never materialised: each lane keeps a running masked sum-of-exponentials and
// the row result is a warp-shuffle reduction.
//
// cosine + moderate temperature ("FAST"): |s| <= 1, so every exponent is taken
// relative to the constant shift 1/tau; sums from different lanes / CTAs / ranks
// then add without rescaling and exp(.) is symmetric in (i, j), which lets the
// backward fold the column-side gradient G^T into the row pass:
//     dL/dn_i = w * sum_j e_ij (c_i + c_j - [pos_ij](q_i + q_j)) n_j
// (c = [finite]/sum_all, q = [finite]/sum_pos) — no transposed pass, no atomics
// and under data parallelism no reduce-scatter of dZ.
// Other similarities / tiny temperatures use running maxima (two accumulators
// per set) like the reference's logsumexp (losses.py:87-95).
#include <cuda_bf16.h>
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace {

constexpr int kTN = 256;       // columns per shared-memory tile == threads per CTA
constexpr int kWarps = 8;
constexpr float kCosEps = 1e-8f;  // F.cosine_similarity eps (losses.py:55)

struct TermF {
  const float *mu, *lv, *eps, *mu_cols, *lv_cols;
  float *z, *stats;
  int snn, ps;
  float* aux;   // [B] SupCon only: n_k (supcon_in) / positive count (supcon_out)
};
struct FwdParams {
  TermF t[2];
  const long long *lab_r, *lab_c;
  long long B, Bg, row_off;
  int D, z_stride, finalize, max_ctas;
  float inv_tau;
  float* scalars;
  unsigned* ticket;
  float* kl_partial;  // [2][max_ctas]
  // column split (small local batches, FFMA kernels): blockIdx.z sweeps columns [z*cps, (z+1)*cps); per-row partial
  // (max_all, sum_all, max_pos, sum_pos) go to part[term][z][B][4] and the last CTA of the grid merges them in z order
  int splits;
  long long cps;
  float* part;
  int loss;   // CLEARVAE_LOSS_*
};
struct TermB {
  const float *mu, *lv, *eps, *mu_cols, *stats_all, *dz;
  float *dmu, *dlv;
  int snn, ps;
  const float* lv_cols;
  const float* aux_all;   // [Bg] SupCon only (see TermF::aux), all global rows
};
struct BwdParams {
  TermB t[2];
  const long long *lab_r, *lab_c;
  long long B, Bg, row_off;
  int D, z_stride;
  float inv_tau;
  const float *scalars, *gscal;
  // column split: partial (dn[DP], csum) per row -> part[term][z][B][DP+1]; the last of the `splits` CTAs of a row block
  // (ticket per (term, blockIdx.x), self-resetting) merges them in z order and runs the row epilogue
  int splits;
  long long cps;
  float* part;
  unsigned* tickets;
  int loss;   // CLEARVAE_LOSS_*
};

enum { SIM_COS = CLEARVAE_SIM_COSINE, SIM_L2 = CLEARVAE_SIM_L2, SIM_ML2 = CLEARVAE_SIM_MODIFIED_L2, SIM_JEF = CLEARVAE_SIM_JEFFREY,
       SIM_MAH = CLEARVAE_SIM_MAHALANOBIS };
// similarities that also read logvar (losses.py:62-84) stage up to three derived per-column arrays next to mu:
//   modified_l2 : u = exp(-lv/2)                      s_ij = -sum_d (mu_j - mu_i)^2 u_i u_j
//   mahalanobis : v = exp(lv)                         s_ij = -sum_d (mu_j - mu_i)^2 / ((v_i + v_j) / 2)
//   jeffrey     : v, w = 1/v, h = 1/(v + 1e-8)        s_ij = -1/4 [ -2D + sum_d ((mu_j - mu_i)^2 (w_i + w_j) + v_j h_i + v_i h_j) ]
// all symmetric in (i, j), so the backward still folds G^T into the row pass (coefficient G_ij + G_ji).
template <int SIM> struct SimLv { static constexpr int N = SIM == SIM_JEF ? 3 : (SIM == SIM_ML2 || SIM == SIM_MAH) ? 1 : 0; };
template <int SIM>
__device__ __forceinline__ void lv_derive(float lv, float (&o)[3]) {
  if (SIM == SIM_ML2) { o[0] = expf(-0.5f * lv); o[1] = o[2] = 0.f; }
  else { const float v = expf(lv); o[0] = v; o[1] = 1.f / v; o[2] = 1.f / (v + 1e-8f); }
}
// one dimension of one pair: contribution to -s (acc), and the derivatives of s w.r.t. the ROW's mu_d and logvar_d
template <int SIM, bool GRAD>
__device__ __forceinline__ void pair_dim(float a, const float (&ri)[3], float x, float c0, float c1, float c2, float& acc, float& ds_da,
                                         float& ds_dlv) {
  const float df = x - a;
  if (SIM == SIM_ML2) {
    const float t = ri[0] * c0;
    acc = fmaf(df * df, t, acc);
    if (GRAD) { ds_da = 2.f * df * t; ds_dlv = 0.5f * df * df * t; }
  } else if (SIM == SIM_MAH) {
    const float iv = 2.f / (ri[0] + c0);          // 1 / V, V = (v_i + v_j) / 2
    acc = fmaf(df * df, iv, acc);
    if (GRAD) { ds_da = 2.f * df * iv; ds_dlv = 0.5f * df * df * iv * iv * ri[0]; }
  } else {
    acc += df * df * (ri[1] + c1) + c0 * ri[2] + ri[0] * c2;
    if (GRAD) {
      ds_da = 0.5f * df * (ri[1] + c1);
      ds_dlv = -0.25f * (-df * df * ri[1] - c0 * ri[2] * ri[2] * ri[0] + ri[0] * c2);
    }
  }
}
template <int SIM>
__device__ __forceinline__ float sim_from_acc(float acc, int D) {
  return SIM == SIM_JEF ? -0.25f * (acc - 2.f * (float)D) : -acc;
}

// ---------------------------------------------------------------------------
// column tile: thread `tid` stages column j0 + tid (normalised for cosine)
// ---------------------------------------------------------------------------
template <int DP, int SIM>
__device__ __forceinline__ void stage_column(const float* __restrict__ cols, const long long* __restrict__ lab,
                                             long long j, long long Bg, int D, float* sN, long long* sL,
                                             const float* __restrict__ lv_cols = nullptr, float* sX = nullptr) {
  float v[DP];
  const int tid = threadIdx.x;
  if (SimLv<SIM>::N > 0) {   // derived logvar arrays [N][DP][kTN] of the logvar-aware similarities
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      float o[3] = {0.f, 0.f, 0.f};
      if (j < Bg && d < D) lv_derive<SIM>(__ldg(lv_cols + j * (long long)D + d), o);
#pragma unroll
      for (int k = 0; k < SimLv<SIM>::N; ++k) sX[(k * DP + d) * kTN + tid] = o[k];
    }
  }
  if (j < Bg) {
    const float* src = cols + j * (long long)D;
    if ((D & 3) == 0) {
#pragma unroll
      for (int d = 0; d < DP; d += 4) {
        if (d < D) {
          float4 q = __ldg(reinterpret_cast<const float4*>(src + d));
          v[d] = q.x; v[d + 1] = q.y; v[d + 2] = q.z; v[d + 3] = q.w;
        } else {
          v[d] = v[d + 1] = v[d + 2] = v[d + 3] = 0.f;
        }
      }
    } else {
#pragma unroll
      for (int d = 0; d < DP; ++d) v[d] = d < D ? __ldg(src + d) : 0.f;
    }
    if (SIM == SIM_COS) {
      float ss = 0.f;
#pragma unroll
      for (int d = 0; d < DP; ++d) ss = fmaf(v[d], v[d], ss);
      const float inv = 1.f / fmaxf(sqrtf(ss), kCosEps);
#pragma unroll
      for (int d = 0; d < DP; ++d) v[d] *= inv;
    }
    sL[tid] = lab[j];
  } else {
#pragma unroll
    for (int d = 0; d < DP; ++d) v[d] = 0.f;
    sL[tid] = 0;
  }
#pragma unroll
  for (int d = 0; d < DP; ++d) sN[d * kTN + tid] = v[d];
}

template <int DP, int RM, int SIM>
__device__ __forceinline__ void load_rows(const float* __restrict__ mu, const long long* __restrict__ lab,
                                          long long row0, long long B, int D, float (&row)[RM][DP],
                                          float (&inv_norm)[RM], long long (&rl)[RM]) {
#pragma unroll
  for (int r = 0; r < RM; ++r) {
    const long long i = row0 + r;
    float ss = 0.f;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      row[r][d] = (i < B && d < D) ? __ldg(mu + i * (long long)D + d) : 0.f;
      ss = fmaf(row[r][d], row[r][d], ss);
    }
    rl[r] = i < B ? lab[i] : 0;
    inv_norm[r] = 1.f;
    if (SIM == SIM_COS) {
      inv_norm[r] = 1.f / fmaxf(sqrtf(ss), kCosEps);
#pragma unroll
      for (int d = 0; d < DP; ++d) row[r][d] *= inv_norm[r];
    }
  }
}

template <int DP, int SIM>
__device__ __forceinline__ float pair_sim(const float (&a)[DP], const float (&x)[DP]) {
  float acc = 0.f;
  if (SIM == SIM_COS) {
#pragma unroll
    for (int d = 0; d < DP; ++d) acc = fmaf(a[d], x[d], acc);
  } else {
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      const float df = x[d] - a[d];
      acc = fmaf(-df, df, acc);
    }
  }
  return acc;
}

// finite-row mean of one term from row stats (losses.py:125-126); thread 0 writes.
__device__ void finalize_term(const float* stats, long long Bg, int term, float* scalars, float* sred) {
  float s = 0.f, c = 0.f;
  for (long long i = threadIdx.x; i < Bg; i += blockDim.x) {
    const float a = __ldcg(stats + 2 * i), b = __ldcg(stats + 2 * i + 1);
    const float l = a - b;
    if (isfinite(l)) { s += l; c += 1.f; }
  }
  s = cv::block_sum<kTN>(s, sred);
  c = cv::block_sum<kTN>(c, sred);
  if (threadIdx.x == 0) {
    scalars[CLEARVAE_S_SUM0 + term] = s;
    scalars[CLEARVAE_S_CNT0 + term] = c;
    scalars[CLEARVAE_S_LOSS0 + term] = s / c;  // 0/0 = nan, like mean of an empty tensor
  }
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
template <int DP, int RM, int SIM, bool FAST>
__global__ void __launch_bounds__(kTN) snn_fwd_kernel(const FwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sN = reinterpret_cast<float*>(smem_raw);                 // [DP][kTN]
  long long* sL = reinterpret_cast<long long*>(sN + DP * kTN);    // [kTN]
  float* sRed = reinterpret_cast<float*>(sL + kTN);               // [kWarps + 1]
  const int term = blockIdx.y;
  const TermF& t = p.t[term];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)blockIdx.x * (kWarps * RM) + warp * RM;
  const int D = p.D;

  // ---- reparameterisation + KL for the rows of this warp (lanes over d)
  if (t.lv != nullptr && blockIdx.z == 0) {
    float kl = 0.f;
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      const long long i = row0 + r;
      if (i < p.B) {
        for (int d = lane; d < D; d += 32) {
          const float m = t.mu[i * D + d], lv = t.lv[i * D + d];
          kl += 1.f + lv - m * m - expf(lv);
          if (t.z != nullptr && t.eps != nullptr) t.z[i * p.z_stride + d] = fmaf(t.eps[i * D + d], expf(0.5f * lv), m);
        }
      }
    }
    kl = cv::block_sum<kTN>(kl, sRed);
    if (threadIdx.x == 0) p.kl_partial[term * p.max_ctas + blockIdx.x] = kl;
  }

  if (t.snn) {
    float row[RM][DP], inv_norm[RM];
    long long rl[RM];
    load_rows<DP, RM, SIM>(t.mu, p.lab_r, row0, p.B, D, row, inv_norm, rl);
    constexpr int NLV = SimLv<SIM>::N;
    float rlv[RM][NLV > 0 ? DP : 1][3];
    if constexpr (NLV > 0) {
#pragma unroll
      for (int r = 0; r < RM; ++r)
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          const long long i = row0 + r;
          rlv[r][d][0] = rlv[r][d][1] = rlv[r][d][2] = 0.f;
          if (i < p.B && d < D) lv_derive<SIM>(__ldg(t.lv + i * (long long)D + d), rlv[r][d]);
        }
    }
    float sa[RM], sp[RM], ma[RM], mp[RM];
    // SupCon row losses (losses.py:140-170) share the sweep; they additionally need the positive count and, for
    // supcon_out, the sum of the raw positive similarities.  The row statistics keep their (a, b) form with
    //   supcon_in : b = lse_pos - log(n_k),  n_k = #positives (-1 under ps: the reference counts the undiagonalised mask)
    //   supcon_out: b = (sum of positive similarities) / #positives        (0/0 = nan: the row is selected out)
    // so that the row loss is a - b and the finite-row mean of `finalize_term` applies unchanged.
    float pc[RM], ps_sum[RM];
#pragma unroll
    for (int r = 0; r < RM; ++r) { sa[r] = sp[r] = 0.f; ma[r] = mp[r] = -INFINITY; pc[r] = ps_sum[r] = 0.f; }
    const bool supcon = p.loss != CLEARVAE_LOSS_SNN;
    const float k2 = p.inv_tau * CV_LOG2E;
    const float* cols = t.mu_cols ? t.mu_cols : t.mu;
    const float* lvc = t.lv_cols ? t.lv_cols : t.lv;
    float* sX = sRed + kWarps + 1;   // [NLV][DP][kTN]
    const long long jbeg = p.splits > 1 ? (long long)blockIdx.z * p.cps : 0;
    const long long jend = p.splits > 1 ? min(p.Bg, jbeg + p.cps) : p.Bg;
    for (long long j0 = jbeg; j0 < jend; j0 += kTN) {
      __syncthreads();
      stage_column<DP, SIM>(cols, p.lab_c, j0 + threadIdx.x, p.Bg, D, sN, sL, lvc, sX);
      __syncthreads();
      for (int jj = lane; jj < kTN; jj += 32) {
        const long long j = j0 + jj;
        if (j >= jend) break;
        float x[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) x[d] = sN[d * kTN + jj];
        const long long lab = sL[jj];
#pragma unroll
        for (int r = 0; r < RM; ++r) {
          float s;
          if constexpr (NLV > 0) {
            float acc = 0.f, u0, u1;
#pragma unroll
            for (int d = 0; d < DP; ++d)
              if (d < D)
                pair_dim<SIM, false>(row[r][d], rlv[r][d], x[d], sX[d * kTN + jj], NLV > 1 ? sX[(DP + d) * kTN + jj] : 0.f,
                                     NLV > 2 ? sX[(2 * DP + d) * kTN + jj] : 0.f, acc, u0, u1);
            s = sim_from_acc<SIM>(acc, D);
          } else {
            s = pair_sim<DP, SIM>(row[r], x);
          }
          const bool cand = (j != p.row_off + row0 + r);
          const bool pos = cand && ((lab == rl[r]) != (t.ps != 0));
          if (supcon) {
            pc[r] += pos ? 1.f : 0.f;
            ps_sum[r] += pos ? s : 0.f;
          }
          if (FAST) {
            const float e = exp2f(fmaf(s, k2, -k2));
            sa[r] += cand ? e : 0.f;
            sp[r] += pos ? e : 0.f;
          } else {
            const float xs = s * p.inv_tau;
            if (cand) cv::lse_push(ma[r], sa[r], xs);
            else if (p.loss == CLEARVAE_LOSS_SUPCON_OUT) cv::lse_push(ma[r], sa[r], -999.f * p.inv_tau);  // losses.py:158
            if (pos) cv::lse_push(mp[r], sp[r], xs);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      float oa, op;
      if (p.splits > 1) {   // partial over this CTA's column range; merged by the last CTA below
        if (FAST) {
          sa[r] = cv::warp_sum(sa[r]);
          sp[r] = cv::warp_sum(sp[r]);
        } else {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            cv::lse_merge(ma[r], sa[r], __shfl_xor_sync(0xffffffffu, ma[r], o), __shfl_xor_sync(0xffffffffu, sa[r], o));
            cv::lse_merge(mp[r], sp[r], __shfl_xor_sync(0xffffffffu, mp[r], o), __shfl_xor_sync(0xffffffffu, sp[r], o));
          }
        }
        const long long i = row0 + r;
        if (lane == 0 && i < p.B)
          reinterpret_cast<float4*>(p.part)[((long long)term * p.splits + blockIdx.z) * p.B + i] = make_float4(ma[r], sa[r], mp[r], sp[r]);
        continue;
      }
      if (FAST) {
        oa = logf(cv::warp_sum(sa[r]));
        op = logf(cv::warp_sum(sp[r]));
      } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          cv::lse_merge(ma[r], sa[r], __shfl_xor_sync(0xffffffffu, ma[r], o), __shfl_xor_sync(0xffffffffu, sa[r], o));
          cv::lse_merge(mp[r], sp[r], __shfl_xor_sync(0xffffffffu, mp[r], o), __shfl_xor_sync(0xffffffffu, sp[r], o));
        }
        oa = ma[r] + logf(sa[r]);  // (-inf) + log(0) = -inf
        op = mp[r] + logf(sp[r]);
      }
      const long long i = row0 + r;
      if (supcon) {
        const float cnt = cv::warp_sum(pc[r]);
        const float ssum = cv::warp_sum(ps_sum[r]);
        float aux;
        if (p.loss == CLEARVAE_LOSS_SUPCON_IN) {
          aux = cnt - (t.ps != 0 ? 1.f : 0.f);
          op = op - logf(aux);           // log(0) = -inf, log(<0) = nan: non-finite rows drop out like the reference's
        } else {
          aux = cnt;
          op = ssum / cnt - (FAST ? p.inv_tau : 0.f);   // FAST: `oa` is relative to the shared shift 1/tau, keep a - b exact
        }
        if (lane == 0 && i < p.B && t.aux != nullptr) t.aux[i] = aux;
      }
      if (lane == 0 && i < p.B) {
        t.stats[2 * i] = oa;
        t.stats[2 * i + 1] = op;
      }
    }
  }

  // ---- last CTA: deterministic reduction of the per-CTA partials
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned tk = atomicAdd(p.ticket, 1u);
    sRed[kWarps] = (tk == gridDim.x * gridDim.y * gridDim.z - 1) ? 1.f : 0.f;
  }
  __syncthreads();
  if (sRed[kWarps] != 0.f) {
    __threadfence();
    if (p.splits > 1) {   // merge the column-split partials in z order -> row statistics
      for (int tt = 0; tt < (int)gridDim.y; ++tt) {
        if (!p.t[tt].snn) continue;
        for (long long i = threadIdx.x; i < p.B; i += kTN) {
          float ma = -INFINITY, sa = 0.f, mp = -INFINITY, sp = 0.f;
          for (int z = 0; z < p.splits; ++z) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(p.part) + ((long long)tt * p.splits + z) * p.B + i);
            if (FAST) { sa += v.y; sp += v.w; }
            else { cv::lse_merge(ma, sa, v.x, v.y); cv::lse_merge(mp, sp, v.z, v.w); }
          }
          p.t[tt].stats[2 * i] = FAST ? logf(sa) : ma + logf(sa);
          p.t[tt].stats[2 * i + 1] = FAST ? logf(sp) : mp + logf(sp);
        }
      }
      __threadfence();
      __syncthreads();
    }
    for (int tt = 0; tt < (int)gridDim.y; ++tt) {
      if (p.t[tt].lv != nullptr) {
        float s = 0.f;
        for (int c = threadIdx.x; c < (int)gridDim.x; c += kTN) s += __ldcg(p.kl_partial + tt * p.max_ctas + c);
        s = cv::block_sum<kTN>(s, sRed);
        if (threadIdx.x == 0) p.scalars[CLEARVAE_S_KL0 + tt] = -0.5f * s / (float)p.B;
      }
      if (p.finalize && p.t[tt].snn) finalize_term(p.t[tt].stats, p.Bg, tt, p.scalars, sRed);
    }
    if (threadIdx.x == 0) *p.ticket = 0u;
  }
}

__global__ void __launch_bounds__(kTN) snn_finalize_kernel(const float* stats, long long Bg, int term, float* scalars) {
  __shared__ float sred[kWarps];
  finalize_term(stats, Bg, term, scalars, sred);
}


// ---------------------------------------------------------------------------
// forward, tensor-core path (cosine, shared-shift): large batches
//
//   S = N_rows * N_cols^T on tcgen05 with a 3xTF32 split (hi*hi + lo*hi + hi*lo, fp32-grade:
//   the temperature multiplies similarity error by 1/tau before the exponential, so a single
//   TF32 / bf16 product would miss the 1e-5 gate), accumulators in TMEM, double buffered so
//   the masked exp / running-sum epilogue of column tile j overlaps the MMA of tile j+1.
//   Per pair the epilogue spends 1 FFMA + 1 MUFU.EX2 + 2 FADD (+ label compare); the kernel is
//   bound by the MUFU pipe (16 ex2/clk/SM), which is the roofline bench.py reports against.
//   Warp roles: 0-3 column-tile producers (normalise + split + UMMA K-major interleave layout),
//   4 MMA issuer, 5-12 epilogue (each warp: 32 TMEM lanes x half of the tile's columns).
// ---------------------------------------------------------------------------
template <int DP> struct TcCfg {
  static constexpr int BN = DP <= 16 ? 256 : 128;     // columns per tile
  static constexpr int KT = 3 * DP;                   // hi | lo | hi  (A)   x   hi | hi | lo  (B)
  static constexpr int A_BYTES = 128 * KT * 4;
  static constexpr int B_BYTES = BN * KT * 4;
  // column-tile ring (operand + labels), decoupled from the two TMEM accumulators so the producers run ahead of the epilogue
  static constexpr int NSB = DP <= 8 ? 4 : 3;
  static constexpr int SMEM = A_BYTES + NSB * B_BYTES + NSB * BN * 8 /*labels lo/hi*/ + 2048 /*barriers, flags, partials*/ + 1024;
};
constexpr int kTcThreads = 13 * 32;

// round-to-nearest, ties away from zero, to the 10-bit tf32 significand.  For finite inputs this is bit-identical to
// cvt.rna.tf32.f32 (sign-magnitude: add half an ulp of the kept field, clear the dropped 13 bits), which sm_100a expands into
// four instructions per value (add, mask, NaN/Inf test, select); every caller passes normalised, finite values.
__device__ __forceinline__ float tf32_rna_finite(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// tf32 head by truncation (one LOP3; cvt.rna.tf32.f32 is a four-instruction sequence on sm_100a).  x - tf32_trunc(x) is exact in
// fp32 and below 2^-10 |x|, so a (head, tail) pair read by the tensor core as two tf32 operands carries x to 2^-20 relative.
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ float bf16_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffff0000u); }
// {lo -> bits [0,16), hi -> bits [16,32)} as round-to-nearest bf16
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// stage one normalised row / column vector as [hi | second | third] K-major interleave chunks
template <int DP>
__device__ __forceinline__ void load_vec(const float* __restrict__ src, bool valid, int D, float (&v)[DP]) {
  if (DP % 4 == 0 && D == DP) {   // contiguous, 16-byte aligned rows: vector loads
#pragma unroll
    for (int c = 0; c < DP / 4; ++c) {
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) q = __ldg(reinterpret_cast<const float4*>(src) + c);
      v[4 * c] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int d = 0; d < DP; ++d) v[d] = (valid && d < D) ? __ldg(src + d) : 0.f;
  }
}
template <int DP, bool TRUNC = false>
__device__ __forceinline__ void stage_split_vals(const float (&v)[DP], unsigned char* base, int rows, int r, bool a_side,
                                                 unsigned char* kmajor2 = nullptr, unsigned char* kmajor2_bf16 = nullptr,
                                                 int nb2 = 0) {
  float ss = 0.f;
#pragma unroll
  for (int d = 0; d < DP; ++d) ss = fmaf(v[d], v[d], ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), kCosEps);
  float hi[DP], lo[DP];
  if (TRUNC) {   // head by truncation, exact remainder (the tensor core drops the remainder's low bits itself)
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      const float n = v[d] * inv;
      hi[d] = tf32_trunc(n);
      lo[d] = n - hi[d];
    }
  } else if (isfinite(ss)) {   // every n is finite: the two-instruction rounding
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      const float n = v[d] * inv;
      hi[d] = tf32_rna_finite(n);
      lo[d] = tf32_rna_finite(n - hi[d]);
    }
  } else {                     // NaN / Inf rows keep the instruction's semantics
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      const float n = v[d] * inv;
      hi[d] = tf32_rna(n);
      lo[d] = tf32_rna(n - hi[d]);
    }
  }
  constexpr int NC = DP / 4;  // 16-byte chunks per D-block
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float4 h4 = make_float4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
    const float4 l4 = make_float4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
    *reinterpret_cast<float4*>(base + (size_t)(c) * rows * 16 + r * 16) = h4;
    *reinterpret_cast<float4*>(base + (size_t)(NC + c) * rows * 16 + r * 16) = a_side ? l4 : h4;
    *reinterpret_cast<float4*>(base + (size_t)(2 * NC + c) * rows * 16 + r * 16) = a_side ? h4 : l4;
  }
  if (kmajor2 != nullptr) {
    // second-GEMM operand [n = (hi | lo) x d][k = column r]: K-major interleave, 2*DP rows
    float* dst = reinterpret_cast<float*>(kmajor2 + (size_t)(r >> 2) * (2 * DP * 16) + (r & 3) * 4);
#pragma unroll
    for (int d = 0; d < DP; ++d) { dst[d * 4] = hi[d]; dst[(DP + d) * 4] = lo[d]; }
  }
  if (kmajor2_bf16 != nullptr) {
    // second-GEMM operand in bf16: [n = part x d (nb2 rows)][k = column r], K-major interleave (16-byte chunk = 8 columns);
    // three parts n0 + n1 + n2 carry the normalised vector to 24 bits
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(kmajor2_bf16 + (size_t)(r >> 3) * (nb2 * 16) + (r & 7) * 2);
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      const float n = v[d] * inv;
      const __nv_bfloat16 n0 = __float2bfloat16_rn(n);
      const float r1 = n - __bfloat162float(n0);
      const __nv_bfloat16 n1 = __float2bfloat16_rn(r1);
      const __nv_bfloat16 n2 = __float2bfloat16_rn(r1 - __bfloat162float(n1));
      dst[d * 8] = n0; dst[(DP + d) * 8] = n1; dst[(2 * DP + d) * 8] = n2;
    }
  }
}
template <int DP>
__device__ __forceinline__ void stage_split(const float* __restrict__ src, bool valid, int D, unsigned char* base, int rows,
                                            int r, bool a_side, unsigned char* kmajor2 = nullptr) {
  float v[DP];
  load_vec<DP>(src, valid, D, v);
  stage_split_vals<DP>(v, base, rows, r, a_side, kmajor2);
}

template <int DP>
__global__ void __launch_bounds__(kTcThreads, 1) snn_fwd_tc_kernel(const FwdParams p) {
  using C = TcCfg<DP>;
  constexpr int BN = C::BN;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (sm100::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;
  unsigned char* sB = smem + C::A_BYTES;
  constexpr int NSB = C::NSB;
  int* sLab = reinterpret_cast<int*>(sB + NSB * C::B_BYTES);          // [NSB][2][BN]: stage, (lo, hi), column
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLab + 2 * NSB * BN);
  uint64_t *b_full = bars, *b_empty = bars + NSB, *l_empty = bars + 2 * NSB, *t_full = bars + 3 * NSB, *t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  int* sFlags = reinterpret_cast<int*>(tmem_slot + 1);                 // [NSB] sticky per-stage "a column label's high word differs"
  float* sRed = reinterpret_cast<float*>(sFlags + 4);                  // [kWarps + 1] + 2*128 partial sums
  float* sPart = sRed + 16;

  const int term = blockIdx.y;
  const TermF& t = p.t[term];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m0 = (long long)blockIdx.x * 128;
  const int D = p.D;
  const int ntiles = (int)((p.Bg + BN - 1) / BN);
  const float* cols = t.mu_cols ? t.mu_cols : t.mu;

  // ---- reparameterisation + KL for this CTA's rows (same arithmetic as the FFMA kernel)
  if (t.lv != nullptr) {
    float kl = 0.f;
    if (threadIdx.x < 256) {
      for (int rr = threadIdx.x >> 5; rr < 128; rr += 8) {
        const long long i = m0 + rr;
        if (i < p.B) {
          for (int d = lane; d < D; d += 32) {
            const float m = t.mu[i * D + d], lv = t.lv[i * D + d];
            kl += 1.f + lv - m * m - expf(lv);
            if (t.z != nullptr && t.eps != nullptr) t.z[i * p.z_stride + d] = fmaf(t.eps[i * D + d], expf(0.5f * lv), m);
          }
        }
      }
    }
    kl = cv::warp_sum(kl);
    if (lane == 0) sPart[warp] = kl;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < 13; ++w) s += sPart[w];
      p.kl_partial[term * p.max_ctas + blockIdx.x] = s;
    }
    __syncthreads();
  }

  if (t.snn) {
    using namespace sm100;
    if (threadIdx.x == 0) {
      for (int b = 0; b < NSB; ++b) { mbar_init(&b_full[b], 128); mbar_init(&b_empty[b], 1); mbar_init(&l_empty[b], 8); sFlags[b] = 0; }
      for (int b = 0; b < 2; ++b) { mbar_init(&t_full[b], 1); mbar_init(&t_empty[b], 8); }
      fence_barrier_init();
    }
    if (warp == 4) { tmem_alloc(tmem_slot, 2 * BN); tmem_relinquish(); }
    // rows: A operand [hi | lo | hi], staged once (threads 0..127 = rows)
    const long long hi_ref = (long long)(p.lab_c[0] >> 32);
    if (threadIdx.x < 128) {
      const long long i = m0 + threadIdx.x;
      stage_split<DP>(t.mu + (i < p.B ? i : 0) * (long long)D, i < p.B, D, sA, 128, threadIdx.x, true);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    fence_proxy_async();
    __syncthreads();

    if (warp < 4) {
      // ================= column-tile producers =================
      // BN / 128 columns per thread; the global loads of tile jt + 1 are issued before tile jt is normalised / split /
      // stored, and the shared-memory ring lets this run NSB - 1 tiles ahead of the MMA + epilogue
      constexpr int CPT = BN / 128;
      float nv[CPT][DP];
      long long nlab[CPT];
      auto prefetch = [&](int jt) {
#pragma unroll
        for (int u = 0; u < CPT; ++u) {
          const long long j = (long long)jt * BN + threadIdx.x + u * 128;
          const bool valid = j < p.Bg;
          load_vec<DP>(cols + (valid ? j : 0) * (long long)D, valid, D, nv[u]);
          nlab[u] = valid ? __ldg(p.lab_c + j) : 0;
        }
      };
      prefetch(0);
      for (int jt = 0; jt < ntiles; ++jt) {
        const int b = jt % NSB;
        const uint32_t ph = ((jt / NSB) & 1) ^ 1;
        float cv_[CPT][DP];
        long long clab[CPT];
#pragma unroll
        for (int u = 0; u < CPT; ++u) {
          clab[u] = nlab[u];
#pragma unroll
          for (int d = 0; d < DP; ++d) cv_[u][d] = nv[u][d];
        }
        if (jt + 1 < ntiles) prefetch(jt + 1);
        mbar_wait(&b_empty[b], ph);     // the MMA of tile jt - NSB has read the operand stage
        mbar_wait(&l_empty[b], ph);     // the epilogue of tile jt - NSB has read the label stage
        unsigned char* bs = sB + b * C::B_BYTES;
        int any_diff = 0;
#pragma unroll
        for (int u = 0; u < CPT; ++u) {
          const int cc = threadIdx.x + u * 128;
          const long long j = (long long)jt * BN + cc;
          const bool valid = j < p.Bg;
          stage_split_vals<DP>(cv_[u], bs, BN, cc, false);
          sLab[(b * 2 + 0) * BN + cc] = (int)(clab[u] & 0xffffffffll);
          sLab[(b * 2 + 1) * BN + cc] = (int)(clab[u] >> 32);
          any_diff |= (valid && (clab[u] >> 32) != hi_ref) ? 1 : 0;
        }
        any_diff = __any_sync(0xffffffffu, any_diff);
        if (lane == 0 && any_diff) atomicOr(&sFlags[b], 1);
        fence_proxy_async();
        mbar_arrive(&b_full[b]);
      }
    } else if (warp == 4) {
      // ================= MMA issuer =================
      if (lane == 0) {
        constexpr uint32_t idesc = instr_desc(kFmtTF32, 128, BN, 0, 0);
        for (int jt = 0; jt < ntiles; ++jt) {
          const int b = jt & 1, sb = jt % NSB;
          mbar_wait(&b_full[sb], (jt / NSB) & 1);
          mbar_wait(&t_empty[b], ((jt >> 1) & 1) ^ 1);   // the epilogue has drained this TMEM accumulator
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB + sb * C::B_BYTES);
#pragma unroll
          for (int k8 = 0; k8 < C::KT / 8; ++k8) {
            const uint64_t ad = smem_desc(a_base + k8 * 2 * (128 * 16), 128 * 16, 128, kLayoutNone);
            const uint64_t bd = smem_desc(b_base + k8 * 2 * (BN * 16), BN * 16, 128, kLayoutNone);
            umma_tf32(tmem_base + b * BN, ad, bd, idesc, k8 != 0 ? 1u : 0u);
          }
          umma_commit(&b_empty[sb]);
          umma_commit(&t_full[b]);
        }
      }
    } else {
      // ================= epilogue: masked exp + running sums =================
      const int ew = warp - 5;                 // 0..7
      const int lane_grp = warp & 3;           // TMEM lanes this warp may touch: 32 * (warp % 4)
      const int half = ew >> 2;                // which half of the tile's columns
      const int r = lane_grp * 32 + lane;      // row inside the tile
      const long long i = m0 + r;
      const long long my_lab = i < p.B ? p.lab_r[i] : 0;
      const int my_lo = (int)(my_lab & 0xffffffffll), my_hi = (int)(my_lab >> 32);
      const bool my_hi_odd = (my_lab >> 32) != hi_ref;  // then equality of the low words is not enough
      const long long diag = p.row_off + i;    // global column index of this row's diagonal
      const bool ps = t.ps != 0;
      const float k2 = p.inv_tau * CV_LOG2E;
      float sa0 = 0.f, sp0 = 0.f, sp1 = 0.f, sp2 = 0.f, sp3 = 0.f;
      float2 sa01 = make_float2(0.f, 0.f), sa23 = make_float2(0.f, 0.f);
      const float2 k2v = make_float2(k2, k2), nk2v = make_float2(-k2, -k2);
      constexpr int HALF = BN / 2;
      for (int jt = 0; jt < ntiles; ++jt) {
        const int b = jt & 1;
        mbar_wait(&t_full[b], (jt >> 1) & 1);
        tc_fence_after();
        const long long jbase = (long long)jt * BN + half * HALF;
        const int sb = jt % NSB;   // shared-memory stage of this tile's labels (b = TMEM accumulator)
        const bool edge = (jbase + HALF > p.Bg) || (diag >= jbase && diag < jbase + HALF) || sFlags[sb] || my_hi_odd;
        const int* lab_lo = sLab + (sb * 2 + 0) * BN + half * HALF;
        const int* lab_hi = sLab + (sb * 2 + 1) * BN + half * HALF;
        // software pipeline: the TMEM load of chunk c+1 is in flight while chunk c is exponentiated
        constexpr int NCH = HALF / 32;
        uint32_t raw[2][32];
        const uint32_t tcol = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(b * BN + half * HALF);
        tmem_ld32(tcol, raw[0]);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const int c0 = ch * 32;
          tmem_ld_wait();
          if (ch + 1 < NCH) tmem_ld32(tcol + (uint32_t)(c0 + 32), raw[(ch + 1) & 1]);
          const uint32_t(&rw)[32] = raw[ch & 1];
          if (!edge) {
#pragma unroll
            for (int q = 0; q < 32; q += 4) {
              const int4 l4 = *reinterpret_cast<const int4*>(lab_lo + c0 + q);
              // packed fp32x2 pipe (FFMA2 / FADD2): half the issue slots for the scale and the running sum, which is
              // what keeps this loop under the 8 issue slots per MUFU.EX2 the roofline allows
              const float2 t01 = __ffma2_rn(make_float2(__uint_as_float(rw[q]), __uint_as_float(rw[q + 1])), k2v, nk2v);
              const float2 t23 = __ffma2_rn(make_float2(__uint_as_float(rw[q + 2]), __uint_as_float(rw[q + 3])), k2v, nk2v);
              const float e0 = ex2_approx(t01.x), e1 = ex2_approx(t01.y), e2 = ex2_approx(t23.x), e3 = ex2_approx(t23.y);
              sa01 = __fadd2_rn(sa01, make_float2(e0, e1));
              sa23 = __fadd2_rn(sa23, make_float2(e2, e3));
              sp0 += ((l4.x == my_lo) != ps) ? e0 : 0.f;
              sp1 += ((l4.y == my_lo) != ps) ? e1 : 0.f;
              sp2 += ((l4.z == my_lo) != ps) ? e2 : 0.f;
              sp3 += ((l4.w == my_lo) != ps) ? e3 : 0.f;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              const long long j = jbase + c0 + q;
              const float e = ex2_approx(fmaf(__uint_as_float(rw[q]), k2, -k2));
              const bool cand = (j < p.Bg) && (j != diag);
              const bool same = (lab_lo[c0 + q] == my_lo) && (lab_hi[c0 + q] == my_hi);
              sa0 += cand ? e : 0.f;
              sp0 += (cand && (same != ps)) ? e : 0.f;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&t_empty[b]); mbar_arrive(&l_empty[sb]); }
      }
      // combine the two column halves of each row
      const float sa = ((sa01.x + sa01.y) + (sa23.x + sa23.y)) + sa0, sp = (sp0 + sp1) + (sp2 + sp3);
      if (half == 1) { sPart[r] = sa; sPart[128 + r] = sp; }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 0 && i < p.B) {
        t.stats[2 * i] = logf(sa + sPart[r]);
        t.stats[2 * i + 1] = logf(sp + sPart[128 + r]);
      }
    }
    sm100::tc_fence_before();
    __syncthreads();
    if (warp == 4) sm100::tmem_dealloc(tmem_base, 2 * BN);
  }

  // ---- last CTA: deterministic reduction of the per-CTA partials (same as the FFMA kernel)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned tk = atomicAdd(p.ticket, 1u);
    sRed[kWarps] = (tk == gridDim.x * gridDim.y - 1) ? 1.f : 0.f;
  }
  __syncthreads();
  if (sRed[kWarps] != 0.f && threadIdx.x < kTN) {
    __threadfence();
    for (int tt = 0; tt < (int)gridDim.y; ++tt) {
      if (p.t[tt].lv != nullptr) {
        float s = 0.f;
        for (int c = threadIdx.x; c < (int)gridDim.x; c += kTN) s += __ldcg(p.kl_partial + tt * p.max_ctas + c);
        s = cv::warp_sum(s);
        if (lane == 0) sPart[warp] = s;
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (threadIdx.x == 0) {
          float a = 0.f;
          for (int w = 0; w < kWarps; ++w) a += sPart[w];
          p.scalars[CLEARVAE_S_KL0 + tt] = -0.5f * a / (float)p.B;
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
      if (p.finalize && p.t[tt].snn) {
        const float* stats = p.t[tt].stats;
        float s = 0.f, c = 0.f;
        for (long long ii = threadIdx.x; ii < p.Bg; ii += kTN) {
          const float l = __ldcg(stats + 2 * ii) - __ldcg(stats + 2 * ii + 1);
          if (isfinite(l)) { s += l; c += 1.f; }
        }
        s = cv::warp_sum(s);
        c = cv::warp_sum(c);
        if (lane == 0) { sPart[warp] = s; sPart[16 + warp] = c; }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (threadIdx.x == 0) {
          float a = 0.f, n = 0.f;
          for (int w = 0; w < kWarps; ++w) { a += sPart[w]; n += sPart[16 + w]; }
          p.scalars[CLEARVAE_S_SUM0 + tt] = a;
          p.scalars[CLEARVAE_S_CNT0 + tt] = n;
          p.scalars[CLEARVAE_S_LOSS0 + tt] = a / n;
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
    }
    if (threadIdx.x == 0) *p.ticket = 0u;
  }
}

template <int DP>
int launch_fwd_tc(const FwdParams& p, int n_terms, cudaStream_t st) {
  auto kern = snn_fwd_tc_kernel<DP>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<DP>::SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  dim3 grid((unsigned)((p.B + 127) / 128), (unsigned)n_terms);
  kern<<<grid, kTcThreads, TcCfg<DP>::SMEM, st>>>(p);
  CV_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------
// Per-row coefficients of the pair gradient from the forward statistics (a, b) [+ aux for the SupCon losses]:
//   snn / supcon_in : c = e^{-a}, q = e^{-lse_pos}   (FAST; the general path keeps the logs and exponentiates per pair)
//                     supcon_in stores b = lse_pos - log(n_k), and log(n_k) carries no gradient
//   supcon_out      : c as above, q = tau / #pos     (the positive part of the row loss is linear in s; 1/tau factored out)
// Non-finite rows (dropped from the mean) get zero coefficients.
template <bool FAST>
__device__ __forceinline__ void row_coef(int loss, float tau, bool fin, float a, float b, float aux, float& c, float& q) {
  if (loss == CLEARVAE_LOSS_SUPCON_IN) b += logf(aux);
  if (FAST) {
    c = fin ? __expf(-a) : 0.f;
    q = fin ? __expf(-b) : 0.f;
  } else {
    c = fin ? a : INFINITY;
    q = fin ? b : INFINITY;
  }
  if (loss == CLEARVAE_LOSS_SUPCON_OUT) q = fin ? tau / aux : 0.f;   // 1/tau is factored out of the whole coefficient
}

template <int DP, int RM, int SIM, bool FAST>
__global__ void __launch_bounds__(kTN) snn_bwd_kernel(const BwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sN = reinterpret_cast<float*>(smem_raw);               // [DP][kTN]
  long long* sL = reinterpret_cast<long long*>(sN + DP * kTN);  // [kTN]
  float* sC = reinterpret_cast<float*>(sL + kTN);               // [kTN]
  float* sQ = sC + kTN;                                         // [kTN]
  float* sX = sQ + kTN;                                         // [NLV][DP][kTN] (logvar-aware similarities)
  const int term = blockIdx.y;
  const TermB& t = p.t[term];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)blockIdx.x * (kWarps * RM) + warp * RM;
  const int D = p.D;
  const float g_kl = p.gscal[term], g_loss = p.gscal[2 + term];
  constexpr int NLV = SimLv<SIM>::N;
  if (!t.snn && blockIdx.z > 0) return;   // KL / reparam-only term: one CTA per row block does the epilogue

  float dn[RM][DP], row[RM][DP], inv_norm[RM], csum[RM];
  float dl[RM][NLV > 0 ? DP : 1], rlv[RM][NLV > 0 ? DP : 1][3];
  long long rl[RM];
#pragma unroll
  for (int r = 0; r < RM; ++r) {
    csum[r] = 0.f;
#pragma unroll
    for (int d = 0; d < DP; ++d) dn[r][d] = 0.f;
    if constexpr (NLV > 0) {
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        const long long i = row0 + r;
        dl[r][d] = 0.f;
        rlv[r][d][0] = rlv[r][d][1] = rlv[r][d][2] = 0.f;
        if (t.snn && i < p.B && d < D) lv_derive<SIM>(__ldg(t.lv + i * (long long)D + d), rlv[r][d]);
      }
    }
  }
  float w = 0.f;
  if (t.snn) {
    load_rows<DP, RM, SIM>(t.mu, p.lab_r, row0, p.B, D, row, inv_norm, rl);
    w = g_loss * p.inv_tau / p.scalars[CLEARVAE_S_CNT0 + term];
    // row-side coefficients from the forward stats of the global rows
    float ci[RM], qi[RM];
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      const long long i = row0 + r;
      float a = INFINITY, b = INFINITY;
      if (i < p.B) {
        a = t.stats_all[2 * (p.row_off + i)];
        b = t.stats_all[2 * (p.row_off + i) + 1];
      }
      const bool fin = isfinite(a - b);
      const float ax = (p.loss != CLEARVAE_LOSS_SNN && i < p.B) ? t.aux_all[p.row_off + i] : 1.f;
      row_coef<FAST>(p.loss, 1.f / p.inv_tau, fin, a, b, ax, ci[r], qi[r]);
    }
    const float k2 = p.inv_tau * CV_LOG2E;
    const float* cols = t.mu_cols ? t.mu_cols : t.mu;
    const float* lvc = t.lv_cols ? t.lv_cols : t.lv;
    const long long jbeg = p.splits > 1 ? (long long)blockIdx.z * p.cps : 0;
    const long long jend = p.splits > 1 ? min(p.Bg, jbeg + p.cps) : p.Bg;
    for (long long j0 = jbeg; j0 < jend; j0 += kTN) {
      __syncthreads();
      stage_column<DP, SIM>(cols, p.lab_c, j0 + threadIdx.x, p.Bg, D, sN, sL, lvc, sX);
      {
        const long long j = j0 + threadIdx.x;
        float a = INFINITY, b = INFINITY;
        if (j < p.Bg) { a = t.stats_all[2 * j]; b = t.stats_all[2 * j + 1]; }
        const bool fin = isfinite(a - b);
        const float ax = (p.loss != CLEARVAE_LOSS_SNN && j < p.Bg) ? t.aux_all[j] : 1.f;
        row_coef<FAST>(p.loss, 1.f / p.inv_tau, fin, a, b, ax, sC[threadIdx.x], sQ[threadIdx.x]);
      }
      __syncthreads();
      for (int jj = lane; jj < kTN; jj += 32) {
        const long long j = j0 + jj;
        if (j >= jend) break;
        float x[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) x[d] = sN[d * kTN + jj];
        const long long lab = sL[jj];
        const float cj = sC[jj], qj = sQ[jj];
#pragma unroll
        for (int r = 0; r < RM; ++r) {
          float s;
          float da[NLV > 0 ? DP : 1], dv[NLV > 0 ? DP : 1];
          if constexpr (NLV > 0) {
            float acc = 0.f;
#pragma unroll
            for (int d = 0; d < DP; ++d) {
              da[d] = dv[d] = 0.f;
              if (d < D)
                pair_dim<SIM, true>(row[r][d], rlv[r][d], x[d], sX[d * kTN + jj], NLV > 1 ? sX[(DP + d) * kTN + jj] : 0.f,
                                    NLV > 2 ? sX[(2 * DP + d) * kTN + jj] : 0.f, acc, da[d], dv[d]);
            }
            s = sim_from_acc<SIM>(acc, D);
          } else {
            s = pair_sim<DP, SIM>(row[r], x);
          }
          const bool cand = (j != p.row_off + row0 + r);
          const bool pos = cand && ((lab == rl[r]) != (t.ps != 0));
          float coef;
          if (p.loss == CLEARVAE_LOSS_SUPCON_OUT) {
            // d(row loss)/ds = e^{s/tau - a} / tau - [pos] / #pos : the positive part is linear in s (losses.py:167)
            const float e = FAST ? exp2f(fmaf(s, k2, -k2)) * (ci[r] + cj) : __expf(s * p.inv_tau - ci[r]) + __expf(s * p.inv_tau - cj);
            coef = e - (pos ? (qi[r] + qj) : 0.f);
          } else if (FAST) {
            const float e = exp2f(fmaf(s, k2, -k2));
            coef = e * ((ci[r] + cj) - (pos ? (qi[r] + qj) : 0.f));
          } else {
            const float xs = s * p.inv_tau;
            coef = __expf(xs - ci[r]) + __expf(xs - cj);
            if (pos) coef -= __expf(xs - qi[r]) + __expf(xs - qj);
          }
          coef = cand ? coef : 0.f;
          if constexpr (NLV > 0) {
#pragma unroll
            for (int d = 0; d < DP; ++d) { dn[r][d] = fmaf(coef, da[d], dn[r][d]); dl[r][d] = fmaf(coef, dv[d], dl[r][d]); }
          } else {
            csum[r] += coef;
#pragma unroll
            for (int d = 0; d < DP; ++d) dn[r][d] = fmaf(coef, x[d], dn[r][d]);
          }
        }
      }
    }
  }

  // ---- epilogue: reduce over lanes, chain through the similarity operand, add KL / reparam grads
  if (t.snn) {
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      csum[r] = cv::warp_sum(csum[r]);
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        dn[r][d] = cv::warp_sum(dn[r][d]);
        if constexpr (NLV > 0) dl[r][d] = cv::warp_sum(dl[r][d]);
      }
    }
  }
  if constexpr (NLV == 0) {
    if (p.splits > 1 && t.snn) {
      // column split: publish this CTA's partial sums; the last CTA of the row block merges them in z order
      constexpr int PS = DP + 1;
      float* part = p.part + ((long long)term * p.splits * p.B) * PS;
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        const long long i = row0 + r;
        if (i >= p.B) continue;
        float* dst = part + ((long long)blockIdx.z * p.B + i) * PS;
#pragma unroll
        for (int d = 0; d < DP; ++d)
          if ((d & 31) == lane) dst[d] = dn[r][d];
        if (lane == 0) dst[DP] = csum[r];
      }
      __shared__ int s_last;
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned* tk = p.tickets + (long long)term * gridDim.x + blockIdx.x;
        s_last = (atomicAdd(tk, 1u) == (unsigned)p.splits - 1) ? 1 : 0;
        if (s_last) *tk = 0u;
      }
      __syncthreads();
      if (!s_last) return;
      __threadfence();
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        const long long i = row0 + r;
        csum[r] = 0.f;
#pragma unroll
        for (int d = 0; d < DP; ++d) dn[r][d] = 0.f;
        if (i >= p.B) continue;
        for (int z = 0; z < p.splits; ++z) {
          const float* src = part + ((long long)z * p.B + i) * PS;
#pragma unroll
          for (int d = 0; d < DP; ++d) dn[r][d] += __ldcg(src + d);
          csum[r] += __ldcg(src + DP);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RM; ++r) {
    const long long i = row0 + r;
    float dot = 0.f;
    if (t.snn) {
#pragma unroll
      for (int d = 0; d < DP; ++d) dot = fmaf(dn[r][d], row[r][d], dot);
    }
    if (i >= p.B) continue;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      if ((d & 31) != lane || d >= D) continue;
      float gm = 0.f, gl = 0.f, gl_snn = 0.f;
      if (t.snn) {
        if constexpr (NLV > 0) {
          gm = w * dn[r][d];          // dn / dl already hold sum_j (G_ij + G_ji) ds_ij/d(mu_i, logvar_i)
          gl_snn = w * dl[r][d];
        } else if (SIM == SIM_COS) {
          // n = mu / max(|mu|, eps): the clamp is outside the graph (SURVEY §8a')
          const bool live = inv_norm[r] < 1.f / kCosEps;
          gm = w * (dn[r][d] - (live ? row[r][d] * dot : 0.f)) * inv_norm[r];
        } else {
          gm = w * 2.f * (dn[r][d] - csum[r] * row[r][d]);
        }
      }
      if (t.lv != nullptr) {
        const float m = t.mu[i * D + d], lv = t.lv[i * D + d];
        const float invB = 1.f / (float)p.B;
        gm = fmaf(g_kl * invB, m, gm);
        gl = g_kl * invB * 0.5f * (expf(lv) - 1.f);
        if (t.dz != nullptr) {
          const float gz = t.dz[i * p.z_stride + d];
          gm += gz;
          if (t.eps != nullptr) gl = fmaf(0.5f * gz * t.eps[i * D + d], expf(0.5f * lv), gl);
        }
      }
      t.dmu[i * D + d] = gm;
      if (t.dlv != nullptr) t.dlv[i * D + d] = gl + gl_snn;
    }
  }
}


// ---------------------------------------------------------------------------
// backward, tensor-core path (cosine, shared-shift): FlashAttention-backward-shaped
//   per column tile:  S = N_rows N_cols^T (3xTF32, TMEM)  ->  epilogue turns S into the coefficient tile
//   P_ij = e_ij ((c_i + c_j) - [pos_ij](q_i + q_j)) and writes it back over S in TMEM (tcgen05.st)
//   ->  second MMA with A = P taken from TMEM:  dN += P [N_hi | N_lo]   (same shared tile, read MN-major)
//   The [128 x 2D] gradient accumulator stays in TMEM for the whole column sweep.
// ---------------------------------------------------------------------------
// optional per-role timeline of CTA (0, 0) (tools/latent_timeline.py, include/clearvae_b200_debug.h): clock64 stamps of the first
// 64 column tiles, 16 slots per tile -- 0/1 producer (stage free, staged), 2/3 MMA (operands seen, S issued), 4/5 MMA (P seen,
// second GEMM issued), 8 epilogue (S seen), 9-11 chunk loaded, 12-14 chunk computed, 15 P published
__device__ long long* g_latent_timeline = nullptr;
#define LAT_TL(jt, slot) do { if (tl != nullptr && (jt) < 64) tl[(jt) * 16 + (slot)] = clock64(); } while (0)

template <int DP> struct TcBwdCfg {
  static constexpr int BN = 96;                       // columns per tile: 2 x (S/P_hi + P_lo) + dN must fit 512 TMEM columns
  static constexpr int KT = 3 * DP;
  static constexpr int A_BYTES = 128 * KT * 4;
  static constexpr int B_BYTES = BN * KT * 4;
  // second GEMM in bf16 (K = 16 columns per instruction): B operand [n0 | n1 | n2]^T (three bf16 parts, 24 bits), K-major,
  // rows padded to a multiple of 16
  static constexpr int NB2 = (3 * DP + 15) / 16 * 16;
  static constexpr int B2_BYTES = BN * NB2 * 2;
  // column-tile ring, decoupled from the two TMEM S/P buffers: the producers (global loads + normalise + split) run
  // NSB - 1 tiles ahead of the epilogue instead of waiting for the second MMA of tile jt - 2 to release their buffer
  // (a stage is held from staging through S, the epilogue and the second GEMM: ~3.5 tile periods)
  static constexpr int NSB = DP <= 8 ? 6 : DP <= 16 ? 5 : 4;
  // The first GEMM's operand (B_BYTES, the larger part) is only needed until S of its tile has been computed: it lives in its own
  // shorter ring, released by the S issuer, so that D = 32 (36 KB per stage) still gets NSB stages of everything else.
  static constexpr int NSB1 = DP <= 16 ? NSB : 2;
  // D = 32: staging a tile (normalise + tf32 split + three bf16 parts of 32 dims per column) takes one thread per column
  // ~2700 cycles, more than the rest of the pipeline needs per tile: two producer sets stage alternate tiles
  static constexpr int NPROD = DP <= 16 ? 1 : 2;
  static constexpr int THREADS = (13 + (NPROD - 1) * 3) * 32;
  // The TMEM accumulator of dN adds with truncation: over a 65536-column sweep (683 tiles x 24 MMAs) the bias reached 3.5e-4 of
  // the gradient's max.  Every FLUSH tiles the accumulator is drained into an fp32 shared-memory copy ([2DP][128], one row per
  // epilogue thread) and restarted, which bounds the chain length (measured error then <= 2e-5 of max at 65536 columns).
  static constexpr int FLUSH = 32;
  static constexpr int ACC_BYTES = DP * 128 * 4;
  static constexpr int SMEM = A_BYTES + NSB1 * B_BYTES + NSB * B2_BYTES + NSB * BN * 16 /*labels lo/hi, c, q*/ + ACC_BYTES + 2048 + 1024;
};

// warp 0 issues the S GEMMs, warps 1-3 are the column producers, 4-11 the epilogue, 12 issues the second GEMMs.  Two issuing
// threads because tcgen05.mma issue blocks while the pipe is busy (tools/latent_timeline.py: ~80 cycles per TS instruction, ~960
// per tile): with one thread the S GEMM of the next tile of one epilogue group queued behind the other group's second GEMM and
// the two groups ran in turns instead of side by side.
// (D = 32 adds warps 13-15 as a second producer set.)
template <int DP>
__global__ void __launch_bounds__(TcBwdCfg<DP>::THREADS, 1) snn_bwd_tc_kernel(const BwdParams p) {
  using namespace sm100;
  using C = TcBwdCfg<DP>;
  constexpr int BN = C::BN;
  constexpr int NSB = C::NSB, NSB1 = C::NSB1, NPROD = C::NPROD;
  // TMEM columns: two S buffers [b * BN, (b + 1) * BN), a ring of NPB coefficient buffers (P0 as bf16 pairs in the first BN / 2
  // words, P1 in the second), the gradient accumulator [128 x NB2] (columns part * DP + d).  With a third P buffer (fits for
  // DP <= 8) an epilogue group's first store no longer waits for the second GEMM of its own previous tile.
  constexpr int NB2 = C::NB2;
  constexpr int NPB = (2 * BN + 3 * BN + NB2 <= 512) ? 3 : 2;
  constexpr uint32_t kPCol = 2 * BN;
  constexpr uint32_t kDnCol = kPCol + NPB * BN;
  static_assert(kDnCol + NB2 <= 512, "TMEM budget");
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;
  unsigned char* sB = smem + C::A_BYTES;
  unsigned char* sB2 = sB + NSB1 * C::B_BYTES;
  int* sLab = reinterpret_cast<int*>(sB2 + NSB * C::B2_BYTES);  // [NSB][2][BN]
  float* sCQ = reinterpret_cast<float*>(sLab + 2 * NSB * BN);   // [NSB][2][BN]  (c_j, q_j)
  float* sAcc = sCQ + 2 * NSB * BN;                             // [DP][128] fp32 copy of the drained dN chunks
  uint64_t* bars = reinterpret_cast<uint64_t*>(sAcc + DP * 128);
  uint64_t *b_full = bars, *b_empty = bars + NSB, *s_full = bars + 2 * NSB, *p_full = s_full + 2, *dn_full = p_full + 3, *dn_taken = dn_full + 1;
  uint64_t *s_empty = dn_taken + 1, *p_empty = s_empty + 2;   // S read by the epilogue / P consumed by the second GEMM
  uint64_t* b1_empty = p_empty + 3;                            // [NSB1] first-GEMM operand consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b1_empty + NSB1);
  constexpr int FLUSH = C::FLUSH;
  int* sFlags = reinterpret_cast<int*>(tmem_slot + 1);          // [NSB]

  const int term = blockIdx.y;
  const TermB& t = p.t[term];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m0 = (long long)blockIdx.x * 128;
  const int D = p.D;
  const int ntiles = (int)((p.Bg + BN - 1) / BN);
  const float* cols = t.mu_cols ? t.mu_cols : t.mu;
  long long* const tl = (blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && (warp == 0 || warp == 1 || warp == 4 || warp == 8 || warp == 12))
                            ? g_latent_timeline : nullptr;

  if (threadIdx.x == 0) {
    for (int b = 0; b < NSB; ++b) { mbar_init(&b_full[b], BN); mbar_init(&b_empty[b], 1); sFlags[b] = 0; }
    for (int b = 0; b < NSB1; ++b) mbar_init(&b1_empty[b], 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], 4); }
    for (int b = 0; b < NPB; ++b) { mbar_init(&p_full[b], 4); mbar_init(&p_empty[b], 1); }
    mbar_init(dn_full, 1);
    mbar_init(dn_taken, 4);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  const long long hi_ref = (long long)(p.lab_c[0] >> 32);
  if (threadIdx.x >= 128 && threadIdx.x < 256) {
    const int rr = threadIdx.x - 128;
    const long long i = m0 + rr;
    stage_split<DP>(t.mu + (i < p.B ? i : 0) * (long long)D, i < p.B, D, sA, 128, rr, true);
#pragma unroll
    for (int c = 0; c < DP; ++c) sAcc[c * 128 + rr] = 0.f;
  }
  if (NB2 > 3 * DP)   // padding rows of the second GEMM's B operand stay zero for the whole sweep
    for (int o = threadIdx.x * 16; o < NSB * C::B2_BYTES; o += C::THREADS * 16) *reinterpret_cast<uint4*>(sB2 + o) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  static_assert(BN == 96, "three producer warps, one thread per column");
  if ((warp >= 1 && warp < 4) || warp >= 13) {
    // ================= column-tile producers (one thread per column; warps 1-3, second set 13-15) =================
    {
      // the global loads of the set's next tile (vector, label, row statistics) are in flight while this one is normalised and stored
      const int pset = warp >= 13 ? 1 : 0;
      const int cc = threadIdx.x - (pset ? 13 * 32 : 32);
      float nv[DP];
      long long nlab;
      float na, nq;
      bool nvalid;
      auto prefetch = [&](int jt) {
        const long long j = (long long)jt * BN + cc;
        nvalid = j < p.Bg;
        load_vec<DP>(cols + (nvalid ? j : 0) * (long long)D, nvalid, D, nv);
        nlab = nvalid ? __ldg(p.lab_c + j) : 0;
        na = nvalid ? __ldg(t.stats_all + 2 * j) : INFINITY;
        nq = nvalid ? __ldg(t.stats_all + 2 * j + 1) : INFINITY;
      };
      if (pset < ntiles) prefetch(pset);
      for (int jt = pset; jt < ntiles; jt += NPROD) {
        const int b = jt % NSB, b1 = jt % NSB1;
        float cvv[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) cvv[d] = nv[d];
        const long long lab = nlab;
        const float a = na, q = nq;
        const bool valid = nvalid;
        if (jt + NPROD < ntiles) prefetch(jt + NPROD);
        mbar_wait(&b_empty[b], ((jt / NSB) & 1) ^ 1);
        mbar_wait(&b1_empty[b1], ((jt / NSB1) & 1) ^ 1);
        LAT_TL(jt, 0);
        stage_split_vals<DP, true>(cvv, sB + b1 * C::B_BYTES, BN, cc, false, nullptr, sB2 + b * C::B2_BYTES, NB2);
        sLab[(b * 2 + 0) * BN + cc] = (int)(lab & 0xffffffffll);
        sLab[(b * 2 + 1) * BN + cc] = (int)(lab >> 32);
        const bool fin = isfinite(a - q);
        sCQ[(b * 2 + 0) * BN + cc] = fin ? __expf(-a) : 0.f;
        sCQ[(b * 2 + 1) * BN + cc] = fin ? __expf(-q) : 0.f;
        const int any_diff = __any_sync(0xffffffffu, (valid && (lab >> 32) != hi_ref) ? 1 : 0);
        if (lane == 0 && any_diff) atomicOr(&sFlags[b], 1);   // sticky per stage: only ever forces the exact (slow) epilogue path
        fence_proxy_async();
        mbar_arrive(&b_full[b]);
        LAT_TL(jt, 1);
      }
    }
  } else if (warp == 0) {
    // ================= S GEMM issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc1 = instr_desc(kFmtTF32, 128, BN, 0, 0);
      // descriptors differ only in their start-address field (bits [0,14) of the low word, 16-byte units): one add per instruction
      const uint64_t a_desc0 = smem_desc(smem_u32(sA), 128 * 16, 128, kLayoutNone);
      const uint64_t b_desc0 = smem_desc(smem_u32(sB), BN * 16, 128, kLayoutNone);
      for (int jt = 0; jt < ntiles; ++jt) {
        const int b = jt & 1, sb = jt % NSB, sb1 = jt % NSB1;
        mbar_wait(&b_full[sb], (jt / NSB) & 1);
        mbar_wait(&s_empty[b], ((jt >> 1) & 1) ^ 1);   // the epilogue has read S of tile jt - 2 out of this buffer
        tc_fence_after();
        LAT_TL(jt, 2);
        const uint64_t bsd = b_desc0 + (uint32_t)(sb1 * (C::B_BYTES >> 4));
#pragma unroll
        for (int k8 = 0; k8 < C::KT / 8; ++k8) {
          const uint64_t ad = a_desc0 + (uint32_t)(k8 * ((2 * (128 * 16)) >> 4));
          const uint64_t bd = bsd + (uint32_t)(k8 * ((2 * (BN * 16)) >> 4));
          umma_tf32(tmem_base + b * BN, ad, bd, idesc1, k8 != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[b]);
        umma_commit(&b1_empty[sb1]);
        LAT_TL(jt, 3);
      }
    }
  } else if (warp == 12) {
    // ================= second GEMM issuer: dN += P [n0 | n1 | n2] =================
    if (lane == 0) {
      constexpr uint32_t idesc2 = instr_desc(kFmtBF16, 128, NB2, 0, 0);
      const uint64_t b2_desc0 = smem_desc(smem_u32(sB2), NB2 * 16, 128, kLayoutNone);
      uint32_t drains = 0;   // accumulator chunks handed to the epilogue so far
      for (int jt = 0; jt < ntiles; ++jt) {
        const int pb = jt % NPB, sb = jt % NSB;
        mbar_wait(&b_full[sb], (jt / NSB) & 1);       // (already implied by P of this tile: S was computed from the same stage)
        mbar_wait(&p_full[pb], (jt / NPB) & 1);
        tc_fence_after();
        LAT_TL(jt, 4);
        const bool restart = (jt % FLUSH) == 0;       // first tile of a chunk: the accumulator starts over
        if (restart && jt > 0) {                      // ... once the epilogue has drained the previous chunk
          mbar_wait(dn_taken, (drains - 1) & 1);
          tc_fence_after();
        }
        const uint64_t b2d = b2_desc0 + (uint32_t)(sb * (C::B2_BYTES >> 4));
        const uint32_t p_col = tmem_base + kPCol + pb * BN;
#pragma unroll
        for (int part = 0; part < 2; ++part) {        // A = P0, then P1 (bf16 head + bf16 tail: 16 significand bits)
#pragma unroll
          for (int k16 = 0; k16 < BN / 16; ++k16) {   // 16 columns = 8 TMEM words of A, two 16-byte K chunks of B per instruction
            const uint64_t bd = b2d + (uint32_t)(k16 * ((2 * (NB2 * 16)) >> 4));
            umma_f16_ts(tmem_base + kDnCol, p_col + part * (BN / 2) + k16 * 8, bd, idesc2,
                        (restart && part == 0 && k16 == 0) ? 0u : 1u);
          }
        }
        umma_commit(&b_empty[sb]);                    // the stage's S GEMM retired before this tile's P existed
        umma_commit(&p_empty[pb]);
        LAT_TL(jt, 5);
        if (((jt + 1) % FLUSH == 0 && jt + 1 < ntiles) || jt + 1 == ntiles) { umma_commit(dn_full); ++drains; }   // chunk complete -> drain
      }
    }
  } else {
    // ================= epilogue: two groups of 4 warps alternate over the column tiles =================
    const int ew = warp - 4, lane_grp = warp & 3, grp = ew >> 2;
    const int r = lane_grp * 32 + lane;
    const long long i = m0 + r;
    const long long my_lab = i < p.B ? p.lab_r[i] : 0;
    const int my_lo = (int)(my_lab & 0xffffffffll), my_hi = (int)(my_lab >> 32);
    const bool my_hi_odd = (my_lab >> 32) != hi_ref;
    const long long diag = p.row_off + i;
    const bool ps = t.ps != 0;
    const float k2 = p.inv_tau * CV_LOG2E;
    float ci = 0.f, qi = 0.f;
    if (i < p.B) {
      const float a = t.stats_all[2 * diag], q = t.stats_all[2 * diag + 1];
      if (isfinite(a - q)) { ci = __expf(-a); qi = __expf(-q); }
    }
    const float2 ci2 = make_float2(ci, ci), qi2 = make_float2(qi, qi), k2v = make_float2(k2, k2), nk2v = make_float2(-k2, -k2),
                 neg1 = make_float2(-1.f, -1.f);
    // group 0 drains the dN accumulator every FLUSH tiles: chunk k is complete once the second MMA of tile (k+1)*FLUSH - 1 has
    // retired (dn_full, phase k & 1); the MMA warp restarts the accumulator only after all four warps have taken it (dn_taken)
    int drained = 0;
    auto drain = [&]() {
      mbar_wait(dn_full, drained & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < NB2; c0 += 16) {
        uint32_t r16[16];
        tmem_ld16(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + kDnCol + (uint32_t)c0, r16);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; ++q)
          if (c0 + q < 3 * DP) sAcc[((c0 + q) % DP) * 128 + r] += __uint_as_float(r16[q]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dn_taken);
      ++drained;
    };
    for (int jt = grp; jt < ntiles; jt += 2) {
      const int b = grp;
      if (grp == 0)
        while ((drained + 1) * FLUSH <= jt && (drained + 1) * FLUSH < ntiles) drain();
      mbar_wait(&s_full[b], (jt >> 1) & 1);
      tc_fence_after();
      LAT_TL(jt, 8);
      const long long jbase = (long long)jt * BN;
      const int sb = jt % NSB;   // shared-memory stage of this column tile (b indexes the TMEM S/P buffer)
      const bool edge = (jbase + BN > p.Bg) || (diag >= jbase && diag < jbase + BN) || sFlags[sb] || my_hi_odd;
      const int* lab_lo = sLab + (sb * 2 + 0) * BN;
      const int* lab_hi = sLab + (sb * 2 + 1) * BN;
      const float* cj = sCQ + (sb * 2 + 0) * BN;
      const float* qj = sCQ + (sb * 2 + 1) * BN;
      const int pb = jt % NPB;
      const uint32_t tcol = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(b * BN);
      const uint32_t pcol = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + kPCol + (uint32_t)(pb * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t rw[32], p0w[16], p1w[16];
        tmem_ld32(tcol + (uint32_t)c0, rw);
        tmem_ld_wait();
        LAT_TL(jt, 9 + c0 / 32);
        if (c0 + 32 >= BN) {            // S is in registers: the buffer's S columns may take tile jt + 2
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[b]);
        }
        if (!edge) {
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            const int4 l4 = *reinterpret_cast<const int4*>(lab_lo + c0 + q);
            const float4 c4 = *reinterpret_cast<const float4*>(cj + c0 + q);
            const float4 q4 = *reinterpret_cast<const float4*>(qj + c0 + q);
            // packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2): ~7 issue slots per pair instead of ~12
            const float2 t01 = __ffma2_rn(make_float2(__uint_as_float(rw[q]), __uint_as_float(rw[q + 1])), k2v, nk2v);
            const float2 t23 = __ffma2_rn(make_float2(__uint_as_float(rw[q + 2]), __uint_as_float(rw[q + 3])), k2v, nk2v);
            const float2 e01 = make_float2(ex2_approx(t01.x), ex2_approx(t01.y));
            const float2 e23 = make_float2(ex2_approx(t23.x), ex2_approx(t23.y));
            const float2 a01 = __fadd2_rn(ci2, make_float2(c4.x, c4.y)), a23 = __fadd2_rn(ci2, make_float2(c4.z, c4.w));
            float2 b01 = __fadd2_rn(qi2, make_float2(q4.x, q4.y)), b23 = __fadd2_rn(qi2, make_float2(q4.z, q4.w));
            b01.x = ((l4.x == my_lo) != ps) ? b01.x : 0.f;
            b01.y = ((l4.y == my_lo) != ps) ? b01.y : 0.f;
            b23.x = ((l4.z == my_lo) != ps) ? b23.x : 0.f;
            b23.y = ((l4.w == my_lo) != ps) ? b23.y : 0.f;
            const float2 x01 = __fmul2_rn(e01, __ffma2_rn(b01, neg1, a01));
            const float2 x23 = __fmul2_rn(e23, __ffma2_rn(b23, neg1, a23));
            // bf16 head by truncation (the packed pair is one PRMT of the two high halves), bf16 tail of the exact remainder
            const float2 h01 = make_float2(bf16_trunc(x01.x), bf16_trunc(x01.y)), h23 = make_float2(bf16_trunc(x23.x), bf16_trunc(x23.y));
            const float2 l01 = __ffma2_rn(h01, neg1, x01), l23 = __ffma2_rn(h23, neg1, x23);
            p0w[q / 2] = __byte_perm(__float_as_uint(x01.x), __float_as_uint(x01.y), 0x7632);
            p0w[q / 2 + 1] = __byte_perm(__float_as_uint(x23.x), __float_as_uint(x23.y), 0x7632);
            p1w[q / 2] = pack_bf16x2(l01.x, l01.y);
            p1w[q / 2 + 1] = pack_bf16x2(l23.x, l23.y);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const long long j = jbase + c0 + q;
            const float e = ex2_approx(fmaf(__uint_as_float(rw[q]), k2, -k2));
            const bool cand = (j < p.Bg) && (j != diag);
            const bool same = (lab_lo[c0 + q] == my_lo) && (lab_hi[c0 + q] == my_hi);
            const float x = cand ? e * ((ci + cj[c0 + q]) - ((same != ps) ? (qi + qj[c0 + q]) : 0.f)) : 0.f;
            rw[q] = __float_as_uint(x);
          }
#pragma unroll
          for (int q = 0; q < 32; q += 2) {
            const float x0 = __uint_as_float(rw[q]), x1 = __uint_as_float(rw[q + 1]);
            p0w[q / 2] = __byte_perm(rw[q], rw[q + 1], 0x7632);
            p1w[q / 2] = pack_bf16x2(x0 - bf16_trunc(x0), x1 - bf16_trunc(x1));
          }
        }
        LAT_TL(jt, 12 + c0 / 32);
        if (c0 == 0) {                  // P columns of this buffer: the second GEMM of tile jt - 2 has retired
          mbar_wait(&p_empty[pb], ((jt / NPB) & 1) ^ 1);
          tc_fence_after();
        }
        tmem_st16(pcol + (uint32_t)(c0 / 2), p0w);
        tmem_st16(pcol + (uint32_t)(BN / 2 + c0 / 2), p1w);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[pb]);
      LAT_TL(jt, 15);
    }
    // ---- final: dN (TMEM) -> chain through the normalisation, add KL / reparam gradients
    if (grp == 0) {
      while ((drained + 1) * FLUSH < ntiles) drain();   // chunks completed after this group's last tile
      mbar_wait(dn_full, drained & 1);                  // the last (partial) chunk
      tc_fence_after();
      float acc[DP];
#pragma unroll
      for (int c = 0; c < DP; ++c) acc[c] = sAcc[c * 128 + r];
#pragma unroll
      for (int c0 = 0; c0 < NB2; c0 += 16) {
        uint32_t r16[16];
        tmem_ld16(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + kDnCol + (uint32_t)c0, r16);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; ++q)
          if (c0 + q < 3 * DP) acc[(c0 + q) % DP] += __uint_as_float(r16[q]);
      }
      if (i < p.B) {
        const float g_kl = p.gscal[term], g_loss = p.gscal[2 + term];
        const float w = g_loss * p.inv_tau / p.scalars[CLEARVAE_S_CNT0 + term];
        float n[DP], mu[DP];
        float ss = 0.f;
#pragma unroll
        for (int d = 0; d < DP; ++d) { mu[d] = d < D ? t.mu[i * D + d] : 0.f; ss = fmaf(mu[d], mu[d], ss); }
        const float inv = 1.f / fmaxf(sqrtf(ss), kCosEps);
        const bool live = inv < 1.f / kCosEps;
        float dot = 0.f;
#pragma unroll
        for (int d = 0; d < DP; ++d) { n[d] = mu[d] * inv; dot = fmaf(acc[d], n[d], dot); }
        const float invB = 1.f / (float)p.B;
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          if (d >= D) continue;
          float gm = w * (acc[d] - (live ? n[d] * dot : 0.f)) * inv, gl = 0.f;
          if (t.lv != nullptr) {
            const float lv = t.lv[i * D + d];
            gm = fmaf(g_kl * invB, mu[d], gm);
            gl = g_kl * invB * 0.5f * (expf(lv) - 1.f);
            if (t.dz != nullptr) {
              const float gz = t.dz[i * p.z_stride + d];
              gm += gz;
              if (t.eps != nullptr) gl = fmaf(0.5f * gz * t.eps[i * D + d], expf(0.5f * lv), gl);
            }
          }
          t.dmu[i * D + d] = gm;
          if (t.dlv != nullptr) t.dlv[i * D + d] = gl;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int DP>
int launch_bwd_tc(const BwdParams& p, int n_terms, cudaStream_t st) {
  auto kern = snn_bwd_tc_kernel<DP>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcBwdCfg<DP>::SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  dim3 grid((unsigned)((p.B + 127) / 128), (unsigned)n_terms);
  kern<<<grid, TcBwdCfg<DP>::THREADS, TcBwdCfg<DP>::SMEM, st>>>(p);
  CV_LAUNCH_CHECK();
  return 0;
}

__global__ void pair_mask_kernel(const long long* lab_r, const long long* lab_c, long long B, long long Bg,
                                 long long row_off, int ps, unsigned char* out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * Bg) return;
  const long long i = idx / Bg, j = idx % Bg;
  const bool cand = (j != row_off + i);
  const bool pos = cand && ((lab_c[j] == lab_r[i]) != (ps != 0));
  out[idx] = (unsigned char)((cand ? 1 : 0) | (pos ? 2 : 0));
}

// ---------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------
template <int DP, int SIM>
constexpr size_t fwd_smem() { return (size_t)DP * kTN * 4 + kTN * 8 + (kWarps + 1) * 4 + (size_t)SimLv<SIM>::N * DP * kTN * 4; }
template <int DP, int SIM>
constexpr size_t bwd_smem() { return (size_t)DP * kTN * 4 + kTN * 8 + 2 * kTN * 4 + (size_t)SimLv<SIM>::N * DP * kTN * 4; }

template <int DP, int RM, int SIM, bool FAST>
int launch_fwd(const FwdParams& p, int n_terms, cudaStream_t st) {
  auto kern = snn_fwd_kernel<DP, RM, SIM, FAST>;
  constexpr size_t smem = fwd_smem<DP, SIM>();
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((unsigned)((p.B + kWarps * RM - 1) / (kWarps * RM)), (unsigned)n_terms, (unsigned)max(1, p.splits));
  kern<<<grid, kTN, smem, st>>>(p);
  CV_LAUNCH_CHECK();
  return 0;
}
template <int DP, int RM, int SIM, bool FAST>
int launch_bwd(const BwdParams& p, int n_terms, cudaStream_t st) {
  auto kern = snn_bwd_kernel<DP, RM, SIM, FAST>;
  constexpr size_t smem = bwd_smem<DP, SIM>();
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((unsigned)((p.B + kWarps * RM - 1) / (kWarps * RM)), (unsigned)n_terms, (unsigned)max(1, p.splits));
  kern<<<grid, kTN, smem, st>>>(p);
  CV_LAUNCH_CHECK();
  return 0;
}

// rows per warp: register-block when there are enough rows to still fill the GPU
inline int pick_rm(int dp, long long B) {
  if (B <= 2048) return 1;
  return dp == 8 ? 4 : (dp == 16 ? 2 : 1);
}

#define CV_DISPATCH(LAUNCH, P, NT, ST, DPV, RMV, SIMV, FASTV)                                    \
  do {                                                                                            \
    if (SIMV == SIM_COS) {                                                                        \
      if (FASTV) return LAUNCH<DPV, RMV, SIM_COS, true>(P, NT, ST);                               \
      return LAUNCH<DPV, RMV, SIM_COS, false>(P, NT, ST);                                         \
    }                                                                                             \
    if (SIMV == SIM_L2) return LAUNCH<DPV, RMV, SIM_L2, false>(P, NT, ST);                        \
  } while (0)
// logvar-aware similarities: one row per warp (their per-dimension state fills the registers), D <= 32
#define CV_DISPATCH_LV(LAUNCH, P, NT, ST, DPV, SIMV)                                             \
  do {                                                                                            \
    if (SIMV == SIM_ML2) return LAUNCH<DPV, 1, SIM_ML2, false>(P, NT, ST);                        \
    if (SIMV == SIM_MAH) return LAUNCH<DPV, 1, SIM_MAH, false>(P, NT, ST);                        \
    if (SIMV == SIM_JEF) return LAUNCH<DPV, 1, SIM_JEF, false>(P, NT, ST);                        \
  } while (0)

#define CV_DISPATCH_D(LAUNCH, P, NT, ST, dp, rm, sim, fast)                                      \
  do {                                                                                            \
    if (sim >= SIM_ML2) {                                                                         \
      if (dp == 8) CV_DISPATCH_LV(LAUNCH, P, NT, ST, 8, sim);                                     \
      if (dp == 16) CV_DISPATCH_LV(LAUNCH, P, NT, ST, 16, sim);                                   \
      if (dp == 32) CV_DISPATCH_LV(LAUNCH, P, NT, ST, 32, sim);                                   \
      return CLEARVAE_EUNSUPPORTED;                                                               \
    }                                                                                             \
    if (dp == 8) { if (rm == 4) CV_DISPATCH(LAUNCH, P, NT, ST, 8, 4, sim, fast); CV_DISPATCH(LAUNCH, P, NT, ST, 8, 1, sim, fast); } \
    if (dp == 16) { if (rm == 2) CV_DISPATCH(LAUNCH, P, NT, ST, 16, 2, sim, fast); CV_DISPATCH(LAUNCH, P, NT, ST, 16, 1, sim, fast); } \
    if (dp == 32) CV_DISPATCH(LAUNCH, P, NT, ST, 32, 1, sim, fast);                               \
    if (dp == 64) CV_DISPATCH(LAUNCH, P, NT, ST, 64, 1, sim, fast);                               \
  } while (0)

inline int pad_d(int D) { return D <= 8 ? 8 : D <= 16 ? 16 : D <= 32 ? 32 : D <= 64 ? 64 : -1; }
// shared-shift path is safe while exp(-2/tau) stays a normal float
inline bool fast_ok(int sim, float tau) { return sim == SIM_COS && tau > 0.f && 2.f / tau <= 80.f; }

// the tensor-core forward pays off once there are enough 128-row tiles to cover the SMs
int g_tc_min_rows = 4096;

int dispatch_fwd(const FwdParams& p, int n_terms, int sim, bool fast, cudaStream_t st) {
  const int dp = pad_d(p.D), rm = pick_rm(dp, p.B);
  if (fast && sim == SIM_COS && dp <= 32 && p.B >= g_tc_min_rows && p.loss == CLEARVAE_LOSS_SNN) {
    bool all_snn = true;  // KL/reparam-only terms have no tile loop; keep them on the FFMA kernel
    for (int i = 0; i < n_terms; ++i) all_snn &= p.t[i].snn != 0;
    if (all_snn) {
      if (dp == 8) return launch_fwd_tc<8>(p, n_terms, st);
      if (dp == 16) return launch_fwd_tc<16>(p, n_terms, st);
      return launch_fwd_tc<32>(p, n_terms, st);
    }
  }
  CV_DISPATCH_D(launch_fwd, p, n_terms, st, dp, rm, sim, fast);
  return CLEARVAE_EUNSUPPORTED;
}
int dispatch_bwd(const BwdParams& p, int n_terms, int sim, bool fast, cudaStream_t st) {
  const int dp = pad_d(p.D), rm = pick_rm(dp, p.B);
  if (fast && sim == SIM_COS && dp <= 32 && p.B >= g_tc_min_rows && p.loss == CLEARVAE_LOSS_SNN) {
    bool all_snn = true;
    for (int i = 0; i < n_terms; ++i) all_snn &= p.t[i].snn != 0;
    if (all_snn) {
      if (dp == 8) return launch_bwd_tc<8>(p, n_terms, st);
      if (dp == 16) return launch_bwd_tc<16>(p, n_terms, st);
      return launch_bwd_tc<32>(p, n_terms, st);
    }
  }
  CV_DISPATCH_D(launch_bwd, p, n_terms, st, dp, rm, sim, fast);
  return CLEARVAE_EUNSUPPORTED;
}

struct WsLayout {
  unsigned ticket;
  unsigned pad[63];
};

inline int max_ctas_for(long long B) { return (int)((B + kWarps - 1) / kWarps); }

// Column split of the FFMA kernels.  A CTA owns 8 rows and sweeps every column, staging each 256-column tile behind two
// barriers: with few rows (a data-parallel shard against the gathered global batch: 1024 x 8192) the grid is one CTA per
// SM and the sweep is a chain of exposed load latencies.  Splitting the columns over blockIdx.z puts 4+ CTAs on every SM.
constexpr int kMaxSplits = 8;
constexpr long long kSplitMaxRows = 4096;   // beyond this the row blocks alone fill the GPU (and the TC path takes over)
inline int pick_splits(long long B, long long Bg, int n_terms, int sim, int loss) {
  if (sim >= SIM_ML2 || B > kSplitMaxRows || loss != CLEARVAE_LOSS_SNN) return 1;
  const long long ctas = (long long)max_ctas_for(B) * n_terms;
  const long long tiles = (Bg + kTN - 1) / kTN;
  long long s = (148 * 8 + ctas - 1) / ctas;   // aim for ~8 resident CTAs per SM
  s = std::min<long long>(s, std::min<long long>(kMaxSplits, tiles));
  return (int)std::max<long long>(1, s);
}
inline long long cols_per_split(long long Bg, int splits) {
  const long long tiles = (Bg + kTN - 1) / kTN;
  return ((tiles + splits - 1) / splits) * kTN;
}
inline size_t fwd_part_bytes(long long B, int n_terms) {
  return B <= kSplitMaxRows ? (size_t)n_terms * kMaxSplits * (size_t)B * 4 * sizeof(float) : 0;
}

}  // namespace

extern "C" {

int clearvae_version(void) { return 100; }

/* tuning / test hook: minimum local batch for the tensor-core latent forward (default 4096) */
void clearvae_set_latent_tc_min_rows(int32_t rows) { g_tc_min_rows = rows; }
int clearvae_debug_latent_timeline(long long* device_buffer) {
  return (int)cudaMemcpyToSymbol(g_latent_timeline, &device_buffer, sizeof(device_buffer));
}

size_t clearvae_latent_workspace_bytes(int64_t B, int64_t Bg, int32_t D, int32_t n_terms) {
  (void)Bg; (void)D; (void)n_terms;
  if (B < 0) return 0;
  const size_t base = sizeof(WsLayout) + (size_t)2 * max_ctas_for(B) * sizeof(float);
  return ((base + 15) & ~(size_t)15) + fwd_part_bytes(B, n_terms);
}

size_t clearvae_latent_bwd_workspace_bytes(int64_t B, int64_t Bg, int32_t D, int32_t n_terms) {
  (void)Bg;
  if (B < 0 || B > kSplitMaxRows || pad_d(D) < 0) return 256;
  return 256 + (size_t)n_terms * max_ctas_for(B) * sizeof(unsigned) + 16 +
         (size_t)n_terms * kMaxSplits * (size_t)B * (pad_d(D) + 1) * sizeof(float);
}

int clearvae_latent_fwd(const clearvae_term_fwd* terms, int32_t n_terms, const int64_t* label_rows,
                        const int64_t* label_cols, int64_t B, int64_t Bg, int64_t row_offset, int32_t D,
                        int32_t z_stride, int32_t sim_fn, int32_t loss_name, float temperature, float* scalars,
                        int32_t finalize, void* workspace, size_t workspace_bytes, void* stream) {
  if (!terms || n_terms < 1 || n_terms > 2 || !scalars || !workspace || B <= 0 || Bg <= 0 || D <= 0) return CLEARVAE_EINVAL;
  if (loss_name < CLEARVAE_LOSS_SNN || loss_name > CLEARVAE_LOSS_SUPCON_OUT) return CLEARVAE_EUNSUPPORTED;
  if (sim_fn < SIM_COS || sim_fn > SIM_MAH) return CLEARVAE_EUNSUPPORTED;
  if (pad_d(D) < 0) return CLEARVAE_EUNSUPPORTED;
  if (workspace_bytes < clearvae_latent_workspace_bytes(B, Bg, D, n_terms)) return CLEARVAE_EWORKSPACE;
  if (row_offset < 0 || row_offset + B > Bg) return CLEARVAE_EINVAL;
  FwdParams p{};
  bool any_snn = false;
  for (int i = 0; i < n_terms; ++i) {
    const clearvae_term_fwd& s = terms[i];
    if (!s.mu) return CLEARVAE_EINVAL;
    if (s.snn_enable && (!s.row_stats || !label_rows)) return CLEARVAE_EINVAL;
    if (s.snn_enable && sim_fn >= SIM_ML2 && !s.logvar) return CLEARVAE_EINVAL;   // these similarities read logvar
    if (s.snn_enable && loss_name != CLEARVAE_LOSS_SNN && !s.row_aux) return CLEARVAE_EINVAL;
    p.t[i] = TermF{s.mu, s.logvar, s.eps, s.mu_cols, s.logvar_cols, s.z, s.row_stats, s.snn_enable, s.ps, s.row_aux};
    any_snn |= s.snn_enable != 0;
  }
  p.lab_r = reinterpret_cast<const long long*>(label_rows);
  p.lab_c = reinterpret_cast<const long long*>(label_cols ? label_cols : label_rows);
  p.B = B; p.Bg = Bg; p.row_off = row_offset; p.D = D; p.z_stride = z_stride; p.finalize = finalize;
  p.max_ctas = max_ctas_for(B);
  p.inv_tau = 1.f / temperature;
  p.loss = loss_name;
  p.scalars = scalars;
  p.ticket = &reinterpret_cast<WsLayout*>(workspace)->ticket;
  p.kl_partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + sizeof(WsLayout));
  {
    const size_t base = (sizeof(WsLayout) + (size_t)2 * max_ctas_for(B) * sizeof(float) + 15) & ~(size_t)15;
    p.part = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + base);
    p.splits = any_snn ? pick_splits(B, Bg, n_terms, sim_fn, loss_name) : 1;
    p.cps = cols_per_split(Bg, p.splits);
  }
  return dispatch_fwd(p, n_terms, sim_fn, fast_ok(sim_fn, temperature), (cudaStream_t)stream);
}

int clearvae_snn_finalize(const float* row_stats_all, int64_t Bg, int32_t term, float* scalars, void* stream) {
  if (!row_stats_all || !scalars || Bg <= 0 || term < 0 || term > 1) return CLEARVAE_EINVAL;
  snn_finalize_kernel<<<1, kTN, 0, (cudaStream_t)stream>>>(row_stats_all, Bg, term, scalars);
  CV_LAUNCH_CHECK();
  return 0;
}

int clearvae_latent_bwd(const clearvae_term_bwd* terms, int32_t n_terms, const int64_t* label_rows,
                        const int64_t* label_cols, int64_t B, int64_t Bg, int64_t row_offset, int32_t D,
                        int32_t z_stride, int32_t sim_fn, int32_t loss_name, float temperature,
                        const float* scalars, const float* gscal, void* stream) {
  return clearvae_latent_bwd_ws(terms, n_terms, label_rows, label_cols, B, Bg, row_offset, D, z_stride, sim_fn, loss_name,
                                temperature, scalars, gscal, nullptr, 0, stream);
}

int clearvae_latent_bwd_ws(const clearvae_term_bwd* terms, int32_t n_terms, const int64_t* label_rows,
                           const int64_t* label_cols, int64_t B, int64_t Bg, int64_t row_offset, int32_t D,
                           int32_t z_stride, int32_t sim_fn, int32_t loss_name, float temperature,
                           const float* scalars, const float* gscal, void* workspace, size_t workspace_bytes, void* stream) {
  if (!terms || n_terms < 1 || n_terms > 2 || !scalars || !gscal || B <= 0 || Bg <= 0 || D <= 0) return CLEARVAE_EINVAL;
  if (loss_name < CLEARVAE_LOSS_SNN || loss_name > CLEARVAE_LOSS_SUPCON_OUT) return CLEARVAE_EUNSUPPORTED;
  if (sim_fn < SIM_COS || sim_fn > SIM_MAH) return CLEARVAE_EUNSUPPORTED;
  if (pad_d(D) < 0) return CLEARVAE_EUNSUPPORTED;
  if (row_offset < 0 || row_offset + B > Bg) return CLEARVAE_EINVAL;
  BwdParams p{};
  for (int i = 0; i < n_terms; ++i) {
    const clearvae_term_bwd& s = terms[i];
    if (!s.mu || !s.dmu) return CLEARVAE_EINVAL;
    if (s.snn_enable && (!s.row_stats_all || !label_rows)) return CLEARVAE_EINVAL;
    if (s.snn_enable && sim_fn >= SIM_ML2 && (!s.logvar || !s.dlogvar)) return CLEARVAE_EINVAL;
    if (s.snn_enable && loss_name != CLEARVAE_LOSS_SNN && !s.row_aux_all) return CLEARVAE_EINVAL;
    p.t[i] = TermB{s.mu, s.logvar, s.eps, s.mu_cols, s.row_stats_all, s.dz, s.dmu, s.dlogvar, s.snn_enable, s.ps, s.logvar_cols,
                   s.row_aux_all};
  }
  p.lab_r = reinterpret_cast<const long long*>(label_rows);
  p.lab_c = reinterpret_cast<const long long*>(label_cols ? label_cols : label_rows);
  p.B = B; p.Bg = Bg; p.row_off = row_offset; p.D = D; p.z_stride = z_stride;
  p.inv_tau = 1.f / temperature;
  p.scalars = scalars; p.gscal = gscal;
  p.loss = loss_name;
  p.splits = 1;
  p.cps = cols_per_split(Bg, 1);
  bool any_snn = false;
  for (int i = 0; i < n_terms; ++i) any_snn |= terms[i].snn_enable != 0;
  if (workspace && any_snn && workspace_bytes >= clearvae_latent_bwd_workspace_bytes(B, Bg, D, n_terms)) {
    p.splits = pick_splits(B, Bg, n_terms, sim_fn, loss_name);
    p.cps = cols_per_split(Bg, p.splits);
    p.tickets = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(workspace) + 256);
    const size_t off = (256 + (size_t)n_terms * max_ctas_for(B) * sizeof(unsigned) + 15) & ~(size_t)15;
    p.part = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + off);
  }
  return dispatch_bwd(p, n_terms, sim_fn, fast_ok(sim_fn, temperature), (cudaStream_t)stream);
}

int clearvae_pair_mask(const int64_t* label_rows, const int64_t* label_cols, int64_t B, int64_t Bg,
                       int64_t row_offset, int32_t ps, uint8_t* out, void* stream) {
  if (!label_rows || !out || B <= 0 || Bg <= 0) return CLEARVAE_EINVAL;
  const long long n = B * Bg;
  pair_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long*>(label_rows),
      reinterpret_cast<const long long*>(label_cols ? label_cols : label_rows), B, Bg, row_offset, ps, out);
  CV_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
