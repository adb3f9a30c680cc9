// Group-evidence accumulation of the ML-VAE / GVAE baselines and its group-wise reparameterisation
// (reference code/src/models/vae.py:159-223), as segmented reductions over label groups.
//
//   group id of row i = gid[i] in [0, G)  (rank of the row's label among the sorted unique labels: `label.unique(sorted=True)`)
//
//   MLVAE (product of Gaussians):  a_i = -lv_i,  L_g = LSE_{i in g} a_i,   mu_g = sum_i mu_i exp(a_i - L_g),   lv_g = -L_g
//   GVAE  (average):               mu_g = mean_{i in g} mu_i,              lv_g = LSE_{i in g} lv_i - log n_g
//   reparam:                       z_i = mu_g(i) + eps_i * exp(lv_g(i) / 2)
//
// One CTA per group scans the batch (B x G row tests; the batches of the path are <= a few thousand rows and G <= a few
// hundred groups), 32 lanes over the latent dimension x 8 row lanes, fixed-order shared-memory reduction over the row lanes
// => bit-reproducible.  Backward of the accumulation is elementwise in the rows given the group results; backward of the
// reparameterisation is the same segmented reduction of the incoming gradient.
#include "common.cuh"

namespace {

constexpr int kRows = 8;   // row lanes per CTA
constexpr int kDMax = 32;  // latent width per head handled by one lane set (D <= 32: VAE 8, VAE64 32)

// ---- segmented reductions: grid = G, block = (32, kRows)
template <int MODE>  // 0: MLVAE evidence, 1: GVAE evidence, 2: reparam backward (sum dz, sum dz*eps)
__global__ void __launch_bounds__(32 * kRows) group_reduce_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                   const long long* __restrict__ gid, long long B, int D,
                                                                   float* __restrict__ out0, float* __restrict__ out1,
                                                                   float* __restrict__ count) {
  __shared__ float s_m[kRows][kDMax], s_s[kRows][kDMax], s_w[kRows][kDMax], s_n[kRows];
  const int g = blockIdx.x, d = threadIdx.x, r = threadIdx.y;
  const bool live = d < D;
  float m = -INFINITY, s = 0.f, w = 0.f, n = 0.f;   // running max / scaled sum-exp / scaled weighted sum (or plain sums)
  for (long long i = r; i < B; i += kRows) {
    if (gid[i] != g) continue;
    n += 1.f;
    if (!live) continue;
    const float x = a[i * D + d], y = b[i * D + d];
    if (MODE == 0) {            // a = mu, b = logvar: accumulate sum mu e^{-lv} and LSE(-lv) with a shared running max
      const float t = -y;
      if (t > m) { const float sc = __expf(m - t); s = s * sc + 1.f; w = w * sc + x; m = t; }
      else { const float e = __expf(t - m); s += e; w = fmaf(x, e, w); }
    } else if (MODE == 1) {     // a = mu, b = logvar: sum mu and LSE(lv)
      w += x;
      if (y > m) { s = s * __expf(m - y) + 1.f; m = y; } else { s += __expf(y - m); }
    } else {                    // a = dz, b = eps: sum dz and sum dz * eps
      s += x;
      w = fmaf(x, y, w);
    }
  }
  if (live) { s_m[r][d] = m; s_s[r][d] = s; s_w[r][d] = w; }
  if (d == 0) s_n[r] = n;
  __syncthreads();
  if (r == 0 && live) {
    float M = -INFINITY, S = 0.f, W = 0.f, N = 0.f;
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      N += s_n[k];
      if (MODE == 2) { S += s_s[k][d]; W += s_w[k][d]; continue; }
      const float mk = s_m[k][d];
      if (mk == -INFINITY) continue;
      if (mk > M) { const float sc = __expf(M - mk); S = S * sc + s_s[k][d]; W = (MODE == 0 ? W * sc : W) + s_w[k][d]; M = mk; }
      else { const float e = __expf(mk - M); S = fmaf(s_s[k][d], e, S); W = MODE == 0 ? fmaf(s_w[k][d], e, W) : W + s_w[k][d]; }
    }
    if (MODE == 0) {            // mu_g = W / S (both relative to e^M), lv_g = -(M + log S)
      out0[g * D + d] = N > 0.f ? W / S : 0.f;
      out1[g * D + d] = N > 0.f ? -(M + logf(S)) : 0.f;
    } else if (MODE == 1) {
      out0[g * D + d] = N > 0.f ? W / N : 0.f;
      out1[g * D + d] = N > 0.f ? M + logf(S) - logf(N) : 0.f;
    } else {
      out0[g * D + d] = S;      // d mu_g
      out1[g * D + d] = W;      // sum dz * eps (the caller's elementwise pass multiplies by sigma_g / 2)
    }
    if (d == 0 && count != nullptr) count[g] = N;
  }
}

// ---- elementwise passes over the rows
template <int MODE>  // 0: reparam forward, 1: MLVAE evidence backward, 2: GVAE evidence backward
__global__ void group_rows_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ eps,
                                  const long long* __restrict__ gid, const float* __restrict__ mu_g, const float* __restrict__ lv_g,
                                  const float* __restrict__ count, const float* __restrict__ dmu_g, const float* __restrict__ dlv_g,
                                  long long B, int D, float* __restrict__ o0, float* __restrict__ o1) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * D) return;
  const long long i = idx / D;
  const int d = (int)(idx - i * D);
  const long long g = gid[i];
  const float mg = mu_g[g * D + d], lg = lv_g[g * D + d];
  if (MODE == 0) {
    o0[idx] = fmaf(eps[idx], __expf(0.5f * lg), mg);
  } else if (MODE == 1) {
    // p_i = exp(-lv_i - L_g) = exp(-lv_i + lv_g);  dmu_i = p dmu_g;  dlv_i = p (dlv_g - dmu_g (mu_i - mu_g))
    const float p = __expf(lg - lv[idx]);
    o0[idx] = p * dmu_g[g * D + d];
    o1[idx] = p * (dlv_g[g * D + d] - dmu_g[g * D + d] * (mu[idx] - mg));
  } else {
    // mu_g = mean: dmu_i = dmu_g / n;  lv_g = LSE(lv) - log n: dlv_i = dlv_g softmax_i = dlv_g exp(lv_i - lv_g - log n)
    const float n = count[g];
    o0[idx] = dmu_g[g * D + d] / n;
    o1[idx] = dlv_g[g * D + d] * __expf(lv[idx] - lg) / n;
  }
}

}  // namespace

extern "C" {

int clearvae_group_evidence_fwd(int32_t mode, const float* mu, const float* logvar, const int64_t* group_id, int64_t B, int32_t D,
                                int32_t G, float* mu_grp, float* logvar_grp, float* count, void* stream) {
  if (!mu || !logvar || !group_id || !mu_grp || !logvar_grp || !count || B <= 0 || G <= 0) return CLEARVAE_EINVAL;
  if (D < 1 || D > kDMax || (mode != CLEARVAE_GROUP_MLVAE && mode != CLEARVAE_GROUP_GVAE)) return CLEARVAE_EUNSUPPORTED;
  const dim3 block(32, kRows);
  const long long* gid = reinterpret_cast<const long long*>(group_id);
  if (mode == CLEARVAE_GROUP_MLVAE) group_reduce_kernel<0><<<G, block, 0, (cudaStream_t)stream>>>(mu, logvar, gid, B, D, mu_grp, logvar_grp, count);
  else group_reduce_kernel<1><<<G, block, 0, (cudaStream_t)stream>>>(mu, logvar, gid, B, D, mu_grp, logvar_grp, count);
  CV_LAUNCH_CHECK();
  return 0;
}

int clearvae_group_evidence_bwd(int32_t mode, const float* mu, const float* logvar, const int64_t* group_id, const float* mu_grp,
                                const float* logvar_grp, const float* count, const float* dmu_grp, const float* dlogvar_grp, int64_t B,
                                int32_t D, float* dmu, float* dlogvar, void* stream) {
  if (!mu || !logvar || !group_id || !mu_grp || !logvar_grp || !count || !dmu_grp || !dlogvar_grp || !dmu || !dlogvar || B <= 0)
    return CLEARVAE_EINVAL;
  if (D < 1 || D > kDMax || (mode != CLEARVAE_GROUP_MLVAE && mode != CLEARVAE_GROUP_GVAE)) return CLEARVAE_EUNSUPPORTED;
  const long long total = (long long)B * D;
  const unsigned grid = (unsigned)((total + 255) / 256);
  const long long* gid = reinterpret_cast<const long long*>(group_id);
  if (mode == CLEARVAE_GROUP_MLVAE)
    group_rows_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(mu, logvar, nullptr, gid, mu_grp, logvar_grp, count, dmu_grp, dlogvar_grp, B, D, dmu, dlogvar);
  else
    group_rows_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(mu, logvar, nullptr, gid, mu_grp, logvar_grp, count, dmu_grp, dlogvar_grp, B, D, dmu, dlogvar);
  CV_LAUNCH_CHECK();
  return 0;
}

int clearvae_group_reparam_fwd(const float* mu_grp, const float* logvar_grp, const float* eps, const int64_t* group_id, int64_t B, int32_t D,
                               float* z, void* stream) {
  if (!mu_grp || !logvar_grp || !eps || !group_id || !z || B <= 0) return CLEARVAE_EINVAL;
  if (D < 1 || D > kDMax) return CLEARVAE_EUNSUPPORTED;
  const long long total = (long long)B * D;
  group_rows_kernel<0><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nullptr, nullptr, eps, reinterpret_cast<const long long*>(group_id),
                                                                                            mu_grp, logvar_grp, nullptr, nullptr, nullptr, B, D, z, nullptr);
  CV_LAUNCH_CHECK();
  return 0;
}

int clearvae_group_reparam_bwd(const float* dz, const float* eps, const int64_t* group_id, int64_t B, int32_t D, int32_t G, float* dmu_grp,
                               float* dz_eps_grp, void* stream) {
  if (!dz || !eps || !group_id || !dmu_grp || !dz_eps_grp || B <= 0 || G <= 0) return CLEARVAE_EINVAL;
  if (D < 1 || D > kDMax) return CLEARVAE_EUNSUPPORTED;
  group_reduce_kernel<2><<<G, dim3(32, kRows), 0, (cudaStream_t)stream>>>(dz, eps, reinterpret_cast<const long long*>(group_id), B, D, dmu_grp, dz_eps_grp,
                                                                          nullptr);
  CV_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
