// Reconstruction term of the ELBO (losses.py:45-47): mean_b sum_{chw} (xhat - x)^2,
// and its gradient.  Pure HBM streaming: 16-byte loads, grid sized to the SM count,
// deterministic two-level reduction (per-CTA partials, last CTA folds them).
#include <algorithm>

#include "common.cuh"

namespace {
constexpr int kNT = 256;

__global__ void __launch_bounds__(kNT) recon_fwd_kernel(const float* __restrict__ xh, const float* __restrict__ x,
                                                        long long n, float invB, float* out, float* partial,
                                                        unsigned* ticket) {
  __shared__ float sred[kNT / 32 + 1];
  float acc = 0.f;
  const long long n4 = n >> 2;
  const float4* a4 = reinterpret_cast<const float4*>(xh);
  const float4* b4 = reinterpret_cast<const float4*>(x);
  for (long long i = (long long)blockIdx.x * kNT + threadIdx.x; i < n4; i += (long long)gridDim.x * kNT) {
    const float4 a = __ldg(a4 + i), b = __ldg(b4 + i);
    float d;
    d = a.x - b.x; acc = fmaf(d, d, acc);
    d = a.y - b.y; acc = fmaf(d, d, acc);
    d = a.z - b.z; acc = fmaf(d, d, acc);
    d = a.w - b.w; acc = fmaf(d, d, acc);
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * kNT + threadIdx.x; i < n; i += (long long)gridDim.x * kNT) {
    const float d = xh[i] - x[i];
    acc = fmaf(d, d, acc);
  }
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
    if (threadIdx.x == 0) { *out = s * invB; *ticket = 0u; }
  }
}

__global__ void __launch_bounds__(kNT) recon_bwd_kernel(const float* __restrict__ xh, const float* __restrict__ x,
                                                        const float* __restrict__ g, long long n, float twoInvB,
                                                        float* __restrict__ dxh) {
  const float s = twoInvB * __ldg(g);
  const long long n4 = n >> 2;
  const float4* a4 = reinterpret_cast<const float4*>(xh);
  const float4* b4 = reinterpret_cast<const float4*>(x);
  float4* o4 = reinterpret_cast<float4*>(dxh);
  for (long long i = (long long)blockIdx.x * kNT + threadIdx.x; i < n4; i += (long long)gridDim.x * kNT) {
    const float4 a = __ldg(a4 + i), b = __ldg(b4 + i);
    o4[i] = make_float4(s * (a.x - b.x), s * (a.y - b.y), s * (a.z - b.z), s * (a.w - b.w));
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * kNT + threadIdx.x; i < n; i += (long long)gridDim.x * kNT)
    dxh[i] = s * (xh[i] - x[i]);
}

inline int grid_for(long long n) {
  long long g = (n / 4 + kNT - 1) / kNT;
  const long long cap = 148 * 8;  // 8 resident CTAs of 256 threads per SM
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}
// Several reparameterisation draws of the same (mu, logvar) in one launch (CLEAR-MIM inner loop, trainer.py:874-888:
// five forwards of an unchanged encoder differ only in their noise).  z[j][b, h*D + d] = mu_h + eps[j][h] * exp(logvar_h / 2),
// the same fmaf / expf expression as the latent kernel's fused reparameterisation (vae.py:56-60).
struct ReparamTable {
  const float* mu[2];
  const float* lv[2];
  const float* eps[CLEARVAE_REPARAM_MAX_DRAWS * 2];
  float* z[CLEARVAE_REPARAM_MAX_DRAWS];
  int heads, draws;
};

__global__ void __launch_bounds__(kNT) reparam_multi_kernel(const __grid_constant__ ReparamTable tb, long long B, int D) {
  const long long n = B * D;
  for (long long i = (long long)blockIdx.x * kNT + threadIdx.x; i < n; i += (long long)gridDim.x * kNT) {
    const long long b = i / D;
    const int d = (int)(i - b * D);
    for (int h = 0; h < tb.heads; ++h) {
      const float m = __ldg(tb.mu[h] + i), sd = expf(0.5f * __ldg(tb.lv[h] + i));
      for (int j = 0; j < tb.draws; ++j)
        tb.z[j][b * (tb.heads * D) + h * D + d] = fmaf(__ldg(tb.eps[j * tb.heads + h] + i), sd, m);
    }
  }
}

}  // namespace

extern "C" {

size_t clearvae_recon_workspace_bytes(void) { return 256 + 148 * 8 * sizeof(float); }

int clearvae_recon_fwd(const float* xhat, const float* x, int64_t B, int64_t per_sample, float* out,
                       void* workspace, size_t workspace_bytes, void* stream) {
  if (!xhat || !x || !out || !workspace || B <= 0 || per_sample <= 0) return CLEARVAE_EINVAL;
  if (workspace_bytes < clearvae_recon_workspace_bytes()) return CLEARVAE_EWORKSPACE;
  if (((uintptr_t)xhat | (uintptr_t)x) & 15) return CLEARVAE_EINVAL;
  const long long n = B * per_sample;
  unsigned* ticket = reinterpret_cast<unsigned*>(workspace);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
  recon_fwd_kernel<<<grid_for(n), kNT, 0, (cudaStream_t)stream>>>(xhat, x, n, 1.f / (float)B, out, partial, ticket);
  CV_LAUNCH_CHECK();
  return 0;
}

int clearvae_recon_bwd(const float* xhat, const float* x, const float* grad_out, int64_t B, int64_t per_sample,
                       float* dxhat, void* stream) {
  if (!xhat || !x || !grad_out || !dxhat || B <= 0 || per_sample <= 0) return CLEARVAE_EINVAL;
  if (((uintptr_t)xhat | (uintptr_t)x | (uintptr_t)dxhat) & 15) return CLEARVAE_EINVAL;
  const long long n = B * per_sample;
  recon_bwd_kernel<<<grid_for(n), kNT, 0, (cudaStream_t)stream>>>(xhat, x, grad_out, n, 2.f / (float)B, dxhat);
  CV_LAUNCH_CHECK();
  return 0;
}

int clearvae_reparam_multi(int32_t heads, int32_t draws, const float* const* mu_host, const float* const* logvar_host,
                           const float* const* eps_host, float* const* z_host, int64_t B, int32_t D, void* stream) {
  if (heads < 1 || heads > 2 || draws < 1 || draws > CLEARVAE_REPARAM_MAX_DRAWS || !mu_host || !logvar_host || !eps_host || !z_host ||
      B < 0 || D < 1)
    return CLEARVAE_EINVAL;
  if (B == 0) return 0;
  ReparamTable tb{};
  for (int h = 0; h < heads; ++h) {
    if (!mu_host[h] || !logvar_host[h]) return CLEARVAE_EINVAL;
    tb.mu[h] = mu_host[h];
    tb.lv[h] = logvar_host[h];
  }
  for (int j = 0; j < draws; ++j) {
    if (!z_host[j]) return CLEARVAE_EINVAL;
    tb.z[j] = z_host[j];
    for (int h = 0; h < heads; ++h) {
      if (!eps_host[j * heads + h]) return CLEARVAE_EINVAL;
      tb.eps[j * heads + h] = eps_host[j * heads + h];
    }
  }
  tb.heads = heads;
  tb.draws = draws;
  const long long n = B * D;
  const int grid = (int)std::min<long long>(148 * 4, (n + kNT - 1) / kNT);
  reparam_multi_kernel<<<grid, kNT, 0, (cudaStream_t)stream>>>(tb, B, D);
  CV_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
