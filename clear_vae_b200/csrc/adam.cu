// Fused multi-tensor Adam (one launch per optimiser step), sm_100a.
//
// Reference: the factories build `torch.optim.Adam(params, lr=...)` with default betas / eps and no weight
// decay (code/src/utils/trainer_utils.py:100,139-140,178-181); `optimizer.step()` is called once per VAE
// update and once per estimator / discriminator iteration (code/src/trainer.py:483, 698-699, 870, 886).
// torch's foreach implementation is ~12 launches plus one scalar kernel per tensor for the step counters;
// here every (param, grad, exp_avg, exp_avg_sq) quadruple of the optimiser is updated by one grid:
//     m <- m + (g - m)(1 - b1);  v <- b2 v + (1 - b2) g^2;
//     p <- p - lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps),     t = step + 1
// The step counters live on the device (CUDA-graph replay): every CTA reads step[0] before taking a ticket,
// the last CTA to finish writes t back to all counters.
#include "common.cuh"

namespace {

constexpr int kMaxT = CLEARVAE_ADAM_MAX_TENSORS;
constexpr int kNT = 256;
constexpr int kChunk = kNT * 8;

struct AdamTable {
  float* p[kMaxT];
  const float* g[kMaxT];
  float* m[kMaxT];
  float* v[kMaxT];
  int numel[kMaxT];
  int chunk_start[kMaxT + 1];
  int n;
};

__global__ void __launch_bounds__(kNT) adam_kernel(const __grid_constant__ AdamTable tb, float* steps, int n_steps, int tick,
                                                   unsigned int* counter, float lr, float b1, float b2, float eps,
                                                   float grad_scale) {
  const float t = steps[0] + 1.f;
  int ti = 0;
  while (ti + 1 < tb.n && (int)blockIdx.x >= tb.chunk_start[ti + 1]) ++ti;
  const int base = ((int)blockIdx.x - tb.chunk_start[ti]) * kChunk;
  const int n = tb.numel[ti];
  float* __restrict__ p = tb.p[ti];
  const float* __restrict__ g = tb.g[ti];
  float* __restrict__ m = tb.m[ti];
  float* __restrict__ v = tb.v[ti];
  const float bc1 = (float)(1.0 - pow((double)b1, (double)t));
  const float bc2s = sqrtf((float)(1.0 - pow((double)b2, (double)t)));
  const float step_size = lr / bc1;
#pragma unroll
  for (int k = 0; k < kChunk / kNT; ++k) {
    const int i = base + k * kNT + threadIdx.x;
    if (i < n) {
      const float gi = g[i] * grad_scale;
      const float mi = m[i] + (gi - m[i]) * (1.f - b1);
      const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
      m[i] = mi;
      v[i] = vi;
      p[i] -= step_size * (mi / (sqrtf(vi) / bc2s + eps));
    }
  }
  if (!tick) return;
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (s_last) {
    for (int i = threadIdx.x; i < n_steps; i += kNT) steps[i] = t;
    if (threadIdx.x == 0) *counter = 0u;
  }
}

}  // namespace

extern "C" {

int clearvae_adam_step(int32_t n_tensors, float* const* params_host, const float* const* grads_host, float* const* exp_avg_host,
                       float* const* exp_avg_sq_host, const int64_t* numel_host, float* steps, int32_t n_steps,
                       unsigned int* counter, float lr, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  if (n_tensors <= 0) return 0;
  if (!params_host || !grads_host || !exp_avg_host || !exp_avg_sq_host || !numel_host || !steps || !counter || n_steps < 1)
    return CLEARVAE_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  for (int t0 = 0; t0 < n_tensors; t0 += kMaxT) {
    AdamTable tb{};
    const int n = std::min(kMaxT, n_tensors - t0);
    long long chunks = 0;
    for (int i = 0; i < n; ++i) {
      const int j = t0 + i;
      if (!params_host[j] || !grads_host[j] || !exp_avg_host[j] || !exp_avg_sq_host[j] || numel_host[j] < 0 ||
          numel_host[j] > 0x7fffffffLL)
        return CLEARVAE_EINVAL;
      tb.p[i] = params_host[j]; tb.g[i] = grads_host[j]; tb.m[i] = exp_avg_host[j]; tb.v[i] = exp_avg_sq_host[j];
      tb.numel[i] = (int)numel_host[j];
      tb.chunk_start[i] = (int)chunks;
      chunks += (numel_host[j] + kChunk - 1) / kChunk;
    }
    tb.chunk_start[n] = (int)chunks;
    tb.n = n;
    if (chunks == 0) continue;
    const int last = t0 + kMaxT >= n_tensors;  // the step counters advance once, after the final group
    adam_kernel<<<(unsigned)chunks, kNT, 0, st>>>(tb, steps, n_steps, last, counter, lr, beta1, beta2, eps, grad_scale);
    CV_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
