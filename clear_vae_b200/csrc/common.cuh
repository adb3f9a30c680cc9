// Shared device helpers for libclearvae_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "clearvae_b200.h"
#include "clearvae_b200_debug.h"

#define CV_LOG2E 1.4426950408889634f

#define CV_LAUNCH_CHECK()                         \
  do {                                            \
    cudaError_t e__ = cudaGetLastError();         \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)

namespace cv {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// merge two (max, sum-of-exp) accumulators; (−inf, 0) is the identity
__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
  float mm = fmaxf(m, m2);
  if (mm == -INFINITY) { m = mm; s = 0.f; return; }
  s = s * __expf(m - mm) + s2 * __expf(m2 - mm);
  m = mm;
}

__device__ __forceinline__ void lse_push(float& m, float& s, float x) {
  if (x > m) {
    s = s * __expf(m - x) + 1.f;  // m == -inf -> s*0 + 1
    m = x;
  } else {
    s += __expf(x - m);
  }
}

// Deterministic block reduction (fixed order): result valid in thread 0.
template <int NT>
__device__ __forceinline__ float block_sum(float v, float* smem /* >= NT/32 floats */) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) smem[w] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) r += smem[i];
  }
  return r;
}

}  // namespace cv
