// Measures the chip-wide ex2 (MUFU) and FFMA issue rates so the latent-loss kernel
// can be placed on a roofline ("MUFU peak is not in MEASURED_PEAKS.json").
#include <cstdio>
#include <cuda_runtime.h>

__global__ void ex2_kernel(float* out, int iters) {
  float a = threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f;
  for (int i = 0; i < iters; ++i) {
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d));
    a -= 1.f; b -= 1.f; c -= 1.f; d -= 1.f;
  }
  if (a + b + c + d == 12345.f) out[0] = a;
}
__global__ void ffma_kernel(float* out, int iters) {
  float a = threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f, e = a + .4f, f = a + .5f, g = a + .6f, h = a + .7f;
  const float m = 1.0001f, k = 1e-4f;
  for (int i = 0; i < iters; ++i) {
    a = fmaf(a, m, k); b = fmaf(b, m, k); c = fmaf(c, m, k); d = fmaf(d, m, k);
    e = fmaf(e, m, k); f = fmaf(f, m, k); g = fmaf(g, m, k); h = fmaf(h, m, k);
  }
  if (a + b + c + d + e + f + g + h == 12345.f) out[0] = a;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float* out; cudaMalloc(&out, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1 << 14, blocks = p.multiProcessorCount * 8, threads = 256;
  for (int rep = 0; rep < 3; ++rep) {
    ex2_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e0); ex2_kernel<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)blocks * threads * iters * 4;
    printf("{\"probe\":\"ex2\",\"sms\":%d,\"gexp_per_s\":%.1f,\"ms\":%.3f}\n", p.multiProcessorCount, n / ms * 1e-6, ms);
    cudaEventRecord(e0); ffma_kernel<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    n = (double)blocks * threads * iters * 8;
    printf("{\"probe\":\"ffma\",\"gffma_per_s\":%.1f,\"ms\":%.3f}\n", n / ms * 1e-6, ms);
  }
  return 0;
}
