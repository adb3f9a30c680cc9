// Fused kernels for the density-ratio total-correlation term of CLEAR-TC-VAE, sm_100a.
//
// Reference: factor_cls = Linear(Z,Z) - ReLU - Linear(Z,1) - Sigmoid (code/src/utils/trainer_utils.py:133-138), used at
//   code/src/trainer.py:664-665   d = factor_cls(z);  mi = relu(log(d / (1 - d))).mean()          (mode BOUND: value + d/dz)
//   code/src/trainer.py:573-587, 680-699   BCELoss(cat[factor_cls(z), factor_cls(factor_shuffling(z))], cat[1, 0])
//                                  with factor_shuffling "permute_1" = [z_c | roll(z_s, -1, 0)]   (mode DISC: value + parameter grads)
// One launch per use: a CTA owns 64 row instances (DISC: 32 rows x {joint, shuffled}); the MLP forward runs with four
// threads per instance (hidden units interleaved), per-instance vectors live in shared memory as [feature][instance],
// parameter gradients are per-CTA outer products and the last CTA sums the per-CTA partials in a fixed order
// (bit-reproducible).  Replaces ~40 ATen launches per training step.
#include "common.cuh"

namespace {

constexpr int kInst = 64;
constexpr int kNT = 256;
constexpr int kLd = kInst + 1;

struct TcParams {
  const float* z; long long ldz;
  int B, Z, D;          // D = Z / 2 (content | style split of factor_shuffling)
  const float *w1, *b1, *w2, *b2;
  int mode;
  int np;
  float* out;           // [np]: out[0] = value; DISC: out[1..] = flat grads (w1, b1, w2, b2)
  float* dz;            // BOUND: unit gradient [B, Z]
  float* partial;
  unsigned int* counter;
};

template <int MAXD>
__global__ void __launch_bounds__(kNT) tc_kernel(const TcParams p) {
  extern __shared__ float sm[];
  float* sW1 = sm;                       // [Z][Z]  (row j = hidden unit)
  float* sB1 = sW1 + MAXD * MAXD;        // [MAXD]
  float* sW2 = sB1 + MAXD;               // [MAXD]
  float* sx = sW2 + MAXD;                // [MAXD][kLd]
  float* sh = sx + MAXD * kLd;           // [MAXD][kLd]
  float* sdh = sh + MAXD * kLd;          // [MAXD][kLd]
  float* sa = sdh + MAXD * kLd;          // [4][kLd] partial logits, then [kLd] d(loss)/d(logit)
  float* sred = sa + 4 * kLd;            // [kNT / 32]
  __shared__ int s_last;

  const int t = threadIdx.x;
  const int B = p.B, Z = p.Z, D = p.D;
  const bool disc = p.mode == CLEARVAE_TC_DISC;
  const int ninst_total = disc ? 2 * B : B;
  const int inst0 = blockIdx.x * kInst;
  const int ninst = min(kInst, ninst_total - inst0);

  for (int i = t; i < Z * Z; i += kNT) sW1[i] = __ldg(p.w1 + i);
  if (t < Z) { sB1[t] = __ldg(p.b1 + t); sW2[t] = __ldg(p.w2 + t); }
  // instance r of this CTA: DISC -> row (inst0 + r) / 2, variant (inst0 + r) & 1 (0 joint, 1 style half rolled by one row)
  for (int idx = t; idx < kInst * Z; idx += kNT) {
    const int r = idx / Z, i = idx - r * Z;
    float v = 0.f;
    if (r < ninst) {
      const int inst = inst0 + r;
      int row = disc ? inst >> 1 : inst;
      if (disc && (inst & 1) && i >= D) row = row + 1 == B ? 0 : row + 1;   // roll(z_s, -1, 0)
      v = __ldg(p.z + (long long)row * p.ldz + i);
    }
    sx[i * kLd + r] = v;
  }
  __syncthreads();

  // ---- phase 1: hidden layer, four threads per instance (hidden units j = part, part + 4, ...)
  {
    const int r = t & (kInst - 1), part = t >> 6;
    float xr[MAXD];
#pragma unroll
    for (int i = 0; i < MAXD; ++i) xr[i] = i < Z ? sx[i * kLd + r] : 0.f;
    float logit = 0.f;
    for (int j = part; j < Z; j += 4) {
      float a = sB1[j];
      const float* w = sW1 + j * Z;
#pragma unroll
      for (int i = 0; i < MAXD; ++i)
        if (i < Z) a = fmaf(w[i], xr[i], a);
      a = fmaxf(a, 0.f);
      sh[j * kLd + r] = a;
      logit = fmaf(sW2[j], a, logit);
    }
    sa[part * kLd + r] = logit;
  }
  __syncthreads();

  // ---- phase 2: logit -> loss term and d(loss)/d(logit), one thread per instance
  float loss = 0.f;
  if (t < kInst) {
    const int r = t;
    float da = 0.f;
    if (r < ninst) {
      const float a = ((sa[r] + sa[kLd + r]) + (sa[2 * kLd + r] + sa[3 * kLd + r])) + __ldg(p.b2);
      const float d = 1.f / (1.f + expf(-a));
      if (!disc) {
        const float v = fmaxf(logf(d / (1.f - d)), 0.f);      // the reference's formula (saturates like it does)
        loss = v;
        da = v > 0.f ? 1.f / (float)B : 0.f;
      } else {
        const float inv = 0.5f / (float)B;
        if ((inst0 + r) & 1) {                                 // shuffled sample, target 0: -log(1 - d), log clamped at -100
          const float lg = logf(1.f - d);
          loss = -fmaxf(lg, -100.f);
          da = lg > -100.f ? inv * d : 0.f;
        } else {                                               // joint sample, target 1: -log(d)
          const float lg = logf(d);
          loss = -fmaxf(lg, -100.f);
          da = lg > -100.f ? -inv * (1.f - d) : 0.f;
        }
      }
    }
    sa[r] = da;   // (reads of all four partial rows by this thread are done)
  }
  const float cta_loss = cv::block_sum<kNT>(loss, sred);
  float* part_out = p.partial + (long long)blockIdx.x * p.np;
  if (t == 0) part_out[0] = cta_loss;
  __syncthreads();

  // ---- phase 3: back through the output layer and the ReLU
  for (int idx = t; idx < kInst * Z; idx += kNT) {
    const int j = idx / kInst, r = idx - j * kInst;
    sdh[j * kLd + r] = sh[j * kLd + r] > 0.f ? sa[r] * sW2[j] : 0.f;
  }
  __syncthreads();

  if (!disc) {
    // gradient w.r.t. z: W1^T dh
    for (int idx = t; idx < kInst * Z; idx += kNT) {
      const int r = idx / Z, i = idx - r * Z;
      if (r >= ninst) continue;
      float a = 0.f;
      for (int j = 0; j < Z; ++j) a = fmaf(sW1[j * Z + i], sdh[j * kLd + r], a);
      p.dz[(long long)(inst0 + r) * Z + i] = a;
    }
  } else {
    // parameter gradients: [w1 (Z*Z) | b1 (Z) | w2 (Z) | b2 (1)], outer products over this CTA's instances
    const int P = Z * Z + 2 * Z + 1;
    for (int e = t; e < P; e += kNT) {
      float acc = 0.f;
      if (e < Z * Z) {
        const int j = e / Z, i = e - j * Z;
        const float *a = sdh + j * kLd, *b = sx + i * kLd;
#pragma unroll 8
        for (int r = 0; r < kInst; ++r) acc = fmaf(a[r], b[r], acc);
      } else if (e < Z * Z + Z) {
        const float* a = sdh + (e - Z * Z) * kLd;
#pragma unroll 8
        for (int r = 0; r < kInst; ++r) acc += a[r];
      } else if (e < Z * Z + 2 * Z) {
        const float* b = sh + (e - Z * Z - Z) * kLd;
#pragma unroll 8
        for (int r = 0; r < kInst; ++r) acc = fmaf(sa[r], b[r], acc);
      } else {
#pragma unroll 8
        for (int r = 0; r < kInst; ++r) acc += sa[r];
      }
      part_out[1 + e] = acc;
    }
  }

  // ---- last CTA: fixed-order sum of the per-CTA partials
  __threadfence();
  __syncthreads();
  if (t == 0) s_last = atomicAdd(p.counter, 1u) == gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int e = t; e < p.np; e += kNT) {
    float s = 0.f;
    for (unsigned c = 0; c < gridDim.x; ++c) s += __ldcg(p.partial + (long long)c * p.np + e);
    if (e == 0) s *= disc ? 0.5f / (float)B : 1.f / (float)B;
    p.out[e] = s;
  }
  if (t == 0) *p.counter = 0u;
}

__global__ void scale_kernel(const float* __restrict__ g, const float* __restrict__ x, long long n, float* __restrict__ y) {
  const float gv = __ldg(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = gv * x[i];
}

template <int MAXD>
constexpr size_t tc_smem() { return (size_t)(MAXD * MAXD + 2 * MAXD + 3 * MAXD * kLd + 4 * kLd + kNT / 32) * sizeof(float); }

template <int MAXD>
int launch_tc(const TcParams& p, int grid, cudaStream_t st) {
  constexpr size_t smem = tc_smem<MAXD>();
  static bool attr_done = false;
  if (!attr_done && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(tc_kernel<MAXD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  tc_kernel<MAXD><<<grid, kNT, smem, st>>>(p);
  CV_LAUNCH_CHECK();
  return 0;
}

inline int tc_np(int mode, int Z) { return mode == CLEARVAE_TC_DISC ? 1 + Z * Z + 2 * Z + 1 : 1; }
inline long long tc_grid(int mode, long long B) { return ((mode == CLEARVAE_TC_DISC ? 2 * B : B) + kInst - 1) / kInst; }

}  // namespace

extern "C" {

size_t clearvae_tc_workspace_bytes(int32_t mode, int64_t B, int32_t Z) {
  if (B <= 0) return 0;
  return 256 + (size_t)tc_grid(mode, B) * tc_np(mode, Z) * sizeof(float);
}

int clearvae_tc_factor(int32_t mode, const float* z, int64_t ldz, int64_t B, int32_t Z, const float* w1, const float* b1,
                       const float* w2, const float* b2, float* out, float* dz_unit, void* workspace, size_t workspace_bytes,
                       void* stream) {
  if (!z || !w1 || !b1 || !w2 || !b2 || !out || !workspace || B <= 0 || ldz < Z) return CLEARVAE_EINVAL;
  if (mode != CLEARVAE_TC_BOUND && mode != CLEARVAE_TC_DISC) return CLEARVAE_EINVAL;
  if (mode == CLEARVAE_TC_BOUND && !dz_unit) return CLEARVAE_EINVAL;
  if (Z < 2 || Z > 64 || (Z & 1) || B > (1 << 24)) return CLEARVAE_EUNSUPPORTED;
  if (workspace_bytes < clearvae_tc_workspace_bytes(mode, B, Z)) return CLEARVAE_EWORKSPACE;
  TcParams p{};
  p.z = z; p.ldz = ldz; p.B = (int)B; p.Z = Z; p.D = Z / 2;
  p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2;
  p.mode = mode; p.np = tc_np(mode, Z);
  p.out = out; p.dz = dz_unit;
  p.counter = reinterpret_cast<unsigned int*>(workspace);
  p.partial = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + 256);
  const int grid = (int)tc_grid(mode, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (Z <= 16) return launch_tc<16>(p, grid, st);
  return launch_tc<64>(p, grid, st);
}

int clearvae_scale(const float* grad_out, const float* x, int64_t n, float* y, void* stream) {
  if (!grad_out || !x || !y || n <= 0) return CLEARVAE_EINVAL;
  const int grid = (int)std::min<long long>((n + 255) / 256, 148 * 4);
  scale_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(grad_out, x, n, y);
  CV_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
