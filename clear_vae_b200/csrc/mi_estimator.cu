// Fused variational-MI estimator kernels (CLUB-S / L1OutUB), sm_100a.
//
// Reference: code/src/models/mi_estimator.py:108-198 (Linear95/CLUB-style Gaussian heads)
//     p_mu     = Linear(Dx,H) - ReLU - Linear(H,Dy)
//     p_logvar = Linear(Dx,H) - ReLU - Linear(H,Dy) - Tanh
// and their three uses on the CLEAR-MIM step (code/src/trainer.py:858, 874-888):
//     mode LEARN : learning_loss = mean_b sum_d [ (mu - y)^2 e^{-lv} + lv ]   -> gradients of the 8 parameters
//     mode CLUB  : CLUBSample.forward  (1/2B) sum_b sum_d e^{-lv} [ (mu - y_perm)^2 - (mu - y)^2 ]
//     mode L1OUT : L1OutUB.forward *as executed* (its [B,B]+[B,B,1] broadcast, mi_estimator.py:181-189):
//                  mean_c pos_c - mean_{b,c} ap_bc - log1p(e^-20/(B-1)), evaluated from the column moments
//                  S1 = sum_c y_c, S2 = sum_c y_c^2             -> gradients w.r.t. x and y
//
// One launch does the whole forward *and* backward of the tiny MLPs: a CTA owns 64 rows, keeps every per-row
// vector in shared memory as [feature][row] (conflict-free for row-parallel and for outer-product access),
// the two networks run on two 64-thread halves, the parameter gradients are per-CTA outer products over the
// CTA's rows, and the last CTA to finish sums the per-CTA partials in a fixed order (deterministic, no float
// atomics on the results).  All sizes are tiny (B*D^2 MACs): the point is 1 launch instead of ~70.
#include "common.cuh"

namespace {

constexpr int kRows = 64;
constexpr int kNT = 256;
constexpr int kLd = kRows + 1;

struct MiParams {
  const float* x;
  const float* y;
  const int64_t* perm;
  int B, Dx, H, Dy;
  long long ldx, ldy;   // row strides (elements) of x / y: column slices of one [B, 2D] latent tensor are read in place
  const float* w[8];  // p_mu.0.weight [H,Dx], p_mu.0.bias, p_mu.2.weight [Dy,H], p_mu.2.bias, then p_logvar.*
  int mode;
  int np;             // length of the reduced output vector
  float* out;         // [np]: out[0] = value; LEARN: out[1..] = flat parameter gradients; L1OUT: out[1..] = E_d, M_d
  float* dx;          // [B,Dx] unit gradient (bound modes)
  float* dy;          // [B,Dy] unit gradient of the direct / permuted terms (bound modes)
  float* partial;     // [gridDim.x][np]
  unsigned int* counter;
  float l1_const;
};

template <int MAXD>
__global__ void __launch_bounds__(kNT) mi_kernel(const MiParams p) {
  extern __shared__ float sm[];
  constexpr int MM = MAXD * MAXD;
  float* sW1 = sm;                 // [2][H*Dx]
  float* sW2 = sW1 + 2 * MM;       // [2][Dy*H]
  float* sB1 = sW2 + 2 * MM;       // [2][MAXD]
  float* sB2 = sB1 + 2 * MAXD;
  float* sS = sB2 + 2 * MAXD;      // L1OUT: S1/n, S2/n; later E, M
  float* sx = sS + 2 * MAXD;       // [MAXD][kLd]
  float* sy = sx + MAXD * kLd;
  float* sy2 = sy + MAXD * kLd;
  float* sh = sy2 + MAXD * kLd;    // [2][MAXD][kLd] hidden activations
  float* so = sh + 2 * MAXD * kLd; // [2][MAXD][kLd] outputs, then d(mu), d(pre-tanh logvar)
  float* sdh = so + 2 * MAXD * kLd;
  float* sred = sdh + 2 * MAXD * kLd;  // [kNT/32]
  __shared__ int s_last;

  const int t = threadIdx.x;
  const int B = p.B, Dx = p.Dx, H = p.H, Dy = p.Dy;
  const int row0 = blockIdx.x * kRows;
  const int nrows = min(kRows, B - row0);
  const float inv_n = 1.f / (float)B;

  // ---- phase 0: parameters, this CTA's rows, (L1OUT) column moments of y
  for (int net = 0; net < 2; ++net) {
    const float* W1 = p.w[net * 4], *b1 = p.w[net * 4 + 1], *W2 = p.w[net * 4 + 2], *b2 = p.w[net * 4 + 3];
    for (int i = t; i < H * Dx; i += kNT) sW1[net * MM + i] = __ldg(W1 + i);
    for (int i = t; i < Dy * H; i += kNT) sW2[net * MM + i] = __ldg(W2 + i);
    if (t < H) sB1[net * MAXD + t] = __ldg(b1 + t);
    if (t < Dy) sB2[net * MAXD + t] = __ldg(b2 + t);
  }
  if (t < 2 * MAXD) sS[t] = 0.f;
  for (int idx = t; idx < kRows * Dx; idx += kNT) {
    const int r = idx / Dx, i = idx - r * Dx;
    sx[i * kLd + r] = r < nrows ? __ldg(p.x + (long long)(row0 + r) * p.ldx + i) : 0.f;
  }
  for (int idx = t; idx < kRows * Dy; idx += kNT) {
    const int r = idx / Dy, d = idx - r * Dy;
    const bool v = r < nrows;
    sy[d * kLd + r] = v ? __ldg(p.y + (long long)(row0 + r) * p.ldy + d) : 0.f;
    if (p.mode == CLEARVAE_MI_CLUB) sy2[d * kLd + r] = v ? __ldg(p.y + __ldg(p.perm + row0 + r) * p.ldy + d) : 0.f;
  }
  __syncthreads();
  if (p.mode == CLEARVAE_MI_L1OUT) {
    // every CTA reduces the full column moments (B*Dy floats from L2): no second launch, no grid sync
    int dp = 1;
    while (dp < Dy) dp <<= 1;
    const int d = t & (dp - 1);
    float a = 0.f, b = 0.f;
    if (d < Dy)
      for (int r = t / dp; r < B; r += kNT / dp) {
        const float v = __ldg(p.y + (long long)r * p.ldy + d);
        a += v;
        b = fmaf(v, v, b);
      }
    if (d < Dy) { atomicAdd(&sS[d], a * inv_n); atomicAdd(&sS[MAXD + d], b * inv_n); }
    __syncthreads();
  }

  // ---- phase 1: both MLP forwards, one (network, row) per thread of the first 128
  if (t < 2 * kRows) {
    const int net = t / kRows, r = t - net * kRows;
    float xr[MAXD], hr[MAXD];
#pragma unroll
    for (int i = 0; i < MAXD; ++i) xr[i] = i < Dx ? sx[i * kLd + r] : 0.f;
#pragma unroll
    for (int j = 0; j < MAXD; ++j) {
      float a = 0.f;
      if (j < H) {
        a = sB1[net * MAXD + j];
        const float* w = sW1 + net * MM + j * Dx;
#pragma unroll
        for (int i = 0; i < MAXD; ++i)
          if (i < Dx) a = fmaf(w[i], xr[i], a);
        a = fmaxf(a, 0.f);
        sh[(net * MAXD + j) * kLd + r] = a;
      }
      hr[j] = a;
    }
    for (int d = 0; d < Dy; ++d) {
      float a = sB2[net * MAXD + d];
      const float* w = sW2 + net * MM + d * H;
#pragma unroll
      for (int j = 0; j < MAXD; ++j)
        if (j < H) a = fmaf(w[j], hr[j], a);
      so[(net * MAXD + d) * kLd + r] = a;
    }
  }
  __syncthreads();
  if (p.mode == CLEARVAE_MI_L1OUT && t < 2 * MAXD) {
    // keep S1/n, S2/n in registers of nobody: copy to the tail of sred is not needed; E/M accumulate in sdh scratch
  }

  // ---- phase 2: per-element loss and output gradients
  float* sE = sdh;  // L1OUT scratch [2][MAXD] (sdh is written in phase 3, after the next barrier)
  if (p.mode == CLEARVAE_MI_L1OUT && t < 2 * MAXD) sE[t] = 0.f;
  if (p.mode == CLEARVAE_MI_L1OUT) __syncthreads();
  float loss = 0.f;
  for (int idx = t; idx < kRows * Dy; idx += kNT) {
    const int d = idx / kRows, r = idx - d * kRows;
    const bool valid = r < nrows;
    const float mu = so[d * kLd + r];
    const float lv = tanhf(so[(MAXD + d) * kLd + r]);
    const float inv = expf(-lv);
    const float y = sy[d * kLd + r];
    const float diff = mu - y;
    float le, dmu, dlv;
    if (p.mode == CLEARVAE_MI_LEARN) {
      const float q = diff * diff * inv;
      le = q + lv;
      dmu = 2.f * inv_n * diff * inv;
      dlv = inv_n * (1.f - q);
    } else if (p.mode == CLEARVAE_MI_CLUB) {
      const float y2 = sy2[d * kLd + r];
      const float d2 = mu - y2;
      le = 0.5f * (d2 * d2 - diff * diff) * inv;
      dmu = inv_n * (y - y2) * inv;
      dlv = -inv_n * le;
      if (valid) {
        atomicAdd(p.dy + (long long)(row0 + r) * Dy + d, inv_n * diff * inv);
        atomicAdd(p.dy + __ldg(p.perm + row0 + r) * Dy + d, -inv_n * d2 * inv);
      }
    } else {
      const float s1 = sS[d], s2 = sS[MAXD + d];
      const float q = s2 - 2.f * mu * s1 + mu * mu;
      le = 0.5f * inv * (q - diff * diff);
      dmu = inv_n * inv * (y - s1);
      dlv = -inv_n * le;
      if (valid) {
        p.dy[(long long)(row0 + r) * Dy + d] = inv_n * diff * inv;
        atomicAdd(&sE[d], inv);
        atomicAdd(&sE[MAXD + d], inv * mu);
      }
    }
    if (!valid) { le = 0.f; dmu = 0.f; dlv = 0.f; }
    loss += le;
    so[d * kLd + r] = dmu;
    so[(MAXD + d) * kLd + r] = dlv * (1.f - lv * lv);
  }
  const float cta_loss = cv::block_sum<kNT>(loss, sred);  // contains __syncthreads(): so[] is complete afterwards
  float* part = p.partial + (long long)blockIdx.x * p.np;
  if (t == 0) part[0] = cta_loss;
  if (p.mode == CLEARVAE_MI_L1OUT && t < Dy) { part[1 + t] = sE[t]; part[1 + Dy + t] = sE[MAXD + t]; }
  __syncthreads();

  // ---- phase 3: back through the second linear layers and the ReLUs
  if (t < 2 * kRows) {
    const int net = t / kRows, r = t - net * kRows;
    float g[MAXD];
#pragma unroll
    for (int d = 0; d < MAXD; ++d) g[d] = d < Dy ? so[(net * MAXD + d) * kLd + r] : 0.f;
    for (int j = 0; j < H; ++j) {
      float a = 0.f;
      const float* w = sW2 + net * MM + j;
#pragma unroll
      for (int d = 0; d < MAXD; ++d)
        if (d < Dy) a = fmaf(w[d * H], g[d], a);
      sdh[(net * MAXD + j) * kLd + r] = sh[(net * MAXD + j) * kLd + r] > 0.f ? a : 0.f;
    }
  }
  __syncthreads();

  if (p.mode != CLEARVAE_MI_LEARN) {
    // gradient w.r.t. x: both first layers transposed
    for (int idx = t; idx < kRows * Dx; idx += kNT) {
      const int r = idx / Dx, i = idx - r * Dx;
      if (r >= nrows) continue;
      float a = 0.f;
      for (int net = 0; net < 2; ++net)
        for (int j = 0; j < H; ++j) a = fmaf(sW1[net * MM + j * Dx + i], sdh[(net * MAXD + j) * kLd + r], a);
      p.dx[(long long)(row0 + r) * Dx + i] = a;
    }
  } else {
    // parameter gradients: per-CTA outer products over this CTA's rows (rows past the batch carry zeros)
    const int Pn = H * Dx + H + Dy * H + Dy;
    for (int e = t; e < 2 * Pn; e += kNT) {
      const int net = e / Pn;
      int q = e - net * Pn;
      const float *a, *b = nullptr;
      if (q < H * Dx) {
        const int j = q / Dx, i = q - j * Dx;
        a = sdh + (net * MAXD + j) * kLd;
        b = sx + i * kLd;
      } else if ((q -= H * Dx) < H) {
        a = sdh + (net * MAXD + q) * kLd;
      } else if ((q -= H) < Dy * H) {
        const int d = q / H, j = q - d * H;
        a = so + (net * MAXD + d) * kLd;
        b = sh + (net * MAXD + j) * kLd;
      } else {
        q -= Dy * H;
        a = so + (net * MAXD + q) * kLd;
      }
      float acc = 0.f;
      if (b != nullptr) {
#pragma unroll 8
        for (int r = 0; r < kRows; ++r) acc = fmaf(a[r], b[r], acc);
      } else {
#pragma unroll 8
        for (int r = 0; r < kRows; ++r) acc += a[r];
      }
      part[1 + e] = acc;
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
    if (e == 0) s = s * inv_n - (p.mode == CLEARVAE_MI_L1OUT ? p.l1_const : 0.f);
    p.out[e] = s;
  }
  if (t == 0) *p.counter = 0u;
}

// gx = g * dx_unit; gy = g * (dy_unit [+ (y * E_d - M_d) / B^2 for L1OUT])
__global__ void mi_bound_bwd_kernel(int mode, const float* __restrict__ g, const float* __restrict__ dxu,
                                    const float* __restrict__ dyu, const float* __restrict__ y, long long ldy,
                                    const float* __restrict__ em, int B, int Dx, int Dy, float* __restrict__ gx,
                                    float* __restrict__ gy) {
  const float gv = __ldg(g);
  const float inv_n2 = 1.f / ((float)B * (float)B);
  const int nx = B * Dx, ny = B * Dy;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx + ny; i += gridDim.x * blockDim.x) {
    if (i < nx) {
      gx[i] = gv * dxu[i];
    } else {
      const int k = i - nx;
      float v = dyu[k];
      if (mode == CLEARVAE_MI_L1OUT) {
        const int row = k / Dy, d = k - row * Dy;
        v += inv_n2 * (y[(long long)row * ldy + d] * em[d] - em[Dy + d]);
      }
      gy[k] = gv * v;
    }
  }
}

template <int MAXD>
constexpr size_t mi_smem_bytes() {
  return (size_t)(4 * MAXD * MAXD + 6 * MAXD + 9 * MAXD * kLd + kNT / 32) * sizeof(float);
}

template <int MAXD>
int launch_mi(const MiParams& p, int grid, cudaStream_t st) {
  constexpr size_t smem = mi_smem_bytes<MAXD>();
  static bool attr_done = false;
  if (!attr_done && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mi_kernel<MAXD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  mi_kernel<MAXD><<<grid, kNT, smem, st>>>(p);
  CV_LAUNCH_CHECK();
  return 0;
}

inline int np_of(int mode, int Dx, int H, int Dy) {
  if (mode == CLEARVAE_MI_LEARN) return 1 + 2 * (H * Dx + H + Dy * H + Dy);
  if (mode == CLEARVAE_MI_L1OUT) return 1 + 2 * Dy;
  return 1;
}

}  // namespace

extern "C" {

size_t clearvae_mi_workspace_bytes(int32_t mode, int64_t B, int32_t Dx, int32_t H, int32_t Dy) {
  if (B <= 0) return 0;
  const long long grid = (B + kRows - 1) / kRows;
  return 256 + (size_t)grid * np_of(mode, Dx, H, Dy) * sizeof(float);
}

int clearvae_mi_estimator(int32_t mode, const float* x, int64_t ldx, const float* y, int64_t ldy, const int64_t* perm, int64_t B,
                          int32_t Dx, int32_t H, int32_t Dy, const float* const* params_host, float* out, float* dx_unit, float* dy_unit,
                          void* workspace, size_t workspace_bytes, void* stream) {
  if (!x || !y || !params_host || !out || !workspace || B <= 0 || ldx < Dx || ldy < Dy) return CLEARVAE_EINVAL;
  if (mode != CLEARVAE_MI_LEARN && mode != CLEARVAE_MI_CLUB && mode != CLEARVAE_MI_L1OUT) return CLEARVAE_EINVAL;
  if (mode == CLEARVAE_MI_CLUB && !perm) return CLEARVAE_EINVAL;
  if (mode != CLEARVAE_MI_LEARN && (!dx_unit || !dy_unit)) return CLEARVAE_EINVAL;
  if (mode == CLEARVAE_MI_L1OUT && B < 2) return CLEARVAE_EINVAL;
  if (Dx < 1 || H < 1 || Dy < 1 || Dx > 32 || H > 32 || Dy > 32 || B > (1 << 24)) return CLEARVAE_EUNSUPPORTED;
  for (int i = 0; i < 8; ++i)
    if (!params_host[i]) return CLEARVAE_EINVAL;
  if (workspace_bytes < clearvae_mi_workspace_bytes(mode, B, Dx, H, Dy)) return CLEARVAE_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  MiParams p{};
  p.x = x; p.y = y; p.perm = perm;
  p.B = (int)B; p.Dx = Dx; p.H = H; p.Dy = Dy;
  p.ldx = ldx; p.ldy = ldy;
  for (int i = 0; i < 8; ++i) p.w[i] = params_host[i];
  p.mode = mode;
  p.np = np_of(mode, Dx, H, Dy);
  p.out = out; p.dx = dx_unit; p.dy = dy_unit;
  p.counter = reinterpret_cast<unsigned int*>(workspace);  // zero on first use, reset by the last CTA of every launch
  p.partial = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + 256);
  p.l1_const = (float)log1p(exp(-20.0) / ((double)B - 1.0));
  if (mode == CLEARVAE_MI_CLUB) {
    cudaError_t e = cudaMemsetAsync(dy_unit, 0, (size_t)B * Dy * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
  }
  const int grid = (int)((B + kRows - 1) / kRows);
  const int md = Dx > H ? (Dx > Dy ? Dx : Dy) : (H > Dy ? H : Dy);
  if (md <= 8) return launch_mi<8>(p, grid, st);
  if (md <= 16) return launch_mi<16>(p, grid, st);
  return launch_mi<32>(p, grid, st);
}

int clearvae_mi_bound_bwd(int32_t mode, const float* grad_out, const float* dx_unit, const float* dy_unit, const float* y,
                          int64_t ldy, const float* out_fwd, int64_t B, int32_t Dx, int32_t Dy, float* gx, float* gy, void* stream) {
  if (!grad_out || !dx_unit || !dy_unit || !gx || !gy || B <= 0) return CLEARVAE_EINVAL;
  if (mode == CLEARVAE_MI_L1OUT && (!y || !out_fwd)) return CLEARVAE_EINVAL;
  const long long n = B * (long long)(Dx + Dy);
  const int grid = (int)std::min<long long>((n + 255) / 256, 148 * 4);
  mi_bound_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mode, grad_out, dx_unit, dy_unit, y, ldy, out_fwd ? out_fwd + 1 : nullptr,
                                                              (int)B, Dx, Dy, gx, gy);
  CV_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
