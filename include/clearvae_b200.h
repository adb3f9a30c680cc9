/*
 * clearvae_b200.h — C ABI of libclearvae_b200.so (sm_100a CUDA, no torch types).
 *
 * The reference (scotsun/clear-vae) has no FFI layer of its own: its hot path is
 * Python calling PyTorch.  Each entry point below replaces the
 * arithmetic of the cited reference lines; the torch custom ops in
 * clear_vae_b200/csrc/torch_binding.cpp are thin shims over exactly these
 * symbols, and INTEGRATION.md shows the ctypes stub a reference maintainer
 * would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - tensors are contiguous; fp32 unless stated; labels are int64;
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises, every
 *     launch is CUDA-graph capturable, no function allocates or frees;
 *   - return 0 on success, a negative CLEARVAE_E* on bad arguments, or a
 *     positive cudaError_t when a launch fails; no exceptions cross the ABI.
 */
#ifndef CLEARVAE_B200_H_
#define CLEARVAE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLEARVAE_OK 0
#define CLEARVAE_EINVAL (-1)      /* null pointer / negative size */
#define CLEARVAE_EUNSUPPORTED (-2) /* shape or mode not built */
#define CLEARVAE_EWORKSPACE (-3)   /* workspace too small */

/* similarity selector: reference `sim_fn` strings (losses.py:111-123) */
enum { CLEARVAE_SIM_COSINE = 0, CLEARVAE_SIM_L2 = 1, CLEARVAE_SIM_MODIFIED_L2 = 2,
       CLEARVAE_SIM_JEFFREY = 3, CLEARVAE_SIM_MAHALANOBIS = 4 };
/* row-loss selector: reference `loss_name` strings (losses.py:124,129-170) */
enum { CLEARVAE_LOSS_SNN = 0, CLEARVAE_LOSS_SUPCON_IN = 1, CLEARVAE_LOSS_SUPCON_OUT = 2 };

int clearvae_version(void);

/* ---------------------------------------------------------------------------
 * Latent-head loss block.
 *
 * One "term" = one contrastive_loss call of the reference
 * (losses.py:98-137: pair mask, pairwise similarity, snn_loss, finite-row
 * mean).  The block handles up to two terms (content, style) in one launch and
 * fuses, per term, the reparameterisation (vae.py:56-60) and the Gaussian KL
 * (losses.py:48-49).
 *
 * Row side = this rank's shard [B, D]; column side = the (all-gathered) global
 * batch [Bg, D].  Single GPU: cols == rows, Bg == B, row_offset == 0.
 * ------------------------------------------------------------------------- */
typedef struct {
  /* row side (local shard) */
  const float* mu;      /* [B, D]                                    */
  const float* logvar;  /* [B, D] or NULL (no KL / reparam)          */
  const float* eps;     /* [B, D] or NULL (no reparam)               */
  /* column side (global batch); NULL means "same as rows" */
  const float* mu_cols; /* [Bg, D]                                   */
  const float* logvar_cols; /* [Bg, D], only for logvar-dependent sims */
  /* outputs */
  float* z;             /* [B, z_stride] slice start, or NULL        */
  float* row_stats;     /* [B, 2]: (lse_all, lse_pos) minus the shared shift — DESIGN.md §3.1 */
  /* configuration */
  int32_t snn_enable;   /* 0: only reparam/KL for this term          */
  int32_t ps;           /* 0: same-label positives, 1: flipped mask  */
  float* row_aux;       /* [B] out, SupCon row losses only (NULL for snn_loss): n_k of losses.py:141 (supcon_in) or the
                           positive count of losses.py:164 (supcon_out); row_stats then holds (lse_all, b) with
                           b = lse_pos - log(n_k) resp. mean positive similarity, so the row loss is always a - b */
} clearvae_term_fwd;

/* scalars written by the forward (float[CLEARVAE_NSCALARS]) */
enum { CLEARVAE_S_KL0 = 0, CLEARVAE_S_KL1 = 1, CLEARVAE_S_LOSS0 = 2, CLEARVAE_S_LOSS1 = 3,
       CLEARVAE_S_SUM0 = 4, CLEARVAE_S_SUM1 = 5, CLEARVAE_S_CNT0 = 6, CLEARVAE_S_CNT1 = 7,
       CLEARVAE_NSCALARS = 8 };

size_t clearvae_latent_workspace_bytes(int64_t B, int64_t Bg, int32_t D, int32_t n_terms);

/* forward: replaces vae.py:56-60 (sample), losses.py:48-49 (KL), losses.py:98-137.
 * `finalize` != 0 : the last CTA also reduces row_stats -> LOSS/SUM/CNT scalars
 *                   (single GPU).  With finalize == 0 call clearvae_snn_finalize
 *                   on the gathered stats.  `workspace` must be zero-initialised
 *                   once; the kernels leave it zeroed. */
int clearvae_latent_fwd(const clearvae_term_fwd* terms_host, int32_t n_terms,
                        const int64_t* label_rows, const int64_t* label_cols,
                        int64_t B, int64_t Bg, int64_t row_offset, int32_t D, int32_t z_stride,
                        int32_t sim_fn, int32_t loss_name, float temperature,
                        float* scalars, int32_t finalize,
                        void* workspace, size_t workspace_bytes, void* stream);

/* reduce gathered row stats [Bg,2] of one term to (loss, sum, count); losses.py:125-126 */
int clearvae_snn_finalize(const float* row_stats_all, int64_t Bg, int32_t term,
                          float* scalars, void* stream);

typedef struct {
  const float* mu;       /* [B, D] */
  const float* logvar;   /* [B, D] or NULL */
  const float* eps;      /* [B, D] or NULL */
  const float* mu_cols;  /* [Bg, D] or NULL (= rows) */
  const float* row_stats_all; /* [Bg, 2] forward stats of ALL global rows */
  const float* dz;       /* [B, z_stride] slice start or NULL: grad wrt z */
  float* dmu;            /* [B, D] out */
  float* dlogvar;        /* [B, D] out or NULL */
  int32_t snn_enable;
  int32_t ps;
  const float* logvar_cols; /* [Bg, D] or NULL (= rows): only the logvar-dependent similarities read it */
  const float* row_aux_all; /* [Bg] row_aux of ALL global rows (SupCon row losses only, else NULL) */
} clearvae_term_bwd;

/* backward of the block. `gscal` (device, float[4]) = upstream grads of
 * (kl0, kl1, loss0, loss1); `scalars` = the forward's scalar buffer (CNT used). */
int clearvae_latent_bwd(const clearvae_term_bwd* terms_host, int32_t n_terms,
                        const int64_t* label_rows, const int64_t* label_cols,
                        int64_t B, int64_t Bg, int64_t row_offset, int32_t D, int32_t z_stride,
                        int32_t sim_fn, int32_t loss_name, float temperature,
                        const float* scalars, const float* gscal, void* stream);

/* Same backward with a workspace (zero-initialised once, left zeroed): lets the FFMA kernels split the column sweep over
 * several CTAs per row block when the local batch is small (a data-parallel shard against the gathered global batch);
 * partial sums are merged in a fixed order, so results stay bit-reproducible.  workspace == NULL: identical to
 * clearvae_latent_bwd. */
size_t clearvae_latent_bwd_workspace_bytes(int64_t B, int64_t Bg, int32_t D, int32_t n_terms);
int clearvae_latent_bwd_ws(const clearvae_term_bwd* terms_host, int32_t n_terms,
                           const int64_t* label_rows, const int64_t* label_cols,
                           int64_t B, int64_t Bg, int64_t row_offset, int32_t D, int32_t z_stride,
                           int32_t sim_fn, int32_t loss_name, float temperature,
                           const float* scalars, const float* gscal, void* workspace, size_t workspace_bytes, void* stream);

/* debug/test: positive/candidate sets of losses.py:107-110,131-135 as bytes
 * (bit0 = candidate j!=i, bit1 = positive), [B, Bg] — same index math as the kernels. */
int clearvae_pair_mask(const int64_t* label_rows, const int64_t* label_cols, int64_t B, int64_t Bg,
                       int64_t row_offset, int32_t ps, uint8_t* out, void* stream);

/* ---------------------------------------------------------------------------
 * Reconstruction term: replaces losses.py:36-47 (F.mse_loss(reduction="none")
 * summed per sample, batch mean) and its backward.  xhat / x are any contiguous
 * [B, per_sample] fp32 views, 16-byte aligned.
 * ------------------------------------------------------------------------- */
size_t clearvae_recon_workspace_bytes(void);
int clearvae_recon_fwd(const float* xhat, const float* x, int64_t B, int64_t per_sample, float* out,
                       void* workspace, size_t workspace_bytes, void* stream);
/* dxhat = 2 (xhat - x) / B * (*grad_out);  grad_out is a device scalar */
int clearvae_recon_bwd(const float* xhat, const float* x, const float* grad_out, int64_t B, int64_t per_sample,
                       float* dxhat, void* stream);

/* ---------------------------------------------------------------------------
 * Convolution-shaped GEMMs (encoder / decoder layers and the linear heads).
 *
 * Replaces nn.Conv2d / nn.ConvTranspose2d / nn.Linear forward and data-gradient
 * as the reference stacks them at vae.py:15-46 and vae.py:113-156, with the
 * preceding BatchNorm-apply + ReLU folded into the operand load ("pre-op") and
 * the following BatchNorm's batch statistics accumulated in the epilogue.
 * Tensor-core path: bf16 operands, fp32 accumulation in TMEM (tcgen05), weights
 * staged by TMA from a packed K-major copy.
 * ------------------------------------------------------------------------- */
typedef struct {
  int32_t transposed;  /* 0: Conv2d [Cout,Cin,k,k]; 1: ConvTranspose2d [Cin,Cout,k,k] */
  int32_t k, stride, pad, out_pad;
  int32_t Cin, Cout, Hin, Win;   /* a Linear(in,out) is k=1,stride=1,pad=0,Hin=Win=1 */
} clearvae_conv_geom;

/* strided view of a 4-D activation, element (n, h, w, c) at ptr + n*sn + h*sh + w*sw + c*sc (in elements) */
typedef struct {
  void* ptr;
  int64_t sn, sh, sw, sc;
  int32_t dtype;  /* CLEARVAE_F32 or CLEARVAE_BF16 */
} clearvae_tensor4;
enum { CLEARVAE_F32 = 0, CLEARVAE_BF16 = 1 };
enum { CLEARVAE_ROLE_FPROP = 0, CLEARVAE_ROLE_DGRAD = 1 };
/* OR into `role` of clearvae_conv_packed_weight_bytes / clearvae_conv_pack_weight / clearvae_conv_gemm: fp32-grade products on
 * the bf16 tensor cores.  Each fp32 operand x is split into three bf16 parts, x = p0 + p1 + p2 (p0 = bf16(x), p1 = bf16(x - p0),
 * p2 = bf16(x - p0 - p1): 24 significand bits), and the six part products down to 2^-24 |a||w| — (0,0) (0,1) (1,0) (1,1) (0,2)
 * (2,0) — are accumulated in fp32 (TMEM): the accuracy class of the reference's own fp32 run (cudnn.allow_tf32 = False) at
 * six times the MMA work.  The packed weight holds [p0 | p1 | p2] (three times the bytes, parts 128-byte aligned). */
#define CLEARVAE_ROLE_SPLIT3 16
/* epilogue of clearvae_conv_gemm */
enum { CLEARVAE_EPI_BIAS_STATS = 0,  /* dst = acc + bias; stats += (sum v, sum v^2) per output channel      */
       CLEARVAE_EPI_MASK_STATS = 1   /* dst = g = acc * [mask_src*mask_scale+mask_shift > 0];
                                        stats += (sum g, sum g*mask_src): ReLU + BatchNorm backward sums    */ };

size_t clearvae_conv_packed_weight_bytes(const clearvae_conv_geom* g, int32_t role);
/* fp32 reference-layout weight -> bf16 packed GEMM operand(s) of (geometry, role) */
int clearvae_conv_pack_weight(const clearvae_conv_geom* g, int32_t role, const float* weight, void* packed, void* stream);

/* All packed operands of a model refreshed in ONE launch (after the optimiser step, for the next step): `build` is pure host
 * code that fills a table (clearvae_conv_pack_multi_table_bytes(n) bytes) describing n (geometry, role, weight, packed buffer)
 * tuples; the caller copies the table to the device once (weights and packed buffers are persistent) and replays `launch`.
 * Split-mode roles are not accepted (their packs stay per weight). */
size_t clearvae_conv_pack_multi_table_bytes(int32_t n_weights);
int clearvae_conv_pack_multi_build(int32_t n_weights, const clearvae_conv_geom* geoms_host, const int32_t* roles_host,
                                   const float* const* weights, void* const* packed, void* table_host, int32_t* n_entries,
                                   int32_t* n_blocks);
int clearvae_conv_pack_multi_launch(const void* table_device, int32_t n_entries, int32_t n_blocks, void* stream);

int clearvae_conv_gemm(const clearvae_conv_geom* g, int32_t role, int64_t batch,
                       const clearvae_tensor4* src, const float* pre_scale, const float* pre_shift, int32_t pre_relu,
                       const void* packed_weight, const float* bias, const clearvae_tensor4* dst, int32_t epilogue,
                       const clearvae_tensor4* mask_src, const float* mask_scale, const float* mask_shift,
                       double* stats, void* stream);

/* weight gradient of the layer (replaces autograd's conv / conv-transpose / linear wgrad):
 * dweight (fp32, reference layout, must be zero or hold the value to accumulate into) +=
 *   sum over pixels of pre(src)[pixel @ tap, cin] * dy[pixel, cout].  `dy` is the gradient
 *   w.r.t. the layer's raw output, indexed like the forward dst. */
int clearvae_conv_wgrad(const clearvae_conv_geom* g, int64_t batch, const clearvae_tensor4* src,
                        const float* pre_scale, const float* pre_shift, int32_t pre_relu,
                        const clearvae_tensor4* dy, float* dweight, void* stream);

/* the same weight gradient in split mode (see CLEARVAE_ROLE_SPLIT3): both operands in three bf16 parts, six products per pixel block */
int clearvae_conv_wgrad_split3(const clearvae_conv_geom* g, int64_t batch, const clearvae_tensor4* src,
                               const float* pre_scale, const float* pre_shift, int32_t pre_relu,
                               const clearvae_tensor4* dy, float* dweight, void* stream);

/* ---------------------------------------------------------------------------
 * Train-mode BatchNorm around the GEMMs (nn.BatchNorm1d/2d at vae.py:17-44,115-154).
 * Flat tensors; channel(idx) = (idx / inner) % C  (inner = 1 for NHWC, H*W for NCHW).
 * `stats` is a [2][C*group] double accumulator filled by GEMM epilogues or
 * clearvae_bn_reduce; finalize / coef consume and clear it.
 * ------------------------------------------------------------------------- */
int clearvae_bn_finalize(double* stats, int32_t C, int32_t group, double count, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, float momentum, float eps, float* scale, float* shift,
                         int32_t expand, float* save_mean, float* save_invstd, int32_t repeat /* momentum updates applied */, void* stream);
/* mode 0: (sum y, sum y^2); mode 1: (sum g', sum g'*y) with g' = g*[act>0] when act != NULL */
int clearvae_bn_reduce(const void* y, int32_t y_dtype, const void* g, int32_t g_dtype, const void* act, int32_t act_dtype,
                       const float* mask_scale, const float* mask_shift /* or: g' = g*[y*scale+shift > 0] */, int64_t total,
                       int32_t C, int64_t inner, int32_t mode, double* stats, void* stream);
size_t clearvae_bn_act_workspace_bytes(void);
/* out = act(raw*scale[ch]+shift[ch]); act 0 none / 1 relu / 2 sigmoid; with `target`, *sse_out = sum((out-target)^2)/batch */
int clearvae_bn_act_fwd(const void* raw, int32_t raw_dtype, const float* scale, const float* shift, int64_t total, int32_t C,
                        int64_t inner, int32_t act, int32_t to_nhwc_C, int32_t to_nhwc_HW /* >0: write [b][hw][c] */, void* out,
                        int32_t out_dtype, const float* target, int64_t batch,
                        float* sse_out, void* workspace, size_t workspace_bytes, void* stream);
/* backward of xhat = sigmoid(bn(raw)) under recon = sum((xhat-x)^2)/batch (+ optional external grad on xhat) */
int clearvae_sigmoid_mse_bwd(const float* xhat, const float* x, const float* grad_recon, const float* grad_ext, const void* raw,
                             int32_t raw_dtype, int64_t total, int32_t C, int64_t inner, int64_t batch, float* g_pre,
                             double* stats, void* stream);
/* coef [3][C]: dy = coef0[ch]*g + coef1[ch]*y + coef2[ch]; also BatchNorm weight / bias gradients */
int clearvae_bn_bwd_coef(double* stats, int32_t C, int32_t group, double count, const float* gamma, const float* save_mean,
                         const float* save_invstd, float* coef, float* dgamma, float* dbeta, void* stream);
int clearvae_bn_bwd_apply(const void* g, int32_t g_dtype, const void* y, int32_t y_dtype, const void* act, int32_t act_dtype,
                          const float* mask_scale, const float* mask_shift, const float* coef, int64_t total, int32_t C,
                          int64_t inner, int32_t to_nhwc /* write [b][hw][c] instead of [b][c][hw] */, void* dy,
                          int32_t dy_dtype, void* stream);
/* clearvae_bn_finalize + BatchNorm-apply + ReLU in one launch (train-mode nn.BatchNorm2d/1d + nn.ReLU, vae.py:17-18,34-35).
 * `stats` must hold 2*C moments followed by one zero-initialised ticket word (2*C + 1 doubles); it is cleared on exit.
 * layout 0: raw bf16 channels-last [.., C]; 1: raw bf16 channel-major [B, C*HW]; both -> act bf16 in the same layout;
 * layout 2: raw fp32 [B, C], C = C0*HW in (c0, hw) order -> act bf16 [B, HW, C0] (decoder fc block);
 * layout 3: raw bf16 channels-last [B, HW, C] -> act bf16 AND a copy of raw (raw_cm_bf16), both channel-major [B, C*HW]
 *           (last encoder block: the reference's Flatten order, vae.py:26).  total = number of raw elements. */
int clearvae_bn_finalize_apply(double* stats, int32_t C, double count, const float* gamma, const float* beta, float* running_mean,
                               float* running_var, float momentum, float eps, int32_t repeat, float* scale, float* shift,
                               int32_t expand, float* save_mean, float* save_invstd, const void* raw, int32_t layout, int64_t total,
                               int32_t HW, void* act_bf16, void* raw_cm_bf16, void* stream);
/* act = relu(raw * scale[c] + shift[c]) on a channels-last bf16 tensor, C % 8 == 0: BatchNorm-apply + ReLU
 * (vae.py:17-18 etc.) materialised once between two GEMMs so their operand loads are plain async copies */
int clearvae_bn_relu_apply(const void* raw_bf16, const float* scale, const float* shift, int64_t total, int32_t C, void* act_bf16,
                           void* stream);
/* out[c] = sum_r x[r][c]  (bias gradients of the linear heads) */
int clearvae_colsum(const float* x, int64_t rows, int32_t cols, float* out, void* stream);

/* Direct (CUDA-core) forward of the two boundary layers whose shapes do not suit 128-row tensor-core tiles:
 * Conv2d(Cin<=4 -> 32) on an NCHW fp32 image (vae.py:16,114) and ConvTranspose2d(32 -> Cout<=4) to an NCHW
 * fp32 image (vae.py:43,153); k in {3,4}, stride 2, pad 1.  `weight` is the fp32 reference-layout tensor.
 * clearvae_conv_direct_supported returns 1 when (geometry, views) match, else callers use clearvae_conv_gemm. */
int clearvae_conv_direct_supported(const clearvae_conv_geom* g, const clearvae_tensor4* src, const clearvae_tensor4* dst);
int clearvae_conv_direct_fwd(const clearvae_conv_geom* g, int64_t batch, const clearvae_tensor4* src, const float* pre_scale,
                             const float* pre_shift, int32_t pre_relu, const float* weight, const float* bias,
                             const clearvae_tensor4* dst, double* stats, void* stream);

/* weight gradient of the same two boundary layers on CUDA cores (fp32 accumulate, fp32 atomics into `dweight`, which the
 * caller zero-fills): src / dy as in clearvae_conv_wgrad — the 32-channel side channels-last bf16, the <=4-channel side
 * NCHW fp32 or bf16.  Returns CLEARVAE_EUNSUPPORTED for any other layer (use clearvae_conv_wgrad). */
int clearvae_conv_direct_wgrad(const clearvae_conv_geom* g, int64_t batch, const clearvae_tensor4* src, const clearvae_tensor4* dy,
                               float* dweight, void* stream);

/* data gradient of the last decoder layer ConvTranspose2d(32 -> C<=4, stride 2, pad 1) (vae.py:43,153): it is the
 * Conv2d(C -> 32) of dy with the same weight tensor, so it runs on the direct first-layer kernel, with the previous block's
 * ReLU mask (mask_src * mask_scale + mask_shift > 0) and the BatchNorm-backward sums (sum g, sum g*y) in the epilogue.
 * dy: NCHW fp32|bf16; dst / mask_src: channels-last [B, Hin, Win, 32]. */
int clearvae_conv_direct_dgrad(const clearvae_conv_geom* g, int64_t batch, const clearvae_tensor4* dy, const float* weight,
                               const clearvae_tensor4* dst, const clearvae_tensor4* mask_src, const float* mask_scale,
                               const float* mask_shift, double* stats, void* stream);

/* decoder fc layer out = z W^T + b (nn.Linear(2D, 2048), vae.py:33,137) in fp32 on CUDA cores, K = 2D <= 64, K % 4 == 0;
 * `stats` (optional, 2*N doubles) accumulates the BatchNorm1d batch moments (sum, sum of squares) of every column */
int clearvae_fc_fwd(const float* z, const float* weight, const float* bias, float* out, double* stats, int64_t B, int32_t K,
                    int32_t N, void* stream);

/* ---------------------------------------------------------------------------
 * Variational MI estimators of CLEAR-MIM (mi_estimator.py:108-198): Gaussian heads
 *   p_mu = Linear(Dx,H)-ReLU-Linear(H,Dy),  p_logvar = Linear(Dx,H)-ReLU-Linear(H,Dy)-Tanh   (Dx, H, Dy <= 32).
 * `params_host` = 8 device pointers in nn.Module.parameters() order
 *   (p_mu.0.weight [H,Dx], p_mu.0.bias, p_mu.2.weight [Dy,H], p_mu.2.bias, p_logvar.0.weight, ...).
 * One launch runs forward and backward of both MLPs:
 *   CLEARVAE_MI_LEARN : out[0] = learning_loss(x, y) (mi_estimator.py:129-131,145-146), out[1..] = its gradient
 *                       w.r.t. the 8 parameters, flattened in the order above;
 *   CLEARVAE_MI_CLUB  : out[0] = CLUBSample.forward(x, y) with permutation `perm` (int64 [B], mi_estimator.py:133-143);
 *   CLEARVAE_MI_L1OUT : out[0] = L1OutUB.forward(x, y) as the reference executes it (mi_estimator.py:170-191),
 *                       out[1 .. 1+2*Dy) = column sums the backward needs;
 *   the bound modes also write the unit gradients d out[0] / dx, d out[0] / dy (direct part) to dx_unit / dy_unit;
 *   clearvae_mi_bound_bwd scales them by the incoming gradient (device scalar) and completes dy for L1OUT.
 * x / y are [B, Dx] / [B, Dy] views with row strides ldx / ldy (elements): column slices of one latent tensor work in place.
 * `workspace` must be zero-initialised once (its first word is a self-resetting ticket counter).
 * ------------------------------------------------------------------------- */
enum { CLEARVAE_MI_LEARN = 0, CLEARVAE_MI_CLUB = 1, CLEARVAE_MI_L1OUT = 2 };
size_t clearvae_mi_workspace_bytes(int32_t mode, int64_t B, int32_t Dx, int32_t H, int32_t Dy);
int clearvae_mi_estimator(int32_t mode, const float* x, int64_t ldx, const float* y, int64_t ldy, const int64_t* perm, int64_t B,
                          int32_t Dx, int32_t H, int32_t Dy, const float* const* params_host, float* out, float* dx_unit, float* dy_unit,
                          void* workspace, size_t workspace_bytes, void* stream);
int clearvae_mi_bound_bwd(int32_t mode, const float* grad_out, const float* dx_unit, const float* dy_unit, const float* y,
                          int64_t ldy, const float* out_fwd, int64_t B, int32_t Dx, int32_t Dy, float* gx, float* gy, void* stream);

/* ---------------------------------------------------------------------------
 * Density-ratio total-correlation term of CLEAR-TC-VAE: factor_cls = Linear(Z,Z)-ReLU-Linear(Z,1)-Sigmoid
 * (trainer_utils.py:133-138), Z even, Z <= 64.  w1 [Z,Z], b1 [Z], w2 [Z] (the [1,Z] weight), b2 [1]; z [B, Z] with row stride ldz.
 *   CLEARVAE_TC_BOUND : out[0] = mean relu(log(d / (1 - d))), d = factor_cls(z) (trainer.py:664-665); dz_unit [B,Z] = d out[0] / dz
 *   CLEARVAE_TC_DISC  : out[0] = BCELoss(cat[factor_cls(z), factor_cls([z_c | roll(z_s, -1, 0)])], cat[1, 0])
 *                       (trainer.py:573-587, 683-694), out[1..] = its gradient w.r.t. (w1, b1, w2, b2), flattened in that order
 * `workspace` zero-initialised once (self-resetting ticket).  clearvae_scale: y = (*grad_out) * x  (device scalar).
 * ------------------------------------------------------------------------- */
enum { CLEARVAE_TC_BOUND = 0, CLEARVAE_TC_DISC = 1 };
size_t clearvae_tc_workspace_bytes(int32_t mode, int64_t B, int32_t Z);
int clearvae_tc_factor(int32_t mode, const float* z, int64_t ldz, int64_t B, int32_t Z, const float* w1, const float* b1,
                       const float* w2, const float* b2, float* out, float* dz_unit, void* workspace, size_t workspace_bytes,
                       void* stream);
int clearvae_scale(const float* grad_out, const float* x, int64_t n, float* y, void* stream);

/* ---------------------------------------------------------------------------
 * Fused multi-tensor Adam: torch.optim.Adam defaults as built by the reference factories
 * (trainer_utils.py:100,139-140,178-181; no weight decay, no amsgrad).  The *_host arrays hold n_tensors
 * device pointers; `steps` = n_steps device floats holding the common step count (all advanced by one),
 * `counter` = one zero-initialised device word (self-resetting); gradients are multiplied by grad_scale.
 * ------------------------------------------------------------------------- */
#define CLEARVAE_ADAM_MAX_TENSORS 64 /* per launch; larger lists are split internally */
int clearvae_adam_step(int32_t n_tensors, float* const* params_host, const float* const* grads_host, float* const* exp_avg_host,
                       float* const* exp_avg_sq_host, const int64_t* numel_host, float* steps, int32_t n_steps,
                       unsigned int* counter, float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);

/* ---------------------------------------------------------------------------
 * Several reparameterisation draws of one (mu, logvar) pair per head in one launch (vae.py:56-60 as called five times
 * by CLEAR-MIM's inner loop, trainer.py:874-888): z_host[j] = [B, heads*D] with
 * z[j][b, h*D + d] = mu[h][b,d] + eps[j*heads + h][b,d] * exp(logvar[h][b,d] / 2).  The *_host arrays hold device pointers.
 * ------------------------------------------------------------------------- */
#define CLEARVAE_REPARAM_MAX_DRAWS 8
int clearvae_reparam_multi(int32_t heads, int32_t draws, const float* const* mu_host, const float* const* logvar_host,
                           const float* const* eps_host, float* const* z_host, int64_t B, int32_t D, void* stream);

/* ---------------------------------------------------------------------------
 * Group evidence of the ML-VAE / GVAE baselines (models/vae.py:159-223): `accumulate_group_evidence` and
 * `groupwise_reparam_each` as segmented reductions over label groups.  group_id[i] in [0, G) is the rank of row i's label among
 * the sorted unique labels of the batch (what `label.unique(sorted=True)` enumerates at vae.py:163); D <= 32.
 *   fwd:  MLVAE  mu_g = sum_i mu_i e^{-lv_i} / sum_i e^{-lv_i},  lv_g = -LSE_i(-lv_i)          (vae.py:174-180)
 *         GVAE   mu_g = mean_i mu_i,                              lv_g = LSE_i(lv_i) - log n_g   (vae.py:181-186)
 *         count[g] = n_g (float)
 *   bwd:  (dmu_grp, dlogvar_grp) [G, D] -> (dmu, dlogvar) [B, D], elementwise in the rows
 *   reparam fwd:  z_i = mu_g(i) + eps_i exp(lv_g(i) / 2)                                         (vae.py:196-209)
 *   reparam bwd:  dmu_grp[g] = sum_{i in g} dz_i,  dz_eps_grp[g] = sum_{i in g} dz_i eps_i  (d lv_g = dz_eps_grp * exp(lv_g / 2) / 2)
 * ------------------------------------------------------------------------- */
enum { CLEARVAE_GROUP_MLVAE = 0, CLEARVAE_GROUP_GVAE = 1 };
int clearvae_group_evidence_fwd(int32_t mode, const float* mu, const float* logvar, const int64_t* group_id, int64_t B, int32_t D,
                                int32_t G, float* mu_grp, float* logvar_grp, float* count, void* stream);
int clearvae_group_evidence_bwd(int32_t mode, const float* mu, const float* logvar, const int64_t* group_id, const float* mu_grp,
                                const float* logvar_grp, const float* count, const float* dmu_grp, const float* dlogvar_grp, int64_t B,
                                int32_t D, float* dmu, float* dlogvar, void* stream);
int clearvae_group_reparam_fwd(const float* mu_grp, const float* logvar_grp, const float* eps, const int64_t* group_id, int64_t B, int32_t D,
                               float* z, void* stream);
int clearvae_group_reparam_bwd(const float* dz, const float* eps, const int64_t* group_id, int64_t B, int32_t D, int32_t G, float* dmu_grp,
                               float* dz_eps_grp, void* stream);

/* ---------------------------------------------------------------------------
 * One-shot collectives over NVLink peer memory (data-parallel step).  They replace the small NCCL
 * exchanges a data-parallel port of the reference loop would issue per step (trainer.py:446-492 under DDP semantics):
 * the all-gather of the similarity operands / labels / row statistics / estimator latents and the parameter-gradient
 * all-reduce.  Every rank owns one buffer of `buffer_bytes` (identical on all ranks): a 4 KB header (arrival flags
 * written by the peers, the device-side call counter) followed by two slots used alternately.  `bases_host[r]` is rank
 * r's buffer as mapped into THIS process (own pointer at [rank]).  All ranks must issue the same sequence of calls on
 * one stream; launches are CUDA-graph capturable.  Setup-time helpers (alloc / export / open / close / error) are the
 * only entry points of the library that allocate or synchronise.
 *   clearvae_peer_gather:    piece k of `bytes_host[k]` bytes (multiple of 4) per rank -> dst_host[k] = [world][bytes] .
 *   clearvae_peer_allreduce: every tensor <- sum over ranks, fixed rank order (bit-identical on all ranks), in place.
 *   clearvae_peer_error:     0 = no poll ever timed out (20 s); otherwise 1 + the rank that was missing.
 * ------------------------------------------------------------------------- */
#define CLEARVAE_PEER_MAX_RANKS 8
#define CLEARVAE_PEER_MAX_PIECES 8
#define CLEARVAE_PEER_HEADER_BYTES 4096
#define CLEARVAE_PEER_HANDLE_BYTES 64
int clearvae_peer_alloc(int64_t bytes, void** ptr);
int clearvae_peer_free(void* ptr);
int clearvae_peer_export(void* ptr, uint8_t* handle_host);
int clearvae_peer_open(const uint8_t* handle_host, void** ptr);
int clearvae_peer_close(void* ptr);
int clearvae_peer_error(const void* local_base, int32_t* err_host);
int clearvae_peer_gather(void* const* bases_host, int32_t world, int32_t rank, int64_t buffer_bytes, int32_t n_pieces,
                         const void* const* src_host, void* const* dst_host, const int64_t* bytes_host, void* stream);
int clearvae_peer_allreduce(void* const* bases_host, int32_t world, int32_t rank, int64_t buffer_bytes, int32_t n_tensors,
                            float* const* tensors_host, const int64_t* numel_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLEARVAE_B200_H_ */
