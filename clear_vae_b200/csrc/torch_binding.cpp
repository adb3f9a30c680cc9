// Thin torch custom-op shim over the C ABI of libclearvae_b200.so.
//
// Ops live in the `clearvae` namespace and are registered for the CUDA
// dispatch key ONLY: calling them with CPU tensors raises (no CPU fallback by
// design).  The shim checks device / dtype / contiguity, allocates outputs from
// the caching allocator, launches on the current stream and never synchronises.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>

#include <vector>

#include "clearvae_b200.h"

namespace {

using at::Tensor;
using OptTensor = std::optional<Tensor>;

void check_f32(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), "clearvae: ", name, " must be a CUDA tensor");
  TORCH_CHECK(t.scalar_type() == at::kFloat, "clearvae: ", name, " must be float32");
  TORCH_CHECK(t.is_contiguous(), "clearvae: ", name, " must be contiguous");
}
void check_i64(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), "clearvae: ", name, " must be a CUDA tensor");
  TORCH_CHECK(t.scalar_type() == at::kLong, "clearvae: ", name, " must be int64");
  TORCH_CHECK(t.is_contiguous(), "clearvae: ", name, " must be contiguous");
}
const float* fptr(const OptTensor& t, const char* name, int64_t rows, int64_t cols) {
  if (!t.has_value() || !t->defined()) return nullptr;
  check_f32(*t, name);
  TORCH_CHECK(t->dim() == 2 && t->size(0) == rows && t->size(1) == cols, "clearvae: bad shape for ", name);
  return t->data_ptr<float>();
}
void check_rc(int rc, const char* what) {
  TORCH_CHECK(rc == 0, "clearvae: ", what, " failed with code ", rc,
              rc > 0 ? " (cudaError)" : rc == CLEARVAE_EUNSUPPORTED ? " (unsupported configuration)"
                                                                   : rc == CLEARVAE_EWORKSPACE ? " (workspace)" : " (invalid argument)");
}
void* cur_stream() { return (void*)at::cuda::getCurrentCUDAStream().stream(); }

// ---------------------------------------------------------------------------
std::tuple<Tensor, Tensor, std::vector<Tensor>> latent_fwd(
    at::TensorList mu, const c10::List<OptTensor>& logvar, const c10::List<OptTensor>& eps,
    const c10::List<OptTensor>& mu_cols, const c10::List<OptTensor>& logvar_cols, const Tensor& label_rows,
    const OptTensor& label_cols, at::IntArrayRef snn, at::IntArrayRef ps, int64_t row_offset, int64_t sim_fn,
    int64_t loss_name, double tau, bool finalize, bool want_z, Tensor workspace) {
  const int n = (int)mu.size();
  TORCH_CHECK(n >= 1 && n <= 2, "clearvae: 1 or 2 terms");
  TORCH_CHECK((int)logvar.size() == n && (int)eps.size() == n && (int)mu_cols.size() == n && (int)snn.size() == n &&
                  (int)ps.size() == n, "clearvae: per-term argument lists must have equal length");
  check_f32(mu[0], "mu");
  const c10::cuda::CUDAGuard guard(mu[0].device());
  const int64_t B = mu[0].size(0), D = mu[0].size(1);
  check_i64(label_rows, "label");
  TORCH_CHECK(label_rows.numel() == B, "clearvae: label must have B entries");
  int64_t Bg = B;
  const int64_t* lab_c = nullptr;
  if (label_cols.has_value() && label_cols->defined()) {
    check_i64(*label_cols, "label_cols");
    Bg = label_cols->numel();
    lab_c = label_cols->data_ptr<int64_t>();
  }
  auto fopt = mu[0].options();
  Tensor z = want_z ? at::empty({B, n * D}, fopt) : Tensor();
  Tensor scalars = at::zeros({CLEARVAE_NSCALARS}, fopt);
  std::vector<Tensor> stats, aux;
  const bool supcon = loss_name != CLEARVAE_LOSS_SNN;   // SupCon row losses carry one more per-row statistic
  clearvae_term_fwd terms[2];
  for (int i = 0; i < n; ++i) {
    check_f32(mu[i], "mu");
    TORCH_CHECK(mu[i].dim() == 2 && mu[i].size(0) == B && mu[i].size(1) == D, "clearvae: mu shapes differ");
    terms[i].mu = mu[i].data_ptr<float>();
    terms[i].logvar = fptr(logvar.get(i), "logvar", B, D);
    terms[i].eps = fptr(eps.get(i), "eps", B, D);
    terms[i].mu_cols = fptr(mu_cols.get(i), "mu_cols", Bg, D);
    terms[i].logvar_cols = (int)logvar_cols.size() == n ? fptr(logvar_cols.get(i), "logvar_cols", Bg, D) : nullptr;
    terms[i].z = want_z ? z.data_ptr<float>() + i * D : nullptr;
    terms[i].snn_enable = (int32_t)snn[i];
    terms[i].ps = (int32_t)ps[i];
    stats.push_back(snn[i] ? at::empty({B, 2}, fopt) : at::empty({0, 2}, fopt));
    terms[i].row_stats = snn[i] ? stats.back().data_ptr<float>() : nullptr;
    terms[i].row_aux = nullptr;
    if (supcon) {
      aux.push_back(snn[i] ? at::empty({B}, fopt) : at::empty({0}, fopt));
      terms[i].row_aux = snn[i] ? aux.back().data_ptr<float>() : nullptr;
    }
  }
  TORCH_CHECK(workspace.is_cuda() && workspace.is_contiguous(), "clearvae: workspace must be a contiguous CUDA tensor");
  check_rc(clearvae_latent_fwd(terms, n, label_rows.data_ptr<int64_t>(), lab_c, B, Bg, row_offset, (int32_t)D,
                               (int32_t)(n * D), (int32_t)sim_fn, (int32_t)loss_name, (float)tau,
                               scalars.data_ptr<float>(), finalize ? 1 : 0, workspace.data_ptr(),
                               (size_t)workspace.nbytes(), cur_stream()),
           "latent_fwd");
  if (!want_z) z = at::empty({0}, fopt);
  for (auto& a : aux) stats.push_back(a);   // [stats_0 .. stats_{n-1}, aux_0 .. aux_{n-1}] for the SupCon row losses
  return {z, scalars, stats};
}

void snn_finalize(const Tensor& stats_all, int64_t term, Tensor scalars) {
  check_f32(stats_all, "stats_all");
  check_f32(scalars, "scalars");
  const c10::cuda::CUDAGuard guard(stats_all.device());
  check_rc(clearvae_snn_finalize(stats_all.data_ptr<float>(), stats_all.size(0), (int32_t)term,
                                 scalars.data_ptr<float>(), cur_stream()),
           "snn_finalize");
}

std::tuple<std::vector<Tensor>, std::vector<Tensor>> latent_bwd(
    at::TensorList mu, const c10::List<OptTensor>& logvar, const c10::List<OptTensor>& eps,
    const c10::List<OptTensor>& mu_cols, const c10::List<OptTensor>& logvar_cols, const c10::List<OptTensor>& stats_all,
    const OptTensor& dz,
    const Tensor& label_rows, const OptTensor& label_cols, at::IntArrayRef snn, at::IntArrayRef ps,
    int64_t row_offset, int64_t sim_fn, int64_t loss_name, double tau, const Tensor& scalars, const Tensor& gscal,
    const OptTensor& workspace) {
  const int n = (int)mu.size();
  TORCH_CHECK(n >= 1 && n <= 2, "clearvae: 1 or 2 terms");
  check_f32(mu[0], "mu");
  const c10::cuda::CUDAGuard guard(mu[0].device());
  const int64_t B = mu[0].size(0), D = mu[0].size(1);
  check_i64(label_rows, "label");
  int64_t Bg = B;
  const int64_t* lab_c = nullptr;
  if (label_cols.has_value() && label_cols->defined()) {
    check_i64(*label_cols, "label_cols");
    Bg = label_cols->numel();
    lab_c = label_cols->data_ptr<int64_t>();
  }
  check_f32(scalars, "scalars");
  check_f32(gscal, "gscal");
  TORCH_CHECK(gscal.numel() >= 4 && scalars.numel() == CLEARVAE_NSCALARS, "clearvae: bad scalar buffers");
  const float* dzp = nullptr;
  if (dz.has_value() && dz->defined()) {
    check_f32(*dz, "dz");
    TORCH_CHECK(dz->size(0) == B && dz->size(1) == n * D, "clearvae: bad dz shape");
    dzp = dz->data_ptr<float>();
  }
  std::vector<Tensor> dmu, dlv;
  clearvae_term_bwd terms[2];
  auto fopt = mu[0].options();
  for (int i = 0; i < n; ++i) {
    check_f32(mu[i], "mu");
    terms[i].mu = mu[i].data_ptr<float>();
    terms[i].logvar = fptr(logvar.get(i), "logvar", B, D);
    terms[i].eps = fptr(eps.get(i), "eps", B, D);
    terms[i].mu_cols = fptr(mu_cols.get(i), "mu_cols", Bg, D);
    terms[i].row_stats_all = snn[i] ? fptr(stats_all.get(i), "stats_all", Bg, 2) : nullptr;
    TORCH_CHECK(!snn[i] || terms[i].row_stats_all, "clearvae: stats_all missing for an enabled term");
    terms[i].row_aux_all = nullptr;
    if (loss_name != CLEARVAE_LOSS_SNN && snn[i]) {
      TORCH_CHECK((int)stats_all.size() == 2 * n, "clearvae: SupCon row losses need [stats..., aux...] from the forward");
      const OptTensor a = stats_all.get(n + i);
      TORCH_CHECK(a.has_value() && a->defined() && a->numel() == Bg, "clearvae: aux_all missing for an enabled term");
      check_f32(*a, "aux_all");
      terms[i].row_aux_all = a->data_ptr<float>();
    }
    terms[i].dz = dzp ? dzp + i * D : nullptr;
    dmu.push_back(at::empty({B, D}, fopt));
    terms[i].dmu = dmu.back().data_ptr<float>();
    if (terms[i].logvar) {
      dlv.push_back(at::empty({B, D}, fopt));
      terms[i].dlogvar = dlv.back().data_ptr<float>();
    } else {
      dlv.push_back(at::empty({0, D}, fopt));
      terms[i].dlogvar = nullptr;
    }
    terms[i].snn_enable = (int32_t)snn[i];
    terms[i].ps = (int32_t)ps[i];
    terms[i].logvar_cols = (int)logvar_cols.size() == n ? fptr(logvar_cols.get(i), "logvar_cols", Bg, D) : nullptr;
  }
  void* ws = nullptr;
  size_t ws_bytes = 0;
  if (workspace.has_value() && workspace->defined()) {
    TORCH_CHECK(workspace->is_cuda() && workspace->is_contiguous(), "clearvae: workspace must be a contiguous CUDA tensor");
    ws = workspace->data_ptr();
    ws_bytes = (size_t)workspace->numel() * workspace->element_size();
  }
  check_rc(clearvae_latent_bwd_ws(terms, n, label_rows.data_ptr<int64_t>(), lab_c, B, Bg, row_offset, (int32_t)D,
                                  (int32_t)(n * D), (int32_t)sim_fn, (int32_t)loss_name, (float)tau,
                                  scalars.data_ptr<float>(), gscal.data_ptr<float>(), ws, ws_bytes, cur_stream()),
           "latent_bwd");
  return {dmu, dlv};
}

Tensor pair_mask(const Tensor& label_rows, const OptTensor& label_cols, int64_t row_offset, int64_t ps) {
  check_i64(label_rows, "label");
  const c10::cuda::CUDAGuard guard(label_rows.device());
  const int64_t B = label_rows.numel();
  int64_t Bg = B;
  const int64_t* lab_c = nullptr;
  if (label_cols.has_value() && label_cols->defined()) {
    check_i64(*label_cols, "label_cols");
    Bg = label_cols->numel();
    lab_c = label_cols->data_ptr<int64_t>();
  }
  Tensor out = at::empty({B, Bg}, label_rows.options().dtype(at::kByte));
  check_rc(clearvae_pair_mask(label_rows.data_ptr<int64_t>(), lab_c, B, Bg, row_offset, (int32_t)ps,
                              out.data_ptr<uint8_t>(), cur_stream()),
           "pair_mask");
  return out;
}

int64_t latent_bwd_workspace_bytes(int64_t B, int64_t Bg, int64_t D, int64_t n_terms) {
  return (int64_t)clearvae_latent_bwd_workspace_bytes(B, Bg, (int32_t)D, (int32_t)n_terms);
}

int64_t latent_workspace_bytes(int64_t B, int64_t Bg, int64_t D, int64_t n_terms) {
  return (int64_t)clearvae_latent_workspace_bytes(B, Bg, (int32_t)D, (int32_t)n_terms);
}

Tensor recon_fwd(const Tensor& xhat, const Tensor& x, Tensor workspace) {
  check_f32(xhat, "xhat");
  check_f32(x, "x");
  TORCH_CHECK(xhat.sizes() == x.sizes() && xhat.dim() >= 1, "clearvae: xhat / x shape mismatch");
  const c10::cuda::CUDAGuard guard(xhat.device());
  const int64_t B = xhat.size(0);
  Tensor out = at::empty({}, xhat.options());
  check_rc(clearvae_recon_fwd(xhat.data_ptr<float>(), x.data_ptr<float>(), B, xhat.numel() / B, out.data_ptr<float>(),
                              workspace.data_ptr(), (size_t)workspace.nbytes(), cur_stream()),
           "recon_fwd");
  return out;
}

Tensor recon_bwd(const Tensor& xhat, const Tensor& x, const Tensor& grad_out) {
  check_f32(xhat, "xhat");
  check_f32(x, "x");
  check_f32(grad_out, "grad_out");
  const c10::cuda::CUDAGuard guard(xhat.device());
  const int64_t B = xhat.size(0);
  Tensor dx = at::empty_like(xhat);
  check_rc(clearvae_recon_bwd(xhat.data_ptr<float>(), x.data_ptr<float>(), grad_out.data_ptr<float>(), B,
                              xhat.numel() / B, dx.data_ptr<float>(), cur_stream()),
           "recon_bwd");
  return dx;
}

int64_t recon_workspace_bytes() { return (int64_t)clearvae_recon_workspace_bytes(); }

// ---------------------------------------------------------------------------
// convolution-shaped GEMMs
// ---------------------------------------------------------------------------
clearvae_conv_geom geom_from(at::IntArrayRef g) {
  TORCH_CHECK(g.size() == 9, "clearvae: conv geometry = [transposed,k,stride,pad,out_pad,Cin,Cout,Hin,Win]");
  clearvae_conv_geom q;
  q.transposed = (int32_t)g[0]; q.k = (int32_t)g[1]; q.stride = (int32_t)g[2]; q.pad = (int32_t)g[3]; q.out_pad = (int32_t)g[4];
  q.Cin = (int32_t)g[5]; q.Cout = (int32_t)g[6]; q.Hin = (int32_t)g[7]; q.Win = (int32_t)g[8];
  return q;
}
// view [N,H,W,C] logical dims through explicit element strides (sn,sh,sw,sc)
clearvae_tensor4 t4(const Tensor& t, at::IntArrayRef strides, const char* name) {
  TORCH_CHECK(t.is_cuda(), "clearvae: ", name, " must be a CUDA tensor");
  TORCH_CHECK(t.scalar_type() == at::kFloat || t.scalar_type() == at::kBFloat16, "clearvae: ", name, " must be fp32 or bf16");
  TORCH_CHECK(strides.size() == 4, "clearvae: 4 strides expected for ", name);
  clearvae_tensor4 q;
  q.ptr = t.numel() > 0 ? t.data_ptr() : nullptr;
  q.sn = strides[0]; q.sh = strides[1]; q.sw = strides[2]; q.sc = strides[3];
  q.dtype = t.scalar_type() == at::kBFloat16 ? CLEARVAE_BF16 : CLEARVAE_F32;
  return q;
}
const float* optf(const OptTensor& t, const char* name) {
  if (!t.has_value() || !t->defined()) return nullptr;
  check_f32(*t, name);
  return t->data_ptr<float>();
}

Tensor conv_pack_weight(at::IntArrayRef geom, int64_t role, const Tensor& weight) {
  check_f32(weight, "weight");
  const c10::cuda::CUDAGuard guard(weight.device());
  auto g = geom_from(geom);
  const size_t bytes = clearvae_conv_packed_weight_bytes(&g, (int32_t)role);
  TORCH_CHECK(bytes > 0, "clearvae: unsupported conv geometry");
  TORCH_CHECK(weight.numel() == (int64_t)g.Cin * g.Cout * g.k * g.k, "clearvae: weight size does not match the geometry");
  Tensor packed = at::empty({(int64_t)(bytes / 2)}, weight.options().dtype(at::kBFloat16));
  check_rc(clearvae_conv_pack_weight(&g, (int32_t)role, weight.data_ptr<float>(), packed.data_ptr(), cur_stream()), "conv_pack_weight");
  return packed;
}

// table of (geometry, role, weight, packed) tuples for conv_pack_multi: built on the host, returned as a CUDA uint8 tensor whose
// first 8 bytes of metadata travel separately (entries, blocks)
std::tuple<Tensor, int64_t, int64_t> conv_pack_multi_build(at::IntArrayRef geoms_flat, at::IntArrayRef roles, at::TensorList weights,
                                                          at::TensorList packed) {
  const size_t n = weights.size();
  TORCH_CHECK(n >= 1 && roles.size() == n && packed.size() == n && geoms_flat.size() == 9 * n, "clearvae: conv_pack_multi_build argument sizes");
  std::vector<clearvae_conv_geom> g(n);
  std::vector<int32_t> r(n);
  std::vector<const float*> w(n);
  std::vector<void*> pk(n);
  for (size_t i = 0; i < n; ++i) {
    g[i] = geom_from(geoms_flat.slice(9 * i, 9));
    r[i] = (int32_t)roles[i];
    check_f32(weights[i], "weight");
    TORCH_CHECK(packed[i].is_cuda() && packed[i].scalar_type() == at::kBFloat16 && packed[i].is_contiguous(), "clearvae: packed buffers must be bf16 CUDA tensors");
    TORCH_CHECK((size_t)packed[i].numel() * 2 >= clearvae_conv_packed_weight_bytes(&g[i], r[i]), "clearvae: packed buffer too small");
    w[i] = weights[i].data_ptr<float>();
    pk[i] = packed[i].data_ptr();
  }
  const size_t bytes = clearvae_conv_pack_multi_table_bytes((int32_t)n);
  Tensor host = at::empty({(int64_t)bytes}, at::TensorOptions().dtype(at::kByte));
  int32_t ne = 0, nb = 0;
  check_rc(clearvae_conv_pack_multi_build((int32_t)n, g.data(), r.data(), w.data(), pk.data(), host.data_ptr(), &ne, &nb), "conv_pack_multi_build");
  return {host, (int64_t)ne, (int64_t)nb};   // host table: the caller pins it and uploads it (stream-ordered, capture-safe)
}

void conv_pack_multi(const Tensor& table, int64_t n_entries, int64_t n_blocks) {
  const c10::cuda::CUDAGuard guard(table.device());
  TORCH_CHECK(table.is_cuda() && table.scalar_type() == at::kByte, "clearvae: the pack table must be a CUDA uint8 tensor");
  check_rc(clearvae_conv_pack_multi_launch(table.data_ptr(), (int32_t)n_entries, (int32_t)n_blocks, cur_stream()), "conv_pack_multi");
}

void conv_gemm(at::IntArrayRef geom, int64_t role, int64_t batch, const Tensor& src, at::IntArrayRef src_strides,
               const OptTensor& pre_scale, const OptTensor& pre_shift, bool pre_relu, const Tensor& packed_weight,
               const OptTensor& bias, Tensor dst, at::IntArrayRef dst_strides, int64_t epilogue, const OptTensor& mask_src,
               at::IntArrayRef mask_strides, const OptTensor& mask_scale, const OptTensor& mask_shift, const OptTensor& stats) {
  const c10::cuda::CUDAGuard guard(src.device());
  auto g = geom_from(geom);
  auto s4 = t4(src, src_strides, "src");
  auto d4 = t4(dst, dst_strides, "dst");
  clearvae_tensor4 m4{};
  const clearvae_tensor4* mp = nullptr;
  if (mask_src.has_value() && mask_src->defined()) { m4 = t4(*mask_src, mask_strides, "mask_src"); mp = &m4; }
  double* st = nullptr;
  if (stats.has_value() && stats->defined()) {
    TORCH_CHECK(stats->is_cuda() && stats->scalar_type() == at::kDouble && stats->is_contiguous(), "clearvae: stats must be a contiguous CUDA float64 tensor");
    st = stats->data_ptr<double>();
  }
  TORCH_CHECK(packed_weight.is_cuda() && packed_weight.scalar_type() == at::kBFloat16, "clearvae: packed weight must be bf16");
  check_rc(clearvae_conv_gemm(&g, (int32_t)role, batch, &s4, optf(pre_scale, "pre_scale"), optf(pre_shift, "pre_shift"),
                              pre_relu ? 1 : 0, packed_weight.data_ptr(), optf(bias, "bias"), &d4, (int32_t)epilogue, mp,
                              optf(mask_scale, "mask_scale"), optf(mask_shift, "mask_shift"), st, cur_stream()),
           "conv_gemm");
}

void conv_wgrad(at::IntArrayRef geom, int64_t batch, const Tensor& src, at::IntArrayRef src_strides, const OptTensor& pre_scale,
                const OptTensor& pre_shift, bool pre_relu, const Tensor& dy, at::IntArrayRef dy_strides, Tensor dweight, bool split3) {
  const c10::cuda::CUDAGuard guard(src.device());
  auto g = geom_from(geom);
  auto s4 = t4(src, src_strides, "src");
  auto y4 = t4(dy, dy_strides, "dy");
  check_f32(dweight, "dweight");
  TORCH_CHECK(dweight.numel() == (int64_t)g.Cin * g.Cout * g.k * g.k, "clearvae: dweight size does not match the geometry");
  check_rc((split3 ? clearvae_conv_wgrad_split3 : clearvae_conv_wgrad)(&g, batch, &s4, optf(pre_scale, "pre_scale"), optf(pre_shift, "pre_shift"),
                                                                       pre_relu ? 1 : 0, &y4, dweight.data_ptr<float>(), cur_stream()),
           "conv_wgrad");
}

bool conv_direct_wgrad(at::IntArrayRef geom, int64_t batch, const Tensor& src, at::IntArrayRef src_strides, const Tensor& dy,
                       at::IntArrayRef dy_strides, Tensor dweight) {
  const c10::cuda::CUDAGuard guard(src.device());
  auto g = geom_from(geom);
  auto s4 = t4(src, src_strides, "src");
  auto y4 = t4(dy, dy_strides, "dy");
  check_f32(dweight, "dweight");
  TORCH_CHECK(dweight.numel() == (int64_t)g.Cin * g.Cout * g.k * g.k, "clearvae: dweight size does not match the geometry");
  const int rc = clearvae_conv_direct_wgrad(&g, batch, &s4, &y4, dweight.data_ptr<float>(), cur_stream());
  if (rc == CLEARVAE_EUNSUPPORTED) return false;
  check_rc(rc, "conv_direct_wgrad");
  return true;
}

double* stats_ptr(const Tensor& t);
bool conv_direct_dgrad(at::IntArrayRef geom, int64_t batch, const Tensor& dy, at::IntArrayRef dy_strides, const Tensor& weight, Tensor dst,
                       at::IntArrayRef dst_strides, const Tensor& mask_src, at::IntArrayRef mask_strides, const OptTensor& mask_scale,
                       const OptTensor& mask_shift, const OptTensor& stats) {
  const c10::cuda::CUDAGuard guard(dy.device());
  auto g = geom_from(geom);
  auto y4 = t4(dy, dy_strides, "dy");
  auto d4 = t4(dst, dst_strides, "dst");
  auto m4 = t4(mask_src, mask_strides, "mask_src");
  check_f32(weight, "weight");
  double* st = nullptr;
  if (stats.has_value() && stats->defined()) st = stats_ptr(*stats);
  const int rc = clearvae_conv_direct_dgrad(&g, batch, &y4, weight.data_ptr<float>(), &d4, &m4, optf(mask_scale, "mask_scale"),
                                            optf(mask_shift, "mask_shift"), st, cur_stream());
  if (rc == CLEARVAE_EUNSUPPORTED) return false;
  check_rc(rc, "conv_direct_dgrad");
  return true;
}

bool fc_fwd(const Tensor& z, const Tensor& weight, const OptTensor& bias, Tensor out, const OptTensor& stats) {
  check_f32(z, "z");
  check_f32(weight, "weight");
  check_f32(out, "out");
  const c10::cuda::CUDAGuard guard(z.device());
  TORCH_CHECK(z.dim() == 2 && weight.dim() == 2 && weight.size(1) == z.size(1) && out.size(0) == z.size(0) && out.size(1) == weight.size(0),
              "clearvae: fc_fwd shape mismatch");
  double* st = nullptr;
  if (stats.has_value() && stats->defined()) {
    TORCH_CHECK(stats->is_cuda() && stats->scalar_type() == at::kDouble && stats->numel() >= 2 * weight.size(0), "clearvae: bad stats buffer");
    st = stats->data_ptr<double>();
  }
  const int rc = clearvae_fc_fwd(z.data_ptr<float>(), weight.data_ptr<float>(), optf(bias, "bias"), out.data_ptr<float>(), st, z.size(0),
                                 (int32_t)z.size(1), (int32_t)weight.size(0), cur_stream());
  if (rc == CLEARVAE_EUNSUPPORTED) return false;
  check_rc(rc, "fc_fwd");
  return true;
}

int dt_of(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.is_contiguous(), "clearvae: ", name, " must be a contiguous CUDA tensor");
  TORCH_CHECK(t.scalar_type() == at::kFloat || t.scalar_type() == at::kBFloat16, "clearvae: ", name, " must be fp32 or bf16");
  return t.scalar_type() == at::kBFloat16 ? CLEARVAE_BF16 : CLEARVAE_F32;
}
double* stats_ptr(const Tensor& t) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kDouble && t.is_contiguous(), "clearvae: stats must be contiguous CUDA float64");
  return t.data_ptr<double>();
}
float* optf_mut(const OptTensor& t, const char* name) { return const_cast<float*>(optf(t, name)); }

std::tuple<Tensor, Tensor, Tensor, Tensor> bn_finalize(Tensor stats, int64_t C, int64_t group, double count, const OptTensor& gamma,
                                                       const OptTensor& beta, const OptTensor& running_mean,
                                                       const OptTensor& running_var, double momentum, double eps, int64_t expand, int64_t repeat) {
  const c10::cuda::CUDAGuard guard(stats.device());
  TORCH_CHECK(stats.numel() >= 2 * C * group, "clearvae: stats too small");
  auto fopt = stats.options().dtype(at::kFloat);
  Tensor scale = at::empty({C * expand}, fopt), shift = at::empty({C * expand}, fopt), mean = at::empty({C}, fopt), invstd = at::empty({C}, fopt);
  check_rc(clearvae_bn_finalize(stats_ptr(stats), (int32_t)C, (int32_t)group, count, optf(gamma, "gamma"), optf(beta, "beta"),
                                optf_mut(running_mean, "running_mean"), optf_mut(running_var, "running_var"), (float)momentum,
                                (float)eps, scale.data_ptr<float>(), shift.data_ptr<float>(), (int32_t)expand,
                                mean.data_ptr<float>(), invstd.data_ptr<float>(), (int32_t)repeat, cur_stream()),
           "bn_finalize");
  return {scale, shift, mean, invstd};
}

void bn_reduce(const Tensor& y, const OptTensor& g, const OptTensor& act, const OptTensor& mscale, const OptTensor& mshift, int64_t C,
               int64_t inner, int64_t mode, Tensor stats) {
  const c10::cuda::CUDAGuard guard(y.device());
  const bool hg = g.has_value() && g->defined(), ha = act.has_value() && act->defined();
  TORCH_CHECK(!hg || g->numel() == y.numel(), "clearvae: g / y size mismatch");
  TORCH_CHECK(stats.numel() >= 2 * C, "clearvae: stats too small");
  check_rc(clearvae_bn_reduce(y.data_ptr(), dt_of(y, "y"), hg ? g->data_ptr() : nullptr, hg ? dt_of(*g, "g") : 0,
                              ha ? act->data_ptr() : nullptr, ha ? dt_of(*act, "act") : 0, optf(mscale, "mask_scale"),
                              optf(mshift, "mask_shift"), y.numel(), (int32_t)C, inner, (int32_t)mode, stats_ptr(stats), cur_stream()),
           "bn_reduce");
}

std::tuple<Tensor, Tensor> bn_act_fwd(const Tensor& raw, const Tensor& scale, const Tensor& shift, int64_t C, int64_t inner,
                                      int64_t act, int64_t tC, int64_t tHW, int64_t out_dtype, const OptTensor& target, int64_t batch,
                                      Tensor workspace) {
  const c10::cuda::CUDAGuard guard(raw.device());
  check_f32(scale, "scale");
  check_f32(shift, "shift");
  Tensor out = at::empty(raw.sizes(), raw.options().dtype(out_dtype == CLEARVAE_BF16 ? at::kBFloat16 : at::kFloat));
  const bool ht = target.has_value() && target->defined();
  Tensor sse = at::empty({}, raw.options().dtype(at::kFloat));
  if (ht) { check_f32(*target, "target"); TORCH_CHECK(target->numel() == raw.numel(), "clearvae: target size mismatch"); }
  check_rc(clearvae_bn_act_fwd(raw.data_ptr(), dt_of(raw, "raw"), scale.data_ptr<float>(), shift.data_ptr<float>(), raw.numel(),
                               (int32_t)C, inner, (int32_t)act, (int32_t)tC, (int32_t)tHW, out.data_ptr(), (int32_t)out_dtype,
                               ht ? target->data_ptr<float>() : nullptr, batch, sse.data_ptr<float>(), workspace.data_ptr(),
                               (size_t)workspace.nbytes(), cur_stream()),
           "bn_act_fwd");
  return {out, sse};
}

Tensor sigmoid_mse_bwd(const Tensor& xhat, const Tensor& x, const OptTensor& grad_recon, const OptTensor& grad_ext, const Tensor& raw,
                       int64_t C, int64_t inner, int64_t batch, Tensor stats) {
  const c10::cuda::CUDAGuard guard(xhat.device());
  check_f32(xhat, "xhat");
  check_f32(x, "x");
  TORCH_CHECK(xhat.numel() == x.numel() && raw.numel() == x.numel(), "clearvae: size mismatch");
  Tensor g = at::empty_like(xhat);
  check_rc(clearvae_sigmoid_mse_bwd(xhat.data_ptr<float>(), x.data_ptr<float>(), optf(grad_recon, "grad_recon"),
                                    optf(grad_ext, "grad_ext"), raw.data_ptr(), dt_of(raw, "raw"), xhat.numel(), (int32_t)C, inner,
                                    batch, g.data_ptr<float>(), stats_ptr(stats), cur_stream()),
           "sigmoid_mse_bwd");
  return g;
}

std::tuple<Tensor, Tensor, Tensor> bn_bwd_coef(Tensor stats, int64_t C, int64_t group, double count, const OptTensor& gamma,
                                               const Tensor& mean, const Tensor& invstd) {
  const c10::cuda::CUDAGuard guard(stats.device());
  check_f32(mean, "mean");
  check_f32(invstd, "invstd");
  auto fopt = mean.options();
  Tensor coef = at::empty({3, C}, fopt), dgamma = at::empty({C}, fopt), dbeta = at::empty({C}, fopt);
  check_rc(clearvae_bn_bwd_coef(stats_ptr(stats), (int32_t)C, (int32_t)group, count, optf(gamma, "gamma"), mean.data_ptr<float>(),
                                invstd.data_ptr<float>(), coef.data_ptr<float>(), dgamma.data_ptr<float>(),
                                dbeta.data_ptr<float>(), cur_stream()),
           "bn_bwd_coef");
  return {coef, dgamma, dbeta};
}

Tensor bn_bwd_apply(const Tensor& g, const Tensor& y, const OptTensor& act, const OptTensor& mscale, const OptTensor& mshift,
                    const Tensor& coef, int64_t C, int64_t inner, bool to_nhwc, int64_t out_dtype) {
  const c10::cuda::CUDAGuard guard(g.device());
  check_f32(coef, "coef");
  TORCH_CHECK(g.numel() == y.numel(), "clearvae: g / y size mismatch");
  const bool ha = act.has_value() && act->defined();
  Tensor dy = at::empty(y.sizes(), y.options().dtype(out_dtype == CLEARVAE_BF16 ? at::kBFloat16 : at::kFloat));
  check_rc(clearvae_bn_bwd_apply(g.data_ptr(), dt_of(g, "g"), y.data_ptr(), dt_of(y, "y"), ha ? act->data_ptr() : nullptr,
                                 ha ? dt_of(*act, "act") : 0, optf(mscale, "mask_scale"), optf(mshift, "mask_shift"),
                                 coef.data_ptr<float>(), g.numel(), (int32_t)C, inner, to_nhwc ? 1 : 0, dy.data_ptr(),
                                 (int32_t)out_dtype, cur_stream()),
           "bn_bwd_apply");
  return dy;
}

std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor> bn_finalize_apply(Tensor stats, int64_t C, double count, const OptTensor& gamma,
                                                                             const OptTensor& beta, const OptTensor& running_mean,
                                                                             const OptTensor& running_var, double momentum, double eps,
                                                                             int64_t expand, int64_t repeat, const Tensor& raw,
                                                                             int64_t layout, int64_t HW) {
  const c10::cuda::CUDAGuard guard(stats.device());
  TORCH_CHECK(stats.numel() >= 2 * C + 1, "clearvae: stats must hold 2*C moments + the ticket word");
  TORCH_CHECK(raw.is_cuda() && raw.is_contiguous(), "clearvae: raw must be a contiguous CUDA tensor");
  TORCH_CHECK(layout == 2 ? raw.scalar_type() == at::kFloat : raw.scalar_type() == at::kBFloat16, "clearvae: raw dtype does not match the layout");
  auto fopt = stats.options().dtype(at::kFloat);
  auto bopt = raw.options().dtype(at::kBFloat16);
  Tensor scale = at::empty({C * expand}, fopt), shift = at::empty({C * expand}, fopt), mean = at::empty({C}, fopt), invstd = at::empty({C}, fopt);
  Tensor act = layout == 3 ? at::empty({raw.size(0), raw.numel() / raw.size(0)}, bopt) : at::empty(raw.sizes(), bopt);
  Tensor raw_cm = layout == 3 ? at::empty_like(act) : at::empty({0}, bopt);
  check_rc(clearvae_bn_finalize_apply(stats_ptr(stats), (int32_t)C, count, optf(gamma, "gamma"), optf(beta, "beta"),
                                      optf_mut(running_mean, "running_mean"), optf_mut(running_var, "running_var"), (float)momentum,
                                      (float)eps, (int32_t)repeat, scale.data_ptr<float>(), shift.data_ptr<float>(), (int32_t)expand,
                                      mean.data_ptr<float>(), invstd.data_ptr<float>(), raw.data_ptr(), (int32_t)layout, raw.numel(),
                                      (int32_t)HW, act.data_ptr(), layout == 3 ? raw_cm.data_ptr() : nullptr, cur_stream()),
           "bn_finalize_apply");
  return {act, scale, shift, mean, invstd, raw_cm};
}

Tensor bn_relu_apply(const Tensor& raw, const Tensor& scale, const Tensor& shift, int64_t C) {
  const c10::cuda::CUDAGuard guard(raw.device());
  TORCH_CHECK(raw.is_cuda() && raw.is_contiguous() && raw.scalar_type() == at::kBFloat16, "clearvae: raw must be contiguous CUDA bf16");
  check_f32(scale, "scale");
  check_f32(shift, "shift");
  TORCH_CHECK(scale.numel() >= C && shift.numel() >= C, "clearvae: scale / shift too small");
  Tensor out = at::empty_like(raw);
  check_rc(clearvae_bn_relu_apply(raw.data_ptr(), scale.data_ptr<float>(), shift.data_ptr<float>(), raw.numel(), (int32_t)C,
                                  out.data_ptr(), cur_stream()),
           "bn_relu_apply");
  return out;
}

Tensor colsum(const Tensor& x) {
  check_f32(x, "x");
  const c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.dim() == 2, "clearvae: colsum expects a matrix");
  Tensor out = at::empty({x.size(1)}, x.options());
  check_rc(clearvae_colsum(x.data_ptr<float>(), x.size(0), (int32_t)x.size(1), out.data_ptr<float>(), cur_stream()), "colsum");
  return out;
}

bool conv_direct_fwd(at::IntArrayRef geom, int64_t batch, const Tensor& src, at::IntArrayRef src_strides, const OptTensor& pre_scale,
                     const OptTensor& pre_shift, bool pre_relu, const Tensor& weight, const OptTensor& bias, Tensor dst,
                     at::IntArrayRef dst_strides, const OptTensor& stats) {
  const c10::cuda::CUDAGuard guard(src.device());
  auto g = geom_from(geom);
  auto s4 = t4(src, src_strides, "src");
  auto d4 = t4(dst, dst_strides, "dst");
  if (!clearvae_conv_direct_supported(&g, &s4, &d4)) return false;
  if (!g.transposed && (pre_relu || (pre_scale.has_value() && pre_scale->defined()))) return false;
  check_f32(weight, "weight");
  double* st = nullptr;
  if (stats.has_value() && stats->defined()) st = stats_ptr(*stats);
  check_rc(clearvae_conv_direct_fwd(&g, batch, &s4, optf(pre_scale, "pre_scale"), optf(pre_shift, "pre_shift"), pre_relu ? 1 : 0,
                                    weight.data_ptr<float>(), optf(bias, "bias"), &d4, st, cur_stream()),
           "conv_direct_fwd");
  return true;
}

// ---------------------------------------------------------------------------
// MI estimators (CLUB-S / L1OutUB) and fused Adam
// ---------------------------------------------------------------------------
void check_rows(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kFloat && t.dim() == 2 && (t.size(1) == 1 || t.stride(1) == 1) && t.stride(0) >= t.size(1),
              "clearvae: ", name, " must be a CUDA float32 [B, D] view with unit column stride");
}

std::tuple<Tensor, Tensor, Tensor> mi_estimator(int64_t mode, const Tensor& x, const Tensor& y, const OptTensor& perm,
                                                at::TensorList params, Tensor workspace) {
  check_rows(x, "x");
  check_rows(y, "y");
  const c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.size(0) == y.size(0), "clearvae: x / y must have equal B");
  TORCH_CHECK(params.size() == 8, "clearvae: 8 estimator parameters expected");
  const int64_t B = x.size(0), Dx = x.size(1), Dy = y.size(1), H = params[0].size(0);
  const float* pp[8];
  const int64_t want[8][2] = {{H, Dx}, {H, 0}, {Dy, H}, {Dy, 0}, {H, Dx}, {H, 0}, {Dy, H}, {Dy, 0}};
  int64_t P = 0;
  for (int i = 0; i < 8; ++i) {
    check_f32(params[i], "estimator parameter");
    TORCH_CHECK(params[i].size(0) == want[i][0] && (want[i][1] == 0 ? params[i].dim() == 1 : params[i].size(1) == want[i][1]),
                "clearvae: estimator parameter ", i, " has the wrong shape");
    pp[i] = params[i].data_ptr<float>();
    P += params[i].numel();
  }
  const int64_t* pm = nullptr;
  if (perm.has_value() && perm->defined()) {
    check_i64(*perm, "perm");
    TORCH_CHECK(perm->numel() == B, "clearvae: perm must have B entries");
    pm = perm->data_ptr<int64_t>();
  }
  auto fopt = x.options();
  const bool learn = mode == CLEARVAE_MI_LEARN;
  Tensor out = at::empty({learn ? 1 + P : mode == CLEARVAE_MI_L1OUT ? 1 + 2 * Dy : 1}, fopt);
  Tensor dx = learn ? at::empty({0}, fopt) : at::empty({B, Dx}, fopt);
  Tensor dy = learn ? at::empty({0}, fopt) : at::empty({B, Dy}, fopt);
  TORCH_CHECK(workspace.is_cuda() && workspace.is_contiguous(), "clearvae: workspace must be a contiguous CUDA tensor");
  check_rc(clearvae_mi_estimator((int32_t)mode, x.data_ptr<float>(), x.stride(0), y.data_ptr<float>(), y.stride(0), pm, B, (int32_t)Dx,
                                 (int32_t)H, (int32_t)Dy, pp, out.data_ptr<float>(), learn ? nullptr : dx.data_ptr<float>(),
                                 learn ? nullptr : dy.data_ptr<float>(), workspace.data_ptr(), (size_t)workspace.nbytes(),
                                 cur_stream()),
           "mi_estimator");
  return {out, dx, dy};
}

std::tuple<Tensor, Tensor> mi_bound_bwd(int64_t mode, const Tensor& grad_out, const Tensor& dx_unit, const Tensor& dy_unit,
                                        const Tensor& y, const Tensor& out_fwd) {
  check_f32(grad_out, "grad_out");
  check_f32(dx_unit, "dx_unit");
  check_f32(dy_unit, "dy_unit");
  check_rows(y, "y");
  check_f32(out_fwd, "out_fwd");
  const c10::cuda::CUDAGuard guard(y.device());
  Tensor gx = at::empty_like(dx_unit), gy = at::empty_like(dy_unit);
  check_rc(clearvae_mi_bound_bwd((int32_t)mode, grad_out.data_ptr<float>(), dx_unit.data_ptr<float>(), dy_unit.data_ptr<float>(),
                                 y.data_ptr<float>(), y.stride(0), out_fwd.data_ptr<float>(), dx_unit.size(0), (int32_t)dx_unit.size(1),
                                 (int32_t)dy_unit.size(1), gx.data_ptr<float>(), gy.data_ptr<float>(), cur_stream()),
           "mi_bound_bwd");
  return {gx, gy};
}

std::tuple<Tensor, Tensor> tc_factor(int64_t mode, const Tensor& z, const Tensor& w1, const Tensor& b1, const Tensor& w2, const Tensor& b2,
                                     Tensor workspace) {
  check_rows(z, "z");
  check_f32(w1, "w1");
  check_f32(b1, "b1");
  check_f32(w2, "w2");
  check_f32(b2, "b2");
  const c10::cuda::CUDAGuard guard(z.device());
  const int64_t B = z.size(0), Z = z.size(1);
  TORCH_CHECK(w1.numel() == Z * Z && b1.numel() == Z && w2.numel() == Z && b2.numel() == 1, "clearvae: factor_cls parameter shapes");
  auto fopt = z.options();
  const bool disc = mode == CLEARVAE_TC_DISC;
  Tensor out = at::empty({disc ? 1 + Z * Z + 2 * Z + 1 : 1}, fopt);
  Tensor dz = disc ? at::empty({0}, fopt) : at::empty({B, Z}, fopt);
  TORCH_CHECK(workspace.is_cuda() && workspace.is_contiguous(), "clearvae: workspace must be a contiguous CUDA tensor");
  check_rc(clearvae_tc_factor((int32_t)mode, z.data_ptr<float>(), z.stride(0), B, (int32_t)Z, w1.data_ptr<float>(), b1.data_ptr<float>(),
                              w2.data_ptr<float>(), b2.data_ptr<float>(), out.data_ptr<float>(), disc ? nullptr : dz.data_ptr<float>(),
                              workspace.data_ptr(), (size_t)workspace.nbytes(), cur_stream()),
           "tc_factor");
  return {out, dz};
}

int64_t tc_workspace_bytes(int64_t mode, int64_t B, int64_t Z) { return (int64_t)clearvae_tc_workspace_bytes((int32_t)mode, B, (int32_t)Z); }

Tensor scale_by(const Tensor& grad_out, const Tensor& x) {
  check_f32(grad_out, "grad_out");
  check_f32(x, "x");
  const c10::cuda::CUDAGuard guard(x.device());
  Tensor y = at::empty_like(x);
  if (x.numel() > 0) check_rc(clearvae_scale(grad_out.data_ptr<float>(), x.data_ptr<float>(), x.numel(), y.data_ptr<float>(), cur_stream()), "scale");
  return y;
}

int64_t mi_workspace_bytes(int64_t mode, int64_t B, int64_t Dx, int64_t H, int64_t Dy) {
  return (int64_t)clearvae_mi_workspace_bytes((int32_t)mode, B, (int32_t)Dx, (int32_t)H, (int32_t)Dy);
}

void adam_step(at::TensorList params, at::TensorList grads, at::TensorList exp_avg, at::TensorList exp_avg_sq, Tensor steps,
               Tensor counter, double lr, double beta1, double beta2, double eps, double grad_scale) {
  const size_t n = params.size();
  TORCH_CHECK(grads.size() == n && exp_avg.size() == n && exp_avg_sq.size() == n, "clearvae: adam lists must have equal length");
  if (n == 0) return;
  const c10::cuda::CUDAGuard guard(params[0].device());
  std::vector<float*> p(n), m(n), v(n);
  std::vector<const float*> g(n);
  std::vector<int64_t> ne(n);
  for (size_t i = 0; i < n; ++i) {
    check_f32(params[i], "param");
    check_f32(grads[i], "grad");
    check_f32(exp_avg[i], "exp_avg");
    check_f32(exp_avg_sq[i], "exp_avg_sq");
    ne[i] = params[i].numel();
    TORCH_CHECK(grads[i].numel() == ne[i] && exp_avg[i].numel() == ne[i] && exp_avg_sq[i].numel() == ne[i],
                "clearvae: adam tensor sizes differ");
    p[i] = params[i].data_ptr<float>(); g[i] = grads[i].data_ptr<float>();
    m[i] = exp_avg[i].data_ptr<float>(); v[i] = exp_avg_sq[i].data_ptr<float>();
  }
  check_f32(steps, "steps");
  TORCH_CHECK(counter.is_cuda() && counter.scalar_type() == at::kInt && counter.numel() >= 1, "clearvae: counter must be a CUDA int32 tensor");
  check_rc(clearvae_adam_step((int32_t)n, p.data(), g.data(), m.data(), v.data(), ne.data(), steps.data_ptr<float>(),
                              (int32_t)steps.numel(), reinterpret_cast<unsigned int*>(counter.data_ptr<int32_t>()), (float)lr,
                              (float)beta1, (float)beta2, (float)eps, (float)grad_scale, cur_stream()),
           "adam_step");
  // the packed bf16 weight copies are refreshed when a master's version counter moves (engine._PackCache)
  for (size_t i = 0; i < n; ++i) params[i].unsafeGetTensorImpl()->bump_version();
}

int64_t bn_act_workspace_bytes() { return (int64_t)clearvae_bn_act_workspace_bytes(); }

std::vector<Tensor> reparam_multi(at::TensorList mu, at::TensorList logvar, at::TensorList eps) {
  const size_t heads = mu.size();
  TORCH_CHECK(heads >= 1 && heads <= 2 && logvar.size() == heads && !eps.empty() && eps.size() % heads == 0,
              "clearvae: reparam_multi takes 1-2 heads and draws x heads noise tensors");
  const size_t draws = eps.size() / heads;
  TORCH_CHECK(draws <= CLEARVAE_REPARAM_MAX_DRAWS, "clearvae: too many draws");
  const c10::cuda::CUDAGuard guard(mu[0].device());
  const int64_t B = mu[0].size(0), D = mu[0].size(1);
  std::vector<const float*> pm(heads), pl(heads), pe(eps.size());
  for (size_t h = 0; h < heads; ++h) {
    pm[h] = fptr(mu[h], "mu", B, D);
    pl[h] = fptr(logvar[h], "logvar", B, D);
  }
  for (size_t i = 0; i < eps.size(); ++i) pe[i] = fptr(eps[i], "eps", B, D);
  std::vector<Tensor> z(draws);
  std::vector<float*> pz(draws);
  for (size_t j = 0; j < draws; ++j) {
    z[j] = at::empty({B, (int64_t)heads * D}, mu[0].options());
    pz[j] = z[j].data_ptr<float>();
  }
  check_rc(clearvae_reparam_multi((int32_t)heads, (int32_t)draws, pm.data(), pl.data(), pe.data(), pz.data(), B, (int32_t)D, cur_stream()),
           "reparam_multi");
  return z;
}

// ---- group evidence (ML-VAE / GVAE baselines)
std::tuple<Tensor, Tensor, Tensor> group_evidence_fwd(int64_t mode, const Tensor& mu, const Tensor& logvar, const Tensor& gid, int64_t G) {
  const c10::cuda::CUDAGuard guard(mu.device());
  const int64_t B = mu.size(0), D = mu.size(1);
  const float* pm = fptr(mu, "mu", B, D);
  const float* pl = fptr(logvar, "logvar", B, D);
  TORCH_CHECK(gid.is_cuda() && gid.scalar_type() == at::kLong && gid.is_contiguous() && gid.numel() == B, "clearvae: group ids must be int64 [B] on CUDA");
  TORCH_CHECK(G >= 1, "clearvae: at least one group");
  Tensor mg = at::empty({G, D}, mu.options()), lg = at::empty({G, D}, mu.options()), cnt = at::empty({G}, mu.options());
  check_rc(clearvae_group_evidence_fwd((int32_t)mode, pm, pl, gid.data_ptr<int64_t>(), B, (int32_t)D, (int32_t)G, mg.data_ptr<float>(),
                                       lg.data_ptr<float>(), cnt.data_ptr<float>(), cur_stream()), "group_evidence_fwd");
  return {mg, lg, cnt};
}

std::tuple<Tensor, Tensor> group_evidence_bwd(int64_t mode, const Tensor& mu, const Tensor& logvar, const Tensor& gid, const Tensor& mg,
                                              const Tensor& lg, const Tensor& cnt, const Tensor& dmg, const Tensor& dlg) {
  const c10::cuda::CUDAGuard guard(mu.device());
  const int64_t B = mu.size(0), D = mu.size(1), G = mg.size(0);
  Tensor dmu = at::empty_like(mu), dlv = at::empty_like(mu);
  check_rc(clearvae_group_evidence_bwd((int32_t)mode, fptr(mu, "mu", B, D), fptr(logvar, "logvar", B, D), gid.data_ptr<int64_t>(),
                                       fptr(mg, "mu_grp", G, D), fptr(lg, "logvar_grp", G, D), cnt.data_ptr<float>(), fptr(dmg, "dmu_grp", G, D),
                                       fptr(dlg, "dlogvar_grp", G, D), B, (int32_t)D, dmu.data_ptr<float>(), dlv.data_ptr<float>(), cur_stream()),
           "group_evidence_bwd");
  return {dmu, dlv};
}

Tensor group_reparam_fwd(const Tensor& mg, const Tensor& lg, const Tensor& eps, const Tensor& gid) {
  const c10::cuda::CUDAGuard guard(eps.device());
  const int64_t B = eps.size(0), D = eps.size(1), G = mg.size(0);
  Tensor z = at::empty_like(eps);
  check_rc(clearvae_group_reparam_fwd(fptr(mg, "mu_grp", G, D), fptr(lg, "logvar_grp", G, D), fptr(eps, "eps", B, D), gid.data_ptr<int64_t>(), B,
                                      (int32_t)D, z.data_ptr<float>(), cur_stream()), "group_reparam_fwd");
  return z;
}

std::tuple<Tensor, Tensor> group_reparam_bwd(const Tensor& dz, const Tensor& eps, const Tensor& gid, int64_t G) {
  const c10::cuda::CUDAGuard guard(eps.device());
  const int64_t B = eps.size(0), D = eps.size(1);
  Tensor a = at::empty({G, D}, eps.options()), b = at::empty({G, D}, eps.options());
  check_rc(clearvae_group_reparam_bwd(fptr(dz, "dz", B, D), fptr(eps, "eps", B, D), gid.data_ptr<int64_t>(), B, (int32_t)D, (int32_t)G,
                                      a.data_ptr<float>(), b.data_ptr<float>(), cur_stream()), "group_reparam_bwd");
  return {a, b};
}

// ---- one-shot collectives over peer memory (bases = every rank's buffer as mapped into this process)
std::vector<void*> peer_bases(at::IntArrayRef bases) {
  TORCH_CHECK(!bases.empty() && bases.size() <= CLEARVAE_PEER_MAX_RANKS, "clearvae: 1..8 peer buffers");
  std::vector<void*> b(bases.size());
  for (size_t i = 0; i < bases.size(); ++i) b[i] = reinterpret_cast<void*>(static_cast<uintptr_t>(bases[i]));
  return b;
}

void peer_gather(at::IntArrayRef bases, int64_t rank, int64_t buffer_bytes, at::TensorList src, at::TensorList dst) {
  const size_t n = src.size();
  TORCH_CHECK(n >= 1 && n <= CLEARVAE_PEER_MAX_PIECES && dst.size() == n, "clearvae: peer_gather takes 1..8 (src, dst) pairs");
  const c10::cuda::CUDAGuard guard(src[0].device());
  auto b = peer_bases(bases);
  std::vector<const void*> s(n);
  std::vector<void*> d(n);
  std::vector<int64_t> nb(n);
  for (size_t i = 0; i < n; ++i) {
    TORCH_CHECK(src[i].is_cuda() && dst[i].is_cuda() && src[i].is_contiguous() && dst[i].is_contiguous(),
                "clearvae: peer_gather needs contiguous CUDA tensors");
    TORCH_CHECK(src[i].scalar_type() == dst[i].scalar_type(), "clearvae: peer_gather dtype mismatch");
    nb[i] = (int64_t)src[i].numel() * (int64_t)src[i].element_size();
    TORCH_CHECK((int64_t)dst[i].numel() * (int64_t)dst[i].element_size() == nb[i] * (int64_t)bases.size(),
                "clearvae: peer_gather dst must hold world x src");
    s[i] = src[i].data_ptr();
    d[i] = dst[i].data_ptr();
  }
  check_rc(clearvae_peer_gather(b.data(), (int32_t)bases.size(), (int32_t)rank, buffer_bytes, (int32_t)n, s.data(), d.data(),
                                nb.data(), cur_stream()), "peer_gather");
}

void peer_allreduce(at::IntArrayRef bases, int64_t rank, int64_t buffer_bytes, at::TensorList tensors) {
  const size_t n = tensors.size();
  if (n == 0) return;
  const c10::cuda::CUDAGuard guard(tensors[0].device());
  auto b = peer_bases(bases);
  std::vector<float*> t(n);
  std::vector<int64_t> ne(n);
  for (size_t i = 0; i < n; ++i) {
    check_f32(tensors[i], "tensor");
    t[i] = tensors[i].data_ptr<float>();
    ne[i] = tensors[i].numel();
  }
  check_rc(clearvae_peer_allreduce(b.data(), (int32_t)bases.size(), (int32_t)rank, buffer_bytes, (int32_t)n, t.data(), ne.data(),
                                   cur_stream()), "peer_allreduce");
}

}  // namespace

TORCH_LIBRARY(clearvae, m) {
  m.def("latent_fwd(Tensor[] mu, Tensor?[] logvar, Tensor?[] eps, Tensor?[] mu_cols, Tensor?[] logvar_cols, Tensor label_rows, "
        "Tensor? label_cols, int[] snn, int[] ps, int row_offset, int sim_fn, int loss_name, float tau, "
        "bool finalize, bool want_z, Tensor(a!) workspace) -> (Tensor, Tensor, Tensor[])");
  m.def("snn_finalize(Tensor stats_all, int term, Tensor(a!) scalars) -> ()");
  m.def("latent_bwd(Tensor[] mu, Tensor?[] logvar, Tensor?[] eps, Tensor?[] mu_cols, Tensor?[] logvar_cols, Tensor?[] stats_all, "
        "Tensor? dz, Tensor label_rows, Tensor? label_cols, int[] snn, int[] ps, int row_offset, int sim_fn, "
        "int loss_name, float tau, Tensor scalars, Tensor gscal, Tensor(a!)? workspace=None) -> (Tensor[], Tensor[])");
  m.def("latent_bwd_workspace_bytes(int B, int Bg, int D, int n_terms) -> int", &latent_bwd_workspace_bytes);
  m.def("pair_mask(Tensor label_rows, Tensor? label_cols, int row_offset, int ps) -> Tensor");
  m.def("latent_workspace_bytes(int B, int Bg, int D, int n_terms) -> int", &latent_workspace_bytes);
  m.def("recon_fwd(Tensor xhat, Tensor x, Tensor(a!) workspace) -> Tensor");
  m.def("recon_bwd(Tensor xhat, Tensor x, Tensor grad_out) -> Tensor");
  m.def("recon_workspace_bytes() -> int", &recon_workspace_bytes);
  m.def("conv_pack_weight(int[] geom, int role, Tensor weight) -> Tensor");
  m.def("conv_pack_multi_build(int[] geoms_flat, int[] roles, Tensor[] weights, Tensor[] packed) -> (Tensor, int, int)");
  m.def("conv_pack_multi(Tensor table, int n_entries, int n_blocks) -> ()");
  m.def("conv_direct_fwd(int[] geom, int batch, Tensor src, int[] src_strides, Tensor? pre_scale, Tensor? pre_shift, bool pre_relu, "
        "Tensor weight, Tensor? bias, Tensor(a!) dst, int[] dst_strides, Tensor(b!)? stats) -> bool");
  m.def("conv_direct_wgrad(int[] geom, int batch, Tensor src, int[] src_strides, Tensor dy, int[] dy_strides, Tensor(a!) dweight) -> bool");
  m.def("conv_direct_dgrad(int[] geom, int batch, Tensor dy, int[] dy_strides, Tensor weight, Tensor(a!) dst, int[] dst_strides, "
        "Tensor mask_src, int[] mask_strides, Tensor? mask_scale, Tensor? mask_shift, Tensor(b!)? stats) -> bool");
  m.def("fc_fwd(Tensor z, Tensor weight, Tensor? bias, Tensor(a!) out, Tensor(b!)? stats) -> bool");
  m.def("conv_wgrad(int[] geom, int batch, Tensor src, int[] src_strides, Tensor? pre_scale, Tensor? pre_shift, bool pre_relu, "
        "Tensor dy, int[] dy_strides, Tensor(a!) dweight, bool split3=False) -> ()");
  m.def("bn_finalize(Tensor(a!) stats, int C, int group, float count, Tensor? gamma, Tensor? beta, Tensor(b!)? running_mean, "
        "Tensor(c!)? running_var, float momentum, float eps, int expand, int repeat) -> (Tensor, Tensor, Tensor, Tensor)");
  m.def("bn_reduce(Tensor y, Tensor? g, Tensor? act, Tensor? mask_scale, Tensor? mask_shift, int C, int inner, int mode, Tensor(a!) stats) -> ()");
  m.def("bn_act_fwd(Tensor raw, Tensor scale, Tensor shift, int C, int inner, int act, int to_nhwc_C, int to_nhwc_HW, int out_dtype, Tensor? target, int batch, "
        "Tensor(a!) workspace) -> (Tensor, Tensor)");
  m.def("sigmoid_mse_bwd(Tensor xhat, Tensor x, Tensor? grad_recon, Tensor? grad_ext, Tensor raw, int C, int inner, int batch, "
        "Tensor(a!) stats) -> Tensor");
  m.def("bn_bwd_coef(Tensor(a!) stats, int C, int group, float count, Tensor? gamma, Tensor mean, Tensor invstd) -> (Tensor, Tensor, Tensor)");
  m.def("bn_bwd_apply(Tensor g, Tensor y, Tensor? act, Tensor? mask_scale, Tensor? mask_shift, Tensor coef, int C, int inner, bool to_nhwc, int out_dtype) -> Tensor");
  m.def("colsum(Tensor x) -> Tensor");
  m.def("bn_relu_apply(Tensor raw, Tensor scale, Tensor shift, int C) -> Tensor");
  m.def("bn_finalize_apply(Tensor(a!) stats, int C, float count, Tensor? gamma, Tensor? beta, Tensor(b!)? running_mean, "
        "Tensor(c!)? running_var, float momentum, float eps, int expand, int repeat, Tensor raw, int layout, int HW) "
        "-> (Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)");
  m.def("mi_estimator(int mode, Tensor x, Tensor y, Tensor? perm, Tensor[] params, Tensor(a!) workspace) -> (Tensor, Tensor, Tensor)");
  m.def("mi_bound_bwd(int mode, Tensor grad_out, Tensor dx_unit, Tensor dy_unit, Tensor y, Tensor out_fwd) -> (Tensor, Tensor)");
  m.def("mi_workspace_bytes(int mode, int B, int Dx, int H, int Dy) -> int", &mi_workspace_bytes);
  m.def("tc_factor(int mode, Tensor z, Tensor w1, Tensor b1, Tensor w2, Tensor b2, Tensor(a!) workspace) -> (Tensor, Tensor)");
  m.def("tc_workspace_bytes(int mode, int B, int Z) -> int", &tc_workspace_bytes);
  m.def("scale_by(Tensor grad_out, Tensor x) -> Tensor");
  m.def("adam_step(Tensor(a!)[] params, Tensor[] grads, Tensor(b!)[] exp_avg, Tensor(c!)[] exp_avg_sq, Tensor(d!) steps, "
        "Tensor(e!) counter, float lr, float beta1, float beta2, float eps, float grad_scale) -> ()");
  m.def("bn_act_workspace_bytes() -> int", &bn_act_workspace_bytes);
  m.def("reparam_multi(Tensor[] mu, Tensor[] logvar, Tensor[] eps) -> Tensor[]");
  m.def("group_evidence_fwd(int mode, Tensor mu, Tensor logvar, Tensor group_id, int G) -> (Tensor, Tensor, Tensor)");
  m.def("group_evidence_bwd(int mode, Tensor mu, Tensor logvar, Tensor group_id, Tensor mu_grp, Tensor logvar_grp, Tensor count, "
        "Tensor dmu_grp, Tensor dlogvar_grp) -> (Tensor, Tensor)");
  m.def("group_reparam_fwd(Tensor mu_grp, Tensor logvar_grp, Tensor eps, Tensor group_id) -> Tensor");
  m.def("group_reparam_bwd(Tensor dz, Tensor eps, Tensor group_id, int G) -> (Tensor, Tensor)");
  m.def("peer_gather(int[] bases, int rank, int buffer_bytes, Tensor[] src, Tensor(a!)[] dst) -> ()");
  m.def("peer_allreduce(int[] bases, int rank, int buffer_bytes, Tensor(a!)[] tensors) -> ()");
  m.def("conv_gemm(int[] geom, int role, int batch, Tensor src, int[] src_strides, Tensor? pre_scale, Tensor? pre_shift, "
        "bool pre_relu, Tensor packed_weight, Tensor? bias, Tensor(a!) dst, int[] dst_strides, int epilogue, "
        "Tensor? mask_src, int[] mask_strides, Tensor? mask_scale, Tensor? mask_shift, Tensor(b!)? stats) -> ()");
}

TORCH_LIBRARY_IMPL(clearvae, CUDA, m) {
  m.impl("group_evidence_fwd", &group_evidence_fwd);
  m.impl("group_evidence_bwd", &group_evidence_bwd);
  m.impl("group_reparam_fwd", &group_reparam_fwd);
  m.impl("group_reparam_bwd", &group_reparam_bwd);
  m.impl("latent_fwd", &latent_fwd);
  m.impl("snn_finalize", &snn_finalize);
  m.impl("latent_bwd", &latent_bwd);
  m.impl("pair_mask", &pair_mask);
  m.impl("recon_fwd", &recon_fwd);
  m.impl("recon_bwd", &recon_bwd);
  m.impl("conv_pack_weight", &conv_pack_weight);
  m.impl("conv_pack_multi_build", &conv_pack_multi_build);
  m.impl("conv_pack_multi", &conv_pack_multi);
  m.impl("conv_gemm", &conv_gemm);
  m.impl("conv_direct_fwd", &conv_direct_fwd);
  m.impl("conv_wgrad", &conv_wgrad);
  m.impl("fc_fwd", &fc_fwd);
  m.impl("conv_direct_dgrad", &conv_direct_dgrad);
  m.impl("conv_direct_wgrad", &conv_direct_wgrad);
  m.impl("bn_finalize", &bn_finalize);
  m.impl("bn_reduce", &bn_reduce);
  m.impl("bn_act_fwd", &bn_act_fwd);
  m.impl("sigmoid_mse_bwd", &sigmoid_mse_bwd);
  m.impl("bn_bwd_coef", &bn_bwd_coef);
  m.impl("bn_bwd_apply", &bn_bwd_apply);
  m.impl("colsum", &colsum);
  m.impl("bn_relu_apply", &bn_relu_apply);
  m.impl("bn_finalize_apply", &bn_finalize_apply);
  m.impl("mi_estimator", &mi_estimator);
  m.impl("tc_factor", &tc_factor);
  m.impl("scale_by", &scale_by);
  m.impl("mi_bound_bwd", &mi_bound_bwd);
  m.impl("adam_step", &adam_step);
  m.impl("reparam_multi", &reparam_multi);
  m.impl("peer_gather", &peer_gather);
  m.impl("peer_allreduce", &peer_allreduce);
}
