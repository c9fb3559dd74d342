// Python bindings (torch extension) for the sm_100a kernels.  Only this file
// includes torch headers; the .cu files are plain CUDA so they rebuild in
// seconds.  Every op runs on the caller's current CUDA stream.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <cstring>
#include <string>
#include <vector>

#include "feed.h"
#include "igemm.h"
#include "ops.h"
#include "tfrecord.h"
#include "vmm.h"

namespace py = pybind11;
using torch::Tensor;

namespace {

inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }

inline void check(cudaError_t e, const char* what) {
  TORCH_CHECK(e == cudaSuccess, "tfos_b200 ", what, ": ", cudaGetErrorString(e));
}
inline void need(const Tensor& t, c10::ScalarType dt, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.scalar_type() == dt, name, " has wrong dtype");
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}
inline const void* optptr(const c10::optional<Tensor>& t) {
  return t.has_value() && t->defined() ? t->data_ptr() : nullptr;
}
inline float* optf(const c10::optional<Tensor>& t) {
  return t.has_value() && t->defined() ? t->data_ptr<float>() : nullptr;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename T>
T geti(const py::dict& d, const char* k, T dflt) {
  return d.contains(k) ? d[k].cast<T>() : dflt;
}
void fill_ints(const py::dict& d, const char* k, int* dst, int n) {
  for (int i = 0; i < n; ++i) dst[i] = 0;
  if (!d.contains(k)) return;
  auto v = d[k].cast<std::vector<int>>();
  TORCH_CHECK(static_cast<int>(v.size()) <= n, k, " has too many entries");
  for (size_t i = 0; i < v.size(); ++i) dst[i] = v[i];
}

tfos::TmapDesc parse_tmap(const py::dict& d) {
  tfos::TmapDesc t;
  std::memset(&t, 0, sizeof(t));
  t.base = reinterpret_cast<void*>(d["base"].cast<uint64_t>());
  auto dims = d["dims"].cast<std::vector<uint64_t>>();
  auto strides = d["strides"].cast<std::vector<uint64_t>>();
  auto box = d["box"].cast<std::vector<uint32_t>>();
  t.rank = static_cast<int>(dims.size());
  TORCH_CHECK(t.rank >= 2 && t.rank <= 4, "tensor map rank must be 2..4");
  TORCH_CHECK(static_cast<int>(strides.size()) == t.rank - 1 &&
                  static_cast<int>(box.size()) == t.rank,
              "tensor map strides/box rank mismatch");
  for (int i = 0; i < t.rank; ++i) {
    t.dims[i] = dims[i];
    t.box[i] = box[i];
    t.elem_strides[i] = 1;
  }
  for (int i = 0; i + 1 < t.rank; ++i) t.strides_bytes[i] = strides[i];
  if (d.contains("elem_strides")) {
    auto es = d["elem_strides"].cast<std::vector<uint32_t>>();
    for (int i = 0; i < t.rank && i < static_cast<int>(es.size()); ++i) t.elem_strides[i] = es[i];
  }
  return t;
}

int64_t plan_fwd(const py::dict& a, const py::dict& b, const py::dict& g, int bn, bool b_mn) {
  tfos::FwdArgs f;
  std::memset(&f, 0, sizeof(f));
  std::memset(&f, 0, sizeof(f));
  f.tiles_w = geti<int>(g, "tiles_w", 1);
  f.tiles_h = geti<int>(g, "tiles_h", 1);
  f.tiles_n = geti<int>(g, "tiles_n", 1);
  f.n_tiles = geti<int>(g, "n_tiles", 1);
  f.box_w = geti<int>(g, "box_w", 128);
  f.box_h = geti<int>(g, "box_h", 1);
  f.box_n = geti<int>(g, "box_n", 1);
  f.mul_w = geti<int>(g, "mul_w", 1);
  f.mul_h = geti<int>(g, "mul_h", 1);
  f.num_taps = geti<int>(g, "num_taps", 1);
  f.k_chunks = geti<int>(g, "k_chunks", 1);
  fill_ints(g, "tap_dw", f.tap_dw, tfos::kMaxTaps);
  fill_ints(g, "tap_dh", f.tap_dh, tfos::kMaxTaps);
  fill_ints(g, "tap_dc", f.tap_dc, tfos::kMaxTaps);
  fill_ints(g, "tap_bk", f.tap_bk, tfos::kMaxTaps);
  fill_ints(g, "tap_bn", f.tap_bn, tfos::kMaxTaps);
  f.lim_w = geti<int>(g, "lim_w", 1);
  f.lim_h = geti<int>(g, "lim_h", 1);
  f.lim_n = geti<int>(g, "lim_n", 1);
  f.OW = geti<int>(g, "OW", f.lim_w);
  f.OH = geti<int>(g, "OH", f.lim_h);
  f.osw = geti<int>(g, "osw", 1);
  f.oow = geti<int>(g, "oow", 0);
  f.osh = geti<int>(g, "osh", 1);
  f.ooh = geti<int>(g, "ooh", 0);
  f.ldo = geti<int>(g, "ldo", 0);
  f.n_valid = geti<int>(g, "n_valid", 0);
  f.relu = geti<int>(g, "relu", 0);
  f.out_fp32 = geti<int>(g, "out_fp32", 0);
  f.accumulate = geti<int>(g, "accumulate", 0);
  f.stem = geti<int>(g, "stem", 0);
  f.acc_mask = reinterpret_cast<const uint8_t*>(geti<uint64_t>(g, "acc_mask", 0));
  f.bias = reinterpret_cast<const float*>(geti<uint64_t>(g, "bias", 0));
  f.col_sum = reinterpret_cast<float*>(geti<uint64_t>(g, "col_sum", 0));
  f.col_sumsq = reinterpret_cast<float*>(geti<uint64_t>(g, "col_sumsq", 0));
  f.out = reinterpret_cast<void*>(geti<uint64_t>(g, "out", 0));
  f.red_x = reinterpret_cast<const void*>(geti<uint64_t>(g, "red_x", 0));
  f.red_mask = reinterpret_cast<const uint8_t*>(geti<uint64_t>(g, "red_mask", 0));
  TORCH_CHECK(f.out != nullptr && f.ldo > 0 && f.n_valid > 0, "igemm fwd: out/ldo/n_valid");
  char err[512] = {0};
  tfos::IGemmPlan* p =
      tfos::igemm_plan_fwd(parse_tmap(a), parse_tmap(b), f, bn, b_mn ? 1 : 0, num_sms(), err, 512);
  TORCH_CHECK(p != nullptr, "igemm_plan_fwd: ", err);
  return reinterpret_cast<int64_t>(p);
}

int64_t plan_wgrad(const py::dict& a, const py::dict& b, const py::dict& g, int bn) {
  tfos::WgradArgs w;
  std::memset(&w, 0, sizeof(w));
  w.tiles_w = geti<int>(g, "tiles_w", 1);
  w.tiles_h = geti<int>(g, "tiles_h", 1);
  w.tiles_n = geti<int>(g, "tiles_n", 1);
  w.box_w = geti<int>(g, "box_w", 128);
  w.box_h = geti<int>(g, "box_h", 1);
  w.box_n = geti<int>(g, "box_n", 1);
  w.box_rows = w.box_w * w.box_h * w.box_n;
  w.mul_w = geti<int>(g, "mul_w", 1);
  w.mul_h = geti<int>(g, "mul_h", 1);
  w.num_taps = geti<int>(g, "num_taps", 1);
  fill_ints(g, "tap_dw", w.tap_dw, tfos::kMaxTaps);
  fill_ints(g, "tap_dh", w.tap_dh, tfos::kMaxTaps);
  fill_ints(g, "tap_dc", w.tap_dc, tfos::kMaxTaps);
  fill_ints(g, "tap_out", w.tap_out, tfos::kMaxTaps);
  w.m_tiles = geti<int>(g, "m_tiles", 1);
  w.n_tiles = geti<int>(g, "n_tiles", 1);
  w.k_splits = geti<int>(g, "k_splits", 1);
  w.m_valid = geti<int>(g, "m_valid", 0);
  w.n_valid = geti<int>(g, "n_valid", 0);
  w.ldw = geti<int>(g, "ldw", 0);
  w.stem = geti<int>(g, "stem", 0);
  w.halo_boxes = geti<int>(g, "halo_boxes", 1);
  w.halo_rows = geti<int>(g, "halo_rows", 21);
  w.halo_hmul = geti<int>(g, "halo_hmul", 2);
  w.halo_h0 = geti<int>(g, "halo_h0", 0);
  w.halo_jmul = geti<int>(g, "halo_jmul", 2);
  w.halo_rowbytes = geti<int>(g, "halo_rowbytes", 2048);
  w.halo_nacc = geti<int>(g, "halo_nacc", 7);
  fill_ints(g, "acc_box", w.acc_box, tfos::kMaxTaps);
  fill_ints(g, "acc_row", w.acc_row, tfos::kMaxTaps);
  fill_ints(g, "box_dw", w.box_dw, 3);
  w.wide = geti<int>(g, "wide", 0);
  w.dw = reinterpret_cast<float*>(geti<uint64_t>(g, "dw", 0));
  TORCH_CHECK(w.dw != nullptr && w.ldw > 0, "igemm wgrad: dw/ldw");
  char err[512] = {0};
  tfos::IGemmPlan* p =
      tfos::igemm_plan_wgrad(parse_tmap(a), parse_tmap(b), w, bn, num_sms(), err, 512);
  TORCH_CHECK(p != nullptr, "igemm_plan_wgrad: ", err);
  return reinterpret_cast<int64_t>(p);
}

void plan_run(int64_t h) {
  check(tfos::igemm_run(reinterpret_cast<tfos::IGemmPlan*>(h), cur_stream()), "igemm_run");
}
void plan_set_reverse(int64_t h, bool flag) {
  tfos::igemm_plan_set_reverse(reinterpret_cast<tfos::IGemmPlan*>(h), flag ? 1 : 0);
}
void bn_set_row_reverse(bool flag) { tfos::bn_set_row_reverse(flag ? 1 : 0); }
void plan_free(int64_t h) { tfos::igemm_plan_free(reinterpret_cast<tfos::IGemmPlan*>(h)); }
py::dict plan_info(int64_t h) {
  auto* p = reinterpret_cast<tfos::IGemmPlan*>(h);
  py::dict d;
  d["kind"] = p->kind;
  d["bn"] = p->bn;
  d["grid"] = p->grid;
  d["total_work"] = p->total_work;
  d["cta_group"] = p->cta_group;
  return d;
}

// ------------------------------------------------------------- elementwise
void bn_stats(const Tensor& x, Tensor sum, Tensor sumsq) {
  need(x, torch::kBFloat16, "x");
  const int C = x.size(-1);
  check(tfos::bn_stats(x.data_ptr(), x.numel() / C, C, sum.data_ptr<float>(),
                       sumsq.data_ptr<float>(), cur_stream()),
        "bn_stats");
}
void bn_finalize(Tensor sum, Tensor sumsq, const Tensor& gamma, const Tensor& beta,
                 c10::optional<Tensor> rm, c10::optional<Tensor> rv, Tensor mean, Tensor invstd,
                 Tensor scale, Tensor shift, double count, double eps, double momentum) {
  check(tfos::bn_finalize(sum.data_ptr<float>(), sumsq.data_ptr<float>(), gamma.data_ptr<float>(),
                          beta.data_ptr<float>(), optf(rm), optf(rv), mean.data_ptr<float>(),
                          invstd.data_ptr<float>(), scale.data_ptr<float>(),
                          shift.data_ptr<float>(), gamma.numel(), count, eps, momentum,
                          cur_stream()),
        "bn_finalize");
}
void bn_inference_coeffs(const Tensor& gamma, const Tensor& beta, const Tensor& rm,
                         const Tensor& rv, Tensor scale, Tensor shift, double eps) {
  check(tfos::bn_inference_coeffs(gamma.data_ptr<float>(), beta.data_ptr<float>(),
                                  rm.data_ptr<float>(), rv.data_ptr<float>(),
                                  scale.data_ptr<float>(), shift.data_ptr<float>(), gamma.numel(),
                                  eps, cur_stream()),
        "bn_inference_coeffs");
}
void bn_apply(const Tensor& x, c10::optional<Tensor> residual, const Tensor& scale,
              const Tensor& shift, Tensor y, int act, c10::optional<Tensor> mask) {
  need(x, torch::kBFloat16, "x");
  need(y, torch::kBFloat16, "y");
  const int C = x.size(-1);
  TORCH_CHECK(C % 8 == 0, "bn_apply: C % 8");
  if (mask.has_value())
    TORCH_CHECK(mask->scalar_type() == torch::kUInt8 && mask->numel() * 8 == x.numel(),
                "bn_apply: mask must be uint8 with numel/8 entries");
  check(tfos::bn_apply(x.data_ptr(), optptr(residual), scale.data_ptr<float>(),
                       shift.data_ptr<float>(), y.data_ptr(),
                       mask.has_value() ? mask->data_ptr<uint8_t>() : nullptr, x.numel() / C, C,
                       act, cur_stream()),
        "bn_apply");
}
void bn_apply_finalize(const Tensor& x, c10::optional<Tensor> residual, Tensor y, int act,
                       c10::optional<Tensor> mask, const Tensor& sum, const Tensor& sumsq,
                       const Tensor& gamma, const Tensor& beta, c10::optional<Tensor> running_mean,
                       c10::optional<Tensor> running_var, Tensor mean, Tensor invstd, Tensor scale,
                       Tensor shift, double count, double eps, double momentum) {
  need(x, torch::kBFloat16, "x");
  need(y, torch::kBFloat16, "y");
  const int C = x.size(-1);
  TORCH_CHECK(C % 8 == 0 && sum.numel() == C && sumsq.numel() == C, "bn_apply_finalize: C");
  tfos::BnFinalize f;
  f.sum = sum.data_ptr<float>();
  f.sumsq = sumsq.data_ptr<float>();
  f.gamma = gamma.data_ptr<float>();
  f.beta = beta.data_ptr<float>();
  f.running_mean = running_mean.has_value() ? running_mean->data_ptr<float>() : nullptr;
  f.running_var = running_var.has_value() ? running_var->data_ptr<float>() : nullptr;
  f.mean = mean.data_ptr<float>();
  f.invstd = invstd.data_ptr<float>();
  f.scale = scale.data_ptr<float>();
  f.shift = shift.data_ptr<float>();
  f.count = static_cast<float>(count);
  f.eps = static_cast<float>(eps);
  f.momentum = static_cast<float>(momentum);
  check(tfos::bn_apply_finalize(x.data_ptr(), optptr(residual), y.data_ptr(),
                                mask.has_value() ? mask->data_ptr<uint8_t>() : nullptr,
                                x.numel() / C, C, act, f, cur_stream()),
        "bn_apply_finalize");
}
void bn_bwd_reduce(const Tensor& dy, const Tensor& x, c10::optional<Tensor> y, const Tensor& mean,
                   const Tensor& invstd, Tensor dgamma, Tensor dbeta, int relu,
                   c10::optional<Tensor> fscale, c10::optional<Tensor> fshift) {
  TORCH_CHECK(relu != 1 || (y.has_value() && y->defined()), "bn_bwd_reduce: relu=1 needs y");
  TORCH_CHECK(relu != 2 || (fscale.has_value() && fshift.has_value()), "bn_bwd_reduce: relu=2 needs scale/shift");
  need(dy, torch::kBFloat16, "dy");
  need(x, torch::kBFloat16, "x");
  const int C = x.size(-1);
  check(tfos::bn_bwd_reduce(dy.data_ptr(), x.data_ptr(), optptr(y), mean.data_ptr<float>(),
                            invstd.data_ptr<float>(), optf(fscale), optf(fshift), x.numel() / C, C,
                            relu, dgamma.data_ptr<float>(), dbeta.data_ptr<float>(), cur_stream()),
        "bn_bwd_reduce");
}
void bn_bwd_apply(const Tensor& dy, const Tensor& x, c10::optional<Tensor> y, const Tensor& gamma,
                  const Tensor& mean, const Tensor& invstd, const Tensor& dgamma,
                  const Tensor& dbeta, Tensor dx, c10::optional<Tensor> dres, int relu,
                  c10::optional<Tensor> fscale, c10::optional<Tensor> fshift,
                  c10::optional<Tensor> sum_g, c10::optional<Tensor> sum_gx) {
  const int C = x.size(-1);
  check(tfos::bn_bwd_apply(dy.data_ptr(), x.data_ptr(), optptr(y), gamma.data_ptr<float>(),
                           mean.data_ptr<float>(), invstd.data_ptr<float>(),
                           dgamma.data_ptr<float>(), dbeta.data_ptr<float>(), optf(fscale),
                           optf(fshift), dx.data_ptr(), const_cast<void*>(optptr(dres)),
                           x.numel() / C, C, relu, optf(sum_g), optf(sum_gx), cur_stream()),
        "bn_bwd_apply");
}
void stem_bn_relu_pool_fwd(const Tensor& x, Tensor y, Tensor idx, const Tensor& sum,
                           const Tensor& sumsq, const Tensor& gamma, const Tensor& beta,
                           Tensor running_mean, Tensor running_var, Tensor mean, Tensor invstd,
                           Tensor scale, Tensor shift, double count, double eps, double momentum) {
  need(x, torch::kBFloat16, "x");
  need(y, torch::kBFloat16, "y");
  TORCH_CHECK(x.dim() == 4 && y.dim() == 4 && idx.numel() == y.numel(), "stem_bn_relu_pool_fwd: shapes");
  tfos::BnFinalize f;
  f.sum = sum.data_ptr<float>();
  f.sumsq = sumsq.data_ptr<float>();
  f.gamma = gamma.data_ptr<float>();
  f.beta = beta.data_ptr<float>();
  f.running_mean = running_mean.data_ptr<float>();
  f.running_var = running_var.data_ptr<float>();
  f.mean = mean.data_ptr<float>();
  f.invstd = invstd.data_ptr<float>();
  f.scale = scale.data_ptr<float>();
  f.shift = shift.data_ptr<float>();
  f.count = static_cast<float>(count);
  f.eps = static_cast<float>(eps);
  f.momentum = static_cast<float>(momentum);
  check(tfos::stem_bn_relu_pool_fwd(x.data_ptr(), y.data_ptr(), idx.data_ptr<uint8_t>(), x.size(0),
                                    x.size(1), x.size(2), x.size(3), y.size(1), y.size(2), f,
                                    cur_stream()),
        "stem_bn_relu_pool_fwd");
}
void stem_pool_bn_bwd(const Tensor& dy_pool, const Tensor& idx, const Tensor& x, const Tensor& gamma,
                      const Tensor& mean, const Tensor& invstd, const Tensor& scale,
                      const Tensor& shift, Tensor dgamma, Tensor dbeta, Tensor dx) {
  need(dy_pool, torch::kBFloat16, "dy_pool");
  need(x, torch::kBFloat16, "x");
  need(dx, torch::kBFloat16, "dx");
  TORCH_CHECK(x.dim() == 4 && dy_pool.dim() == 4 && dx.numel() == x.numel(), "stem_pool_bn_bwd: shapes");
  check(tfos::stem_pool_bn_bwd(dy_pool.data_ptr(), idx.data_ptr<uint8_t>(), x.data_ptr(),
                               gamma.data_ptr<float>(), mean.data_ptr<float>(),
                               invstd.data_ptr<float>(), scale.data_ptr<float>(),
                               shift.data_ptr<float>(), dgamma.data_ptr<float>(),
                               dbeta.data_ptr<float>(), dx.data_ptr(), x.size(0), x.size(1),
                               x.size(2), x.size(3), dy_pool.size(1), dy_pool.size(2), cur_stream()),
        "stem_pool_bn_bwd");
}
void add_act(const Tensor& a, c10::optional<Tensor> b, Tensor out, int act) {
  need(a, torch::kBFloat16, "a");
  TORCH_CHECK(a.numel() % 8 == 0, "add_act: numel % 8");
  check(tfos::add_act(a.data_ptr(), optptr(b), out.data_ptr(), a.numel(), act, cur_stream()),
        "add_act");
}
void relu_bwd(const Tensor& dy, const Tensor& y, Tensor dx) {
  need(dy, torch::kBFloat16, "dy");
  TORCH_CHECK(dy.numel() % 8 == 0, "relu_bwd: numel % 8");
  check(tfos::relu_bwd(dy.data_ptr(), y.data_ptr(), dx.data_ptr(), dy.numel(), cur_stream()),
        "relu_bwd");
}
void colsum(const Tensor& x, Tensor out) {
  need(x, torch::kBFloat16, "x");
  const int C = x.size(-1);
  TORCH_CHECK(C % 8 == 0, "colsum: C % 8");
  check(tfos::colsum(x.data_ptr(), x.numel() / C, C, out.data_ptr<float>(), cur_stream()),
        "colsum");
}
void maxpool_fwd(const Tensor& x, Tensor y, c10::optional<Tensor> idx, int k, int stride, int pad) {
  need(x, torch::kBFloat16, "x");
  check(tfos::maxpool_fwd(x.data_ptr(), y.data_ptr(),
                          idx.has_value() ? idx->data_ptr<uint8_t>() : nullptr, x.size(0),
                          x.size(1), x.size(2), x.size(3), y.size(1), y.size(2), k, stride, pad,
                          cur_stream()),
        "maxpool_fwd");
}
void maxpool_bwd(const Tensor& dy, const Tensor& idx, Tensor dx, int k, int stride, int pad) {
  check(tfos::maxpool_bwd(dy.data_ptr(), idx.data_ptr<uint8_t>(), dx.data_ptr(), dx.size(0),
                          dx.size(1), dx.size(2), dx.size(3), dy.size(1), dy.size(2), k, stride,
                          pad, cur_stream()),
        "maxpool_bwd");
}
void avgpool_fwd(const Tensor& x, Tensor y) {
  need(x, torch::kBFloat16, "x");
  const int N = x.size(0), C = x.size(-1);
  check(tfos::avgpool_fwd(x.data_ptr(), y.data_ptr(), N, x.numel() / N / C, C, cur_stream()),
        "avgpool_fwd");
}
void avgpool_bwd(const Tensor& dy, Tensor dx) {
  const int N = dx.size(0), C = dx.size(-1);
  check(tfos::avgpool_bwd(dy.data_ptr(), dx.data_ptr(), N, dx.numel() / N / C, C, cur_stream()),
        "avgpool_bwd");
}
void softmax_xent(const Tensor& logits, const Tensor& labels, c10::optional<Tensor> dlogits,
                  Tensor loss_sum, c10::optional<Tensor> correct, int V, double scale) {
  TORCH_CHECK(logits.is_cuda() && logits.is_contiguous(), "logits");
  need(labels, torch::kInt32, "labels");
  const int ld = logits.size(-1);
  const long long rows = logits.numel() / ld;
  const bool fp32 = logits.scalar_type() == torch::kFloat32;
  TORCH_CHECK(fp32 || logits.scalar_type() == torch::kBFloat16, "logits dtype");
  int ldd = 0;
  void* dl = nullptr;
  if (dlogits.has_value() && dlogits->defined()) {
    need(*dlogits, torch::kBFloat16, "dlogits");
    ldd = dlogits->size(-1);
    dl = dlogits->data_ptr();
  }
  check(tfos::softmax_xent(logits.data_ptr(), fp32 ? 1 : 0, labels.data_ptr<int>(), dl,
                           loss_sum.data_ptr<float>(), optf(correct), rows, V, ld, ldd, scale,
                           cur_stream()),
        "softmax_xent");
}
void decode_normalize(const Tensor& in, Tensor out, int wofs, std::vector<double> mean,
                      std::vector<double> stdv) {
  need(in, torch::kUInt8, "in");
  need(out, torch::kBFloat16, "out");
  float m[3], is[3];
  for (int i = 0; i < 3; ++i) {
    m[i] = mean[i % mean.size()];
    is[i] = 1.f / stdv[i % stdv.size()];
  }
  check(tfos::decode_normalize(in.data_ptr<uint8_t>(), out.data_ptr(), in.size(0), in.size(1),
                               in.size(2), in.size(3), out.size(2), out.size(3), wofs, m, is,
                               cur_stream()),
        "decode_normalize");
}
void cast_f32_bf16(const Tensor& in, Tensor out) {
  need(in, torch::kFloat32, "in");
  need(out, torch::kBFloat16, "out");
  check(tfos::cast_f32_bf16(in.data_ptr<float>(), out.data_ptr(), in.numel(), cur_stream()),
        "cast_f32_bf16");
}
void conv3x3_c1_fwd(const Tensor& x, const Tensor& w, c10::optional<Tensor> bias, Tensor y,
                    bool relu) {
  need(x, torch::kBFloat16, "x");
  check(tfos::conv3x3_c1_fwd(x.data_ptr(), w.data_ptr(), optf(bias), y.data_ptr(), x.size(0),
                             x.size(1), x.size(2), y.size(3), relu ? 1 : 0, cur_stream()),
        "conv3x3_c1_fwd");
}
void conv3x3_c1_wgrad(const Tensor& x, const Tensor& dy, Tensor dw, c10::optional<Tensor> dbias) {
  check(tfos::conv3x3_c1_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr<float>(), optf(dbias),
                               x.size(0), x.size(1), x.size(2), dy.size(3), cur_stream()),
        "conv3x3_c1_wgrad");
}
void depthwise3x3_fwd(const Tensor& x, const Tensor& w, c10::optional<Tensor> bias, Tensor y,
                      int stride, int act) {
  need(x, torch::kBFloat16, "x");
  check(tfos::depthwise3x3_fwd(x.data_ptr(), w.data_ptr(), optf(bias), y.data_ptr(), x.size(0),
                               x.size(1), x.size(2), x.size(3), stride, act, cur_stream()),
        "depthwise3x3_fwd");
}
void copy_channels(const Tensor& src, Tensor dst, int C, int src_off, int dst_off) {
  need(src, torch::kBFloat16, "src");
  need(dst, torch::kBFloat16, "dst");
  const int sld = src.size(-1), dld = dst.size(-1);
  TORCH_CHECK(C % 8 == 0 && src_off % 8 == 0 && dst_off % 8 == 0 && sld % 8 == 0 && dld % 8 == 0,
              "copy_channels: channel counts / offsets must be multiples of 8");
  TORCH_CHECK(src.numel() / sld == dst.numel() / dld, "copy_channels: pixel counts differ");
  check(tfos::copy_channels(src.data_ptr(), dst.data_ptr(), src.numel() / sld, C, sld, src_off,
                            dld, dst_off, cur_stream()),
        "copy_channels");
}
void pixel_xent(const Tensor& logits, const Tensor& labels, c10::optional<Tensor> dlogits,
                Tensor loss_sum, c10::optional<Tensor> correct, int V, double scale) {
  need(logits, torch::kBFloat16, "logits");
  need(labels, torch::kInt32, "labels");
  TORCH_CHECK(logits.size(-1) == 8 && V <= 8, "pixel_xent expects 8 padded channels");
  check(tfos::pixel_xent(logits.data_ptr(), labels.data_ptr<int>(),
                         const_cast<void*>(optptr(dlogits)), loss_sum.data_ptr<float>(),
                         optf(correct), logits.numel() / 8, V, scale, cur_stream()),
        "pixel_xent");
}

// --------------------------------------------------- collectives / optimizer
void fill_ptrs(const std::vector<uint64_t>& v, void** dst, int world) {
  TORCH_CHECK(static_cast<int>(v.size()) >= world, "peer pointer list shorter than world");
  for (int i = 0; i < tfos::kMaxRanks; ++i)
    dst[i] = i < world ? reinterpret_cast<void*>(v[i]) : nullptr;
}

void allreduce_opt(const py::dict& d) {
  tfos::AllreduceOptArgs a;
  std::memset(&a, 0, sizeof(a));
  a.master = reinterpret_cast<float*>(d["master"].cast<uint64_t>());
  a.state1 = reinterpret_cast<float*>(geti<uint64_t>(d, "state1", 0));
  a.state2 = reinterpret_cast<float*>(geti<uint64_t>(d, "state2", 0));
  a.hyper = reinterpret_cast<const float*>(d["hyper"].cast<uint64_t>());
  a.begin = d["begin"].cast<long long>();
  a.end = d["end"].cast<long long>();
  a.decay_end = geti<long long>(d, "decay_end", 0);
  a.state_offset = 0;
  a.world = geti<int>(d, "world", 1);
  a.rank = geti<int>(d, "rank", 0);
  a.slot = geti<int>(d, "slot", 0);
  a.zero_grads = geti<int>(d, "zero_grads", 0);
  fill_ptrs(d["grads"].cast<std::vector<uint64_t>>(), reinterpret_cast<void**>(a.grads), a.world);
  fill_ptrs(d["weights"].cast<std::vector<uint64_t>>(), reinterpret_cast<void**>(a.weights),
            a.world);
  if (a.world > 1) {
    fill_ptrs(d["flags"].cast<std::vector<uint64_t>>(), reinterpret_cast<void**>(a.flags), a.world);
    a.epoch = reinterpret_cast<uint32_t*>(d["epoch"].cast<uint64_t>());
    a.block_counter = reinterpret_cast<uint32_t*>(d["block_counter"].cast<uint64_t>());
  }
  if (d.contains("aux32")) {
    fill_ptrs(d["aux32"].cast<std::vector<uint64_t>>(), reinterpret_cast<void**>(a.aux32), a.world);
    a.aux_begin = geti<long long>(d, "aux_begin", 0);
  }
  a.grads_mc = reinterpret_cast<const float*>(geti<uint64_t>(d, "grads_mc", 0));
  a.weights_mc = reinterpret_cast<__nv_bfloat16*>(geti<uint64_t>(d, "weights_mc", 0));
  a.aux32_mc = reinterpret_cast<float*>(geti<uint64_t>(d, "aux32_mc", 0));
  const int opt = geti<int>(d, "opt", 0);
  const int grid = geti<int>(d, "grid", 32);
  check(tfos::allreduce_opt(a, opt, grid, cur_stream(), geti<int>(d, "phase", 0)), "allreduce_opt");
}

tfos::BcastArgs parse_bcast(const py::dict& d) {
  tfos::BcastArgs a;
  std::memset(&a, 0, sizeof(a));
  a.world = geti<int>(d, "world", 1);
  a.rank = geti<int>(d, "rank", 0);
  a.root = geti<int>(d, "root", 0);
  a.slot = geti<int>(d, "slot", 0);
  a.bytes = geti<long long>(d, "bytes", 0);
  if (d.contains("bufs"))
    fill_ptrs(d["bufs"].cast<std::vector<uint64_t>>(), a.bufs, a.world);
  fill_ptrs(d["flags"].cast<std::vector<uint64_t>>(), reinterpret_cast<void**>(a.flags), a.world);
  a.epoch = reinterpret_cast<uint32_t*>(d["epoch"].cast<uint64_t>());
  a.block_counter = reinterpret_cast<uint32_t*>(d["block_counter"].cast<uint64_t>());
  return a;
}
void bcast_pull(const py::dict& d) {
  check(tfos::bcast_pull(parse_bcast(d), geti<int>(d, "grid", 32), cur_stream()), "bcast_pull");
}
void set_flag_timeout_ms(double ms) {
  check(tfos::set_flag_timeout_ns(static_cast<unsigned long long>(ms * 1e6)), "set_flag_timeout_ms");
}
void flag_barrier(const py::dict& d) {
  check(tfos::flag_barrier(parse_bcast(d), cur_stream()), "flag_barrier");
}
void ps_push_dense(uint64_t w_ps, const Tensor& g, const Tensor& hyper) {
  need(g, torch::kFloat32, "g");
  check(tfos::ps_push_dense(reinterpret_cast<float*>(w_ps), g.data_ptr<float>(), g.numel(),
                            hyper.data_ptr<float>(), cur_stream()),
        "ps_push_dense");
}
void ps_push_sparse(uint64_t w_ps, const Tensor& g_rows, const Tensor& idx, const Tensor& hyper) {
  need(g_rows, torch::kFloat32, "g_rows");
  need(idx, torch::kInt32, "idx");
  check(tfos::ps_push_sparse(reinterpret_cast<float*>(w_ps), g_rows.data_ptr<float>(),
                             idx.data_ptr<int>(), g_rows.size(0), g_rows.size(1),
                             hyper.data_ptr<float>(), cur_stream()),
        "ps_push_sparse");
}
void ps_pull(uint64_t w_ps, c10::optional<Tensor> w_local, c10::optional<Tensor> w_bf16,
             long long n) {
  check(tfos::ps_pull(reinterpret_cast<const float*>(w_ps), optf(w_local),
                      const_cast<void*>(optptr(w_bf16)), n, cur_stream()),
        "ps_pull");
}

template <typename T>
T* ptr_of(const py::dict& d, const char* k) {
  return reinterpret_cast<T*>(geti<uint64_t>(d, k, 0));
}
void ps_apply(const py::dict& d) {
  tfos::PsApplyArgs a;
  std::memset(&a, 0, sizeof(a));
  a.master = ptr_of<float>(d, "master");
  a.state1 = ptr_of<float>(d, "state1");
  a.state2 = ptr_of<float>(d, "state2");
  a.wbf16 = ptr_of<__nv_bfloat16>(d, "wbf16");
  a.slot = ptr_of<const float>(d, "slot");
  a.hyper = ptr_of<const float>(d, "hyper");
  a.n = geti<long long>(d, "n", 0);
  a.lo = geti<long long>(d, "lo", 0);
  a.decay_end = geti<long long>(d, "decay_end", 0);
  a.ema_begin = geti<long long>(d, "ema_begin", 0);
  a.applied_flag = ptr_of<uint32_t>(d, "applied_flag");
  a.seq = geti<uint32_t>(d, "seq", 0);
  a.block_counter = ptr_of<uint32_t>(d, "block_counter");
  check(tfos::ps_apply(a, geti<int>(d, "opt", 0), cur_stream()), "ps_apply");
}
void ps_push_slot(const py::dict& d) {
  tfos::PsPushArgs a;
  std::memset(&a, 0, sizeof(a));
  a.slot = ptr_of<float>(d, "slot");
  a.grads = ptr_of<const float>(d, "grads");
  a.running = ptr_of<const float>(d, "running");
  a.running_pulled = ptr_of<const float>(d, "running_pulled");
  a.n = geti<long long>(d, "n", 0);
  a.lo = geti<long long>(d, "lo", 0);
  a.total = geti<long long>(d, "total", 0);
  a.applied_flag = ptr_of<const uint32_t>(d, "applied_flag");
  a.need_applied = geti<uint32_t>(d, "need_applied", 0);
  a.ready_flag = ptr_of<uint32_t>(d, "ready_flag");
  a.seq = geti<uint32_t>(d, "seq", 0);
  a.block_counter = ptr_of<uint32_t>(d, "block_counter");
  check(tfos::ps_push_slot(a, geti<int>(d, "grid", 64), cur_stream()), "ps_push_slot");
}
void ps_pull_model(const py::dict& d) {
  tfos::PsPullArgs a;
  std::memset(&a, 0, sizeof(a));
  a.wbf16 = ptr_of<const __nv_bfloat16>(d, "wbf16");
  a.master = ptr_of<const float>(d, "master");
  a.weights = ptr_of<__nv_bfloat16>(d, "weights");
  a.aux32 = ptr_of<float>(d, "aux32");
  a.running = ptr_of<float>(d, "running");
  a.running_pulled = ptr_of<float>(d, "running_pulled");
  a.n = geti<long long>(d, "n", 0);
  a.lo = geti<long long>(d, "lo", 0);
  a.total = geti<long long>(d, "total", 0);
  a.decay_end = geti<long long>(d, "decay_end", 0);
  check(tfos::ps_pull_model(a, geti<int>(d, "grid", 64), cur_stream()), "ps_pull_model");
}

// ------------------------------------------------ symmetric memory (CUDA IPC)
// A symmetric buffer is a plain cudaMalloc allocation whose IPC handle is
// exchanged through the reservation/node_meta channel; peers map it with
// cudaIpcOpenMemHandle and address it directly from kernels.
py::tuple symm_alloc(long long bytes) {
  void* p = nullptr;
  check(cudaMalloc(&p, bytes), "symm_alloc cudaMalloc");
  check(cudaMemset(p, 0, bytes), "symm_alloc memset");
  cudaIpcMemHandle_t h;
  check(cudaIpcGetMemHandle(&h, p), "cudaIpcGetMemHandle");
  return py::make_tuple(reinterpret_cast<uint64_t>(p),
                        py::bytes(reinterpret_cast<const char*>(&h), sizeof(h)));
}
uint64_t symm_open(const std::string& handle) {
  TORCH_CHECK(handle.size() == sizeof(cudaIpcMemHandle_t), "bad IPC handle size");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle.data(), sizeof(h));
  void* p = nullptr;
  check(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
  return reinterpret_cast<uint64_t>(p);
}
void symm_close(uint64_t p) { cudaIpcCloseMemHandle(reinterpret_cast<void*>(p)); }
void symm_free(uint64_t p) { cudaFree(reinterpret_cast<void*>(p)); }
Tensor tensor_from_ptr(uint64_t ptr, std::vector<int64_t> sizes, const std::string& dtype) {
  c10::ScalarType dt = torch::kUInt8;
  if (dtype == "bf16") dt = torch::kBFloat16;
  else if (dtype == "f32") dt = torch::kFloat32;
  else if (dtype == "i32") dt = torch::kInt32;
  else if (dtype == "u8") dt = torch::kUInt8;
  else TORCH_CHECK(false, "unknown dtype ", dtype);
  int dev = 0;
  cudaGetDevice(&dev);
  // target_device: the pointer may be a peer's memory mapped into this process (CUDA IPC or an
  // imported cuMem allocation, whose pointer attributes name the OWNING device); it is addressed
  // from the current device, so the tensor is declared there and torch's ownership check skipped
  const c10::Device here(torch::kCUDA, static_cast<c10::DeviceIndex>(dev));
  return at::for_blob(reinterpret_cast<void*>(ptr), sizes)
      .options(torch::TensorOptions().dtype(dt).device(here))
      .target_device(here)
      .make_tensor();
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "tensorflowonspark_b200 native sm_100a kernels";
  m.def("igemm_plan_fwd", &plan_fwd);
  m.def("igemm_plan_wgrad", &plan_wgrad);
  m.def("igemm_run", &plan_run);
  m.def("igemm_free", &plan_free);
  m.def("igemm_set_reverse", &plan_set_reverse);
  m.def("bn_set_row_reverse", &bn_set_row_reverse);
  m.def("igemm_info", &plan_info);
  m.def("bn_stats", &bn_stats);
  m.def("bn_finalize", &bn_finalize);
  m.def("bn_inference_coeffs", &bn_inference_coeffs);
  m.def("bn_apply", &bn_apply, py::arg("x"), py::arg("residual"), py::arg("scale"),
        py::arg("shift"), py::arg("y"), py::arg("act"), py::arg("mask") = py::none());
  m.def("bn_apply_finalize", &bn_apply_finalize);
  m.def("bn_bwd_reduce", &bn_bwd_reduce);
  m.def("bn_bwd_apply", &bn_bwd_apply, py::arg("dy"), py::arg("x"), py::arg("y"), py::arg("gamma"),
        py::arg("mean"), py::arg("invstd"), py::arg("dgamma"), py::arg("dbeta"), py::arg("dx"),
        py::arg("dres"), py::arg("relu"), py::arg("fscale"), py::arg("fshift"),
        py::arg("sum_g") = py::none(), py::arg("sum_gx") = py::none());
  m.def("stem_bn_relu_pool_fwd", &stem_bn_relu_pool_fwd);
  m.def("stem_pool_bn_bwd", &stem_pool_bn_bwd);
  m.def("add_act", &add_act);
  m.def("relu_bwd", &relu_bwd);
  m.def("colsum", &colsum);
  m.def("maxpool_fwd", &maxpool_fwd);
  m.def("maxpool_bwd", &maxpool_bwd);
  m.def("avgpool_fwd", &avgpool_fwd);
  m.def("avgpool_bwd", &avgpool_bwd);
  m.def("softmax_xent", &softmax_xent);
  m.def("decode_normalize", &decode_normalize);
  m.def("cast_f32_bf16", &cast_f32_bf16);
  m.def("conv3x3_c1_fwd", &conv3x3_c1_fwd);
  m.def("conv3x3_c1_wgrad", &conv3x3_c1_wgrad);
  m.def("depthwise3x3_fwd", &depthwise3x3_fwd);
  m.def("copy_channels", &copy_channels);
  m.def("pixel_xent", &pixel_xent);
  m.def("allreduce_opt", &allreduce_opt);
  m.def("bcast_pull", &bcast_pull);
  m.def("flag_barrier", &flag_barrier);
  m.def("set_flag_timeout_ms", &set_flag_timeout_ms);
  m.def("ps_push_dense", &ps_push_dense);
  m.def("ps_push_sparse", &ps_push_sparse);
  m.def("ps_pull", &ps_pull);
  m.def("ps_apply", &ps_apply);
  m.def("ps_push_slot", &ps_push_slot);
  m.def("ps_pull_model", &ps_pull_model);
  m.def("symm_alloc", &symm_alloc);
  m.def("symm_open", &symm_open);
  m.def("symm_close", &symm_close);
  m.def("symm_free", &symm_free);
  m.def("tensor_from_ptr", &tensor_from_ptr);
  m.def("num_sms", &num_sms);
  tfos::bind_feed(m);
  tfos::bind_tfrecord(m);
  tfos::bind_vmm(m);
}
