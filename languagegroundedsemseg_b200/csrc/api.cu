// C-ABI dispatch for the convolution entry points (include/lgs_b200.h).
#include <string>

#include "common.cuh"

namespace lgs {
int conv_fwd_simt(const void* in, int64_t n_in, int c_in, const void* w, int w_layout, int K, int c_out,
                  const int32_t* table, int64_t n_out, int reverse_k, const float* bias, void* out, int dtype,
                  cudaStream_t stream);
int conv_wgrad_simt(const void* in, int c_in, const void* gout, int64_t n_out, int c_out, const int32_t* table, int K,
                    float* gw, int dtype, cudaStream_t stream);
// tcgen05 path (conv_tc.cu): returns LGS_E_UNSUPPORTED when the shape is outside what it was built for
int conv_fwd_tc(const void* in, int64_t n_in, int c_in, const void* w, int K, int c_out, const int32_t* table,
                int64_t n_out, int reverse_k, const float* bias, void* out, int dtype, int precise, cudaStream_t stream);
int conv_tc_shape_ok(int c_in, int c_out, int dtype);
int weight_prep(const float* w, int K, int c_in, int c_out, int nsplit, void* fwd, void* bwd, int dtype,
                cudaStream_t stream);
int weight_prep_batch(const int64_t* desc, int n_layers, int64_t total_tiles, int nsplit, int dtype, cudaStream_t stream);
int conv_wgrad_tc(const void* in, int64_t n_in, int c_in, const void* gout, int64_t n_out, int c_out,
                  const int32_t* table, int K, float* gw, int dtype, cudaStream_t stream);
bool tc_built();
// bf16x3 path (conv_bx3.cu)
int conv_bx3_shape_ok(int c_in, int c_out);
int conv_fwd_bx3(const void* in, int c_in, const void* in2, int c_in2, const void* w, int K, int c_out, const int32_t* table,
                 int64_t n_out, int reverse_k, const float* bias, float* out, double* stats, cudaStream_t stream);
int bx3_tune(const char* key, int value);
int weight_prep_bx3(const float* w, int K, int c_in, int c_out, void* fwd, void* bwd, cudaStream_t stream);
int weight_prep_bx3_batch(const int64_t* desc, int n_layers, int64_t total_tiles, cudaStream_t stream);
int conv_wgrad_stem(const void* in, int c_in, const void* gout, int64_t n_out, int c_out, const int32_t* table, int K, float* gw,
                    int dtype, cudaStream_t stream);
// neighbourhood-cache path (conv_nb.cu, nbplan.cu)
int conv_nb_shape_ok(int c_in, int c_in2, int c_out, int K);
int conv_fwd_nb(const void* in, int c_in, const void* in2, int c_in2, const void* w, int K, int c_out, const void* d_plan,
                int64_t n_out, int reverse_k, const float* bias, const float* addend, float* out, double* stats, cudaStream_t stream);
}  // namespace lgs

using namespace lgs;

extern "C" {

int lgs_has_tc(void) { return tc_built() ? 1 : 0; }

int lgs_tune(const char* key, int32_t value) {
  if (!key) return fail(LGS_E_INVALID, "lgs_tune: null key");
  if (std::string(key) == "pdl") {
    g_pdl.store(value ? 1 : 0);
    return LGS_OK;
  }
  if (bx3_tune(key, value)) return LGS_OK;
  if (nb_tune(key, value)) return LGS_OK;
  return fail(LGS_E_INVALID, "lgs_tune: unknown key %s", key);
}

int lgs_conv_tc_supported(int32_t c_in, int32_t c_out, int32_t dtype) { return conv_tc_shape_ok(c_in, c_out, dtype); }

int64_t lgs_weight_bx3_elems(int32_t K, int32_t c_rows, int32_t c_reduced) {
  return int64_t(K) * c_rows * ((c_reduced + 31) / 32) * 64;
}

int lgs_weight_prep(const float* d_weight, int32_t K, int32_t c_in, int32_t c_out, int32_t nsplit, void* d_fwd,
                    void* d_bwd, int32_t dtype, void* stream_) {
  LGS_TRACE("lgs_weight_prep %p %d %d %d %d %p %p %d %p", (const void*)d_weight, (int)K, (int)c_in, (int)c_out, (int)nsplit, (const void*)d_fwd, (const void*)d_bwd, (int)dtype, (const void*)stream_);
  if (K < 1 || c_in < 1 || c_out < 1 || nsplit < 1 || nsplit > 3 || !d_weight || (!d_fwd && !d_bwd))
    return fail(LGS_E_INVALID, "lgs_weight_prep: bad arguments");
  if (dtype != LGS_F32 && dtype != LGS_BF16) return fail(LGS_E_INVALID, "lgs_weight_prep: dtype %d", dtype);
  if (nsplit == 3) {
    if (dtype != LGS_F32) return fail(LGS_E_INVALID, "lgs_weight_prep: nsplit 3 (LGS_W_BX3) serves LGS_F32 features");
    return weight_prep_bx3(d_weight, K, c_in, c_out, d_fwd, d_bwd, static_cast<cudaStream_t>(stream_));
  }
  return weight_prep(d_weight, K, c_in, c_out, nsplit, d_fwd, d_bwd, dtype, static_cast<cudaStream_t>(stream_));
}

int lgs_weight_prep_batch(const int64_t* d_desc, int32_t n_layers, int64_t total_tiles, int32_t nsplit, int32_t dtype,
                          void* stream_) {
  LGS_TRACE("lgs_weight_prep_batch %p %d %lld %d %d %p", (const void*)d_desc, (int)n_layers, (long long)total_tiles, (int)nsplit, (int)dtype, (const void*)stream_);
  if (n_layers < 0 || total_tiles < 0 || total_tiles >= (int64_t(1) << 31) || nsplit < 1 || nsplit > 3 ||
      (n_layers > 0 && !d_desc))
    return fail(LGS_E_INVALID, "lgs_weight_prep_batch: bad arguments");
  if (dtype != LGS_F32 && dtype != LGS_BF16) return fail(LGS_E_INVALID, "lgs_weight_prep_batch: dtype %d", dtype);
  if (nsplit == 3) {
    if (dtype != LGS_F32) return fail(LGS_E_INVALID, "lgs_weight_prep_batch: nsplit 3 (LGS_W_BX3) serves LGS_F32 features");
    return weight_prep_bx3_batch(d_desc, n_layers, total_tiles, static_cast<cudaStream_t>(stream_));
  }
  return weight_prep_batch(d_desc, n_layers, total_tiles, nsplit, dtype, static_cast<cudaStream_t>(stream_));
}

int lgs_conv_fwd(const void* d_in, int64_t n_in, int32_t c_in, const void* d_weight, int32_t weight_layout, int32_t K,
                 int32_t c_out,
                 const int32_t* d_table, int64_t n_out, int32_t reverse_k, const float* d_bias, void* d_out,
                 int32_t dtype, int32_t algo, void* stream_) {
  LGS_TRACE("lgs_conv_fwd %p %lld %d %p %d %d %d %p %lld %d %p %p %d %d %p", (const void*)d_in, (long long)n_in, (int)c_in, (const void*)d_weight, (int)weight_layout, (int)K, (int)c_out, (const void*)d_table, (long long)n_out, (int)reverse_k, (const void*)d_bias, (const void*)d_out, (int)dtype, (int)algo, (const void*)stream_);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_in < 0 || n_out < 0 || c_in < 1 || c_out < 1 || K < 1 || K > 27)
    return fail(LGS_E_INVALID, "lgs_conv_fwd: bad sizes n_in=%lld n_out=%lld c_in=%d c_out=%d K=%d", (long long)n_in,
                (long long)n_out, c_in, c_out, K);
  if (!d_table && (K != 1 || n_in != n_out))
    return fail(LGS_E_INVALID, "lgs_conv_fwd: NULL table needs K == 1 and n_in == n_out");
  if (dtype != LGS_F32 && dtype != LGS_BF16) return fail(LGS_E_INVALID, "lgs_conv_fwd: dtype %d", dtype);
  if (weight_layout != LGS_W_KCN && weight_layout != LGS_W_KNC && weight_layout != LGS_W_KNC_SPLIT && weight_layout != LGS_W_BX3)
    return fail(LGS_E_INVALID, "lgs_conv_fwd: weight_layout %d", weight_layout);
  if (algo == LGS_ALGO_BX3 || weight_layout == LGS_W_BX3) {
    if (algo != LGS_ALGO_BX3 || weight_layout != LGS_W_BX3 || dtype != LGS_F32)
      return fail(LGS_E_INVALID, "lgs_conv_fwd: LGS_ALGO_BX3 needs LGS_W_BX3 weights and LGS_F32 features");
    if ((n_out && (!d_in && n_in)) || !d_weight || (n_out && !d_out)) return fail(LGS_E_INVALID, "lgs_conv_fwd: null pointer");
    const int rc = conv_fwd_bx3(d_in, c_in, nullptr, 0, d_weight, K, c_out, d_table, n_out, reverse_k, d_bias,
                                static_cast<float*>(d_out), nullptr, stream);
    if (rc == LGS_E_UNSUPPORTED) return fail(rc, "lgs_conv_fwd: shape %d->%d not supported by LGS_ALGO_BX3", c_in, c_out);
    return rc;
  }
  if (algo == LGS_ALGO_TC3) {
    // 3xTF32: hi/lo-split K-major weights only; the caller checks lgs_conv_tc_supported() before choosing this form
    if (weight_layout != LGS_W_KNC_SPLIT || dtype != LGS_F32)
      return fail(LGS_E_INVALID, "lgs_conv_fwd: LGS_ALGO_TC3 needs LGS_W_KNC_SPLIT weights and LGS_F32 features");
    const int rc = conv_fwd_tc(d_in, n_in, c_in, d_weight, K, c_out, d_table, n_out, reverse_k, d_bias, d_out, dtype, 1,
                               stream);
    if (rc == LGS_E_UNSUPPORTED) return fail(rc, "lgs_conv_fwd: shape %d->%d not supported by LGS_ALGO_TC3", c_in, c_out);
    return rc;
  }
  if (weight_layout == LGS_W_KNC_SPLIT) return fail(LGS_E_INVALID, "lgs_conv_fwd: split weights need LGS_ALGO_TC3");
  if ((n_out && (!d_in && n_in)) || !d_weight || (n_out && !d_out)) return fail(LGS_E_INVALID, "lgs_conv_fwd: null pointer");
  if (algo == LGS_ALGO_TC) {
    const int rc = weight_layout == LGS_W_KNC
                       ? conv_fwd_tc(d_in, n_in, c_in, d_weight, K, c_out, d_table, n_out, reverse_k, d_bias, d_out,
                                     dtype, 0, stream)
                       : LGS_E_UNSUPPORTED;
    if (rc != LGS_E_UNSUPPORTED) return rc;
    // shape outside the tensor-core kernel's envelope (e.g. c_in = 3): the SIMT kernel takes it
  } else if (algo != LGS_ALGO_SIMT) {
    return fail(LGS_E_INVALID, "lgs_conv_fwd: algo %d", algo);
  }
  return conv_fwd_simt(d_in, n_in, c_in, d_weight, weight_layout, K, c_out, d_table, n_out, reverse_k, d_bias, d_out,
                       dtype, stream);
}

int lgs_conv_wgrad(const void* d_in, int64_t n_in, int32_t c_in, const void* d_grad_out, int64_t n_out, int32_t c_out,
                   const int32_t* d_table, int32_t K, float* d_grad_w, int32_t dtype, int32_t algo, void* stream_) {
  LGS_TRACE("lgs_conv_wgrad %p %lld %d %p %lld %d %p %d %p %d %d %p", (const void*)d_in, (long long)n_in, (int)c_in, (const void*)d_grad_out, (long long)n_out, (int)c_out, (const void*)d_table, (int)K, (const void*)d_grad_w, (int)dtype, (int)algo, (const void*)stream_);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_in < 0 || n_out < 0 || c_in < 1 || c_out < 1 || K < 1 || K > 27)
    return fail(LGS_E_INVALID, "lgs_conv_wgrad: bad sizes");
  if (!d_table && (K != 1 || n_in != n_out))
    return fail(LGS_E_INVALID, "lgs_conv_wgrad: NULL table needs K == 1 and n_in == n_out");
  if (dtype != LGS_F32 && dtype != LGS_BF16) return fail(LGS_E_INVALID, "lgs_conv_wgrad: dtype %d", dtype);
  if (!d_grad_w) return fail(LGS_E_INVALID, "lgs_conv_wgrad: null pointer");
  if (algo == LGS_ALGO_TC) {
    const int rc = conv_wgrad_tc(d_in, n_in, c_in, d_grad_out, n_out, c_out, d_table, K, d_grad_w, dtype, stream);
    if (rc != LGS_E_UNSUPPORTED) return rc;
  } else if (algo == LGS_ALGO_TC3 || algo == LGS_ALGO_BX3) {
    // the stem (c_in 4, c_out 32, K 27): exact fp32 SIMT kernel, 4x faster than 128 MMA lanes for 4 channels
    const int rs = conv_wgrad_stem(d_in, c_in, d_grad_out, n_out, c_out, d_table, K, d_grad_w, dtype, stream);
    if (rs != LGS_E_UNSUPPORTED) return rs;
    // weight gradients: single-pass TF32 products with fp32 accumulation (sums over ~1e5 rows)
    const int rc = conv_wgrad_tc(d_in, n_in, c_in, d_grad_out, n_out, c_out, d_table, K, d_grad_w, dtype, stream);
    if (rc != LGS_E_UNSUPPORTED) return rc;
  } else if (algo != LGS_ALGO_SIMT) {
    return fail(LGS_E_INVALID, "lgs_conv_wgrad: algo %d", algo);
  }
  return conv_wgrad_simt(d_in, c_in, d_grad_out, n_out, c_out, d_table, K, d_grad_w, dtype, stream);
}


int lgs_conv_fwd2(const float* d_in, int32_t c_in, const float* d_in2, int32_t c_in2, int64_t n_in, const void* d_weight,
                  int32_t K, int32_t c_out, const int32_t* d_table, int64_t n_out, int32_t reverse_k, const float* d_bias,
                  float* d_out, double* d_bn_sums, void* stream_) {
  LGS_TRACE("lgs_conv_fwd2 %p %d %p %d %lld %p %d %d %p %lld %d %p %p %p %p", (const void*)d_in, (int)c_in, (const void*)d_in2, (int)c_in2, (long long)n_in, (const void*)d_weight, (int)K, (int)c_out, (const void*)d_table, (long long)n_out, (int)reverse_k, (const void*)d_bias, (const void*)d_out, (const void*)d_bn_sums, (const void*)stream_);
  if (n_in < 0 || n_out < 0 || c_in < 1 || c_in2 < 0 || c_out < 1 || K < 1 || K > 27)
    return fail(LGS_E_INVALID, "lgs_conv_fwd2: bad sizes");
  if (!d_table && (K != 1 || n_in != n_out)) return fail(LGS_E_INVALID, "lgs_conv_fwd2: NULL table needs K == 1 and n_in == n_out");
  if ((n_out && n_in && (!d_in || (c_in2 > 0 && !d_in2))) || !d_weight || (n_out && !d_out))
    return fail(LGS_E_INVALID, "lgs_conv_fwd2: null pointer");
  const int rc = conv_fwd_bx3(d_in, c_in, c_in2 > 0 ? d_in2 : nullptr, c_in2, d_weight, K, c_out, d_table, n_out, reverse_k,
                              d_bias, d_out, d_bn_sums, static_cast<cudaStream_t>(stream_));
  if (rc == LGS_E_UNSUPPORTED) return fail(rc, "lgs_conv_fwd2: shape %d+%d->%d not supported", c_in, c_in2, c_out);
  return rc;
}


int lgs_conv_fwd3(const float* d_in, int32_t c_in, const float* d_in2, int32_t c_in2, int64_t n_in, const void* d_weight,
                  int32_t K, int32_t c_out, const int32_t* d_table, const void* d_plan, int64_t n_out, int32_t reverse_k,
                  const float* d_bias, float* d_out, double* d_bn_sums, void* stream_) {
  return lgs_conv_fwd4(d_in, c_in, d_in2, c_in2, n_in, d_weight, K, c_out, d_table, d_plan, n_out, reverse_k, d_bias, nullptr, d_out,
                       d_bn_sums, stream_);
}

int lgs_conv_fwd4(const float* d_in, int32_t c_in, const float* d_in2, int32_t c_in2, int64_t n_in, const void* d_weight,
                  int32_t K, int32_t c_out, const int32_t* d_table, const void* d_plan, int64_t n_out, int32_t reverse_k,
                  const float* d_bias, const float* d_addend, float* d_out, double* d_bn_sums, void* stream_) {
  if (d_addend && d_bn_sums) return fail(LGS_E_INVALID, "lgs_conv_fwd4: BatchNorm sums are those of the convolution alone (no addend)");
  if (!d_plan || !lgs_nbplan_supported(n_out, K) || n_in != n_out || !conv_nb_shape_ok(c_in, c_in2, c_out, K)) {
    const int rc = lgs_conv_fwd2(d_in, c_in, d_in2, c_in2, n_in, d_weight, K, c_out, d_table, n_out, reverse_k, d_bias, d_out, d_bn_sums, stream_);
    if (rc != LGS_OK || !d_addend) return rc;
    return lgs_add(d_out, d_addend, d_out, n_out * int64_t(c_out), stream_);
  }
  LGS_TRACE("lgs_conv_fwd4 %p %d %p %d %lld %p %d %d %p %p %lld %d %p %p %p %p %p", (const void*)d_in, (int)c_in, (const void*)d_in2, (int)c_in2, (long long)n_in, (const void*)d_weight, (int)K, (int)c_out, (const void*)d_table, (const void*)d_plan, (long long)n_out, (int)reverse_k, (const void*)d_bias, (const void*)d_addend, (const void*)d_out, (const void*)d_bn_sums, (const void*)stream_);
  if (!d_in || (c_in2 > 0 && !d_in2) || !d_weight || !d_out) return fail(LGS_E_INVALID, "lgs_conv_fwd4: null pointer");
  const int rc = conv_fwd_nb(d_in, c_in, c_in2 > 0 ? d_in2 : nullptr, c_in2, d_weight, K, c_out, d_plan, n_out, reverse_k, d_bias, d_addend,
                             d_out, d_bn_sums, static_cast<cudaStream_t>(stream_));
  if (rc == LGS_E_UNSUPPORTED) {
    const int r2 = lgs_conv_fwd2(d_in, c_in, d_in2, c_in2, n_in, d_weight, K, c_out, d_table, n_out, reverse_k, d_bias, d_out, d_bn_sums, stream_);
    if (r2 != LGS_OK || !d_addend) return r2;
    return lgs_add(d_out, d_addend, d_out, n_out * int64_t(c_out), stream_);
  }
  return rc;
}

}  // extern "C"
