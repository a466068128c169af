// C-ABI dispatch for the convolution entry points (include/lgs_b200.h).
#include "common.cuh"

namespace lgs {
int conv_fwd_simt(const void* in, int64_t n_in, int c_in, const void* w, int w_layout, int K, int c_out,
                  const int32_t* table, int64_t n_out, int reverse_k, const float* bias, void* out, int dtype,
                  cudaStream_t stream);
int conv_wgrad_simt(const void* in, int c_in, const void* gout, int64_t n_out, int c_out, const int32_t* table, int K,
                    float* gw, int dtype, cudaStream_t stream);
// tcgen05 path (conv_tc.cu): returns LGS_E_UNSUPPORTED when the shape is outside what it was built for
int conv_fwd_tc(const void* in, int64_t n_in, int c_in, const void* w, int K, int c_out, const int32_t* table,
                int64_t n_out, int reverse_k, const float* bias, void* out, int dtype, cudaStream_t stream);
int conv_wgrad_tc(const void* in, int64_t n_in, int c_in, const void* gout, int64_t n_out, int c_out,
                  const int32_t* table, int K, float* gw, int dtype, cudaStream_t stream);
bool tc_built();
}  // namespace lgs

using namespace lgs;

extern "C" {

int lgs_has_tc(void) { return tc_built() ? 1 : 0; }

int lgs_conv_fwd(const void* d_in, int64_t n_in, int32_t c_in, const void* d_weight, int32_t weight_layout, int32_t K,
                 int32_t c_out,
                 const int32_t* d_table, int64_t n_out, int32_t reverse_k, const float* d_bias, void* d_out,
                 int32_t dtype, int32_t algo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_in < 0 || n_out < 0 || c_in < 1 || c_out < 1 || K < 1 || K > 27)
    return fail(LGS_E_INVALID, "lgs_conv_fwd: bad sizes n_in=%lld n_out=%lld c_in=%d c_out=%d K=%d", (long long)n_in,
                (long long)n_out, c_in, c_out, K);
  if (!d_table && (K != 1 || n_in != n_out))
    return fail(LGS_E_INVALID, "lgs_conv_fwd: NULL table needs K == 1 and n_in == n_out");
  if (dtype != LGS_F32 && dtype != LGS_BF16) return fail(LGS_E_INVALID, "lgs_conv_fwd: dtype %d", dtype);
  if (weight_layout != LGS_W_KCN && weight_layout != LGS_W_KNC)
    return fail(LGS_E_INVALID, "lgs_conv_fwd: weight_layout %d", weight_layout);
  if ((n_out && (!d_in && n_in)) || !d_weight || (n_out && !d_out)) return fail(LGS_E_INVALID, "lgs_conv_fwd: null pointer");
  if (algo == LGS_ALGO_TC) {
    const int rc = weight_layout == LGS_W_KNC
                       ? conv_fwd_tc(d_in, n_in, c_in, d_weight, K, c_out, d_table, n_out, reverse_k, d_bias, d_out,
                                     dtype, stream)
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
  } else if (algo != LGS_ALGO_SIMT) {
    return fail(LGS_E_INVALID, "lgs_conv_wgrad: algo %d", algo);
  }
  return conv_wgrad_simt(d_in, c_in, d_grad_out, n_out, c_out, d_table, K, d_grad_w, dtype, stream);
}

}  // extern "C"
