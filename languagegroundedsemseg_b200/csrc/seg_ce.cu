// Fused softmax cross-entropy over the per-point class logits (mean over the points whose label is not `ignore`),
// forward AND gradient in one pass over the [n, C] logit matrix.
//   Replaces nn.CrossEntropyLoss(ignore_index=...) at lib/train_test/pl_BaselineTrainer.py:343,350 (criterion built by
//   loss_by_name, :96): ATen runs log_softmax, nll_loss, their two backward kernels and a reduction — five passes over
//   a 119 MB matrix at 150 K points x 200 classes (~0.5 ms); here the matrix is read once and its gradient written once.
//   launch 1 (one block)   n_valid = #{i : y_i != ignore}; clears the loss accumulator and the completion ticket
//   launch 2 (warp per row) loss_i = logsumexp(x_i) - x_i[y_i];  dlogits_i = (softmax(x_i) - onehot(y_i)) / n_valid
//                           block partial sums -> one fp64 atomic; the last block to finish writes loss = sum / n_valid
// Labels outside [0, C) other than `ignore` are an error in torch (device assert); here they are treated as ignored.
#include <algorithm>

#include "common.cuh"

namespace lgs {

constexpr int SCE_THREADS = 256;          // 8 rows per block
constexpr int SCE_MAXV = 8;               // float4 per lane: C <= 1024

__global__ void __launch_bounds__(1024)
seg_ce_count_kernel(const int64_t* __restrict__ labels, int64_t n, int c, int64_t ignore, double* __restrict__ ws) {
  pdl_grid_sync();
  __shared__ int part[32];
  int cnt = 0;
  // 8 independent loads in flight per thread: one block walks 1.2 MB of labels, serial loads cost 32 us at 150 K points
  for (int64_t i0 = threadIdx.x; i0 < n; i0 += int64_t(blockDim.x) * 8) {
    int64_t y[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int64_t i = i0 + int64_t(u) * blockDim.x;
      y[u] = i < n ? __ldg(labels + i) : ignore;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) cnt += (y[u] != ignore && y[u] >= 0 && y[u] < c) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) tot += part[w];
    ws[0] = 0.0;                                   // loss accumulator
    ws[1] = double(tot);                           // n_valid
    reinterpret_cast<unsigned long long*>(ws)[2] = 0ull;   // completion ticket
  }
}

__global__ void __launch_bounds__(SCE_THREADS)
seg_ce_kernel(const float* __restrict__ logits, int64_t n, int c, const int64_t* __restrict__ labels, int64_t ignore,
              double* __restrict__ ws, float* __restrict__ loss_out, float* __restrict__ dlogits) {
  pdl_grid_sync();
  __shared__ float wsum[SCE_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c4 = c >> 2;
  const float inv_valid = float(1.0 / ws[1]);      // written by the count kernel (stream order)
  float my_loss = 0.f;
  // persistent grid: a warp walks rows with the grid's stride (one loss atomic per block instead of one per 8 rows)
  for (int64_t row = int64_t(blockIdx.x) * (SCE_THREADS / 32) + warp; row < n; row += int64_t(gridDim.x) * (SCE_THREADS / 32)) {
    const int64_t y = __ldg(labels + row);
    const bool valid = y != ignore && y >= 0 && y < c;
    float4* drow = dlogits ? reinterpret_cast<float4*>(dlogits + row * c) : nullptr;
    if (!valid) {
      if (drow)
        for (int j = lane; j < c4; j += 32) drow[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      const float4* xrow = reinterpret_cast<const float4*>(logits + row * c);
      float4 v[SCE_MAXV];
      float m = -INFINITY;
#pragma unroll
      for (int u = 0; u < SCE_MAXV; ++u) {
        const int j = lane + 32 * u;
        if (j < c4) {
          v[u] = __ldg(xrow + j);
          m = fmaxf(fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)), m);
        }
      }
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float s = 0.f, xy = 0.f;
      const int yq = int(y) >> 2, ye = int(y) & 3;
#pragma unroll
      for (int u = 0; u < SCE_MAXV; ++u) {
        const int j = lane + 32 * u;
        if (j < c4) {
          if (j == yq) xy = ye == 0 ? v[u].x : ye == 1 ? v[u].y : ye == 2 ? v[u].z : v[u].w;
          v[u].x = __expf(v[u].x - m); v[u].y = __expf(v[u].y - m);
          v[u].z = __expf(v[u].z - m); v[u].w = __expf(v[u].w - m);
          s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        xy += __shfl_xor_sync(0xffffffffu, xy, o);    // only one lane holds a non-zero value
      }
      my_loss += __logf(s) + m - xy;
      if (drow) {
        const float k = inv_valid / s;
#pragma unroll
        for (int u = 0; u < SCE_MAXV; ++u) {
          const int j = lane + 32 * u;
          if (j < c4) {
            float4 g = make_float4(v[u].x * k, v[u].y * k, v[u].z * k, v[u].w * k);
            if (j == yq) {
              if (ye == 0) g.x -= inv_valid; else if (ye == 1) g.y -= inv_valid;
              else if (ye == 2) g.z -= inv_valid; else g.w -= inv_valid;
            }
            drow[j] = g;
          }
        }
      }
    }
  }
  if (lane == 0) wsum[warp] = my_loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < SCE_THREADS / 32; ++w) tot += double(wsum[w]);
    atomicAdd(ws, tot);
    __threadfence();
    const unsigned long long ticket = atomicAdd(reinterpret_cast<unsigned long long*>(ws) + 2, 1ull);
    if (ticket == gridDim.x - 1) {                 // every block's partial sum is in
      __threadfence();
      const double sum = *reinterpret_cast<volatile double*>(ws);
      *loss_out = float(sum / ws[1]);
    }
  }
}

}  // namespace lgs

using namespace lgs;

extern "C" {

int lgs_seg_ce_supported(int32_t c) { return (c >= 4 && (c & 3) == 0 && c <= 32 * 4 * SCE_MAXV) ? 1 : 0; }

int lgs_seg_ce(const float* d_logits, int64_t n, int32_t c, const int64_t* d_labels, int64_t ignore_label,
               double* d_ws /*[4]*/, float* d_loss, float* d_grad_logits, void* stream_) {
  LGS_TRACE("lgs_seg_ce %p %lld %d %p %lld %p %p %p %p", (const void*)d_logits, (long long)n, (int)c, (const void*)d_labels, (long long)ignore_label, (const void*)d_ws, (const void*)d_loss, (const void*)d_grad_logits, (const void*)stream_);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n < 1 || c < 1) return fail(LGS_E_INVALID, "lgs_seg_ce: bad sizes n=%lld c=%d", (long long)n, c);
  if (!lgs_seg_ce_supported(c)) return fail(LGS_E_UNSUPPORTED, "lgs_seg_ce: c=%d (need c %% 4 == 0, 4 <= c <= %d)", c, 128 * SCE_MAXV);
  if (!d_logits || !d_labels || !d_ws || !d_loss) return fail(LGS_E_INVALID, "lgs_seg_ce: null pointer");
  if ((reinterpret_cast<uintptr_t>(d_logits) & 15) || (reinterpret_cast<uintptr_t>(d_grad_logits) & 15))
    return fail(LGS_E_UNSUPPORTED, "lgs_seg_ce: logits / gradient rows must be 16-byte aligned");
  LGS_LAUNCH_PDL(seg_ce_count_kernel, 1, 1024, 0, stream, d_labels, n, c, ignore_label, d_ws);
  const unsigned sce_grid = unsigned(std::min<int64_t>(cdiv(n, SCE_THREADS / 32), 148 * 8));
  LGS_LAUNCH_PDL(seg_ce_kernel, sce_grid, SCE_THREADS, 0, stream, d_logits, n, c, d_labels,
             ignore_label, d_ws, d_loss, d_grad_logits);
  return LGS_OK;
}

}  // extern "C"
