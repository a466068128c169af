// Native step driver: a straight-line program of engine calls, built once per network and executed
// per training step by ONE C-ABI call — no Python, no autograd graph, no per-layer tensor objects on the critical path.
//
// The reference's trainer runs forward, loss, backward layer by layer through Python (lib/train_test/pl_BaselineTrainer.py:
// 157-160, 288-309 under PyTorch Lightning); the facade reproduces that faithfully and costs ~14.8 ms of host time per
// Res16UNet34C step, as much as the GPU work itself.  Here the host side of a step is a loop over ~450 fixed-size op
// records that calls the SAME entry points the facade calls (lgs_conv_fwd2, lgs_conv_wgrad, lgs_bn_fwd2, lgs_bn_bwd,
// lgs_seg_ce, ...), so the library's call recorder (lgs_trace_begin) shows the program's calls next to the facade's and
// the CPU tests compare them.
//
// Memory: every intermediate (activations kept for backward, gradients) lives in ONE caller-provided arena; a buffer is
// (level, channels): its row count is the number of voxels of that U-Net level in THIS batch, known only at run time, so
// offsets are assigned per run by a bump allocator (no reuse inside a step: 180 GB of HBM3e hold ~40 scenes' worth).
// Parameters, gradients, BatchNorm buffers, weight operands and kernel-map tables are EXTERNAL pointers (owned by PyTorch)
// passed in a table per run, so the nn.Module keeps owning its state (optimiser, checkpoints, DDP buckets unchanged).
#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#include <cstdlib>

#include "common.cuh"

namespace lgs {

// ---- small kernels the program needs besides the engine's own ------------------------------------------------------
// dst[r, dst_col0 + c] = src[r, src_col0 + c]  for c < cols  (float4 when everything is 16-byte aligned):
// `cat` (models/res16unet.py:237,247,257,267), its backward column split, c_in 3 -> 4 padding, gradient slicing
__global__ void __launch_bounds__(256) copy2d_kernel(const float* __restrict__ src, int64_t src_ld, const float* dummy,
                                                     float* __restrict__ dst, int64_t dst_ld, int64_t rows, int cols) {
  pdl_grid_sync();
  (void)dummy;
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols;
    const int c = int(i - r * cols);
    dst[r * dst_ld + c] = __ldg(src + r * src_ld + c);
  }
}
__global__ void __launch_bounds__(256) copy2d_v4_kernel(const float4* __restrict__ src, int64_t src_ld4, float4* __restrict__ dst,
                                                        int64_t dst_ld4, int64_t rows, int cols4) {
  pdl_grid_sync();
  const int64_t total = rows * cols4;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols4;
    const int c = int(i - r * cols4);
    dst[r * dst_ld4 + c] = __ldg(src + r * src_ld4 + c);
  }
}
// out = a + b (float4 stream): gradient accumulation where two consumers meet (residual branches, skip connections)
__global__ void __launch_bounds__(256) zero_v4_kernel(float4* __restrict__ p, int64_t n4) {
  pdl_grid_sync();
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x)
    p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
__global__ void __launch_bounds__(256) add_v4_kernel(const float4* a, const float4* __restrict__ b, float4* out,   // a may alias out

                                                     int64_t n4) {
  pdl_grid_sync();
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    const float4 x = a[i], y = __ldg(b + i);
    out[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
  }
}
// out[c] = sum_r g[r, c]  (bias gradient of the classifier, models/res16unet.py:193): block = 32 x 8, fp32 partials, atomics
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ g, int64_t rows, int c, float* __restrict__ out) {
  pdl_grid_sync();
  __shared__ float sh[8][33];
  const int ch = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (ch < c)
    for (int64_t r = int64_t(blockIdx.y) * 8 + threadIdx.y; r < rows; r += int64_t(gridDim.y) * 8) acc += __ldg(g + r * c + ch);
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += sh[j][threadIdx.x];
    atomicAdd(out + ch, t);
  }
}

// vectorised variant (c % 4 == 0, 16-byte aligned rows): the matrix as one stream of float4, a block owns 128 consecutive rows,
// its T threads (the largest multiple of c / 4 <= 256) stride over them so that a thread always owns the same four columns
__global__ void __launch_bounds__(256) colsum_v4_kernel(const float4* __restrict__ g, int64_t rows, int c4, int T, float* __restrict__ out) {
  pdl_grid_sync();
  __shared__ float4 sh[256];
  const int t = threadIdx.x;
  const int64_t i0 = int64_t(blockIdx.x) * 128 * c4, i1 = min(rows, int64_t(blockIdx.x + 1) * 128) * c4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t < T) {
    for (int64_t i = i0 + t; i < i1; i += int64_t(4) * T) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (i + int64_t(u) * T < i1) ? __ldg(g + i + int64_t(u) * T) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc.x += v[u].x, acc.y += v[u].y, acc.z += v[u].z, acc.w += v[u].w;
    }
  }
  sh[t] = acc;
  __syncthreads();
  if (t < c4) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int u = t; u < T; u += c4) s.x += sh[u].x, s.y += sh[u].y, s.z += sh[u].z, s.w += sh[u].w;
    atomicAdd(out + 4 * t, s.x);
    atomicAdd(out + 4 * t + 1, s.y);
    atomicAdd(out + 4 * t + 2, s.z);
    atomicAdd(out + 4 * t + 3, s.w);
  }
}


}  // namespace lgs

using namespace lgs;

namespace lgs {
int zero_fill_async(void* p, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return LGS_OK;
  static const bool use_memset = getenv("LGS_ZERO_MEMSET") != nullptr;      // A/B knob
  if (use_memset || (bytes & 15) || (reinterpret_cast<uintptr_t>(p) & 15) || !g_pdl.load(std::memory_order_relaxed)) {
    LGS_CUDA(cudaMemsetAsync(p, 0, bytes, stream));
    return LGS_OK;
  }
  const int64_t n4 = int64_t(bytes >> 4);
  const unsigned blocks = unsigned(std::min<int64_t>(cdiv(n4, 256 * 2), 148 * 4));
  LGS_LAUNCH_PDL(zero_v4_kernel, std::max(1u, blocks), 256, 0, stream, static_cast<float4*>(p), n4);
  return LGS_OK;
}
}  // namespace lgs

extern "C" {

int lgs_copy2d(const float* d_src, int64_t src_ld, float* d_dst, int64_t dst_ld, int64_t rows, int32_t cols, void* stream_) {
  LGS_TRACE("lgs_copy2d %p %lld %p %lld %lld %d %p", (const void*)d_src, (long long)src_ld, (const void*)d_dst, (long long)dst_ld, (long long)rows, (int)cols, (const void*)stream_);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows < 0 || cols < 0 || src_ld < cols || dst_ld < cols) return fail(LGS_E_INVALID, "lgs_copy2d: bad sizes");
  if (rows == 0 || cols == 0) return LGS_OK;
  if (!d_src || !d_dst) return fail(LGS_E_INVALID, "lgs_copy2d: null pointer");
  const bool v4 = !((reinterpret_cast<uintptr_t>(d_src) | reinterpret_cast<uintptr_t>(d_dst)) & 15) && !(src_ld & 3) && !(dst_ld & 3) && !(cols & 3);
  const int64_t work = v4 ? rows * (cols / 4) : rows * cols;
  const unsigned blocks = unsigned(std::min<int64_t>(cdiv(work, 256 * 4), 148 * 8));
  if (v4) {
    LGS_LAUNCH_PDL(copy2d_v4_kernel, std::max(1u, blocks), 256, 0, stream, reinterpret_cast<const float4*>(d_src), src_ld / 4,
               reinterpret_cast<float4*>(d_dst), dst_ld / 4, rows, cols / 4);
  } else {
    LGS_LAUNCH_PDL(copy2d_kernel, std::max(1u, blocks), 256, 0, stream, d_src, src_ld, nullptr, d_dst, dst_ld, rows, cols);
  }
  return LGS_OK;
}

int lgs_add(const float* d_a, const float* d_b, float* d_out, int64_t n, void* stream_) {
  LGS_TRACE("lgs_add %p %p %p %lld %p", (const void*)d_a, (const void*)d_b, (const void*)d_out, (long long)n, (const void*)stream_);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n < 0 || (n & 3)) return fail(LGS_E_INVALID, "lgs_add: n must be a non-negative multiple of 4");
  if (n == 0) return LGS_OK;
  if (!d_a || !d_b || !d_out || ((reinterpret_cast<uintptr_t>(d_a) | reinterpret_cast<uintptr_t>(d_b) | reinterpret_cast<uintptr_t>(d_out)) & 15))
    return fail(LGS_E_INVALID, "lgs_add: null or unaligned pointer");
  const unsigned blocks = unsigned(std::min<int64_t>(cdiv(n / 4, 256 * 4), 148 * 8));
  LGS_LAUNCH_PDL(add_v4_kernel, std::max(1u, blocks), 256, 0, stream, reinterpret_cast<const float4*>(d_a), reinterpret_cast<const float4*>(d_b),
             reinterpret_cast<float4*>(d_out), n / 4);
  return LGS_OK;
}

int lgs_colsum(const float* d_g, int64_t rows, int32_t c, float* d_out, void* stream_) {
  LGS_TRACE("lgs_colsum %p %lld %d %p %p", (const void*)d_g, (long long)rows, (int)c, (const void*)d_out, (const void*)stream_);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows < 0 || c < 1 || !d_out) return fail(LGS_E_INVALID, "lgs_colsum: bad arguments");
  {
    const int rz = zero_fill_async(d_out, size_t(c) * sizeof(float), stream);
    if (rz != LGS_OK) return rz;
  }
  if (rows == 0) return LGS_OK;
  if ((c & 3) == 0 && c <= 1024 && !(reinterpret_cast<uintptr_t>(d_g) & 15)) {
    const int c4 = c >> 2, T = (256 / c4) * c4;
    LGS_LAUNCH_PDL(colsum_v4_kernel, unsigned(cdiv(rows, 128)), 256, 0, stream, reinterpret_cast<const float4*>(d_g), rows, c4, T, d_out);
    return LGS_OK;
  }
  const dim3 grid{unsigned((c + 31) / 32), unsigned(std::min<int64_t>(cdiv(rows, 8 * 16), 148 * 4)), 1u}, block{32u, 8u, 1u};
  LGS_LAUNCH_PDL(colsum_kernel, grid, block, 0, stream, d_g, rows, c, d_out);
  return LGS_OK;
}

// ---- the program -------------------------------------------------------------------------------------------------------
struct lgs_program {
  std::vector<int64_t> ops;     // [n_ops][LGS_PROGRAM_OP_WORDS]
  std::vector<int64_t> bufs;    // [n_bufs][4] = kind (0 external slot, 1 arena), level (-1: rows = 1), channels, element bytes
  int32_t n_ops = 0, n_bufs = 0, n_levels = 0, n_ext = 0;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int bn_half = 0;              // which half of the BatchNorm accumulator scratch is all-zero / current
  std::vector<int64_t> offsets; // per run
};

static inline int64_t align256(int64_t x) { return (x + 255) & ~int64_t(255); }
// BatchNorm statistics come out of the convolution's epilogue only on maps that fill the machine without splitting the
// kernel offsets over CTAs (partial sums cannot feed the statistics); smaller maps keep the separate statistics pass
constexpr int64_t kFuseStatsMinRows = 148 * 128;

int lgs_program_create(const int64_t* ops, int32_t n_ops, const int64_t* bufs, int32_t n_bufs, int32_t n_levels, int32_t n_ext,
                       lgs_program** out) {
  if (!ops || !bufs || !out || n_ops < 0 || n_bufs < 0 || n_levels < 1 || n_ext < 0) return fail(LGS_E_INVALID, "lgs_program_create: bad arguments");
  for (int32_t b = 0; b < n_bufs; ++b) {
    const int64_t* r = bufs + int64_t(b) * 4;
    if ((r[0] == 0 && (r[1] < 0 || r[1] >= n_ext)) || (r[0] == 1 && (r[1] < -1 || r[1] >= n_levels || r[2] < 1 || r[3] < 1)) || r[0] < 0 || r[0] > 1)
      return fail(LGS_E_INVALID, "lgs_program_create: bad buffer record %d", b);
  }
  lgs_program* p = new lgs_program;
  p->ops.assign(ops, ops + int64_t(n_ops) * LGS_PROGRAM_OP_WORDS);
  p->bufs.assign(bufs, bufs + int64_t(n_bufs) * 4);
  p->n_ops = n_ops, p->n_bufs = n_bufs, p->n_levels = n_levels, p->n_ext = n_ext;
  p->offsets.resize(size_t(n_bufs));
  *out = p;
  return LGS_OK;
}

void lgs_program_destroy(lgs_program* p) {
  if (!p) return;
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->ev_join) cudaEventDestroy(p->ev_join);
  delete p;
}

int64_t lgs_program_arena_bytes(const lgs_program* p, const int64_t* level_rows) {
  if (!p || !level_rows) return -1;
  int64_t off = 0;
  for (int32_t b = 0; b < p->n_bufs; ++b) {
    const int64_t* r = p->bufs.data() + int64_t(b) * 4;
    if (r[0] != 1) continue;
    const int64_t rows = r[1] < 0 ? 1 : level_rows[r[1]];
    off += align256(rows * r[2] * r[3]);
  }
  return off;
}

void lgs_program_reset(lgs_program* p) {
  if (p) p->bn_half = 0;
}

int lgs_program_run(lgs_program* p, int32_t op_begin, int32_t op_end, const int64_t* level_rows, void* const* ext,
                    void* d_arena, int64_t arena_bytes, void* d_bn_scratch, void* stream_, void* side_stream_) {
  return lgs_program_run2(p, op_begin, op_end, level_rows, ext, d_arena, arena_bytes, d_bn_scratch, stream_, side_stream_, 0);
}

int lgs_program_run2(lgs_program* p, int32_t op_begin, int32_t op_end, const int64_t* level_rows, void* const* ext,
                     void* d_arena, int64_t arena_bytes, void* d_bn_scratch, void* stream_, void* side_stream_, int32_t flags) {
  if (!p || !level_rows || !ext || op_begin < 0 || op_end > p->n_ops || op_begin > op_end)
    return fail(LGS_E_INVALID, "lgs_program_run: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_), side = static_cast<cudaStream_t>(side_stream_);
  const bool tracing = g_trace_on.load(std::memory_order_relaxed) != 0;
  // buffer addresses of this run
  std::vector<uint8_t*> addr(size_t(p->n_bufs));
  {
    int64_t off = 0;
    for (int32_t b = 0; b < p->n_bufs; ++b) {
      const int64_t* r = p->bufs.data() + int64_t(b) * 4;
      if (r[0] == 0) {
        addr[b] = static_cast<uint8_t*>(ext[r[1]]);
      } else {
        const int64_t rows = r[1] < 0 ? 1 : level_rows[r[1]];
        addr[b] = static_cast<uint8_t*>(d_arena) + off;
        off += align256(rows * r[2] * r[3]);
      }
    }
    if (off > arena_bytes) return fail(LGS_E_INVALID, "lgs_program_run: arena of %lld bytes, %lld needed", (long long)arena_bytes, (long long)off);
  }
  auto P = [&](int64_t id) -> void* { return id < 0 ? nullptr : addr[size_t(id)]; };
  auto rows_of = [&](int64_t lvl) -> int64_t { return lvl < 0 ? 1 : level_rows[lvl]; };
  double* scratch = static_cast<double*>(d_bn_scratch);
  auto half = [&](int h) { return scratch + size_t(h) * 16384; };
  bool forked = false;
  if (!tracing && side && !p->ev_fork) {
    LGS_CUDA(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    LGS_CUDA(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
  }
  int rc = LGS_OK;
  for (int32_t i = op_begin; i < op_end && rc == LGS_OK; ++i) {
    const int64_t* o = p->ops.data() + int64_t(i) * LGS_PROGRAM_OP_WORDS;
    switch (o[0]) {
      case LGS_OP_WEIGHT_PREP:   // 1 desc, 2 n_layers, 3 total_tiles, 4 nsplit
        rc = lgs_weight_prep_batch(static_cast<const int64_t*>(P(o[1])), int32_t(o[2]), o[3], int32_t(o[4]), LGS_F32, stream);
        break;
      case LGS_OP_CONV: {        // 1 in, 2 c_in, 3 in2, 4 c_in2, 5 lvl_in, 6 weight, 7 K, 8 c_out, 9 table, 10 lvl_out, 11 reverse_k, 12 bias, 13 out, 14 stats flag, 15 plan, 16 addend
        double* sums = (o[14] && rows_of(o[10]) >= kFuseStatsMinRows) ? half(p->bn_half) : nullptr;
        rc = lgs_conv_fwd4(static_cast<const float*>(P(o[1])), int32_t(o[2]), static_cast<const float*>(P(o[3])), int32_t(o[4]), rows_of(o[5]),
                           P(o[6]), int32_t(o[7]), int32_t(o[8]), static_cast<const int32_t*>(P(o[9])), P(o[15]), rows_of(o[10]), int32_t(o[11]),
                           static_cast<const float*>(P(o[12])), static_cast<const float*>(P(o[16])), static_cast<float*>(P(o[13])), sums, stream);
        break;
      }
      case LGS_OP_WGRAD: {       // 1 in, 2 c_in, 3 lvl_in, 4 gout, 5 c_out, 6 lvl_out, 7 table, 8 K, 9 gw, 10 algo, 11 on side stream
        cudaStream_t s = stream;
        if (o[11] && side && !tracing) {
          LGS_CUDA(cudaEventRecord(p->ev_fork, stream));
          LGS_CUDA(cudaStreamWaitEvent(side, p->ev_fork, 0));
          s = side;
          forked = true;
        }
        rc = lgs_conv_wgrad(P(o[1]), rows_of(o[3]), int32_t(o[2]), P(o[4]), rows_of(o[6]), int32_t(o[5]), static_cast<const int32_t*>(P(o[7])),
                            int32_t(o[8]), static_cast<float*>(P(o[9])), LGS_F32, int32_t(o[10]), s);
        break;
      }
      case LGS_OP_BN_FWD: {      // 1 x, 2 res, 3 lvl, 4 c, 5 gamma, 6 beta, 7 eps bits, 8 momentum bits, 9 relu, 10 rm, 11 rv, 12 z, 13 stats [2,c], 14 nbt, 15 stats_ready
        float eps, mom;
        const uint32_t eb = uint32_t(o[7]), mb = uint32_t(o[8]);
        memcpy(&eps, &eb, 4), memcpy(&mom, &mb, 4);
        float* st = static_cast<float*>(P(o[13]));
        rc = lgs_bn_fwd2(static_cast<const float*>(P(o[1])), static_cast<const float*>(P(o[2])), rows_of(o[3]), int32_t(o[4]),
                         static_cast<const float*>(P(o[5])), static_cast<const float*>(P(o[6])), eps, mom, int32_t(o[9]),
                         static_cast<float*>(P(o[10])), static_cast<float*>(P(o[11])), static_cast<float*>(P(o[12])), st, st + o[4],
                         half(p->bn_half), half(1 - p->bn_half), static_cast<int64_t*>(P(o[14])),
                         (o[15] && rows_of(o[3]) >= kFuseStatsMinRows) ? 1 : 0, stream);
        if (rc == LGS_OK) p->bn_half ^= 1;
        break;
      }
      case LGS_OP_BN_BWD: {      // 1 x, 2 z, 3 dz, 4 lvl, 5 c, 6 gamma, 7 stats [2,c], 8 relu, 9 dx, 10 dres, 11 dgamma, 12 dbeta
        const float* st = static_cast<const float*>(P(o[7]));
        rc = lgs_bn_bwd(static_cast<const float*>(P(o[1])), static_cast<const float*>(P(o[2])), static_cast<const float*>(P(o[3])), rows_of(o[4]),
                        int32_t(o[5]), static_cast<const float*>(P(o[6])), st, st + o[5], int32_t(o[8]), static_cast<float*>(P(o[9])),
                        static_cast<float*>(P(o[10])), static_cast<float*>(P(o[11])), static_cast<float*>(P(o[12])), half(p->bn_half),
                        half(1 - p->bn_half), stream);
        if (rc == LGS_OK) p->bn_half ^= 1;
        break;
      }
      case LGS_OP_COPY2D:        // 1 src, 2 src_ld, 3 src_col0, 4 dst, 5 dst_ld, 6 dst_col0, 7 lvl (rows) or -1 with 8 = rows, 9 cols
        rc = lgs_copy2d(static_cast<const float*>(P(o[1])) + o[3], o[2], static_cast<float*>(P(o[4])) + o[6], o[5],
                        o[7] >= 0 ? rows_of(o[7]) : o[8], int32_t(o[9]), stream);
        break;
      case LGS_OP_ADD:           // 1 a, 2 b, 3 out, 4 lvl, 5 c
        rc = lgs_add(static_cast<const float*>(P(o[1])), static_cast<const float*>(P(o[2])), static_cast<float*>(P(o[3])), rows_of(o[4]) * o[5], stream);
        break;
      case LGS_OP_SEG_CE:        // 1 logits, 2 lvl, 3 c, 4 labels, 5 ignore, 6 ws, 7 loss, 8 grad
        rc = lgs_seg_ce(static_cast<const float*>(P(o[1])), rows_of(o[2]), int32_t(o[3]), static_cast<const int64_t*>(P(o[4])), o[5],
                        static_cast<double*>(P(o[6])), static_cast<float*>(P(o[7])), static_cast<float*>(P(o[8])), stream);
        break;
      case LGS_OP_COLSUM:        // 1 g, 2 lvl, 3 c, 4 out
        rc = lgs_colsum(static_cast<const float*>(P(o[1])), rows_of(o[2]), int32_t(o[3]), static_cast<float*>(P(o[4])), stream);
        break;
      case LGS_OP_JOIN:          // the training stream waits for everything issued on the side stream so far
        if (forked && !tracing) {
          LGS_CUDA(cudaEventRecord(p->ev_join, side));
          LGS_CUDA(cudaStreamWaitEvent(stream, p->ev_join, 0));
          forked = false;
        }
        break;
      default:
        rc = fail(LGS_E_INVALID, "lgs_program_run: unknown op %lld at %d", (long long)o[0], i);
    }
  }
  if (forked && !tracing && !(flags & LGS_RUN_NO_JOIN)) {     // never leave side-stream work unjoined behind a returning call
    cudaEventRecord(p->ev_join, side);
    cudaStreamWaitEvent(stream, p->ev_join, 0);
  }
  return rc;
}

}  // extern "C"
