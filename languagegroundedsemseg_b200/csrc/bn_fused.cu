// Fused training-mode BatchNorm (+ residual add) (+ ReLU) over the rows of a feature matrix [n, c] (SURVEY.md §8 f,
// rank 1).  Replaces, per layer, ATen's collect_statistics + transform_input + relu (+ add) in forward and
// threshold_backward + backward_reduce + backward_elemt (+ add) in backward by 2 + 2 launches that each read the
// feature matrix once:
//   fwd  (1) column sums / sums of squares (fp32 per thread -> fp64 atomics per block)
//        (2) z = relu?( gamma * (x - mean) * invstd + beta [+ residual] ),  running stats updated by block 0
//   bwd  (3) dy = dz * (z > 0)?;  column sums of dy and dy * xhat
//        (4) dx = gamma * invstd * (dy - mean(dy) - xhat * mean(dy * xhat));  d_residual = dy
// Reference semantics: ME.MinkowskiBatchNorm = nn.BatchNorm1d on .F (models/modules/common.py:17-19), eps 1e-5,
// biased variance for normalisation, unbiased for the running estimate.
#include <cstdlib>

#include "common.cuh"

namespace lgs {

constexpr int BN_TX = 32;     // channels per block (x): one warp reads 128 contiguous bytes of a row
constexpr int BN_TY = 8;      // row lanes per block (y)
constexpr int BN_ROWS = 256;  // rows per block
constexpr int BN_COPIES = 8;  // interleaved copies of the fp64 accumulators (block b adds into copy b % 8): less contention
constexpr int BN_SCRATCH_DOUBLES = 2 * BN_COPIES * 1024;   // one accumulator scratch: [BN_COPIES][2 * c], c <= 1024

__device__ __forceinline__ void bn_block_reduce(float a, float b, int ch, int c, double* sums) {
  __shared__ float s1[BN_TY][BN_TX + 1], s2[BN_TY][BN_TX + 1];
  s1[threadIdx.y][threadIdx.x] = a;
  s2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
    double da = 0.0, db = 0.0;
#pragma unroll
    for (int j = 0; j < BN_TY; ++j) {
      da += s1[j][threadIdx.x];
      db += s2[j][threadIdx.x];
    }
    double* dst = sums + size_t(blockIdx.y % BN_COPIES) * 2 * c;
    atomicAdd(dst + ch, da);
    atomicAdd(dst + c + ch, db);
  }
}

// grid (ceil(c/32), ceil(n/BN_ROWS)); block (32, 8); U independent row loads in flight per thread
template <int U>
__global__ void __launch_bounds__(BN_TX * BN_TY)
bn_stats_kernel(const float* __restrict__ x, int64_t n, int c, double* __restrict__ sums /*[BN_COPIES][2c]*/) {
  pdl_grid_sync();
  const int ch = blockIdx.x * BN_TX + threadIdx.x;
  const int64_t r0 = int64_t(blockIdx.y) * BN_ROWS + threadIdx.y;
  const int64_t r1 = min(n, int64_t(blockIdx.y + 1) * BN_ROWS);
  float a = 0.f, b = 0.f;
  if (ch < c) {
    const float* p = x + ch;
    for (int64_t r = r0; r < r1; r += U * BN_TY) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = (r + u * BN_TY < r1) ? __ldg(p + (r + u * BN_TY) * c) : 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        a += v[u];
        b = fmaf(v[u], v[u], b);
      }
    }
  }
  bn_block_reduce(a, b, ch, c, sums);
}

// ---- vectorised statistics (default) ---------------------------------------------------------------------------
// The feature matrix is one flat stream of float4: a block owns BN_VROWS consecutive rows and its T threads
// (T = the largest multiple of c/4 <= 256) stride over them, so a thread always sees the same 4 channels, every load is
// 16 bytes and a warp reads 512 contiguous bytes.  U loads are in flight per thread.
constexpr int BN_VROWS = 128;   // small blocks: 8 resident per SM keep ~128 KB of loads in flight
constexpr int BN_VTHREADS = 256;

// partial sums of the threads that own the same channel quad -> fp64 atomics (one per channel and block)
__device__ __forceinline__ void bn_vec_reduce(const float (&a)[4], const float (&b)[4], int c4, int T, int c, double* sums) {
  __shared__ float sh[BN_VTHREADS][9];
  const int t = threadIdx.x;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    sh[t][e] = a[e];
    sh[t][4 + e] = b[e];
  }
  __syncthreads();
  // c4 * 8 (quad, component) pairs, each summed over the T / c4 threads of that quad
  for (int j = t; j < c4 * 8; j += BN_VTHREADS) {
    const int q = j >> 3, e = j & 7;
    double acc = 0.0;
    for (int u = q; u < T; u += c4) acc += sh[u][e];
    double* dst = sums + size_t(blockIdx.x % BN_COPIES) * 2 * c;
    atomicAdd(dst + (e < 4 ? 0 : c) + q * 4 + (e & 3), acc);
  }
}

template <int U>
__global__ void __launch_bounds__(BN_VTHREADS)
bn_stats_vec_kernel(const float4* __restrict__ x, int64_t n, int c, int T, double* __restrict__ sums /*[BN_COPIES][2c]*/) {
  pdl_grid_sync();
  const int c4 = c >> 2;
  const int64_t i0 = int64_t(blockIdx.x) * BN_VROWS * c4;
  const int64_t i1 = min(n, int64_t(blockIdx.x + 1) * BN_VROWS) * c4;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  if (int(threadIdx.x) < T) {
    for (int64_t i = i0 + threadIdx.x; i < i1; i += int64_t(U) * T) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = (i + int64_t(u) * T < i1) ? __ldg(x + i + int64_t(u) * T) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        a[0] += v[u].x; a[1] += v[u].y; a[2] += v[u].z; a[3] += v[u].w;
        b[0] = fmaf(v[u].x, v[u].x, b[0]); b[1] = fmaf(v[u].y, v[u].y, b[1]);
        b[2] = fmaf(v[u].z, v[u].z, b[2]); b[3] = fmaf(v[u].w, v[u].w, b[3]);
      }
    }
  }
  bn_vec_reduce(a, b, c4, T, c, sums);
}

template <int U>
__global__ void __launch_bounds__(BN_VTHREADS)
bn_bwd_stats_vec_kernel(const float4* __restrict__ x, const float4* __restrict__ z, const float4* __restrict__ dz, int64_t n,
                        int c, int T, const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
                        double* __restrict__ sums /*[BN_COPIES][2c]: sum dy, sum dy*xhat*/) {
  pdl_grid_sync();
  const int c4 = c >> 2;
  const int64_t i0 = int64_t(blockIdx.x) * BN_VROWS * c4;
  const int64_t i1 = min(n, int64_t(blockIdx.x + 1) * BN_VROWS) * c4;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  if (int(threadIdx.x) < T) {
    const int q = int(threadIdx.x) % c4;      // T is a multiple of c4 and i0 a multiple of c4: the quad never changes
    const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + q);
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd) + q);
    for (int64_t i = i0 + threadIdx.x; i < i1; i += int64_t(U) * T) {
      float4 g[U], zz[U], v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = i + int64_t(u) * T < i1;
        const int64_t j = i + int64_t(u) * T;
        g[u] = ok ? __ldg(dz + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        zz[u] = (ok && relu) ? __ldg(z + j) : make_float4(1.f, 1.f, 1.f, 1.f);
        v[u] = ok ? __ldg(x + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float d0 = zz[u].x > 0.f ? g[u].x : 0.f, d1 = zz[u].y > 0.f ? g[u].y : 0.f;
        const float d2 = zz[u].z > 0.f ? g[u].z : 0.f, d3 = zz[u].w > 0.f ? g[u].w : 0.f;
        a[0] += d0; a[1] += d1; a[2] += d2; a[3] += d3;
        b[0] = fmaf(d0, (v[u].x - m.x) * is.x, b[0]);
        b[1] = fmaf(d1, (v[u].y - m.y) * is.y, b[1]);
        b[2] = fmaf(d2, (v[u].z - m.z) * is.z, b[2]);
        b[3] = fmaf(d3, (v[u].w - m.w) * is.w, b[3]);
      }
    }
  }
  bn_vec_reduce(a, b, c4, T, c, sums);
}

// mean / invstd from the sums; block (0,0) also updates the running statistics
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ res, int64_t n, int c,
                const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta,
                float eps, int relu, float* __restrict__ z, float* __restrict__ save_mean, float* __restrict__ save_invstd,
                float* running_mean, float* running_var, float momentum, long long* num_batches_tracked,
                double* __restrict__ zero_next) {
  pdl_grid_sync();
  extern __shared__ float sh[];   // scale[c], shift[c]
  float* scale = sh;
  float* shift = sh + c;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int j = 0; j < BN_COPIES; ++j) {
      s1 += sums[size_t(j) * 2 * c + ch];
      s2 += sums[size_t(j) * 2 * c + c + ch];
    }
    const double mean = s1 / double(n);
    double var = s2 / double(n) - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = float(1.0 / sqrt(var + double(eps)));
    const float g = gamma ? gamma[ch] : 1.f, bta = beta ? beta[ch] : 0.f;
    scale[ch] = g * invstd;
    shift[ch] = bta - float(mean) * g * invstd;
    if (blockIdx.x == 0) {
      save_mean[ch] = float(mean);
      save_invstd[ch] = invstd;
      if (running_mean) {
        const double unbiased = n > 1 ? var * double(n) / double(n - 1) : var;
        running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * float(mean);
        running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * float(unbiased);
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches_tracked) *num_batches_tracked += 1;
  // the other half of the caller's accumulator scratch is cleared here for the NEXT BatchNorm launch on this stream
  // (its previous readers finished before this kernel started), so no memset node is needed per call
  if (zero_next)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < BN_SCRATCH_DOUBLES; i += gridDim.x * blockDim.x) zero_next[i] = 0.0;
  __syncthreads();
  const int64_t total4 = (n * int64_t(c)) >> 2;   // c % 4 == 0 (checked by the caller)
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total4; i += int64_t(gridDim.x) * blockDim.x) {
    const int ch = int((i << 2) % c);
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    v.x = fmaf(v.x, scale[ch], shift[ch]);
    v.y = fmaf(v.y, scale[ch + 1], shift[ch + 1]);
    v.z = fmaf(v.z, scale[ch + 2], shift[ch + 2]);
    v.w = fmaf(v.w, scale[ch + 3], shift[ch + 3]);
    if (res) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(res) + i);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (relu) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    }
    reinterpret_cast<float4*>(z)[i] = v;
  }
}

// sums of dy and dy * xhat, dy = dz * (z > 0) when relu
template <int U>
__global__ void __launch_bounds__(BN_TX * BN_TY)
bn_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ z, const float* __restrict__ dz, int64_t n, int c,
                    const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
                    double* __restrict__ sums /*[BN_COPIES][2c]: sum dy, sum dy*xhat*/) {
  pdl_grid_sync();
  const int ch = blockIdx.x * BN_TX + threadIdx.x;
  const int64_t r0 = int64_t(blockIdx.y) * BN_ROWS + threadIdx.y;
  const int64_t r1 = min(n, int64_t(blockIdx.y + 1) * BN_ROWS);
  float a = 0.f, b = 0.f;
  if (ch < c) {
    const float m = mean[ch], is = invstd[ch];
    for (int64_t r = r0; r < r1; r += U * BN_TY) {
      float g[U], zz[U], v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = r + u * BN_TY < r1;
        const int64_t i = (r + u * BN_TY) * c + ch;
        g[u] = ok ? __ldg(dz + i) : 0.f;
        zz[u] = (ok && relu) ? __ldg(z + i) : 1.f;
        v[u] = ok ? __ldg(x + i) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float dy = zz[u] > 0.f ? g[u] : 0.f;
        a += dy;
        b = fmaf(dy, (v[u] - m) * is, b);
      }
    }
  }
  bn_block_reduce(a, b, ch, c, sums);
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ z, const float* __restrict__ dz, int64_t n, int c,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const double* __restrict__ sums, int relu, float* __restrict__ dx, float* __restrict__ dres,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, double* __restrict__ zero_next) {
  pdl_grid_sync();
  extern __shared__ float sh[];   // k1[c] = gamma*invstd, k2[c] = mean(dy), k3[c] = mean(dy*xhat), m[c], is[c]
  float* k1 = sh;
  float* k2 = sh + c;
  float* k3 = sh + 2 * c;
  float* mm = sh + 3 * c;
  float* is = sh + 4 * c;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const float g = gamma ? gamma[ch] : 1.f;
    k1[ch] = g * invstd[ch];
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int j = 0; j < BN_COPIES; ++j) {
      s1 += sums[size_t(j) * 2 * c + ch];
      s2 += sums[size_t(j) * 2 * c + c + ch];
    }
    k2[ch] = float(s1 / double(n));
    k3[ch] = float(s2 / double(n));
    mm[ch] = mean[ch];
    is[ch] = invstd[ch];
    if (blockIdx.x == 0) {
      if (dgamma) dgamma[ch] = float(s2);
      if (dbeta) dbeta[ch] = float(s1);
    }
  }
  if (zero_next)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < BN_SCRATCH_DOUBLES; i += gridDim.x * blockDim.x) zero_next[i] = 0.0;
  __syncthreads();
  const int64_t total4 = (n * int64_t(c)) >> 2;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total4; i += int64_t(gridDim.x) * blockDim.x) {
    const int ch = int((i << 2) % c);
    float4 g = __ldg(reinterpret_cast<const float4*>(dz) + i);
    if (relu) {
      const float4 zz = __ldg(reinterpret_cast<const float4*>(z) + i);
      if (!(zz.x > 0.f)) g.x = 0.f;
      if (!(zz.y > 0.f)) g.y = 0.f;
      if (!(zz.z > 0.f)) g.z = 0.f;
      if (!(zz.w > 0.f)) g.w = 0.f;
    }
    if (dres) reinterpret_cast<float4*>(dres)[i] = g;
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + i);
    float4 o;
    o.x = k1[ch] * (g.x - k2[ch] - (xv.x - mm[ch]) * is[ch] * k3[ch]);
    o.y = k1[ch + 1] * (g.y - k2[ch + 1] - (xv.y - mm[ch + 1]) * is[ch + 1] * k3[ch + 1]);
    o.z = k1[ch + 2] * (g.z - k2[ch + 2] - (xv.z - mm[ch + 2]) * is[ch + 2] * k3[ch + 2]);
    o.w = k1[ch + 3] * (g.w - k2[ch + 3] - (xv.w - mm[ch + 3]) * is[ch + 3] * k3[ch + 3]);
    reinterpret_cast<float4*>(dx)[i] = o;
  }
}

}  // namespace lgs

using namespace lgs;

// threads of a vectorised statistics block: the largest multiple of c/4 that fits (c <= 1024 => c/4 <= 256)
static inline int bn_vec_threads(int c) { return (BN_VTHREADS / (c >> 2)) * (c >> 2); }
static inline bool bn_use_vec(const void* a, const void* b, const void* d, int c) {
  const bool off = getenv("LGS_BN_SCALAR") != nullptr;
  const uintptr_t bits = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(d);
  return !off && (bits & 15) == 0 && (c & 3) == 0 && c >= 4 && c <= 1024;
}

extern "C" {

int lgs_bn_fwd2(const float* d_x, const float* d_residual, int64_t n, int32_t c, const float* d_gamma, const float* d_beta,
                float eps, float momentum, int32_t relu, float* d_running_mean, float* d_running_var, float* d_z,
                float* d_save_mean, float* d_save_invstd, double* d_scratch, double* d_scratch_next,
                int64_t* d_num_batches_tracked, int32_t stats_ready, void* stream_);

int lgs_bn_fwd(const float* d_x, const float* d_residual, int64_t n, int32_t c, const float* d_gamma, const float* d_beta,
               float eps, float momentum, int32_t relu, float* d_running_mean, float* d_running_var, float* d_z,
               float* d_save_mean, float* d_save_invstd, double* d_scratch /*[16c]*/, double* d_scratch_next /*[16384] or NULL*/,
               int64_t* d_num_batches_tracked, void* stream_) {
  return lgs_bn_fwd2(d_x, d_residual, n, c, d_gamma, d_beta, eps, momentum, relu, d_running_mean, d_running_var, d_z, d_save_mean,
                     d_save_invstd, d_scratch, d_scratch_next, d_num_batches_tracked, 0, stream_);
}

int lgs_bn_fwd2(const float* d_x, const float* d_residual, int64_t n, int32_t c, const float* d_gamma, const float* d_beta,
                float eps, float momentum, int32_t relu, float* d_running_mean, float* d_running_var, float* d_z,
                float* d_save_mean, float* d_save_invstd, double* d_scratch /*[16c]*/, double* d_scratch_next /*[16384] or NULL*/,
                int64_t* d_num_batches_tracked, int32_t stats_ready, void* stream_) {
  if (stats_ready) LGS_TRACE("lgs_bn_fwd2 %p %p %lld %d %p %p %.9g %.9g %d %p %p %p %p %p %p %p %p %d %p", (const void*)d_x, (const void*)d_residual, (long long)n, (int)c, (const void*)d_gamma, (const void*)d_beta, (double)eps, (double)momentum, (int)relu, (const void*)d_running_mean, (const void*)d_running_var, (const void*)d_z, (const void*)d_save_mean, (const void*)d_save_invstd, (const void*)d_scratch, (const void*)d_scratch_next, (const void*)d_num_batches_tracked, (int)stats_ready, (const void*)stream_);
  LGS_TRACE("lgs_bn_fwd %p %p %lld %d %p %p %.9g %.9g %d %p %p %p %p %p %p %p %p %p", (const void*)d_x, (const void*)d_residual, (long long)n, (int)c, (const void*)d_gamma, (const void*)d_beta, (double)eps, (double)momentum, (int)relu, (const void*)d_running_mean, (const void*)d_running_var, (const void*)d_z, (const void*)d_save_mean, (const void*)d_save_invstd, (const void*)d_scratch, (const void*)d_scratch_next, (const void*)d_num_batches_tracked, (const void*)stream_);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n < 1 || c < 4 || (c & 3) || c > 1024) return fail(LGS_E_UNSUPPORTED, "lgs_bn_fwd: n=%lld c=%d (need c %% 4 == 0, c <= 1024)", (long long)n, c);
  if (!d_x || !d_z || !d_save_mean || !d_save_invstd || !d_scratch) return fail(LGS_E_INVALID, "lgs_bn_fwd: null pointer");
  // with d_scratch_next the caller guarantees d_scratch is already zero (cleared by the previous call's apply kernel)
  if (!d_scratch_next && !stats_ready) LGS_CUDA(cudaMemsetAsync(d_scratch, 0, size_t(BN_COPIES) * 2 * c * sizeof(double), stream));
  const dim3 grid{unsigned((c + BN_TX - 1) / BN_TX), unsigned(cdiv(n, BN_ROWS)), 1u}, block{BN_TX, BN_TY, 1u};
  static const int unroll = getenv("LGS_BN_UNROLL") ? atoi(getenv("LGS_BN_UNROLL")) : 4;
  if (stats_ready) {
    // the column sums were accumulated by the producing convolution's epilogue (lgs_conv_fwd2, d_bn_sums): no statistics pass
  } else if (bn_use_vec(d_x, nullptr, nullptr, c)) {
    LGS_LAUNCH_PDL(bn_stats_vec_kernel<4>, unsigned(cdiv(n, BN_VROWS)), BN_VTHREADS, 0, stream,
               reinterpret_cast<const float4*>(d_x), n, c, bn_vec_threads(c), d_scratch);
  } else if (unroll == 1) {
    LGS_LAUNCH_PDL(bn_stats_kernel<1>, grid, block, 0, stream, d_x, n, c, d_scratch);
  } else {
    LGS_LAUNCH_PDL(bn_stats_kernel<4>, grid, block, 0, stream, d_x, n, c, d_scratch);
  }
  const int64_t total4 = n * c / 4;
  int blocks = int(cdiv(total4, 256 * 4));
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  LGS_LAUNCH_PDL(bn_apply_kernel, blocks, 256, size_t(2 * c) * sizeof(float), stream, d_x, d_residual, n, c, d_scratch, d_gamma, d_beta,
             eps, relu, d_z, d_save_mean, d_save_invstd, d_running_mean, d_running_var, momentum,
             reinterpret_cast<long long*>(d_num_batches_tracked), d_scratch_next);
  return LGS_OK;
}

int lgs_bn_bwd(const float* d_x, const float* d_z, const float* d_dz, int64_t n, int32_t c, const float* d_gamma,
               const float* d_save_mean, const float* d_save_invstd, int32_t relu, float* d_dx, float* d_dresidual,
               float* d_dgamma, float* d_dbeta, double* d_scratch /*[16c]*/, double* d_scratch_next /*[16384] or NULL*/,
               void* stream_) {
  LGS_TRACE("lgs_bn_bwd %p %p %p %lld %d %p %p %p %d %p %p %p %p %p %p %p", (const void*)d_x, (const void*)d_z, (const void*)d_dz, (long long)n, (int)c, (const void*)d_gamma, (const void*)d_save_mean, (const void*)d_save_invstd, (int)relu, (const void*)d_dx, (const void*)d_dresidual, (const void*)d_dgamma, (const void*)d_dbeta, (const void*)d_scratch, (const void*)d_scratch_next, (const void*)stream_);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n < 1 || c < 4 || (c & 3) || c > 1024) return fail(LGS_E_UNSUPPORTED, "lgs_bn_bwd: n=%lld c=%d", (long long)n, c);
  if (!d_x || !d_dz || !d_dx || !d_save_mean || !d_save_invstd || !d_scratch || (relu && !d_z))
    return fail(LGS_E_INVALID, "lgs_bn_bwd: null pointer");
  if (!d_scratch_next) LGS_CUDA(cudaMemsetAsync(d_scratch, 0, size_t(BN_COPIES) * 2 * c * sizeof(double), stream));
  const dim3 grid{unsigned((c + BN_TX - 1) / BN_TX), unsigned(cdiv(n, BN_ROWS)), 1u}, block{BN_TX, BN_TY, 1u};
  static const int unroll = getenv("LGS_BN_UNROLL") ? atoi(getenv("LGS_BN_UNROLL")) : 4;
  if (bn_use_vec(d_x, relu ? d_z : nullptr, d_dz, c) && !(reinterpret_cast<uintptr_t>(d_save_mean) & 15) &&
      !(reinterpret_cast<uintptr_t>(d_save_invstd) & 15)) {
    LGS_LAUNCH_PDL(bn_bwd_stats_vec_kernel<2>, unsigned(cdiv(n, BN_VROWS)), BN_VTHREADS, 0, stream,
               reinterpret_cast<const float4*>(d_x), reinterpret_cast<const float4*>(d_z),
               reinterpret_cast<const float4*>(d_dz), n, c, bn_vec_threads(c), d_save_mean, d_save_invstd, relu, d_scratch);
  } else if (unroll == 1) {
    LGS_LAUNCH_PDL(bn_bwd_stats_kernel<1>, grid, block, 0, stream, d_x, d_z, d_dz, n, c, d_save_mean, d_save_invstd, relu, d_scratch);
  } else {
    LGS_LAUNCH_PDL(bn_bwd_stats_kernel<4>, grid, block, 0, stream, d_x, d_z, d_dz, n, c, d_save_mean, d_save_invstd, relu, d_scratch);
  }
  const int64_t total4 = n * c / 4;
  int blocks = int(cdiv(total4, 256 * 4));
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  LGS_LAUNCH_PDL(bn_bwd_apply_kernel, blocks, 256, size_t(5 * c) * sizeof(float), stream, d_x, d_z, d_dz, n, c, d_save_mean,
             d_save_invstd, d_gamma, d_scratch, relu, d_dx, d_dresidual, d_dgamma, d_dbeta, d_scratch_next);
  return LGS_OK;
}

}  // extern "C"
