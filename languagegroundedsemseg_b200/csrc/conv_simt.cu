// LGS_ALGO_SIMT: exact-fp32 output-stationary sparse convolution (forward / dgrad) and wgrad.
// This is the parity anchor of the engine (fp32 FMA, deterministic forward: no atomics in fwd/dgrad) and the
// fallback for shapes the tcgen05 path does not take.  Replaces ME matmul/matmul2 "DIRECT_GEMM" kernels
//  with an implicit GEMM: per 64-row output tile, loop over the K offsets, gather the neighbour
// rows through the kernel-map table into shared memory and accumulate in registers; one store per output element.
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"

namespace lgs {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int AS_STRIDE = BM + 4;

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// 4 consecutive elements starting at p[0], elements >= valid are zero; `vec` = 4-element alignment holds
template <typename T>
__device__ __forceinline__ float4 load4(const T* __restrict__ p, int valid, bool vec) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid <= 0) return r;
  if (vec && valid >= 4) {
    if constexpr (sizeof(T) == 4) {
      r = __ldg(reinterpret_cast<const float4*>(p));
    } else {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
      const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
      r = make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
    }
    return r;
  }
  r.x = to_f<T>(p[0]);
  if (valid > 1) r.y = to_f<T>(p[1]);
  if (valid > 2) r.z = to_f<T>(p[2]);
  if (valid > 3) r.w = to_f<T>(p[3]);
  return r;
}

template <typename T, bool W_KNC>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const T* __restrict__ in, int c_in, const T* __restrict__ w, int K, int c_out,
                 const int32_t* __restrict__ table, int64_t n_out, int reverse_k, const float* __restrict__ bias,
                 T* __restrict__ out) {
  __shared__ __align__(16) float As[BK][AS_STRIDE];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ int32_t rows[BM];

  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int64_t m0 = int64_t(blockIdx.x) * BM;
  const int n0 = blockIdx.y * BN;
  const bool vec_in = (c_in & 3) == 0, vec_w = (c_out & 3) == 0;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int a_row = t >> 2, a_chunk = t & 3;  // A loader: 64 rows x 4 chunks of 4 channels
  const int b_kk = t >> 4, b_n4 = t & 15;     // B loader: 16 k-rows x 16 chunks of 4 columns

  for (int k = 0; k < K; ++k) {
    const int kk_tab = reverse_k ? (K - 1 - k) : k;
    int32_t myrow = -1;
    if (t < BM) {
      const int64_t o = m0 + t;
      if (o < n_out) myrow = table ? __ldg(table + int64_t(kk_tab) * n_out + o) : int32_t(o);
      rows[t] = myrow;
    }
    if (!__syncthreads_or(myrow >= 0)) continue;  // no neighbour at this offset anywhere in the tile
    const int32_t r = rows[a_row];
    const T* wk = w + size_t(k) * c_in * c_out;
    for (int c0 = 0; c0 < c_in; c0 += BK) {
      const int ca = c0 + a_chunk * 4;
      const float4 av = (r >= 0) ? load4<T>(in + size_t(r) * c_in + ca, c_in - ca, vec_in)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
      As[a_chunk * 4 + 0][a_row] = av.x;
      As[a_chunk * 4 + 1][a_row] = av.y;
      As[a_chunk * 4 + 2][a_row] = av.z;
      As[a_chunk * 4 + 3][a_row] = av.w;
      if constexpr (W_KNC) {
        // weights stored [K, c_out, c_in]: same access pattern as the A loader (row = output channel)
        const int n = n0 + a_row;
        const float4 bv = (n < c_out) ? load4<T>(wk + size_t(n) * c_in + ca, c_in - ca, vec_in)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        Bs[a_chunk * 4 + 0][a_row] = bv.x;
        Bs[a_chunk * 4 + 1][a_row] = bv.y;
        Bs[a_chunk * 4 + 2][a_row] = bv.z;
        Bs[a_chunk * 4 + 3][a_row] = bv.w;
      } else {
        const int cb = c0 + b_kk, nb = n0 + b_n4 * 4;
        const float4 bv = (cb < c_in) ? load4<T>(wk + size_t(cb) * c_out + nb, c_out - nb, vec_w)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(&Bs[b_kk][b_n4 * 4]) = bv;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t o = m0 + ty * 4 + i;
    if (o >= n_out) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < c_out) out[size_t(o) * c_out + n] = from_f<T>(acc[i][j] + (bias ? __ldg(bias + n) : 0.f));
    }
  }
}

// wgrad: grid = (ci_tiles * co_tiles, K, row_chunks); 64x64 tile of dW[k] per CTA, reduction over a chunk of rows.
template <typename T>
__global__ void __launch_bounds__(256)
wgrad_simt_kernel(const T* __restrict__ in, int c_in, const T* __restrict__ gout, int c_out,
                  const int32_t* __restrict__ table, int64_t n_out, int64_t rows_per_chunk, int co_tiles,
                  float* __restrict__ gw) {
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int ci0 = (blockIdx.x / co_tiles) * BM, co0 = (blockIdx.x % co_tiles) * BN;
  const int k = blockIdx.y;
  const int64_t r_begin = int64_t(blockIdx.z) * rows_per_chunk;
  const int64_t r_end = min(n_out, r_begin + rows_per_chunk);
  const bool vec_in = (c_in & 3) == 0, vec_out = (c_out & 3) == 0;
  const int l_r = t >> 4, l_c4 = t & 15;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t r0 = r_begin; r0 < r_end; r0 += BK) {
    const int64_t o = r0 + l_r;
    int32_t src = -1;
    if (o < r_end) src = table ? __ldg(table + int64_t(k) * n_out + o) : int32_t(o);
    if (!__syncthreads_or(src >= 0)) continue;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
    if (src >= 0) {
      const int ca = ci0 + l_c4 * 4, cb = co0 + l_c4 * 4;
      av = load4<T>(in + size_t(src) * c_in + ca, c_in - ca, vec_in);
      bv = load4<T>(gout + size_t(o) * c_out + cb, c_out - cb, vec_out);
    }
    *reinterpret_cast<float4*>(&As[l_r][l_c4 * 4]) = av;
    *reinterpret_cast<float4*>(&Bs[l_r][l_c4 * 4]) = bv;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* gwk = gw + size_t(k) * c_in * c_out;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + ty * 4 + i;
    if (ci >= c_in) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co < c_out && acc[i][j] != 0.f) atomicAdd(gwk + size_t(ci) * c_out + co, acc[i][j]);
    }
  }
}

template <typename T>
static int launch_conv_simt(const void* in, int c_in, const void* w, int w_layout, int K, int c_out,
                            const int32_t* table, int64_t n_out, int reverse_k, const float* bias, void* out,
                            cudaStream_t stream) {
  dim3 grid(unsigned(cdiv(n_out, BM)), unsigned(cdiv(c_out, BN)));
  if (w_layout == LGS_W_KNC) {
    LGS_LAUNCH((conv_simt_kernel<T, true>), grid, 256, 0, stream, static_cast<const T*>(in), c_in,
               static_cast<const T*>(w), K, c_out, table, n_out, reverse_k, bias, static_cast<T*>(out));
  } else {
    LGS_LAUNCH((conv_simt_kernel<T, false>), grid, 256, 0, stream, static_cast<const T*>(in), c_in,
               static_cast<const T*>(w), K, c_out, table, n_out, reverse_k, bias, static_cast<T*>(out));
  }
  return LGS_OK;
}

int conv_fwd_simt(const void* in, int64_t n_in, int c_in, const void* w, int w_layout, int K, int c_out,
                  const int32_t* table, int64_t n_out, int reverse_k, const float* bias, void* out, int dtype,
                  cudaStream_t stream) {
  (void)n_in;
  if (n_out == 0) return LGS_OK;
  if (dtype == LGS_F32)
    return launch_conv_simt<float>(in, c_in, w, w_layout, K, c_out, table, n_out, reverse_k, bias, out, stream);
  return launch_conv_simt<__nv_bfloat16>(in, c_in, w, w_layout, K, c_out, table, n_out, reverse_k, bias, out, stream);
}

// ---- wgrad of the stem: c_in = 4 (3 colour channels + pad), c_out = 32, K = 27 (res16unet.py:38 conv0p1s1) ------------------
// 149 K rows x 27 offsets x 4 x 32 = 0.5 GMAC of exact fp32 FMAs: the tensor-core wgrad spends 230 us on it (128 MMA lanes for
// 4 input channels, one 16-byte cp.async per (row, offset)); here a block stages 64 output rows — their 27 neighbour rows of X
// (one float4 each) and their dY rows — in shared memory with coalesced loads and thread (co, offset group) keeps its
// 4 offsets x 4 channels of dW in registers over all of the block's tiles; one atomicAdd per dW element and block at the end.
constexpr int WS_TILE = 64, WS_K = 27;
__global__ void __launch_bounds__(256) wgrad_stem_kernel(const float4* __restrict__ x, const float* __restrict__ gy,
                                                         const int32_t* __restrict__ table, int64_t n_out,
                                                         float* __restrict__ gw /*[27][4][32]*/) {
  __shared__ float4 xs[WS_K][WS_TILE];
  __shared__ float dys[WS_TILE][32];
  const int tid = threadIdx.x, co = tid & 31, kg = tid >> 5;      // offsets kg, kg + 8, kg + 16, kg + 24 (< 27)
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const int64_t tiles = (n_out + WS_TILE - 1) / WS_TILE;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t o0 = t * WS_TILE;
    __syncthreads();                                              // the previous tile's reads are done
    for (int i = tid; i < WS_K * WS_TILE; i += 256) {
      const int k = i / WS_TILE, r = i - k * WS_TILE;
      const int64_t o = o0 + r;
      const int32_t j = o < n_out ? __ldg(table + int64_t(k) * n_out + o) : -1;
      xs[k][r] = j >= 0 ? __ldg(x + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = tid; i < WS_TILE * 32; i += 256) {
      const int r = i >> 5;
      dys[r][i & 31] = o0 + r < n_out ? __ldg(gy + (o0 + r) * 32 + (i & 31)) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < WS_TILE; ++r) {
      const float d = dys[r][co];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int k = kg + 8 * a;
        if (k < WS_K) {
          const float4 v = xs[k][r];                              // same address for the whole warp: broadcast
          acc[a][0] = fmaf(v.x, d, acc[a][0]);
          acc[a][1] = fmaf(v.y, d, acc[a][1]);
          acc[a][2] = fmaf(v.z, d, acc[a][2]);
          acc[a][3] = fmaf(v.w, d, acc[a][3]);
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int k = kg + 8 * a;
    if (k < WS_K)
#pragma unroll
      for (int b = 0; b < 4; ++b) atomicAdd(gw + (k * 4 + b) * 32 + co, acc[a][b]);
  }
}

// 1 if conv_wgrad_stem takes this layer (fp32, K = 27, c_in = 4, c_out = 32, 16-byte aligned rows)
int conv_wgrad_stem(const void* in, int c_in, const void* gout, int64_t n_out, int c_out, const int32_t* table, int K, float* gw,
                    int dtype, cudaStream_t stream) {
  if (dtype != LGS_F32 || K != WS_K || c_in != 4 || c_out != 32 || !table || (reinterpret_cast<uintptr_t>(in) & 15))
    return LGS_E_UNSUPPORTED;
  {
    const int rz = zero_fill_async(gw, size_t(K) * c_in * c_out * sizeof(float), stream);
    if (rz != LGS_OK) return rz;
  }
  if (n_out == 0) return LGS_OK;
  const int64_t tiles = cdiv(n_out, WS_TILE);
  const unsigned grid = unsigned(std::min<int64_t>(tiles, 148 * 4));
  LGS_LAUNCH(wgrad_stem_kernel, grid, 256, 0, stream, static_cast<const float4*>(in), static_cast<const float*>(gout), table, n_out, gw);
  return LGS_OK;
}

int conv_wgrad_simt(const void* in, int c_in, const void* gout, int64_t n_out, int c_out, const int32_t* table, int K,
                    float* gw, int dtype, cudaStream_t stream) {
  LGS_CUDA(cudaMemsetAsync(gw, 0, size_t(K) * c_in * c_out * sizeof(float), stream));
  if (n_out == 0) return LGS_OK;
  const int ci_tiles = int(cdiv(c_in, BM)), co_tiles = int(cdiv(c_out, BN));
  const int64_t tiles = int64_t(ci_tiles) * co_tiles * K;
  int64_t chunks = cdiv(148 * 6, tiles);
  const int64_t max_chunks = cdiv(n_out, 512);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  if (chunks > 65535) chunks = 65535;
  int64_t rows_per_chunk = cdiv(cdiv(n_out, chunks), BK) * BK;
  chunks = cdiv(n_out, rows_per_chunk);
  dim3 grid(unsigned(ci_tiles * co_tiles), unsigned(K), unsigned(chunks));
  if (dtype == LGS_F32) {
    LGS_LAUNCH(wgrad_simt_kernel<float>, grid, 256, 0, stream, static_cast<const float*>(in), c_in,
               static_cast<const float*>(gout), c_out, table, n_out, rows_per_chunk, co_tiles, gw);
  } else {
    LGS_LAUNCH(wgrad_simt_kernel<__nv_bfloat16>, grid, 256, 0, stream, static_cast<const __nv_bfloat16*>(in), c_in,
               static_cast<const __nv_bfloat16*>(gout), c_out, table, n_out, rows_per_chunk, co_tiles, gw);
  }
  return LGS_OK;
}

}  // namespace lgs
