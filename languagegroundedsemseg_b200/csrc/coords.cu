// Coordinate manager kernels: cuckoo-hash coordinate maps, strided maps, output-stationary kernel maps,
// on-device voxel quantisation.  Integer / byte work, L2-latency bound: one thread per coordinate (or per
// (offset, coordinate) probe), coalesced int4 coordinate loads, warp-ballot counting.
//
// Replaces: ME insert_and_map_kernel + concurrent_unordered_map, stride_map kernels,
// count_kernel + preallocated_kernel_map_iteration + thrust sort; lib/voxelizer.py:138-142.
#include <cstring>
#include <mutex>
#include <string>

#include "common.cuh"

namespace lgs {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_trace_on{0};
std::atomic<int> g_pdl{1};
static std::mutex g_trace_mu;
static std::string g_trace_buf;

void trace_record(const char* fmt, ...) {
  char line[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(line, sizeof(line), fmt, ap);
  va_end(ap);
  std::lock_guard<std::mutex> lock(g_trace_mu);
  g_trace_buf.append(line);
  g_trace_buf.push_back('\n');
}

constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanBlock * kScanItems;  // 2048 rows per scan block

__device__ __forceinline__ int32_t floor_to(int32_t v, int32_t q) {
  // floor(v / q) * q for q > 0 (C division truncates toward zero)
  int32_t r = v % q;
  return v - (r < 0 ? r + q : r);
}

__device__ __forceinline__ int4 load_quantised(const int32_t* __restrict__ coords, int64_t i, int32_t quant) {
  int4 c = __ldg(reinterpret_cast<const int4*>(coords) + i);  // (b,x,y,z): one 16-byte load per row
  if (quant > 1) {
    c.y = floor_to(c.y, quant);
    c.z = floor_to(c.z, quant);
    c.w = floor_to(c.w, quant);
  }
  return c;
}

// ---- 1. cuckoo insert ---------------------------------------------------------------------------------
// flags[0]: range error, flags[1]: eviction chain too long.
__global__ void __launch_bounds__(256) insert_kernel(const int32_t* __restrict__ coords, int64_t n, int32_t quant,
                                                     unsigned long long* keys, uint32_t mask, int32_t* flags) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int4 c = load_quantised(coords, i, quant);
  if (!key_in_range(c.x, c.y, c.z, c.w)) {
    flags[0] = 1;
    return;
  }
  uint64_t cur = pack_key(c.x, c.y, c.z, c.w);
  // cheap de-dup: already resident?  (racy by design — a duplicate slipping through is harmless, lookups
  // always resolve to the first match in probe order)
#pragma unroll
  for (int j = 0; j < kNumHashes; ++j) {
    const uint64_t k = keys[hash_slot(cur, j, mask)];
    if (k == cur) return;
    if (j == 0 && k == kEmptyKey) break;
  }
  int j = 0;
  for (int it = 0; it < kMaxEvictions; ++it) {
    const uint32_t s = hash_slot(cur, j, mask);
    const uint64_t old = atomicExch(keys + s, (unsigned long long)cur);
    if (old == kEmptyKey || old == cur) return;
    // re-home the evicted key at the hash function after the one it was sitting at
    cur = old;
    int jj = 0;
#pragma unroll
    for (int t = 0; t < kNumHashes; ++t)
      if (hash_slot(cur, t, mask) == s) jj = t;
    j = (jj + 1) % kNumHashes;
  }
  flags[1] = 1;
}

// ---- 2. first occurrence wins: atomicMin(row id) at the key's slot ------------------------------------
__global__ void __launch_bounds__(256) claim_kernel(const int32_t* __restrict__ coords, int64_t n, int32_t quant,
                                                    const uint64_t* __restrict__ keys, int32_t* vals, uint32_t mask,
                                                    int32_t* __restrict__ slot_of_row) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int4 c = load_quantised(coords, i, quant);
  int32_t s = -1;
  if (key_in_range(c.x, c.y, c.z, c.w)) s = cuckoo_find(keys, mask, pack_key(c.x, c.y, c.z, c.w));
  slot_of_row[i] = s;
  if (s >= 0) atomicMin(vals + s, int32_t(i));
}

// ---- 3. exclusive scan of "row i is the first occurrence" --------------------------------------------
__global__ void __launch_bounds__(kScanBlock) scan_block_kernel(const int32_t* __restrict__ slot_of_row,
                                                                const int32_t* __restrict__ vals, int64_t n,
                                                                int32_t* __restrict__ local_rank,
                                                                int32_t* __restrict__ block_total) {
  __shared__ int32_t warp_sums[kScanBlock / 32];
  const int64_t base = blockIdx.x * int64_t(kScanTile) + threadIdx.x * kScanItems;
  int32_t f[kScanItems];
  int32_t sum = 0;
#pragma unroll
  for (int t = 0; t < kScanItems; ++t) {
    const int64_t i = base + t;
    int32_t v = 0;
    if (i < n) {
      const int32_t s = slot_of_row[i];
      v = (s >= 0 && vals[s] == int32_t(i)) ? 1 : 0;
    }
    f[t] = sum;  // exclusive within the thread
    sum += v;
  }
  // warp inclusive scan of per-thread sums
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int32_t y = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += y;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int32_t w = lane < kScanBlock / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += y;
    }
    if (lane < kScanBlock / 32) warp_sums[lane] = w;  // inclusive
  }
  __syncthreads();
  const int32_t thread_off = (inc - sum) + (warp ? warp_sums[warp - 1] : 0);
#pragma unroll
  for (int t = 0; t < kScanItems; ++t) {
    const int64_t i = base + t;
    if (i < n) local_rank[i] = thread_off + f[t];
  }
  if (threadIdx.x == kScanBlock - 1) block_total[blockIdx.x] = thread_off + sum;
}

// single block: exclusive scan of block totals in place, total -> block_total[nb] and n_unique
__global__ void __launch_bounds__(1024) scan_totals_kernel(int32_t* block_total, int32_t nb, int32_t* n_unique) {
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int32_t base = 0; base < nb; base += 1024) {
    const int32_t i = base + threadIdx.x;
    const int32_t v = i < nb ? block_total[i] : 0;
    int32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += y;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int32_t w = warp_sums[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const int32_t excl = carry + (inc - v) + (warp ? warp_sums[warp - 1] : 0);
    if (i < nb) block_total[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    block_total[nb] = carry;
    *n_unique = carry;
  }
}

// ---- 4. emit unique rows + inverse map, then relabel the table with unique-row ids ---------------------
__global__ void __launch_bounds__(256) emit_kernel(const int32_t* __restrict__ coords, int64_t n, int32_t quant,
                                                   const int32_t* __restrict__ slot_of_row,
                                                   const int32_t* __restrict__ vals,
                                                   const int32_t* __restrict__ local_rank,
                                                   const int32_t* __restrict__ block_off,
                                                   int32_t* __restrict__ out_coords, int32_t* __restrict__ unique_index,
                                                   int32_t* __restrict__ inverse) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int32_t s = slot_of_row[i];
  if (s < 0) {
    if (inverse) inverse[i] = -1;
    return;
  }
  const int32_t first = vals[s];
  const int32_t u = local_rank[first] + block_off[first / kScanTile];
  if (inverse) inverse[i] = u;
  if (first == int32_t(i)) {
    reinterpret_cast<int4*>(out_coords)[u] = load_quantised(coords, i, quant);
    if (unique_index) unique_index[u] = int32_t(i);
  }
}

__global__ void __launch_bounds__(256) relabel_kernel(int64_t n, const int32_t* __restrict__ slot_of_row,
                                                      int32_t* vals, const int32_t* __restrict__ local_rank,
                                                      const int32_t* __restrict__ block_off) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int32_t s = slot_of_row[i];
  if (s >= 0 && vals[s] == int32_t(i)) vals[s] = local_rank[i] + block_off[i / kScanTile];
}

// ---- kernel map: one probe per (offset, output row) ----------------------------------------------------
__global__ void __launch_bounds__(256) kmap_kernel(const int32_t* __restrict__ out_coords, int64_t n_out,
                                                   const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals,
                                                   uint32_t mask, int32_t ks, int32_t step, int32_t K,
                                                   int32_t* __restrict__ table, int32_t* counts) {
  __shared__ int32_t s_counts[64];
  if (threadIdx.x < 64) s_counts[threadIdx.x] = 0;
  __syncthreads();
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const int64_t total = int64_t(K) * n_out;
  int32_t k = -1;
  bool hit = false;
  if (idx < total) {
    k = int32_t(idx / n_out);
    const int64_t o = idx - int64_t(k) * n_out;
    const int4 c = __ldg(reinterpret_cast<const int4*>(out_coords) + o);
    const int32_t centre = (ks & 1) ? ks / 2 : 0;
    const int32_t ix = k % ks, iy = (k / ks) % ks, iz = k / (ks * ks);  // x fastest (App. A.5)
    const int32_t x = c.y + (ix - centre) * step, y = c.z + (iy - centre) * step, z = c.w + (iz - centre) * step;
    int32_t row = -1;
    if (key_in_range(c.x, x, y, z)) {
      const int32_t s = cuckoo_find(keys, mask, pack_key(c.x, x, y, z));
      if (s >= 0) row = __ldg(vals + s);
    }
    table[idx] = row;
    hit = row >= 0;
  }
  // warp-ballot count; a warp spans at most two offsets when n_out >= 32, else fall back to per-lane adds
  const int32_t k0 = __shfl_sync(0xffffffffu, k, 0);
  const uint32_t m0 = __ballot_sync(0xffffffffu, hit && k == k0);
  const uint32_t m1 = __ballot_sync(0xffffffffu, hit && k == k0 + 1);
  const uint32_t rest = __ballot_sync(0xffffffffu, hit && k != k0 && k != k0 + 1);
  if ((threadIdx.x & 31) == 0) {
    if (m0) atomicAdd(&s_counts[k0], __popc(m0));
    if (m1) atomicAdd(&s_counts[k0 + 1], __popc(m1));
  }
  if (rest && hit && k != k0 && k != k0 + 1) atomicAdd(&s_counts[k], 1);
  __syncthreads();
  if (threadIdx.x < K && s_counts[threadIdx.x]) atomicAdd(counts + threadIdx.x, s_counts[threadIdx.x]);
}

__global__ void __launch_bounds__(256) kmap_transpose_kernel(const int32_t* __restrict__ table, int64_t total,
                                                             int64_t n_out, int64_t n_in,
                                                             int32_t* __restrict__ table_t) {
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int32_t i = table[idx];
  if (i < 0) return;
  const int64_t k = idx / n_out;
  table_t[k * n_in + i] = int32_t(idx - k * n_out);
}

// ---- voxelise: float64 affine, fixed evaluation order, no FMA ------------------------------------------
struct Affine {
  double m[12];
};

__global__ void __launch_bounds__(256) voxelize_kernel(const float* __restrict__ xyz, int64_t n, Affine A,
                                                       int32_t batch, int32_t* __restrict__ coords) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
  int32_t q[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double v = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, A.m[4 * j]), __dmul_rn(y, A.m[4 * j + 1])),
                                         __dmul_rn(z, A.m[4 * j + 2])),
                               A.m[4 * j + 3]);
    q[j] = int32_t(floor(v));
  }
  reinterpret_cast<int4*>(coords)[i] = make_int4(batch, q[0], q[1], q[2]);
}

}  // namespace lgs

using namespace lgs;

extern "C" {

int lgs_version(void) { return 100; }
const char* lgs_last_error(void) { return g_err; }
uint64_t lgs_launch_count(void) { return g_launches.load(); }

int lgs_trace_begin(void) {
  std::lock_guard<std::mutex> lock(g_trace_mu);
  g_trace_buf.clear();
  g_trace_on.store(1);
  return LGS_OK;
}

int64_t lgs_trace_end(char* buf, int64_t capacity) {
  g_trace_on.store(0);
  std::lock_guard<std::mutex> lock(g_trace_mu);
  const int64_t need = int64_t(g_trace_buf.size()) + 1;
  if (buf && capacity >= need) memcpy(buf, g_trace_buf.c_str(), size_t(need));
  return need;
}
int32_t lgs_coord_limit(void) { return kCoordLimit; }

int64_t lgs_hash_capacity(int64_t n) {
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  return cap;
}

int64_t lgs_coordmap_scratch_elems(int64_t n) { return 2 * n + cdiv(n, kScanTile) + 16; }

int lgs_coordmap_build(const int32_t* d_coords, int64_t n, int32_t quant, uint64_t* d_table_keys,
                       int32_t* d_table_vals, int64_t capacity, int32_t* d_out_coords, int32_t* d_unique_index,
                       int32_t* d_inverse, int32_t* d_scratch, int32_t* d_n_unique, int64_t* h_n_unique,
                       void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n < 0 || quant < 1 || capacity < 2 * n || (capacity & (capacity - 1)) || capacity > (int64_t(1) << 31))
    return fail(LGS_E_INVALID, "lgs_coordmap_build: n=%lld quant=%d capacity=%lld (need pow2 >= 2n)", (long long)n,
                quant, (long long)capacity);
  if (n > (int64_t(1) << 30)) return fail(LGS_E_INVALID, "lgs_coordmap_build: n too large (max 2^30 rows)");
  if (!d_table_keys || !d_table_vals || !d_scratch || !d_n_unique || (n && (!d_coords || !d_out_coords)))
    return fail(LGS_E_INVALID, "lgs_coordmap_build: null pointer");
  LGS_CUDA(cudaMemsetAsync(d_table_keys, 0xFF, size_t(capacity) * 8, stream));
  LGS_CUDA(cudaMemsetAsync(d_table_vals, 0x7F, size_t(capacity) * 4, stream));
  const int32_t nb = int32_t(cdiv(n, kScanTile));
  int32_t* slot_of_row = d_scratch;
  int32_t* local_rank = d_scratch + n;
  int32_t* block_tot = d_scratch + 2 * n;  // nb + 1 totals, then 2 flags
  int32_t* flags = block_tot + nb + 2;
  LGS_CUDA(cudaMemsetAsync(block_tot, 0, size_t(nb + 8) * 4, stream));
  const uint32_t mask = uint32_t(capacity - 1);
  if (n > 0) {
    const int grid = int(cdiv(n, 256));
    LGS_LAUNCH(insert_kernel, grid, 256, 0, stream, d_coords, n, quant,
               reinterpret_cast<unsigned long long*>(d_table_keys), mask, flags);
    LGS_LAUNCH(claim_kernel, grid, 256, 0, stream, d_coords, n, quant, d_table_keys, d_table_vals, mask,
               slot_of_row);
    LGS_LAUNCH(scan_block_kernel, nb, kScanBlock, 0, stream, slot_of_row, d_table_vals, n, local_rank, block_tot);
  }
  LGS_LAUNCH(scan_totals_kernel, 1, 1024, 0, stream, block_tot, nb, d_n_unique);
  if (n > 0) {
    const int grid = int(cdiv(n, 256));
    LGS_LAUNCH(emit_kernel, grid, 256, 0, stream, d_coords, n, quant, slot_of_row, d_table_vals, local_rank,
               block_tot, d_out_coords, d_unique_index, d_inverse);
    LGS_LAUNCH(relabel_kernel, grid, 256, 0, stream, n, slot_of_row, d_table_vals, local_rank, block_tot);
  }
  if (h_n_unique) {
    int32_t h[3] = {0, 0, 0};
    LGS_CUDA(cudaMemcpyAsync(&h[0], d_n_unique, 4, cudaMemcpyDeviceToHost, stream));
    LGS_CUDA(cudaMemcpyAsync(&h[1], flags, 8, cudaMemcpyDeviceToHost, stream));
    LGS_CUDA(cudaStreamSynchronize(stream));
    if (h[1]) return fail(LGS_E_RANGE, "coordinate outside packable range (batch < %d, |xyz| < %d)", kBatchLimit,
                          kCoordLimit);
    if (h[2]) return fail(LGS_E_HASH_FULL, "cuckoo eviction chain exceeded %d: retry with larger capacity",
                          kMaxEvictions);
    *h_n_unique = h[0];
  }
  return LGS_OK;
}

int lgs_kmap_build(const int32_t* d_out_coords, int64_t n_out, const uint64_t* d_in_table_keys,
                   const int32_t* d_in_table_vals, int64_t in_capacity, int32_t ksize, int32_t in_tensor_stride,
                   int32_t dilation, int32_t* d_table, int32_t* d_counts, void* stream_) {
  LGS_TRACE("lgs_kmap_build %p %lld %p %p %lld %d %d %d %p %p %p", (const void*)d_out_coords, (long long)n_out, (const void*)d_in_table_keys, (const void*)d_in_table_vals, (long long)in_capacity, (int)ksize, (int)in_tensor_stride, (int)dilation, (const void*)d_table, (const void*)d_counts, (const void*)stream_);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (ksize < 1 || ksize > 3 || in_tensor_stride < 1 || dilation < 1 || n_out < 0 ||
      (in_capacity & (in_capacity - 1)) || in_capacity < 1)
    return fail(LGS_E_INVALID, "lgs_kmap_build: ksize=%d (1..3) stride=%d dilation=%d cap=%lld", ksize,
                in_tensor_stride, dilation, (long long)in_capacity);
  const int32_t K = ksize * ksize * ksize;
  LGS_CUDA(cudaMemsetAsync(d_counts, 0, size_t(K) * 4, stream));
  const int64_t total = int64_t(K) * n_out;
  if (total == 0) return LGS_OK;
  LGS_LAUNCH(kmap_kernel, int(cdiv(total, 256)), 256, 0, stream, d_out_coords, n_out, d_in_table_keys,
             d_in_table_vals, uint32_t(in_capacity - 1), ksize, in_tensor_stride * dilation, K, d_table, d_counts);
  return LGS_OK;
}

int lgs_kmap_transpose(const int32_t* d_table, int32_t K, int64_t n_out, int64_t n_in, int32_t* d_table_t,
                       void* stream_) {
  LGS_TRACE("lgs_kmap_transpose %p %d %lld %lld %p %p", (const void*)d_table, (int)K, (long long)n_out, (long long)n_in, (const void*)d_table_t, (const void*)stream_);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (K < 1 || n_out < 0 || n_in < 0) return fail(LGS_E_INVALID, "lgs_kmap_transpose: bad sizes");
  LGS_CUDA(cudaMemsetAsync(d_table_t, 0xFF, size_t(K) * n_in * 4, stream));
  const int64_t total = int64_t(K) * n_out;
  if (total == 0) return LGS_OK;
  LGS_LAUNCH(kmap_transpose_kernel, int(cdiv(total, 256)), 256, 0, stream, d_table, total, n_out, n_in, d_table_t);
  return LGS_OK;
}

int lgs_voxelize_affine(const float* d_xyz, int64_t n, const double* h_M, int32_t batch, int32_t* d_coords,
                        void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n < 0 || !h_M || batch < 0 || batch >= kBatchLimit) return fail(LGS_E_INVALID, "lgs_voxelize_affine: bad args");
  if (n == 0) return LGS_OK;
  Affine A;
  for (int i = 0; i < 12; ++i) A.m[i] = h_M[i];
  LGS_LAUNCH(voxelize_kernel, int(cdiv(n, 256)), 256, 0, stream, d_xyz, n, A, batch, d_coords);
  return LGS_OK;
}

}  // extern "C"
