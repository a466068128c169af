// Shared host/device helpers of the lgs_b200 engine (error slot, launch counter, key packing, cuckoo probes).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cstdio>
#include <cstdarg>
#include <utility>

#include "../../include/lgs_b200.h"

namespace lgs {

// ---- host-side plumbing ------------------------------------------------------------------------------
extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define LGS_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::lgs::fail(LGS_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

// every kernel launch goes through here so lgs_launch_count() is the library's own claim
#define LGS_LAUNCH(kernel, grid, block, smem, stream, ...)                                          \
  do {                                                                                              \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                     \
    ::lgs::g_launches.fetch_add(1, std::memory_order_relaxed);                                      \
    LGS_CUDA(cudaGetLastError());                                                                   \
  } while (0)

// ---- programmatic dependent launch --------------------------------------------------------------------
// A kernel launched with LGS_LAUNCH_PDL may be scheduled while the previous kernel of its stream is still running (its
// CTAs become resident as the predecessor's drain) — the launch latency and the drain / fill gap between two dependent
// kernels (2-3 us, ~500 launches per training step) overlap.  Contract: such a kernel calls pdl_grid_sync() before it
// touches global memory (griddepcontrol.wait returns when the predecessor grid has completed and its writes are
// visible), which also lets ITS successor launch early.  lgs_tune("pdl", 0) switches to plain launches.
extern std::atomic<int> g_pdl;
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl.load(std::memory_order_relaxed) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
// zero-fill as a KERNEL (float4 stores; bytes a multiple of 16, 16-byte aligned): a cudaMemsetAsync node between two kernels
// breaks their programmatic dependency, so the kernel behind it pays the whole launch + drain gap
int zero_fill_async(void* p, size_t bytes, cudaStream_t stream);
#define LGS_LAUNCH_PDL(kernel, grid, block, smem, stream, ...)                                      \
  do {                                                                                              \
    LGS_CUDA(::lgs::launch_pdl(kernel, dim3(grid), dim3(block), size_t(smem), (stream), __VA_ARGS__)); \
    ::lgs::g_launches.fetch_add(1, std::memory_order_relaxed);                                      \
  } while (0)
#endif

// ---- call recorder (lgs_trace_begin / lgs_trace_end) --------------------------------------------------
// While recording, a traced entry point appends one line "name arg arg ..." and returns LGS_OK WITHOUT touching the
// GPU, so the host side of the engine (bindings, the facade's call sequence and arguments) can be exercised and
// compared on a machine with no GPU.  Off (one relaxed atomic load per call) unless lgs_trace_begin() was called.
extern std::atomic<int> g_trace_on;
void trace_record(const char* fmt, ...);
#define LGS_TRACE(...)                                              \
  do {                                                              \
    if (::lgs::g_trace_on.load(std::memory_order_relaxed)) {        \
      ::lgs::trace_record(__VA_ARGS__);                             \
      return LGS_OK;                                                \
    }                                                               \
  } while (0)

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// geometry of a neighbourhood plan (nbplan.cu): offsets in int32 words from the start of the plan buffer
struct NbGeom {
  int K, tm, rt, RS, umax;
  int64_t S, off_order, off_ucount, off_uniq, off_loc, words;
};
NbGeom nb_geometry(int64_t n_out, int K);
int nb_tune(const char* key, int value);
int nb_min_rows();
int nb_disabled();
int nb_target_ctas();
int nb_no_split();

// ---- coordinate keys ---------------------------------------------------------------------------------
// 64-bit key: batch 10 bit | x 18 bit | y 18 bit | z 18 bit (two's complement fields).
constexpr int kCoordBits = 18;
constexpr int32_t kCoordLimit = (1 << (kCoordBits - 1)) - 64;  // margin so +-offset probes never wrap
constexpr int32_t kBatchLimit = 1023;
constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr int kNumHashes = 4;
constexpr int kMaxEvictions = 256;

__host__ __device__ __forceinline__ uint64_t pack_key(int32_t b, int32_t x, int32_t y, int32_t z) {
  const uint64_t m = (1ull << kCoordBits) - 1;
  return (uint64_t(uint32_t(b)) << (3 * kCoordBits)) | ((uint64_t(uint32_t(x)) & m) << (2 * kCoordBits)) |
         ((uint64_t(uint32_t(y)) & m) << kCoordBits) | (uint64_t(uint32_t(z)) & m);
}

__host__ __device__ __forceinline__ bool key_in_range(int32_t b, int32_t x, int32_t y, int32_t z) {
  return b >= 0 && b < kBatchLimit && x > -kCoordLimit && x < kCoordLimit && y > -kCoordLimit &&
         y < kCoordLimit && z > -kCoordLimit && z < kCoordLimit;
}

// murmur3 finaliser over key ^ per-function seed
__host__ __device__ __forceinline__ uint32_t hash_slot(uint64_t key, int j, uint32_t mask) {
  const uint64_t seeds[kNumHashes] = {0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull,
                                      0xD6E8FEB86659FD93ull};
  uint64_t h = key ^ seeds[j];
  h ^= h >> 33;
  h *= 0xff51afd7ed558ccdull;
  h ^= h >> 33;
  h *= 0xc4ceb9fe1a85ec53ull;
  h ^= h >> 33;
  return uint32_t(h) & mask;
}

// Slot holding `key`, or -1.  Probe order j = 0..3; an EMPTY first-choice slot proves absence (slots are never
// emptied once filled and every key's first placement attempt is its j = 0 slot).
__device__ __forceinline__ int32_t cuckoo_find(const uint64_t* __restrict__ keys, uint32_t mask, uint64_t key) {
#pragma unroll
  for (int j = 0; j < kNumHashes; ++j) {
    const uint32_t s = hash_slot(key, j, mask);
    const uint64_t k = __ldg(keys + s);
    if (k == key) return int32_t(s);
    if (j == 0 && k == kEmptyKey) return -1;
  }
  return -1;
}

}  // namespace lgs
