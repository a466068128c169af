// Neighbourhood plan of a same-map 3x3x3 kernel map (the large coordinate maps of the U-Net).
//
// The output-stationary neighbour table says, for output row o and offset k, which input row to gather.  Processing 128
// CONSECUTIVE rows per tile (conv_bx3.cu) gathers every (row, offset) pair separately: 27 x 128 rows x 128 B per channel
// block through cp.async, although rows that are close in space share most of their neighbours.  The plan regroups the
// output rows into spatially compact supertiles (order = rows sorted along a Morton curve over coarse cells), lists the
// UNIQUE input rows each supertile touches (`uniq`, a few hundred instead of 27 x rows: measured 415 on average for 256
// rows of the synthetic ScanNet-shaped scene, i.e. 9x fewer gathered rows) and rewrites the table in supertile-local
// indices (`loc`, 16 bit).  conv_nb.cu loads a supertile's unique rows into shared memory ONCE per channel block and
// builds every offset's A operand from that cache.  Row order of the tensors does not change: only the grouping of output
// rows into CTAs does, so results per row are bit-identical to any other grouping.
//
// Replaces nothing in MinkowskiEngine (its kernel maps are per-offset pair lists consumed by gather -> GEMM -> scatter,
// ME src/convolution_kernel.cu [ME-upstream]); this is the B200-native formulation of the same map.
//
// Integer work, one pass over the rows per kernel, no sort:
//   1. nb_extent_kernel   min / max of (b, x, y, z)                                  (atomicMin / atomicMax)
//   2. nb_bin_kernel      row -> bin = batch bits | Morton(cell), cell = (c - min) >> shift; histogram
//   3. nb_scan_*_kernel   exclusive scan of the 2^18 bin counts (block totals, then blocks scan with their offsets)
//   4. nb_scatter_kernel  order_tmp[bin_start + cursor++] = row;  nb_rank_kernel: rows of a bin sorted by row id, so the
//                         order (the composition of every supertile) is deterministic
//   5. nb_plan_kernel     one CTA per supertile: slots re-ordered by COLOUR (below), shared-memory hash set of the
//                         neighbour rows -> ids by slot order (block scan) -> uniq[], loc[][].  The SET of unique rows is
//                         deterministic, their local numbering is not (linear-probing slots depend on arrival order);
//                         nothing downstream depends on it
//
// Colours (bank-conflict-free cache reads without a register permutation).  conv_nb.cu's row threads read 16-byte chunks of
// ARBITRARY cached rows; eight lanes that read the same chunk index of eight rows with a 128-byte pitch would collide in
// one bank group.  Every voxel gets a colour g = (x + 3y + 5z) / step mod 8, a cached row stores logical chunk c at position
// (c + g) mod 8, and the slots of a supertile are ordered so that each aligned group of eight slots (a quarter-warp) holds
// eight DIFFERENT colours wherever the supertile's colour histogram allows it (rows beyond a colour's share fill the
// holes and cost a 2-way conflict).  Translation by a kernel offset adds the same constant to all eight colours, so the
// eight neighbour rows of a quarter-warp are again pairwise different in colour for every offset: lane j reads position
// (i + g_j) mod 8 in step i and the eight reads fall into eight different bank groups.  Missing neighbours read the zero
// row at the position their colour WOULD have.  loc = idx | g << 12 (idx 0xFFF = no neighbour), uniq = row | g << 28.
#include <algorithm>
#include <string>

#include "common.cuh"

namespace lgs {
namespace nbp {

constexpr int kBinBits = 18;
constexpr int kBins = 1 << kBinBits;
constexpr int kHash = 4096;             // shared-memory hash slots per supertile (load <= 0.25 at UMAX 1024)
constexpr uint32_t kMagic = 0x4E42504Cu;

__global__ void nb_init_kernel(int32_t* ext, int32_t* bins) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4) ext[i] = INT32_MAX;
  else if (i < 8) ext[i] = INT32_MIN;
  for (int j = i; j < 2 * kBins; j += gridDim.x * blockDim.x) bins[j] = 0;
}

__global__ void __launch_bounds__(256) nb_extent_kernel(const int32_t* __restrict__ coords, int64_t n, int32_t* ext) {
  int4 lo = make_int4(INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX), hi = make_int4(INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int4 c = __ldg(reinterpret_cast<const int4*>(coords) + i);
    lo.x = min(lo.x, c.x), lo.y = min(lo.y, c.y), lo.z = min(lo.z, c.z), lo.w = min(lo.w, c.w);
    hi.x = max(hi.x, c.x), hi.y = max(hi.y, c.y), hi.z = max(hi.z, c.z), hi.w = max(hi.w, c.w);
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    lo.x = min(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, d)), lo.y = min(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, d));
    lo.z = min(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, d)), lo.w = min(lo.w, __shfl_xor_sync(0xffffffffu, lo.w, d));
    hi.x = max(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, d)), hi.y = max(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, d));
    hi.z = max(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, d)), hi.w = max(hi.w, __shfl_xor_sync(0xffffffffu, hi.w, d));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(ext + 0, lo.x), atomicMin(ext + 1, lo.y), atomicMin(ext + 2, lo.z), atomicMin(ext + 3, lo.w);
    atomicMax(ext + 4, hi.x), atomicMax(ext + 5, hi.y), atomicMax(ext + 6, hi.z), atomicMax(ext + 7, hi.w);
  }
}

__device__ __forceinline__ int bits_for(uint32_t extent) {   // smallest b with extent < 2^b
  return extent == 0 ? 0 : 32 - __clz(extent);
}
__device__ __forceinline__ uint32_t spread3(uint32_t v) {    // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x30000ffu;
  v = (v | (v << 8)) & 0x300f00fu;
  v = (v | (v << 4)) & 0x30c30c3u;
  v = (v | (v << 2)) & 0x9249249u;
  return v;
}

// bin of a row: [batch bits | Morton of the cell].  The cell size (a power of two, in coordinate units) is the smallest for
// which batch bits + 3 x cell bits fit kBinBits; with one row per bin the order is the exact Morton order.
__device__ __forceinline__ uint32_t bin_of(int4 c, const int32_t* __restrict__ ext) {
  const int bb = bits_for(uint32_t(ext[4] - ext[0]));
  const int per_axis = (kBinBits - bb) / 3;
  const uint32_t ex = uint32_t(ext[5] - ext[1]), ey = uint32_t(ext[6] - ext[2]), ez = uint32_t(ext[7] - ext[3]);
  const int need = max(bits_for(ex), max(bits_for(ey), bits_for(ez)));
  const int shift = max(0, need - per_axis);
  const uint32_t x = uint32_t(c.y - ext[1]) >> shift, y = uint32_t(c.z - ext[2]) >> shift, z = uint32_t(c.w - ext[3]) >> shift;
  const uint32_t m = spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
  return ((uint32_t(c.x - ext[0]) << (3 * per_axis)) | m) & uint32_t(kBins - 1);
}

__global__ void __launch_bounds__(256) nb_bin_kernel(const int32_t* __restrict__ coords, int64_t n, const int32_t* __restrict__ ext,
                                                     int32_t* __restrict__ bin_of_row, int32_t* bins) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint32_t b = bin_of(__ldg(reinterpret_cast<const int4*>(coords) + i), ext);
  bin_of_row[i] = int32_t(b);
  atomicAdd(bins + b, 1);
}

// exclusive scan of the kBins counts in place, two launches of kBins / 1024 blocks x 1024 threads (one bin per thread):
// block totals, then every block scans the totals of the blocks before it (256 values) and its own 1024 bins
constexpr int kScanBlocks = kBins / 1024;
__device__ __forceinline__ int32_t block_inclusive_scan_1024(int32_t v, int32_t* warp_sums) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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
  return inc + (warp ? warp_sums[warp - 1] : 0);
}
__global__ void __launch_bounds__(1024) nb_scan_totals_kernel(const int32_t* __restrict__ bins, int32_t* __restrict__ totals) {
  __shared__ int32_t warp_sums[32];
  const int32_t inc = block_inclusive_scan_1024(bins[blockIdx.x * 1024 + threadIdx.x], warp_sums);
  if (threadIdx.x == 1023) totals[blockIdx.x] = inc;
}
__global__ void __launch_bounds__(1024) nb_scan_kernel(int32_t* bins, const int32_t* __restrict__ totals) {
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t base;
  int32_t before = 0;
  for (int i = threadIdx.x; i < int(blockIdx.x); i += 1024) before += totals[i];     // kScanBlocks <= 1024: one value per thread
  const int32_t all = block_inclusive_scan_1024(before, warp_sums);
  if (threadIdx.x == 1023) base = all;
  __syncthreads();
  const int32_t off = base;
  __syncthreads();
  const int32_t v = bins[blockIdx.x * 1024 + threadIdx.x];
  const int32_t inc = block_inclusive_scan_1024(v, warp_sums);
  bins[blockIdx.x * 1024 + threadIdx.x] = off + inc - v;
}

__global__ void __launch_bounds__(256) nb_scatter_kernel(const int32_t* __restrict__ bin_of_row, int64_t n,
                                                         const int32_t* __restrict__ bin_start, int32_t* cursor,
                                                         int32_t* __restrict__ order_tmp) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int32_t b = bin_of_row[i];
  order_tmp[bin_start[b] + atomicAdd(cursor + b, 1)] = int32_t(i);
}

// rows of one bin in ascending row id (the scatter's arrival order is arbitrary): rank = rows of the bin with a lower id.
// Bins hold a few dozen rows; a bin larger than kRankCap keeps its arrival order (still a valid plan).
constexpr int kRankCap = 2048;
__global__ void __launch_bounds__(256) nb_rank_kernel(const int32_t* __restrict__ bin_of_row, int64_t n, int64_t n_pad,
                                                      const int32_t* __restrict__ bin_start, const int32_t* __restrict__ cursor,
                                                      const int32_t* __restrict__ order_tmp, int32_t* __restrict__ order) {
  const int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (p >= n_pad) return;
  if (p >= n) {
    order[p] = -1;
    return;
  }
  const int32_t row = order_tmp[p];
  const int32_t b = bin_of_row[row];
  const int32_t s = bin_start[b], cnt = cursor[b];
  int32_t rank = int32_t(p) - s;
  if (cnt <= kRankCap) {
    rank = 0;
    for (int32_t j = 0; j < cnt; ++j) rank += order_tmp[s + j] < row ? 1 : 0;
  }
  order[s + rank] = row;
}

__device__ __forceinline__ uint32_t hash_row(int32_t j) {
  uint32_t h = uint32_t(j) * 0x9E3779B1u;
  return (h ^ (h >> 15)) & uint32_t(kHash - 1);
}

// One CTA (256 threads) per supertile of RS = tm * rt row slots (slot q = t * rt + i; slots past the map hold order = -1).
__device__ __forceinline__ int colour_of(int4 c, int step) {
  return ((c.y / step) + 3 * (c.z / step) + 5 * (c.w / step)) & 7;     // coordinates are exact multiples of the tensor stride
}

__global__ void __launch_bounds__(256) nb_plan_kernel(const int32_t* __restrict__ coords, int step,
                                                      const int32_t* __restrict__ table, int K, int64_t n_out,
                                                      int32_t* __restrict__ order, int RS, int umax,
                                                      int32_t* __restrict__ ucount, int32_t* __restrict__ uniq,
                                                      uint16_t* __restrict__ loc, int32_t* hdr) {
  __shared__ int32_t keys[kHash];
  __shared__ uint16_t ids[kHash];
  __shared__ uint8_t gcol[kHash];
  __shared__ int32_t warp_sums[8];
  __shared__ int32_t total;
  __shared__ int32_t ord[256];
  __shared__ int8_t col_in[256], own_col[256];
  __shared__ uint8_t taken[256], ovf[256];
  const int s = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < kHash; i += 256) keys[i] = -1;
  // phase 0: slots of the supertile re-ordered by colour: the r-th row of colour c goes to slot 8 r + c while octets last;
  // the rest (and the padding of the last supertile) fill the free slots in order.  Deterministic.
  {
    int32_t o = -1;
    int c = -1;
    if (tid < RS) {
      o = order[int64_t(s) * RS + tid];
      if (o >= 0) c = colour_of(__ldg(reinterpret_cast<const int4*>(coords) + o), step);
    }
    col_in[tid] = int8_t(c);
    taken[tid] = 0;
    ord[tid] = -1;
    own_col[tid] = int8_t(tid & 7);
    __syncthreads();
    const int octets = RS >> 3;
    int rank = 0;
    if (c >= 0)
      for (int q = 0; q < tid; ++q) rank += col_in[q] == c ? 1 : 0;
    const bool primary = c >= 0 && rank < octets;
    if (primary) {
      const int slot = 8 * rank + c;
      ord[slot] = o;
      own_col[slot] = int8_t(c);
      taken[slot] = 1;
    }
    ovf[tid] = (c >= 0 && !primary) ? 1 : 0;
    __syncthreads();
    if (c >= 0 && !primary) {
      int before = 0;                                   // overflow rows ahead of this one
      for (int q = 0; q < tid; ++q) before += ovf[q];
      int slot = 0;
      for (int seen = -1; slot < RS; ++slot) {
        if (!taken[slot] && ++seen == before) break;
      }
      ord[slot] = o;                                    // distinct `before` -> distinct free slots: no race
      own_col[slot] = int8_t(c);
    }
    __syncthreads();
    if (tid < RS) order[int64_t(s) * RS + tid] = ord[tid];
  }
  __syncthreads();
  const int work = K * RS;
  // phase 1: insert every neighbour row into the hash set
  for (int w = tid; w < work; w += 256) {
    const int k = w / RS, q = w - k * RS;
    const int32_t o = ord[q];
    if (o < 0) continue;
    const int32_t j = __ldg(table + int64_t(k) * n_out + o);
    if (j < 0) continue;
    uint32_t h = hash_row(j);
    for (int probe = 0; probe < kHash; ++probe) {
      const int32_t prev = atomicCAS(keys + h, -1, j);
      if (prev == -1 || prev == j) break;
      h = (h + 1) & uint32_t(kHash - 1);
    }
  }
  __syncthreads();
  // phase 2: ids in slot order: block exclusive scan of the occupancy, 16 slots per thread
  int32_t occ = 0;
  const int base = tid * (kHash / 256);
#pragma unroll
  for (int i = 0; i < kHash / 256; ++i) occ += keys[base + i] >= 0 ? 1 : 0;
  const int lane = tid & 31, warp = tid >> 5;
  int32_t inc = occ;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int32_t y = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += y;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (tid == 0) {
    int32_t run = 0;
    for (int w = 0; w < 8; ++w) {
      const int32_t v = warp_sums[w];
      warp_sums[w] = run;
      run += v;
    }
    total = run;
  }
  __syncthreads();
  int32_t id = (inc - occ) + warp_sums[warp];
#pragma unroll
  for (int i = 0; i < kHash / 256; ++i) {
    const int32_t key = keys[base + i];
    if (key >= 0) {
      const int g = colour_of(__ldg(reinterpret_cast<const int4*>(coords) + key), step);
      ids[base + i] = uint16_t(min(id, 0xFFFF));
      gcol[base + i] = uint8_t(g);
      if (id < umax) uniq[int64_t(s) * umax + id] = key | (g << 28);
      ++id;
    }
  }
  __syncthreads();
  if (tid == 0) {
    ucount[s] = min(total, umax);
    if (total > umax) atomicOr(hdr + 8, 1);
    atomicMax(hdr + 9, total);
  }
  // phase 3: the table in local indices: idx | colour << 12; idx 0xFFF = no neighbour / padded slot / did not fit, with the
  // colour the neighbour would have (own colour + the offset's colour shift), so that the zero-row read is conflict-free too
  uint16_t* lc = loc + int64_t(s) * K * RS;
  for (int w = tid; w < work; w += 256) {
    const int k = w / RS, q = w - k * RS;
    const int32_t o = ord[q];
    const int dk = (k % 3 - 1) + 3 * ((k / 3) % 3 - 1) + 5 * (k / 9 - 1);                  // x fastest, as in kmap_kernel
    uint16_t v = uint16_t(0xFFF | (((int(own_col[q]) + dk) & 7) << 12));
    if (o >= 0) {
      const int32_t j = __ldg(table + int64_t(k) * n_out + o);
      if (j >= 0) {
        uint32_t h = hash_row(j);
        int probe = 0;
        while (keys[h] != j && probe < kHash) h = (h + 1) & uint32_t(kHash - 1), ++probe;   // bounded: a full set drops rows
        if (probe < kHash) {
          const uint16_t lid = ids[h];
          if (int(lid) < umax) v = uint16_t(lid | (uint16_t(gcol[h]) << 12));
        }
      }
    }
    lc[w] = v;
  }
}

}  // namespace nbp

// ---- plan geometry: a pure function of (n_out, K) and the knobs, so the builder and the kernels agree without a header ----
static int g_nb_rt = 0, g_nb_umax = 0, g_nb_min_rows = 0, g_nb_off = 0, g_nb_target = 0, g_nb_no_split = 0;
int nb_tune(const char* key, int value) {
  const std::string s(key);
  if (s == "nb_rt") g_nb_rt = value;
  else if (s == "nb_umax") g_nb_umax = value;
  else if (s == "nb_min_rows") g_nb_min_rows = value;
  else if (s == "nb_off") g_nb_off = value;
  else if (s == "nb_target_ctas") g_nb_target = value;
  else if (s == "nb_no_split") g_nb_no_split = value;
  else return 0;
  return 1;
}

NbGeom nb_geometry(int64_t n_out, int K) {
  NbGeom g;
  g.K = K;
  g.tm = 2;
  g.umax = std::min(4000, g_nb_umax > 0 ? g_nb_umax : 640);
  // rows per tile.  Large maps: whole waves of 296 CTAs (2 resident per SM), tiles as full as the wave count allows.
  // Small maps (fewer than 148 x 128 rows, where the BatchNorm statistics are not fused either): full 128-row tiles — conv_nb.cu finds its parallelism by splitting the
  // reduction over CTAs instead of by emptying MMA lanes.
  const int64_t slots = 296;
  const int64_t full = cdiv(n_out, int64_t(g.tm) * 128);
  const int64_t waves = std::max<int64_t>(1, cdiv(full, slots));
  int rt = int(cdiv(cdiv(n_out, waves * slots), int64_t(g.tm)));
  rt = std::min(128, std::max(32, rt));
  if (n_out < 148 * 128) rt = int(std::min<int64_t>(128, std::max<int64_t>(8, cdiv(n_out, int64_t(g.tm)))));
  if (g_nb_rt > 0) rt = std::min(128, std::max(8, g_nb_rt));
  rt = std::min(128, (rt + 7) & ~7);                // whole octets of slots per tile (colour groups, see the header comment)
  g.rt = rt;
  g.RS = g.tm * rt;
  g.S = cdiv(n_out, int64_t(g.RS));
  g.off_order = 16;
  g.off_ucount = g.off_order + g.S * g.RS;
  g.off_uniq = g.off_ucount + ((g.S + 3) & ~int64_t(3));
  g.off_loc = g.off_uniq + g.S * int64_t(g.umax);
  g.words = g.off_loc + (g.S * int64_t(K) * g.RS + 1) / 2;
  g.words = (g.words + 3) & ~int64_t(3);
  return g;
}

int nb_min_rows() { return g_nb_min_rows > 0 ? g_nb_min_rows : 256; }
int nb_target_ctas() { return g_nb_target > 0 ? g_nb_target : 296; }
int nb_no_split() { return g_nb_no_split; }
int nb_disabled() { return g_nb_off; }

}  // namespace lgs

using namespace lgs;

extern "C" {

int64_t lgs_nbplan_bytes(int64_t n_out, int32_t K) {
  if (n_out <= 0 || K < 1) return 0;
  return nb_geometry(n_out, K).words * 4;
}

int64_t lgs_nbplan_scratch_bytes(int64_t n_out) {
  if (n_out <= 0) return 0;
  const int64_t n_pad = n_out + 2 * 128 * 4;       // order_tmp / bin_of_row
  return (8 + 1024 + 2 * int64_t(nbp::kBins) + 2 * n_pad) * 4;
}

int lgs_nbplan_supported(int64_t n_out, int32_t K) {
  return (!nb_disabled() && K == 27 && n_out >= nb_min_rows() && n_out < (int64_t(1) << 28)) ? 1 : 0;
}

int lgs_nbplan_build(const int32_t* d_out_coords, int64_t n_out, const int32_t* d_table, int32_t K, int32_t step, void* d_plan,
                     void* d_scratch, int32_t* h_status, void* stream_) {
  LGS_TRACE("lgs_nbplan_build %p %lld %p %d %d %p %p %p", (const void*)d_out_coords, (long long)n_out, (const void*)d_table, (int)K, (int)step, (const void*)d_plan, (const void*)d_scratch, (const void*)stream_);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d_out_coords || !d_table || !d_plan || !d_scratch || n_out <= 0 || K != 27 || step < 1)
    return fail(LGS_E_INVALID, "lgs_nbplan_build: bad arguments");
  if (n_out >= (int64_t(1) << 28)) return fail(LGS_E_UNSUPPORTED, "lgs_nbplan_build: map too large");
  const NbGeom g = nb_geometry(n_out, K);
  int32_t* plan = static_cast<int32_t*>(d_plan);
  int32_t* scr = static_cast<int32_t*>(d_scratch);
  int32_t* ext = scr;
  int32_t* totals = scr + 8;
  int32_t* bins = scr + 8 + 1024;
  int32_t* cursor = bins + nbp::kBins;
  int32_t* bin_of_row = cursor + nbp::kBins;
  int32_t* order_tmp = bin_of_row + (n_out + 1024);
  const int64_t n_pad = g.S * g.RS;
  const int32_t hdr[16] = {int32_t(nbp::kMagic), K, int32_t(n_out), g.tm, g.rt, g.RS, int32_t(g.S), g.umax, 0, 0, 0, 0, 0, 0, 0, 0};
  LGS_CUDA(cudaMemcpyAsync(plan, hdr, sizeof(hdr), cudaMemcpyHostToDevice, stream));
  LGS_LAUNCH(nbp::nb_init_kernel, 512, 1024, 0, stream, ext, bins);
  const unsigned rb = unsigned(cdiv(n_out, 256));
  LGS_LAUNCH(nbp::nb_extent_kernel, std::min(rb, 148u * 8u), 256, 0, stream, d_out_coords, n_out, ext);
  LGS_LAUNCH(nbp::nb_bin_kernel, rb, 256, 0, stream, d_out_coords, n_out, ext, bin_of_row, bins);
  LGS_LAUNCH(nbp::nb_scan_totals_kernel, nbp::kScanBlocks, 1024, 0, stream, bins, totals);
  LGS_LAUNCH(nbp::nb_scan_kernel, nbp::kScanBlocks, 1024, 0, stream, bins, totals);
  LGS_LAUNCH(nbp::nb_scatter_kernel, rb, 256, 0, stream, bin_of_row, n_out, bins, cursor, order_tmp);
  LGS_LAUNCH(nbp::nb_rank_kernel, unsigned(cdiv(n_pad, 256)), 256, 0, stream, bin_of_row, n_out, n_pad, bins, cursor, order_tmp,
             plan + g.off_order);
  LGS_LAUNCH(nbp::nb_plan_kernel, unsigned(g.S), 256, 0, stream, d_out_coords, step, d_table, K, n_out, plan + g.off_order, g.RS, g.umax,
             plan + g.off_ucount, plan + g.off_uniq, reinterpret_cast<uint16_t*>(plan + g.off_loc), plan);
  if (h_status) {
    int32_t st[2] = {0, 0};
    LGS_CUDA(cudaMemcpyAsync(st, plan + 8, sizeof(st), cudaMemcpyDeviceToHost, stream));
    LGS_CUDA(cudaStreamSynchronize(stream));
    h_status[0] = st[0], h_status[1] = st[1];
  }
  return LGS_OK;
}

/* geometry of the plan lgs_nbplan_build writes for (n_out, K): out[0..7] = tm, rt, RS, S, umax, order / ucount / uniq offsets
 * in int32 words; out[8] = loc offset (uint16 array starts at that word); for tests and tools */
int lgs_nbplan_geometry(int64_t n_out, int32_t K, int64_t* out) {
  if (!out || n_out <= 0) return fail(LGS_E_INVALID, "lgs_nbplan_geometry: bad arguments");
  const NbGeom g = nb_geometry(n_out, K);
  out[0] = g.tm, out[1] = g.rt, out[2] = g.RS, out[3] = g.S, out[4] = g.umax, out[5] = g.off_order, out[6] = g.off_ucount,
  out[7] = g.off_uniq, out[8] = g.off_loc;
  return LGS_OK;
}

}  // extern "C"
