// LGS_ALGO_TC: output-stationary sparse convolution on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// One CTA owns a tile of 128 output rows and up to 256 output channels.  For every kernel offset k that has at least
// one neighbour in the tile, and every 128-byte block of input channels:
//   * 4 producer warps gather the neighbour rows  in[table[k][o]]  with 16-byte cp.async (zero-fill for missing
//     neighbours) straight into a 128B-swizzled K-major shared-memory tile  A[128 rows][128 B];
//   * one thread TMA-loads the matching weight block  W^T[k][n][128 B]  (K-major, 128B swizzle) — the B operand;
//   * one thread issues tcgen05.mma (M=128, N=c_out, K=32 B per instruction) accumulating into TMEM over ALL offsets
//     and channel blocks, so the [128 x c_out] fp32 accumulator never leaves the SM until the tile is done;
//   * the producer warps then drain TMEM (tcgen05.ld) and store each output row once — no atomics, deterministic.
// The stages form an mbarrier ring (full: 128 cp.async-completion arrivals + TMA transaction bytes; empty:
// tcgen05.commit).
//
// Math modes: fp32 features -> kind::tf32 (operands are read as TF32, fp32 accumulate); bf16 features -> kind::f16.
// Shapes outside the envelope (row bytes not a multiple of 16, e.g. c_in = 3) return LGS_E_UNSUPPORTED and the caller
// falls back to the exact SIMT kernel.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "tcgen05.cuh"

namespace lgs {

namespace tc {

constexpr int BM = 128;          // output rows per CTA (UMMA M)
constexpr int THREADS = 192;     // warps 0-3: gather producers + epilogue; warp 4: MMA issuer; warp 5: weight TMA
constexpr int KBLOCK_BYTES = 128;
constexpr int A_STAGE_BYTES = BM * KBLOCK_BYTES;  // 16 KB
constexpr int MIN_STAGES = 3;
constexpr int MAX_STAGES = 8;


struct Params {
  const uint8_t* in;        // [n_in, c_in] features (fp32 or bf16)
  int64_t n_in;
  int32_t row_bytes;        // c_in * elem size
  int32_t K;
  int32_t c_out;            // total output channels
  const int32_t* table;     // [K, n_out] or nullptr (identity, K == 1)
  int64_t n_out;
  int32_t reverse_k;
  const float* bias;
  void* out;                // [n_out, c_out]
  int32_t num_kb;           // 128-byte channel blocks per offset
  int32_t n_tile;           // MMA N of one CTA (multiple of 16, <= 256)
  int32_t stages;
  int32_t tmem_cols;
  int32_t b_stage_bytes;    // n_tile * 128
  int32_t k_per_split;      // offsets handled by one CTA (blockIdx.z selects the range); < K => partial sums, red.add
  int32_t k_splits;
};

// PRECISE (fp32 features only): 3xTF32 error-compensated product.  Four extra warps derive from every landed A tile the
// residual lo = a - trunc_tf32(a) (second tile); the tensor core itself truncates the raw tile to hi (kind::tf32 ignores the
// low 13 mantissa bits of both operands — verified bit-exactly by scripts/dev_trunc.py); the weights arrive pre-split (hi, lo
// stacked, see lgs_weight_prep); three MMAs per K step accumulate hi*hi + lo*hi + hi*lo, i.e. fp32-grade products
// (relative error ~2^-21) with fp32 accumulation in TMEM.
//
// The hot loops are written for ONE warp per SM sub-partition (no latency hiding by other warps): everything that does
// not change per stage is hoisted — row pointers once per offset, destination offsets once per thread, descriptor
// high words once per kernel, stage index / phase carried incrementally (no division).

template <bool BF16, bool PRECISE>
__global__ void __launch_bounds__(PRECISE ? THREADS + 128 : THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x ([A hi | A lo] [B hi | B lo])] [idx K*128 int32] [klist 32] [kflag 32] [barriers] [tmem ptr]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NSPLIT = PRECISE ? 2 : 1;
  constexpr int NTHREADS = PRECISE ? THREADS + 128 : THREADS;
  const int stages = p.stages;
  const uint32_t stage_bytes = NSPLIT * (A_STAGE_BYTES + p.b_stage_bytes);
  int32_t* sidx = reinterpret_cast<int32_t*>(smem + size_t(stages) * stage_bytes);
  int32_t* klist = sidx + p.k_per_split * BM;
  int32_t* kflag = klist + 32;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(kflag + 32);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* split_bar = empty_bar + MAX_STAGES;
  uint64_t* acc_bar = split_bar + MAX_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_bar + 1);
  int32_t* nk_smem = reinterpret_cast<int32_t*>(tmem_ptr_smem + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = int64_t(blockIdx.x) * BM;
  const int n0 = blockIdx.y * p.n_tile;
  // split-K over kernel offsets for small coordinate maps: this CTA owns offsets [kbeg, kbeg + kcnt)
  const int kbeg = blockIdx.z * p.k_per_split;
  const int kcnt = min(p.k_per_split, p.K - kbeg);
  const bool partial = p.k_splits > 1;
  const uint32_t smem_base = smem_u32(smem);

  // ---- prologue: neighbour table of this tile -> shared memory (async, all offsets in flight at once) ----
  {
    const int64_t rows_valid = (p.n_out - m0) < int64_t(BM) ? (p.n_out - m0) : int64_t(BM);
    for (int e = tid; e < kcnt * BM; e += NTHREADS) {
      const int k = kbeg + (e >> 7), r = e & (BM - 1);
      if (r < rows_valid && p.table) {
        cp_async4(smem_u32(sidx + e), p.table + int64_t(p.reverse_k ? p.K - 1 - k : k) * p.n_out + m0 + r);
      } else {
        sidx[e] = r < rows_valid ? int32_t(m0 + r) : -1;
      }
    }
  }
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar + s, 128 + 1);
      mbar_init(empty_bar + s, 1);
      mbar_init(split_bar + s, 128);
    }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(uint32_t(p.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 5 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
  cp_async_wait_all();
  __syncthreads();
  // which offsets have any neighbour in this tile
  for (int k = warp; k < kcnt; k += NTHREADS / 32) {
    bool any = false;
#pragma unroll
    for (int j = 0; j < BM / 32; ++j) any |= sidx[k * BM + lane + 32 * j] >= 0;
    const uint32_t b = __ballot_sync(0xffffffffu, any);
    if (lane == 0) kflag[k] = b != 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    int nk = 0;
    for (int k = 0; k < kcnt; ++k)
      if (kflag[k]) klist[nk++] = k;   // local offset index; the weight block is kbeg + k
    *nk_smem = nk;
  }
  __syncthreads();
  const int nk = *nk_smem;
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int num_kb = p.num_kb;
  const int total = nk * num_kb;

  if (warp < 4) {
    // =================================== gather producers ===================================
    const int chunk = tid & 7, rbase = tid >> 3;  // 8 lanes cover one 128-byte row segment; 16 rows per pass
    uint32_t dst_off[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = rbase + 16 * i;
      dst_off[i] = uint32_t(r * KBLOCK_BYTES + ((chunk ^ (r & 7)) << 4));
    }
    const int row_bytes = p.row_bytes;
    const uint8_t* in_chunk = p.in + chunk * 16;
    int s = 0;
    uint32_t ph = 0;
    uint32_t a_base = smem_base;
    for (int a = 0; a < nk; ++a) {
      const int32_t* idx = sidx + klist[a] * BM + rbase;
      const uint8_t* src[8];
      uint32_t okmask = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int32_t row = idx[16 * i];
        okmask |= (row >= 0 ? 1u : 0u) << i;
        src[i] = in_chunk + size_t(row >= 0 ? row : 0) * row_bytes;
      }
      int col = chunk * 16;
      for (int kb = 0; kb < num_kb; ++kb, col += KBLOCK_BYTES) {
        mbar_wait(empty_bar + s, ph ^ 1);
        const uint32_t m = col < row_bytes ? okmask : 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) cp_async16(a_base + dst_off[i], src[i] + kb * KBLOCK_BYTES, ((m >> i) & 1u) ? 16u : 0u);
        // hardware arrives on full[s] when this thread's copies have landed: nothing blocks, every free stage of the
        // ring is in flight.  The generic->async proxy fence is issued by the consumer after it observes the barrier.
        cp_async_mbar_arrive_noinc(full_bar + s);
        a_base += stage_bytes;
        if (++s == stages) {
          s = 0;
          ph ^= 1;
          a_base = smem_base;
        }
      }
    }

    // =================================== epilogue ===================================
    if (total > 0) {
      mbar_wait(acc_bar, 0);
      tc_fence_after();
    }
    const int64_t o = m0 + warp * 32 + lane;
    const int ncols = min(p.n_tile, p.c_out - n0);
    const float* bias = (partial && blockIdx.z != 0) ? nullptr : p.bias;
    for (int c0 = 0; c0 < ((partial && total == 0) ? 0 : ncols); c0 += 32) {
      uint32_t v[32];
      if (total > 0) {
        tmem_ld32(tmem_base + (uint32_t(warp * 32) << 16) + uint32_t(c0), v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (o < p.n_out) {
        if constexpr (BF16) {
          __nv_bfloat16* orow = static_cast<__nv_bfloat16*>(p.out) + size_t(o) * p.c_out + n0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            if (c0 + j + 1 < ncols) {
              const float x0 = __uint_as_float(v[j]) + (bias ? __ldg(bias + n0 + c0 + j) : 0.f);
              const float x1 = __uint_as_float(v[j + 1]) + (bias ? __ldg(bias + n0 + c0 + j + 1) : 0.f);
              *reinterpret_cast<__nv_bfloat162*>(orow + j) = __floats2bfloat162_rn(x0, x1);
            } else if (c0 + j < ncols) {
              orow[j] = __float2bfloat16(__uint_as_float(v[j]) + (bias ? __ldg(bias + n0 + c0 + j) : 0.f));
            }
          }
        } else {
          float* orow = static_cast<float*>(p.out) + size_t(o) * p.c_out + n0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (c0 + j + 3 < ncols) {
              float4 x;
              x.x = __uint_as_float(v[j]) + (bias ? __ldg(bias + n0 + c0 + j) : 0.f);
              x.y = __uint_as_float(v[j + 1]) + (bias ? __ldg(bias + n0 + c0 + j + 1) : 0.f);
              x.z = __uint_as_float(v[j + 2]) + (bias ? __ldg(bias + n0 + c0 + j + 2) : 0.f);
              x.w = __uint_as_float(v[j + 3]) + (bias ? __ldg(bias + n0 + c0 + j + 3) : 0.f);
              if (partial) red_add_v4(orow + j, x.x, x.y, x.z, x.w);
              else *reinterpret_cast<float4*>(orow + j) = x;
            } else {
              for (int jj = j; jj < j + 4; ++jj)
                if (c0 + jj < ncols) {
                  const float x1 = __uint_as_float(v[jj]) + (bias ? __ldg(bias + n0 + c0 + jj) : 0.f);
                  if (partial) atomicAdd(orow + jj, x1);
                  else orow[jj] = x1;
                }
            }
          }
        }
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // =================================== MMA issuer (warp-uniform loop, one elected lane issues) ================
    {
      // instruction descriptor: D fp32, A/B tf32 (or bf16), both K-major, N = n_tile, M = 128
      const uint32_t fmt = BF16 ? 1u : 2u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(p.n_tile >> 3) << 17) | (uint32_t(BM >> 4) << 24);
      // shared-memory descriptor: hi word constant (SBO 1024 B, version 1, SWIZZLE_128B); lo word = addr>>4 | LBO(1)<<16
      const uint32_t desc_hi = uint32_t(1024 >> 4) | (1u << 14) | (2u << 29);
      const uint32_t b_off = NSPLIT * A_STAGE_BYTES;
      const int row_bytes = p.row_bytes;
      int s = 0;
      uint32_t ph = 0, first = 0;
      uint32_t a_base = smem_base;
      for (int a = 0; a < nk; ++a) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar + s, ph);          // gathered rows (cp.async) and weights (TMA) have landed
          if constexpr (PRECISE) mbar_wait(split_bar + s, ph);   // ... and the hi/lo split of A is done
          fence_proxy_async();  // generic-proxy writes (cp.async / splitters) -> visible to the tensor core's reads
          tc_fence_after();
          const int valid = min(KBLOCK_BYTES, row_bytes - kb * KBLOCK_BYTES);
          const int ksteps = (valid + 31) >> 5;  // 32 bytes of K per instruction (8 tf32 / 16 bf16)
          const uint32_t a_lo32 = ((a_base >> 4) & 0x3FFF) | (1u << 16);
          const uint32_t b_lo32 = (((a_base + b_off) >> 4) & 0x3FFF) | (1u << 16);
          if (elect_one()) {
#pragma unroll 4
            for (int j = 0; j < ksteps; ++j) {
              const uint64_t a_hi = desc_from(a_lo32 + 2 * j, desc_hi), b_hi = desc_from(b_lo32 + 2 * j, desc_hi);
              umma<BF16>(tmem_base, a_hi, b_hi, idesc, first | uint32_t(j));
              if constexpr (PRECISE) {
                umma<BF16>(tmem_base, desc_from(a_lo32 + (A_STAGE_BYTES >> 4) + 2 * j, desc_hi), b_hi, idesc, 1u);
                umma<BF16>(tmem_base, a_hi, desc_from(b_lo32 + (p.b_stage_bytes >> 4) + 2 * j, desc_hi), idesc, 1u);
              }
            }
            umma_commit(empty_bar + s);  // frees the stage once these MMAs have read it
          }
          __syncwarp();
          first = 1;
          a_base += stage_bytes;
          if (++s == stages) {
            s = 0;
            ph ^= 1;
            a_base = smem_base;
          }
        }
      }
      if (total > 0 && elect_one()) umma_commit(acc_bar);
    }
    __syncwarp();
  } else if (warp == 5) {
    // =================================== weight TMA producer (warp-uniform loop, one elected lane issues) ======
    {
      const int kelems = KBLOCK_BYTES / (BF16 ? 2 : 4);
      const uint32_t b_off = NSPLIT * A_STAGE_BYTES;
      const uint32_t tx = NSPLIT * p.b_stage_bytes;
      int s = 0;
      uint32_t ph = 0;
      uint32_t b_dst = smem_base + b_off;
      for (int a = 0; a < nk; ++a) {
        const int row = (kbeg + klist[a]) * p.c_out + n0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar + s, ph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(full_bar + s, tx);
            tma_load_2d(b_dst, &tmap_w, full_bar + s, kb * kelems, row);
            if constexpr (PRECISE)   // lo halves are stacked after the K*c_out hi rows
              tma_load_2d(b_dst + p.b_stage_bytes, &tmap_w, full_bar + s, kb * kelems, p.K * p.c_out + row);
          }
          __syncwarp();
          b_dst += stage_bytes;
          if (++s == stages) {
            s = 0;
            ph ^= 1;
            b_dst = smem_base + b_off;
          }
        }
      }
    }
    __syncwarp();
  }
  if constexpr (PRECISE) {
    if (warp >= 6) {
      // =================================== hi/lo splitters (128 threads) ===================================
      const int st = tid - THREADS;
      int s = 0;
      uint32_t ph = 0;
      uint8_t* a_ptr = smem;
      for (int it = 0; it < total; ++it) {
        mbar_wait(full_bar + s, ph);
        float4* a_hi = reinterpret_cast<float4*>(a_ptr) + st;
        float4* a_lo = reinterpret_cast<float4*>(a_ptr + A_STAGE_BYTES) + st;
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = a_hi[128 * i];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(v[i].x) & 0xFFFFE000u);
          h.y = __uint_as_float(__float_as_uint(v[i].y) & 0xFFFFE000u);
          h.z = __uint_as_float(__float_as_uint(v[i].z) & 0xFFFFE000u);
          h.w = __uint_as_float(__float_as_uint(v[i].w) & 0xFFFFE000u);
          l.x = v[i].x - h.x;
          l.y = v[i].y - h.y;
          l.z = v[i].z - h.z;
          l.w = v[i].w - h.w;
          a_lo[128 * i] = l;
        }
        fence_proxy_async();
        mbar_arrive(split_bar + s);
        a_ptr += stage_bytes;
        if (++s == stages) {
          s = 0;
          ph ^= 1;
          a_ptr = smem;
        }
      }
    }
  }

  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(p.tmem_cols))
                 : "memory");
  }
}

// =============================================================================================================
// Multi-tile variant for large coordinate maps: one CTA owns TM (<= 4) consecutive 128-row tiles and keeps TM
// accumulators in TMEM (TM * n_tile <= 512 columns).  Loop order  offset k -> channel block kb -> tile t :
// every weight block W[k][kb] is TMA-loaded ONCE per CTA and feeds TM MMAs groups, cutting the L2->SM weight traffic
// (the dominant stream of the single-tile kernel: 27 * Cin * Cout * 4 B per 128 rows) by TM.  A and B live in separate
// mbarrier rings.  Table entries are read straight from global memory by the producers, one offset ahead.
// PRECISE (3xTF32): the gathered tile only LANDS in shared memory; four splitter warps (thread <-> row = TMEM lane) read
// their row, split it into hi = trunc_tf32(a) and lo = a - hi and write both with tcgen05.st into a TMEM ring after the
// accumulators; the MMAs run in TS form (A from TMEM), so shared memory serves only the cp.async landing, one read by the
// splitters and the weight operand: 74 KB per 16 KB stage instead of 138 KB (SS form with a second `lo` tile).
// =============================================================================================================
struct Params2 {
  const uint8_t* in;
  int32_t row_bytes;
  int32_t K;
  int32_t c_out;
  const int32_t* table;
  int64_t n_out;
  int32_t reverse_k;
  const float* bias;
  void* out;
  int32_t num_kb;
  int32_t n_tile;        // output channels per CTA: padded c_out, or an equal slice of it when c_out > 256 (blockIdx.y)
  int32_t TM;            // tiles per CTA
  int32_t rt;            // output rows per tile (<= 128): rows beyond it are empty MMA lanes, so that TM*rt*gridDim.x can
                         // match n_out on a grid that is a multiple of the SM count (no partial last wave)
  int32_t a_stages, b_stages;
  int32_t tmem_cols;
  int32_t b_stage_bytes;
  int32_t ta_stages;     // PRECISE: TMEM ring of split A operands (64 columns each: hi | lo), after the TM accumulators
  int32_t ta_col0;       // first TMEM column of that ring
};

constexpr int MAX_A_STAGES = 12, MAX_B_STAGES = 4, MAX_TA_STAGES = 4;
constexpr int T2_PROD_WARPS = 8;                       // gather producers (and epilogue)
constexpr int T2_THREADS = (T2_PROD_WARPS + 2) * 32;    // + MMA issuer warp + weight TMA warp

template <bool BF16, bool PRECISE>
__global__ void __launch_bounds__(PRECISE ? T2_THREADS + 128 : T2_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmap_w, const Params2 p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NSPLIT = PRECISE ? 2 : 1;
  constexpr uint32_t A_BYTES = A_STAGE_BYTES;                   // raw gathered tile; PRECISE: its hi/lo split lives in TMEM
  const uint32_t b_bytes = NSPLIT * p.b_stage_bytes;             // [B hi | B lo]
  const int SA = p.a_stages, SB = p.b_stages, TM = p.TM;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + size_t(SA) * A_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(b_ring + size_t(SB) * b_bytes);
  uint64_t* a_empty = a_full + MAX_A_STAGES;
  uint64_t* b_full = a_empty + MAX_A_STAGES;
  uint64_t* b_empty = b_full + MAX_B_STAGES;
  uint64_t* ta_full = b_empty + MAX_B_STAGES;     // PRECISE: split A operand of a stage is in TMEM
  uint64_t* ta_empty = ta_full + MAX_TA_STAGES;   // PRECISE: MMAs reading that TMEM stage have completed
  uint64_t* acc_bar = ta_empty + MAX_TA_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = int64_t(blockIdx.x) * (int64_t(TM) * p.rt);
  const int n0 = blockIdx.y * p.n_tile;          // output-channel slice of this CTA (c_out > 256 is sliced)
  const int K = p.K, num_kb = p.num_kb;

  if (tid == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(a_full + s, T2_PROD_WARPS * 32);   // cp.async-completion arrivals of the producer threads
      mbar_init(a_empty + s, PRECISE ? 128 : 1);   // PRECISE: freed by the 128 splitter threads once they have read it
    }
    for (int s = 0; s < MAX_TA_STAGES; ++s) {
      mbar_init(ta_full + s, 128);
      mbar_init(ta_empty + s, 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(b_full + s, 1);
      mbar_init(b_empty + s, 1);
    }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == T2_PROD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(uint32_t(p.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == T2_PROD_WARPS + 1 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < T2_PROD_WARPS) {
    {
      // =================================== gather producers: 8 warps, cp.async ===================================
      // Two producer warps per SM sub-partition hide each other's issue latency; 8 lanes cover one 128-byte row
      // segment, 32 rows per pass, 4 passes per 128-row tile.  Table entries are read straight from global memory one
      // offset ahead (8 lanes share an address -> one sector, broadcast).
      const int chunk = tid & 7, rbase = tid >> 3;          // rbase 0..31
      uint32_t dst_off[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = rbase + 32 * i;
        dst_off[i] = uint32_t(r * KBLOCK_BYTES + ((chunk ^ (r & 7)) << 4));
      }
      const int row_bytes = p.row_bytes;
      const uint8_t* in_chunk = p.in + chunk * 16;
      const uint32_t a_ring_base = smem_u32(a_ring);
      auto load_idx = [&](int k, int32_t (&v)[4][4]) {
        const int32_t* trow = p.table ? p.table + int64_t(p.reverse_k ? K - 1 - k : k) * p.n_out : nullptr;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = rbase + 32 * i;
            const int64_t o = m0 + int64_t(t) * p.rt + r;
            v[t][i] = (t < TM && r < p.rt && o < p.n_out) ? (trow ? __ldg(trow + o) : int32_t(o)) : -1;
          }
        }
      };
      int32_t cur[4][4], nxt[4][4];
      load_idx(0, cur);
      int s = 0;
      uint32_t ph = 0;
      for (int k = 0; k < K; ++k) {
        if (k + 1 < K) load_idx(k + 1, nxt);        // in flight while this offset is gathered
        for (int kb = 0; kb < num_kb; ++kb) {
          const int col = kb * KBLOCK_BYTES + chunk * 16;
          const bool col_ok = col < row_bytes;
          const uint8_t* in_kb = in_chunk + kb * KBLOCK_BYTES;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (t < TM) {
              mbar_wait(a_empty + s, ph ^ 1);
              const uint32_t a_base = a_ring_base + uint32_t(s) * A_BYTES;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int32_t row = cur[t][i];
                const bool ok = col_ok && row >= 0;
                cp_async16(a_base + dst_off[i], in_kb + size_t(ok ? row : 0) * row_bytes, ok ? 16u : 0u);
              }
              cp_async_mbar_arrive_noinc(a_full + s);
              if (++s == SA) {
                s = 0;
                ph ^= 1;
              }
            }
          }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
          for (int i = 0; i < 4; ++i) cur[t][i] = nxt[t][i];
      }
    }

    // =================================== epilogue ===================================
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int ncols = min(p.n_tile, p.c_out - n0);
    const int lq = warp & 3;                       // TMEM lane quadrant this warp may read
    for (int t = warp >> 2; t < TM; t += T2_PROD_WARPS / 4) {
      const int64_t o = m0 + int64_t(t) * p.rt + lq * 32 + lane;
      const bool row_ok = lq * 32 + lane < p.rt && o < p.n_out;
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (uint32_t(lq * 32) << 16) + uint32_t(t * p.n_tile + c0), v);
        if (row_ok) {
          if constexpr (BF16) {
            __nv_bfloat16* orow = static_cast<__nv_bfloat16*>(p.out) + size_t(o) * p.c_out + n0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              if (c0 + j + 1 < ncols) {
                const float x0 = __uint_as_float(v[j]) + (p.bias ? __ldg(p.bias + n0 + c0 + j) : 0.f);
                const float x1 = __uint_as_float(v[j + 1]) + (p.bias ? __ldg(p.bias + n0 + c0 + j + 1) : 0.f);
                *reinterpret_cast<__nv_bfloat162*>(orow + j) = __floats2bfloat162_rn(x0, x1);
              } else if (c0 + j < ncols) {
                orow[j] = __float2bfloat16(__uint_as_float(v[j]) + (p.bias ? __ldg(p.bias + n0 + c0 + j) : 0.f));
              }
            }
          } else {
            float* orow = static_cast<float*>(p.out) + size_t(o) * p.c_out + n0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (c0 + j + 3 < ncols) {
                float4 x;
                x.x = __uint_as_float(v[j]) + (p.bias ? __ldg(p.bias + n0 + c0 + j) : 0.f);
                x.y = __uint_as_float(v[j + 1]) + (p.bias ? __ldg(p.bias + n0 + c0 + j + 1) : 0.f);
                x.z = __uint_as_float(v[j + 2]) + (p.bias ? __ldg(p.bias + n0 + c0 + j + 2) : 0.f);
                x.w = __uint_as_float(v[j + 3]) + (p.bias ? __ldg(p.bias + n0 + c0 + j + 3) : 0.f);
                *reinterpret_cast<float4*>(orow + j) = x;
              } else {
                for (int jj = j; jj < j + 4; ++jj)
                  if (c0 + jj < ncols) orow[jj] = __uint_as_float(v[jj]) + (p.bias ? __ldg(p.bias + n0 + c0 + jj) : 0.f);
              }
            }
          }
        }
      }
    }
    tc_fence_before();
  } else if (warp == T2_PROD_WARPS) {
    // =================================== MMA issuer (warp-uniform loop, one elected lane issues) ================
    {
      const uint32_t fmt = BF16 ? 1u : 2u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(p.n_tile >> 3) << 17) | (uint32_t(BM >> 4) << 24);
      const uint32_t desc_hi = uint32_t(1024 >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_ring_base = smem_u32(a_ring), b_ring_base = smem_u32(b_ring);
      const int row_bytes = p.row_bytes;
      int sa = 0, sb = 0, tsa = 0;
      uint32_t pha = 0, phb = 0, phta = 0;
      (void)sa; (void)pha; (void)tsa; (void)phta; (void)a_ring_base;
      for (int k = 0; k < K; ++k) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(b_full + sb, phb);
          if constexpr (PRECISE) fence_proxy_async();
          const uint32_t b_base = b_ring_base + uint32_t(sb) * b_bytes;
          const uint32_t b_lo32 = ((b_base >> 4) & 0x3FFF) | (1u << 16);
          const int valid = min(KBLOCK_BYTES, row_bytes - kb * KBLOCK_BYTES);
          const int ksteps = (valid + 31) >> 5;
          const uint32_t acc_flag = (k | kb) ? 1u : 0u;
          for (int t = 0; t < TM; ++t) {
            const uint32_t d_addr = tmem_base + uint32_t(t * p.n_tile);
            if constexpr (PRECISE) {
              // A operand (hi | lo) comes from TMEM, written by the splitter warps; only the weights are read from smem
              mbar_wait(ta_full + tsa, phta);
              tc_fence_after();
              const uint32_t a_tm = tmem_base + uint32_t(p.ta_col0 + tsa * 64);
              if (elect_one()) {
#pragma unroll 4
                for (int j = 0; j < ksteps; ++j) {
                  const uint64_t b_hi = desc_from(b_lo32 + 2 * j, desc_hi);
                  umma_ts_tf32(d_addr, a_tm + 8 * j, b_hi, idesc, acc_flag | uint32_t(j));
                  umma_ts_tf32(d_addr, a_tm + 32 + 8 * j, b_hi, idesc, 1u);
                  umma_ts_tf32(d_addr, a_tm + 8 * j, desc_from(b_lo32 + (p.b_stage_bytes >> 4) + 2 * j, desc_hi), idesc, 1u);
                }
                umma_commit(ta_empty + tsa);
              }
              __syncwarp();
              if (++tsa == p.ta_stages) {
                tsa = 0;
                phta ^= 1;
              }
            } else {
              mbar_wait(a_full + sa, pha);
              fence_proxy_async();   // generic-proxy writes (cp.async) -> visible to the tensor core's reads
              tc_fence_after();
              const uint32_t a_base = a_ring_base + uint32_t(sa) * A_BYTES;
              const uint32_t a_lo32 = ((a_base >> 4) & 0x3FFF) | (1u << 16);
              if (elect_one()) {
#pragma unroll 4
                for (int j = 0; j < ksteps; ++j)
                  umma<BF16>(d_addr, desc_from(a_lo32 + 2 * j, desc_hi), desc_from(b_lo32 + 2 * j, desc_hi), idesc,
                             acc_flag | uint32_t(j));
                umma_commit(a_empty + sa);
              }
              __syncwarp();
              if (++sa == SA) {
                sa = 0;
                pha ^= 1;
              }
            }
          }
          if (elect_one()) umma_commit(b_empty + sb);
          __syncwarp();
          if (++sb == SB) {
            sb = 0;
            phb ^= 1;
          }
        }
      }
      if (elect_one()) umma_commit(acc_bar);
    }
    __syncwarp();
  } else if (warp == T2_PROD_WARPS + 1) {
    // =================================== weight TMA producer (warp-uniform loop, one elected lane issues) ======
    {
      const int kelems = KBLOCK_BYTES / (BF16 ? 2 : 4);
      const uint32_t b_ring_base = smem_u32(b_ring);
      int sb = 0;
      uint32_t phb = 0;
      for (int k = 0; k < K; ++k) {
        const int row = k * p.c_out + n0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(b_empty + sb, phb ^ 1);
          const uint32_t b_dst = b_ring_base + uint32_t(sb) * b_bytes;
          if (elect_one()) {
            mbar_expect_tx(b_full + sb, b_bytes);
            tma_load_2d(b_dst, &tmap_w, b_full + sb, kb * kelems, row);
            if constexpr (PRECISE) tma_load_2d(b_dst + p.b_stage_bytes, &tmap_w, b_full + sb, kb * kelems, K * p.c_out + row);
          }
          __syncwarp();
          if (++sb == SB) {
            sb = 0;
            phb ^= 1;
          }
        }
      }
    }
    __syncwarp();
  }
  if constexpr (PRECISE) {
    if (warp >= T2_PROD_WARPS + 2) {
      // =================================== hi/lo splitters (128 threads): smem tile -> TMEM ===================================
      // thread <-> row of the tile (TMEM lane): read the row's 8 swizzled 16-byte chunks, split, tcgen05.st hi and lo.
      const int lq = warp & 3;                     // TMEM lane quadrant this warp may access
      const int r = lq * 32 + lane;
      const uint32_t lane_addr = uint32_t(lq * 32) << 16;
      const int total = K * num_kb * TM;
      int s = 0, ts = 0;
      uint32_t ph = 0, pht = 0;
      for (int it = 0; it < total; ++it) {
        mbar_wait(a_full + s, ph);                 // gathered rows have landed (cp.async completion arrivals)
        const uint8_t* a_row = a_ring + size_t(s) * A_BYTES + r * KBLOCK_BYTES;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(a_row + ((c ^ (r & 7)) << 4));
          const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t h = __float_as_uint(f[e]) & 0xFFFFE000u;
            hi[c * 4 + e] = h;
            lo[c * 4 + e] = __float_as_uint(f[e] - __uint_as_float(h));
          }
        }
        mbar_arrive(a_empty + s);                  // the shared-memory stage can be refilled
        mbar_wait(ta_empty + ts, pht ^ 1);         // MMAs that read this TMEM stage last time are done
        tc_fence_after();
        const uint32_t ta = tmem_base + lane_addr + uint32_t(p.ta_col0 + ts * 64);
        tmem_st32(ta, hi);
        tmem_st32(ta + 32, lo);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(ta_full + ts);
        if (++s == SA) {
          s = 0;
          ph ^= 1;
        }
        if (++ts == p.ta_stages) {
          ts = 0;
          pht ^= 1;
        }
      }
    }
  }

  __syncthreads();
  if (warp == T2_PROD_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(p.tmem_cols))
                 : "memory");
  }
}

// ---- host side -------------------------------------------------------------------------------------------

}  // namespace tc

bool tc_built() { return true; }

// One launch per layer: from W [K, c_in, c_out] (fp32) produce the tensor-core operand forms
//   fwd  [nsplit, K, c_out, c_in]  (per-offset transpose = K-major B of the forward GEMM)
//   bwd  [nsplit, K, c_in, c_out]  (K-major B of the dgrad GEMM: W itself)
// nsplit = 1: plain copy / transpose (bf16 or fp32 -> TF32 read by the hardware);
// nsplit = 2: hi = RN_tf32(w), lo = RN_tf32(w - hi) for the 3xTF32 path.
template <typename T>
__device__ __forceinline__ void wp_store(T* fwd_or_bwd, int64_t idx, int64_t total, float v, int nsplit) {
  if constexpr (sizeof(T) == 2) {
    fwd_or_bwd[idx] = __float2bfloat16(v);
  } else if (nsplit == 1) {
    fwd_or_bwd[idx] = v;
  } else {
    uint32_t hb, lb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
    const float hi = __uint_as_float(hb);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(v - hi));
    fwd_or_bwd[idx] = hi;
    fwd_or_bwd[total + idx] = __uint_as_float(lb);
  }
}

// grid = (ceil(c_out/32), ceil(c_in/32), K), block = (32, 8): 32x32 tile through shared memory so that both the
// [c_in][c_out] reads and the transposed [c_out][c_in] writes are coalesced
template <typename T>
__global__ void __launch_bounds__(256) weight_prep_kernel(const float* __restrict__ w, int K, int c_in, int c_out,
                                                          int nsplit, T* __restrict__ fwd, T* __restrict__ bwd) {
  __shared__ float tile[32][33];
  const int k = blockIdx.z;
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
  const int64_t total = int64_t(K) * c_in * c_out;
  const float* wk = w + int64_t(k) * c_in * c_out;
#pragma unroll
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int ci = ci0 + r, co = co0 + threadIdx.x;
    float v = 0.f;
    if (ci < c_in && co < c_out) {
      v = wk[int64_t(ci) * c_out + co];
      if (bwd) wp_store<T>(bwd, int64_t(k) * c_in * c_out + int64_t(ci) * c_out + co, total, v, nsplit);
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  if (!fwd) return;
#pragma unroll
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int co = co0 + r, ci = ci0 + threadIdx.x;
    if (ci < c_in && co < c_out)
      wp_store<T>(fwd, int64_t(k) * c_in * c_out + int64_t(co) * c_in + ci, total, tile[threadIdx.x][r], nsplit);
  }
}

// All layers of a network in ONE launch (the per-layer launches are latency-bound: 63 of them per training step).
// desc [n_layers][8] int64 = {w, fwd, bwd, K, c_in, c_out, first tile of the layer, unused}; a block finds its layer by
// binary search over the first-tile column, then does the same 32x32 tile as weight_prep_kernel.
template <typename T>
__global__ void __launch_bounds__(256) weight_prep_batch_kernel(const int64_t* __restrict__ desc, int n_layers, int nsplit) {
  __shared__ float tile[32][33];
  const int64_t b = blockIdx.x;
  int lo = 0, hi = n_layers - 1;
  while (lo < hi) {                       // last layer whose first tile <= b
    const int mid = (lo + hi + 1) >> 1;
    if (desc[mid * 8 + 6] <= b) lo = mid; else hi = mid - 1;
  }
  const int64_t* d = desc + lo * 8;
  const float* w = reinterpret_cast<const float*>(d[0]);
  T* fwd = reinterpret_cast<T*>(d[1]);
  T* bwd = reinterpret_cast<T*>(d[2]);
  const int K = int(d[3]), c_in = int(d[4]), c_out = int(d[5]);
  (void)K;
  const int tx = (c_out + 31) / 32, ty = (c_in + 31) / 32;
  int64_t t = b - d[6];
  const int co0 = int(t % tx) * 32;
  t /= tx;
  const int ci0 = int(t % ty) * 32;
  const int k = int(t / ty);
  const int64_t total = int64_t(d[3]) * c_in * c_out;
  const float* wk = w + int64_t(k) * c_in * c_out;
#pragma unroll
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int ci = ci0 + r, co = co0 + threadIdx.x;
    float v = 0.f;
    if (ci < c_in && co < c_out) {
      v = wk[int64_t(ci) * c_out + co];
      if (bwd) wp_store<T>(bwd, int64_t(k) * c_in * c_out + int64_t(ci) * c_out + co, total, v, nsplit);
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  if (!fwd) return;
#pragma unroll
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int co = co0 + r, ci = ci0 + threadIdx.x;
    if (ci < c_in && co < c_out)
      wp_store<T>(fwd, int64_t(k) * c_in * c_out + int64_t(co) * c_in + ci, total, tile[threadIdx.x][r], nsplit);
  }
}

int weight_prep_batch(const int64_t* desc, int n_layers, int64_t total_tiles, int nsplit, int dtype, cudaStream_t stream) {
  if (n_layers == 0 || total_tiles == 0) return LGS_OK;
  const dim3 block{32u, 8u, 1u};
  if (dtype == LGS_BF16) {
    LGS_LAUNCH(weight_prep_batch_kernel<__nv_bfloat16>, unsigned(total_tiles), block, 0, stream, desc, n_layers, 1);
  } else {
    LGS_LAUNCH(weight_prep_batch_kernel<float>, unsigned(total_tiles), block, 0, stream, desc, n_layers, nsplit);
  }
  return LGS_OK;
}

int weight_prep(const float* w, int K, int c_in, int c_out, int nsplit, void* fwd, void* bwd, int dtype,
                cudaStream_t stream) {
  const int64_t total = int64_t(K) * c_in * c_out;
  if (total == 0) return LGS_OK;
  const dim3 grid{unsigned((c_out + 31) / 32), unsigned((c_in + 31) / 32), unsigned(K)};
  const dim3 block{32u, 8u, 1u};
  if (dtype == LGS_BF16) {
    LGS_LAUNCH(weight_prep_kernel<__nv_bfloat16>, grid, block, 0, stream, w, K, c_in, c_out, 1,
               static_cast<__nv_bfloat16*>(fwd), static_cast<__nv_bfloat16*>(bwd));
  } else {
    LGS_LAUNCH(weight_prep_kernel<float>, grid, block, 0, stream, w, K, c_in, c_out, nsplit,
               static_cast<float*>(fwd), static_cast<float*>(bwd));
  }
  return LGS_OK;
}

// Weights arrive K-major for this path: d_weight_nk = [K, c_out, c_in] (c_in contiguous).
// 0 if the tensor-core kernels take this shape (the facade asks before choosing the weight layout)
int conv_tc_shape_ok(int c_in, int c_out, int dtype) {
  const int es = dtype == LGS_BF16 ? 2 : 4;
  const int row_bytes = c_in * es;
  if (row_bytes % 16 != 0 || row_bytes < 16) return 0;
  if (dtype == LGS_BF16 ? (c_out % 2 != 0) : (c_out % 4 != 0)) return 0;
  return 1;
}

int conv_fwd_tc(const void* in, int64_t n_in, int c_in, const void* w_nk, int K, int c_out, const int32_t* table,
                int64_t n_out, int reverse_k, const float* bias, void* out, int dtype, int precise, cudaStream_t stream) {
  using namespace tc;
  if (dtype == LGS_BF16) precise = 0;
  const int nsplit = precise ? 2 : 1;
  const int es = dtype == LGS_BF16 ? 2 : 4;
  const int row_bytes = c_in * es;
  if (row_bytes % 16 != 0 || row_bytes < 16) return LGS_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(w_nk) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15))
    return LGS_E_UNSUPPORTED;
  if (dtype == LGS_BF16 ? (c_out % 2 != 0) : (c_out % 4 != 0)) return LGS_E_UNSUPPORTED;  // vector stores
  if (int64_t(K) * c_out * 2 >= (int64_t(1) << 31)) return LGS_E_UNSUPPORTED;
  if (!conv_tc_shape_ok(c_in, c_out, dtype)) return LGS_E_UNSUPPORTED;
  if (n_out == 0) return LGS_OK;
  EncodeTiledFn encode = get_encode();
  if (!encode) return fail(LGS_E_CUDA, "cuTensorMapEncodeTiled entry point not available");

  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    const void* fns[] = {(const void*)conv_tc_kernel<false, false>, (const void*)conv_tc_kernel<true, false>,
                         (const void*)conv_tc_kernel<false, true>,  (const void*)conv_tc2_kernel<false, false>,
                         (const void*)conv_tc2_kernel<true, false>, (const void*)conv_tc2_kernel<false, true>};
    for (const void* f : fns)
      if (attr_err == cudaSuccess)
        attr_err = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) return fail(LGS_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));

  // ---- large maps: multi-tile CTAs sharing each weight block across TM row tiles --------------------------------
  {
    const int c_pad_all = ((c_out + 15) / 16) * 16;
    // output-channel slices of <= 256 columns (multiples of 16); the last slice may be narrower (544 = 192 + 192 + 160):
    // its surplus MMA columns read the next offset's weight rows / TMA zero fill and are never stored
    const int n_slices = (c_pad_all + 255) / 256;
    const int c_pad2 = ((((c_pad_all + n_slices - 1) / n_slices) + 15) / 16) * 16;     // columns per CTA
    const int64_t tiles = cdiv(n_out, BM);
    int TM = 0, RT = BM;
    const int ta_min_cols = precise ? 128 : 0;   // PRECISE keeps >= 2 split-A stages (64 columns each) in TMEM
    if (tiles * n_slices >= 148 && n_in > 0 && !getenv("LGS_TC_NO_MULTI")) {
      // Cost of a decomposition = tile-rows an SM processes in sequence.  (a) full 128-row tiles: waves * TM.
      // (b) balanced: a grid of w * 148 CTAs (exactly w full waves) with TM tiles of rt = ceil(rows per CTA / TM) <= 128
      // rows: the gather stream shrinks with rt, the MMAs do not (empty lanes), hence the floor.  A map of 301 tiles
      // (level 1 of a 150 K-voxel scene) costs 3 as 128-row tiles (148 + 148 + 5 CTAs) but 3 * 87/128 = 2.04 balanced.
      // Ties go to the larger TM (less weight traffic).
      const bool balance = !getenv("LGS_TC_NO_BALANCE");
      const double mma_floor = precise ? 0.6 : 0.35;
      double best = -1.0;
      const bool ts1 = precise && getenv("LGS_TC_TS1");   // TM = 1 otherwise means conv_tc_kernel below
      for (int tm = 1; tm <= 4; ++tm) {
        if (tm * c_pad2 + ta_min_cols > 512) break;
        if (tm != 3) {
          const double cost = double(cdiv(cdiv(tiles, tm) * n_slices, 148) * tm);
          if (best < 0 || cost <= best) best = cost, TM = tm, RT = BM;
        }
        if (!balance || (tm == 1 && !ts1)) continue;
        for (int w = 1; w <= 64; ++w) {
          const int64_t gx = std::max<int64_t>(1, int64_t(148) * w / n_slices);
          const int64_t rt = cdiv(cdiv(n_out, gx), int64_t(tm));
          if (rt > BM) continue;
          const double cost = double(w) * tm * std::max(double(rt) / BM, mma_floor);
          if (cost <= best * 0.97) best = cost, TM = tm, RT = int(rt);   // must beat full tiles by 3 %
          break;
        }
      }
      if (const char* e = getenv("LGS_TC_TM")) {      // tuning / test override, clamped to what TMEM holds
        TM = std::min(4, std::max(1, atoi(e))), RT = BM;
        while (TM > 1 && TM * c_pad2 + ta_min_cols > 512) --TM;
      }
      if (const char* e = getenv("LGS_TC_RT")) RT = std::min(BM, std::max(8, atoi(e)));
    }
    if (TM >= 2 || (TM == 1 && precise && getenv("LGS_TC_TS1"))) {
      Params2 q;
      q.in = static_cast<const uint8_t*>(in);
      q.row_bytes = row_bytes;
      q.K = K;
      q.c_out = c_out;
      q.table = table;
      q.n_out = n_out;
      q.reverse_k = reverse_k;
      q.bias = bias;
      q.out = out;
      q.num_kb = (row_bytes + KBLOCK_BYTES - 1) / KBLOCK_BYTES;
      q.n_tile = c_pad2;
      q.TM = TM;
      q.rt = RT;
      q.b_stage_bytes = c_pad2 * KBLOCK_BYTES;
      q.ta_col0 = ((TM * c_pad2 + 31) / 32) * 32;
      q.ta_stages = precise ? (512 - q.ta_col0) / 64 : 0;
      if (q.ta_stages > MAX_TA_STAGES) q.ta_stages = MAX_TA_STAGES;
      int cols = 32;
      while (cols < (precise ? q.ta_col0 + q.ta_stages * 64 : TM * c_pad2)) cols <<= 1;
      q.tmem_cols = cols;
      const int a_bytes = A_STAGE_BYTES, b_bytes = nsplit * q.b_stage_bytes;
      const int fixed2 = (2 * MAX_A_STAGES + 2 * MAX_B_STAGES + 2 * MAX_TA_STAGES + 1) * 8 + 16 + 1024;
      int sb = 2;
      int sa = (227 * 1024 - fixed2 - sb * b_bytes) / a_bytes;
      if (sa > MAX_A_STAGES) sa = MAX_A_STAGES;
      if (sa >= 6 && (227 * 1024 - fixed2 - 3 * b_bytes) / a_bytes >= 5) {   // room for a third weight stage
        sb = 3;
        sa = (227 * 1024 - fixed2 - sb * b_bytes) / a_bytes;
        if (sa > MAX_A_STAGES) sa = MAX_A_STAGES;
      }
      if (sa >= 2) {
        q.a_stages = sa;
        q.b_stages = sb;
        const size_t smem2 = size_t(sa) * a_bytes + size_t(sb) * b_bytes + fixed2;
        CUtensorMap tmap2;
        const cuuint64_t gdim2[2] = {cuuint64_t(c_in), cuuint64_t(nsplit) * cuuint64_t(K) * cuuint64_t(c_out)};
        const cuuint64_t gstride2[1] = {cuuint64_t(row_bytes)};
        const cuuint32_t box2[2] = {cuuint32_t(KBLOCK_BYTES / es), cuuint32_t(c_pad2)};
        const cuuint32_t estr2[2] = {1, 1};
        const CUresult cr2 = encode(&tmap2, dtype == LGS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                    2, const_cast<void*>(w_nk), gdim2, gstride2, box2, estr2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr2 != CUDA_SUCCESS) return fail(LGS_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", int(cr2));
        const dim3 grid2{unsigned(cdiv(n_out, int64_t(TM) * RT)), unsigned(n_slices), 1u};
        if (dtype == LGS_BF16) {
          LGS_LAUNCH((conv_tc2_kernel<true, false>), grid2, T2_THREADS, smem2, stream, tmap2, q);
        } else if (precise) {
          LGS_LAUNCH((conv_tc2_kernel<false, true>), grid2, T2_THREADS + 128, smem2, stream, tmap2, q);
        } else {
          LGS_LAUNCH((conv_tc2_kernel<false, false>), grid2, T2_THREADS, smem2, stream, tmap2, q);
        }
        return LGS_OK;
      }
    }
  }

  Params p;
  p.in = static_cast<const uint8_t*>(in);
  p.n_in = n_in;
  p.row_bytes = row_bytes;
  p.K = K;
  p.c_out = c_out;
  p.table = table;
  p.n_out = n_out;
  p.reverse_k = reverse_k;
  p.bias = bias;
  p.out = out;
  p.num_kb = (row_bytes + KBLOCK_BYTES - 1) / KBLOCK_BYTES;
  const int c_pad = ((c_out + 15) / 16) * 16;
  const int64_t m_tiles = cdiv(n_out, BM);
  // Work decomposition.  Large maps: one CTA per 128-row tile (all offsets, all channels <= 256) — deterministic,
  // no atomics.  Small maps (fewer tiles than SMs; the coarse U-Net levels, where the WEIGHTS are the traffic):
  // split the kernel offsets (and, if still short, the output channels) over CTAs; partial sums meet in global memory
  // through red.add on a zeroed output.
  int n_tiles = (c_pad + 255) / 256;
  p.n_tile = ((((c_pad + n_tiles - 1) / n_tiles) + 15) / 16) * 16;   // <= 256; the last slice may be narrower (544 = 192+192+160)
  int k_splits = 1;
  // CTAs to aim for when a map has fewer tiles than SMs: two of these CTAs fit on an SM (<= 100 KB of shared memory each),
  // and 296 measured 5-15 % faster than 148 on the 8.5 K-row level (scripts/dev_small_layers.py), equal elsewhere
  int split_target = 296;
  if (const char* e = getenv("LGS_TC_SPLIT_TARGET")) split_target = std::max(1, atoi(e));
  if (m_tiles * n_tiles <= split_target / 2) {
    int want = int(cdiv(split_target, m_tiles * n_tiles));
    if (dtype == LGS_F32 && K > 1) {
      k_splits = want < K ? want : K;
      want = int(cdiv(want, k_splits));
    }
    if (want > 1 && n_tiles == 1 && c_pad >= 128) {   // split channels too (N >= 64 per CTA)
      int ns = c_pad / 64;
      if (ns > want) ns = want;
      while (ns > 1 && (c_pad % (ns * 16)) != 0) --ns;
      if (ns > 1) {
        p.n_tile = c_pad / ns;
        n_tiles = ns;
      }
    }
  }
  p.k_per_split = (K + k_splits - 1) / k_splits;
  k_splits = (K + p.k_per_split - 1) / p.k_per_split;
  p.k_splits = k_splits;
  p.b_stage_bytes = p.n_tile * KBLOCK_BYTES;
  int cols = 32;
  while (cols < p.n_tile) cols <<= 1;
  p.tmem_cols = cols;
  const int stage_bytes = nsplit * (A_STAGE_BYTES + p.b_stage_bytes);
  const int fixed = p.k_per_split * BM * 4 + 64 * 4 + (3 * MAX_STAGES + 1) * 8 + 16 + 1024;  // idx, lists, barriers, align
  int stages = (100 * 1024 - fixed) / stage_bytes;            // try to leave room for two CTAs per SM
  if (stages < MIN_STAGES + 1) stages = (227 * 1024 - fixed) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (const char* e = getenv("LGS_TC_STAGES")) { const int v = atoi(e); if (v >= 2 && v < stages) stages = v; }  // tuning knob
  if (stages < (precise ? 2 : MIN_STAGES)) return LGS_E_UNSUPPORTED;
  p.stages = stages;
  const size_t smem_bytes = size_t(stages) * stage_bytes + fixed;
  if (k_splits > 1) LGS_CUDA(cudaMemsetAsync(out, 0, size_t(n_out) * c_out * sizeof(float), stream));

  // tensor map over W^T viewed as [K * c_out rows, c_in] with box {128 B of channels, n_tile rows}, 128B swizzle
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {cuuint64_t(c_in), cuuint64_t(nsplit) * cuuint64_t(K) * cuuint64_t(c_out)};
  const cuuint64_t gstride[1] = {cuuint64_t(row_bytes)};
  const cuuint32_t box[2] = {cuuint32_t(KBLOCK_BYTES / es), cuuint32_t(p.n_tile)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = encode(&tmap, dtype == LGS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                             2, const_cast<void*>(w_nk), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(LGS_E_CUDA, "cuTensorMapEncodeTiled failed (%d) c_in=%d c_out=%d K=%d", int(cr), c_in, c_out, K);

  const dim3 grid{unsigned(m_tiles), unsigned(n_tiles), unsigned(k_splits)};
  if (dtype == LGS_BF16) {
    LGS_LAUNCH((conv_tc_kernel<true, false>), grid, THREADS, smem_bytes, stream, tmap, p);
  } else if (precise) {
    LGS_LAUNCH((conv_tc_kernel<false, true>), grid, THREADS + 128, smem_bytes, stream, tmap, p);
  } else {
    LGS_LAUNCH((conv_tc_kernel<false, false>), grid, THREADS, smem_bytes, stream, tmap, p);
  }
  return LGS_OK;
}

// =============================================================================================================
// wgrad on tensor cores:  dW[k] (c_in x c_out) = sum_o  X[table[k][o], :]^T (outer) dY[o, :]
//
// The reduction dimension of the GEMM is the (gathered) row index, so both operands are "MN-major": a shared-memory
// tile [rows][128 B of channels] with the 128B swizzle is exactly the canonical MN-major layout (rows = K, 8-row
// groups 1024 B apart, channel blocks LBO apart) — the same tiles the forward kernel gathers, read through a
// different descriptor.  One MMA consumes 8 rows (tf32) / 16 rows (bf16).
//
// CTA = (chunk of rows, group of G offsets).  Per tile of R rows: TMA loads the dY tile once (shared by the G offsets),
// the producers gather X for each offset, the MMA thread accumulates into G*MC accumulators in TMEM (MC = c_in/128
// lane chunks).  At the end the accumulators are reduced into global dW with vector red.add.
// =============================================================================================================
namespace tcw {
using namespace tc;

constexpr int MAX_WA_STAGES = 8;

struct WParams {
  const uint8_t* in;       // X [n_in, c_in]
  int32_t in_row_bytes;
  const int32_t* table;    // [K, n_out] or nullptr
  int64_t n_out;
  int32_t K, c_in, c_out;
  float* gw;               // [K, c_in, c_out] fp32, pre-zeroed
  int32_t n0_step;         // output-channel slice per CTA (blockIdx.z selects it); == n_cols unless channels are split
  int32_t R;               // rows per tile (32 / 64 / 128)
  int32_t nb_in, nb_out;   // 128-byte channel blocks of X / dY
  int32_t G;               // offsets per CTA
  int32_t MC;              // 128-lane chunks of c_in handled by one CTA
  int32_t MC_total;        // all chunks of c_in (blockIdx.z / n_splits selects the CTA's first chunk)
  int32_t n_splits;        // output-channel slices
  int32_t n_cols;          // accumulator columns (c_out padded to 16)
  int32_t tmem_cols;
  int64_t rows_per_chunk;  // multiple of R
  int32_t a_stages;        // gather ring depth
  int32_t b_bufs;          // dY tile buffers (2 = double buffered, 1 when shared memory is tight)
  int32_t nb_in_real;      // channel blocks actually gathered per stage (the MMA may read up to MC*4 blocks: the tail
                           // aliases the next stage / buffer = finite garbage in accumulator rows >= c_in, never stored)
};

// MN-major operand descriptor over a tile [rows][128 B of channels].
//   16-bit types: SWIZZLE_128B (16-byte chunks XOR row%8), swizzle atom = 8 rows  -> SBO 1024 B
//   32-bit types: SWIZZLE_128B_BASE32B (32-byte chunks XOR row%4), atom = 4 rows   -> SBO  512 B
// (CuTe: Layout_MN_SW128_Atom / Layout_MN_SW128_32B_Atom; the transposing read of 32-bit elements needs the 32 B base.)
template <bool BF16>
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((saddr >> 4) & 0x3FFF);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;            // LBO: stride between 128-byte channel blocks
  d |= uint64_t((BF16 ? 1024 : 512) >> 4) << 32;              // SBO: stride between swizzle atoms along the rows
  d |= uint64_t(1) << 46;
  d |= uint64_t(BF16 ? 2 : 1) << 61;
  return d;
}
// position of 16-byte chunk c of row r inside the 128-byte row
template <bool BF16>
__device__ __forceinline__ uint32_t mn_chunk_pos(uint32_t c, uint32_t r) {
  if constexpr (BF16) return c ^ (r & 7);
  return ((((c >> 1) ^ (r & 3)) << 1) | (c & 1));
}

template <bool BF16>
__global__ void __launch_bounds__(THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_gy, const WParams p) {
  pdl_grid_sync();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int blk_bytes = p.R * KBLOCK_BYTES;             // one channel block of one tile
  const int a_stage_bytes = p.nb_in_real * blk_bytes;
  const int A_STAGES = p.a_stages;
  const int b_buf_bytes = p.nb_out * blk_bytes;
  // B (dY) buffers first, gather ring last: the MMA's over-read of a stage's missing tail blocks stays inside the ring
  // or runs into the barrier words / padding that follow it (finite garbage, see WParams::nb_in_real)
  uint8_t* b_smem = smem;
  uint8_t* a_smem = smem + size_t(p.b_bufs) * size_t(b_buf_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_smem + size_t(A_STAGES) * a_stage_bytes + size_t(p.nb_in - p.nb_in_real) * blk_bytes);
  uint64_t* afull = bars;                       // [MAX_WA_STAGES]
  uint64_t* aempty = bars + MAX_WA_STAGES;      // [MAX_WA_STAGES]
  uint64_t* bfull = bars + 2 * MAX_WA_STAGES;   // [2]
  uint64_t* bempty = bfull + 2;                 // [2]
  uint64_t* acc_bar = bempty + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t r_begin = int64_t(blockIdx.x) * p.rows_per_chunk;
  const int64_t r_end = min(p.n_out, r_begin + p.rows_per_chunk);
  const int k0 = blockIdx.y * p.G;
  const int n0 = (blockIdx.z % p.n_splits) * p.n0_step;   // first output channel of this CTA's slice
  const int mc0 = (blockIdx.z / p.n_splits) * p.MC;       // first 128-channel chunk of the input channels it owns
  const int mc_cnt = min(p.MC, p.MC_total - mc0);
  const int in_col0 = mc0 * 128 * (BF16 ? 2 : 4);         // byte offset of that chunk inside a feature row
  const int ncols_valid = min(p.n_cols, p.c_out - n0);
  const int g_count = min(p.G, p.K - k0);
  const int n_tiles = int((r_end - r_begin + p.R - 1) / p.R);

  if (tid == 0) {
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(afull + s, 128);
      mbar_init(aempty + s, 1);
    }
    for (int s = 0; s < p.b_bufs; ++s) {
      mbar_init(bfull + s, 1);
      mbar_init(bempty + s, 1);
    }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(uint32_t(p.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 5 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_gy) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < 4) {
    // =================================== gather producers ===================================
    const int chunk = tid & 7, rbase = tid >> 3;
    const int passes = p.R / 16;
    uint32_t dst_off[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = rbase + 16 * i;
      dst_off[i] = uint32_t(r * KBLOCK_BYTES) + (mn_chunk_pos<BF16>(chunk, r) << 4);
    }
    const int row_bytes = p.in_row_bytes;
    const uint8_t* in_chunk = p.in + in_col0 + chunk * 16;
    const uint32_t a_smem_base = smem_u32(a_smem);
    int s = 0;
    uint32_t ph = 0;
    // table entries of step (t, g), read one step ahead so their latency hides behind the previous step's copies
    auto load_rows = [&](int t, int g, int32_t (&v)[8]) {
      const int32_t* trow = p.table ? p.table + int64_t(k0 + g) * p.n_out : nullptr;
      const int64_t row0 = r_begin + int64_t(t) * p.R + rbase;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t o = row0 + 16 * i;
        v[i] = (i < passes && o < r_end) ? (trow ? ldg_nc_ordered(trow + o) : int32_t(o)) : -1;
      }
    };
    // table entries are read TWO steps ahead: a load issued behind a stage's 24 cp.async per thread drains through the same
    // LSU queue and came back only after the next stage had started (25 % of the producers' samples, profiles/r2_wgrad_L0_96_ncu.txt)
    int32_t cur[8], nxt[8], nx2[8];
    const int total_steps = n_tiles * g_count;
    auto load_step = [&](int step, int32_t (&v)[8]) {
      if (step < total_steps) load_rows(step / g_count, step % g_count, v);
    };
    load_step(0, cur);
    load_step(1, nxt);
    int step = 0;
    for (int t = 0; t < n_tiles; ++t) {
      for (int g = 0; g < g_count; ++g, ++step) {
        load_step(step + 2, nx2);
        mbar_wait(aempty + s, ph ^ 1);
        uint32_t a_base = a_smem_base + uint32_t(s) * uint32_t(a_stage_bytes);
        int col = chunk * 16;
        for (int kb = 0; kb < p.nb_in_real; ++kb, col += KBLOCK_BYTES, a_base += blk_bytes) {
          const bool col_ok = in_col0 + col < row_bytes;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (i < passes) {
              const bool ok = col_ok && cur[i] >= 0;
              cp_async16(a_base + dst_off[i], in_chunk + size_t(ok ? cur[i] : 0) * row_bytes + kb * KBLOCK_BYTES, ok ? 16u : 0u);
            }
          }
        }
        cp_async_mbar_arrive_noinc(afull + s);
        if (++s == A_STAGES) {
          s = 0;
          ph ^= 1;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = nxt[i], nxt[i] = nx2[i];
      }
    }
    // =================================== epilogue: TMEM -> red.add into dW ===================================
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    for (int g = 0; g < (n_tiles > 0 ? g_count : 0); ++g) {
      for (int mc = 0; mc < mc_cnt; ++mc) {
        const int ci = (mc0 + mc) * 128 + warp * 32 + lane;
        const uint32_t col_base = uint32_t((g * p.MC + mc) * p.n_cols);
        for (int c0 = 0; c0 < ncols_valid; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + (uint32_t(warp * 32) << 16) + col_base + uint32_t(c0), v);
          if (ci < p.c_in) {
            float* dst = p.gw + (size_t(k0 + g) * p.c_in + ci) * p.c_out + n0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (c0 + j + 3 < ncols_valid) {
                red_add_v4(dst + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                           __uint_as_float(v[j + 3]));
              } else {
                for (int jj = j; jj < j + 4; ++jj)
                  if (c0 + jj < ncols_valid) atomicAdd(dst + jj, __uint_as_float(v[jj]));
              }
            }
          }
        }
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // =================================== MMA issuer (warp-uniform loop, one elected lane issues) ================
    {
      const uint32_t fmt = BF16 ? 1u : 2u;
      const int m_dim = 128;  // always whole 128-lane chunks (channels beyond c_in are zero-filled by the gather)
      // D fp32, A/B tf32|bf16, BOTH MN-major (bits 15, 16), N = n_cols, M = m_dim
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) |
                             (uint32_t(p.n_cols >> 3) << 17) | (uint32_t(m_dim >> 4) << 24);
      const int rows_per_mma = BF16 ? 16 : 8;
      const int mmas = p.R / rows_per_mma;
      const uint32_t mma_step16 = uint32_t(rows_per_mma * KBLOCK_BYTES) >> 4;   // K step (8 or 16 rows) in 16-byte units
      const int mc_blocks = 128 * (BF16 ? 2 : 4) / KBLOCK_BYTES;     // channel blocks per 128-lane chunk (4 fp32 / 2 bf16)
      // MN-major descriptor: hi word constant (SBO, version, layout); lo word = addr>>4 | (LBO>>4)<<16
      const uint32_t desc_hi = uint32_t((BF16 ? 1024 : 512) >> 4) | (1u << 14) | (uint32_t(BF16 ? 2 : 1) << 29);
      const uint32_t lbo16 = (uint32_t(blk_bytes) >> 4) << 16;
      const uint32_t a_smem_base = smem_u32(a_smem), b_smem_base = smem_u32(b_smem);
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int bs = p.b_bufs == 2 ? (t & 1) : 0;
        mbar_wait(bfull + bs, (p.b_bufs == 2 ? (t >> 1) : t) & 1);
        const uint32_t b_lo32 = (((b_smem_base + uint32_t(bs) * uint32_t(b_buf_bytes)) >> 4) & 0x3FFF) | lbo16;
        for (int g = 0; g < g_count; ++g) {
          mbar_wait(afull + s, ph);
          fence_proxy_async();
          tc_fence_after();
          const uint32_t a_base = a_smem_base + uint32_t(s) * uint32_t(a_stage_bytes);
          if (elect_one()) {
            for (int mc = 0; mc < mc_cnt; ++mc) {
              const uint32_t d_addr = tmem_base + uint32_t((g * p.MC + mc) * p.n_cols);
              const uint32_t a_lo32 = (((a_base + uint32_t(mc * mc_blocks) * uint32_t(blk_bytes)) >> 4) & 0x3FFF) | lbo16;
#pragma unroll 4
              for (int j = 0; j < mmas; ++j)
                umma<BF16>(d_addr, desc_from(a_lo32 + j * mma_step16, desc_hi), desc_from(b_lo32 + j * mma_step16, desc_hi),
                           idesc, (t > 0 || j > 0) ? 1u : 0u);
            }
            umma_commit(aempty + s);
          }
          __syncwarp();
          if (++s == A_STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        if (elect_one()) umma_commit(bempty + bs);
        __syncwarp();
      }
      if (elect_one()) umma_commit(acc_bar);
    }
    __syncwarp();
  } else {
    // =================================== dY tile TMA producer (warp-uniform loop, one elected lane issues) ======
    {
      const int kelems = KBLOCK_BYTES / (BF16 ? 2 : 4);
      for (int t = 0; t < n_tiles; ++t) {
        const int bs = p.b_bufs == 2 ? (t & 1) : 0;
        mbar_wait(bempty + bs, ((p.b_bufs == 2 ? (t >> 1) : t) & 1) ^ 1);
        const int64_t row0 = r_begin + int64_t(t) * p.R;
        if (elect_one()) {
          mbar_expect_tx(bfull + bs, uint32_t(b_buf_bytes));
          for (int kb = 0; kb < p.nb_out; ++kb)
            tma_load_2d(smem_u32(b_smem + size_t(bs) * b_buf_bytes + kb * blk_bytes), &tmap_gy, bfull + bs,
                        n0 + kb * kelems, int32_t(row0));
        }
        __syncwarp();
      }
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(p.tmem_cols))
                 : "memory");
  }
}

}  // namespace tcw

int conv_wgrad_tc(const void* in, int64_t n_in, int c_in, const void* gout, int64_t n_out, int c_out,
                  const int32_t* table, int K, float* gw, int dtype, cudaStream_t stream) {
  using namespace tcw;
  (void)n_in;
  const int es = dtype == LGS_BF16 ? 2 : 4;
  const int in_row_bytes = c_in * es, out_row_bytes = c_out * es;
  if (in_row_bytes % 16 != 0 || out_row_bytes % 16 != 0) return LGS_E_UNSUPPORTED;
  if (c_out % 4 != 0) return LGS_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(gout) & 15) ||
      (reinterpret_cast<uintptr_t>(gw) & 15))
    return LGS_E_UNSUPPORTED;
  if (n_out >= (int64_t(1) << 31)) return LGS_E_UNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  if (!encode) return fail(LGS_E_CUDA, "cuTensorMapEncodeTiled entry point not available");

  WParams p;
  p.in = static_cast<const uint8_t*>(in);
  p.in_row_bytes = in_row_bytes;
  p.table = table;
  p.n_out = n_out;
  p.K = K;
  p.c_in = c_in;
  p.c_out = c_out;
  p.gw = gw;
  p.nb_in = (in_row_bytes + KBLOCK_BYTES - 1) / KBLOCK_BYTES;
  p.nb_out = (out_row_bytes + KBLOCK_BYTES - 1) / KBLOCK_BYTES;
  // Decomposition over channels: output channels in slices of <= 256 columns (MMA N), input channels in groups of MC
  // 128-lane chunks with MC * n_cols <= 512 TMEM columns.  Slicing the INPUT channels is preferred: each CTA then gathers
  // only its own channels (the gather is the expensive stream; the dY tile is a plain TMA load).
  p.MC_total = (c_in + 127) / 128;
  const int c_pad_all = ((c_out + 15) / 16) * 16;
  const int n_splits = (c_pad_all + 255) / 256;
  p.n_cols = ((((c_pad_all + n_splits - 1) / n_splits) + 15) / 16) * 16;
  p.n0_step = p.n_cols;
  p.n_splits = n_splits;
  const int blocks_per_mc = 128 * es / KBLOCK_BYTES;   // channel blocks per 128-lane chunk: 4 (fp32) / 2 (bf16)
  const int nb_out_alloc = (p.n_cols * es + KBLOCK_BYTES - 1) / KBLOCK_BYTES;
  // Choose (MC, R, dY buffers): per-step handshakes dominate, so the tallest row tile wins (R = 128 with two gather stages
  // beats R = 64 with six: 0.34 vs 0.56 ms on the 96->96 L0 layer); ties go to the larger MC.  The MMA reads whole
  // 128-lane chunks; blocks a CTA does not gather (c_in not a multiple of 128) alias the next stage: finite garbage in
  // accumulator rows >= c_in, never stored.
  int mc_max = 512 / p.n_cols;
  if (mc_max > p.MC_total) mc_max = p.MC_total;
  if (mc_max > 4) mc_max = 4;
  int R = 0, a_stages = 0, b_bufs = 2, nb_in_alloc = 0, nb_in_real = 0;
  p.MC = 0;
  const int cand[4][2] = {{128, 2}, {128, 1}, {64, 2}, {32, 2}};
  for (int mc = mc_max; mc >= 1; --mc) {
    const int alloc = mc * blocks_per_mc;
    const int real = p.nb_in < alloc ? p.nb_in : alloc;
    for (int ci = 0; ci < 4; ++ci) {
      const size_t blk = size_t(cand[ci][0]) * KBLOCK_BYTES;
      const size_t fixed_b = size_t(cand[ci][1]) * nb_out_alloc * blk + size_t(alloc - real) * blk;
      if (fixed_b >= 216 * 1024) continue;
      const int st = int((216 * 1024 - fixed_b) / (real * blk));
      if (st < 2) continue;
      if (cand[ci][0] > R) {
        R = cand[ci][0];
        b_bufs = cand[ci][1];
        a_stages = st;
        p.MC = mc;
        nb_in_alloc = alloc;
        nb_in_real = real;
      }
      break;   // first (tallest) candidate that fits for this mc
    }
  }
  if (p.MC == 0) return LGS_E_UNSUPPORTED;
  const int m_slices = (p.MC_total + p.MC - 1) / p.MC;
  if (R < 32 || a_stages < 2) return LGS_E_UNSUPPORTED;
  if (a_stages > MAX_WA_STAGES) a_stages = MAX_WA_STAGES;
  p.R = R;
  p.a_stages = a_stages;
  p.b_bufs = b_bufs;
  p.nb_in_real = nb_in_real;
  int G = 512 / (p.MC * p.n_cols);
  if (G < 1) return LGS_E_UNSUPPORTED;
  if (G > K) G = K;
  // balance the offset groups (27 offsets, G = 5 -> 6 groups of 4/5)
  const int groups = (K + G - 1) / G;
  G = (K + groups - 1) / groups;
  p.G = G;
  int cols = 32;
  while (cols < G * p.MC * p.n_cols) cols <<= 1;
  p.tmem_cols = cols;

  {
    const int rz = zero_fill_async(gw, size_t(K) * c_in * c_out * sizeof(float), stream);
    if (rz != LGS_OK) return rz;
  }
  if (n_out == 0) return LGS_OK;

  // LGS_WGRAD_WAVES (default 4): CTAs' worth of work per SM; more = shorter CTAs (finer interleaving with the training stream's
  // kernels, which cannot share an SM with a 214 KB wgrad CTA) but more red.add traffic.  Step time 10.86-10.93 ms at 2,
  // 10.77 at 3 / 4 / 8 on one box (gpurun_out/r2an)
  static const int wg_waves = getenv("LGS_WGRAD_WAVES") ? std::max(1, atoi(getenv("LGS_WGRAD_WAVES"))) : 4;
  int64_t chunks = (148 * wg_waves) / (groups * n_splits * m_slices);   // one CTA resident per SM at a time (TMEM)
  const int64_t max_chunks = cdiv(n_out, int64_t(R) * 4);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  p.rows_per_chunk = cdiv(cdiv(n_out, chunks), R) * R;
  chunks = cdiv(n_out, p.rows_per_chunk);
  if (!getenv("LGS_WGRAD_NO_BALANCE")) {
    // One CTA per SM is resident (TMEM): a grid a few CTAs over a multiple of 148 pays a whole extra wave (76 x 2 = 152
    // CTAs for the 32 -> 32 layers of a 38 K-row map).  Search the chunk count for the cheapest waves x (tiles + 2 for the
    // red.add epilogue) and keep the choice above unless another one is clearly better.
    const int64_t per = int64_t(groups) * n_splits * m_slices;
    auto cost_of = [&](int64_t c, int64_t& rows, int64_t& act) {
      rows = cdiv(cdiv(n_out, c), R) * R;
      act = cdiv(n_out, rows);
      return double(cdiv(act * per, 148)) * double(rows / R + 2);
    };
    int64_t rows = 0, act = 0;
    double best = cost_of(chunks, rows, act);
    int64_t best_rows = p.rows_per_chunk, best_chunks = chunks;
    const int64_t c_hi = std::min<int64_t>(max_chunks, cdiv(148 * 2 * wg_waves, per));   // chunks stay >= 4 tiles tall
    for (int64_t c = 1; c <= c_hi; ++c) {
      const double cost = cost_of(c, rows, act);
      if (cost < best * 0.97) best = cost, best_rows = rows, best_chunks = act;
    }
    p.rows_per_chunk = best_rows;
    chunks = best_chunks;
  }

  // kernel-side sizes use nb_in (gathered) for the producer loop but the stage stride must cover what the MMA reads
  WParams q = p;
  const int blk_bytes = R * KBLOCK_BYTES;
  // encode stage strides through nb_in/nb_out: the kernel computes a_stage_bytes = nb_in * blk_bytes, so pass the
  // allocation counts and keep the number of gathered blocks implicit in in_row_bytes (col_ok masks the rest)
  q.nb_in = nb_in_alloc;
  q.nb_out = nb_out_alloc;
  const size_t smem_bytes = size_t(a_stages * nb_in_real + (nb_in_alloc - nb_in_real) + b_bufs * nb_out_alloc) * blk_bytes +
                            (2 * MAX_WA_STAGES + 8) * 8 + 1024;

  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {cuuint64_t(c_out), cuuint64_t(n_out)};
  const cuuint64_t gstride[1] = {cuuint64_t(out_row_bytes)};
  const cuuint32_t box[2] = {cuuint32_t(KBLOCK_BYTES / es), cuuint32_t(R)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = encode(&tmap, dtype == LGS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                             2, const_cast<void*>(gout), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             dtype == LGS_BF16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(LGS_E_CUDA, "cuTensorMapEncodeTiled (wgrad) failed (%d)", int(cr));

  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) return fail(LGS_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));

  const dim3 grid{unsigned(chunks), unsigned(groups), unsigned(n_splits * m_slices)};
  if (dtype == LGS_BF16) {
    LGS_LAUNCH_PDL(wgrad_tc_kernel<true>, grid, THREADS, smem_bytes, stream, tmap, q);
  } else {
    LGS_LAUNCH_PDL(wgrad_tc_kernel<false>, grid, THREADS, smem_bytes, stream, tmap, q);
  }
  return LGS_OK;
}

}  // namespace lgs
