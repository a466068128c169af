// conv_nb: 3x3x3 same-map sparse convolution on the large coordinate maps, bf16x3 products (LGS_ALGO_BX3 numerics), with the
// gather served from a shared-memory NEIGHBOURHOOD CACHE instead of one cp.async per (row, offset) pair.
//
// conv_bx3.cu is bound by the number of (row, offset) slots it copies: 27 x 128 rows x 128 B per channel block and tile,
// ~45 % of them zero-fill, a warp-level cp.async costing ~17 cycles whatever its lanes do (profiles/r2_bx3_kernel_variants.txt)
// — 870 cycles per (tile, offset, channel block) stage where the three MMAs need 288.  The neighbourhood plan (nbplan.cu)
// groups output rows into spatially compact supertiles whose 27-neighbourhoods overlap almost entirely: ~415 unique input
// rows per 256 output rows instead of 3 800 gathered ones.  This kernel therefore
//   - loads the supertile's unique rows ONCE per 32-channel block, splits them into bf16 hi / lo THERE (once per unique row and
//     channel block instead of once per (row, offset): 16x less conversion work — the first version split in the row threads and
//     was bound by the ALU pipe, 52 % active against 40 % tensor pipe, profiles/r2_conv_nb_v1_ncu.txt) and stores them in the
//     shared-memory cache as [32 x hi | 32 x lo] = 128 bytes per row, chunk c at position (c + colour) mod 8,
//   - builds every offset's A operand from that cache: row thread <-> output row <-> TMEM lane reads its neighbour's 128 bytes
//     by LOCAL index (plan `loc`: 12-bit index + 3-bit colour) with eight 16-byte loads and tcgen05.st's them behind the
//     accumulators (TS-form MMAs as in conv_bx3.cu) — no arithmetic on the data.  The plan orders the slots of a supertile so
//     that the eight lanes of a quarter-warp read rows of eight different colours for every offset, i.e. eight different bank
//     groups in every step (nbplan.cu, "Colours"),
//   - runs TWO CTAs per SM (256 threads, <= 113 KB of shared memory, 256 TMEM columns each): one CTA's cache refill between
//     channel blocks hides behind the other CTA's MMAs.
// Loop nest per CTA: channel block -> offset -> row tile (TM = 2); one weight block [n_tile x (hi|lo)] per (channel block,
// offset) by TMA, shared by both row tiles.  Output rows are written to their own positions (`order`), so tensors keep
// MinkowskiEngine's row order; every output row accumulates its 27 x c_in products in the same sequence as in conv_bx3.cu.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "tcgen05.cuh"

namespace lgs {
namespace nb {
using namespace tc;

constexpr int BM = 128;
constexpr int ROW_BYTES = 128;                    // fp32 bytes of one channel block (32 channels) per cached row
constexpr int THREADS = 256;                      // warps 0-3 row threads (fill, A operand, epilogue), 4 and 6 MMA, 5 weight TMA, 7 fill
constexpr int FILL_THREADS = 160;                 // warps 0-3 and 7 refill the cache together between channel blocks
constexpr int TM = 2;                             // row tiles per supertile (nb_geometry's tm)
constexpr int FILL_UNROLL = 6;
constexpr int MAX_B = 4, MAX_TA = 4;
constexpr int TA_COLS = 32;

struct Params {
  const uint8_t* in;
  const uint8_t* in2;
  int32_t row_bytes, row_bytes2;
  int32_t nkb1, num_kb;
  int32_t K, c_out;
  const int32_t* order;      // [S][RS] output row of every slot, -1 = padding
  const int32_t* ucount;     // [S]
  const int32_t* uniq;       // [S][umax] input rows of the supertile | colour << 28
  const uint16_t* loc;       // [S][K][RS] local index of slot's neighbour at offset k (0xFFF = none) | colour << 12
  int32_t RS, rt, umax;
  int32_t reverse_k;
  int32_t kb_per_split, kb_splits, k_per_split;   // blockIdx.z = offset split * kb_splits + channel-block split (small maps)
  const float* bias;
  const float* addend;       // [n_out][c_out] added to the result (dgrad + residual-path gradient), or NULL
  float* out;
  int32_t n_tile, b_stages, ta_stages, ta_col0, tmem_cols, b_stage_bytes;
  double* stats;
};

__device__ __forceinline__ void umma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void split2(float e0, float e1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(e0, e1);
  const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
  const float r0 = e0 - __uint_as_float(hb << 16);
  const float r1 = e1 - __uint_as_float(hb & 0xFFFF0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
  hi = hb;
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void lds128(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void fill_bar() { asm volatile("bar.sync 1, %0;" ::"n"(FILL_THREADS) : "memory"); }

// Cache refill of one channel block by the 192 fill threads: unique row u, 4-channel group c4 -> 16 bytes of fp32 from global
// memory -> bf16 hi / lo pairs -> two 8-byte stores at the row's colour-rotated chunk positions.
__device__ __forceinline__ void fill_cache(const uint8_t* in, const uint8_t* in2, int row_bytes, int row_bytes2, int nkb1, int kb, int ft,
                                           int U, const int32_t* s_uniq, uint32_t cache_base) {
  const int c4 = ft & 7, rsub = ft >> 3;             // 8 threads cover one 128-byte row segment; 24 rows per pass
  const bool second = kb >= nkb1;
  const int kbl = second ? kb - nkb1 : kb;
  const int rb = second ? row_bytes2 : row_bytes;
  const bool col_ok = kbl * ROW_BYTES + c4 * 16 < rb;
  const uint8_t* src = (second ? in2 : in) + kbl * ROW_BYTES + c4 * 16;
  constexpr int STEP = FILL_THREADS / 8;
#pragma unroll 1
  for (int u0 = rsub; u0 < U; u0 += STEP * FILL_UNROLL) {
    float4 v[FILL_UNROLL];
    int32_t e[FILL_UNROLL];
#pragma unroll
    for (int j = 0; j < FILL_UNROLL; ++j) {
      const int u = u0 + j * STEP;
      e[j] = u < U ? s_uniq[u] : -1;
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e[j] >= 0 && col_ok) v[j] = __ldg(reinterpret_cast<const float4*>(src + size_t(e[j] & 0x0FFFFFFF) * rb));
    }
#pragma unroll
    for (int j = 0; j < FILL_UNROLL; ++j) {
      if (e[j] < 0) continue;
      const int u = u0 + j * STEP;
      const uint32_t g = uint32_t(e[j]) >> 28;
      uint32_t h0, l0, h1, l1;
      split2(v[j].x, v[j].y, h0, l0);
      split2(v[j].z, v[j].w, h1, l1);
      const uint32_t row = cache_base + uint32_t(u) * ROW_BYTES + uint32_t((c4 & 1) << 3);
      sts64(row + (((uint32_t(c4 >> 1) + g) & 7u) << 4), h0, h1);
      sts64(row + (((uint32_t(4 + (c4 >> 1)) + g) & 7u) << 4), l0, l1);
    }
  }
}

// Epilogue of one row warp, out of line so that its register needs (three 32-word arrays for the fused BatchNorm sums) do not
// spill loop-invariant values of the main loops: TMEM accumulators -> (+ bias) -> output rows at their own positions.
__device__ __noinline__ void epilogue_rows(uint32_t tmem_lane_base, int n_tile, int ncols, int c_out, int n0, const float* bias,
                                           const float* addend, float* out, const int32_t* order_s, int rt, int r, bool want_stats,
                                           float* s_stats, bool partial) {
  const int lane = threadIdx.x & 31;
  for (int t = 0; t < TM; ++t) {
    const int32_t o = r < rt ? __ldg(order_s + t * rt + r) : -1;
    const bool row_ok = o >= 0;
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_lane_base + uint32_t(t * n_tile + c0), v);
      if (bias) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj)
          if (c0 + jj < ncols) v[jj] = __float_as_uint(__uint_as_float(v[jj]) + __ldg(bias + n0 + c0 + jj));
      }
      if (addend && row_ok) {                       // out = conv + addend (this CTA is the one that adds it: blockIdx.z == 0)
        const float* arow = addend + size_t(o) * c_out + n0 + c0;
#pragma unroll
        for (int jj = 0; jj < 32; jj += 4) {
          if (c0 + jj + 3 < ncols) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(arow + jj));
            v[jj] = __float_as_uint(__uint_as_float(v[jj]) + a.x);
            v[jj + 1] = __float_as_uint(__uint_as_float(v[jj + 1]) + a.y);
            v[jj + 2] = __float_as_uint(__uint_as_float(v[jj + 2]) + a.z);
            v[jj + 3] = __float_as_uint(__uint_as_float(v[jj + 3]) + a.w);
          } else {
            for (int e = jj; e < jj + 4; ++e)
              if (c0 + e < ncols) v[e] = __float_as_uint(__uint_as_float(v[e]) + __ldg(arow + e));
          }
        }
      }
      if (row_ok) {
        float* orow = out + size_t(o) * c_out + n0 + c0;
#pragma unroll
        for (int jj = 0; jj < 32; jj += 4) {
          if (c0 + jj + 3 < ncols) {
            if (partial) {                            // this CTA holds a partial sum over channel blocks / offsets
              red_add_v4(orow + jj, __uint_as_float(v[jj]), __uint_as_float(v[jj + 1]), __uint_as_float(v[jj + 2]), __uint_as_float(v[jj + 3]));
            } else {
              float4 x;
              x.x = __uint_as_float(v[jj]);
              x.y = __uint_as_float(v[jj + 1]);
              x.z = __uint_as_float(v[jj + 2]);
              x.w = __uint_as_float(v[jj + 3]);
              *reinterpret_cast<float4*>(orow + jj) = x;
            }
          } else {
            for (int e = jj; e < jj + 4; ++e)
              if (c0 + e < ncols) {
                if (partial) atomicAdd(orow + e, __uint_as_float(v[e]));
                else orow[e] = __uint_as_float(v[e]);
              }
          }
        }
      }
      if (want_stats) {
        // BatchNorm statistics of the output from the accumulator registers (as in conv_bx3.cu): a butterfly over the
        // warp's 32 rows leaves lane L with the sums of column c0 + L
        float s1[32], s2[32];
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const float x = row_ok ? __uint_as_float(v[jj]) : 0.f;
          s1[jj] = x;
          s2[jj] = x * x;
        }
#pragma unroll
        for (int w2 = 16; w2 >= 1; w2 >>= 1) {
          const bool upper = (lane & w2) != 0;
#pragma unroll
          for (int jj = 0; jj < w2; ++jj) {
            const float a1 = upper ? s1[jj] : s1[jj + w2], a2 = upper ? s2[jj] : s2[jj + w2];
            const float k1 = upper ? s1[jj + w2] : s1[jj], k2 = upper ? s2[jj + w2] : s2[jj];
            s1[jj] = k1 + __shfl_xor_sync(0xffffffffu, a1, w2);
            s2[jj] = k2 + __shfl_xor_sync(0xffffffffu, a2, w2);
          }
        }
        if (c0 + lane < ncols) {
          atomicAdd(s_stats + c0 + lane, s1[0]);
          atomicAdd(s_stats + 256 + c0 + lane, s2[0]);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(THREADS, 2) conv_nb_kernel(const __grid_constant__ CUtensorMap tmap_w, const Params p) {
  pdl_grid_sync();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int SB = p.b_stages, K = p.K, umax = p.umax;
  // this CTA's share of the reduction: channel blocks [kb0, kb1) x kernel offsets [k0, k1) (everything on large maps)
  const int zs = blockIdx.z / p.kb_splits, zb = blockIdx.z - zs * p.kb_splits;
  const int kb0 = zb * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
  const int k0 = zs * p.k_per_split, k1 = min(K, k0 + p.k_per_split);
  const bool partial = gridDim.z > 1;
  const int NBUF = p.ta_stages / TM;                                // TMEM A buffers per row tile (tile t owns stages t + TM i)
  const uint32_t b_bytes = uint32_t(p.b_stage_bytes);
  uint8_t* b_ring = smem;                                           // 1024-aligned (SWIZZLE_128B TMA destination)
  uint8_t* cache = b_ring + size_t(SB) * b_bytes;                   // [(umax + 1)][128 B], row umax = zeros
  int32_t* s_uniq = reinterpret_cast<int32_t*>(cache + size_t(umax + 1) * ROW_BYTES);
  float* s_stats = reinterpret_cast<float*>(s_uniq + umax);         // [2][256]
  uint64_t* b_full = reinterpret_cast<uint64_t*>(s_stats + 512);
  uint64_t* b_empty = b_full + MAX_B;
  uint64_t* ta_full = b_empty + MAX_B;
  uint64_t* ta_empty = ta_full + MAX_TA;
  uint64_t* acc_bar = ta_empty + MAX_TA;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t s = blockIdx.x;
  const int n0 = blockIdx.y * p.n_tile;
  const int U = min(__ldg(p.ucount + s), umax);

  if (tid == 0) {
    for (int i = 0; i < MAX_B; ++i) {
      mbar_init(b_full + i, 1);
      mbar_init(b_empty + i, TM);                    // one commit per MMA issuer warp
    }
    for (int i = 0; i < MAX_TA; ++i) {
      mbar_init(ta_full + i, 128);
      mbar_init(ta_empty + i, 1);
    }
    mbar_init(acc_bar, TM);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(uint32_t(p.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 5 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
  for (int u = tid; u < U; u += THREADS) s_uniq[u] = __ldg(p.uniq + s * umax + u);
  if (tid < 32) reinterpret_cast<float*>(cache + size_t(umax) * ROW_BYTES)[tid] = 0.f;     // the "no neighbour" row
  if (p.stats)
    for (int i = tid; i < 512; i += THREADS) s_stats[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < 4) {
    // =================================== row threads: cache row -> TMEM (no arithmetic on the data) ===================
    const int r = warp * 32 + lane;                 // row of the tile = TMEM lane (warp w may touch lanes [32w, 32w + 32))
    const uint32_t lane_addr = uint32_t(warp * 32) << 16;
    const uint32_t cache_base = smem_u32(cache);
    const bool row_in_tile = r < p.rt;
    const uint32_t none = 0xFFFu | (uint32_t(lane & 7) << 12);
    // index stream of this row slot: loc[kk][t * rt + r], kk ascending (forward) or descending (dgrad), fetched one offset ahead
    const int64_t kstride = p.reverse_k ? -int64_t(p.RS) : int64_t(p.RS);
    const uint16_t* lp0 = p.loc + s * int64_t(K) * p.RS + (p.reverse_k ? int64_t(K - 1 - k0) : int64_t(k0)) * p.RS + min(r, p.rt - 1);
    const int rt = p.rt;
    int buf = 0;
    uint32_t pht = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      if (kb > kb0) fill_bar();                     // every row thread is done reading the previous channel block
      fill_cache(p.in, p.in2, p.row_bytes, p.row_bytes2, p.nkb1, kb, tid, U, s_uniq, cache_base);
      const uint16_t* lp = lp0;
      uint32_t nx[TM];
#pragma unroll
      for (int t = 0; t < TM; ++t) nx[t] = row_in_tile ? uint32_t(__ldg(lp + t * rt)) : none;
      fill_bar();                                   // this channel block of every unique row is in the cache
#pragma unroll 1
      for (int ki = k0; ki < k1; ++ki) {
        uint32_t cur[TM];
#pragma unroll
        for (int t = 0; t < TM; ++t) cur[t] = nx[t];
        if (ki + 1 < k1) {
          lp += kstride;
#pragma unroll
          for (int t = 0; t < TM; ++t) nx[t] = row_in_tile ? uint32_t(__ldg(lp + t * rt)) : none;
        }
#pragma unroll
        for (int t = 0; t < TM; ++t) {
          const uint32_t j = min(cur[t] & 0xFFFu, uint32_t(umax));
          const uint32_t g = cur[t] >> 12;
          const uint32_t row_addr = cache_base + j * ROW_BYTES;
          uint32_t w[32];                           // [0,16): hi pairs, [16,32): lo pairs
#pragma unroll
          for (int i = 0; i < 8; ++i)
            lds128(row_addr + (((uint32_t(i) + g) & 7u) << 4), w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
          const int ts = t + buf * TM;
          mbar_wait(ta_empty + ts, pht ^ 1);        // MMAs that read this TMEM stage last time are done
          tc_fence_after();
          tmem_st32(tmem_base + lane_addr + uint32_t(p.ta_col0 + ts * TA_COLS), w);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(ta_full + ts);
        }
        if (++buf == NBUF) {
          buf = 0;
          pht ^= 1;
        }
      }
    }

    // =================================== epilogue ===================================
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    epilogue_rows(tmem_base + lane_addr, p.n_tile, min(p.n_tile, p.c_out - n0), p.c_out, n0, blockIdx.z == 0 ? p.bias : nullptr,
                  blockIdx.z == 0 ? p.addend : nullptr, p.out, p.order + s * p.RS, rt, r, p.stats != nullptr, s_stats, partial);
    tc_fence_before();
  } else if (warp == 4 || warp == 6) {
    // =================================== MMA issuers: warp 4 <-> row tile 0, warp 6 <-> row tile 1 ================
    // (a single issuing warp needed ~90 dependent instructions per 288-cycle stage and was ~80 % busy: profiles/r2_conv_nb_v2_ncu.txt)
    const int t = warp == 4 ? 0 : 1;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(p.n_tile >> 3) << 17) | (uint32_t(BM >> 4) << 24);
    const uint32_t desc_hi = uint32_t(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
    const uint32_t b_lo_base = ((smem_u32(b_ring) >> 4) & 0x3FFF) | (1u << 16);
    const uint32_t b_lo_step = b_bytes >> 4;
    const uint32_t d_addr = tmem_base + uint32_t(t * p.n_tile);
    const uint32_t a_base = tmem_base + uint32_t(p.ta_col0 + t * TA_COLS);
    int sb = 0, buf = 0;
    uint32_t phb = 0, phta = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      const bool second = kb >= p.nkb1;
      const int valid = min(ROW_BYTES, second ? p.row_bytes2 - (kb - p.nkb1) * ROW_BYTES : p.row_bytes - kb * ROW_BYTES);
      const bool two = valid > 64;                   // 16 channels (64 bytes of fp32) per instruction, 1 or 2 steps
#pragma unroll 1
      for (int ki = k0; ki < k1; ++ki) {
        const int ts = t + buf * TM;
        const uint32_t b_lo32 = b_lo_base + uint32_t(sb) * b_lo_step;
        const uint32_t a_tm = a_base + uint32_t(buf * TM * TA_COLS);
        mbar_wait(b_full + sb, phb);
        mbar_wait(ta_full + ts, phta);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t b_hi0 = desc_from(b_lo32, desc_hi), b_lo0 = desc_from(b_lo32 + 4, desc_hi);
          umma_ts_f16(d_addr, a_tm, b_hi0, idesc, (ki > k0 || kb > kb0) ? 1u : 0u);
          umma_ts_f16(d_addr, a_tm + 16, b_hi0, idesc, 1u);
          umma_ts_f16(d_addr, a_tm, b_lo0, idesc, 1u);
          if (two) {
            const uint64_t b_hi1 = desc_from(b_lo32 + 2, desc_hi), b_lo1 = desc_from(b_lo32 + 6, desc_hi);
            umma_ts_f16(d_addr, a_tm + 8, b_hi1, idesc, 1u);
            umma_ts_f16(d_addr, a_tm + 24, b_hi1, idesc, 1u);
            umma_ts_f16(d_addr, a_tm + 8, b_lo1, idesc, 1u);
          }
          umma_commit(ta_empty + ts);
          umma_commit(b_empty + sb);
        }
        __syncwarp();
        if (++buf == NBUF) {
          buf = 0;
          phta ^= 1;
        }
        if (++sb == SB) {
          sb = 0;
          phb ^= 1;
        }
      }
    }
    if (elect_one()) umma_commit(acc_bar);
    __syncwarp();
  } else if (warp == 5) {
    // =================================== weight TMA producer ===================================
    const uint32_t b_ring_base = smem_u32(b_ring);
    int sb = 0;
    uint32_t phb = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      for (int ki = k0; ki < k1; ++ki) {
        mbar_wait(b_empty + sb, phb ^ 1);
        if (elect_one()) {
          mbar_expect_tx(b_full + sb, b_bytes);
          tma_load_2d(b_ring_base + uint32_t(sb) * b_bytes, &tmap_w, b_full + sb, kb * 64, ki * p.c_out + n0);
        }
        __syncwarp();
        if (++sb == SB) {
          sb = 0;
          phb ^= 1;
        }
      }
    }
  } else {
    // =================================== cache fill helper (warp 7; the row threads fill with it) ==========
    const uint32_t cache_base = smem_u32(cache);
    for (int kb = kb0; kb < kb1; ++kb) {
      if (kb > kb0) fill_bar();
      fill_cache(p.in, p.in2, p.row_bytes, p.row_bytes2, p.nkb1, kb, tid - 96, U, s_uniq, cache_base);
      fill_bar();
    }
  }

  __syncthreads();
  if (p.stats) {
    const int ncols = min(p.n_tile, p.c_out - n0);
    double* dst = p.stats + size_t(blockIdx.x & 7) * 2 * p.c_out + n0;
    for (int i = tid; i < ncols; i += THREADS) {
      atomicAdd(dst + i, double(s_stats[i]));
      atomicAdd(dst + p.c_out + i, double(s_stats[256 + i]));
    }
  }
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(p.tmem_cols))
                 : "memory");
  }
}

}  // namespace nb

int conv_bx3_shape_ok(int c_in, int c_out);

// 1 if conv_fwd_nb takes this layer on a map with a neighbourhood plan
int conv_nb_shape_ok(int c_in, int c_in2, int c_out, int K) {
  if (K != 27 || !conv_bx3_shape_ok(c_in, c_out)) return 0;
  if (c_in2 && (c_in % 32 != 0 || c_in2 % 4 != 0 || c_in2 < 4)) return 0;
  if (c_out % 16 != 0 || c_out > 1024) return 0;      // 1024: the BatchNorm accumulator scratch holds 8 x 2 x 1024 doubles
  return 1;
}

// in2 / c_in2: optional second gather source as in conv_fwd_bx3.  d_plan: lgs_nbplan_build output for (n_out, K).
int conv_fwd_nb(const void* in, int c_in, const void* in2, int c_in2, const void* w, int K, int c_out, const void* d_plan,
                int64_t n_out, int reverse_k, const float* bias, const float* addend, float* out, double* stats, cudaStream_t stream) {
  using namespace nb;
  if (!conv_nb_shape_ok(c_in, in2 ? c_in2 : 0, c_out, K)) return LGS_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(in2) & 15) || (reinterpret_cast<uintptr_t>(w) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(d_plan) & 15) || (reinterpret_cast<uintptr_t>(addend) & 15))
    return LGS_E_UNSUPPORTED;
  if (n_out == 0) return LGS_OK;
  EncodeTiledFn encode = get_encode();
  if (!encode) return fail(LGS_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(conv_nb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
  });
  if (attr_err != cudaSuccess) return fail(LGS_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));

  const NbGeom g = nb_geometry(n_out, K);
  const int32_t* plan = static_cast<const int32_t*>(d_plan);
  Params q;
  q.in = static_cast<const uint8_t*>(in);
  q.in2 = static_cast<const uint8_t*>(in2);
  q.row_bytes = c_in * 4;
  q.row_bytes2 = in2 ? c_in2 * 4 : 0;
  q.nkb1 = (q.row_bytes + ROW_BYTES - 1) / ROW_BYTES;
  q.num_kb = q.nkb1 + (in2 ? (q.row_bytes2 + ROW_BYTES - 1) / ROW_BYTES : 0);
  q.K = K;
  q.c_out = c_out;
  q.order = plan + g.off_order;
  q.ucount = plan + g.off_ucount;
  q.uniq = plan + g.off_uniq;
  q.loc = reinterpret_cast<const uint16_t*>(plan + g.off_loc);
  if (g.tm != TM) return fail(LGS_E_INVALID, "conv_fwd_nb: plan geometry tm %d", g.tm);
  q.RS = g.RS, q.rt = g.rt, q.umax = g.umax;
  q.reverse_k = reverse_k;
  q.bias = bias;
  q.addend = addend;
  q.out = out;
  q.stats = stats;
  // output channels per CTA: TM accumulators + >= 2 split-A stages within 256 TMEM columns (two CTAs share an SM's 512)
  const int max_nt = ((256 - 2 * TA_COLS) / g.tm) & ~15;          // 96 for TM = 2
  const int ns = (c_out + max_nt - 1) / max_nt;
  const int nt = ((((c_out + ns - 1) / ns) + 15) / 16) * 16;
  q.n_tile = nt;
  q.b_stage_bytes = nt * ROW_BYTES;
  q.ta_col0 = ((g.tm * nt + 31) / 32) * 32;
  q.ta_stages = std::min(MAX_TA, (256 - q.ta_col0) / TA_COLS) / TM * TM;      // whole buffers per row tile
  int cols = 32;
  while (cols < q.ta_col0 + q.ta_stages * TA_COLS) cols <<= 1;
  q.tmem_cols = cols;
  const size_t fixed = size_t(g.umax + 1) * ROW_BYTES + size_t(g.umax) * 4 + 2048 + (2 * MAX_B + 2 * MAX_TA + 1) * 8 + 16 + 1024;
  int sb = 3;
  while (sb > 2 && fixed + size_t(sb) * q.b_stage_bytes > 113 * 1024) --sb;
  if (fixed + size_t(sb) * q.b_stage_bytes > 113 * 1024) return LGS_E_UNSUPPORTED;
  q.b_stages = sb;
  const size_t smem = fixed + size_t(sb) * q.b_stage_bytes;

  // Small maps (fewer supertiles than SMs): the reduction over channel blocks and kernel offsets is split over CTAs
  // (blockIdx.z) until about two CTAs per SM are busy; partial sums meet through red.global.add on a zeroed output.
  // Channel blocks first (every split fills only its own block of the cache), then offsets.  Never when the BatchNorm
  // statistics are wanted from the epilogue (they need complete sums).
  const int n_slices = (c_out + nt - 1) / nt;
  int kb_splits = 1, k_splits = 1;
  const int64_t base_ctas = g.S * n_slices;
  if (!stats && base_ctas < 148 && !nb_no_split()) {
    // cheapest (waves of nb_target_ctas CTAs) x (fixed cost of a CTA + its stages): a second wave costs a whole CTA time
    const int64_t slots = nb_target_ctas();
    double best = 1e30;
    for (int kbs = 1; kbs <= q.num_kb; ++kbs) {
      const int kbp = (q.num_kb + kbs - 1) / kbs;
      if ((q.num_kb + kbp - 1) / kbp != kbs) continue;
      for (int ks = 1; ks <= 9; ++ks) {
        const int kp = (K + ks - 1) / ks;
        if ((K + kp - 1) / kp != ks) continue;
        const int64_t ctas = base_ctas * kbs * ks;
        const double cost = double(cdiv(ctas, slots)) * (6000.0 + 3000.0 * kbp + 800.0 * kbp * kp * TM) + 40.0 * kbs * ks;
        if (cost < best) best = cost, kb_splits = kbs, k_splits = ks;
      }
    }
  }
  q.kb_per_split = (q.num_kb + kb_splits - 1) / kb_splits;
  q.kb_splits = (q.num_kb + q.kb_per_split - 1) / q.kb_per_split;
  q.k_per_split = (K + k_splits - 1) / k_splits;
  k_splits = (K + q.k_per_split - 1) / q.k_per_split;
  const unsigned gz = unsigned(q.kb_splits * k_splits);
  if (gz > 1) {
    const int rz = zero_fill_async(out, size_t(n_out) * c_out * sizeof(float), stream);
    if (rz != LGS_OK) return rz;
  }

  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {cuuint64_t(q.num_kb) * 64, cuuint64_t(K) * cuuint64_t(c_out)};
  const cuuint64_t gstride[1] = {cuuint64_t(q.num_kb) * 128};
  const cuuint32_t box[2] = {64, cuuint32_t(nt)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(LGS_E_CUDA, "cuTensorMapEncodeTiled (nb) failed (%d) c_in=%d c_out=%d K=%d", int(cr), c_in + c_in2, c_out, K);
  const dim3 grid{unsigned(g.S), unsigned(n_slices), gz};
  LGS_LAUNCH_PDL(conv_nb_kernel, grid, THREADS, smem, stream, tmap, q);
  return LGS_OK;
}

}  // namespace lgs
