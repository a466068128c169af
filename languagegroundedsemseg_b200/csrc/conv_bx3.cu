// LGS_ALGO_BX3: output-stationary sparse convolution for fp32 features with error-compensated BF16 products
// ("bf16x3": a = a_hi + a_lo, w = w_hi + w_lo as bf16 pairs; a_hi*w_hi + a_lo*w_hi + a_hi*w_lo, fp32 accumulation in
// TMEM).  Products are exact to ~2^-16 relative (dropped: a_lo*w_lo and the bits below the 16 kept), features and
// accumulators keep fp32 range and storage.  Versus the 3xTF32 path of conv_tc.cu (2^-21) this halves the tensor-core
// time and the weight traffic: kind::f16 MMAs retire 16 K-elements per instruction where kind::tf32 retires 8, and a
// hi/lo weight pair is 4 bytes instead of 8.  With three TF32 MMAs per K step the old kernel was co-limited by the tensor
// pipe (576 of ~1100 cycles per 16 KB stage), the L2->SM gather (607) and shared-memory bandwidth (531); here the gather is
// the only stream near its limit (tensor 288, shared memory ~390 cycles per stage).
//
// Structure (one CTA = TM <= 4 row tiles of rt <= 128 output rows x n_tile output channels [x a range of kernel offsets]):
//   8 producer warps   gather neighbour rows in[table[k][o]] with 16-byte cp.async (zero-fill for missing neighbours)
//                      into a ring of 128B-swizzled [128 rows][128 B] stages; later the epilogue (TMEM -> global)
//   4 splitter warps   thread <-> row <-> TMEM lane: read the row's 32 floats, split into bf16 hi / lo (round to nearest),
//                      tcgen05.st both (16 + 16 packed columns) into a TMEM ring behind the accumulators
//   1 TMA warp         weight blocks [n_tile][hi x32 | lo x32] (bf16, 128 B per output channel) -> shared-memory ring
//   1 MMA warp         TS-form tcgen05.mma kind::f16 (A from TMEM, B from shared memory), 3 MMAs per 16 channels,
//                      TM accumulators in TMEM; every weight block is loaded once per CTA and feeds TM row tiles
// Two gather sources (in | in2) serve `cat`-free decoder convolutions: channel blocks [0, nkb1) come from `in`, the rest
// from `in2` (models/res16unet.py:237-268 concatenates the up-sampled tensor with the encoder skip before block5..8).
// Small coordinate maps split the kernel offsets (blockIdx.z) and output channels (blockIdx.y) over CTAs; offset splits
// meet through red.global.add on a zeroed output.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <string>

#include "common.cuh"
#include "tcgen05.cuh"

namespace lgs {
namespace bx3 {
using namespace tc;

constexpr int BM = 128;
constexpr int KBLOCK_BYTES = 128;                 // fp32 bytes of one channel block (32 channels) per gathered row
constexpr int A_BYTES = BM * KBLOCK_BYTES;        // 16 KB per gather stage
constexpr int MAX_A = 12, MAX_B = 4, MAX_TA = 8;
constexpr int PROD_WARPS = 8;
constexpr int THREADS = (PROD_WARPS + 2 + 4) * 32;  // producers + MMA + TMA + splitters = 448
constexpr int TA_COLS = 32;                       // TMEM columns of one split A stage: 16 (hi, packed bf16x2) + 16 (lo)

struct Params {
  const uint8_t* in;
  const uint8_t* in2;
  int32_t row_bytes, row_bytes2;
  int32_t nkb1, num_kb;
  int32_t K, c_out;
  const int32_t* table;
  int64_t n_out;
  int32_t reverse_k;
  const float* bias;
  float* out;
  int32_t n_tile, TM, rt;
  int32_t a_stages, b_stages, ta_stages, ta_col0, tmem_cols, b_stage_bytes;
  int32_t k_per_split, k_splits;
  double* stats;            // optional BatchNorm accumulators [8][2][c_out] (column sums / sums of squares of the output), or null
};

__device__ __forceinline__ void umma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// (hi, lo) bf16x2 words of two consecutive floats; element e0 in the low half
__device__ __forceinline__ void split2(float e0, float e1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(e0, e1);
  const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
  const float r0 = e0 - __uint_as_float(hb << 16);
  const float r1 = e1 - __uint_as_float(hb & 0xFFFF0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
  hi = hb;
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Variants of this pipeline measured on the B200 and NOT kept (profiles/r2_bx3_kernel_variants.txt): 8 splitter warps in
// two groups (292 us vs 282 on the dominant 96 -> 96 launch: the splitters are not the limiter), 4 producer warps with 48 KB
// stages (297 us; in isolation that producer shape gathers 20 % faster, scripts/micro/gather_bulk_probe.cu, but the overlap
// with the splitters / MMAs got worse), one mbarrier arrival per warp with cp.async commit groups (415 us: two stages of
// copies in flight per warp are too few), compacting the present rows before the copies (438 us: a dependent shared-memory
// read in front of every cp.async).  A warp-level cp.async costs ~16-18 cycles whether its lanes copy, zero-fill or are
// predicated off, so the gather stream is bounded by (row, offset) SLOTS, not by bytes: ~500 cycles per 16 KB stage alone.
__global__ void __launch_bounds__(THREADS, 1) conv_bx3_kernel(const __grid_constant__ CUtensorMap tmap_w, const Params p) {
  pdl_grid_sync();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int SA = p.a_stages, SB = p.b_stages, TM = p.TM, STA = p.ta_stages;
  const uint32_t b_bytes = uint32_t(p.b_stage_bytes);
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + size_t(SA) * A_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(b_ring + size_t(SB) * b_bytes);
  uint64_t* a_empty = a_full + MAX_A;
  uint64_t* b_full = a_empty + MAX_A;
  uint64_t* b_empty = b_full + MAX_B;
  uint64_t* ta_full = b_empty + MAX_B;
  uint64_t* ta_empty = ta_full + MAX_TA;
  uint64_t* acc_bar = ta_empty + MAX_TA;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_bar + 1);
  float* s_stats = reinterpret_cast<float*>(tmem_ptr_smem + 2);      // [2][256] per-CTA column sums / sums of squares

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = int64_t(blockIdx.x) * (int64_t(TM) * p.rt);
  const int n0 = blockIdx.y * p.n_tile;
  const int kbeg = blockIdx.z * p.k_per_split;
  const int kcnt = min(p.k_per_split, p.K - kbeg);
  const bool partial = p.k_splits > 1;
  const int K = p.K, num_kb = p.num_kb;

  if (tid == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(a_full + s, PROD_WARPS * 32);
      mbar_init(a_empty + s, 128);
    }
    for (int s = 0; s < MAX_TA; ++s) {
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
  if (warp == PROD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(uint32_t(p.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == PROD_WARPS + 1 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
  if (p.stats)
    for (int i = tid; i < 512; i += THREADS) s_stats[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < PROD_WARPS) {
    {
      // =================================== gather producers ===================================
      const int chunk = tid & 7, rbase = tid >> 3;          // 8 lanes cover one 128-byte row segment; 32 rows per pass
      uint32_t dst_off[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = rbase + 32 * i;
        dst_off[i] = uint32_t(r * KBLOCK_BYTES + ((chunk ^ (r & 7)) << 4));
      }
      const uint32_t a_ring_base = smem_u32(a_ring);
      auto load_idx = [&](int k, int32_t (&v)[4][4]) {
        const int32_t* trow = p.table ? p.table + int64_t(p.reverse_k ? K - 1 - k : k) * p.n_out : nullptr;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = rbase + 32 * i;
            const int64_t o = m0 + int64_t(t) * p.rt + r;
            v[t][i] = (t < TM && r < p.rt && o < p.n_out) ? (trow ? ldg_nc_ordered(trow + o) : int32_t(o)) : -1;
          }
        }
      };
      int32_t cur[4][4], nxt[4][4];
      if (kcnt > 0) load_idx(kbeg, cur);
      int s = 0;
      uint32_t ph = 0;
      for (int ki = 0; ki < kcnt; ++ki) {
        if (ki + 1 < kcnt) load_idx(kbeg + ki + 1, nxt);    // in flight while this offset is gathered
        for (int kb = 0; kb < num_kb; ++kb) {
          const bool second = kb >= p.nkb1;
          const int kbl = second ? kb - p.nkb1 : kb;
          const int rb = second ? p.row_bytes2 : p.row_bytes;
          const bool col_ok = kbl * KBLOCK_BYTES + chunk * 16 < rb;
          const uint8_t* src = (second ? p.in2 : p.in) + kbl * KBLOCK_BYTES + chunk * 16;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (t < TM) {
              mbar_wait(a_empty + s, ph ^ 1);
              const uint32_t a_base = a_ring_base + uint32_t(s) * A_BYTES;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int32_t row = cur[t][i];
                const bool ok = col_ok && row >= 0;
                cp_async16(a_base + dst_off[i], src + size_t(ok ? row : 0) * rb, ok ? 16u : 0u);
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
    if (kcnt > 0) {
      mbar_wait(acc_bar, 0);
      tc_fence_after();
      const int ncols = min(p.n_tile, p.c_out - n0);
      const int lq = warp & 3;                       // TMEM lane quadrant this warp may read
      const float* bias = (partial && blockIdx.z != 0) ? nullptr : p.bias;
      for (int t = warp >> 2; t < TM; t += PROD_WARPS / 4) {
        const int64_t o = m0 + int64_t(t) * p.rt + lq * 32 + lane;
        const bool row_ok = lq * 32 + lane < p.rt && o < p.n_out;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + (uint32_t(lq * 32) << 16) + uint32_t(t * p.n_tile + c0), v);
          if (bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < ncols) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldg(bias + n0 + c0 + j));
          }
          if (row_ok) {
            float* orow = p.out + size_t(o) * p.c_out + n0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (c0 + j + 3 < ncols) {
                if (partial) {
                  red_add_v4(orow + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                             __uint_as_float(v[j + 3]));
                } else {
                  float4 x;
                  x.x = __uint_as_float(v[j]);
                  x.y = __uint_as_float(v[j + 1]);
                  x.z = __uint_as_float(v[j + 2]);
                  x.w = __uint_as_float(v[j + 3]);
                  *reinterpret_cast<float4*>(orow + j) = x;
                }
              } else {
                for (int jj = j; jj < j + 4; ++jj)
                  if (c0 + jj < ncols) {
                    if (partial) atomicAdd(orow + jj, __uint_as_float(v[jj]));
                    else orow[jj] = __uint_as_float(v[jj]);
                  }
              }
            }
          }
          if (p.stats) {
            // BatchNorm statistics of the output straight from the accumulator registers: a butterfly over the warp's 32
            // rows leaves lane j with the sum (and sum of squares) of column c0 + j -> per-CTA partials in shared memory
            float s1[32], s2[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float x = row_ok ? __uint_as_float(v[j]) : 0.f;
              s1[j] = x;
              s2[j] = x * x;
            }
#pragma unroll
            for (int w = 16; w >= 1; w >>= 1) {
              const bool upper = (lane & w) != 0;
#pragma unroll
              for (int j = 0; j < w; ++j) {
                // lanes with bit w set keep the upper half of the remaining columns, the others the lower half
                const float a1 = upper ? s1[j] : s1[j + w], a2 = upper ? s2[j] : s2[j + w];
                const float k1 = upper ? s1[j + w] : s1[j], k2 = upper ? s2[j + w] : s2[j];
                s1[j] = k1 + __shfl_xor_sync(0xffffffffu, a1, w);
                s2[j] = k2 + __shfl_xor_sync(0xffffffffu, a2, w);
              }
            }
            // lane L now holds column c0 + L (a lane keeps the half of the columns its own bit selects at every step)
            if (c0 + lane < ncols) {
              atomicAdd(s_stats + c0 + lane, s1[0]);
              atomicAdd(s_stats + 256 + c0 + lane, s2[0]);
            }
          }
        }
      }
      tc_fence_before();
    }
  } else if (warp == PROD_WARPS) {
    // =================================== MMA issuer (warp-uniform loop, one elected lane issues) ================
    {
      // instruction descriptor: D fp32 (bit 4), A / B bf16 (format 1), K-major B, N = n_tile, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(p.n_tile >> 3) << 17) | (uint32_t(BM >> 4) << 24);
      const uint32_t desc_hi = uint32_t(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
      const uint32_t b_ring_base = smem_u32(b_ring);
      int sb = 0, tsa = 0;
      uint32_t phb = 0, phta = 0;
      for (int ki = 0; ki < kcnt; ++ki) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(b_full + sb, phb);
          const uint32_t b_base = b_ring_base + uint32_t(sb) * b_bytes;
          const uint32_t b_lo32 = ((b_base >> 4) & 0x3FFF) | (1u << 16);
          const bool second = kb >= p.nkb1;
          const int valid = min(KBLOCK_BYTES, second ? p.row_bytes2 - (kb - p.nkb1) * KBLOCK_BYTES : p.row_bytes - kb * KBLOCK_BYTES);
          const int ksteps = (valid + 63) >> 6;          // 16 channels (64 bytes of fp32) per instruction
          const uint32_t acc_flag = (ki | kb) ? 1u : 0u;
          for (int t = 0; t < TM; ++t) {
            const uint32_t d_addr = tmem_base + uint32_t(t * p.n_tile);
            mbar_wait(ta_full + tsa, phta);
            tc_fence_after();
            const uint32_t a_tm = tmem_base + uint32_t(p.ta_col0 + tsa * TA_COLS);
            if (elect_one()) {
#pragma unroll 2
              for (int j = 0; j < ksteps; ++j) {
                // B stage row n: [w_hi x32 | w_lo x32] bf16; 16 channels = 32 bytes = 2 sixteen-byte units
                const uint64_t b_hi = desc_from(b_lo32 + 2 * j, desc_hi);
                const uint64_t b_lo = desc_from(b_lo32 + 4 + 2 * j, desc_hi);
                umma_ts_f16(d_addr, a_tm + 8 * j, b_hi, idesc, acc_flag | uint32_t(j));
                umma_ts_f16(d_addr, a_tm + 16 + 8 * j, b_hi, idesc, 1u);
                umma_ts_f16(d_addr, a_tm + 8 * j, b_lo, idesc, 1u);
              }
              umma_commit(ta_empty + tsa);
            }
            __syncwarp();
            if (++tsa == STA) {
              tsa = 0;
              phta ^= 1;
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
      if (kcnt > 0 && elect_one()) umma_commit(acc_bar);
    }
    __syncwarp();
  } else if (warp == PROD_WARPS + 1) {
    // =================================== weight TMA producer ===================================
    {
      const uint32_t b_ring_base = smem_u32(b_ring);
      int sb = 0;
      uint32_t phb = 0;
      for (int ki = 0; ki < kcnt; ++ki) {
        const int row = (kbeg + ki) * p.c_out + n0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(b_empty + sb, phb ^ 1);
          const uint32_t b_dst = b_ring_base + uint32_t(sb) * b_bytes;
          if (elect_one()) {
            mbar_expect_tx(b_full + sb, b_bytes);
            tma_load_2d(b_dst, &tmap_w, b_full + sb, kb * 64, row);
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
  } else {
    // =================================== hi/lo splitters (128 threads): smem tile -> TMEM ===================================
    // thread <-> row of the tile (TMEM lane; a warp may only touch the lane quadrant (warp index in the CTA) % 4)
    const int lq = warp & 3;
    const int r = lq * 32 + lane;
    const uint32_t lane_addr = uint32_t(lq * 32) << 16;
    const int total = kcnt * num_kb * TM;
    uint32_t sw_off[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) sw_off[c] = uint32_t(r * KBLOCK_BYTES + ((c ^ (r & 7)) << 4));
    int s = 0, ts = 0;
    uint32_t ph = 0, pht = 0;
    for (int it = 0; it < total; ++it) {
      mbar_wait(a_full + s, ph);                 // gathered rows have landed (cp.async completion arrivals)
      const uint8_t* a_tile = a_ring + size_t(s) * A_BYTES;
      uint32_t w[32];                            // [0,16): hi pairs, [16,32): lo pairs
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(a_tile + sw_off[c]);
        split2(v.x, v.y, w[2 * c], w[16 + 2 * c]);
        split2(v.z, v.w, w[2 * c + 1], w[16 + 2 * c + 1]);
      }
      mbar_arrive(a_empty + s);                  // the shared-memory stage can be refilled
      mbar_wait(ta_empty + ts, pht ^ 1);         // MMAs that read this TMEM stage last time are done
      tc_fence_after();
      tmem_st32(tmem_base + lane_addr + uint32_t(p.ta_col0 + ts * TA_COLS), w);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(ta_full + ts);
      if (++s == SA) {
        s = 0;
        ph ^= 1;
      }
      if (++ts == STA) {
        ts = 0;
        pht ^= 1;
      }
    }
  }

  __syncthreads();
  if (p.stats) {
    // one fp64 atomic per column and CTA into one of the 8 interleaved accumulator copies (the layout lgs_bn_fwd reads)
    const int ncols = min(p.n_tile, p.c_out - n0);
    double* dst = p.stats + size_t(blockIdx.x & 7) * 2 * p.c_out + n0;
    for (int i = tid; i < ncols; i += THREADS) {
      atomicAdd(dst + i, double(s_stats[i]));
      atomicAdd(dst + p.c_out + i, double(s_stats[256 + i]));
    }
  }
  if (warp == PROD_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(p.tmem_cols))
                 : "memory");
  }
}

// ---- weight operand form -------------------------------------------------------------------------------------------
// From W [K, c_in, c_out] (fp32 parameter):
//   fwd [K][c_out][ceil(c_in/32)][64]  bf16: per output channel and 32-channel block  [hi x32 | lo x32] of W[k][kb*32+e][n]
//   bwd [K][c_in][ceil(c_out/32)][64]  bf16: the same for dgrad (reduction over the output channels)
// hi = RN_bf16(w), lo = RN_bf16(w - hi); channels beyond the real count are zero.
// grid: one block (32 x 8 threads) per 32x32 tile of W[k]; a block finds its layer by binary search over desc[.][6].
__device__ __forceinline__ void bx3_store(__nv_bfloat16* dst, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  dst[0] = h;
  dst[32] = __float2bfloat16_rn(v - __bfloat162float(h));
}

__device__ __forceinline__ void weight_prep_bx3_tile(const float* __restrict__ w, __nv_bfloat16* __restrict__ fwd,
                                                     __nv_bfloat16* __restrict__ bwd, int c_in, int c_out, int64_t t) {
  __shared__ float tile[32][33];
  const int tx = (c_out + 31) / 32, ty = (c_in + 31) / 32;
  const int cob = int(t % tx);
  t /= tx;
  const int cib = int(t % ty);
  const int k = int(t / ty);
  const int co0 = cob * 32, ci0 = cib * 32;
  const float* wk = w + int64_t(k) * c_in * c_out;
#pragma unroll
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int ci = ci0 + r, co = co0 + threadIdx.x;
    const float v = (ci < c_in && co < c_out) ? wk[int64_t(ci) * c_out + co] : 0.f;
    // bwd row = input channel ci, block = cob, element = threadIdx.x (output channel inside the block)
    if (bwd && ci < c_in) bx3_store(bwd + ((int64_t(k) * c_in + ci) * tx + cob) * 64 + threadIdx.x, v);
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  if (!fwd) return;
#pragma unroll
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int co = co0 + r;
    // fwd row = output channel co, block = cib, element = threadIdx.x (input channel inside the block)
    if (co < c_out) bx3_store(fwd + ((int64_t(k) * c_out + co) * ty + cib) * 64 + threadIdx.x, tile[threadIdx.x][r]);
  }
}

__global__ void __launch_bounds__(256) weight_prep_bx3_batch_kernel(const int64_t* __restrict__ desc, int n_layers) {
  pdl_grid_sync();
  const int64_t b = blockIdx.x;
  int lo = 0, hi = n_layers - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (desc[mid * 8 + 6] <= b) lo = mid; else hi = mid - 1;
  }
  const int64_t* d = desc + lo * 8;
  weight_prep_bx3_tile(reinterpret_cast<const float*>(d[0]), reinterpret_cast<__nv_bfloat16*>(d[1]),
                       reinterpret_cast<__nv_bfloat16*>(d[2]), int(d[4]), int(d[5]), b - d[6]);
}

__global__ void __launch_bounds__(256) weight_prep_bx3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ fwd,
                                                              __nv_bfloat16* __restrict__ bwd, int c_in, int c_out) {
  weight_prep_bx3_tile(w, fwd, bwd, c_in, c_out, int64_t(blockIdx.x));
}

}  // namespace bx3

int weight_prep_bx3_batch(const int64_t* desc, int n_layers, int64_t total_tiles, cudaStream_t stream) {
  if (n_layers == 0 || total_tiles == 0) return LGS_OK;
  const dim3 block{32u, 8u, 1u};
  LGS_LAUNCH_PDL(bx3::weight_prep_bx3_batch_kernel, unsigned(total_tiles), block, 0, stream, desc, n_layers);
  return LGS_OK;
}

int weight_prep_bx3(const float* w, int K, int c_in, int c_out, void* fwd, void* bwd, cudaStream_t stream) {
  const int64_t tiles = int64_t(K) * ((c_in + 31) / 32) * ((c_out + 31) / 32);
  if (tiles == 0) return LGS_OK;
  const dim3 block{32u, 8u, 1u};
  LGS_LAUNCH(bx3::weight_prep_bx3_kernel, unsigned(tiles), block, 0, stream, w, static_cast<__nv_bfloat16*>(fwd),
             static_cast<__nv_bfloat16*>(bwd), c_in, c_out);
  return LGS_OK;
}

int conv_bx3_shape_ok(int c_in, int c_out) { return (c_in % 4 == 0 && c_in >= 4 && c_out % 4 == 0) ? 1 : 0; }

// tuning knobs: read from the environment once (this function runs ~125 times per training step), overridable at run time
// through lgs_tune("bx3_tm" | "bx3_rt" | "bx3_ks" | "bx3_ns" | "bx3_sa" | "bx3_no_balance", value); 0 = heuristic
struct Bx3Knobs {
  int tm, rt, ks, ns, sa, nobalance;
  Bx3Knobs() {
    auto geti = [](const char* n, int d) { const char* e = getenv(n); return e ? atoi(e) : d; };
    tm = geti("LGS_BX3_TM", 0);
    rt = geti("LGS_BX3_RT", 0);
    ks = geti("LGS_BX3_KS", 0);
    ns = geti("LGS_BX3_NS", 0);
    sa = geti("LGS_BX3_SA", 0);
    nobalance = geti("LGS_BX3_NO_BALANCE", 0);
  }
};
static Bx3Knobs& bx3_knobs() {
  static Bx3Knobs k;
  return k;
}
int bx3_tune(const char* key, int value) {
  Bx3Knobs& k = bx3_knobs();
  const std::string s(key);
  if (s == "bx3_tm") k.tm = value;
  else if (s == "bx3_rt") k.rt = value;
  else if (s == "bx3_ks") k.ks = value;
  else if (s == "bx3_ns") k.ns = value;
  else if (s == "bx3_sa") k.sa = value;
  else if (s == "bx3_no_balance") k.nobalance = value;
  else return 0;
  return 1;
}

// in2 / c_in2: optional second gather source (channels [c_in, c_in + c_in2) of the convolution's input), c_in % 32 == 0 then.
// w: LGS_W_BX3 operand of the FULL input width c_in + c_in2.  stats: optional BatchNorm column sums of the output (pre-zeroed).
int conv_fwd_bx3(const void* in, int c_in, const void* in2, int c_in2, const void* w, int K, int c_out, const int32_t* table,
                 int64_t n_out, int reverse_k, const float* bias, float* out, double* stats, cudaStream_t stream) {
  using namespace bx3;
  if (!conv_bx3_shape_ok(c_in, c_out)) return LGS_E_UNSUPPORTED;
  if (in2 && (c_in % 32 != 0 || c_in2 % 4 != 0 || c_in2 < 4)) return LGS_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(in2) & 15) || (reinterpret_cast<uintptr_t>(w) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15))
    return LGS_E_UNSUPPORTED;
  if (int64_t(K) * c_out >= (int64_t(1) << 31)) return LGS_E_UNSUPPORTED;
  if (n_out == 0) return LGS_OK;
  EncodeTiledFn encode = get_encode();
  if (!encode) return fail(LGS_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(conv_bx3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) return fail(LGS_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  const Bx3Knobs& kn = bx3_knobs();

  Params q;
  q.in = static_cast<const uint8_t*>(in);
  q.in2 = static_cast<const uint8_t*>(in2);
  q.row_bytes = c_in * 4;
  q.row_bytes2 = in2 ? c_in2 * 4 : 0;
  q.nkb1 = (q.row_bytes + KBLOCK_BYTES - 1) / KBLOCK_BYTES;
  q.num_kb = q.nkb1 + (in2 ? (q.row_bytes2 + KBLOCK_BYTES - 1) / KBLOCK_BYTES : 0);
  q.K = K;
  q.c_out = c_out;
  q.table = table;
  q.n_out = n_out;
  q.reverse_k = reverse_k;
  q.bias = bias;
  q.out = out;
  q.stats = stats;

  // ---- decomposition: (n_slices, TM, rt, k_splits) minimising  waves x bytes one CTA moves through the L2->SM path ----
  const int c_pad = ((c_out + 15) / 16) * 16;
  const int64_t tiles = cdiv(n_out, BM);
  const int base_slices = (c_pad + 255) / 256;
  double best = -1.0;
  int bTM = 1, bRT = BM, bKS = 1, bNS = base_slices, bNT = 0;
  int64_t bGX = 1;
  for (int nsm = 1; nsm <= 4; nsm <<= 1) {
    const int ns = base_slices * nsm;
    const int nt = ((((c_pad + ns - 1) / ns) + 15) / 16) * 16;
    if (nsm > 1 && nt < 64) break;
    if (kn.ns && ns != kn.ns) continue;
    const int ns_real = (c_pad + nt - 1) / nt;
    for (int tm = 1; tm <= 4; ++tm) {
      if (tm * nt + 2 * TA_COLS > 512) break;
      if (kn.tm && tm != kn.tm) continue;
      if (int64_t(tm - 1) * BM >= n_out && tm > 1) break;
      for (int ks = 1; ks <= K; ++ks) {
        const int kper = (K + ks - 1) / ks;
        if ((K + kper - 1) / kper != ks) continue;          // only distinct splits
        if (kn.ks && ks != kn.ks) continue;
        if (ks > 1 && stats) break;                          // fused statistics need complete sums
        const int64_t gx_full = cdiv(tiles, tm);
        for (int bal = 0; bal < 2; ++bal) {
          int64_t gx = gx_full;
          int rt = BM;
          if (bal) {
            if (kn.nobalance) break;
            const int64_t others = int64_t(ns_real) * ks;
            const int64_t w = cdiv(gx_full * others, 148);
            gx = std::max<int64_t>(gx_full, int64_t(148) * w / others);
            if (gx == gx_full) break;
            rt = int(cdiv(cdiv(n_out, gx), int64_t(tm)));
            if (rt > BM) break;
            gx = cdiv(n_out, int64_t(tm) * rt);
          }
          const int64_t ctas = gx * ns_real * ks;
          const double waves = double(cdiv(ctas, 148));
          const double stage_rows = std::max(double(rt) / BM, 0.45);     // empty MMA lanes still cost tensor time
          const double gather = double(tm) * stage_rows * A_BYTES * kper * q.num_kb;
          const double weights = double(nt) * KBLOCK_BYTES * kper * q.num_kb;
          const double epi = double(tm) * rt * nt * 4.0 * (ks > 1 ? 3.0 : 1.0);
          const double cost = waves * (gather + weights + epi + 160.0 * 1024);
          if (best < 0 || cost < best * 0.999) {
            best = cost, bTM = tm, bRT = rt, bKS = ks, bNS = ns_real, bNT = nt, bGX = gx;
          }
        }
      }
    }
  }
  if (best < 0) return LGS_E_UNSUPPORTED;
  if (kn.rt) {
    bRT = std::min(BM, std::max(8, kn.rt));
    bGX = cdiv(n_out, int64_t(bTM) * bRT);
  }
  q.TM = bTM;
  q.rt = bRT;
  q.n_tile = bNT;
  q.k_per_split = (K + bKS - 1) / bKS;
  q.k_splits = bKS;
  q.b_stage_bytes = bNT * KBLOCK_BYTES;
  q.ta_col0 = ((bTM * bNT + 31) / 32) * 32;
  q.ta_stages = std::min(MAX_TA, (512 - q.ta_col0) / TA_COLS);
  int cols = 32;
  while (cols < q.ta_col0 + q.ta_stages * TA_COLS) cols <<= 1;
  q.tmem_cols = cols;
  const int fixed = (2 * MAX_A + 2 * MAX_B + 2 * MAX_TA + 1) * 8 + 16 + 2048 + 1024;   // barriers, TMEM pointer, statistics, alignment
  int sb = 3;
  int sa = (227 * 1024 - fixed - sb * q.b_stage_bytes) / A_BYTES;
  if (sa < 4) {
    sb = 2;
    sa = (227 * 1024 - fixed - sb * q.b_stage_bytes) / A_BYTES;
  }
  if (sa > MAX_A) sa = MAX_A;
  if (kn.sa && kn.sa < sa) sa = std::max(2, kn.sa);
  if (sa < 2) return LGS_E_UNSUPPORTED;
  q.a_stages = sa;
  q.b_stages = sb;
  const size_t smem = size_t(sa) * A_BYTES + size_t(sb) * q.b_stage_bytes + fixed;

  if (q.k_splits > 1) {
    const int rz = zero_fill_async(out, size_t(n_out) * c_out * sizeof(float), stream);
    if (rz != LGS_OK) return rz;
  }

  // tensor map over the BX3 operand viewed as bf16 [K * c_out rows][num_kb * 64], box {64 (= 128 B), n_tile rows}
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {cuuint64_t(q.num_kb) * 64, cuuint64_t(K) * cuuint64_t(c_out)};
  const cuuint64_t gstride[1] = {cuuint64_t(q.num_kb) * 128};
  const cuuint32_t box[2] = {64, cuuint32_t(bNT)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(LGS_E_CUDA, "cuTensorMapEncodeTiled (bx3) failed (%d) c_in=%d c_out=%d K=%d", int(cr), c_in + c_in2, c_out, K);
  const dim3 grid{unsigned(bGX), unsigned(bNS), unsigned(q.k_splits)};
  LGS_LAUNCH_PDL(conv_bx3_kernel, grid, THREADS, smem, stream, tmap, q);
  return LGS_OK;
}

}  // namespace lgs
