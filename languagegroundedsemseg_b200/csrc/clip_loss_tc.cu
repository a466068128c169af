// CLIP text-anchor cross-entropy on the 5th-generation tensor cores (north_star: "a single tcgen05 GEMM +
// fused softmax-cross-entropy kernel").  Replaces lib/losses/ContrastiveLanguageLoss.py:224-237 (expand to (N,200,C) + bmm +
// CrossEntropyLoss) and lib/losses/utils.py:99-103 (feature_sim argmax).
//
// One CTA owns 128 points (rows of F) and runs two GEMMs back to back with everything in between on chip:
//   phase 1   S_raw[128 x a] = F_tile @ An^T              K = c   (3xTF32: hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM)
//             * the F tile arrives by TMA (128B-swizzled, K-major) in 32-channel blocks;
//             * four "row warps" (thread <-> row <-> TMEM lane) read their row from shared memory, accumulate |F_i|^2, split the
//               values into hi = trunc_tf32(f), lo = f - hi and tcgen05.st both into a TMEM ring; the MMAs run in TS form
//               (A from TMEM), the pre-split anchors (hi | lo, lgs weight_prep) are the TMA-loaded B operand;
//   epilogue 1  the same row warps tcgen05.ld their row of S, scale by 1/|F_i|, and do the softmax cross-entropy, the argmax
//             and G = softmax - onehot in registers; G (hi | lo) goes straight back into TMEM over the S columns;
//   phase 2   dF_raw[128 x c] = G @ An                     K = a   (3xTF32, TS form: A = G from TMEM, B = An^T blocks by TMA),
//             in output-channel chunks of 96 columns;
//   epilogue 2  dF = (dF_raw - (G.S) F_hat) / |F|  -> one store per element.
// The [n, a] logits never reach HBM (optional grad_logits output for the learned anchor projection excepted).
// TMEM columns (512 allocated): [0,n1) S then G_hi | [n1,2 n1) G_lo | [2 n1, 2 n1 + 96) dF accumulator; during phase 1 the
// split-F ring (64 columns per stage) lives in [round32(n1), ...), which phase 2 reuses once all phase-1 MMAs are complete.
#include <cuda.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace lgs {

int weight_prep(const float* w, int K, int c_in, int c_out, int nsplit, void* fwd, void* bwd, int dtype,
                cudaStream_t stream);  // conv_tc.cu

namespace ctc {

using namespace tc;

constexpr int BM = 128;
constexpr int KB_BYTES = 128;  // one 128B-swizzle span: 32 fp32 of the reduction dimension
constexpr int KB_ELEMS = 32;
constexpr int A_BYTES = BM * KB_BYTES;
constexpr int A_STAGES = 4;
constexpr int B_STAGES = 2;
constexpr int TA_STAGES_MAX = 4;
constexpr int THREADS = 192;  // warps 0-3: row warps (split, softmax, epilogues); warp 4: MMA issuer; warp 5: TMA
constexpr int A_MAX = 208;    // anchors per launch: 2 * ceil16(a) + 96 TMEM columns <= 512
constexpr int N2_MAX = 96;
constexpr float kNormEps = 1e-12f;  // F.normalize default eps

struct Params {
  const float* feats;
  int64_t n;
  int32_t c, a;
  const int64_t* labels;
  int64_t ignore_label;
  float* loss;
  float* grad_feats;
  int32_t* pred;
  float* grad_logits;
  int32_t n1;        // phase-1 MMA N: ceil16(a)
  int32_t kp;        // phase-2 K: ceil8(a)
  int32_t num_kb1;   // ceil(c / 32)
  int32_t num_kb2;   // ceil(kp / 32)
  int32_t n2;        // phase-2 MMA N (output-channel chunk): min(96, ceil16(c))
  int32_t n_chunks;  // ceil(c / n2)
  int32_t b_bytes;   // one B stage: 2 * max(n1, n2) * 128  (hi | lo)
  int32_t ta_col0, ta_stages;
  int32_t acc2_col;  // 2 * n1
};

template <int W>
__device__ __forceinline__ void ld_cols(uint32_t addr, uint32_t (&v)[32]) {
  if constexpr (W == 32) {
    tmem_ld32(addr, v);
  } else {
    uint32_t t[16];
    tmem_ld16(addr, t);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      v[j] = t[j];
      v[j + 16] = 0u;
    }
  }
}
template <int W>
__device__ __forceinline__ void st_cols(uint32_t addr, const uint32_t (&v)[32]) {
  if constexpr (W == 32) {
    tmem_st32(addr, v);
  } else {
    uint32_t t[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) t[j] = v[j];
    tmem_st16(addr, t);
  }
}

__global__ void __launch_bounds__(THREADS, 1)
clip_ce_tc_kernel(const __grid_constant__ CUtensorMap tmap_f, const __grid_constant__ CUtensorMap tmap_an,
                  const __grid_constant__ CUtensorMap tmap_ant, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + size_t(A_STAGES) * A_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(b_ring + size_t(B_STAGES) * p.b_bytes);
  uint64_t* a_empty = a_full + A_STAGES;
  uint64_t* b_full = a_empty + A_STAGES;
  uint64_t* b_empty = b_full + B_STAGES;
  uint64_t* ta_full = b_empty + B_STAGES;
  uint64_t* ta_empty = ta_full + TA_STAGES_MAX;
  uint64_t* acc1_bar = ta_empty + TA_STAGES_MAX;   // phase-1 accumulator complete
  uint64_t* g_full = acc1_bar + 1;                 // G (hi | lo) of all 128 rows is in TMEM
  uint64_t* acc2_full = g_full + 1;                // one dF chunk accumulated
  uint64_t* acc2_empty = acc2_full + 1;            // ... and drained by the row warps
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc2_empty + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = int64_t(blockIdx.x) * BM;
  const int c = p.c, a = p.a, n1 = p.n1, n2 = p.n2;
  const bool need_g = p.grad_feats != nullptr;     // phase 2 runs only when dF is wanted

  if (tid == 0) {
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(a_full + s, 1);
      mbar_init(a_empty + s, 128);
    }
    for (int s = 0; s < B_STAGES; ++s) {
      mbar_init(b_full + s, 1);
      mbar_init(b_empty + s, 1);
    }
    for (int s = 0; s < TA_STAGES_MAX; ++s) {
      mbar_init(ta_full + s, 128);
      mbar_init(ta_empty + s, 1);
    }
    mbar_init(acc1_bar, 1);
    mbar_init(g_full, 128);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 5 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_f) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_an) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_ant) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < 4) {
    // ============================ row warps: thread <-> row of the tile <-> TMEM lane ============================
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = uint32_t(warp * 32) << 16;
    const int64_t gi = m0 + r;
    // ---- phase 1: split the landed F blocks into the TMEM ring, accumulate the row norm ----
    float ss = 0.f;
    {
      int s = 0, ts = 0;
      uint32_t ph = 0, pht = 0;
      for (int kb = 0; kb < p.num_kb1; ++kb) {
        mbar_wait(a_full + s, ph);
        const uint8_t* a_row = a_ring + size_t(s) * A_BYTES + r * KB_BYTES;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(a_row + ((q ^ (r & 7)) << 4));
          const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            ss = fmaf(f[e], f[e], ss);
            const uint32_t h = __float_as_uint(f[e]) & 0xFFFFE000u;
            hi[q * 4 + e] = h;
            lo[q * 4 + e] = __float_as_uint(f[e] - __uint_as_float(h));
          }
        }
        mbar_arrive(a_empty + s);               // the shared-memory stage can be refilled
        mbar_wait(ta_empty + ts, pht ^ 1);      // MMAs that read this TMEM stage last time are done
        tc_fence_after();
        const uint32_t ta = tmem_base + lane_addr + uint32_t(p.ta_col0 + ts * 64);
        tmem_st32(ta, hi);
        tmem_st32(ta + 32, lo);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(ta_full + ts);
        if (++s == A_STAGES) {
          s = 0;
          ph ^= 1;
        }
        if (++ts == p.ta_stages) {
          ts = 0;
          pht ^= 1;
        }
      }
    }
    // ---- epilogue 1: cosine logits -> softmax cross-entropy, argmax, G = softmax - onehot (back into TMEM) ----
    mbar_wait(acc1_bar, 0);
    tc_fence_after();
    const float inv = 1.f / fmaxf(sqrtf(ss), kNormEps);
    const int64_t y = gi < p.n ? p.labels[gi] : p.ignore_label;
    const bool valid = gi < p.n && y != p.ignore_label;
    const uint32_t s_addr = tmem_base + lane_addr;
    float mx = -INFINITY;
    int amax = 0;
    for (int c0 = 0; c0 < n1; c0 += 32) {
      uint32_t v[32];
      if (c0 + 32 <= n1) ld_cols<32>(s_addr + c0, v);
      else ld_cols<16>(s_addr + c0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float s = __uint_as_float(v[j]) * inv;
        if (c0 + j < a && s > mx) {
          mx = s;
          amax = c0 + j;
        }
      }
    }
    float se = 0.f, sy = 0.f;
    for (int c0 = 0; c0 < n1; c0 += 32) {
      uint32_t v[32];
      if (c0 + 32 <= n1) ld_cols<32>(s_addr + c0, v);
      else ld_cols<16>(s_addr + c0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float s = __uint_as_float(v[j]) * inv;
        if (c0 + j < a) {
          se += __expf(s - mx);
          if (valid && int64_t(c0 + j) == y) sy = s;
        }
      }
    }
    if (gi < p.n) {
      if (p.loss) p.loss[gi] = valid ? (logf(se) + mx - sy) : 0.f;
      if (p.pred) p.pred[gi] = amax;
    }
    float sd = 0.f;  // sum_j g_j S_j
    if (need_g || p.grad_logits) {
      const float inv_se = 1.f / se;
      for (int c0 = 0; c0 < n1; c0 += 32) {
        const bool full = c0 + 32 <= n1;
        uint32_t v[32], hi[32], lo[32];
        if (full) ld_cols<32>(s_addr + c0, v);
        else ld_cols<16>(s_addr + c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float s = __uint_as_float(v[j]) * inv;
          float g = 0.f;
          if (valid && c0 + j < a) {
            g = __expf(s - mx) * inv_se - (int64_t(c0 + j) == y ? 1.f : 0.f);
            sd = fmaf(g, s, sd);
          }
          const uint32_t h = __float_as_uint(g) & 0xFFFFE000u;
          hi[j] = h;
          lo[j] = __float_as_uint(g - __uint_as_float(h));
          v[j] = __float_as_uint(g);
        }
        if (need_g) {
          if (full) {
            st_cols<32>(s_addr + c0, hi);
            st_cols<32>(s_addr + n1 + c0, lo);
          } else {
            st_cols<16>(s_addr + c0, hi);
            st_cols<16>(s_addr + n1 + c0, lo);
          }
        }
        if (p.grad_logits && gi < p.n) {
          float* grow = p.grad_logits + size_t(gi) * a + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (c0 + j + 3 < a)
              *reinterpret_cast<float4*>(grow + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                  __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
      }
    }
    if (need_g) {
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(g_full);
      // ---- epilogue 2: dF = (G @ An - (G.S) F_hat) / |F|, one 96-channel chunk at a time ----
      const uint32_t acc2 = tmem_base + lane_addr + uint32_t(p.acc2_col);
      const float sdi = sd * inv;
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        mbar_wait(acc2_full, uint32_t(ch & 1));
        tc_fence_after();
        for (int c0 = 0; c0 < n2; c0 += 32) {
          const int col0 = ch * n2 + c0;
          if (col0 >= c) break;
          uint32_t v[32];
          if (c0 + 32 <= n2) ld_cols<32>(acc2 + c0, v);
          else ld_cols<16>(acc2 + c0, v);
          if (gi < p.n) {
            const float* frow = p.feats + size_t(gi) * c + col0;
            float* orow = p.grad_feats + size_t(gi) * c + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (col0 + j < c) {   // c % 4 == 0 and col0 % 4 == 0: whole float4 groups
                const float4 f = __ldg(reinterpret_cast<const float4*>(frow + j));
                float4 o;
                o.x = (__uint_as_float(v[j]) - sdi * f.x) * inv;
                o.y = (__uint_as_float(v[j + 1]) - sdi * f.y) * inv;
                o.z = (__uint_as_float(v[j + 2]) - sdi * f.z) * inv;
                o.w = (__uint_as_float(v[j + 3]) - sdi * f.w) * inv;
                *reinterpret_cast<float4*>(orow + j) = o;
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(acc2_empty);
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // ============================ MMA issuer (warp-uniform loops, one elected lane issues) ============================
    const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(n1 >> 3) << 17) | (uint32_t(BM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(n2 >> 3) << 17) | (uint32_t(BM >> 4) << 24);
    const uint32_t desc_hi = uint32_t(1024 >> 4) | (1u << 14) | (2u << 29);
    const uint32_t b_ring_base = smem_u32(b_ring);
    int sb = 0, ts = 0;
    uint32_t phb = 0, pht = 0;
    for (int kb = 0; kb < p.num_kb1; ++kb) {
      mbar_wait(b_full + sb, phb);
      mbar_wait(ta_full + ts, pht);
      tc_fence_after();
      const uint32_t b_base = b_ring_base + uint32_t(sb) * uint32_t(p.b_bytes);
      const uint32_t b_lo32 = ((b_base >> 4) & 0x3FFF) | (1u << 16);
      const uint32_t b_lo_off = uint32_t(n1 * KB_BYTES) >> 4;
      const int ksteps = (min(KB_ELEMS, c - kb * KB_ELEMS) + 7) >> 3;
      const uint32_t a_tm = tmem_base + uint32_t(p.ta_col0 + ts * 64);
      if (elect_one()) {
#pragma unroll 4
        for (int j = 0; j < ksteps; ++j) {
          const uint64_t b_hi = desc_from(b_lo32 + 2 * j, desc_hi);
          umma_ts_tf32(tmem_base, a_tm + 8 * j, b_hi, idesc1, (kb | j) ? 1u : 0u);
          umma_ts_tf32(tmem_base, a_tm + 32 + 8 * j, b_hi, idesc1, 1u);
          umma_ts_tf32(tmem_base, a_tm + 8 * j, desc_from(b_lo32 + b_lo_off + 2 * j, desc_hi), idesc1, 1u);
        }
        umma_commit(ta_empty + ts);
        umma_commit(b_empty + sb);
      }
      __syncwarp();
      if (++sb == B_STAGES) {
        sb = 0;
        phb ^= 1;
      }
      if (++ts == p.ta_stages) {
        ts = 0;
        pht ^= 1;
      }
    }
    if (elect_one()) umma_commit(acc1_bar);
    __syncwarp();
    if (need_g) {
      mbar_wait(g_full, 0);
      tc_fence_after();
      const uint32_t acc2 = tmem_base + uint32_t(p.acc2_col);
      const uint32_t b_lo_off = uint32_t(n2 * KB_BYTES) >> 4;
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        if (ch > 0) {
          mbar_wait(acc2_empty, uint32_t((ch - 1) & 1));
          tc_fence_after();
        }
        for (int kb = 0; kb < p.num_kb2; ++kb) {
          mbar_wait(b_full + sb, phb);
          const uint32_t b_base = b_ring_base + uint32_t(sb) * uint32_t(p.b_bytes);
          const uint32_t b_lo32 = ((b_base >> 4) & 0x3FFF) | (1u << 16);
          const int ksteps = min(KB_ELEMS, p.kp - kb * KB_ELEMS) >> 3;   // kp % 8 == 0
          if (elect_one()) {
#pragma unroll 4
            for (int j = 0; j < ksteps; ++j) {
              const uint32_t acol = uint32_t(kb * KB_ELEMS + 8 * j);
              const uint64_t b_hi = desc_from(b_lo32 + 2 * j, desc_hi);
              umma_ts_tf32(acc2, tmem_base + acol, b_hi, idesc2, (kb | j) ? 1u : 0u);
              umma_ts_tf32(acc2, tmem_base + uint32_t(n1) + acol, b_hi, idesc2, 1u);
              umma_ts_tf32(acc2, tmem_base + acol, desc_from(b_lo32 + b_lo_off + 2 * j, desc_hi), idesc2, 1u);
            }
            umma_commit(b_empty + sb);
          }
          __syncwarp();
          if (++sb == B_STAGES) {
            sb = 0;
            phb ^= 1;
          }
        }
        if (elect_one()) umma_commit(acc2_full);
        __syncwarp();
      }
    }
  } else {
    // ============================ TMA producer (warp-uniform loops, one elected lane issues) ============================
    const uint32_t a_ring_base = smem_u32(a_ring), b_ring_base = smem_u32(b_ring);
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    for (int kb = 0; kb < p.num_kb1; ++kb) {
      mbar_wait(a_empty + sa, pha ^ 1);
      if (elect_one()) {
        mbar_expect_tx(a_full + sa, A_BYTES);
        tma_load_2d(a_ring_base + uint32_t(sa) * A_BYTES, &tmap_f, a_full + sa, kb * KB_ELEMS, int32_t(m0));
      }
      __syncwarp();
      mbar_wait(b_empty + sb, phb ^ 1);
      if (elect_one()) {
        const uint32_t b_dst = b_ring_base + uint32_t(sb) * uint32_t(p.b_bytes);
        mbar_expect_tx(b_full + sb, uint32_t(2 * n1 * KB_BYTES));
        tma_load_2d(b_dst, &tmap_an, b_full + sb, kb * KB_ELEMS, 0);                        // hi rows [0, n1)
        tma_load_2d(b_dst + uint32_t(n1 * KB_BYTES), &tmap_an, b_full + sb, kb * KB_ELEMS, a);   // lo rows stacked after
      }
      __syncwarp();
      if (++sa == A_STAGES) {
        sa = 0;
        pha ^= 1;
      }
      if (++sb == B_STAGES) {
        sb = 0;
        phb ^= 1;
      }
    }
    if (need_g) {
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        for (int kb = 0; kb < p.num_kb2; ++kb) {
          mbar_wait(b_empty + sb, phb ^ 1);
          if (elect_one()) {
            const uint32_t b_dst = b_ring_base + uint32_t(sb) * uint32_t(p.b_bytes);
            mbar_expect_tx(b_full + sb, uint32_t(2 * n2 * KB_BYTES));
            tma_load_2d(b_dst, &tmap_ant, b_full + sb, kb * KB_ELEMS, ch * n2);
            tma_load_2d(b_dst + uint32_t(n2 * KB_BYTES), &tmap_ant, b_full + sb, kb * KB_ELEMS, c + ch * n2);
          }
          __syncwarp();
          if (++sb == B_STAGES) {
            sb = 0;
            phb ^= 1;
          }
        }
      }
    }
  }

  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static int encode2d(EncodeTiledFn encode, CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint32_t box_rows) {
  const cuuint64_t gdim[2] = {cuuint64_t(inner), cuuint64_t(rows)};
  const cuuint64_t gstride[1] = {cuuint64_t(inner) * 4};
  const cuuint32_t box[2] = {cuuint32_t(KB_ELEMS), cuuint32_t(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return cr == CUDA_SUCCESS ? 0 : int(cr);
}

}  // namespace ctc

int clip_ce_tc_shape_ok(int c, int a) {
  return c >= 4 && c % 4 == 0 && a >= 4 && a % 4 == 0 && a <= ctc::A_MAX && c <= 65536;
}

int64_t clip_ce_tc_ws_elems(int c, int a) { return 4 * int64_t(a) * c; }

int clip_ce_tc(const float* feats, int64_t n, int c, const float* anchors_n, int a, const int64_t* labels,
               int64_t ignore_label, float* loss, float* grad_feats, int32_t* pred, float* grad_logits, float* ws,
               cudaStream_t stream) {
  using namespace ctc;
  if (!clip_ce_tc_shape_ok(c, a)) return LGS_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(feats) & 15) || (reinterpret_cast<uintptr_t>(ws) & 15) ||
      (reinterpret_cast<uintptr_t>(grad_feats) & 15) || (reinterpret_cast<uintptr_t>(grad_logits) & 15))
    return LGS_E_UNSUPPORTED;
  if (n >= (int64_t(1) << 31) - BM) return LGS_E_UNSUPPORTED;
  if (n == 0) return LGS_OK;
  EncodeTiledFn encode = get_encode();
  if (!encode) return fail(LGS_E_CUDA, "cuTensorMapEncodeTiled entry point not available");

  // anchors -> tensor-core operand forms (one launch): ws_an = [hi | lo] of An [a, c] (K-major B of phase 1),
  // ws_ant = [hi | lo] of An^T [c, a] (K-major B of phase 2)
  float* ws_an = ws;
  float* ws_ant = ws + 2 * int64_t(a) * c;
  const int rc = weight_prep(anchors_n, 1, a, c, 2, ws_ant, ws_an, LGS_F32, stream);
  if (rc != LGS_OK) return rc;

  Params p;
  p.feats = feats;
  p.n = n;
  p.c = c;
  p.a = a;
  p.labels = labels;
  p.ignore_label = ignore_label;
  p.loss = loss;
  p.grad_feats = grad_feats;
  p.pred = pred;
  p.grad_logits = grad_logits;
  p.n1 = ((a + 15) / 16) * 16;
  p.kp = ((a + 7) / 8) * 8;
  p.num_kb1 = (c + KB_ELEMS - 1) / KB_ELEMS;
  p.num_kb2 = (p.kp + KB_ELEMS - 1) / KB_ELEMS;
  p.n2 = ((c + 15) / 16) * 16 < N2_MAX ? ((c + 15) / 16) * 16 : N2_MAX;
  p.n_chunks = (c + p.n2 - 1) / p.n2;
  p.b_bytes = 2 * (p.n1 > p.n2 ? p.n1 : p.n2) * KB_BYTES;
  p.ta_col0 = ((p.n1 + 31) / 32) * 32;
  p.ta_stages = (512 - p.ta_col0) / 64 < TA_STAGES_MAX ? (512 - p.ta_col0) / 64 : TA_STAGES_MAX;
  p.acc2_col = 2 * p.n1;
  if (p.acc2_col + p.n2 > 512 || p.ta_stages < 2) return LGS_E_UNSUPPORTED;

  CUtensorMap tf, tan, tant;
  int er = encode2d(encode, &tf, feats, uint64_t(c), uint64_t(n), BM);
  if (!er) er = encode2d(encode, &tan, ws_an, uint64_t(c), uint64_t(2) * a, uint32_t(p.n1));
  if (!er) er = encode2d(encode, &tant, ws_ant, uint64_t(a), uint64_t(2) * c, uint32_t(p.n2));
  if (er) return fail(LGS_E_CUDA, "lgs_clip_ce_tc: cuTensorMapEncodeTiled failed (%d)", er);

  const size_t smem = size_t(A_STAGES) * A_BYTES + size_t(B_STAGES) * p.b_bytes +
                      (2 * A_STAGES + 2 * B_STAGES + 2 * TA_STAGES_MAX + 4) * 8 + 16 + 1024;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(clip_ce_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) return fail(LGS_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  LGS_LAUNCH(clip_ce_tc_kernel, unsigned(cdiv(n, BM)), THREADS, smem, stream, tf, tan, tant, p);
  return LGS_OK;
}

}  // namespace lgs

using namespace lgs;

extern "C" {

int lgs_clip_ce_tc_supported(int32_t c, int32_t a) { return clip_ce_tc_shape_ok(c, a); }

int64_t lgs_clip_ce_tc_ws_elems(int32_t c, int32_t a) { return clip_ce_tc_ws_elems(c, a); }

int lgs_clip_ce_tc(const float* d_feats, int64_t n, int32_t c, const float* d_anchors_n, int32_t a,
                   const int64_t* d_labels, int64_t ignore_label, float* d_loss, float* d_grad_feats, int32_t* d_pred,
                   float* d_grad_logits, float* d_ws, void* stream_) {
  LGS_TRACE("lgs_clip_ce_tc %p %lld %d %p %d %p %lld %p %p %p %p %p %p", (const void*)d_feats, (long long)n, (int)c, (const void*)d_anchors_n, (int)a, (const void*)d_labels, (long long)ignore_label, (const void*)d_loss, (const void*)d_grad_feats, (const void*)d_pred, (const void*)d_grad_logits, (const void*)d_ws, (const void*)stream_);
  if (n < 0 || c < 1 || a < 1) return fail(LGS_E_INVALID, "lgs_clip_ce_tc: bad sizes n=%lld c=%d a=%d", (long long)n, c, a);
  if (!clip_ce_tc_shape_ok(c, a))
    return fail(LGS_E_UNSUPPORTED, "lgs_clip_ce_tc: needs c %% 4 == 0, a %% 4 == 0, a <= %d (got c=%d a=%d); use lgs_clip_ce",
                ctc::A_MAX, c, a);
  if (n == 0) return LGS_OK;
  if (!d_feats || !d_anchors_n || !d_labels || !d_ws) return fail(LGS_E_INVALID, "lgs_clip_ce_tc: null pointer");
  const int rc = clip_ce_tc(d_feats, n, c, d_anchors_n, a, d_labels, ignore_label, d_loss, d_grad_feats, d_pred,
                            d_grad_logits, d_ws, static_cast<cudaStream_t>(stream_));
  if (rc == LGS_E_UNSUPPORTED) return fail(rc, "lgs_clip_ce_tc: buffers must be 16-byte aligned and n < 2^31");
  return rc;
}

}  // extern "C"
