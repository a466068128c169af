// Fused CLIP text-anchor losses.
//   clip_ce_kernel   : per 32-row tile  S = normalize(F) @ An^T  ->  softmax cross-entropy, argmax, and the
//                      gradient w.r.t. F in the same launch (the [n,a] logits never touch HBM unless asked for).
//   clip_hinge_kernel: one warp per point, 1 + n_neg anchor dot products, loss + gradient.
// Replaces lib/losses/ContrastiveLanguageLoss.py:224-237 / :97-194 and lib/losses/utils.py:99-103, which
// materialise (N,200,C) expansions and loop over classes on the host.
#include "common.cuh"

namespace lgs {

constexpr int CE_ROWS = 32;    // rows per CTA
constexpr int CE_CK = 32;      // channel chunk
constexpr int CE_AMAX = 256;   // anchors supported per launch
constexpr int CE_NJ = CE_AMAX / 16;
constexpr int CE_LD = CE_CK + 1;
constexpr int CE_GLD = CE_AMAX + 1;
constexpr float kNormEps = 1e-12f;  // F.normalize default eps

struct CeSmem {
  float Fs[CE_ROWS][CE_LD];
  float As[CE_AMAX][CE_LD];
  float Gs[CE_ROWS][CE_GLD];
  float inv_norm[CE_ROWS];
  float sdot[CE_ROWS];
};

__global__ void __launch_bounds__(256)
clip_ce_kernel(const float* __restrict__ feats, int64_t n, int c, const float* __restrict__ anchors, int a,
               const int64_t* __restrict__ labels, int64_t ignore_label, float* __restrict__ loss,
               float* __restrict__ grad_feats, int32_t* __restrict__ pred, float* __restrict__ grad_logits) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CeSmem& sm = *reinterpret_cast<CeSmem*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int tx = t & 15, ty = t >> 4;  // ty: 16 row pairs; tx: column residue mod 16
  const int64_t m0 = int64_t(blockIdx.x) * CE_ROWS;

  // ---- row norms: warp w handles rows w*4 .. w*4+3 ----
  for (int rr = 0; rr < 4; ++rr) {
    const int row = warp * 4 + rr;
    const int64_t i = m0 + row;
    float ss = 0.f;
    if (i < n)
      for (int ch = lane; ch < c; ch += 32) {
        const float v = __ldg(feats + size_t(i) * c + ch);
        ss = fmaf(v, v, ss);
      }
#pragma unroll
    for (int d = 16; d; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
    if (lane == 0) sm.inv_norm[row] = 1.f / fmaxf(sqrtf(ss), kNormEps);
  }

  // ---- phase 1: S tile (2 rows x 16 columns per thread) ----
  float acc[2][CE_NJ];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < CE_NJ; ++j) acc[i][j] = 0.f;

  for (int c0 = 0; c0 < c; c0 += CE_CK) {
    __syncthreads();
    for (int e = t; e < CE_ROWS * CE_CK; e += 256) {
      const int row = e / CE_CK, kk = e % CE_CK;
      const int64_t i = m0 + row;
      sm.Fs[row][kk] = (i < n && c0 + kk < c) ? __ldg(feats + size_t(i) * c + c0 + kk) : 0.f;
    }
    for (int e = t; e < CE_AMAX * CE_CK; e += 256) {
      const int col = e / CE_CK, kk = e % CE_CK;
      sm.As[col][kk] = (col < a && c0 + kk < c) ? __ldg(anchors + size_t(col) * c + c0 + kk) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < CE_CK; ++kk) {
      const float f0 = sm.Fs[ty * 2][kk], f1 = sm.Fs[ty * 2 + 1][kk];
#pragma unroll
      for (int j = 0; j < CE_NJ; ++j) {
        const float b = sm.As[tx + 16 * j][kk];
        acc[0][j] = fmaf(f0, b, acc[0][j]);
        acc[1][j] = fmaf(f1, b, acc[1][j]);
      }
    }
  }

  // ---- softmax cross-entropy per row (16 lanes share a row) ----
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = ty * 2 + i;
    const int64_t gi = m0 + row;
    const float inv = sm.inv_norm[row];
    const int64_t y = gi < n ? labels[gi] : ignore_label;
    const bool valid = gi < n && y != ignore_label;
    float mx = -INFINITY;
    int amax = 0;
#pragma unroll
    for (int j = 0; j < CE_NJ; ++j) {
      const int col = tx + 16 * j;
      acc[i][j] *= inv;
      if (col < a && acc[i][j] > mx) {
        mx = acc[i][j];
        amax = col;
      }
    }
#pragma unroll
    for (int d = 8; d; d >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, d);
      const int oa = __shfl_xor_sync(0xffffffffu, amax, d);
      if (om > mx || (om == mx && oa < amax)) {
        mx = om;
        amax = oa;
      }
    }
    float se = 0.f, sy = 0.f;
#pragma unroll
    for (int j = 0; j < CE_NJ; ++j) {
      const int col = tx + 16 * j;
      if (col < a) {
        se += __expf(acc[i][j] - mx);
        if (valid && col == y) sy = acc[i][j];
      }
    }
#pragma unroll
    for (int d = 8; d; d >>= 1) {
      se += __shfl_xor_sync(0xffffffffu, se, d);
      sy += __shfl_xor_sync(0xffffffffu, sy, d);
    }
    const float lse = logf(se) + mx;
    if (tx == 0 && gi < n) {
      if (loss) loss[gi] = valid ? (lse - sy) : 0.f;
      if (pred) pred[gi] = amax;
    }
    // g = softmax - onehot (0 for ignored rows); sdot = sum_j g_j S_j
    float sd = 0.f;
    const float inv_se = 1.f / se;
#pragma unroll
    for (int j = 0; j < CE_NJ; ++j) {
      const int col = tx + 16 * j;
      float g = 0.f;
      if (valid && col < a) {
        g = __expf(acc[i][j] - mx) * inv_se - (col == y ? 1.f : 0.f);
        sd = fmaf(g, acc[i][j], sd);
      }
      sm.Gs[row][col] = g;
      if (grad_logits && gi < n && col < a) grad_logits[size_t(gi) * a + col] = g;
    }
#pragma unroll
    for (int d = 8; d; d >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, d);
    if (tx == 0) sm.sdot[row] = sd;
  }
  if (!grad_feats) return;

  // ---- phase 2: dF = (G @ An - sdot * f_hat) * inv_norm ; thread -> channel lane, 4 rows ----
  const int kk = lane, rbase = warp * 4;
  for (int c0 = 0; c0 < c; c0 += CE_CK) {
    __syncthreads();
    for (int e = t; e < CE_AMAX * CE_CK; e += 256) {
      const int col = e / CE_CK, k2 = e % CE_CK;
      sm.As[col][k2] = (col < a && c0 + k2 < c) ? __ldg(anchors + size_t(col) * c + c0 + k2) : 0.f;
    }
    __syncthreads();
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < a; ++j) {
      const float b = sm.As[j][kk];
#pragma unroll
      for (int r = 0; r < 4; ++r) d[r] = fmaf(sm.Gs[rbase + r][j], b, d[r]);
    }
    const int ch = c0 + kk;
    if (ch < c) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int64_t gi = m0 + rbase + r;
        if (gi < n) {
          const float inv = sm.inv_norm[rbase + r];
          const float fh = __ldg(feats + size_t(gi) * c + ch) * inv;
          grad_feats[size_t(gi) * c + ch] = (d[r] - sm.sdot[rbase + r] * fh) * inv;
        }
      }
    }
  }
}

// one warp per point
__global__ void __launch_bounds__(256)
clip_hinge_kernel(const float* __restrict__ feats, int64_t n, int c, const float* __restrict__ anchors, int a,
                  const int64_t* __restrict__ labels, const int32_t* __restrict__ neg_ids, int n_neg,
                  int64_t ignore_label, float pos_thresh, float neg_thresh, float neg_weight,
                  float* __restrict__ pos_loss, float* __restrict__ neg_loss, float* __restrict__ grad_feats) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  if (i >= n) return;
  const int64_t y = labels[i];
  const bool valid = y != ignore_label && y >= 0 && y < a;
  const float* f = feats + size_t(i) * c;
  if (!valid) {
    if (lane == 0) {
      if (pos_loss) pos_loss[i] = fmaxf(-pos_thresh, 0.f);
      // reference: feat_dist zeroes the distance of ignored rows, then relu(neg_thresh - 0)
      if (neg_loss) neg_loss[i] = fmaxf(neg_thresh, 0.f);
    }
    if (grad_feats)
      for (int ch = lane; ch < c; ch += 32) grad_feats[size_t(i) * c + ch] = 0.f;
    return;
  }
  float ss = 0.f, dp = 0.f;
  const float* ay = anchors + size_t(y) * c;
  for (int ch = lane; ch < c; ch += 32) {
    const float v = __ldg(f + ch);
    ss = fmaf(v, v, ss);
    dp = fmaf(v, __ldg(ay + ch), dp);
  }
  float dn = 0.f;
  for (int j = 0; j < n_neg; ++j) {
    const float* an = anchors + size_t(__ldg(neg_ids + size_t(i) * n_neg + j)) * c;
    for (int ch = lane; ch < c; ch += 32) dn = fmaf(__ldg(f + ch), __ldg(an + ch), dn);
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, d);
    dp += __shfl_xor_sync(0xffffffffu, dp, d);
    dn += __shfl_xor_sync(0xffffffffu, dn, d);
  }
  const float inv = 1.f / fmaxf(sqrtf(ss), kNormEps);
  const float s_pos = dp * inv, s_neg = dn * inv / float(n_neg);  // mean cosine over negatives
  const float pl = (1.f - s_pos) - pos_thresh, nl = neg_thresh - (1.f - s_neg);
  if (lane == 0) {
    if (pos_loss) pos_loss[i] = fmaxf(pl, 0.f);
    if (neg_loss) neg_loss[i] = fmaxf(nl, 0.f);
  }
  if (!grad_feats) return;
  // d(loss_i)/dS: -1 on the positive if active; +neg_weight/n_neg on each negative if active
  const float cp = pl > 0.f ? -1.f : 0.f, cn = nl > 0.f ? neg_weight / float(n_neg) : 0.f;
  const float sdot = cp * s_pos + cn * float(n_neg) * s_neg;  // sum_j coef_j * S_ij
  for (int ch = lane; ch < c; ch += 32) {
    float d = cp * __ldg(ay + ch);
    for (int j = 0; j < n_neg; ++j)
      d = fmaf(cn, __ldg(anchors + size_t(__ldg(neg_ids + size_t(i) * n_neg + j)) * c + ch), d);
    grad_feats[size_t(i) * c + ch] = (d - sdot * __ldg(f + ch) * inv) * inv;
  }
}

}  // namespace lgs

using namespace lgs;

extern "C" {

int lgs_clip_ce(const float* d_feats, int64_t n, int32_t c, const float* d_anchors_n, int32_t a,
                const int64_t* d_labels, int64_t ignore_label, float* d_loss, float* d_grad_feats, int32_t* d_pred,
                float* d_grad_logits, void* stream_) {
  LGS_TRACE("lgs_clip_ce %p %lld %d %p %d %p %lld %p %p %p %p %p", (const void*)d_feats, (long long)n, (int)c, (const void*)d_anchors_n, (int)a, (const void*)d_labels, (long long)ignore_label, (const void*)d_loss, (const void*)d_grad_feats, (const void*)d_pred, (const void*)d_grad_logits, (const void*)stream_);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n < 0 || c < 1 || a < 1) return fail(LGS_E_INVALID, "lgs_clip_ce: bad sizes n=%lld c=%d a=%d", (long long)n, c, a);
  if (a > CE_AMAX) return fail(LGS_E_UNSUPPORTED, "lgs_clip_ce: at most %d anchors per call (got %d)", CE_AMAX, a);
  if (n == 0) return LGS_OK;
  static bool attr_set = false;
  if (!attr_set) {
    LGS_CUDA(cudaFuncSetAttribute(clip_ce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(CeSmem))));
    attr_set = true;
  }
  LGS_LAUNCH(clip_ce_kernel, unsigned(cdiv(n, CE_ROWS)), 256, sizeof(CeSmem), stream, d_feats, n, c, d_anchors_n, a,
             d_labels, ignore_label, d_loss, d_grad_feats, d_pred, d_grad_logits);
  return LGS_OK;
}

int lgs_clip_hinge(const float* d_feats, int64_t n, int32_t c, const float* d_anchors_n, int32_t a,
                   const int64_t* d_labels, const int32_t* d_neg_ids, int32_t n_neg, int64_t ignore_label,
                   float pos_thresh, float neg_thresh, float neg_weight, float* d_pos_loss, float* d_neg_loss,
                   float* d_grad_feats, void* stream_) {
  LGS_TRACE("lgs_clip_hinge %p %lld %d %p %d %p %p %d %lld %.9g %.9g %.9g %p %p %p %p", (const void*)d_feats, (long long)n, (int)c, (const void*)d_anchors_n, (int)a, (const void*)d_labels, (const void*)d_neg_ids, (int)n_neg, (long long)ignore_label, (double)pos_thresh, (double)neg_thresh, (double)neg_weight, (const void*)d_pos_loss, (const void*)d_neg_loss, (const void*)d_grad_feats, (const void*)stream_);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n < 0 || c < 1 || a < 1 || n_neg < 1) return fail(LGS_E_INVALID, "lgs_clip_hinge: bad sizes");
  if (n == 0) return LGS_OK;
  LGS_LAUNCH(clip_hinge_kernel, unsigned(cdiv(n * 32, 256)), 256, 0, stream, d_feats, n, c, d_anchors_n, a, d_labels,
             d_neg_ids, n_neg, ignore_label, pos_thresh, neg_thresh, neg_weight, d_pos_loss, d_neg_loss, d_grad_feats);
  return LGS_OK;
}

}  // extern "C"
