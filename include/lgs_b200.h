/*
 * lgs_b200 — C ABI of the B200-native sparse-voxel convolution engine.
 *
 * Drop-in boundary: the reference (RozDavid/LanguageGroundedSemseg) enters this path only through
 * the Python names of the un-vendored package MinkowskiEngine 0.5.4; ME's own native boundary is the pybind module
 * MinkowskiEngineBackend._C.  Each entry point below names the ME backend call it replaces and the reference call
 * site (relative to /root/reference) that exercises it.
 *
 * Conventions
 *   - plain pointers and sizes, no torch types.  Every pointer prefixed d_ is DEVICE memory owned by the caller
 *     (PyTorch's caching allocator in the facade); the engine allocates nothing and keeps no state between calls,
 *     so a coordinate map is just the buffers the caller holds (coords + cuckoo table).
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it.  No thread-local CUDA state:
 *     forward runs on the Python thread, backward on PyTorch's autograd thread.
 *   - return 0 on success, negative LGS_E_* otherwise; lgs_last_error() gives a per-thread message.
 *   - functions taking a host out-pointer (h_*) synchronise `stream` before returning.
 */
#ifndef LGS_B200_H
#define LGS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGS_OK 0
#define LGS_E_INVALID (-1)      /* bad argument */
#define LGS_E_CUDA (-2)         /* CUDA runtime error */
#define LGS_E_RANGE (-3)        /* coordinate outside the packable range, see lgs_coord_limit() */
#define LGS_E_HASH_FULL (-4)    /* cuckoo eviction chain exceeded its bound: retry with a larger capacity */
#define LGS_E_UNSUPPORTED (-5)  /* shape / dtype / algo combination not built */

/* feature dtypes */
#define LGS_F32 0
#define LGS_BF16 1
/* conv algorithms */
#define LGS_ALGO_SIMT 0  /* fp32 FMA, exact: the parity anchor */
#define LGS_ALGO_TC 1    /* tcgen05 tensor cores, TMEM accumulators (TF32 for LGS_F32 features, BF16 for LGS_BF16) */
#define LGS_ALGO_TC3 2   /* tcgen05, 3xTF32 error-compensated products (hi*hi + lo*hi + hi*lo): 2^-21 products, LGS_F32 only */
#define LGS_ALGO_BX3 3   /* tcgen05, bf16x3 error-compensated products (bf16 hi/lo pairs, fp32 accumulate): 2^-16 products at half
                            the tensor-core time and weight traffic of LGS_ALGO_TC3; LGS_F32 features only */
/* weight layouts */
#define LGS_W_KCN 0      /* [K, c_in, c_out]: MinkowskiEngine's parameter layout */
#define LGS_W_KNC 1      /* [K, c_out, c_in]: per-offset transpose (the K-major B operand the tensor-core path loads by TMA) */
#define LGS_W_KNC_SPLIT 2 /* [2, K, c_out, c_in]: TF32 hi / lo halves of LGS_W_KNC, for LGS_ALGO_TC3 (see lgs_weight_prep) */
#define LGS_W_BX3 3      /* bf16 [K, c_out, ceil(c_in/32), 64]: per output channel and 32-channel block [hi x32 | lo x32],
                            zero-padded channels, for LGS_ALGO_BX3 (lgs_weight_prep with nsplit = 3) */

int lgs_version(void);
const char* lgs_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t lgs_launch_count(void);

/* Call recorder (host-side testing without a GPU).  Between lgs_trace_begin() and lgs_trace_end() the compute entry points
 * (weight prep, conv fwd / wgrad, bn fwd / bwd, seg_ce, kmap build / transpose, clip losses) append one text line
 * "name arg arg ..." each (pointers as %p) and return LGS_OK WITHOUT launching anything.  lgs_trace_end stops recording,
 * copies the NUL-terminated text into buf if capacity allows, and returns the number of bytes needed.  Process-wide. */
int lgs_trace_begin(void);
int64_t lgs_trace_end(char* buf, int64_t capacity);
/* 1 if the library was built with the tcgen05 path */
int lgs_has_tc(void);
/* Override a work-decomposition heuristic of the tensor-core kernels (tests and tuning; 0 restores the heuristic).
 * Keys: "bx3_tm" (row tiles per CTA), "bx3_rt" (rows per tile), "bx3_ks" (kernel-offset splits), "bx3_ns" (output-channel
 * slices), "bx3_sa" (gather ring depth), "bx3_no_balance".  Process-wide; the same knobs are read once from the environment
 * variables LGS_BX3_TM ... at first use. */
int lgs_tune(const char* key, int32_t value);

/* ---------------------------------------------------------------------------------------------------------
 * Coordinate maps.   Replaces ME CoordinateMapManagerGPU_c10::insert_and_map / ::stride
 *   call sites: SparseTensor(input, coords) lib/train_test/pl_BaselineTrainer.py:300;
 *               stride-2 convs models/res16unet.py:49,66,83,100; quantisation lib/voxelizer.py:142.
 * Coordinates are int32 [n,4] = (batch, x, y, z).  Keys are packed to 64 bit: batch in [0, 1023),
 * |x|,|y|,|z| < lgs_coord_limit().
 * --------------------------------------------------------------------------------------------------------- */
int32_t lgs_coord_limit(void);
/* slots (power of two, >= 2n) a cuckoo table for n keys needs; keys: uint64[cap], vals: int32[cap] */
int64_t lgs_hash_capacity(int64_t n);
/* int32 scratch elements lgs_coordmap_build needs for n rows */
int64_t lgs_coordmap_scratch_elems(int64_t n);

/* Build a coordinate map from n rows: each row is floored to a multiple of `quant` (the new tensor stride; 1 = keep),
 * inserted into the cuckoo table, duplicates collapse and the FIRST row (lowest index) wins.  Unique rows are
 * numbered in order of first occurrence — the order ME's CPU manager produces, so a
 * duplicate-free input keeps its row order.
 *   d_out_coords   [n,4]  first n_unique rows valid (quantised coordinates)
 *   d_unique_index [n]    first n_unique valid: input row each unique row came from (ascending)
 *   d_inverse      [n]    unique row of every input row
 *   d_n_unique     [1]    device copy of the count;  h_n_unique: host copy (stream is synchronised)        */
int lgs_coordmap_build(const int32_t* d_coords, int64_t n, int32_t quant,
                       uint64_t* d_table_keys, int32_t* d_table_vals, int64_t capacity,
                       int32_t* d_out_coords, int32_t* d_unique_index, int32_t* d_inverse,
                       int32_t* d_scratch, int32_t* d_n_unique, int64_t* h_n_unique, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Kernel maps.   Replaces ME CoordinateMapManagerGPU_c10::kernel_map (cached per (keys, ks, stride, dilation))
 *   call sites: every conv()/conv_tr() models/modules/common.py:195,228.
 * Output-stationary neighbour table: d_table[k*n_out + o] = row i of the INPUT map with
 * C_in[i] == C_out[o] + off_k, or -1.  off_k enumerates x fastest (k = ix + ks*iy + ks^2*iz); odd ks is centred,
 * even ks starts at 0; offsets are multiples of in_tensor_stride*dilation (App. A.5).  d_counts[k] = pairs of
 * offset k.  The ME pair list of offset k is {(table[k][o], o) : table[k][o] >= 0}.
 * --------------------------------------------------------------------------------------------------------- */
int lgs_kmap_build(const int32_t* d_out_coords, int64_t n_out,
                   const uint64_t* d_in_table_keys, const int32_t* d_in_table_vals, int64_t in_capacity,
                   int32_t ksize, int32_t in_tensor_stride, int32_t dilation,
                   int32_t* d_table, int32_t* d_counts, void* stream);
/* Transposed table: d_table_t[k*n_in + i] = o where d_table[k*n_out + o] == i, else -1 (each (k,i) has at most one o).
 * Used by MinkowskiConvolutionTranspose (common.py:228) and by dgrad of strided convs.                      */
int lgs_kmap_transpose(const int32_t* d_table, int32_t K, int64_t n_out, int64_t n_in, int32_t* d_table_t,
                       void* stream);

/* 1 if the tensor-core kernels (LGS_ALGO_TC / TC3) take a c_in -> c_out layer of this feature dtype */
int lgs_conv_tc_supported(int32_t c_in, int32_t c_out, int32_t dtype);

/* Tensor-core operand forms of one layer's weights W [K,c_in,c_out] (fp32 parameter), one launch:
 *   d_fwd [nsplit,K,c_out,c_in] for the forward GEMM, d_bwd [nsplit,K,c_in,c_out] for dgrad (either may be NULL).
 *   nsplit 1: cast/transposed copy (LGS_W_KNC);  nsplit 2: hi = RN_tf32(w), lo = RN_tf32(w - hi) (LGS_W_KNC_SPLIT).
 * Outputs have the feature dtype.  (ME keeps one fp32 [K,Cin,Cout] kernel and re-reads it per offset.) */
int lgs_weight_prep(const float* d_weight, int32_t K, int32_t c_in, int32_t c_out, int32_t nsplit,
                    void* d_fwd, void* d_bwd, int32_t dtype, void* stream);
/*   nsplit 3 (dtype LGS_F32): the LGS_W_BX3 operand forms, bf16 elements:
 *   d_fwd [K, c_out, ceil(c_in/32), 64], d_bwd [K, c_in, ceil(c_out/32), 64]; lgs_weight_bx3_elems(K, rows, reduced) each. */
int64_t lgs_weight_bx3_elems(int32_t K, int32_t c_rows, int32_t c_reduced);

/* The same for every layer of a network in ONE launch (63 per-layer launches per Res16UNet34C step otherwise).
 *   d_desc: device int64 [n_layers][8] = { d_weight, d_fwd, d_bwd (0 = skip), K, c_in, c_out, first_tile, 0 } where a
 *   layer owns K * ceil(c_in/32) * ceil(c_out/32) consecutive tiles from first_tile (ascending); total_tiles = their sum.
 *   nsplit / dtype as in lgs_weight_prep, common to all layers. */
int lgs_weight_prep_batch(const int64_t* d_desc, int32_t n_layers, int64_t total_tiles, int32_t nsplit, int32_t dtype,
                          void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Sparse convolution.   Replaces ME ConvolutionForwardGPU / ConvolutionBackwardGPU (and ...Transpose...)
 *   call sites: models/modules/common.py:195-203, 228-236; autograd backward of the same.
 *   out[o,:] = sum_k in[table[kk][o], :] @ W[k]  (+ bias),   kk = reverse_k ? K-1-k : k
 * d_table == NULL means the identity map with K == 1 (1x1x1 convs, models/resnet.py:95-101, res16unet.py:193).
 * dgrad is the same call on the transposed problem: in := grad_out, the SAME weight buffer read as LGS_W_KNC
 * (W[k] is [c_in,c_out] = [c_out',c_in']), table := transposed table, or the same table with reverse_k = 1 when the
 * in and out maps coincide and ks is odd.
 * W has the feature dtype (fp32 for LGS_F32, bf16 for LGS_BF16); out too; accumulation is fp32.
 * LGS_ALGO_SIMT reads LGS_W_KCN or LGS_W_KNC.  LGS_ALGO_TC needs LGS_W_KNC and a shape lgs_conv_tc_supported()
 * accepts; any other LGS_ALGO_TC request is served by the SIMT kernel (still on the GPU; there is no CPU path).
 * LGS_ALGO_TC3 needs LGS_W_KNC_SPLIT weights, LGS_F32 features and a supported shape, and fails otherwise.
 * --------------------------------------------------------------------------------------------------------- */
int lgs_conv_fwd(const void* d_in, int64_t n_in, int32_t c_in,
                 const void* d_weight, int32_t weight_layout, int32_t K, int32_t c_out,
                 const int32_t* d_table, int64_t n_out, int32_t reverse_k,
                 const float* d_bias, void* d_out, int32_t dtype, int32_t algo, void* stream);

/* LGS_ALGO_BX3 with its two extensions (fp32 features, LGS_W_BX3 weights of the full input width c_in + c_in2):
 *   - two gather sources: input channels [0, c_in) come from d_in [n_in, c_in], channels [c_in, c_in + c_in2) from d_in2
 *     [n_in, c_in2] (c_in % 32 == 0) — the convolution of cat(d_in, d_in2) without materialising the concatenation
 *     (ME.cat at models/res16unet.py:237,247,257,267 feeding block5..block8).  c_in2 == 0: one source.
 *   - d_bn_sums (may be NULL): BatchNorm accumulators [8][2][c_out] doubles (the d_scratch layout of lgs_bn_fwd, zero on
 *     entry) that receive the column sums and sums of squares of the OUTPUT from the accumulator registers, so the
 *     BatchNorm that follows (models/modules/common.py:17-19) needs no statistics pass: lgs_bn_fwd(..., stats_ready).  */
int lgs_conv_fwd2(const float* d_in, int32_t c_in, const float* d_in2, int32_t c_in2, int64_t n_in, const void* d_weight,
                  int32_t K, int32_t c_out, const int32_t* d_table, int64_t n_out, int32_t reverse_k, const float* d_bias,
                  float* d_out, double* d_bn_sums, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Neighbourhood plans (large same-map 3x3x3 kernel maps; replaces nothing in ME — its kernel maps are per-offset pair lists,
 * this is a second form of the same map for the cache-based kernel).   Call sites served: the 3x3x3 stride-1 convolutions of
 * the fine U-Net levels, models/modules/common.py:195-203 via models/modules/resnet_block.py:23-34 and res16unet.py:38.
 * A plan regroups the output rows of a kernel map into spatially compact supertiles (Morton order over coarse cells), lists
 * the unique input rows each supertile touches and rewrites the table in supertile-local indices (uint16: 12-bit index, 0xFFF =
 * no neighbour, | 3-bit bank colour << 12; uniq entries carry the row's colour << 28); lgs_conv_fwd3 then loads each
 * supertile's rows into shared memory once per channel block instead of once per (row, offset) pair.  step = tensor stride x
 * dilation of the map (the spacing of the kernel offsets, as in lgs_kmap_build).
 * Tensor row order is unchanged; results equal lgs_conv_fwd2's (same products, same accumulation order per row).
 *   lgs_nbplan_supported: 1 if a plan is worth building for this map (K == 27, n_out >= 256 unless tuned).  On maps with fewer
 *   supertiles than SMs lgs_conv_fwd3 splits the reduction (channel blocks, then kernel offsets) over CTAs; partial sums meet
 *   through red.global.add on a zeroed output (not when d_bn_sums is given).
 *   lgs_nbplan_build: d_plan lgs_nbplan_bytes() bytes, d_scratch lgs_nbplan_scratch_bytes() bytes (both 16-byte aligned);
 *   h_status (host, may be NULL; synchronises the stream when given): [0] = 1 if some supertile touches more unique rows than
 *   the cache holds (the plan must then NOT be used), [1] = largest unique-row count of a supertile.
 *   lgs_tune keys: "nb_rt" rows per tile, "nb_umax" cache rows, "nb_min_rows", "nb_off", "nb_target_ctas", "nb_no_split".
 * --------------------------------------------------------------------------------------------------------- */
int lgs_nbplan_supported(int64_t n_out, int32_t K);
int64_t lgs_nbplan_bytes(int64_t n_out, int32_t K);
int64_t lgs_nbplan_scratch_bytes(int64_t n_out);
int lgs_nbplan_build(const int32_t* d_out_coords, int64_t n_out, const int32_t* d_table, int32_t K, int32_t step, void* d_plan,
                     void* d_scratch, int32_t* h_status, void* stream);
/* out[0..8] = tm, rt, RS, S, umax, then the int32-word offsets of order [S][RS], ucount [S], uniq [S][umax] and loc
 * (uint16 [S][K][RS]) inside the plan buffer (tests, tools) */
int lgs_nbplan_geometry(int64_t n_out, int32_t K, int64_t* out);
/* lgs_conv_fwd2 with an optional plan of (d_table, n_out): with d_plan != NULL, n_in == n_out and a supported shape the
 * neighbourhood-cache kernel runs; otherwise exactly lgs_conv_fwd2. */
int lgs_conv_fwd3(const float* d_in, int32_t c_in, const float* d_in2, int32_t c_in2, int64_t n_in, const void* d_weight,
                  int32_t K, int32_t c_out, const int32_t* d_table, const void* d_plan, int64_t n_out, int32_t reverse_k,
                  const float* d_bias, float* d_out, double* d_bn_sums, void* stream);
/* lgs_conv_fwd3 + d_addend: out = conv(in) (+ bias) + addend, addend fp32 [n_out, c_out] or NULL — the dgrad of a residual
 * block's first convolution plus the gradient of the identity path (models/modules/resnet_block.py:41-57 backward) without a
 * separate add pass.  Fused into the neighbourhood-cache kernel's epilogue; any other path runs lgs_conv_fwd2 + lgs_add in
 * place.  d_bn_sums and d_addend exclude each other. */
int lgs_conv_fwd4(const float* d_in, int32_t c_in, const float* d_in2, int32_t c_in2, int64_t n_in, const void* d_weight,
                  int32_t K, int32_t c_out, const int32_t* d_table, const void* d_plan, int64_t n_out, int32_t reverse_k,
                  const float* d_bias, const float* d_addend, float* d_out, double* d_bn_sums, void* stream);

/* grad_w[k] = sum_o in[table[k][o], :]^T (outer) grad_out[o, :]   -> fp32 [K,c_in,c_out], overwritten. */
int lgs_conv_wgrad(const void* d_in, int64_t n_in, int32_t c_in,
                   const void* d_grad_out, int64_t n_out, int32_t c_out,
                   const int32_t* d_table, int32_t K,
                   float* d_grad_w, int32_t dtype, int32_t algo, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fused training-mode BatchNorm (+ residual add) (+ ReLU) on feature rows [n,c] fp32, c % 4 == 0.
 *   Replaces the ATen kernels under ME.MinkowskiBatchNorm (= nn.BatchNorm1d on .F, models/modules/common.py:17-19),
 *   MinkowskiReLU (models/res16unet.py:194) and `out += residual` (models/modules/resnet_block.py:54)   [SURVEY §8 f-1]
 *   fwd: z = relu?( gamma * (x - mean_batch) / sqrt(var_batch + eps) + beta [+ residual] ); running stats (may be NULL)
 *        updated in place with `momentum` (biased variance normalises, unbiased variance feeds the running estimate).
 *   bwd: dy = dz * (z > 0)?;  dx (BatchNorm backward through the batch statistics), d_residual = dy (may be NULL),
 *        dgamma, dbeta.   d_scratch: 16c doubles (8 interleaved copies of the 2c column sums).
 *   d_scratch_next == NULL: the call clears d_scratch itself (one memset node).  d_scratch_next != NULL (16384 doubles):
 *   the caller guarantees d_scratch is all zero on entry and the call leaves d_scratch_next all zero on exit — callers
 *   alternate two scratch halves per stream, so a chain of BatchNorm calls needs no memset nodes at all.
 *   d_num_batches_tracked (int64, may be NULL) is incremented by the forward kernel (nn.BatchNorm1d bookkeeping).
 * --------------------------------------------------------------------------------------------------------- */
int lgs_bn_fwd(const float* d_x, const float* d_residual, int64_t n, int32_t c, const float* d_gamma, const float* d_beta,
               float eps, float momentum, int32_t relu, float* d_running_mean, float* d_running_var, float* d_z,
               float* d_save_mean, float* d_save_invstd, double* d_scratch, double* d_scratch_next,
               int64_t* d_num_batches_tracked, void* stream);
/* lgs_bn_fwd with stats_ready != 0: d_scratch already holds the column sums / sums of squares of d_x in the accumulator
 * layout ([8][2][c] doubles; lgs_conv_fwd2 writes them from the convolution's epilogue), so only the apply kernel runs. */
int lgs_bn_fwd2(const float* d_x, const float* d_residual, int64_t n, int32_t c, const float* d_gamma, const float* d_beta,
                float eps, float momentum, int32_t relu, float* d_running_mean, float* d_running_var, float* d_z,
                float* d_save_mean, float* d_save_invstd, double* d_scratch, double* d_scratch_next,
                int64_t* d_num_batches_tracked, int32_t stats_ready, void* stream);
int lgs_bn_bwd(const float* d_x, const float* d_z, const float* d_dz, int64_t n, int32_t c, const float* d_gamma,
               const float* d_save_mean, const float* d_save_invstd, int32_t relu, float* d_dx, float* d_dresidual,
               float* d_dgamma, float* d_dbeta, double* d_scratch, double* d_scratch_next, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fused softmax cross-entropy over per-point class logits [n,c] fp32 (c % 4 == 0, c <= 1024), mean over the points
 * whose label != ignore_label; forward and d loss / d logits in one pass (two launches, no memset).
 *   Replaces nn.CrossEntropyLoss(ignore_index=...) at lib/train_test/pl_BaselineTrainer.py:343,350 (five ATen passes over
 *   the logit matrix).  d_ws: 4 doubles of workspace; d_loss: one float; d_grad_logits [n,c] may be NULL (evaluation).
 * --------------------------------------------------------------------------------------------------------- */
int lgs_seg_ce_supported(int32_t c);
int lgs_seg_ce(const float* d_logits, int64_t n, int32_t c, const int64_t* d_labels, int64_t ignore_label,
               double* d_ws, float* d_loss, float* d_grad_logits, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * CLIP text-anchor loss.   Replaces lib/losses/ContrastiveLanguageLoss.py:224-237 (+ feat_dist :206-222) and
 * lib/losses/utils.py:99-103 (feature_sim argmax), fused:
 *   S = normalize(F) @ An^T  (An already L2-normalised, [a,c]);  loss_i = CE(S_i, y_i), 0 where y_i == ignore;
 *   d_grad_feats[i] = d loss_i / d F_i;  d_pred[i] = argmax_j S_ij;  optional d_grad_logits [n,a] = softmax - onehot
 *   (for the learned anchor projection, models/clip_models.py:197-200).  Any output pointer may be NULL.
 * --------------------------------------------------------------------------------------------------------- */
int lgs_clip_ce(const float* d_feats, int64_t n, int32_t c, const float* d_anchors_n, int32_t a,
                const int64_t* d_labels, int64_t ignore_label,
                float* d_loss, float* d_grad_feats, int32_t* d_pred, float* d_grad_logits, void* stream);

/* The same loss on the tensor cores: S = F @ An^T as a 3xTF32 tcgen05 GEMM (fp32-grade products, fp32 accumulation in TMEM)
 * with the row norms, softmax cross-entropy, argmax and G = softmax - onehot fused into the TMEM epilogue, and
 * dF = (G @ An - (G.S) F_hat) / |F| as a second tcgen05 GEMM whose A operand (G) never leaves TMEM.
 * Shapes: c % 4 == 0, a % 4 == 0, a <= 208 (lgs_clip_ce_tc_supported); other shapes -> lgs_clip_ce (same results).
 * d_ws: caller-allocated scratch of lgs_clip_ce_tc_ws_elems(c, a) floats (hi/lo-split anchors, both operand forms). */
int lgs_clip_ce_tc_supported(int32_t c, int32_t a);
int64_t lgs_clip_ce_tc_ws_elems(int32_t c, int32_t a);
int lgs_clip_ce_tc(const float* d_feats, int64_t n, int32_t c, const float* d_anchors_n, int32_t a,
                   const int64_t* d_labels, int64_t ignore_label,
                   float* d_loss, float* d_grad_feats, int32_t* d_pred, float* d_grad_logits, float* d_ws, void* stream);

/* Hinge variant (ContrastiveLanguageLoss.py:184-192, 'cos' distance :87-93): negatives' anchor ids are an input
 * [n,n_neg] (the reference draws them on the host, :131-138).
 *   pos_i = relu(1 - S[i,y_i] - pos_thresh);  neg_i = relu(neg_thresh - (1 - mean_j S[i,neg_ij]));  0 where ignored. */
int lgs_clip_hinge(const float* d_feats, int64_t n, int32_t c, const float* d_anchors_n, int32_t a,
                   const int64_t* d_labels, const int32_t* d_neg_ids, int32_t n_neg, int64_t ignore_label,
                   float pos_thresh, float neg_thresh, float neg_weight,
                   float* d_pos_loss, float* d_neg_loss, float* d_grad_feats, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Voxelisation.   Replaces lib/voxelizer.py:138-139 (float64 affine + floor); the de-duplication of :142
 * (ME.utils.sparse_quantize) is lgs_coordmap_build with quant = 1 on the result.
 *   d_coords[i] = (batch, floor(((x*M[j][0] + y*M[j][1]) + z*M[j][2]) + M[j][3]) for j = 0..2), float64, no FMA.
 * h_M: 12 doubles (row-major 3x4) on the HOST.
 * --------------------------------------------------------------------------------------------------------- */
int lgs_voxelize_affine(const float* d_xyz, int64_t n, const double* h_M, int32_t batch, int32_t* d_coords,
                        void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Row-wise helpers (fp32): strided 2-D copy (ME.cat at models/res16unet.py:237-267 and its backward split, channel padding),
 * elementwise sum of two gradient matrices, column sums (bias gradient of `final`, models/res16unet.py:193).
 * --------------------------------------------------------------------------------------------------------- */
int lgs_copy2d(const float* d_src, int64_t src_ld, float* d_dst, int64_t dst_ld, int64_t rows, int32_t cols, void* stream);
int lgs_add(const float* d_a, const float* d_b, float* d_out, int64_t n, void* stream);
int lgs_colsum(const float* d_g, int64_t rows, int32_t c, float* d_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Native step driver.   Replaces the per-layer Python dispatch of one training step: forward + loss + backward of the
 * reference trainer (lib/train_test/pl_BaselineTrainer.py:157-160, 288-309; the network's forward at
 * models/res16unet.py:196-270) become ONE call that walks a straight-line program of the entry points above.
 *   ops   int64 [n_ops][LGS_PROGRAM_OP_WORDS]: word 0 = LGS_OP_*, the others are that op's arguments — buffer ids, channel
 *         counts, level ids, flags (languagegroundedsemseg_b200/program.py builds them from the network's modules and
 *         documents every word; csrc/program.cu is the interpreter).
 *   bufs  int64 [n_bufs][4] = { kind, level | slot, channels, element bytes }: kind 0 = external pointer taken from
 *         ext[slot] at run time (parameters, gradients, BatchNorm buffers, weight operands, kernel-map tables, labels,
 *         loss), kind 1 = intermediate of rows(level) x channels elements placed in the arena (level -1: one row).
 * lgs_program_run executes ops [op_begin, op_end) — ranges let the caller launch a gradient all-reduce bucket between
 * two parts of backward — on `stream`, weight gradients flagged for it on `side_stream` (joined before the call returns).
 *   level_rows [n_levels] voxels per U-Net level of THIS batch;  d_arena: lgs_program_arena_bytes() bytes of scratch;
 *   d_bn_scratch: 2 x 16384 doubles, all zero before the first run (the program alternates the halves like the facade).
 * --------------------------------------------------------------------------------------------------------- */
#define LGS_PROGRAM_OP_WORDS 18
#define LGS_OP_WEIGHT_PREP 1
#define LGS_OP_CONV 2
#define LGS_OP_WGRAD 3
#define LGS_OP_BN_FWD 4
#define LGS_OP_BN_BWD 5
#define LGS_OP_COPY2D 6
#define LGS_OP_ADD 7
#define LGS_OP_SEG_CE 8
#define LGS_OP_COLSUM 9
#define LGS_OP_JOIN 10
typedef struct lgs_program lgs_program;
int lgs_program_create(const int64_t* ops, int32_t n_ops, const int64_t* bufs, int32_t n_bufs, int32_t n_levels, int32_t n_ext,
                       lgs_program** out);
void lgs_program_destroy(lgs_program* p);
int64_t lgs_program_arena_bytes(const lgs_program* p, const int64_t* level_rows);
/* forget the BatchNorm scratch alternation state (after the caller re-zeroed the scratch, e.g. following a failed run) */
void lgs_program_reset(lgs_program* p);
int lgs_program_run(lgs_program* p, int32_t op_begin, int32_t op_end, const int64_t* level_rows, void* const* ext,
                    void* d_arena, int64_t arena_bytes, void* d_bn_scratch, void* stream, void* side_stream);
/* lgs_program_run with flags: LGS_RUN_NO_JOIN (1) = leave the side stream unjoined when the range ends (the caller makes its
 * collective wait on the side stream itself and runs the last range with flags 0, which joins). */
#define LGS_RUN_NO_JOIN 1
int lgs_program_run2(lgs_program* p, int32_t op_begin, int32_t op_end, const int64_t* level_rows, void* const* ext,
                    void* d_arena, int64_t arena_bytes, void* d_bn_scratch, void* stream, void* side_stream, int32_t flags);

#ifdef __cplusplus
}
#endif
#endif /* LGS_B200_H */
