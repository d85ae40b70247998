// Host-callable launchers of the engine's kernels (internal; the public surface is
// include/text2loc_b200.h).  Every launcher enqueues on `st`, never synchronises, and returns
// cudaGetLastError() of its launches.  All pointers are device pointers.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace t2l {

struct Launches {  // running count for t2l_launch_count()
  int64_t n = 0;
};

// ---- linear.cu ---------------------------------------------------------------------------
enum LinearPath { kPathSimt = 0, kPathUmma = 1 };

struct Linear {
  const float* A; long lda;   // [M, K]
  const float* W; long ldw;   // [N, K] (torch Linear layout)
  const float* bias;          // [N] or null
  float* C; long ldc;         // [M, N]  (segmax: [M/32, N])
  int M, N, K;
  int act = 0;                // 0 none, 1 relu
  const float* residual = nullptr; long ldr = 0;  // added after act (store mode only)
  int residual_half = 0;      // residual points at __half data (ldr in elements)
  int segmax = 0;             // max over 32-row groups after relu
  const float* side = nullptr; long lds = 0;      // segmax: elementwise max with side[g, :]
  int round_out = 0;          // round stored values to tf32 (rna)
  int passes = 1;             // 3: A and W are [hi | lo] tf32 planes of logical width K (row pitch >= 2K):
                              //    hi*hi + lo*hi + hi*lo in one accumulator, fp32-level accuracy on the tensor cores
  int half_ops = 0;           // A and W point at __half data (lda, ldw in elements): tcgen05 kind::f16, fp32 accumulate
  int out_half = 0;           // C points at __half data (ldc in elements); values saturate at +-half_max
  float half_max = 65504.f;
  int reg_epilogue = 0;       // out_half + residual: force the register-staged epilogue instead of the TMA one (A/B switch)
  int split_out = 0;          // fp32 store mode without residual: C is [M, ldc >= 2N] and receives the tf32 [hi | lo] planes of the result
                              // (hi at column c, lo at column N + c): the A operand of a following passes=3 GEMM, no split kernel
};
// tf32 tensor-core path: needs K-major operands with 16-byte aligned rows (lda, ldw % 4 == 0),
// N % 32 == 0.  K tails are zero-filled by TMA.
cudaError_t linear_umma(const Linear& l, cudaStream_t st, Launches* lc);
// planes[r, 0:K] = rna_tf32(x[r]), planes[r, K:2K] = rna_tf32(x[r] - hi): operand of a passes=3 GEMM
cudaError_t split_tf32_planes(const float* x, long ldx, float* planes, int rows, int K, cudaStream_t st, Launches* lc);
// exact fp32 path: any shape.
cudaError_t linear_simt(const Linear& l, cudaStream_t st, Launches* lc);

// out[g, c] = max over rows g*32..g*32+31 of C (already >= 0), optionally max'ed with side[g, c]
cudaError_t segmax32(const float* C, long ldc, float* out, long ldo, const float* side, long lds, int groups, int N,
                     cudaStream_t st, Launches* lc);

// ---- geometry.cu (a1: fps + radius of pointnet2.py:26-30, position-only) --------------------
struct Geometry {           // per object, all levels
  uint8_t* fps1; uint8_t* fps2; uint8_t* fps3;   // [n,128] [n,64] [n,32] local indices into the level's dense set
  float* cpos1; float* cpos2; float* cpos3;      // [n,128,3] [n,64,3] [n,32,3] centroid positions
  uint8_t* nbr1; uint8_t* nbr2; uint8_t* nbr3;   // [n,128,32] [n,64,32] [n,32,32]
  uint8_t* cnt1; uint8_t* cnt2; uint8_t* cnt3;   // [n,128] [n,64] [n,32]
};
// dist_fma: evaluate squared distances with fused multiply-adds (common.cuh::sqdist) instead of one rounding per operation
cudaError_t fps_all_levels(const float* pts, int n_obj, const Geometry& g, bool dist_fma, cudaStream_t st, Launches* lc);
cudaError_t ball_query_all_levels(const float* pts, int n_obj, const Geometry& g, bool dist_fma, cudaStream_t st, Launches* lc);

// ---- pointnet.cu --------------------------------------------------------------------------
// sa_obj2.cu's per-point operand of SA1: Qx16[n*256, 32] = fp16(W1x . rgb + b1 + W1p . (pos - o)), o = the object's point 0
constexpr float kQxMax = 32752.f;  // Qx and v are bounded by half of fp16's range, so Qx - v never overflows
cudaError_t sa1_qx16(const float* pts, int n_obj, const float* w1x, const float* w1p, const float* b1, __half* qx16, cudaStream_t st, Launches* lc);
// columns C .. C+7 of x [n*P, ldx] = tf32 hi | lo split of (pos_r - pos of the object's row 0) | 0 0  (pos [n*P, 3])
cudaError_t append_pos_cols(const float* pos, int n_obj, int P, float* x, int ldx, int C, cudaStream_t st, Launches* lc);
// Fused PointConv layer (sa_obj2.cu): gather, first-layer edge activations, second Linear on the tensor cores and the per-centroid
// max in ONE kernel.  W2 resident in tensor memory (TS-mode MMAs), self-loop edges folded in as one extra tile per object, SA1
// tiles paired.  Needs the self-loop source of every object.
struct SaObj2 {
  const __half* Qx16; int C1; int C2;    // Qx16 = fp16 (W1x x_j + b1 + W1p (pos_j - o)), o = the object's point 0; |.| <= 32752
  const float* cpos;                     // [n*M, 3]
  const uint8_t* nbr; const uint8_t* cnt;
  const int32_t* loop_src_obj; const int32_t* loop_half;  // [n] source object / half of its dense points
  const float* Wp;                       // [C1,4]
  const __half* W2h; const float* b2;    // [C2, C1] fp16, [C2]
  float* out; int ldo;                   // [n*M, ldo], ldo >= C2
  int n_obj, P, M;
  int bisect = 0;                        // timing bisect (t2l_debug_sa_bisect): 1 = skip the accumulator drain, 2 = skip the MMAs; results invalid
};
cudaError_t sa_obj2(const SaObj2& a, cudaStream_t st, Launches* lc);
// GA input as fp16 rows of 264 halfs: A16[n*32, 264] = [x3 (256) | cpos3 (3) | 0 x 5] (K padded to a 16-byte multiple)
cudaError_t ga_concat_half(const float* x3, const float* cpos3, int n_obj, __half* A, cudaStream_t st, Launches* lc);

// ---- rowops.cu ----------------------------------------------------------------------------
// y[r, 0:d] (row pitch ldy) = x[r] / max(||x[r]||, 1e-12)      (F.normalize)
cudaError_t l2_normalize_rows(const float* x, long ldx, float* y, long ldy, int rows, int d, cudaStream_t st, Launches* lc);
// y = LayerNorm(x) * w + b, eps 1e-5, one warp per row
// (optional y_half: an fp16 copy of the output, the A operand of a following fp16 tensor-core GEMM)
// (optional y_planes [rows, 2d]: the tf32 hi | lo planes of the output, as split_tf32_planes would produce them)
cudaError_t layer_norm_rows(const float* x, float* y, const float* w, const float* b, int rows, int d, cudaStream_t st, Launches* lc,
                            __half* y_half = nullptr, float* y_planes = nullptr);
// the same on fp16 rows in and out (d = 1024): the token layer's fp16 residual stream
cudaError_t layer_norm_half_rows(const __half* x, __half* y, const float* w, const float* b, int rows, int d, cudaStream_t st, Launches* lc);
// y = fp16(x), saturating at +-65504; n % 4 == 0
cudaError_t to_half_rows(const float* x, __half* y, long n, cudaStream_t st, Launches* lc);
// Unmasked multi-head self attention on packed QKV rows [n_seq*S, 3d] (q | k | v), head h uses
// columns [h*hd, (h+1)*hd); out [n_seq*S, d].  softmax(q k^T / sqrt(hd)) v, fp32.
// out_planes (optional, [rows, 2d]): tf32 hi | lo planes of the output for a following passes=3 GEMM; `out` may then be nullptr
cudaError_t mha_small(const float* qkv, float* out, int n_seq, int S, int d, int n_heads, cudaStream_t st, Launches* lc, int round_out = 0,
                      float* out_planes = nullptr);
// Cross attention with the same core: q [n_seq*Sq, ldq], k / v [n_seq*Sk, ldkv] -> out [n_seq*Sq, d]; Sk <= 32, head dim 32 / 64 / 256
cudaError_t mha_cross_small(const float* q, long ldq, const float* k, const float* v, long ldkv, float* out, int n_seq, int Sq, int Sk, int d,
                            int n_heads, cudaStream_t st, Launches* lc, int round_out = 0, float* out_planes = nullptr);
// Same contract for d = 1024, 4 heads of 256 (the token layer): warp-level mma.sync tf32 tiles.
// round_out: 0 fp32, 1 fp32 rounded to tf32, 2 `out` is __half [rows, 1024].
// half_in: qkv is __half [rows, 3072].
cudaError_t mha_tc256(const void* qkv, float* out, int n_seq, int S, cudaStream_t st, Launches* lc, int round_out = 0, int half_in = 0);
// y[g, :] = max over the S rows of group g of LayerNorm(x[g*S + s, :]) (d = 1024): norm2 + max over tokens without
// materialising the normalised rows
cudaError_t layer_norm_max_rows(const float* x, float* y, const float* w, const float* b, int groups, int S, int d, cudaStream_t st,
                                Launches* lc);
cudaError_t layer_norm_max_half_rows(const __half* x, float* y, const float* w, const float* b, int groups, int S, int d, cudaStream_t st,
                                     Launches* lc);
// y[g, :] = max over the S rows of group g
cudaError_t max_over_rows(const float* x, float* y, int groups, int S, int d, cudaStream_t st, Launches* lc);
// Intra-cell attention without the duplicate padding rows (rowops.cu::mha_seq64_kernel): cell b owns packed rows
// row_ptr[b] .. row_ptr[b+1]) = its min(n_b, slots) objects + one row for all slots - n_b zero-padded slots
cudaError_t mha_cells64(const float* qkv, float* out, int n_cells, const int32_t* row_ptr_dev, const int32_t* cell_ptr_dev, int slots, int d,
                        int n_heads, cudaStream_t st, Launches* lc, float* out_planes = nullptr);
cudaError_t scatter_objects_ragged(const float* emb, const int32_t* cell_ptr_dev, const int32_t* row_ptr_dev, int n_cells, float* X, cudaStream_t st,
                                   Launches* lc);
cudaError_t max_over_rows_ragged(const float* x, const int32_t* row_ptr_dev, float* y, int n_cells, cudaStream_t st, Launches* lc);
// cat[:, d:4d] = [normalize(color_enc(mean rgb)) | normalize(pos_enc(centre)) | normalize(num_enc((count-mean)/std))], d = 256 or 128
// from meta [n, 7]; w1[i] [64, ld 4], b1[i] [64], w2[i] [256, 64], b2[i] [256] for i = colour, position, count
// (models/object_encoder.py:122-145)
cudaError_t side_encoders(const float* meta, int n_obj, const float* const* w1, const float* const* b1, const float* const* w2, const float* const* b2,
                          float* cat, int d, cudaStream_t st, Launches* lc);
// dst[p * G + g, :] = src[(idx ? idx[p] : p0 + p) * G + g, :] for p < n_groups, g < G (rows of d floats)
cudaError_t gather_row_groups(const float* src, const int32_t* idx, int p0, int n_groups, int G, int d, float* dst, cudaStream_t st, Launches* lc);
// y = a + b (elementwise)
cudaError_t add_rows(const float* a, const float* b, float* y, long n, cudaStream_t st, Launches* lc);
// text: [S, nq] row order helpers are not needed: rows are kept query-major (q*S + s)

// ---- search.cu ----------------------------------------------------------------------------
struct SearchDb {
  const float* D = nullptr;        // caller's fp32 rows [N, 256]
  __nv_bfloat16* planes = nullptr; // [N, 512]  hi | lo
  __half* plane16 = nullptr;       // [N, 256]  fp16(d * 2^-ex): rows scaled so that the largest norm lies in [0.5, 1)
  float* max_norm = nullptr;       // [1] max row norm (device)
  float* scale = nullptr;          // [1] 2^ex, undoes the scaling of plane16 (device)
  int64_t n_rows = 0, row_offset = 0;
};
struct SearchWork {
  __nv_bfloat16* q_planes;  // [nq_cap, 512]
  __half* q16;              // [nq_cap, 256] fp16 rows, each scaled by its own power of two
  float* q_scale;           // [nq_cap] 2^ex per query
  float* q_norm;            // [nq_cap]
  float* cand_score;        // [nq_cap, splits_cap, 16]
  int32_t* cand_idx;        // [nq_cap, splits_cap, 16]
  float* cand_thr;          // [nq_cap, splits_cap]
  int32_t* flags;           // [nq_cap]  2 = needs the exact rescan (second-pass buffer overflowed)
  // second pass for queries whose proof failed: collect every row whose approximate score can still reach the top-k
  int32_t* n_fail;          // [1]
  int32_t* fail_ids;        // [nq_cap]
  float* fail_thr;          // [nq_cap]   k-th exact score of pass 1 minus the error bound, rounded down
  __nv_bfloat16* q2_planes; // [nq_cap, 512] planes of the failed queries, compacted
  int32_t* cand2_idx;       // [nq_cap, kPass2Cap]
  int32_t* cand2_cnt;       // [nq_cap]
  int nq_cap, splits_cap;
};
constexpr int kPass2Cap = 256;  // candidates per failed query in the second pass (all splits together)
cudaError_t search_prepare_db(const SearchDb& db, cudaStream_t st, Launches* lc);
// first_pass_bf16x3: rank the candidates with the three-pass bf16 hi|lo product instead of the single fp16 pass (better when
// most score gaps are below the fp16 bound); the result is the same either way
cudaError_t search_topk(const SearchDb& db, const SearchWork& w, const float* Q, int nq, int k, int64_t* out_idx,
                        double* out_score, int32_t* out_n_fallback, bool first_pass_bf16x3, cudaStream_t st, Launches* lc);
cudaError_t search_topk_exact(const SearchDb& db, const float* Q, int nq, int k, int64_t* out_idx, double* out_score,
                              const int32_t* only_flagged, cudaStream_t st, Launches* lc);
// shard g's lists start at idx_all + g * shard_stride / score_all + g * shard_stride (elements)
cudaError_t merge_topk(const int64_t* idx_all, const double* score_all, long shard_stride, int n_shards, int nq, int k, int64_t* out_idx,
                       double* out_score, cudaStream_t st, Launches* lc);
// run := top-k of (run, new) by (score desc, row asc), in place; empty slots carry idx -1
cudaError_t merge_running_topk(int64_t* run_idx, double* run_score, const int64_t* new_idx, const double* new_score, int nq, int k,
                               cudaStream_t st, Launches* lc);

// ---- synthgen.cu ------------------------------------------------------------------------------
// counter-based synthetic objects [first_obj, first_obj + n_obj): pts [n_obj, 256, 6], meta [n_obj, 7]
cudaError_t synth_cells(uint64_t seed, long first_obj, long n_obj, float* pts, float* meta, cudaStream_t st, Launches* lc);


// ---- bookkeeping.cu -----------------------------------------------------------------------
// Per-query accuracy rows of eval_epoch / run_coarse (training/coarse.py:131-150, evaluation/utils.py:31-54).
struct TopkAccuracy {
  const int64_t* idx; int nq, k;    // [nq, k] retrieved database rows, best first (-1 = empty slot)
  const int64_t* target_row;        // [nq] database row of the query's own cell (-1: not in the database) or null
  const double* query_xy;           // [nq, 2] pose_w[0:2]
  const double* cell_xy;            // [N, 2] position credited to each database row (cell centre / predicted in-cell position)
  const int32_t* query_scene;       // [nq] scene code or null
  const int32_t* cell_scene;        // [N] scene code or null (null: no cross-scene masking)
  int32_t top_k[8]; int n_top;      // ascending list of k values (by value: the launch carries them)
  double threshs[8]; int n_thr;     // distance thresholds
  uint8_t* hit;                     // [nq, n_top] or null
  uint8_t* within;                  // [nq, n_top, n_thr] or null: min(dists[0:k]) <= thresh
  double* dists;                    // [nq, k] or null
};
cudaError_t topk_accuracy(const TopkAccuracy& a, cudaStream_t st, Launches* lc);

}  // namespace t2l
