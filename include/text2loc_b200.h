/* text2loc_b200 -- C ABI of the B200-native coarse cell-retrieval engine.
 *
 * Drop-in boundary for Text2Loc's global place-recognition path.  The reference has no FFI;
 * its seam is two Python methods and two functions.  Each entry point
 * below names the reference code it replaces; the Python mirror in text2loc_b200/ binds these
 * with ctypes (INTEGRATION.md shows the binding a reference maintainer would add).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; t2l_last_error() gives the text.
 *     Nothing throws across the ABI.
 *   - data buffers are CALLER-OWNED DEVICE memory unless a parameter says "host"; the engine
 *     owns weights, the prepared database planes and its workspace.
 *   - all work is enqueued on the caller's cudaStream_t (passed as void*).  Device-wide synchronisation
 *     happens only in t2l_create, t2l_finalize_weights, t2l_destroy and when a call needs MORE workspace
 *     than any call before it (the arena is re-allocated: cudaFree synchronises); t2l_reserve sizes the
 *     workspace up front so that steady-state calls never do.
 *   - an engine is bound to one device and is not thread-safe; distinct engines are independent.
 *     Every entry point switches to the engine's device and restores the caller's current device.
 *   - the workspace is shared by all calls of an engine: calls issued on one stream are ordered by the
 *     stream; a call on a different stream first waits (cudaStreamWaitEvent) for the previous call.
 *   - sm_100a only.  There is no CPU or other fallback: on a non-Blackwell device t2l_create fails.
 */
#ifndef TEXT2LOC_B200_H
#define TEXT2LOC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct t2l_engine t2l_engine;

#define T2L_EMBED_DIM 256    /* coarse_embed_dim, evaluation/args.py:55 */
#define T2L_T5_DIM 1024      /* t5-large d_model, README.md:127 */
#define T2L_NUM_POINTS 256   /* pointnet_numpoints, evaluation/args.py:58 */
#define T2L_OBJECT_SLOTS 28  /* object_size, evaluation/args.py:68 */
#define T2L_FINE_DIM 128     /* fine_embed_dim, evaluation/args.py:41 */
#define T2L_MAX_TOPK 12      /* fast path; the reference uses max(top_k) = 10, evaluation/args.py:20 */

/* Engine for CUDA device `device`.  Replaces CellRetrievalNetwork.__init__ + .to(device)
 * (models/cell_retrieval.py:14-54, evaluation/coarse.py:118-124). */
int t2l_create(int device, t2l_engine** out);
void t2l_destroy(t2l_engine* e);
const char* t2l_last_error(const t2l_engine* e); /* e may be NULL: error of the last failed t2l_create */
int t2l_version(void);

/* Weights.  Replaces load_state_dict(strict=False) (evaluation/coarse.py:123).  `name` is one of
 * the folded-layer names listed in text2loc_b200/weights.py (BatchNorm already folded into the
 * preceding Linear on the host, eval-mode affine: models/language_encoder.py:28-31); `data` is a
 * HOST fp32 row-major [rows, cols] array.  t2l_finalize_weights uploads derived copies (tf32-
 * rounded operands) and synchronises the device once. */
int t2l_set_weight(t2l_engine* e, const char* name, const float* data_host, int rows, int cols);
int t2l_finalize_weights(t2l_engine* e);

/* Optional: size the engine's workspace once for the largest calls to come (objects / cells per t2l_encode_cells call,
 * sentences x tokens per text call, queries per search), so that no later call re-allocates.  Synchronises the device. */
int t2l_reserve(t2l_engine* e, int max_objects, int max_cells, int max_sentences, int max_tokens_per_sentence, int max_queries);

/* CellRetrievalNetwork.encode_objects (models/cell_retrieval.py:65-110) = PointNet2 features2
 * (models/pointcloud/pointnet2.py:80-90) -> ObjectEncoder.forward (models/object_encoder.py:92-149)
 * -> intra-cell attention, max over slots, L2 normalise.
 *   pts          device f32 [n_objects, 256, 6]   xyz | rgb of each object's 256-sample (16-byte aligned)
 *   meta         device f32 [n_objects, 7]        mean rgb | centre | raw point count
 *   cell_ptr     HOST   i32 [n_cells + 1]         object range of each cell
 *   out          device f32 [n_cells, 256]        unit rows */
int t2l_encode_cells(t2l_engine* e, const float* pts, const float* meta, const int32_t* cell_ptr_host,
                     int n_cells, float* out, void* stream);

/* Intermediate of the same call, for parity tests: PointNet2.forward(...).features2 and the FPS /
 * ball-query index sets (any pointer may be NULL).
 *   features2 device f32 [n_objects, 256];  fps1/2/3 device u8 [n_objects, 128|64|32] local indices
 *   nbr1/2/3  device u8 [n_objects, 128|64|32, 32] (slots >= cnt undefined);  cnt1/2/3 device u8 */
int t2l_encode_objects_debug(t2l_engine* e, const float* pts, const int32_t* cell_ptr_host, int n_cells,
                             float* features2, uint8_t* fps1, uint8_t* fps2, uint8_t* fps3,
                             uint8_t* nbr1, uint8_t* nbr2, uint8_t* nbr3,
                             uint8_t* cnt1, uint8_t* cnt2, uint8_t* cnt3, void* stream);

/* CellRetrievalNetwork.encode_text after the frozen T5 (models/language_encoder.py:125-148,
 * models/cell_retrieval.py:57-63).
 *   t5   device f32 [n_queries * n_sent, n_tok, 1024]  T5 last_hidden_state (pads included, unmasked)
 *   out  device f32 [n_queries, 256]                   unit rows */
int t2l_encode_text(t2l_engine* e, const float* t5, int n_queries, int n_sent, int n_tok, float* out, void* stream);

/* The same in two stages, so a caller can stream token features in chunks (H2D overlapped with compute)
 * and run the sentence stage once over everything:
 *   tokens:    t5 device f32 [n_sentences, n_tok, 1024] -> pooled device f32 [n_sentences, 1024]
 *              (intra_module + max over tokens, models/language_encoder.py:130-133)
 *   sentences: pooled [n_queries * n_sent, 1024] -> out [n_queries, 256]
 *              (inter_mlp, inter_module with `x += layer(x)`, max over sentences, normalise, :137-148) */
int t2l_encode_text_tokens(t2l_engine* e, const float* t5, int n_sentences, int n_tok, float* pooled, void* stream);
int t2l_encode_text_sentences(t2l_engine* e, const float* pooled, int n_queries, int n_sent, float* out, void* stream);
/* Token stage on T5 states delivered as fp16 (device, raw 16-bit words, [n_sentences, n_tok, 1024], 16-byte aligned): half the
 * bytes to ship and no conversion kernel.  The token layer computes on fp16 operand copies either way; here the residual of
 * its out-projection is read from the fp16 input as well. */
int t2l_encode_text_tokens_f16(t2l_engine* e, const void* t5_half, int n_sentences, int n_tok, float* pooled, void* stream);

/* ---- fine stage: CrossMatch.forward (models/cross_matcher.py:83-129; evaluation/pipeline.py:113-116 calls it once per query) ----
 * An engine serves the fine stage when the weights set on it are CrossMatch's (text2loc_b200/weights.py detects the state
 * dict by its offset MLP): ObjectEncoder and LanguageEncoder(is_fine) at d = T2L_FINE_DIM, two cascaded pairs of
 * TransformerDecoderLayers, the offset MLP.  The coarse entry points then fail, and vice versa.
 *
 * t2l_fine_offsets = CrossMatch.forward for a batch of (cell, description) pairs, every cell padded / cut to the same number of
 * objects (pad_size):  pts / meta / cell_ptr_host as t2l_encode_cells;  t5 device f32 [n_cells * n_hints, n_tok, 1024] = T5
 * states of each pair's hint sentences;  offsets device f32 [n_cells, 2].
 * The three stages are exposed separately so that a caller can encode every database cell's objects ONCE and match many
 * (query, cell) pairs against them (run_fine batched over queries):
 *   t2l_fine_encode_objects  -> obj_emb device f32 [n_objects, 128]   F.normalize(ObjectEncoder(...)) (:98-108)
 *   t2l_fine_encode_hints    -> hints device f32 [n_sentences, 128]   LanguageEncoder(is_fine) after T5 (language_encoder.py:130-140)
 *   t2l_fine_match           pair p uses the n_obj rows of cell pair_cell[p] and the n_hints rows of query pair_query[p]
 *                            (device i32 [n_pairs]; NULL = identity) -> offsets [n_pairs, 2] (:113-127) */
int t2l_fine_offsets(t2l_engine* e, const float* pts, const float* meta, const int32_t* cell_ptr_host, int n_cells, const float* t5,
                     int n_hints, int n_tok, float* offsets, void* stream);
int t2l_fine_encode_objects(t2l_engine* e, const float* pts, const float* meta, const int32_t* cell_ptr_host, int n_cells,
                            float* obj_emb, void* stream);
int t2l_fine_encode_hints(t2l_engine* e, const float* t5, int n_sentences, int n_tok, float* hints, void* stream);
int t2l_fine_match(t2l_engine* e, const float* obj_emb, const int32_t* pair_cell, const float* hints, const int32_t* pair_query,
                   int n_pairs, int n_obj, int n_hints, float* offsets, void* stream);

/* Database side of eval_epoch's search loop (training/coarse.py:81-84,105-113): registers this
 * rank's shard of cell embeddings.  D device f32 [n_rows, 256]; the engine keeps a reference to D
 * (it must stay alive and unchanged) and builds its bf16 hi/lo operand planes.  row_offset is
 * added to returned indices (global row id of local row 0). */
int t2l_db_build(t2l_engine* e, const float* D, int64_t n_rows, int64_t row_offset, void* stream);

/* eval_epoch's per-query `scores = D @ q; argsort(-scores)[:k]` (training/coarse.py:119-125) for a
 * whole query batch: tensor-core candidate pass, exact fp64 re-rank, margin proof, exact rescan
 * of the queries whose proof fails.  Order: score descending, row index ascending on ties.
 *   Q device f32 [nq, 256];  out_idx device i64 [nq, k];  out_score device f64 [nq, k]
 *   out_n_fallback device i32 [1] (may be NULL): queries that needed the exact rescan
 * If the shard has fewer than k rows the tail is filled with idx -1 / score -inf. */
int t2l_search_topk(t2l_engine* e, const float* Q, int nq, int k, int64_t* out_idx, double* out_score,
                    int32_t* out_n_fallback, void* stream);

/* Streamed databases (BASELINE configs[3]; the reference's encode loop training/coarse.py:105-113 followed by :119-125,
 * without ever holding the whole database): search the registered shard and fold its top-k into the caller's running
 * lists IN PLACE.  run_idx device i64 [nq, k] / run_score device f64 [nq, k], initialised by the caller to -1 / -inf.
 * After the last chunk the running lists equal t2l_search_topk over the concatenated database, bit for bit (same fp64
 * scores, same (score desc, row asc) order). */
int t2l_search_topk_accumulate(t2l_engine* e, const float* Q, int nq, int k, int64_t* run_idx, double* run_score,
                               int32_t* out_n_fallback, void* stream);

/* Bench / test input tooling: counter-based synthetic cells [first_cell, first_cell + n_cells) x obj_per_cell objects x
 * 256 points, generated on the device (BASELINE configs[3]'s 98 GB of points never exist at once).  Every value is a pure
 * function of (seed, global object index): any chunking gives the same bytes; oracle/synthgen.py restates it in numpy.
 *   pts device f32 [n_cells * obj_per_cell, 256, 6], meta device f32 [n_cells * obj_per_cell, 7] */
int t2l_synth_cells(t2l_engine* e, uint64_t seed, int64_t first_cell, int n_cells, int obj_per_cell, float* pts, float* meta,
                    void* stream);

/* Exact fp64 scan only (no tensor-core pass); same contract.  Used for k > T2L_MAX_TOPK. */
int t2l_search_topk_exact(t2l_engine* e, const float* Q, int nq, int k, int64_t* out_idx, double* out_score, void* stream);

/* Merge of per-shard top-k lists after the all-gather:
 *   idx_all device i64 [n_shards, nq, k], score_all device f64 [n_shards, nq, k] -> global top-k,
 *   same (score desc, index asc) order, independent of the number of shards. */
int t2l_merge_topk(t2l_engine* e, const int64_t* idx_all, const double* score_all, int n_shards, int nq, int k,
                   int64_t* out_idx, double* out_score, void* stream);

/* The same merge on PACKED per-shard results: shard g contributes one contiguous block of 2 * nq * k 8-byte words,
 * [idx i64 [nq, k] | score f64 [nq, k]], i.e. what a single all-gather of each rank's (idx, score) pair delivers. */
int t2l_merge_topk_packed(t2l_engine* e, const void* packed_all, int n_shards, int nq, int k, int64_t* out_idx, double* out_score,
                          void* stream);

/* Accuracy bookkeeping of eval_epoch / run_coarse for a whole query batch (training/coarse.py:131-150: top-k hit and
 * close-by accuracy; evaluation/utils.py:31-54: calc_sample_accuracies).  For every query q and every k in top_k:
 *   hit[q, k]       = target_row[q] in idx[q, 0:k]
 *   within[q, k, t] = min(dists[q, 0:k]) <= threshs[t],  dists[q, j] = |query_xy[q] - cell_xy[idx[q, j]]| (float64,
 *                     evaluated as numpy does: sqrt(dx*dx + dy*dy)); +inf for empty slots and, when scene codes are
 *                     given, for rows of another scene
 *   idx device i64 [nq, k];  target_row device i64 [nq] or NULL;  query_xy device f64 [nq, 2];  cell_xy device f64 [N, 2]
 *   query_scene / cell_scene device i32 [nq] / [N], both or neither;  top_k HOST i32 [n_top] ascending (<= 8);
 *   threshs HOST f64 [n_thr] (<= 8);  hit device u8 [nq, n_top], within device u8 [nq, n_top, n_thr], dists device
 *   f64 [nq, k] -- any of the three may be NULL. */
int t2l_topk_accuracy(t2l_engine* e, const int64_t* idx, int nq, int k, const int64_t* target_row, const double* query_xy,
                      const double* cell_xy, const int32_t* query_scene, const int32_t* cell_scene,
                      const int32_t* top_k_host, int n_top, const double* threshs_host, int n_thr,
                      uint8_t* hit, uint8_t* within, double* dists, void* stream);

/* Number of kernels this engine has launched since creation (bench.py's gpu_launches). */
int64_t t2l_launch_count(const t2l_engine* e);

/* Test hook: C[M,N] = act(A[M,K] * W[N,K]^T + bias) through the tcgen05 tf32 GEMM (path=1) or the
 * fp32 SIMT GEMM (path=0).  act: 0 none, 1 relu.  segmax != 0: rows are max-reduced in groups of 32
 * after relu (C is [M/32, N]).  All pointers device. */
int t2l_debug_linear(t2l_engine* e, int path, const float* A, int lda, const float* W, int ldw, const float* bias,
                     float* C, int ldc, int M, int N, int K, int act, int segmax, void* stream);

/* Test hook for the fp16-operand tcgen05 GEMM the token layer runs on (kind::f16, fp32 accumulate):
 * A device f16 [M, lda], W device f16 [N, ldw] (as raw 16-bit words), bias device f32 [N] or NULL,
 * C device f32 [M, ldc] or, with out_half != 0, f16 [M, ldc] (saturating at +-65504). */
int t2l_debug_linear_f16(t2l_engine* e, const void* A, int lda, const void* W, int ldw, const float* bias,
                         void* C, int ldc, int M, int N, int K, int act, int out_half, void* stream);

/* Test hook for the fp16 residual stream of the token layer: C = fp16(A W^T + bias + R) with A f16 [M, lda], W f16 [N, ldw],
 * R f16 [M, ldr], C f16 [M, ldc] (acc + bias + R summed in fp32, one rounding, saturating).  reg_epilogue = 0: residual and
 * output move by TMA (the product path when M > 128 and N % 256 == 0); != 0: the register-staged epilogue. */
int t2l_debug_linear_f16_residual(t2l_engine* e, const void* A, int lda, const void* W, int ldw, const float* bias, const void* R,
                                  int ldr, void* C, int ldc, int M, int N, int K, int reg_epilogue, void* stream);

/* Test hook for the small-sequence attention core (intra-cell and sentence-level layers): qkv device f32 [n_seq*S, 3d] packed
 * (q | k | v), head h = columns [h*d/n_heads, ...), no mask; out device f32 [n_seq*S, d] = softmax(q k^T / sqrt(hd)) v. */
int t2l_debug_mha(t2l_engine* e, const float* qkv, float* out, int n_seq, int S, int d, int n_heads, void* stream);

/* Test hook for cross attention with the same cores (nn.TransformerDecoderLayer.multihead_attn, models/cross_matcher.py:113-115):
 * q device f32 [n_seq*Sq, d], kv device f32 [n_seq*Sk, 2d] packed (k | v), out device f32 [n_seq*Sq, d]; Sk <= 32. */
int t2l_debug_mha_cross(t2l_engine* e, const float* q, const float* kv, float* out, int n_seq, int Sq, int Sk, int d, int n_heads,
                        void* stream);

/* Test hook for the intra-cell attention core on packed rows without duplicate padding rows: cell b owns rows
 * row_ptr[b] .. row_ptr[b+1]) = its min(n_b, slots) objects + (if n_b < slots) one row standing for the slots - n_b
 * zero-padded slots of the reference's [B, slots, d] tensor; n_b = cell_ptr[b+1] - cell_ptr[b].  All pointers device. */
int t2l_debug_mha_cells(t2l_engine* e, const float* qkv, float* out, int n_cells, const int32_t* row_ptr_dev,
                        const int32_t* cell_ptr_dev, int slots, int d, int n_heads, void* stream);

/* Timing bisect of the fused set-abstraction kernel (profiling only; embeddings are INVALID while mode != 0):
 * 1 = the epilogue releases the accumulators without reading them, 2 = the MMA issuer skips the MMAs, 3 = both, 0 = normal. */
int t2l_debug_sa_bisect(t2l_engine* e, int mode);

#ifdef __cplusplus
}
#endif
#endif /* TEXT2LOC_B200_H */
