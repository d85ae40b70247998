// Query x database cosine-similarity top-k (training/coarse.py:119-125).
//
// The reference scores in float64 on the host, one GEMV + full argsort per query.  Here:
//   1. candidate pass on the tensor cores, ONE K=256 pass over fp16 copies of Q and D (rows scaled by exact powers of two
//      into fp16's normal range): |approx - q.d| <= kEps16 * ||q|| * max||d||, a rigorous worst-case bound (derivation at
//      kEps16).  The epilogue keeps, per query row and per database split, the 16 best approximate scores and the threshold
//      below which it dropped.  Exactness never depended on the precision of this pass -- steps 2-4 establish it -- so the
//      bf16 hi|lo three-pass product (hi*hi + hi*lo + lo*hi, |err| <= kEpsRel ...) that round 1 used for the candidates is
//      now the SECOND pass only; it is still the better first pass for tightly clustered databases (score gaps below
//      2 kEps16), so the engine falls back to it when most queries of the previous call failed the fp16 proof
//      (search_topk's `first_pass_bf16x3`; api.cu decides);
//   2. exact re-rank: fp64 dot products of the fp32 originals for those candidates, ordered by
//      (score desc, row asc);
//   3. proof: if every split's drop threshold + error bound is below the k-th exact score, no
//      dropped row can belong to the top-k and the query is done;
//   4. otherwise (tight clusters, e.g. near-collinear embeddings) a second tensor-core pass over just
//      those queries COLLECTS every row whose approximate score is >= (k-th exact score so far) - bound:
//      any true top-k member satisfies that, so an exact re-rank of the collected rows is the exact
//      answer.  Only if more than 256 rows qualify is the query rescanned exhaustively in fp64.
// Returned indices are therefore those of the fp64 stable-order oracle by construction.
#include "ops.h"
#include "umma_gemm.cuh"

namespace t2l {

constexpr int kCand = 16;         // width of the exact-scan / merge lists (k <= kCand)
constexpr int kList = 12;         // approximate candidates kept per register list; a query row has 2 lists per database split (24 candidates).
                                  // Measured at 32 768 x 100 000 / x 12 500: 16 -> 1.92 / 0.62 ms; 8 -> 2.30 / 0.54 ms (a fifth of the queries then fail
                                  // the proof: one half of the columns often holds 8 of the top ~12)
constexpr int kMaxSplits = 16;
constexpr int kMaxK = 12;
// bf16 hi|lo three-pass product: x = hi + lo + r with |r| <= 2^-16 |x| (two roundings at u = 2^-8).  The dropped terms
// lo*lo, r_q*d, q*r_d sum to <= 3 * 2^-16 |q_i||d_i| per element, <= 4.6e-5 ||q|| ||d|| per dot product (Cauchy-Schwarz);
// the 768 products are exact in fp32 (8 x 8 significand bits) and their fp32 accumulation, worst case with truncating
// adds, costs <= 768 * 2^-23 = 9.2e-5 of sum|terms|.  Total 1.37e-4, used with a 1.15x margin.
constexpr float kEpsRel = 1.6e-4f;
// fp16 single pass on rows scaled by exact powers of two so that ||q~||, max||d~|| lie in [0.5, 1): fl(x) = x (1 + delta) with
// |delta| <= u = 2^-11 in the normal range, absolute error <= 2^-25 below it.  Per element
//   |q~ d~ - fl(q~) fl(d~)| <= (2u + u^2) |q~||d~| + 2^-25 (|q~| + |d~|) + ...,
// summed over 256 elements: <= (2^-10 + 2^-22) ||q~|| ||d~|| + 2^-25 (||q~||_1 + ||d~||_1) <= 9.77e-4 ||q~|| ||d~|| + 1e-6
// (||x||_1 <= 16 ||x||_2); the 256 products are exact in fp32 (11 x 11 bits), truncating fp32 accumulation adds
// <= 256 * 2^-23 = 3.1e-5 of sum|terms| <= ||q~|| ||d~||.  With ||q~|| max||d~|| >= 1/4 the absolute 1e-6 is <= 4e-6 relative:
// total <= 1.012e-3 ||q|| max||d|| after undoing the (exact) scaling.
constexpr float kEps16 = 1.05e-3f;

// ---- fp32 rows -> bf16 hi | lo planes (+ row norms) ---------------------------------------------
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ x, long rows, __nv_bfloat16* __restrict__ planes,
                                                           float* __restrict__ norms, float* __restrict__ max_norm) {
  const long r = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + r * kEmbed;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kEmbed / 32; ++i) {
    const int c = i * 32 + lane;
    const float v = xr[c];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    planes[r * 2 * kEmbed + c] = hi;
    planes[r * 2 * kEmbed + kEmbed + c] = lo;
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) {
    const float n = sqrtf(ss);
    if (norms) norms[r] = n;
    if (max_norm) atomicMax(reinterpret_cast<unsigned int*>(max_norm), __float_as_uint(n));  // n >= 0
  }
}

// ---- fp32 rows -> fp16 rows scaled by a power of two ------------------------------------------------
// scale = 2^-ex with ex from frexp(norm): norm * scale lies in [0.5, 1) (1 for an all-zero operand), so every element is
// <= 1 in magnitude and the scaling itself is exact.  kPerRow: each row by its own norm (queries); else all rows by
// *ref_norm (the database's largest row norm).  scale_out receives 2^ex, the factor that undoes the scaling.
template <bool kPerRow>
__global__ void __launch_bounds__(256) scaled_half_rows_kernel(const float* __restrict__ x, long rows, const float* __restrict__ ref_norm,
                                                               __half* __restrict__ out, float* __restrict__ scale_out) {
  const long r = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + r * kEmbed);
  const float4 a = xr[2 * lane], b = xr[2 * lane + 1];  // 8 consecutive elements per lane
  float norm;
  if (kPerRow) {
    float ss = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
    norm = sqrtf(warp_sum(ss)) * 1.000001f;  // the fp32 sum may round down: never let the scaled norm exceed 1
  } else {
    norm = ref_norm[0] * 1.000001f;
  }
  int ex = 0;
  if (norm > 0.f && norm < INFINITY) frexpf(norm, &ex);
  const float sc = ldexpf(1.f, -ex);
  if (lane == 0 && (kPerRow || r == 0)) scale_out[kPerRow ? r : 0] = ldexpf(1.f, ex);
  const __half2 h0 = __floats2half2_rn(a.x * sc, a.y * sc), h1 = __floats2half2_rn(a.z * sc, a.w * sc),
                h2 = __floats2half2_rn(b.x * sc, b.y * sc), h3 = __floats2half2_rn(b.z * sc, b.w * sc);
  uint4 u;
  u.x = *reinterpret_cast<const uint32_t*>(&h0); u.y = *reinterpret_cast<const uint32_t*>(&h1);
  u.z = *reinterpret_cast<const uint32_t*>(&h2); u.w = *reinterpret_cast<const uint32_t*>(&h3);
  reinterpret_cast<uint4*>(out + r * kEmbed)[lane] = u;
}

cudaError_t search_prepare_db(const SearchDb& db, cudaStream_t st, Launches* lc) {
  cudaError_t e = cudaMemsetAsync(db.max_norm, 0, sizeof(float), st);
  if (e != cudaSuccess || db.n_rows <= 0) return e;
  if (lc) lc->n += 2;
  const unsigned grid = static_cast<unsigned>((db.n_rows + 7) / 8);
  split_planes_kernel<<<grid, 256, 0, st>>>(db.D, db.n_rows, db.planes, nullptr, db.max_norm);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  scaled_half_rows_kernel<false><<<grid, 256, 0, st>>>(db.D, db.n_rows, db.max_norm, db.plane16, db.scale);
  return cudaGetLastError();
}

// ---- candidate epilogue -------------------------------------------------------------------------
// Each epilogue thread owns one query row (and one of the two column sets, below) and keeps its kList = 12 best approximate scores
// of the current split in REGISTERS as a descending sorted list; inserting is a branch-free 12-step compare-exchange chain
// taken only when a score beats the list minimum (the drop threshold).  Everything the thread ever dropped is therefore <=
// the final minimum, which is what the re-rank proof needs.
//
// With a single K = 256 pass per tile the epilogue has only as many cycles as the MMAs take (2 048 per 128 x 256 tile, and
// draining the accumulator from tensor memory at 64 B/clk already uses all of them), and one warp per SM sub-partition --
// latency-bound on its dependent compare chains -- needed ~3x that (measured: 2.9 ms of epilogue behind 1.0 ms of MMAs at
// 32 768 x 100 000).  Hence: (1) TWO epilogue warp sets (warps 4..7 take the even 32-column chunks of a tile, warps 8..11 the
// odd ones; every row then has two lists per split, 2 x 12 = 24 candidates; a shorter list means fewer insertions while it warms
// up, ~L (1 + ln(n / L)), and a shorter chain per insertion); (2) a 3-input max tree (~16 instructions) decides the
// common "nothing in this chunk beats the minimum" case before any per-element work; (3) two inlined copies of chunk() instead
// of eight (instruction cache).  Measured: 1.92 ms for the whole search at 32 768 x 100 000 (3.5 ms with the three-pass
// candidates of round 1).  A variant that parked hits in shared memory and inserted them once per tile for all 32 rows
// together was slower (2.7 ms): its per-chunk parking code outweighed the chains it saved.
struct TopKEpi {
  struct Params {
    float* cand_score;  // [nq, n_lists, kList]     n_lists = n_splits * kSets: one list per (database split, column set)
    int32_t* cand_idx;  // [nq, n_lists, kList]     -1 = empty
    float* cand_thr;    // [nq, n_lists]            -inf = nothing was dropped
    int nq, n_db, n_splits;
  };
  static constexpr int kSets = 2;                          // two epilogue warps per row: even / odd 32-column chunks
  static constexpr int kSmemBytes = kSets * 32 * 128 * 4;  // one 32-float column per epilogue thread (candidate staging)
  static constexpr bool kCompactLoop = true;
  const Params& p;
  float* s_v;
  int t;  // 0..127: row inside this CTA's 128-row tile
  int set;
  float ls[kList];
  int li[kList];
  bool active;

  __device__ TopKEpi(const Params& p_, uint8_t* smem, int ew, int lane, int, int set_)
      : p(p_), s_v(reinterpret_cast<float*>(smem) + set_ * 32 * 128), t(ew * 32 + lane), set(set_), active(false) {}
  __device__ void prefetch_unit(int, int) {}
  __device__ void begin_unit(int m_tile, int) {
    active = (m_tile * 128 + t) < p.nq;
#pragma unroll
    for (int i = 0; i < kList; ++i) { ls[i] = active ? -INFINITY : INFINITY; li[i] = -1; }
  }
  __device__ void begin_tile(int, int, int) {}
  __device__ __forceinline__ void insert(float x, int xi) {
#pragma unroll
    for (int i = 0; i < kList; ++i) {  // x sinks through the descending list; the old minimum falls out
      const bool gt = x > ls[i];
      const float s_keep = gt ? x : ls[i], s_next = gt ? ls[i] : x;
      const int i_keep = gt ? xi : li[i], i_next = gt ? li[i] : xi;
      ls[i] = s_keep; li[i] = i_keep; x = s_next; xi = i_next;
    }
  }
  __device__ void chunk(int, int, int, int col0, float (&v)[32]) {
    if (col0 + 32 > p.n_db) {  // database tail: zero-filled rows must not compete
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = (col0 + i < p.n_db) ? v[i] : -INFINITY;
    }
    const float thr = ls[kList - 1];
    float m8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m8[i] = fmaxf(fmaxf(v[4 * i], v[4 * i + 1]), fmaxf(v[4 * i + 2], v[4 * i + 3]));
    const float vmax = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
    if (!(vmax > thr)) return;  // the common case once the list has warmed up
    // rare path: park the 32 scores in shared memory so ONE copy of the insertion chain can walk the
    // set bits with a dynamic index (32 unrolled copies thrashed the instruction cache: 3x slower)
    uint32_t mask = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) mask |= (v[i] > thr) ? (1u << i) : 0u;
#pragma unroll
    for (int i = 0; i < 32; ++i) s_v[i * 128 + t] = v[i];
    while (mask) {
      const int i = __ffs(mask) - 1;
      mask &= mask - 1;
      const float x = s_v[i * 128 + t];
      if (x > ls[kList - 1]) insert(x, col0 + i);
    }
  }
  __device__ void end_unit(int m_tile, int split) {
    if (!active) return;
    const long row = static_cast<long>(m_tile) * 128 + t;
    const long list = row * (p.n_splits * kSets) + split * kSets + set;
#pragma unroll
    for (int i = 0; i < kList; ++i) {
      p.cand_score[list * kList + i] = ls[i];
      p.cand_idx[list * kList + i] = li[i];
    }
    p.cand_thr[list] = (li[kList - 1] >= 0) ? ls[kList - 1] : -INFINITY;  // list not full: nothing was dropped
  }
};

// ---- canonical fp64 dot: the ONE definition of an exact score -------------------------------------
// lane l accumulates elements l, l+32, ..., l+224 in that order, then a fixed xor butterfly.
// Products of two fp32 values are exact in fp64, so only the 8+5 additions round.
__device__ __forceinline__ double warp_dot256(const float* __restrict__ a, const float* __restrict__ b, int lane) {
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < kEmbed / 32; ++i) acc = fma(static_cast<double>(a[i * 32 + lane]), static_cast<double>(b[i * 32 + lane]), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

__device__ __forceinline__ bool better(double sa, long ia, double sb, long ib) {  // (score desc, index asc); idx < 0 = empty
  if (ia < 0) return false;
  if (ib < 0) return true;
  return sa > sb || (sa == sb && ia < ib);
}

// Select the top-k of `n` candidates held in per-warp smem arrays; writes them in order.
// Entries are consumed (their index is set to -1).
__device__ void warp_select_topk(double* s_sc, long* s_ix, int n, int k, int lane, double* out_sc, long* out_ix) {
  for (int r = 0; r < k; ++r) {
    double bs = 0.0;
    long bi = -1;
    int bp = -1;
    for (int c = lane; c < n; c += 32)
      if (better(s_sc[c], s_ix[c], bs, bi)) { bs = s_sc[c]; bi = s_ix[c]; bp = c; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double os = __shfl_xor_sync(0xffffffffu, bs, o);
      const long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int op = __shfl_xor_sync(0xffffffffu, bp, o);
      if (better(os, oi, bs, bi)) { bs = os; bi = oi; bp = op; }
    }
    if (lane == 0) {
      out_sc[r] = (bi >= 0) ? bs : -INFINITY;
      out_ix[r] = bi;
      if (bp >= 0) s_ix[bp] = -1;
    }
    __syncwarp();
  }
}

// The same selection for n <= 32 candidates with one candidate per lane: a lane's output position is the number of candidates
// that beat it (the order (score desc, row asc) is total: a row appears once).  n broadcasts instead of k rounds of a
// five-step (score, row, position) butterfly.
__device__ void warp_rank_select(const double* s_sc, const long* s_ix, int n, int k, int lane, double* out_sc, long* out_ix) {
  const double ms = lane < n ? s_sc[lane] : 0.0;
  const long mi = lane < n ? s_ix[lane] : -1;
  int rank = 0;
  for (int j = 0; j < n; ++j) {
    const double os = __shfl_sync(0xffffffffu, ms, j);
    const long oi = __shfl_sync(0xffffffffu, mi, j);
    if (better(os, oi, ms, mi)) ++rank;
  }
  const int n_valid = __popc(__ballot_sync(0xffffffffu, mi >= 0));
  if (lane < k && lane >= n_valid) { out_sc[lane] = -INFINITY; out_ix[lane] = -1; }
  if (mi >= 0 && rank < k) { out_sc[rank] = ms; out_ix[rank] = mi; }
  __syncwarp();
}

// ---- exact re-rank + proof (one warp per query) ----------------------------------------------------
constexpr int kRerankWarps = 4;

__global__ void __launch_bounds__(kRerankWarps * 32) rerank_kernel(const float* __restrict__ Q, const float* __restrict__ D, const float* __restrict__ cand_score,
                                                                   const int32_t* __restrict__ cand_idx, const float* __restrict__ cand_thr,
                                                                   const float* __restrict__ q_norm, const float* __restrict__ max_norm,
                                                                   const float* __restrict__ q_scale, const float* __restrict__ db_scale, float eps_cand,
                                                                   int nq, int n_splits,
                                                                   int k, long row_offset, int64_t* __restrict__ out_idx, double* __restrict__ out_score,
                                                                   int32_t* __restrict__ flags, int32_t* __restrict__ n_fail, int32_t* __restrict__ fail_ids,
                                                                   float* __restrict__ fail_thr) {
  __shared__ double s_sc[kRerankWarps][kMaxSplits * kCand];
  __shared__ long s_ix[kRerankWarps][kMaxSplits * kCand];
  __shared__ double s_osc[kRerankWarps][kMaxK];
  __shared__ long s_oix[kRerankWarps][kMaxK];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * kRerankWarps + w;
  if (q >= nq) return;
  int C = n_splits * kList;
  const float* qr = Q + static_cast<long>(q) * kEmbed;
  const float qn = q_norm[q] * 1.000001f, dn = max_norm[0] * 1.000001f;  // fp32 norms may be rounded down
  const double eps = static_cast<double>(eps_cand) * qn * dn;            // error bound of the candidate pass that ran
  const double unscale = q_scale ? static_cast<double>(q_scale[q]) * static_cast<double>(db_scale[0]) : 1.0;  // exact powers of two
  const int32_t* ci = cand_idx + static_cast<long>(q) * C;
  // Pre-filter (<= 32 candidates, one per lane): with t the k-th largest APPROXIMATE score, k candidates have an exact score
  // >= t - eps, so a candidate whose approximate score is below t - 2 eps has an exact score strictly below all of them and
  // cannot enter (or tie into) the top-k.  About half of the 24 candidates of a query go: half the row gathers and fp64 dots.
  __shared__ int32_t s_keep[kRerankWarps][32];
  if (C <= 32) {
    const int my_idx = lane < C ? ci[lane] : -1;
    const double my_a = lane < C && my_idx >= 0 ? static_cast<double>(cand_score[static_cast<long>(q) * C + lane]) * unscale : 0.0;
    int rank = 0;
    for (int j = 0; j < C; ++j) {
      const double oa = __shfl_sync(0xffffffffu, my_a, j);
      const int oi = __shfl_sync(0xffffffffu, my_idx, j);
      if (oi >= 0 && (oa > my_a || (oa == my_a && j < lane))) ++rank;
    }
    const unsigned valid = __ballot_sync(0xffffffffu, my_idx >= 0);
    bool keep = my_idx >= 0;
    if (__popc(valid) > k) {
      const unsigned at_k = __ballot_sync(0xffffffffu, my_idx >= 0 && rank == k - 1);  // exactly one lane: the ranks of valid lanes are distinct
      const double t = __shfl_sync(0xffffffffu, my_a, __ffs(at_k) - 1);
      keep = keep && !(my_a < t - 2.000001 * eps);
    }
    const unsigned kept = __ballot_sync(0xffffffffu, keep);
    if (keep) s_keep[w][__popc(kept & ((1u << lane) - 1u))] = my_idx;
    __syncwarp();
    C = __popc(kept);
    ci = s_keep[w];
  }
  // The query's elements stay in registers, and four candidates are in flight at a time (their row reads and fp64 chains are
  // independent): one candidate after the other, each dot waited for its own loads and eight dependent fp64 FMAs -- 0.21 ms per
  // 32 768 queries x 24 candidates, as much as the candidate GEMM itself at 12 500 rows.  Same operations in the same
  // order as warp_dot256, so the scores are bit-identical to every other exact score of the engine.
  double qd[kEmbed / 32];
#pragma unroll
  for (int i = 0; i < kEmbed / 32; ++i) qd[i] = static_cast<double>(qr[i * 32 + lane]);
  for (int c0 = 0; c0 < C; c0 += 4) {
    int idx[4];
    double acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      idx[u] = (c0 + u < C) ? ci[c0 + u] : -1;  // warp-uniform
      acc[u] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < kEmbed / 32; ++i) {
      float dv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) dv[u] = idx[u] >= 0 ? D[static_cast<long>(idx[u]) * kEmbed + i * 32 + lane] : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fma(qd[i], static_cast<double>(dv[u]), acc[u]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c0 + u < C) { s_sc[w][c0 + u] = idx[u] >= 0 ? acc[u] : 0.0; s_ix[w][c0 + u] = idx[u]; }
    }
  }
  __syncwarp();
  if (C <= 32) warp_rank_select(s_sc[w], s_ix[w], C, k, lane, s_osc[w], s_oix[w]);
  else warp_select_topk(s_sc[w], s_ix[w], C, k, lane, s_osc[w], s_oix[w]);
  // proof: every dropped row has approx <= thr_s, hence exact <= thr_s + eps; it cannot enter
  // (or tie into) the top-k if thr_s + eps < k-th exact score.
  const double kth = s_osc[w][k - 1];
  bool fail = false;
  for (int s = lane; s < n_splits; s += 32) {
    const float thr = cand_thr[static_cast<long>(q) * n_splits + s];
    if (thr != -INFINITY && !(static_cast<double>(thr) * unscale + eps < kth)) fail = true;
  }
  fail = __any_sync(0xffffffffu, fail);
  if (lane < k) {
    const long ix = s_oix[w][lane];
    out_idx[static_cast<long>(q) * k + lane] = ix >= 0 ? ix + row_offset : -1;
    out_score[static_cast<long>(q) * k + lane] = s_osc[w][lane];
  }
  if (lane == 0) {
    flags[q] = 0;
    if (fail) {  // queue for the second pass: rows with approx >= kth - eps are the only possible top-k members
      const int slot = atomicAdd(n_fail, 1);
      fail_ids[slot] = q;
      // the second pass is always the bf16 hi|lo product on unscaled rows: its own bound applies
      fail_thr[slot] = __double2float_rd(kth - static_cast<double>(kEpsRel) * qn * dn);
    }
  }
}

// ---- second pass: compact the failed queries' operand planes -----------------------------------------
__global__ void __launch_bounds__(256) gather_fail_planes_kernel(const __nv_bfloat16* __restrict__ q_planes, const int32_t* __restrict__ fail_ids,
                                                                 const int32_t* __restrict__ n_fail, __nv_bfloat16* __restrict__ q2_planes,
                                                                 int32_t* __restrict__ cand_cnt) {
  const int slot = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (slot >= *n_fail) return;
  const int lane = threadIdx.x & 31;
  if (lane == 0) cand_cnt[slot] = 0;
  const uint4* src = reinterpret_cast<const uint4*>(q_planes + static_cast<long>(fail_ids[slot]) * 2 * kEmbed);
  uint4* dst = reinterpret_cast<uint4*>(q2_planes + static_cast<long>(slot) * 2 * kEmbed);
  dst[lane] = src[lane];
  dst[lane + 32] = src[lane + 32];
}

// Epilogue of the second pass: no ranking, just the row indices whose approximate score reaches the
// query's threshold.  One kPass2Cap-entry buffer per query shared by all database splits (slots are
// claimed with an atomic counter; the count may exceed the capacity = overflow).  The order in which
// rows land in the buffer is irrelevant: the re-rank orders them by (exact score, row).
struct CollectEpi {
  struct Params {
    const float* thr;       // [slots]
    const int32_t* n_rows;  // [1] live slots
    int32_t* cand_idx;      // [slots, kPass2Cap]
    int32_t* cand_cnt;      // [slots]  zeroed by gather_fail_planes_kernel
    int n_db;
  };
  static constexpr int kSmemBytes = 0;
  static constexpr bool kCompactLoop = false;
  static constexpr int kSets = 1;
  const Params& p;
  int t;
  float T;
  long row;
  __device__ CollectEpi(const Params& p_, uint8_t*, int ew, int lane, int, int) : p(p_), t(ew * 32 + lane), T(INFINITY), row(0) {}
  __device__ void prefetch_unit(int, int) {}
  __device__ void begin_unit(int m_tile, int) {
    row = static_cast<long>(m_tile) * 128 + t;
    T = (row < *p.n_rows) ? p.thr[row] : INFINITY;
  }
  __device__ void begin_tile(int, int, int) {}
  __device__ void chunk(int, int, int, int col0, float (&v)[32]) {
    uint32_t mask = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) mask |= (v[i] >= T && col0 + i < p.n_db) ? (1u << i) : 0u;
    while (mask) {
      const int i = __ffs(mask) - 1;
      mask &= mask - 1;
      const int pos = atomicAdd(p.cand_cnt + row, 1);
      if (pos < kPass2Cap) p.cand_idx[row * kPass2Cap + pos] = col0 + i;
    }
  }
  __device__ void end_unit(int, int) {}
};

// Exact re-rank of the collected rows (one warp per failed query); overflow -> flag for the exhaustive scan.
__global__ void __launch_bounds__(kRerankWarps * 32) rerank2_kernel(const float* __restrict__ Q, const float* __restrict__ D, const int32_t* __restrict__ fail_ids,
                                                                    const int32_t* __restrict__ n_fail, const int32_t* __restrict__ cand_idx,
                                                                    const int32_t* __restrict__ cand_cnt, int k, long row_offset,
                                                                    int64_t* __restrict__ out_idx, double* __restrict__ out_score, int32_t* __restrict__ flags) {
  __shared__ double s_sc[kRerankWarps][kPass2Cap];
  __shared__ long s_ix[kRerankWarps][kPass2Cap];
  __shared__ double s_osc[kRerankWarps][kMaxK];
  __shared__ long s_oix[kRerankWarps][kMaxK];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = blockIdx.x * kRerankWarps + w;
  if (slot >= *n_fail) return;
  const int q = fail_ids[slot];
  const int n = cand_cnt[slot];  // warp-uniform
  if (n > kPass2Cap) {
    if (lane == 0) flags[q] = 2;
    return;
  }
  for (int i = lane; i < n; i += 32) s_ix[w][i] = cand_idx[static_cast<long>(slot) * kPass2Cap + i];
  __syncwarp();
  const float* qr = Q + static_cast<long>(q) * kEmbed;
  for (int c = 0; c < n; ++c) {
    const double s = warp_dot256(qr, D + s_ix[w][c] * kEmbed, lane);
    if (lane == 0) s_sc[w][c] = s;
  }
  __syncwarp();
  warp_select_topk(s_sc[w], s_ix[w], n, k, lane, s_osc[w], s_oix[w]);
  if (lane < k) {
    const long ix = s_oix[w][lane];
    out_idx[static_cast<long>(q) * k + lane] = ix >= 0 ? ix + row_offset : -1;
    out_score[static_cast<long>(q) * k + lane] = s_osc[w][lane];
  }
}

// ---- exact fp64 scan (one CTA per query) -------------------------------------------------------------
constexpr int kScanWarps = 8;

__global__ void __launch_bounds__(kScanWarps * 32) exact_scan_kernel(const float* __restrict__ Q, const float* __restrict__ D, long n_rows, int nq, int k,
                                                                     long row_offset, const int32_t* __restrict__ only_flagged,
                                                                     int64_t* __restrict__ out_idx, double* __restrict__ out_score) {
  __shared__ double s_sc[kScanWarps * kCand];
  __shared__ long s_ix[kScanWarps * kCand];
  __shared__ double s_osc[kCand];
  __shared__ long s_oix[kCand];
  const int q = blockIdx.x;
  if (q >= nq || (only_flagged && !only_flagged[q])) return;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* my_sc = s_sc + w * kCand;
  long* my_ix = s_ix + w * kCand;
  if (lane < kCand) { my_sc[lane] = -INFINITY; my_ix[lane] = -1; }
  __syncwarp();
  const float* qr = Q + static_cast<long>(q) * kEmbed;
  int cnt = 0;
  for (long r = w; r < n_rows; r += kScanWarps) {
    const double s = warp_dot256(qr, D + r * kEmbed, lane);
    // per-warp list sorted by (score desc, row asc); rows arrive in ascending order, so a row
    // that ties an entry ranks after it
    if (cnt < k || s > my_sc[k - 1]) {
      if (lane == 0) {
        int pos = cnt < k ? cnt : k - 1;
        while (pos > 0 && my_sc[pos - 1] < s) { my_sc[pos] = my_sc[pos - 1]; my_ix[pos] = my_ix[pos - 1]; --pos; }
        my_sc[pos] = s;
        my_ix[pos] = r;
      }
      if (cnt < k) ++cnt;
      __syncwarp();
    }
  }
  __syncthreads();
  if (w == 0) {
    warp_select_topk(s_sc, s_ix, kScanWarps * kCand, k, lane, s_osc, s_oix);
    if (lane < k) {
      const long ix = s_oix[lane];
      out_idx[static_cast<long>(q) * k + lane] = ix >= 0 ? ix + row_offset : -1;
      out_score[static_cast<long>(q) * k + lane] = s_osc[lane];
    }
  }
}

cudaError_t search_topk_exact(const SearchDb& db, const float* Q, int nq, int k, int64_t* out_idx, double* out_score,
                              const int32_t* only_flagged, cudaStream_t st, Launches* lc) {
  if (nq <= 0) return cudaSuccess;
  if (k < 1 || k > kCand) return cudaErrorInvalidValue;
  if (lc) lc->n++;
  exact_scan_kernel<<<nq, kScanWarps * 32, 0, st>>>(Q, db.D, db.n_rows, nq, k, db.row_offset, only_flagged, out_idx, out_score);
  return cudaGetLastError();
}

// ---- the search ----------------------------------------------------------------------------------------
cudaError_t search_topk(const SearchDb& db, const SearchWork& w, const float* Q, int nq, int k, int64_t* out_idx,
                        double* out_score, int32_t* out_n_fallback, bool first_pass_bf16x3, cudaStream_t st, Launches* lc) {
  if (nq <= 0) return cudaSuccess;
  if (k < 1 || k > kMaxK || nq > w.nq_cap) return cudaErrorInvalidValue;
  cudaError_t e;
  if ((e = cudaMemsetAsync(w.n_fail, 0, sizeof(int32_t), st)) != cudaSuccess) return e;
  if (out_n_fallback && (e = cudaMemsetAsync(out_n_fallback, 0, sizeof(int32_t), st)) != cudaSuccess) return e;
  if (db.n_rows <= 0) return search_topk_exact(db, Q, nq, k, out_idx, out_score, nullptr, st, lc);

  using Cfg = GemmCfg<256, kOpBf16, 2>;   // CTA pairs: 256 queries x 256 database rows per tile
  using Cfg16 = GemmCfg<256, kOpF16, 2>;  // same tile on the fp16 rows
  constexpr int kTileM = Cfg::BLOCK_M * Cfg::CTA_GROUP;
  if (lc) lc->n += 4;
  split_planes_kernel<<<(nq + 7) / 8, 256, 0, st>>>(Q, nq, w.q_planes, w.q_norm, nullptr);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (!first_pass_bf16x3) {
    scaled_half_rows_kernel<true><<<(nq + 7) / 8, 256, 0, st>>>(Q, nq, nullptr, w.q16, w.q_scale);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }

  CUtensorMap ta, tb;
  if (make_operand_map(&ta, w.q_planes, kOpBf16, nq, 2 * kEmbed, 2 * kEmbed, Cfg::BLOCK_M)) return cudaErrorInvalidValue;
  if (make_operand_map(&tb, db.planes, kOpBf16, db.n_rows, 2 * kEmbed, 2 * kEmbed, Cfg::LOAD_N)) return cudaErrorInvalidValue;
  GemmShape s;
  s.M = nq;
  s.N = static_cast<int>(db.n_rows);
  s.m_tiles = (nq + kTileM - 1) / kTileM;
  s.n_tiles = (s.N + Cfg::BLOCK_N - 1) / Cfg::BLOCK_N;
  // database splits: only as many as it takes to give every CTA pair a unit; each split costs every
  // query another ~16 (1 + ln(rows/16)) list insertions and another 16 candidates to re-rank
  int want = (tma_api().num_sms / Cfg::CTA_GROUP) / s.m_tiles;
  int n_splits = want < 1 ? 1 : want;
  if (n_splits > s.n_tiles) n_splits = s.n_tiles;
  if (n_splits > kMaxSplits / TopKEpi::kSets) n_splits = kMaxSplits / TopKEpi::kSets;  // the re-rank handles kMaxSplits lists per query
  if (n_splits > w.splits_cap / TopKEpi::kSets) n_splits = w.splits_cap / TopKEpi::kSets;
  s.tiles_per_split = (s.n_tiles + n_splits - 1) / n_splits;
  n_splits = (s.n_tiles + s.tiles_per_split - 1) / s.tiles_per_split;
  s.n_splits = n_splits;
  s.ks.n_pass = 3;
  s.ks.kb_per_pass = kEmbed / Cfg::BLOCK_K;
  s.ks.a_off[0] = 0;      s.ks.b_off[0] = 0;       // hi * hi
  s.ks.a_off[1] = 0;      s.ks.b_off[1] = kEmbed;  // hi * lo
  s.ks.a_off[2] = kEmbed; s.ks.b_off[2] = 0;       // lo * hi
  TopKEpi::Params ep{w.cand_score, w.cand_idx, w.cand_thr, nq, s.N, n_splits};
  if (first_pass_bf16x3) {
    if ((e = launch_umma_gemm<Cfg, TopKEpi>(ta, tb, s, ep, st)) != cudaSuccess) return e;
  } else {
    CUtensorMap ta16, tb16;
    if (make_operand_map(&ta16, w.q16, kOpF16, nq, kEmbed, kEmbed, Cfg16::BLOCK_M)) return cudaErrorInvalidValue;
    if (make_operand_map(&tb16, db.plane16, kOpF16, db.n_rows, kEmbed, kEmbed, Cfg16::LOAD_N)) return cudaErrorInvalidValue;
    GemmShape s16 = s;
    s16.ks.n_pass = 1;
    s16.ks.kb_per_pass = kEmbed / Cfg16::BLOCK_K;
    s16.ks.a_off[0] = 0; s16.ks.b_off[0] = 0;
    if ((e = launch_umma_gemm<Cfg16, TopKEpi>(ta16, tb16, s16, ep, st)) != cudaSuccess) return e;
  }

  rerank_kernel<<<(nq + kRerankWarps - 1) / kRerankWarps, kRerankWarps * 32, 0, st>>>(
      Q, db.D, w.cand_score, w.cand_idx, w.cand_thr, w.q_norm, db.max_norm, first_pass_bf16x3 ? nullptr : w.q_scale, db.scale,
      first_pass_bf16x3 ? kEpsRel : kEps16, nq, n_splits * TopKEpi::kSets, k, db.row_offset, out_idx, out_score, w.flags, w.n_fail, w.fail_ids,
      w.fail_thr);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;

  // ---- second pass over the queries whose proof failed (count lives on the device: no host sync; the GEMM
  // skips tiles beyond it).  Same operands, same splits; the epilogue collects instead of ranking.
  if (lc) lc->n += 3;
  gather_fail_planes_kernel<<<(nq + 7) / 8, 256, 0, st>>>(w.q_planes, w.fail_ids, w.n_fail, w.q2_planes, w.cand2_cnt);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  CUtensorMap ta2;
  if (make_operand_map(&ta2, w.q2_planes, kOpBf16, nq, 2 * kEmbed, 2 * kEmbed, Cfg::BLOCK_M)) return cudaErrorInvalidValue;
  // Few queries fail, so the second pass has few query tiles: it is split along the DATABASE instead (the collecting
  // epilogue keeps no per-split state), enough units for every CTA pair even when a single 256-query tile is live.
  GemmShape s2 = s;
  s2.m_rows_dev = w.n_fail;
  {
    int splits2 = tma_api().num_sms / Cfg::CTA_GROUP;
    if (splits2 > s.n_tiles) splits2 = s.n_tiles;
    s2.tiles_per_split = (s.n_tiles + splits2 - 1) / splits2;
    s2.n_splits = (s.n_tiles + s2.tiles_per_split - 1) / s2.tiles_per_split;
  }
  CollectEpi::Params ep2{w.fail_thr, w.n_fail, w.cand2_idx, w.cand2_cnt, s.N};
  if ((e = launch_umma_gemm<Cfg, CollectEpi>(ta2, tb, s2, ep2, st)) != cudaSuccess) return e;
  rerank2_kernel<<<(nq + kRerankWarps - 1) / kRerankWarps, kRerankWarps * 32, 0, st>>>(
      Q, db.D, w.fail_ids, w.n_fail, w.cand2_idx, w.cand2_cnt, k, db.row_offset, out_idx, out_score, w.flags);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (out_n_fallback && (e = cudaMemcpyAsync(out_n_fallback, w.n_fail, sizeof(int32_t), cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
  // exhaustive fp64 rescan only where the second pass overflowed its candidate buffer
  return search_topk_exact(db, Q, nq, k, out_idx, out_score, w.flags, st, lc);
}

// ---- merge of per-shard lists (one warp per query) -----------------------------------------------------
__global__ void __launch_bounds__(128) merge_kernel(const int64_t* __restrict__ idx_all, const double* __restrict__ score_all, long shard_stride,
                                                    int n_shards, int nq, int k, int64_t* __restrict__ out_idx, double* __restrict__ out_score) {
  __shared__ double s_sc[4][kMaxSplits * kCand];
  __shared__ long s_ix[4][kMaxSplits * kCand];
  __shared__ double s_osc[4][kCand];
  __shared__ long s_oix[4][kCand];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + w;
  if (q >= nq) return;
  const int C = n_shards * k;
  for (int c = lane; c < C; c += 32) {
    const int g = c / k, j = c % k;
    const long src = g * shard_stride + static_cast<long>(q) * k + j;
    s_sc[w][c] = score_all[src];
    s_ix[w][c] = idx_all[src];
  }
  __syncwarp();
  warp_select_topk(s_sc[w], s_ix[w], C, k, lane, s_osc[w], s_oix[w]);
  if (lane < k) {
    out_idx[static_cast<long>(q) * k + lane] = s_oix[w][lane];
    out_score[static_cast<long>(q) * k + lane] = s_osc[w][lane];
  }
}

// ---- running top-k of a streamed database (one warp per query): merge a chunk's list into the running list, in place ----
__global__ void __launch_bounds__(128) merge_running_kernel(int64_t* __restrict__ run_idx, double* __restrict__ run_score, const int64_t* __restrict__ new_idx,
                                                            const double* __restrict__ new_score, int nq, int k) {
  __shared__ double s_sc[4][2 * kCand];
  __shared__ long s_ix[4][2 * kCand];
  __shared__ double s_osc[4][kCand];
  __shared__ long s_oix[4][kCand];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + w;
  if (q >= nq) return;
  for (int c = lane; c < 2 * k; c += 32) {
    const bool from_new = c >= k;
    const long src = static_cast<long>(q) * k + (from_new ? c - k : c);
    s_sc[w][c] = from_new ? new_score[src] : run_score[src];
    s_ix[w][c] = from_new ? new_idx[src] : run_idx[src];
  }
  __syncwarp();
  warp_select_topk(s_sc[w], s_ix[w], 2 * k, k, lane, s_osc[w], s_oix[w]);
  if (lane < k) {
    run_idx[static_cast<long>(q) * k + lane] = s_oix[w][lane];
    run_score[static_cast<long>(q) * k + lane] = s_osc[w][lane];
  }
}

cudaError_t merge_running_topk(int64_t* run_idx, double* run_score, const int64_t* new_idx, const double* new_score, int nq, int k,
                               cudaStream_t st, Launches* lc) {
  if (nq <= 0) return cudaSuccess;
  if (k < 1 || k > kCand) return cudaErrorInvalidValue;
  if (lc) lc->n++;
  merge_running_kernel<<<(nq + 3) / 4, 128, 0, st>>>(run_idx, run_score, new_idx, new_score, nq, k);
  return cudaGetLastError();
}

cudaError_t merge_topk(const int64_t* idx_all, const double* score_all, long shard_stride, int n_shards, int nq, int k, int64_t* out_idx,
                       double* out_score, cudaStream_t st, Launches* lc) {
  if (nq <= 0) return cudaSuccess;
  if (k < 1 || k > kCand || n_shards < 1 || n_shards * k > kMaxSplits * kCand) return cudaErrorInvalidValue;
  if (lc) lc->n++;
  merge_kernel<<<(nq + 3) / 4, 128, 0, st>>>(idx_all, score_all, shard_stride, n_shards, nq, k, out_idx, out_score);
  return cudaGetLastError();
}

}  // namespace t2l
