// Row-wise fp32 kernels around the GEMMs: F.normalize, LayerNorm, the small-sequence unmasked
// multi-head attention core, max pooling over a sequence, object scatter into the padded cell
// tensor.  All are HBM/L2-bound streaming kernels: one warp per row, 128-bit accesses where the
// row length allows.
#include <cuda_fp16.h>

#include "ops.h"
#include "common.cuh"

namespace t2l {

// ---- F.normalize(x, dim=-1): x / max(||x||_2, 1e-12) ------------------------------------------
__global__ void l2_normalize_kernel(const float* __restrict__ x, long ldx, float* __restrict__ y, long ldy, int rows, int d) {
  const long r = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + r * ldx;
  float ss = 0.f;
  for (int c = lane; c < d; c += 32) ss = fmaf(xr[c], xr[c], ss);
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  float* yr = y + r * ldy;
  for (int c = lane; c < d; c += 32) yr[c] = xr[c] * inv;
}

cudaError_t l2_normalize_rows(const float* x, long ldx, float* y, long ldy, int rows, int d, cudaStream_t st, Launches* lc) {
  if (rows <= 0) return cudaSuccess;
  if (lc) lc->n++;
  l2_normalize_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, ldx, y, ldy, rows, d);
  return cudaGetLastError();
}

__device__ __forceinline__ float sat_half(float x) { return fminf(fmaxf(x, -65504.f), 65504.f); }

// ---- fp32 -> fp16 rows (saturating): A operand of the fp16 tensor-core layers -------------------
__global__ void __launch_bounds__(256) to_half_kernel(const float4* __restrict__ x, uint2* __restrict__ y, long n4) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = x[i];
  const __half2 h0 = __floats2half2_rn(sat_half(v.x), sat_half(v.y)), h1 = __floats2half2_rn(sat_half(v.z), sat_half(v.w));
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&h0);
  u.y = *reinterpret_cast<const uint32_t*>(&h1);
  y[i] = u;
}

cudaError_t to_half_rows(const float* x, __half* y, long n, cudaStream_t st, Launches* lc) {
  if (n <= 0) return cudaSuccess;
  if (n % 4) return cudaErrorInvalidValue;
  if (lc) lc->n++;
  to_half_kernel<<<static_cast<unsigned>((n / 4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<uint2*>(y), n / 4);
  return cudaGetLastError();
}

// ---- LayerNorm (eps 1e-5, biased variance), one warp per row, d <= 1024 ----------------------
template <int D>
__global__ void __launch_bounds__(256) layer_norm_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ w,
                                                         const float* __restrict__ b, int rows, __half* __restrict__ yh, float* __restrict__ yp) {
  constexpr int R = D / 128;  // float4 per lane
  const long r = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + r * D);
  float4 v[R];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    v[i] = xr[i * 32 + lane];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, dd = v[i].w - mean;
    q += (a * a + bb * bb) + (c * c + dd * dd);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + 1e-5f);
  float4* yr = reinterpret_cast<float4*>(y + r * D);
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const float4 ww = reinterpret_cast<const float4*>(w)[i * 32 + lane];
    const float4 bv = reinterpret_cast<const float4*>(b)[i * 32 + lane];
    float4 o;
    o.x = (v[i].x - mean) * rstd * ww.x + bv.x;
    o.y = (v[i].y - mean) * rstd * ww.y + bv.y;
    o.z = (v[i].z - mean) * rstd * ww.z + bv.z;
    o.w = (v[i].w - mean) * rstd * ww.w + bv.w;
    yr[i * 32 + lane] = o;
    if (yp) {  // tf32 hi | lo planes for a following three-pass GEMM (what split_tf32_kernel would write)
      const float4 hi = make_float4(round_tf32(o.x), round_tf32(o.y), round_tf32(o.z), round_tf32(o.w));
      float4* pr = reinterpret_cast<float4*>(yp + r * 2 * D);
      pr[i * 32 + lane] = hi;
      pr[D / 4 + i * 32 + lane] = make_float4(round_tf32(o.x - hi.x), round_tf32(o.y - hi.y), round_tf32(o.z - hi.z), round_tf32(o.w - hi.w));
    }
    if (yh) {  // fp16 copy for the next tensor-core GEMM's A operand
      const __half2 h0 = __floats2half2_rn(sat_half(o.x), sat_half(o.y)), h1 = __floats2half2_rn(sat_half(o.z), sat_half(o.w));
      uint2 u;
      u.x = *reinterpret_cast<const uint32_t*>(&h0);
      u.y = *reinterpret_cast<const uint32_t*>(&h1);
      reinterpret_cast<uint2*>(yh + r * D)[i * 32 + lane] = u;
    }
  }
}

cudaError_t layer_norm_rows(const float* x, float* y, const float* w, const float* b, int rows, int d, cudaStream_t st, Launches* lc,
                            __half* y_half, float* y_planes) {
  if (rows <= 0) return cudaSuccess;
  if (lc) lc->n++;
  const unsigned grid = (rows + 7) / 8;
  if (d == 256) layer_norm_kernel<256><<<grid, 256, 0, st>>>(x, y, w, b, rows, y_half, y_planes);
  else if (d == 128) layer_norm_kernel<128><<<grid, 256, 0, st>>>(x, y, w, b, rows, y_half, y_planes);
  else if (d == 1024) layer_norm_kernel<1024><<<grid, 256, 0, st>>>(x, y, w, b, rows, y_half, y_planes);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// ---- the same on the fp16 residual stream of the token layer: rows in as fp16, out as fp16 only ----------------
// (the normalised row is the next GEMM's A operand AND its residual; statistics and the affine map in fp32).
// Lane l holds columns i * 256 + 8 l .. + 7, i = 0..3: 16-byte loads and stores, 512 contiguous bytes per warp instruction.
__device__ __forceinline__ void unpack_half8(const uint4& u, float (&f)[8]) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&u.z)), d = __half22float2(*reinterpret_cast<const __half2*>(&u.w));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// mean and 1/std of one 1024-wide fp16 row spread over a warp (v[i][e] = column i * 256 + 8 lane + e)
__device__ __forceinline__ void half_row_load(const __half* __restrict__ row, int lane, uint4 (&raw)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) raw[i] = reinterpret_cast<const uint4*>(row)[i * 32 + lane];
}
__device__ __forceinline__ void half_row_stats(const uint4 (&raw)[4], float (&v)[4][8], float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    unpack_half8(raw[i], v[i]);
    s += ((v[i][0] + v[i][1]) + (v[i][2] + v[i][3])) + ((v[i][4] + v[i][5]) + (v[i][6] + v[i][7]));
  }
  mean = warp_sum(s) * (1.f / 1024);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float a = v[i][e] - mean;
      q = fmaf(a, a, q);
    }
  rstd = rsqrtf(warp_sum(q) * (1.f / 1024) + 1e-5f);
}

__global__ void __launch_bounds__(256) layer_norm_half_kernel(const __half* __restrict__ x, __half* __restrict__ y, const float* __restrict__ w,
                                                              const float* __restrict__ b, int rows) {
  const long r = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  float v[4][8], mean, rstd;
  uint4 raw[4];
  half_row_load(x + r * 1024, lane, raw);
  half_row_stats(raw, v, mean, rstd);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 w0 = reinterpret_cast<const float4*>(w)[i * 64 + lane * 2], w1 = reinterpret_cast<const float4*>(w)[i * 64 + lane * 2 + 1];
    const float4 b0 = reinterpret_cast<const float4*>(b)[i * 64 + lane * 2], b1 = reinterpret_cast<const float4*>(b)[i * 64 + lane * 2 + 1];
    const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __half2 h = __floats2half2_rn(sat_half((v[i][2 * e] - mean) * rstd * ww[2 * e] + bb[2 * e]),
                                          sat_half((v[i][2 * e + 1] - mean) * rstd * ww[2 * e + 1] + bb[2 * e + 1]));
      o[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
    reinterpret_cast<uint4*>(y + r * 1024)[i * 32 + lane] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

cudaError_t layer_norm_half_rows(const __half* x, __half* y, const float* w, const float* b, int rows, int d, cudaStream_t st, Launches* lc) {
  if (rows <= 0) return cudaSuccess;
  if (d != 1024) return cudaErrorInvalidValue;
  if (lc) lc->n++;
  layer_norm_half_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, y, w, b, rows);
  return cudaGetLastError();
}

// ---- small-sequence attention core -----------------------------------------------------------
// One warp per (sequence, head, query row).  S <= 32.  Lanes split the head dimension for the
// q.k dot products (coalesced row reads, warp-sum), lane j then holds score j for the softmax,
// and lanes split the head dimension again for P.V.  Key/value rows are fetched four at a time
// so the L2 latency of one group hides behind the arithmetic of the previous one.  No mask:
// padded slots/tokens attend like real ones, exactly as the reference
// (cell_retrieval.py:101-103, language_encoder.py:130-131).
// Generalised to cross attention (nn.TransformerDecoderLayer.multihead_attn, models/cross_matcher.py:113-115): queries come from
// q [n_seq * Sq, ldq], keys and values from k / v [n_seq * Sk, ldkv]; self attention passes the three slices of one packed buffer.
template <int HD>
__global__ void __launch_bounds__(256) mha_small_kernel(const float* __restrict__ qp, long ldq, const float* __restrict__ kp, const float* __restrict__ vp,
                                                        long ldkv, float* __restrict__ out, long n_rows_total, int Sq, int S, int d, int n_heads,
                                                        float scale, int round_out, float* __restrict__ out_planes) {
  constexpr int R = HD / 32;
  constexpr int U = 4;
  // 32-bit index math: the 64-bit runtime divisions this replaced cost more instructions than the attention itself
  const unsigned wid = blockIdx.x * 8u + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= static_cast<unsigned>(n_rows_total) * static_cast<unsigned>(n_heads)) return;
  const int h = static_cast<int>(wid % static_cast<unsigned>(n_heads));
  const long row = wid / static_cast<unsigned>(n_heads);  // seq * Sq + i
  const long seq0 = static_cast<long>((static_cast<unsigned>(row) / static_cast<unsigned>(Sq)) * static_cast<unsigned>(S));  // first key row
  const long ld = ldkv;
  const float* q = qp + row * ldq + h * HD;
  const float* kbase = kp + seq0 * ld + h * HD + lane;
  const float* vbase = vp + seq0 * ld + h * HD + lane;
  float qv[R];
#pragma unroll
  for (int r = 0; r < R; ++r) qv[r] = q[r * 32 + lane];
  float my_score = -INFINITY;
  for (int j0 = 0; j0 < S; j0 += U) {
    float kv[U][R];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = min(j0 + u, S - 1);
#pragma unroll
      for (int r = 0; r < R; ++r) kv[u][r] = kbase[j * ld + r * 32];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float p = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) p = fmaf(qv[r], kv[u][r], p);
      p = warp_sum(p) * scale;
      if (lane == j0 + u && j0 + u < S) my_score = p;
    }
  }
  const float mx = warp_max(my_score);
  const float e = (lane < S) ? expf(my_score - mx) : 0.f;
  const float prob = e / warp_sum(e);
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  for (int j0 = 0; j0 < S; j0 += U) {
    float vv[U][R];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = min(j0 + u, S - 1);
#pragma unroll
      for (int r = 0; r < R; ++r) vv[u][r] = vbase[j * ld + r * 32];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float pj = (j0 + u < S) ? __shfl_sync(0xffffffffu, prob, (j0 + u) & 31) : 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fmaf(pj, vv[u][r], acc[r]);
    }
  }
  if (out) {
    float* o = out + row * d + h * HD;
#pragma unroll
    for (int r = 0; r < R; ++r) o[r * 32 + lane] = round_out ? round_tf32(acc[r]) : acc[r];
  }
  if (out_planes) {  // tf32 hi | lo planes [rows, 2d] for a following three-pass GEMM
    float* o = out_planes + row * 2 * d + h * HD;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float hi = round_tf32(acc[r]);
      o[r * 32 + lane] = hi;
      o[d + r * 32 + lane] = round_tf32(acc[r] - hi);
    }
  }
}

// Sequence-per-warp attention core (head_dim 64 or 32; at most 32 queries and 32 keys per sequence): ONE warp per
// (sequence, head), lane i = query row i.  K, V and Q of the head are staged in shared memory once (coalesced rows); a lane
// then walks the keys with its q row in registers and K / V rows arriving as warp-wide broadcasts, HD independent FMAs per
// key.  The row-per-warp kernel above reads every K and V row S times and spends a 5-step shuffle reduction per (row, key):
// ~4x the instructions (intra-cell layers: 0.28 ms per layer and 2 048 cells; fine-stage decoder layers with 32-wide heads:
// 42 % of the match stage -- profiles/r02).  Self attention (q, k, v slices of one packed buffer) and cross attention
// (queries from one buffer, keys / values from the memory buffer; nn.TransformerDecoderLayer.multihead_attn).
//
// Ragged form (seq_ptr != nullptr, self attention) for the intra-cell layers: the reference zero-pads every cell to 28 object
// slots and attends WITHOUT a mask (cell_retrieval.py:81-103), so the 28 - n padded slots of a cell carry identical rows
// through both layers.  The engine keeps ONE representative of them: cell b owns rows seq_ptr[b] .. seq_ptr[b+1]) = its
// min(n, slots) objects followed, if n < slots, by the padding row, which counts `slots - n` times as a KEY (its softmax
// weight is multiplied by that count) and once as a query.  Same mathematics, 9 rows instead of 28 at 8 objects per cell.
constexpr int kSeqWarps = 4;
template <int HD>
struct SeqAttn {
  static constexpr int QP = HD + 4;                            // Q / O row pitch: conflict-free row-per-lane 16-byte reads
  static constexpr int kWarpFloats = 2 * 32 * HD + 32 * QP;    // K | V | Q rows; the Q block is reused for scores [key][lane] and O
  static constexpr int kSmemBytes = kSeqWarps * kWarpFloats * 4;
  static constexpr int LPR = HD / 4;                           // lanes per staged row (16 bytes each)
  static constexpr int RPI = 32 / LPR;                         // rows per staging instruction
  static_assert(32 * 32 <= 32 * QP, "score buffer fits the Q block");
};

template <int HD>
__global__ void __launch_bounds__(kSeqWarps * 32, 2) mha_seq_kernel(const float* __restrict__ qp, long ldq, const float* __restrict__ kp,
                                                                    const float* __restrict__ vp, long ldkv, float* __restrict__ out, long ldo,
                                                                    int n_units, int Sq_fixed, int Sk_fixed, int n_heads, float scale, int round_out,
                                                                    const int32_t* __restrict__ seq_ptr, const int32_t* __restrict__ cell_ptr, int slots,
                                                                    float* __restrict__ out_planes) {
  using A = SeqAttn<HD>;
  extern __shared__ __align__(16) float seq_smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Ks = seq_smem + w * A::kWarpFloats;
  float* Vs = Ks + 32 * HD;
  float* Qs = Vs + 32 * HD;
  const int rr = lane / A::LPR, c4 = (lane % A::LPR) * 4;
  for (int unit = blockIdx.x * kSeqWarps + w; unit < n_units; unit += gridDim.x * kSeqWarps) {
    const int h = unit % n_heads;
    const int seq = unit / n_heads;
    long qrow0 = static_cast<long>(seq) * Sq_fixed, krow0 = static_cast<long>(seq) * Sk_fixed;
    int Sq = Sq_fixed, Sk = Sk_fixed;
    float last_weight = 1.f;  // multiplicity of the last key
    if (seq_ptr) {
      qrow0 = krow0 = seq_ptr[seq];
      Sq = Sk = seq_ptr[seq + 1] - seq_ptr[seq];
      const int n_obj = cell_ptr[seq + 1] - cell_ptr[seq];
      if (n_obj < slots) last_weight = static_cast<float>(slots - n_obj);
    }
    const float* qb = qp + qrow0 * ldq + h * HD + c4;
    const float* kb = kp + krow0 * ldkv + h * HD + c4;
    const float* vb = vp + krow0 * ldkv + h * HD + c4;
    for (int j = rr; j < Sq; j += A::RPI) *reinterpret_cast<float4*>(Qs + j * A::QP + c4) = *reinterpret_cast<const float4*>(qb + j * ldq);
    for (int j = rr; j < Sk; j += A::RPI) {
      *reinterpret_cast<float4*>(Ks + j * HD + c4) = *reinterpret_cast<const float4*>(kb + j * ldkv);
      *reinterpret_cast<float4*>(Vs + j * HD + c4) = *reinterpret_cast<const float4*>(vb + j * ldkv);
    }
    __syncwarp();
    const int i = lane < Sq ? lane : Sq - 1;  // idle lanes shadow the last row (no divergence); they store nothing
    float q[HD];
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      const float4 t = *reinterpret_cast<const float4*>(Qs + i * A::QP + c * 4);
      q[4 * c] = t.x; q[4 * c + 1] = t.y; q[4 * c + 2] = t.z; q[4 * c + 3] = t.w;
    }
    __syncwarp();  // every lane holds its q row: Qs becomes the score buffer [key][lane]
    float mx = -INFINITY;
#pragma unroll 2
    for (int j = 0; j < Sk; ++j) {
      float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        const float4 k4 = *reinterpret_cast<const float4*>(Ks + j * HD + c * 4);  // same address in every lane: broadcast
        p0 = fmaf(q[4 * c], k4.x, p0); p1 = fmaf(q[4 * c + 1], k4.y, p1); p2 = fmaf(q[4 * c + 2], k4.z, p2); p3 = fmaf(q[4 * c + 3], k4.w, p3);
      }
      const float sc = ((p0 + p1) + (p2 + p3)) * scale;
      Qs[j * 32 + lane] = sc;
      mx = fmaxf(mx, sc);
    }
    float sum = 0.f;
    for (int j = 0; j < Sk; ++j) {
      float e = expf(Qs[j * 32 + lane] - mx);
      if (j == Sk - 1) e *= last_weight;
      Qs[j * 32 + lane] = e;
      sum += e;
    }
    const float inv = 1.f / sum;
    float acc[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] = 0.f;
#pragma unroll 2
    for (int j = 0; j < Sk; ++j) {
      const float pj = Qs[j * 32 + lane] * inv;
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        const float4 v4 = *reinterpret_cast<const float4*>(Vs + j * HD + c * 4);
        acc[4 * c] = fmaf(pj, v4.x, acc[4 * c]); acc[4 * c + 1] = fmaf(pj, v4.y, acc[4 * c + 1]);
        acc[4 * c + 2] = fmaf(pj, v4.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(pj, v4.w, acc[4 * c + 3]);
      }
    }
    __syncwarp();  // scores are dead: Qs becomes the output staging [row][QP]
    if (lane < Sq) {
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        float4 o = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
        if (round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
        *reinterpret_cast<float4*>(Qs + lane * A::QP + c * 4) = o;
      }
    }
    __syncwarp();
    if (out) {
      float* obase = out + qrow0 * ldo + h * HD + c4;
      for (int j = rr; j < Sq; j += A::RPI) *reinterpret_cast<float4*>(obase + static_cast<long>(j) * ldo) = *reinterpret_cast<const float4*>(Qs + j * A::QP + c4);
    }
    if (out_planes) {  // tf32 hi | lo planes [rows, 2 ldo] for a following three-pass GEMM
      float* pbase = out_planes + qrow0 * 2 * ldo + h * HD + c4;
      for (int j = rr; j < Sq; j += A::RPI) {
        const float4 o = *reinterpret_cast<const float4*>(Qs + j * A::QP + c4);
        const float4 hi = make_float4(round_tf32(o.x), round_tf32(o.y), round_tf32(o.z), round_tf32(o.w));
        *reinterpret_cast<float4*>(pbase + static_cast<long>(j) * 2 * ldo) = hi;
        *reinterpret_cast<float4*>(pbase + static_cast<long>(j) * 2 * ldo + ldo) =
            make_float4(round_tf32(o.x - hi.x), round_tf32(o.y - hi.y), round_tf32(o.z - hi.z), round_tf32(o.w - hi.w));
      }
    }
    __syncwarp();  // before the next unit's staging overwrites the buffers
  }
}

template <int HD>
static cudaError_t launch_seq(const float* q, long ldq, const float* k, const float* v, long ldkv, float* out, int n_seq, int Sq, int Sk, int d,
                              int n_heads, int round_out, const int32_t* seq_ptr, const int32_t* cell_ptr, int slots, cudaStream_t st, Launches* lc,
                              float* out_planes = nullptr) {
  static bool configured_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured_dev[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(mha_seq_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SeqAttn<HD>::kSmemBytes);
    if (e != cudaSuccess) return e;
    configured_dev[dev & 63] = true;
  }
  if (lc) lc->n++;
  const int n_units = n_seq * n_heads;
  const int blocks = (n_units + kSeqWarps - 1) / kSeqWarps;
  mha_seq_kernel<HD><<<blocks, kSeqWarps * 32, SeqAttn<HD>::kSmemBytes, st>>>(q, ldq, k, v, ldkv, out, d, n_units, Sq, Sk, n_heads,
                                                                               1.f / sqrtf(static_cast<float>(HD)), round_out, seq_ptr, cell_ptr, slots,
                                                                               out_planes);
  return cudaGetLastError();
}

cudaError_t mha_cells64(const float* qkv, float* out, int n_cells, const int32_t* row_ptr_dev, const int32_t* cell_ptr_dev, int slots, int d,
                        int n_heads, cudaStream_t st, Launches* lc, float* out_planes) {
  if (n_cells <= 0) return cudaSuccess;
  if (d != 64 * n_heads || slots < 1 || slots > 32 || static_cast<long>(n_cells) * n_heads >= (1L << 31)) return cudaErrorInvalidValue;
  return launch_seq<64>(qkv, 3L * d, qkv + d, qkv + 2 * d, 3L * d, out, n_cells, 0, 0, d, n_heads, 0, row_ptr_dev, cell_ptr_dev, slots, st, lc, out_planes);
}

cudaError_t mha_cross_small(const float* q, long ldq, const float* k, const float* v, long ldkv, float* out, int n_seq, int Sq, int Sk, int d,
                            int n_heads, cudaStream_t st, Launches* lc, int round_out, float* out_planes) {
  if (n_seq <= 0) return cudaSuccess;
  if (Sk > 32 || Sk < 1 || Sq < 1) return cudaErrorInvalidValue;
  const long rows = static_cast<long>(n_seq) * Sq;
  const int hd = d / n_heads;
  if (rows * n_heads >= (1L << 31)) return cudaErrorInvalidValue;
  // sequence-per-warp core: 32-wide heads (fine-stage decoder layers) always, 64-wide heads for the longer sequences
  const bool aligned = !(ldq % 4) && !(ldkv % 4) && !(d % 4) &&
                       !((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(out_planes)) & 15);
  if (aligned && Sq <= 32 && hd * n_heads == d) {
    // (Packing two or four short sequences into one warp -- all 32 lanes busy at 16 or 6 query rows -- was built and measured: 1.71
    // against 1.53 ms per fine-stage pass.  The core is bound by its loads and stores at eight warps per SM, not by FMA issue.)
    if (hd == 32) return launch_seq<32>(q, ldq, k, v, ldkv, out, n_seq, Sq, Sk, d, n_heads, round_out, nullptr, nullptr, 0, st, lc, out_planes);
    if (hd == 64 && Sq > 16) return launch_seq<64>(q, ldq, k, v, ldkv, out, n_seq, Sq, Sk, d, n_heads, round_out, nullptr, nullptr, 0, st, lc, out_planes);
  }
  if (lc) lc->n++;
  const unsigned grid = static_cast<unsigned>((rows * n_heads + 7) / 8);
  const float scale = 1.f / sqrtf(static_cast<float>(hd));
  if (hd == 64) mha_small_kernel<64><<<grid, 256, 0, st>>>(q, ldq, k, v, ldkv, out, rows, Sq, Sk, d, n_heads, scale, round_out, out_planes);
  else if (hd == 32) mha_small_kernel<32><<<grid, 256, 0, st>>>(q, ldq, k, v, ldkv, out, rows, Sq, Sk, d, n_heads, scale, round_out, out_planes);
  else if (hd == 256) mha_small_kernel<256><<<grid, 256, 0, st>>>(q, ldq, k, v, ldkv, out, rows, Sq, Sk, d, n_heads, scale, round_out, out_planes);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t mha_small(const float* qkv, float* out, int n_seq, int S, int d, int n_heads, cudaStream_t st, Launches* lc, int round_out,
                      float* out_planes) {
  return mha_cross_small(qkv, 3L * d, qkv + d, qkv + 2 * d, 3L * d, out, n_seq, S, S, d, n_heads, st, lc, round_out, out_planes);
}

// ---- tensor-core attention core for the token layer (head_dim 256) ---------------------------------
// One warp per (sentence, head): scores = Q K^T and O = P V as warp-level mma.sync m16n8k8 tf32 tiles,
// softmax in the accumulator fragments.  Fragment index tricks keep every global access a
// coalesced 128-bit one and need no shuffles between the two products:
//   * the k index of an MMA is a dummy summation index, so a lane's 8 contiguous floats of a Q/K row
//     serve as the (k = t, k = t+4) elements of four successive k-steps;
//   * the score accumulator layout (row g, keys 2t / 2t+1) is re-read as the A fragment of P V by
//     declaring key 8j+2t <-> k = t and key 8j+2t+1 <-> k = t+4, and loading V rows in that order;
//   * output columns are permuted the same way, so each lane ends with 8 contiguous floats of a row.
// Operands are rounded to tf32 (rna) like every other tensor-core layer of the token encoder.
// NO mask: padded tokens attend like real ones (language_encoder.py:130-131).
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tf32_bits(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// 8 / 4 consecutive elements of a qkv row as fp32 (fp16 -> fp32 is exact, and every fp16 value is a tf32 value)
template <bool HALF_IN>
__device__ __forceinline__ void ld8(const void* base, long off, float4& a, float4& b) {
  if (HALF_IN) {
    const uint4 u = *reinterpret_cast<const uint4*>(static_cast<const __half*>(base) + off);
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&u.z)), f3 = __half22float2(*reinterpret_cast<const __half2*>(&u.w));
    a = make_float4(f0.x, f0.y, f1.x, f1.y);
    b = make_float4(f2.x, f2.y, f3.x, f3.y);
  } else {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + off);
    a = p[0];
    b = p[1];
  }
}
template <bool HALF_IN>
__device__ __forceinline__ float4 ld4(const void* base, long off) {
  if (HALF_IN) {
    const uint2 u = *reinterpret_cast<const uint2*>(static_cast<const __half*>(base) + off);
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(f0.x, f0.y, f1.x, f1.y);
  }
  return *reinterpret_cast<const float4*>(static_cast<const float*>(base) + off);
}

template <int MT, bool HALF_IN>  // MT 16-row query tiles (S <= 16 * MT), 2 * MT key tiles of 8
__global__ void __launch_bounds__(128) mha_tc256_kernel(const void* __restrict__ qkv, float* __restrict__ out, long n_seq, int S, float scale,
                                                        int round_out) {
  constexpr int HD = 256, D = 1024, LD = 3 * D, NT = 2 * MT;
  const long wid = static_cast<long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (wid >= n_seq * 4) return;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int h = static_cast<int>(wid & 3);
  const long base = (wid >> 2) * S * LD + h * HD;
  auto row_off = [&](int r, int which) { return base + static_cast<long>(min(r, S - 1)) * LD + which * D; };  // clamp: rows >= S are never stored / are masked

  // ---- scores[MT*16, NT*8] = Q K^T
  float sc[MT][NT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) sc[m][n][i] = 0.f;
#pragma unroll 2
  for (int s = 0; s < HD / 32; ++s) {  // 32 head-dim columns per step: lane owns columns 32 s + 8 t .. + 7 of its rows
    float4 qa[MT][2][2], kb[NT][2];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) ld8<HALF_IN>(qkv, row_off(16 * m + 8 * hi + g, 0) + 32 * s + 8 * t, qa[m][hi][0], qa[m][hi][1]);
#pragma unroll
    for (int n = 0; n < NT; ++n) ld8<HALF_IN>(qkv, row_off(8 * n + g, 1) + 32 * s + 8 * t, kb[n][0], kb[n][1]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // four k-steps: elements (2j, 2j+1) of the lane's 8 floats
      auto pick = [&](const float4 (&v)[2], int e) { const float* f = reinterpret_cast<const float*>(v); return tf32_bits(f[e]); };
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const uint32_t a0 = pick(qa[m][0], 2 * j), a1 = pick(qa[m][1], 2 * j), a2 = pick(qa[m][0], 2 * j + 1), a3 = pick(qa[m][1], 2 * j + 1);
#pragma unroll
        for (int n = 0; n < NT; ++n) mma_tf32_16x8x8(sc[m][n], a0, a1, a2, a3, pick(kb[n], 2 * j), pick(kb[n], 2 * j + 1));
      }
    }
  }
  // ---- softmax over keys: row (16 m + g [+8]) lives in the 4 lanes of a quad; keys 8 n + 2 t, + 1
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      float mx = -INFINITY;
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float& v = sc[m][n][2 * hi + e];
          v = (8 * n + 2 * t + e < S) ? v * scale : -INFINITY;
          mx = fmaxf(mx, v);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float& v = sc[m][n][2 * hi + e];
          v = expf(v - mx);
          sum += v;
        }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = 1.f / sum;
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) sc[m][n][2 * hi + e] *= inv;
    }
  // ---- O = P V, 32 output columns at a time; lane (g, t) ends with columns 32 s + 8 t .. + 7 of rows g, g + 8
  uint32_t pa[MT][NT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) {  // C layout (c0,c1 | c2,c3) -> A layout (a0 = c0, a1 = c2, a2 = c1, a3 = c3) under the key permutation
      pa[m][n][0] = tf32_bits(sc[m][n][0]); pa[m][n][1] = tf32_bits(sc[m][n][2]);
      pa[m][n][2] = tf32_bits(sc[m][n][1]); pa[m][n][3] = tf32_bits(sc[m][n][3]);
    }
#pragma unroll 1
  for (int s = 0; s < HD / 32; ++s) {
    float o[MT][4][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) o[m][j][i] = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {  // k-step over keys 8 n .. 8 n + 7: k = t <-> key 8 n + 2 t, k = t + 4 <-> key 8 n + 2 t + 1
      const float4 v0 = ld4<HALF_IN>(qkv, row_off(8 * n + 2 * t, 2) + 32 * s + 4 * g);
      const float4 v1 = ld4<HALF_IN>(qkv, row_off(8 * n + 2 * t + 1, 2) + 32 * s + 4 * g);
      const float f0[4] = {v0.x, v0.y, v0.z, v0.w}, f1[4] = {v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // n-tile j: output column (n index g) = 32 s + 4 g + j
        const uint32_t b0 = tf32_bits(f0[j]), b1 = tf32_bits(f1[j]);
#pragma unroll
        for (int m = 0; m < MT; ++m) mma_tf32_16x8x8(o[m][j], pa[m][n][0], pa[m][n][1], pa[m][n][2], pa[m][n][3], b0, b1);
      }
    }
    // accumulator (row g | g+8, n = 2 t | 2 t + 1) of tile j is column 32 s + 8 t + j | + 4 + j
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        const int r = 16 * m + 8 * hi + g;
        if (r < S) {
          float4 lo = make_float4(o[m][0][2 * hi], o[m][1][2 * hi], o[m][2][2 * hi], o[m][3][2 * hi]);
          float4 up = make_float4(o[m][0][2 * hi + 1], o[m][1][2 * hi + 1], o[m][2][2 * hi + 1], o[m][3][2 * hi + 1]);
          const long o_off = ((wid >> 2) * S + r) * D + h * HD + 32 * s + 8 * t;
          if (round_out == 2) {  // fp16 output (out is __half [rows, D]): attention outputs are convex combinations of V rows
            const __half2 h0 = __floats2half2_rn(sat_half(lo.x), sat_half(lo.y)), h1 = __floats2half2_rn(sat_half(lo.z), sat_half(lo.w));
            const __half2 h2 = __floats2half2_rn(sat_half(up.x), sat_half(up.y)), h3 = __floats2half2_rn(sat_half(up.z), sat_half(up.w));
            uint4 u;
            u.x = *reinterpret_cast<const uint32_t*>(&h0); u.y = *reinterpret_cast<const uint32_t*>(&h1);
            u.z = *reinterpret_cast<const uint32_t*>(&h2); u.w = *reinterpret_cast<const uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(out) + o_off) = u;
            continue;
          }
          if (round_out) {
            lo = make_float4(round_tf32(lo.x), round_tf32(lo.y), round_tf32(lo.z), round_tf32(lo.w));
            up = make_float4(round_tf32(up.x), round_tf32(up.y), round_tf32(up.z), round_tf32(up.w));
          }
          float4* dst = reinterpret_cast<float4*>(out + o_off);
          dst[0] = lo;
          dst[1] = up;
        }
      }
  }
}

cudaError_t mha_tc256(const void* qkv, float* out, int n_seq, int S, cudaStream_t st, Launches* lc, int round_out, int half_in) {
  if (n_seq <= 0) return cudaSuccess;
  if (S < 1 || S > 32) return cudaErrorInvalidValue;
  if (lc) lc->n++;
  const unsigned grid = static_cast<unsigned>((static_cast<long>(n_seq) * 4 + 3) / 4);
  const float scale = 1.f / 16.f;  // 1 / sqrt(256)
  if (half_in) {
    if (S <= 16) mha_tc256_kernel<1, true><<<grid, 128, 0, st>>>(qkv, out, n_seq, S, scale, round_out);
    else mha_tc256_kernel<2, true><<<grid, 128, 0, st>>>(qkv, out, n_seq, S, scale, round_out);
  } else {
    if (S <= 16) mha_tc256_kernel<1, false><<<grid, 128, 0, st>>>(qkv, out, n_seq, S, scale, round_out);
    else mha_tc256_kernel<2, false><<<grid, 128, 0, st>>>(qkv, out, n_seq, S, scale, round_out);
  }
  return cudaGetLastError();
}

// ---- LayerNorm over the rows of each group, then max over the group's rows: y[g, :] = max_s LN(x[g*S + s, :]) ----
// (the token layer's norm2 followed by the max over tokens, language_encoder.py:131-133: the normalised
// [tokens, 1024] tensor is never written).  One warp per group.
__global__ void __launch_bounds__(256) layer_norm_max_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ w,
                                                             const float* __restrict__ b, int groups, int S) {
  constexpr int D = 1024, R = D / 128;
  const long g = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (g >= groups) return;
  const int lane = threadIdx.x & 31;
  float4 ww[R], bv[R], m[R];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    ww[i] = reinterpret_cast<const float4*>(w)[i * 32 + lane];
    bv[i] = reinterpret_cast<const float4*>(b)[i * 32 + lane];
    m[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  }
  for (int s = 0; s < S; ++s) {
    const float4* xr = reinterpret_cast<const float4*>(x + (g * S + s) * D);
    float4 v[R];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      v[i] = xr[i * 32 + lane];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(sum) * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, dd = v[i].w - mean;
      q += (a * a + bb * bb) + (c * c + dd * dd);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + 1e-5f);
#pragma unroll
    for (int i = 0; i < R; ++i) {
      m[i].x = fmaxf(m[i].x, (v[i].x - mean) * rstd * ww[i].x + bv[i].x);
      m[i].y = fmaxf(m[i].y, (v[i].y - mean) * rstd * ww[i].y + bv[i].y);
      m[i].z = fmaxf(m[i].z, (v[i].z - mean) * rstd * ww[i].z + bv[i].z);
      m[i].w = fmaxf(m[i].w, (v[i].w - mean) * rstd * ww[i].w + bv[i].w);
    }
  }
  float4* yr = reinterpret_cast<float4*>(y + g * D);
#pragma unroll
  for (int i = 0; i < R; ++i) yr[i * 32 + lane] = m[i];
}

// fp16 rows in (the token layer's fp16 residual stream), fp32 maxima out.  One CTA of four warps per group: warp w takes
// rows w, w + 4, ... with the next row's loads in flight while the current one is reduced, and the four partial maxima meet in
// shared memory (one warp per group walked its rows one dependent reduction at a time: 1.8 TB/s, profiles/r02).
__global__ void __launch_bounds__(128) layer_norm_max_half_kernel(const __half* __restrict__ x, float* __restrict__ y, const float* __restrict__ w,
                                                                  const float* __restrict__ b, int groups, int S) {
  __shared__ float part[4][1024];
  const long g = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float m[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) m[i][e] = -INFINITY;
  uint4 raw[4], nxt[4];
  if (wid < S) half_row_load(x + (g * S + wid) * 1024, lane, raw);
  for (int s = wid; s < S; s += 4) {
    if (s + 4 < S) half_row_load(x + (g * S + s + 4) * 1024, lane, nxt);
    float v[4][8], mean, rstd;
    half_row_stats(raw, v, mean, rstd);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {  // weight and bias from L1 (4 KB each, shared by every row): keeps the kernel at 4+ CTAs per SM
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(w) + i * 64 + lane * 2 + h), b4 = __ldg(reinterpret_cast<const float4*>(b) + i * 64 + lane * 2 + h);
        m[i][4 * h + 0] = fmaxf(m[i][4 * h + 0], (v[i][4 * h + 0] - mean) * rstd * w4.x + b4.x);
        m[i][4 * h + 1] = fmaxf(m[i][4 * h + 1], (v[i][4 * h + 1] - mean) * rstd * w4.y + b4.y);
        m[i][4 * h + 2] = fmaxf(m[i][4 * h + 2], (v[i][4 * h + 2] - mean) * rstd * w4.z + b4.z);
        m[i][4 * h + 3] = fmaxf(m[i][4 * h + 3], (v[i][4 * h + 3] - mean) * rstd * w4.w + b4.w);
      }
      raw[i] = nxt[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    reinterpret_cast<float4*>(part[wid])[i * 64 + lane * 2] = make_float4(m[i][0], m[i][1], m[i][2], m[i][3]);
    reinterpret_cast<float4*>(part[wid])[i * 64 + lane * 2 + 1] = make_float4(m[i][4], m[i][5], m[i][6], m[i][7]);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 2; ++i) {  // 128 threads x 2 float4 = 1024 columns
    const int c4 = i * 128 + threadIdx.x;
    float4 o = reinterpret_cast<const float4*>(part[0])[c4];
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float4 t = reinterpret_cast<const float4*>(part[k])[c4];
      o.x = fmaxf(o.x, t.x); o.y = fmaxf(o.y, t.y); o.z = fmaxf(o.z, t.z); o.w = fmaxf(o.w, t.w);
    }
    reinterpret_cast<float4*>(y + g * 1024)[c4] = o;
  }
}

cudaError_t layer_norm_max_half_rows(const __half* x, float* y, const float* w, const float* b, int groups, int S, int d, cudaStream_t st,
                                     Launches* lc) {
  if (groups <= 0) return cudaSuccess;
  if (d != 1024 || S < 1) return cudaErrorInvalidValue;
  if (lc) lc->n++;
  layer_norm_max_half_kernel<<<groups, 128, 0, st>>>(x, y, w, b, groups, S);
  return cudaGetLastError();
}

cudaError_t layer_norm_max_rows(const float* x, float* y, const float* w, const float* b, int groups, int S, int d, cudaStream_t st,
                                Launches* lc) {
  if (groups <= 0) return cudaSuccess;
  if (d != 1024 || S < 1) return cudaErrorInvalidValue;
  if (lc) lc->n++;
  layer_norm_max_kernel<<<(groups + 7) / 8, 256, 0, st>>>(x, y, w, b, groups, S);
  return cudaGetLastError();
}

// ---- max over the rows of a group ---------------------------------------------------------------
__global__ void max_over_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int groups, int S, int d4) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long>(groups) * d4) return;
  const long g = i / d4;
  const int c = static_cast<int>(i % d4);
  const float4* p = reinterpret_cast<const float4*>(x) + g * S * d4 + c;
  float4 m = p[0];
  for (int s = 1; s < S; ++s) {
    const float4 v = p[static_cast<long>(s) * d4];
    m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
  }
  reinterpret_cast<float4*>(y)[i] = m;
}

cudaError_t max_over_rows(const float* x, float* y, int groups, int S, int d, cudaStream_t st, Launches* lc) {
  if (groups <= 0) return cudaSuccess;
  if (d % 4) return cudaErrorInvalidValue;
  if (lc) lc->n++;
  const long n = static_cast<long>(groups) * (d / 4);
  max_over_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(x, y, groups, S, d / 4);
  return cudaGetLastError();
}

// ---- objects -> packed attention rows of the intra-cell layers ---------------------------------------
// (the reference builds a zero-padded [B, 28, 256] tensor, cell_retrieval.py:85-98)
// Cell b -> rows row_ptr[b] ..: its min(n, 28) normalised objects, then ONE
// zero row standing for all 28 - n padded slots (see mha_seq64_kernel).  One warp per cell.
__global__ void scatter_objects_ragged_kernel(const float* __restrict__ emb, const int32_t* __restrict__ cell_ptr, const int32_t* __restrict__ row_ptr,
                                              int n_cells, float* __restrict__ X) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= n_cells) return;
  const int lane = threadIdx.x & 31;
  const int n = cell_ptr[b + 1] - cell_ptr[b];
  const int rows = row_ptr[b + 1] - row_ptr[b];
  for (int s = 0; s < rows; ++s) {
    float4* dst = reinterpret_cast<float4*>(X + static_cast<long>(row_ptr[b] + s) * kEmbed);
    if (s >= n) {
      dst[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
      dst[lane + 32] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const float4* src = reinterpret_cast<const float4*>(emb + static_cast<long>(cell_ptr[b] + s) * kEmbed);
    const float4 a = src[lane], c = src[lane + 32];
    float ss = (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w) + (c.x * c.x + c.y * c.y) + (c.z * c.z + c.w * c.w);
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    dst[lane] = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
    dst[lane + 32] = make_float4(c.x * inv, c.y * inv, c.z * inv, c.w * inv);
  }
}

cudaError_t scatter_objects_ragged(const float* emb, const int32_t* cell_ptr_dev, const int32_t* row_ptr_dev, int n_cells, float* X, cudaStream_t st,
                                   Launches* lc) {
  if (n_cells <= 0) return cudaSuccess;
  if (lc) lc->n++;
  scatter_objects_ragged_kernel<<<(n_cells + 7) / 8, 256, 0, st>>>(emb, cell_ptr_dev, row_ptr_dev, n_cells, X);
  return cudaGetLastError();
}

// y[b, :] = max over the rows row_ptr[b] .. row_ptr[b+1]) of x (d = 256); one warp per cell
__global__ void max_over_rows_ragged_kernel(const float* __restrict__ x, const int32_t* __restrict__ row_ptr, float* __restrict__ y, int n_cells) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= n_cells) return;
  const int lane = threadIdx.x & 31;
  const int r0 = row_ptr[b], r1 = row_ptr[b + 1];
  float4 m0 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), m1 = m0;
  for (int r = r0; r < r1; ++r) {
    const float4* src = reinterpret_cast<const float4*>(x + static_cast<long>(r) * kEmbed);
    const float4 a = src[lane], c = src[lane + 32];
    m0.x = fmaxf(m0.x, a.x); m0.y = fmaxf(m0.y, a.y); m0.z = fmaxf(m0.z, a.z); m0.w = fmaxf(m0.w, a.w);
    m1.x = fmaxf(m1.x, c.x); m1.y = fmaxf(m1.y, c.y); m1.z = fmaxf(m1.z, c.z); m1.w = fmaxf(m1.w, c.w);
  }
  float4* dst = reinterpret_cast<float4*>(y + static_cast<long>(b) * kEmbed);
  dst[lane] = m0;
  dst[lane + 32] = m1;
}

cudaError_t max_over_rows_ragged(const float* x, const int32_t* row_ptr_dev, float* y, int n_cells, cudaStream_t st, Launches* lc) {
  if (n_cells <= 0) return cudaSuccess;
  if (lc) lc->n++;
  max_over_rows_ragged_kernel<<<(n_cells + 7) / 8, 256, 0, st>>>(x, row_ptr_dev, y, n_cells);
  return cudaGetLastError();
}

// ---- colour / position / point-count encoders of ObjectEncoder.forward in ONE kernel ---------------------------
// (models/object_encoder.py:122-145: three get_mlp([k, 64, 256]) stacks = Linear+BN(folded)+ReLU twice, each followed by
// F.normalize).  The seven SIMT GEMM launches + three normalisations they replace moved 4 096 x 256 floats each and were
// launch/latency-bound (24 us apiece, profiles/r01).  Thread c of a 256-thread block owns output channel c: its 64-float
// second-layer weight row lives in registers and is reused for every object the block visits; the 64 hidden units of
// eight objects at a time are staged in shared memory.  Same operation order as linear_simt (fma chain over k, then bias).
struct SideEncoders {
  const float* w1[3]; const float* b1[3]; const float* w2[3]; const float* b2[3];  // w1 [64, ld 4], w2 [256, 64]
};

template <int D>  // embedding dim = threads per block: 256 (coarse) or 128 (fine stage)
__global__ void __launch_bounds__(D) side_encoders_kernel(const float* __restrict__ meta, int n, SideEncoders w, float* __restrict__ cat) {
  constexpr int OB = 8;
  constexpr int OPP = D / 64;  // objects whose 64 hidden units one pass of the block computes
  __shared__ float hid[OB][64];
  __shared__ float red[D / 32][OB];
  const int c = threadIdx.x, lane = c & 31, wp = c >> 5;
  const float mean = static_cast<float>(1826.6844940968194), std_ = static_cast<float>(2516.8905096993817);
  for (int enc = 0; enc < 3; ++enc) {
    float w2r[64];
#pragma unroll
    for (int k = 0; k < 64; k += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(w.w2[enc] + c * 64 + k));
      w2r[k] = v.x; w2r[k + 1] = v.y; w2r[k + 2] = v.z; w2r[k + 3] = v.w;
    }
    const float b2c = __ldg(w.b2[enc] + c);
    const int h = c & 63;
    const float4 w1h = __ldg(reinterpret_cast<const float4*>(w.w1[enc] + h * 4));
    const float b1h = __ldg(w.b1[enc] + h);
    for (int o0 = blockIdx.x * OB; o0 < n; o0 += gridDim.x * OB) {
#pragma unroll
      for (int pass = 0; pass < OB / OPP; ++pass) {  // D threads = OPP objects x 64 hidden units per pass
        const int ol = pass * OPP + (c >> 6), o = o0 + ol;
        float x = 0.f;
        if (o < n) {
          const float* m = meta + static_cast<long>(o) * 7;
          if (enc == 2) {
            x = fmaf(__fdiv_rn(__fsub_rn(m[6], mean), std_), w1h.x, 0.f);  // (count - mean) / std (object_encoder.py:141-143)
          } else {
            x = fmaf(m[enc * 3 + 0], w1h.x, 0.f);
            x = fmaf(m[enc * 3 + 1], w1h.y, x);
            x = fmaf(m[enc * 3 + 2], w1h.z, x);
          }
          x = fmaxf(x + b1h, 0.f);
        }
        hid[ol][h] = x;
      }
      __syncthreads();
      float acc[OB];
#pragma unroll
      for (int ol = 0; ol < OB; ++ol) acc[ol] = 0.f;
#pragma unroll
      for (int k = 0; k < 64; ++k)
#pragma unroll
        for (int ol = 0; ol < OB; ++ol) acc[ol] = fmaf(hid[ol][k], w2r[k], acc[ol]);
#pragma unroll
      for (int ol = 0; ol < OB; ++ol) {
        acc[ol] = fmaxf(acc[ol] + b2c, 0.f);
        const float ss = warp_sum(acc[ol] * acc[ol]);
        if (lane == 0) red[wp][ol] = ss;
      }
      __syncthreads();
#pragma unroll
      for (int ol = 0; ol < OB; ++ol) {
        if (o0 + ol >= n) break;
        float ss = 0.f;
#pragma unroll
        for (int q = 0; q < D / 32; ++q) ss += red[q][ol];
        const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize
        cat[static_cast<long>(o0 + ol) * (4 * D) + D * (enc + 1) + c] = acc[ol] * inv;
      }
      __syncthreads();
    }
  }
}

cudaError_t side_encoders(const float* meta, int n_obj, const float* const* w1, const float* const* b1, const float* const* w2, const float* const* b2,
                          float* cat, int d, cudaStream_t st, Launches* lc) {
  if (n_obj <= 0) return cudaSuccess;
  if (lc) lc->n++;
  SideEncoders w;
  for (int i = 0; i < 3; ++i) { w.w1[i] = w1[i]; w.b1[i] = b1[i]; w.w2[i] = w2[i]; w.b2[i] = b2[i]; }
  const int batches = (n_obj + 7) / 8;
  const int grid = batches < 296 ? batches : 296;
  if (d == 256) side_encoders_kernel<256><<<grid, 256, 0, st>>>(meta, n_obj, w, cat);
  else if (d == 128) side_encoders_kernel<128><<<grid, 128, 0, st>>>(meta, n_obj, w, cat);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// dst[(p * G + g), :] = src[(idx ? idx[p] : p0 + p) * G + g, :]: the G rows of group idx[p] for every pair p (fine stage: the
// objects of the pair's cell, the hints of the pair's query)
__global__ void gather_row_groups_kernel(const float4* __restrict__ src, const int32_t* __restrict__ idx, int p0, long n_rows, int G, int d4,
                                         float4* __restrict__ dst) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_rows * d4) return;
  const long row = i / d4;
  const int c = static_cast<int>(i - row * d4);
  const long p = row / G;
  const int g = static_cast<int>(row - p * G);
  const long s = (idx ? static_cast<long>(idx[p]) : p0 + p) * G + g;
  dst[i] = src[s * d4 + c];
}

cudaError_t gather_row_groups(const float* src, const int32_t* idx, int p0, int n_groups, int G, int d, float* dst, cudaStream_t st, Launches* lc) {
  if (n_groups <= 0) return cudaSuccess;
  if (d % 4) return cudaErrorInvalidValue;
  if (lc) lc->n++;
  const long n = static_cast<long>(n_groups) * G * (d / 4);
  gather_row_groups_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(src), idx, p0, static_cast<long>(n_groups) * G, G,
                                                                                 d / 4, reinterpret_cast<float4*>(dst));
  return cudaGetLastError();
}

__global__ void add_rows_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ y, long n4) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 p = a[i], q = b[i];
  y[i] = make_float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w);
}

cudaError_t add_rows(const float* a, const float* b, float* y, long n, cudaStream_t st, Launches* lc) {
  if (n <= 0) return cudaSuccess;
  if (n % 4) return cudaErrorInvalidValue;
  if (lc) lc->n++;
  add_rows_kernel<<<static_cast<unsigned>((n / 4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                                                           reinterpret_cast<float4*>(y), n / 4);
  return cudaGetLastError();
}

}  // namespace t2l
