// Row-wise fp32 kernels around the GEMMs: F.normalize, LayerNorm, the small-sequence unmasked
// multi-head attention core, max pooling over a sequence, object scatter into the padded cell
// tensor.  All are HBM/L2-bound streaming kernels: one warp per row, 128-bit accesses where the
// row length allows.
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

// ---- LayerNorm (eps 1e-5, biased variance), one warp per row, d <= 1024 ----------------------
template <int D>
__global__ void __launch_bounds__(256) layer_norm_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ w,
                                                         const float* __restrict__ b, int rows) {
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
  }
}

cudaError_t layer_norm_rows(const float* x, float* y, const float* w, const float* b, int rows, int d, cudaStream_t st, Launches* lc) {
  if (rows <= 0) return cudaSuccess;
  if (lc) lc->n++;
  const unsigned grid = (rows + 7) / 8;
  if (d == 256) layer_norm_kernel<256><<<grid, 256, 0, st>>>(x, y, w, b, rows);
  else if (d == 1024) layer_norm_kernel<1024><<<grid, 256, 0, st>>>(x, y, w, b, rows);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// ---- small-sequence attention core -----------------------------------------------------------
// One warp per (sequence, head, query row).  S <= 32.  Lanes split the head dimension for the
// q.k dot products (coalesced row reads, warp-sum), lane j then holds score j for the softmax,
// and lanes split the head dimension again for P.V.  Key/value rows are fetched four at a time
// so the L2 latency of one group hides behind the arithmetic of the previous one.  No mask:
// padded slots/tokens attend like real ones, exactly as the reference
// (cell_retrieval.py:101-103, language_encoder.py:130-131).
template <int HD>
__global__ void __launch_bounds__(256) mha_small_kernel(const float* __restrict__ qkv, float* __restrict__ out, long n_rows_total, int S, int d,
                                                        int n_heads, float scale, int round_out) {
  constexpr int R = HD / 32;
  constexpr int U = 4;
  const long wid = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= n_rows_total * n_heads) return;
  const int h = static_cast<int>(wid % n_heads);
  const long row = wid / n_heads;  // seq * S + i
  const long seq0 = (row / S) * S;
  const long ld = 3L * d;
  const float* q = qkv + row * ld + h * HD;
  const float* kbase = qkv + seq0 * ld + d + h * HD + lane;
  const float* vbase = qkv + seq0 * ld + 2 * d + h * HD + lane;
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
  float* o = out + row * d + h * HD;
#pragma unroll
  for (int r = 0; r < R; ++r) o[r * 32 + lane] = round_out ? round_tf32(acc[r]) : acc[r];
}

cudaError_t mha_small(const float* qkv, float* out, int n_seq, int S, int d, int n_heads, cudaStream_t st, Launches* lc, int round_out) {
  if (n_seq <= 0) return cudaSuccess;
  if (S > 32 || S < 1) return cudaErrorInvalidValue;
  if (lc) lc->n++;
  const long rows = static_cast<long>(n_seq) * S;
  const int hd = d / n_heads;
  const unsigned grid = static_cast<unsigned>((rows * n_heads + 7) / 8);
  const float scale = 1.f / sqrtf(static_cast<float>(hd));
  if (hd == 64) mha_small_kernel<64><<<grid, 256, 0, st>>>(qkv, out, rows, S, d, n_heads, scale, round_out);
  else if (hd == 256) mha_small_kernel<256><<<grid, 256, 0, st>>>(qkv, out, rows, S, d, n_heads, scale, round_out);
  else return cudaErrorInvalidValue;
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

// ---- objects -> zero-padded [B, 28, 256], each row normalised ---------------------------------------
__global__ void scatter_objects_kernel(const float* __restrict__ emb, const int32_t* __restrict__ cell_ptr, int n_cells, float* __restrict__ X) {
  const long slot = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (slot >= static_cast<long>(n_cells) * kObjectSlots) return;
  const int lane = threadIdx.x & 31;
  const int b = static_cast<int>(slot / kObjectSlots), s = static_cast<int>(slot % kObjectSlots);
  const int n = cell_ptr[b + 1] - cell_ptr[b];
  float4* dst = reinterpret_cast<float4*>(X + slot * kEmbed);
  if (s >= n) {  // padding slot (objects beyond 28 are dropped by construction of the loop bound)
    dst[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    dst[lane + 32] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float4* src = reinterpret_cast<const float4*>(emb + static_cast<long>(cell_ptr[b] + s) * kEmbed);
  const float4 a = src[lane], c = src[lane + 32];
  float ss = (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w) + (c.x * c.x + c.y * c.y) + (c.z * c.z + c.w * c.w);
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  dst[lane] = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
  dst[lane + 32] = make_float4(c.x * inv, c.y * inv, c.z * inv, c.w * inv);
}

cudaError_t scatter_objects(const float* emb, const int32_t* cell_ptr_dev, int n_cells, float* X, cudaStream_t st, Launches* lc) {
  if (n_cells <= 0) return cudaSuccess;
  if (lc) lc->n++;
  const long slots = static_cast<long>(n_cells) * kObjectSlots;
  scatter_objects_kernel<<<static_cast<unsigned>((slots + 7) / 8), 256, 0, st>>>(emb, cell_ptr_dev, n_cells, X);
  return cudaGetLastError();
}

// ---- (count - mean) / std  (models/object_encoder.py:44-45,141-143; fp32 like the reference) ---------
__global__ void num_feature_kernel(const float* __restrict__ meta, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float mean = static_cast<float>(1826.6844940968194), std_ = static_cast<float>(2516.8905096993817);
  out[i] = __fdiv_rn(__fsub_rn(meta[static_cast<long>(i) * 7 + 6], mean), std_);
}

cudaError_t num_feature(const float* meta, int n_obj, float* out, cudaStream_t st, Launches* lc) {
  if (n_obj <= 0) return cudaSuccess;
  if (lc) lc->n++;
  num_feature_kernel<<<(n_obj + 255) / 256, 256, 0, st>>>(meta, n_obj, out);
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
