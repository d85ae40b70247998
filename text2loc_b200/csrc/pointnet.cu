// Feature side of the set-abstraction layers (models/pointcloud/pointnet2.py:31-37, PyG PointConv):
//   message(j -> i) = local_nn([x_j, pos_j - pos_i]),  out_i = max_j message.
// The first Linear of local_nn is split so the 32x-redundant part runs once per POINT:
//   W1 [x_j, dpos] + b = (W1x x_j) + W1p (pos_j - pos_i) + b
// Px = W1x x_j is a dense GEMM over points (linear.cu); this file adds the exact fp32
// position term per EDGE, applies ReLU and lays the edge rows out for the second-layer GEMM
// whose epilogue does the per-centroid max (gemm_epilogues.cuh::SegMaxEpi).
#include <cuda_fp16.h>

#include "ops.h"
#include "common.cuh"

namespace t2l {

__global__ void extract_rgb_kernel(const float* __restrict__ pts, long n_pts, float* __restrict__ x0) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_pts) return;
  const float* p = pts + i * 6;
  *reinterpret_cast<float4*>(x0 + i * 4) = make_float4(p[3], p[4], p[5], 0.f);
}

cudaError_t extract_rgb(const float* pts, int n_obj, float* x0, cudaStream_t st, Launches* lc) {
  const long n = static_cast<long>(n_obj) * kPoints;
  if (n <= 0) return cudaSuccess;
  if (lc) lc->n++;
  extract_rgb_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(pts, n, x0);
  return cudaGetLastError();
}

// SA1's per-point first Linear for the object-resident path: Px16[i, :] = fp16(W1x . rgb_i + b1), 32 channels, K = 3.
// One thread per (point, 8 channels): reads the point's rgb straight from pts, writes 16 bytes.  (The generic SIMT GEMM
// spent 112 us per 4 096 objects on this K = 3 layer; this is a 92 MB streaming pass.)
__global__ void __launch_bounds__(256) px1_kernel(const float* __restrict__ pts, long n_pts, const float* __restrict__ w1x /*[32, ld 4]*/,
                                                  const float* __restrict__ b1, __half* __restrict__ px16) {
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long i = idx >> 2;
  const int q = static_cast<int>(idx & 3);
  if (i >= n_pts) return;
  const float r = pts[i * 6 + 3], g = pts[i * 6 + 4], b = pts[i * 6 + 5];
  uint32_t out[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = q * 8 + p * 2 + e;
      const float4 w = __ldg(reinterpret_cast<const float4*>(w1x + c * 4));
      float acc = fmaf(r, w.x, 0.f);
      acc = fmaf(g, w.y, acc);
      acc = fmaf(b, w.z, acc);
      v[e] = fminf(fmaxf(acc + __ldg(b1 + c), -65504.f), 65504.f);
    }
    const __half2 h = __floats2half2_rn(v[0], v[1]);
    out[p] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(px16 + i * 32 + q * 8) = make_uint4(out[0], out[1], out[2], out[3]);
}

cudaError_t sa1_px16(const float* pts, int n_obj, const float* w1x, const float* b1, __half* px16, cudaStream_t st, Launches* lc) {
  const long n = static_cast<long>(n_obj) * kPoints;
  if (n <= 0) return cudaSuccess;
  if (lc) lc->n++;
  px1_kernel<<<static_cast<unsigned>((n * 4 + 255) / 256), 256, 0, st>>>(pts, n, w1x, b1, px16);
  return cudaGetLastError();
}

// SA1's per-point half of the first Linear for sa_obj2.cu: Qx[i, :] = fp16(W1x . rgb_i + b1 + W1p . (pos_i - o)), o = the
// object's point 0, clamped to +-32752 so that Qx - v stays finite in fp16.  One thread per (point, 8 channels).
__global__ void __launch_bounds__(256) qx1_kernel(const float* __restrict__ pts, long n_pts, const float* __restrict__ w1x /*[32, ld 4]*/,
                                                  const float* __restrict__ w1p /*[32, ld 4]*/, const float* __restrict__ b1,
                                                  __half* __restrict__ qx16) {
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long i = idx >> 2;
  const int q = static_cast<int>(idx & 3);
  if (i >= n_pts) return;
  const float* pi = pts + i * 6;
  const float* po = pts + (i / kPoints) * kPoints * 6;  // the object's point 0
  const float r = pi[3], g = pi[4], b = pi[5];
  const float dx = pi[0] - po[0], dy = pi[1] - po[1], dz = pi[2] - po[2];
  uint32_t out[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = q * 8 + p * 2 + e;
      const float4 w = __ldg(reinterpret_cast<const float4*>(w1x + c * 4));
      const float4 wp = __ldg(reinterpret_cast<const float4*>(w1p + c * 4));
      float acc = fmaf(r, w.x, 0.f);
      acc = fmaf(g, w.y, acc);
      acc = fmaf(b, w.z, acc);
      acc += __ldg(b1 + c);
      acc = fmaf(wp.x, dx, acc);
      acc = fmaf(wp.y, dy, acc);
      acc = fmaf(wp.z, dz, acc);
      v[e] = fminf(fmaxf(acc, -kQxMax), kQxMax);
    }
    const __half2 h = __floats2half2_rn(v[0], v[1]);
    out[p] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(qx16 + i * 32 + q * 8) = make_uint4(out[0], out[1], out[2], out[3]);
}

cudaError_t sa1_qx16(const float* pts, int n_obj, const float* w1x, const float* w1p, const float* b1, __half* qx16, cudaStream_t st, Launches* lc) {
  const long n = static_cast<long>(n_obj) * kPoints;
  if (n <= 0) return cudaSuccess;
  if (lc) lc->n++;
  qx1_kernel<<<static_cast<unsigned>((n * 4 + 255) / 256), 256, 0, st>>>(pts, n, w1x, w1p, b1, qx16);
  return cudaGetLastError();
}

// Position columns of the per-point Linear of levels 2 and 3: row r of x [n*P, ldx] (the previous level's output, C
// feature columns) gets columns C .. C+7 = [hi(d) | lo(d) | 0 0], d = pos_r - o (o = the object's row 0), hi/lo the tf32
// split (hi + lo = d up to 2^-22 |d|), so that the tf32 GEMM against [W1x | W1p | W1p] adds W1p . d at fp32 accuracy.
__global__ void __launch_bounds__(256) pos_cols_kernel(const float* __restrict__ pos /*[n*P, 3]*/, long rows, int P, float* __restrict__ x, int ldx, int C) {
  const long r = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const long r0 = (r / P) * P;
  const float dx = pos[r * 3 + 0] - pos[r0 * 3 + 0], dy = pos[r * 3 + 1] - pos[r0 * 3 + 1], dz = pos[r * 3 + 2] - pos[r0 * 3 + 2];
  const float hx = round_tf32(dx), hy = round_tf32(dy), hz = round_tf32(dz);
  float4* dst = reinterpret_cast<float4*>(x + r * ldx + C);
  dst[0] = make_float4(hx, hy, hz, round_tf32(dx - hx));
  dst[1] = make_float4(round_tf32(dy - hy), round_tf32(dz - hz), 0.f, 0.f);
}

cudaError_t append_pos_cols(const float* pos, int n_obj, int P, float* x, int ldx, int C, cudaStream_t st, Launches* lc) {
  const long rows = static_cast<long>(n_obj) * P;
  if (rows <= 0) return cudaSuccess;
  if ((C & 3) || (ldx & 3) || ldx < C + 8) return cudaErrorInvalidValue;
  if (lc) lc->n++;
  pos_cols_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, st>>>(pos, rows, P, x, ldx, C);
  return cudaGetLastError();
}

// One warp per centroid; lane l owns channels l, l+32, ... (coalesced row reads and writes).
// The 33 edge rows are independent, so they are processed four at a time with all gathers
// (4 x (position + C1/32 feature loads)) issued before the first use: the v1 kernel walked the
// rows one by one and was bound by exposed L2 latency (profiles/r01: 12-16 % of DRAM bandwidth).
template <int C1>
__global__ void __launch_bounds__(256) edge_gather_kernel(EdgeGather a) {
  constexpr int R = C1 / 32;
  constexpr int U = 4;
  const float* __restrict__ Px = a.Px;
  const float* __restrict__ dense_pos = a.dense_pos;
  float* __restrict__ H = a.H;
  float* __restrict__ Hself = a.Hself;
  const int lane = threadIdx.x & 31;
  const long cen = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);  // o * M + m
  if (cen >= static_cast<long>(a.n_obj) * a.M) return;
  const int o = static_cast<int>(cen / a.M), m = static_cast<int>(cen % a.M);
  float wx[R], wy[R], wz[R], b[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = r * 32 + lane;
    wx[r] = a.Wp[c * 4 + 0]; wy[r] = a.Wp[c * 4 + 1]; wz[r] = a.Wp[c * 4 + 2];  // rows padded to 4 floats
    b[r] = a.b1[c];
  }
  const float cx = a.cpos[cen * 3 + 0], cy = a.cpos[cen * 3 + 1], cz = a.cpos[cen * 3 + 2];
  const int cnt = a.cnt[cen];
  const long self_row = static_cast<long>(a.loop_src_obj[o]) * a.P + a.loop_half[o] * a.M + m;
  const int my_nbr = a.nbr[cen * kMaxNbr + lane];
  for (int s0 = 0; s0 <= kMaxNbr; s0 += U) {
    long src[U];
    float dx[U], dy[U], dz[U], px[U][R];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int s = s0 + u;
      const int nb = __shfl_sync(0xffffffffu, my_nbr, s & 31);
      src[u] = (s < cnt) ? static_cast<long>(o) * a.P + nb : self_row;  // empty slots replicate the self-loop edge
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float* dp = dense_pos + src[u] * a.dense_stride;
      dx[u] = dp[0]; dy[u] = dp[1]; dz[u] = dp[2];
#pragma unroll
      for (int r = 0; r < R; ++r) px[u][r] = Px[src[u] * C1 + r * 32 + lane];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int s = s0 + u;
      if (s > kMaxNbr) break;
      const float ex = dx[u] - cx, ey = dy[u] - cy, ez = dz[u] - cz;  // pos_j - pos_i (exact fp32 subtraction)
      float* dst = (s < kMaxNbr) ? H + (cen * kMaxNbr + s) * C1 : Hself + cen * C1;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float v = px[u][r] + b[r];
        v = fmaf(wx[r], ex, v);
        v = fmaf(wy[r], ey, v);
        v = fmaf(wz[r], ez, v);
        dst[r * 32 + lane] = round_tf32(fmaxf(v, 0.f));  // operand of the tf32 second-layer GEMM
      }
    }
  }
}

cudaError_t edge_gather(const EdgeGather& a, cudaStream_t st, Launches* lc) {
  const long n_cen = static_cast<long>(a.n_obj) * a.M;
  if (n_cen <= 0) return cudaSuccess;
  if (lc) lc->n++;
  const unsigned grid = static_cast<unsigned>((n_cen + 7) / 8);
  switch (a.C1) {
    case 32: edge_gather_kernel<32><<<grid, 256, 0, st>>>(a); break;
    case 128: edge_gather_kernel<128><<<grid, 256, 0, st>>>(a); break;
    case 256: edge_gather_kernel<256><<<grid, 256, 0, st>>>(a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// Self-loop edge rows only (one per centroid): A operand of the small side GEMM next to sa_fused.
template <int C1>
__global__ void __launch_bounds__(256) self_edge_kernel(EdgeGather a) {
  constexpr int R = C1 / 32;
  const int lane = threadIdx.x & 31;
  const long cen = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (cen >= static_cast<long>(a.n_obj) * a.M) return;
  const int o = static_cast<int>(cen / a.M), m = static_cast<int>(cen % a.M);
  const long src = static_cast<long>(a.loop_src_obj[o]) * a.P + a.loop_half[o] * a.M + m;
  const float* dp = a.dense_pos + src * a.dense_stride;
  const float ex = dp[0] - a.cpos[cen * 3 + 0], ey = dp[1] - a.cpos[cen * 3 + 1], ez = dp[2] - a.cpos[cen * 3 + 2];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = r * 32 + lane;
    float v = a.Px16 ? __half2float(a.Px16[src * C1 + c]) : a.Px[src * C1 + c] + a.b1[c];  // the fp16 Px already contains b1
    v = fmaf(a.Wp[c * 4 + 0], ex, v);
    v = fmaf(a.Wp[c * 4 + 1], ey, v);
    v = fmaf(a.Wp[c * 4 + 2], ez, v);
    if (a.Hself16) a.Hself16[cen * C1 + c] = __float2half_rn(fminf(fmaxf(v, 0.f), 65504.f));
    else a.Hself[cen * C1 + c] = round_tf32(fmaxf(v, 0.f));
  }
}

// fp16 in / fp16 out variant for the object-resident path: one thread per (centroid, 8 channels) -- 16 bytes of Px16 in,
// 16 bytes of Hself16 out, no idle lanes at C1 = 32 (the warp-per-centroid kernel above took ~85 us per level and
// 4 096 objects for ~100 MB of traffic).  Px16 already contains b1.
// (C1 / 8 and M are powers of two: shifts.)  Every thread handles kSelfUnroll centroids a quarter of the tensor apart with all
// their loads issued before the first use: the kernel is a 3-deep dependent load chain (object -> source row -> data), and with
// one item per thread it ran at ~1 TB/s (13.8 waves of blocks, each waiting out the chain).
constexpr int kSelfUnroll = 4;
__global__ void __launch_bounds__(256) self_edge16_kernel(EdgeGather a, unsigned n_items, unsigned per_thread_stride, int chunk_shift /* log2(C1 / 8) */,
                                                          int m_shift /* log2(M) */) {
  const unsigned base = blockIdx.x * blockDim.x + threadIdx.x;
  if (base >= per_thread_stride) return;
  long cen[kSelfUnroll], src[kSelfUnroll];
  int q[kSelfUnroll];
  bool live[kSelfUnroll];
#pragma unroll
  for (int j = 0; j < kSelfUnroll; ++j) {
    const unsigned idx = base + j * per_thread_stride;
    live[j] = idx < n_items;
    const unsigned id = live[j] ? idx : base;
    cen[j] = id >> chunk_shift;
    q[j] = static_cast<int>(id & ((1u << chunk_shift) - 1u));
    const int o = static_cast<int>(cen[j] >> m_shift), m = static_cast<int>(cen[j] & ((1 << m_shift) - 1));
    src[j] = static_cast<long>(__ldg(a.loop_src_obj + o)) * a.P + __ldg(a.loop_half + o) * a.M + m;
  }
  float ex[kSelfUnroll], ey[kSelfUnroll], ez[kSelfUnroll];
  uint4 raw[kSelfUnroll];
#pragma unroll
  for (int j = 0; j < kSelfUnroll; ++j) {
    const float* dp = a.dense_pos + src[j] * a.dense_stride;
    ex[j] = dp[0] - a.cpos[cen[j] * 3 + 0];
    ey[j] = dp[1] - a.cpos[cen[j] * 3 + 1];
    ez[j] = dp[2] - a.cpos[cen[j] * 3 + 2];
    raw[j] = *reinterpret_cast<const uint4*>(a.Px16 + src[j] * a.C1 + q[j] * 8);
  }
#pragma unroll
  for (int j = 0; j < kSelfUnroll; ++j) {
    if (!live[j]) continue;
    const uint32_t rw[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
    uint32_t out[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const float2 px = __half22float2(*reinterpret_cast<const __half2*>(&rw[p]));
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.Wp) + q[j] * 8 + 2 * p), w1 = __ldg(reinterpret_cast<const float4*>(a.Wp) + q[j] * 8 + 2 * p + 1);
      float v0 = fmaf(w0.x, ex[j], px.x), v1 = fmaf(w1.x, ex[j], px.y);
      v0 = fmaf(w0.y, ey[j], v0); v1 = fmaf(w1.y, ey[j], v1);
      v0 = fmaf(w0.z, ez[j], v0); v1 = fmaf(w1.z, ez[j], v1);
      const __half2 h = __floats2half2_rn(fminf(fmaxf(v0, 0.f), 65504.f), fminf(fmaxf(v1, 0.f), 65504.f));
      out[p] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(a.Hself16 + cen[j] * a.C1 + q[j] * 8) = make_uint4(out[0], out[1], out[2], out[3]);
  }
}

cudaError_t self_edge_rows(const EdgeGather& a, cudaStream_t st, Launches* lc) {
  const long n_cen = static_cast<long>(a.n_obj) * a.M;
  if (n_cen <= 0) return cudaSuccess;
  if (lc) lc->n++;
  if (a.Px16 && a.Hself16) {
    const long n_items = n_cen * (a.C1 / 8);
    auto log2i = [](int v) { int s = 0; while ((1 << s) < v) ++s; return s; };
    if (n_items >= (1L << 31) || (1 << log2i(a.C1 / 8)) != a.C1 / 8 || (1 << log2i(a.M)) != a.M) return cudaErrorInvalidValue;
    const unsigned per_thread = static_cast<unsigned>((n_items + kSelfUnroll - 1) / kSelfUnroll);
    self_edge16_kernel<<<(per_thread + 255) / 256, 256, 0, st>>>(a, static_cast<unsigned>(n_items), per_thread, log2i(a.C1 / 8), log2i(a.M));
    return cudaGetLastError();
  }
  const unsigned grid = static_cast<unsigned>((n_cen + 7) / 8);
  switch (a.C1) {
    case 32: self_edge_kernel<32><<<grid, 256, 0, st>>>(a); break;
    case 128: self_edge_kernel<128><<<grid, 256, 0, st>>>(a); break;
    case 256: self_edge_kernel<256><<<grid, 256, 0, st>>>(a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// GlobalAbstractionLayer input torch.cat((x, pos), dim=1) (pointnet2.py:46), K padded 259 -> 260
__global__ void ga_concat_kernel(const float* __restrict__ x3, const float* __restrict__ cpos3, long rows, float* __restrict__ A) {
  const long r = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* src = reinterpret_cast<const float4*>(x3 + r * 256);
  float4* dst = reinterpret_cast<float4*>(A + r * 260);
  dst[lane] = src[lane];
  dst[lane + 32] = src[lane + 32];
  if (lane == 0) dst[64] = make_float4(cpos3[r * 3 + 0], cpos3[r * 3 + 1], cpos3[r * 3 + 2], 0.f);
}

// fp16 variant: A16[n*32, 264] = [x3 (256) | cpos3 (3) | 0 x 5]
__global__ void ga_concat_half_kernel(const float* __restrict__ x3, const float* __restrict__ cpos3, long rows, __half* __restrict__ A) {
  const long r = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* src = reinterpret_cast<const float4*>(x3 + r * 256);
  const float4 a = src[2 * lane], b = src[2 * lane + 1];  // 8 consecutive channels per lane -> one 16-byte store
  const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w), h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
  uint4 u;
  u.x = *reinterpret_cast<const uint32_t*>(&h0); u.y = *reinterpret_cast<const uint32_t*>(&h1);
  u.z = *reinterpret_cast<const uint32_t*>(&h2); u.w = *reinterpret_cast<const uint32_t*>(&h3);
  reinterpret_cast<uint4*>(A + r * 264)[lane] = u;
  if (lane == 0) {
    const __half2 p0 = __floats2half2_rn(cpos3[r * 3 + 0], cpos3[r * 3 + 1]), p1 = __floats2half2_rn(cpos3[r * 3 + 2], 0.f);
    uint4 t = make_uint4(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1), 0u, 0u);
    reinterpret_cast<uint4*>(A + r * 264)[32] = t;
  }
}

cudaError_t ga_concat_half(const float* x3, const float* cpos3, int n_obj, __half* A, cudaStream_t st, Launches* lc) {
  const long rows = static_cast<long>(n_obj) * 32;
  if (rows <= 0) return cudaSuccess;
  if (lc) lc->n++;
  ga_concat_half_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, st>>>(x3, cpos3, rows, A);
  return cudaGetLastError();
}

cudaError_t ga_concat(const float* x3, const float* cpos3, int n_obj, float* A, cudaStream_t st, Launches* lc) {
  const long rows = static_cast<long>(n_obj) * 32;
  if (rows <= 0) return cudaSuccess;
  if (lc) lc->n++;
  ga_concat_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, st>>>(x3, cpos3, rows, A);
  return cudaGetLastError();
}

}  // namespace t2l
