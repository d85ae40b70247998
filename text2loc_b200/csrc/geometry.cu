// Position-only part of the three set-abstraction layers (models/pointcloud/pointnet2.py:26-30):
// farthest point sampling 256 -> 128 -> 64 -> 32 and the three ball queries (r = 0.2/0.3/0.4,
// <= 32 neighbours).  Index choices are fp32 comparisons, so distances are evaluated exactly as
// the oracle does (sqdist_nofma) and ties resolve to the lowest index (oracle/pyg_ops.py).
#include "ops.h"
#include "common.cuh"

namespace t2l {

// One warp per object.  Level L has P points held in registers, P/32 per lane, point j owned by
// lane j % 32 at register j / 32 (so register order == index order inside a lane).
template <int P, bool kFma>
__device__ __forceinline__ void fps_level(const float* __restrict__ pos_s /*smem [P*3]*/, int lane,
                                          uint8_t* __restrict__ idx_out /*[P/2]*/, float* __restrict__ cpos_s /*smem [(P/2)*3]*/,
                                          float* __restrict__ cpos_out /*[(P/2)*3]*/) {
  constexpr int R = P / 32;
  constexpr int M = P / 2;
  float px[R], py[R], pz[R], dist[R];
#pragma unroll
  for (int t = 0; t < R; ++t) {
    const int j = t * 32 + lane;
    px[t] = pos_s[j * 3 + 0]; py[t] = pos_s[j * 3 + 1]; pz[t] = pos_s[j * 3 + 2];
    dist[t] = __int_as_float(0x7f800000);  // +inf
  }
  int cur = 0;
  for (int s = 0; s < M; ++s) {
    const float cx = pos_s[cur * 3 + 0], cy = pos_s[cur * 3 + 1], cz = pos_s[cur * 3 + 2];
    if (lane == 0) {
      idx_out[s] = static_cast<uint8_t>(cur);
      cpos_s[s * 3 + 0] = cx; cpos_s[s * 3 + 1] = cy; cpos_s[s * 3 + 2] = cz;
      cpos_out[s * 3 + 0] = cx; cpos_out[s * 3 + 1] = cy; cpos_out[s * 3 + 2] = cz;
    }
    if (s == M - 1) break;
    float best = -1.f;
    int best_j = 0;
#pragma unroll
    for (int t = 0; t < R; ++t) {
      dist[t] = fminf(dist[t], sqdist<kFma>(px[t], py[t], pz[t], cx, cy, cz));
      if (dist[t] > best) { best = dist[t]; best_j = t * 32 + lane; }  // strict: first max within the lane
    }
    // distances are >= +0, so their bit patterns order as unsigned integers
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(best));
    const uint32_t cand = (__float_as_uint(best) == wmax) ? static_cast<uint32_t>(best_j) : 0xffffffffu;
    cur = static_cast<int>(__reduce_min_sync(0xffffffffu, cand));  // lowest index among the maxima
  }
  __syncwarp();
}

constexpr int kFpsWarps = 4;

template <bool kFma>
__global__ void __launch_bounds__(kFpsWarps * 32) fps_kernel(const float* __restrict__ pts, int n_obj, Geometry g) {
  __shared__ float s_pos[kFpsWarps][kPoints * 3];
  __shared__ float s_c1[kFpsWarps][128 * 3];
  __shared__ float s_c2[kFpsWarps][64 * 3];
  __shared__ float s_c3[kFpsWarps][32 * 3];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long o = static_cast<long>(blockIdx.x) * kFpsWarps + w;
  if (o >= n_obj) return;
  const float* p = pts + o * kPoints * 6;
  for (int j = lane; j < kPoints; j += 32) {
    s_pos[w][j * 3 + 0] = p[j * 6 + 0]; s_pos[w][j * 3 + 1] = p[j * 6 + 1]; s_pos[w][j * 3 + 2] = p[j * 6 + 2];
  }
  __syncwarp();
  fps_level<256, kFma>(s_pos[w], lane, g.fps1 + o * 128, s_c1[w], g.cpos1 + o * 128 * 3);
  fps_level<128, kFma>(s_c1[w], lane, g.fps2 + o * 64, s_c2[w], g.cpos2 + o * 64 * 3);
  fps_level<64, kFma>(s_c2[w], lane, g.fps3 + o * 32, s_c3[w], g.cpos3 + o * 32 * 3);
}

cudaError_t fps_all_levels(const float* pts, int n_obj, const Geometry& g, bool dist_fma, cudaStream_t st, Launches* lc) {
  if (n_obj <= 0) return cudaSuccess;
  if (lc) lc->n++;
  if (dist_fma) fps_kernel<true><<<(n_obj + kFpsWarps - 1) / kFpsWarps, kFpsWarps * 32, 0, st>>>(pts, n_obj, g);
  else fps_kernel<false><<<(n_obj + kFpsWarps - 1) / kFpsWarps, kFpsWarps * 32, 0, st>>>(pts, n_obj, g);
  return cudaGetLastError();
}

// Ball query: one warp per centroid scans its object's dense points in ascending index, 32 at
// a time, and keeps the first 32 with d < r*r (strict).  r*r is the double product rounded to
// fp32, as torch-cluster passes it.
template <int P, int M, bool kFma>
__device__ __forceinline__ void ball_level(const float* __restrict__ dense_s /*smem [P*3]*/, const float* __restrict__ cpos /*gmem [M*3]*/,
                                           float r2, int warp, int n_warps, int lane, uint8_t* __restrict__ nbr /*[M*32]*/,
                                           uint8_t* __restrict__ cnt /*[M]*/) {
  for (int m = warp; m < M; m += n_warps) {
    const float cx = cpos[m * 3 + 0], cy = cpos[m * 3 + 1], cz = cpos[m * 3 + 2];
    int base = 0;
    for (int t = 0; t < P / 32 && base < kMaxNbr; ++t) {
      const int j = t * 32 + lane;
      const float d = sqdist<kFma>(dense_s[j * 3 + 0], dense_s[j * 3 + 1], dense_s[j * 3 + 2], cx, cy, cz);
      const bool in = d < r2;
      const uint32_t ballot = __ballot_sync(0xffffffffu, in);
      const int rank = base + __popc(ballot & ((1u << lane) - 1u));
      if (in && rank < kMaxNbr) nbr[m * kMaxNbr + rank] = static_cast<uint8_t>(j);
      base += __popc(ballot);
    }
    if (lane == 0) cnt[m] = static_cast<uint8_t>(min(base, kMaxNbr));
  }
}

template <bool kFma>
__global__ void __launch_bounds__(256) ball_kernel(const float* __restrict__ pts, int n_obj, Geometry g, float r2_1, float r2_2, float r2_3) {
  __shared__ float s_d[kPoints * 3];
  const long o = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* p = pts + o * kPoints * 6;
  for (int j = threadIdx.x; j < kPoints; j += blockDim.x) {
    s_d[j * 3 + 0] = p[j * 6 + 0]; s_d[j * 3 + 1] = p[j * 6 + 1]; s_d[j * 3 + 2] = p[j * 6 + 2];
  }
  __syncthreads();
  ball_level<256, 128, kFma>(s_d, g.cpos1 + o * 128 * 3, r2_1, warp, 8, lane, g.nbr1 + o * 128 * 32, g.cnt1 + o * 128);
  __syncthreads();
  for (int j = threadIdx.x; j < 128 * 3; j += blockDim.x) s_d[j] = g.cpos1[o * 128 * 3 + j];
  __syncthreads();
  ball_level<128, 64, kFma>(s_d, g.cpos2 + o * 64 * 3, r2_2, warp, 8, lane, g.nbr2 + o * 64 * 32, g.cnt2 + o * 64);
  __syncthreads();
  for (int j = threadIdx.x; j < 64 * 3; j += blockDim.x) s_d[j] = g.cpos2[o * 64 * 3 + j];
  __syncthreads();
  ball_level<64, 32, kFma>(s_d, g.cpos3 + o * 32 * 3, r2_3, warp, 8, lane, g.nbr3 + o * 32 * 32, g.cnt3 + o * 32);
}

cudaError_t ball_query_all_levels(const float* pts, int n_obj, const Geometry& g, bool dist_fma, cudaStream_t st, Launches* lc) {
  if (n_obj <= 0) return cudaSuccess;
  if (lc) lc->n++;
  // radii of pointnet2.py:57-59; r*r in double then rounded to fp32
  const float r1 = static_cast<float>(0.2 * 0.2), r2 = static_cast<float>(0.3 * 0.3), r3 = static_cast<float>(0.4 * 0.4);
  if (dist_fma) ball_kernel<true><<<n_obj, 256, 0, st>>>(pts, n_obj, g, r1, r2, r3);
  else ball_kernel<false><<<n_obj, 256, 0, st>>>(pts, n_obj, g, r1, r2, r3);
  return cudaGetLastError();
}

}  // namespace t2l
