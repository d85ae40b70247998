// Per-point side of the set-abstraction layers (models/pointcloud/pointnet2.py:31-37, PyG PointConv):
//   message(j -> i) = local_nn([x_j, pos_j - pos_i]),  out_i = max_j message.
// The first Linear of local_nn is split so that its 32x-redundant part runs once per POINT and the rest once per CENTROID:
//   W1 [x_j, pos_j - pos_i] + b = (W1x x_j + b + W1p (pos_j - o)) - W1p (pos_i - o),  o = the object's point 0.
// This file produces the per-point operand Qx (fp16): SA1 straight from pts (K = 3 + 3, exact fp32 SIMT); for SA2 / SA3 it
// appends the tf32 hi | lo position columns to the previous level's output so that ONE tensor-core GEMM against
// [W1x | W1p | W1p] (linear.cu) yields Qx.  The per-centroid part, the second Linear and the max live in sa_obj2.cu.  Also
// the input rows of the global-abstraction MLP.
#include <cuda_fp16.h>

#include "ops.h"
#include "common.cuh"

namespace t2l {

// SA1's per-point half of the first Linear for sa_obj2.cu: Qx[i, :] = fp16(W1x . rgb_i + b1 + W1p . (pos_i - o)), o = the
// object's point 0, clamped to +-32752 so that Qx - v stays finite in fp16.  One thread per point, all 32 channels: the
// weights sit in shared memory and every lane of a warp reads the same address (broadcast).  (One thread per (point, 8
// channels) with the weights through __ldg made each weight load touch four 128-byte lines per warp: 132 us per 1M points
// against a 15 us HBM floor, bound by L1 tag lookups -- profiles/r02.)
__global__ void __launch_bounds__(256) qx1_kernel(const float* __restrict__ pts, long n_pts, const float* __restrict__ w1x /*[32, ld 4]*/,
                                                  const float* __restrict__ w1p /*[32, ld 4]*/, const float* __restrict__ b1,
                                                  __half* __restrict__ qx16) {
  __shared__ __align__(16) float wsm[32 * 8];  // per channel: rgb weights (3) | bias | position weights (3) | 0
  for (int t = threadIdx.x; t < 32 * 8; t += blockDim.x) {
    const int c = t >> 3, k = t & 7;
    wsm[t] = k < 3 ? w1x[c * 4 + k] : (k == 3 ? b1[c] : (k < 7 ? w1p[c * 4 + k - 4] : 0.f));
  }
  __syncthreads();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_pts) return;
  const float2* pi = reinterpret_cast<const float2*>(pts + i * 6);               // 24-byte rows: 8-byte aligned
  const float2* po = reinterpret_cast<const float2*>(pts + (i / kPoints) * kPoints * 6);  // the object's point 0
  const float2 a0 = pi[0], a1 = pi[1], a2 = pi[2], o0 = __ldg(po), o1 = __ldg(po + 1);
  const float dx = a0.x - o0.x, dy = a0.y - o0.y, dz = a1.x - o1.x;
  const float r = a1.y, g = a2.x, b = a2.y;
  uint4* dst = reinterpret_cast<uint4*>(qx16 + i * 32);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t out[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = q * 8 + p * 2 + e;
        const float4 w = *reinterpret_cast<const float4*>(wsm + c * 8), wp = *reinterpret_cast<const float4*>(wsm + c * 8 + 4);
        float acc = fmaf(r, w.x, 0.f);  // same order of operations as before: bit-identical Qx
        acc = fmaf(g, w.y, acc);
        acc = fmaf(b, w.z, acc);
        acc += w.w;
        acc = fmaf(wp.x, dx, acc);
        acc = fmaf(wp.y, dy, acc);
        acc = fmaf(wp.z, dz, acc);
        v[e] = fminf(fmaxf(acc, -kQxMax), kQxMax);
      }
      const __half2 h = __floats2half2_rn(v[0], v[1]);
      out[p] = *reinterpret_cast<const uint32_t*>(&h);
    }
    dst[q] = make_uint4(out[0], out[1], out[2], out[3]);
  }
}

cudaError_t sa1_qx16(const float* pts, int n_obj, const float* w1x, const float* w1p, const float* b1, __half* qx16, cudaStream_t st, Launches* lc) {
  const long n = static_cast<long>(n_obj) * kPoints;
  if (n <= 0) return cudaSuccess;
  if (lc) lc->n++;
  qx1_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(pts, n, w1x, w1p, b1, qx16);
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

// GlobalAbstractionLayer input torch.cat((x, pos), dim=1) (pointnet2.py:46) as fp16 rows: A16[n*32, 264] = [x3 (256) | cpos3 (3) | 0 x 5]
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

}  // namespace t2l
