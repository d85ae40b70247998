// Epilogue policies for umma_gemm_kernel (see umma_gemm.cuh for the interface).
//
// Global loads the epilogue needs (bias, the SegMax side input) are issued in begin_tile(), i.e.
// BEFORE the warp blocks on the accumulator barrier, so their latency hides behind the MMAs of
// the tile instead of sitting on the epilogue's critical path (profiles/r01: the v1 epilogue spent
// ~1.2k cycles per 32-column chunk in long-scoreboard stalls on exactly these loads).
#pragma once

#include <cuda_fp16.h>

#include "common.cuh"

namespace t2l {

constexpr float kHalfMax = 65504.f;  // stores to fp16 saturate instead of producing inf

enum : int { kActNone = 0, kActRelu = 1 };
constexpr int kEpiBiasSmem = 4 * 256 * 4;  // one 256-float bias slice per epilogue warp

// C[row, col] = act(acc + bias[col]) (+ residual[row, col]);  optional tf32 rounding of the
// stored value when the consumer is another tf32 GEMM (round-to-nearest instead of the
// truncation the tensor core applies to a raw fp32 operand).
struct StoreEpi {
  struct Params {
    float* C;
    long ldc;
    const float* bias;      // [N] or nullptr
    const float* residual;  // [M, ldr] or nullptr
    long ldr;
    int M, N;
    int act;
    int round_out;
    int out_half;  // C is __half [M, ldc]: the consumer is an fp16 tensor-core GEMM (same 11-bit significand as tf32)
  };
  static constexpr int kSmemBytes = kEpiBiasSmem;
  const Params& p;
  float* s_bias;
  int ew, lane, block_n;
  int bias_col0 = -1;  // column slice currently staged in s_bias (tiles of one N column share it)
  __device__ StoreEpi(const Params& p_, uint8_t* smem, int ew_, int lane_, int block_n_)
      : p(p_), s_bias(reinterpret_cast<float*>(smem) + ew_ * 256), ew(ew_), lane(lane_), block_n(block_n_) {}
  __device__ void begin_unit(int, int) {}
  __device__ void end_unit(int, int) {}
  __device__ void begin_tile(int, int, int col0) {
    if (col0 == bias_col0) return;
    bias_col0 = col0;
    __syncwarp();
    for (int i = lane; i < block_n; i += 32) s_bias[i] = (p.bias && col0 + i < p.N) ? __ldg(p.bias + col0 + i) : 0.f;
    __syncwarp();
  }
  __device__ void chunk(int m_tile, int, int c, int col0, float (&v)[32]) {
    const long row = static_cast<long>(m_tile) * 128 + ew * 32 + lane;
    if (row >= p.M || col0 >= p.N) return;
    float4 r[8];
    if (p.residual) {
      const float4* res = reinterpret_cast<const float4*>(p.residual + row * p.ldr + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = __ldg(res + i);  // all eight loads in flight before the first use
    }
    const float4* sb = reinterpret_cast<const float4*>(s_bias + c * 32);
    uint2 hv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = sb[i];  // same address in every lane: broadcast
      float4 o = make_float4(v[4 * i] + b.x, v[4 * i + 1] + b.y, v[4 * i + 2] + b.z, v[4 * i + 3] + b.w);
      if (p.act == kActRelu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      if (p.residual) { o.x += r[i].x; o.y += r[i].y; o.z += r[i].z; o.w += r[i].w; }
      if (p.out_half) {  // 8 bytes per step; the lane's 32 columns are 64 contiguous bytes
        o.x = fminf(fmaxf(o.x, -kHalfMax), kHalfMax); o.y = fminf(fmaxf(o.y, -kHalfMax), kHalfMax);
        o.z = fminf(fmaxf(o.z, -kHalfMax), kHalfMax); o.w = fminf(fmaxf(o.w, -kHalfMax), kHalfMax);
        const __half2 h0 = __floats2half2_rn(o.x, o.y), h1 = __floats2half2_rn(o.z, o.w);
        uint2 u;
        u.x = *reinterpret_cast<const uint32_t*>(&h0);
        u.y = *reinterpret_cast<const uint32_t*>(&h1);
        hv[i] = u;
      } else {
        if (p.round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
        reinterpret_cast<float4*>(p.C + row * p.ldc + col0)[i] = o;
      }
    }
    if (p.out_half) {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.C) + row * p.ldc + col0);
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = make_uint4(hv[2 * i].x, hv[2 * i].y, hv[2 * i + 1].x, hv[2 * i + 1].y);
    }
  }
};

// out[g, col] = max over the 32 rows of group g of relu(acc + bias[col]), optionally also
// max'ed with side[g, col].  A group is 32 consecutive rows = exactly one epilogue warp's TMEM
// lane quadrant, so the reduction is a warp reduction: a 5-stage exchange butterfly in which
// every stage halves the columns a lane still holds (16+8+4+2+1 = 31 shuffles for all 32
// columns), after which lane l owns the maximum of column l.  (v1 used 32 redux.sync per chunk,
// which serialised at ~44 cycles each -- profiles/r01.)
// This is PointConv's max aggregation (pointnet2.py:35) and GA's global_max_pool (:48).
struct SegMaxEpi {
  struct Params {
    float* out;
    long ldo;
    const float* bias;
    const float* side;  // [groups, lds] or nullptr
    long lds;
    int M, N;           // M % 32 == 0
    int round_out;
  };
  static constexpr int kSmemBytes = kEpiBiasSmem;
  const Params& p;
  float* s_bias;
  int ew, lane, block_n;
  float side_v[8];  // this lane's side value for each 32-column chunk of the tile
  int bias_col0 = -1;
  __device__ SegMaxEpi(const Params& p_, uint8_t* smem, int ew_, int lane_, int block_n_)
      : p(p_), s_bias(reinterpret_cast<float*>(smem) + ew_ * 256), ew(ew_), lane(lane_), block_n(block_n_) {}
  __device__ void begin_unit(int, int) {}
  __device__ void end_unit(int, int) {}
  __device__ void begin_tile(int m_tile, int, int col0) {
    const long row0 = static_cast<long>(m_tile) * 128 + ew * 32;
    const bool live = row0 < p.M;
#pragma unroll
    for (int c = 0; c < 8; ++c) {  // issued first: consumed only after the accumulator wait
      const int col = col0 + c * 32 + lane;
      side_v[c] = (p.side && live && c * 32 < block_n && col < p.N) ? __ldg(p.side + (row0 >> 5) * p.lds + col) : 0.f;
    }
    if (col0 == bias_col0) return;
    bias_col0 = col0;
    __syncwarp();
    for (int i = lane; i < block_n; i += 32) s_bias[i] = (col0 + i < p.N) ? __ldg(p.bias + col0 + i) : 0.f;
    __syncwarp();
  }
  __device__ void chunk(int m_tile, int, int c, int col0, float (&v)[32]) {
    const long row0 = static_cast<long>(m_tile) * 128 + ew * 32;
    if (row0 >= p.M || col0 >= p.N) return;  // warp-uniform
    const long g = row0 >> 5;
    float x[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) reinterpret_cast<float4*>(x)[i] = reinterpret_cast<const float4*>(s_bias + c * 32)[i];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = fmaxf(v[i] + x[i], 0.f);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool upper = (lane & off) != 0;  // this lane keeps the upper half of the surviving columns
#pragma unroll
      for (int i = 0; i < off; ++i) {
        const float send = upper ? x[i] : x[i + off];
        const float mine = upper ? x[i + off] : x[i];
        x[i] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, off));
      }
    }
    float keep = fmaxf(x[0], side_v[c]);  // side values are post-ReLU (>= 0); 0 when absent
    if (p.round_out) keep = round_tf32(keep);
    p.out[g * p.ldo + col0 + lane] = keep;
  }
};

}  // namespace t2l
