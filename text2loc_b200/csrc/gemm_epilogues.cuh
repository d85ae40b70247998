// Epilogue policies for umma_gemm_kernel (see umma_gemm.cuh for the interface).
#pragma once

#include "common.cuh"

namespace t2l {

enum : int { kActNone = 0, kActRelu = 1 };

// C[row, col] = act(acc + bias[col]) (+ residual[row, col]);  optional tf32 rounding of the
// stored value when the consumer is another tf32 GEMM (round-to-nearest instead of the
// truncation the tensor core would apply to a raw fp32 operand).
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
  };
  static constexpr int kSmemBytes = 0;
  const Params& p;
  int ew, lane;
  __device__ StoreEpi(const Params& p_, uint8_t*, int ew_, int lane_) : p(p_), ew(ew_), lane(lane_) {}
  __device__ void begin_unit(int, int) {}
  __device__ void end_unit(int, int) {}
  __device__ void chunk(int m_tile, int, int col0, float (&v)[32]) {
    const long row = static_cast<long>(m_tile) * 128 + ew * 32 + lane;
    if (row >= p.M || col0 >= p.N) return;
    float* dst = p.C + row * p.ldc + col0;
    const float* res = p.residual ? p.residual + row * p.ldr + col0 : nullptr;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 o;
      float* of = reinterpret_cast<float*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x = v[i + j];
        if (p.bias) x += __ldg(p.bias + col0 + i + j);
        if (p.act == kActRelu) x = fmaxf(x, 0.f);
        of[j] = x;
      }
      if (res) {
        const float4 r = *reinterpret_cast<const float4*>(res + i);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      if (p.round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
      *reinterpret_cast<float4*>(dst + i) = o;
    }
  }
};

// out[g, col] = max over the 32 rows of group g of relu(acc + bias[col]), optionally also
// max'ed with side[g, col].  A group is 32 consecutive rows = exactly one epilogue warp's TMEM
// lane quadrant, so the reduction is a warp reduction.  Post-ReLU values are >= +0, whose
// IEEE bit patterns order like unsigned integers, so redux.sync.max.u32 does the max.
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
  static constexpr int kSmemBytes = 0;
  const Params& p;
  int ew, lane;
  __device__ SegMaxEpi(const Params& p_, uint8_t*, int ew_, int lane_) : p(p_), ew(ew_), lane(lane_) {}
  __device__ void begin_unit(int, int) {}
  __device__ void end_unit(int, int) {}
  __device__ void chunk(int m_tile, int, int col0, float (&v)[32]) {
    const long row0 = static_cast<long>(m_tile) * 128 + ew * 32;
    if (row0 >= p.M || col0 >= p.N) return;  // warp-uniform
    const long g = row0 >> 5;
    float keep = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float x = fmaxf(v[i] + __ldg(p.bias + col0 + i), 0.f);
      const uint32_t m = __reduce_max_sync(0xffffffffu, __float_as_uint(x) & 0x7fffffffu);  // mask: relu may leave -0
      if (lane == i) keep = __uint_as_float(m);
    }
    if (p.side) keep = fmaxf(keep, __ldg(p.side + g * p.lds + col0 + lane));
    if (p.round_out) keep = round_tf32(keep);
    p.out[g * p.ldo + col0 + lane] = keep;
  }
};

}  // namespace t2l
