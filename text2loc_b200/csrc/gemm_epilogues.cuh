// Epilogue policies for umma_gemm_kernel (see umma_gemm.cuh for the interface).
//
// Global loads the epilogue needs (bias, the SegMax side input) are issued in begin_tile(), i.e.
// BEFORE the warp blocks on the accumulator barrier, so their latency hides behind the MMAs of
// the tile instead of sitting on the epilogue's critical path (profiles/r01: the v1 epilogue spent
// ~1.2k cycles per 32-column chunk in long-scoreboard stalls on exactly these loads).
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace t2l {

constexpr float kHalfMax = 65504.f;  // stores to fp16 saturate instead of producing inf

enum : int { kActNone = 0, kActRelu = 1 };
constexpr int kEpiBiasSmem = 4 * 256 * 4;  // one 256-float bias slice per epilogue warp

// C[row, col] = act(acc + bias[col]) (+ residual[row, col]);  optional tf32 rounding of the
// stored value when the consumer is another tf32 GEMM (round-to-nearest instead of the
// truncation the tensor core applies to a raw fp32 operand), or an fp16 store.
//
// A TMEM load hands every lane 32 columns of ITS OWN row, so a direct store makes each warp
// instruction touch 32 different 128-byte lines (32 LSU wavefronts per STG.128; at K = 1024 the
// stores and the residual loads of a tile then cost as many cycles as its MMAs -- profiles/r01,
// out-proj at 577 TFLOP/s).  Each warp therefore transposes its 32 x 32 block through a private
// 4 KB shared-memory tile (XOR-swizzled 16-byte chunks, conflict-free both ways) and goes to global
// memory with full lines: 8 lanes per 128-byte row slice (fp32) or 4 lanes per 64-byte slice (fp16).
// The residual is read in that same coalesced layout and added after the transpose.
// The residual (out-proj, FFN2) is the one operand the epilogue fetches from DRAM.  Fetching it chunk by chunk after the
// accumulator is ready put a full DRAM round trip on every 32-column chunk (~3.4k cycles each, 100 us of a 127 us
// out-proj launch -- profiles/r01).  So: the kernel announces the NEXT unit to prefetch_unit(), which pulls that tile's
// residual rows into L2 (cp.async.bulk.prefetch.L2, one 1 KB row slice per lane) a whole tile ahead; the chunk loop then
// keeps the residual of chunk c+1 in flight in registers while chunk c is transposed and stored.
struct StoreParams {
  float* C;
  long ldc;
  const float* bias;      // [N] or nullptr
  const float* residual;  // [M, ldr] or nullptr
  long ldr;
  int M, N;
  int act;
  int round_out;
  int out_half;  // C is __half [M, ldc]: the consumer is an fp16 tensor-core GEMM (same 11-bit significand as tf32)
  float half_max = kHalfMax;  // fp16 stores saturate at +-half_max
  int res_half = 0;           // residual points at __half data [M, ldr] (the token layer fed with fp16 T5 states)
  int split_out = 0;          // (fp32, no residual) store the tf32 hi | lo planes of the result: hi at column c, lo at column N + c
};

template <bool kHalfOut, bool kResidual>
struct StoreEpiT {
  using Params = StoreParams;
  static constexpr int kStageBytes = 32 * 128;  // one 32 x 32 fp32 block per epilogue warp
  // The residual variants (out-proj, FFN2) are bound by how many residual bytes their epilogue keeps in flight -- one chunk
  // (4 KB) per warp -- not by the MMAs (out-proj: 105 us against a 52 us HBM floor, profiles/r01).  They run TWO epilogue warp
  // sets (even / odd 32-column chunks), which doubles the loads in flight; their bias then comes straight from L1 (__ldg, a
  // broadcast) instead of a per-warp shared-memory slice, which keeps the second set's transpose tiles inside the 227 KB.
  static constexpr int kSets = (kResidual || kHalfOut) ? 2 : 1;  // (the fp16-output variant too: see DESIGN.md section 4, QKV / FFN1)
  static constexpr bool kSmemBias = kSets == 1;
  static constexpr int kSmemBytes = (kSmemBias ? kEpiBiasSmem : 0) + 4 * kSets * kStageBytes;
  static constexpr bool kCompactLoop = kSets == 2;  // two inlined chunk() copies keep the 384-thread variant inside its 168 registers
  const Params& p;
  float* s_bias;
  uint8_t* stage;
  int ew, lane, block_n, set;
  int bias_col0 = -1;  // column slice currently staged in s_bias (tiles of one N column share it)
  float4 res[8];       // residual of the chunk about to be processed (coalesced layout: row 4 j + lane / 8, chunk lane % 8)
  __device__ StoreEpiT(const Params& p_, uint8_t* smem, int ew_, int lane_, int block_n_, int set_)
      : p(p_), s_bias(reinterpret_cast<float*>(smem) + ew_ * 256), stage(smem + (kSmemBias ? kEpiBiasSmem : 0) + (set_ * 4 + ew_) * kStageBytes),
        ew(ew_), lane(lane_), block_n(block_n_), set(set_) {}
  __device__ void begin_unit(int, int) {}
  __device__ void end_unit(int, int) {}
  __device__ __forceinline__ void load_res(long row0, int col0) {
    const int rr = lane >> 3, ch = lane & 7;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long row = row0 + 4 * j + rr;
      if (!(row < p.M && col0 < p.N)) {
        res[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else if (p.res_half) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p.residual) + row * p.ldr + col0) + ch);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        res[j] = make_float4(a.x, a.y, b.x, b.y);
      } else {
        res[j] = __ldg(reinterpret_cast<const float4*>(p.residual + row * p.ldr + col0) + ch);
      }
    }
  }
  __device__ void prefetch_unit(int m_tile, int col0) {
    if (!kResidual || set != 0) return;  // one set pulls the next tile's residual rows into L2 for both
    const long row = static_cast<long>(m_tile) * 128 + ew * 32 + lane;
    if (row < p.M && col0 < p.N) {
      const int esz = p.res_half ? 2 : 4;
      const int bytes = min(block_n, p.N - col0) * esz;
      const char* src = reinterpret_cast<const char*>(p.residual) + (row * p.ldr + col0) * esz;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
    }
  }
  __device__ void begin_tile(int m_tile, int, int col0) {
    if (kResidual) load_res(static_cast<long>(m_tile) * 128 + ew * 32, col0 + set * 32);  // my first chunk, in flight while the MMAs finish
    if (!kSmemBias || col0 == bias_col0) return;
    bias_col0 = col0;
    __syncwarp();
    for (int i = lane; i < block_n; i += 32) s_bias[i] = (p.bias && col0 + i < p.N) ? __ldg(p.bias + col0 + i) : 0.f;
    __syncwarp();
  }
  __device__ void chunk(int m_tile, int, int c, int col0, float (&v)[32]) {
    const long row0 = static_cast<long>(m_tile) * 128 + ew * 32;
    if (row0 >= p.M || col0 >= p.N) return;  // warp-uniform
    const float4* sb = reinterpret_cast<const float4*>(s_bias + c * 32);
    const float4* gb = reinterpret_cast<const float4*>(p.bias + col0);  // (only dereferenced when p.bias != nullptr)
    auto bias4 = [&](int i) -> float4 {  // columns 4 i .. 4 i + 3 of this chunk: the same address in every lane (broadcast)
      if (kSmemBias) return sb[i];
      return p.bias ? __ldg(gb + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    if (kHalfOut && !kResidual) {
      // own row -> 64 bytes = four 16-byte chunks at position i ^ ((row >> 1) & 3)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float o[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 b = bias4(2 * i + h);
          o[4 * h + 0] = v[8 * i + 4 * h + 0] + b.x; o[4 * h + 1] = v[8 * i + 4 * h + 1] + b.y;
          o[4 * h + 2] = v[8 * i + 4 * h + 2] + b.z; o[4 * h + 3] = v[8 * i + 4 * h + 3] + b.w;
        }
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float x = o[2 * q], y = o[2 * q + 1];
          if (p.act == kActRelu) { x = fmaxf(x, 0.f); y = fmaxf(y, 0.f); }
          x = fminf(fmaxf(x, -p.half_max), p.half_max);
          y = fminf(fmaxf(y, -p.half_max), p.half_max);
          const __half2 hh = __floats2half2_rn(x, y);
          w[q] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        *reinterpret_cast<uint4*>(stage + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      __syncwarp();
      __half* Ch = reinterpret_cast<__half*>(p.C);
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // 8 rows x 64 bytes per instruction
        const int r = 8 * j + (lane >> 2), ch = lane & 3;
        const uint4 t = *reinterpret_cast<const uint4*>(stage + r * 64 + ((ch ^ ((r >> 1) & 3)) << 4));
        if (row0 + r < p.M) *reinterpret_cast<uint4*>(Ch + (row0 + r) * p.ldc + col0 + ch * 8) = t;
      }
      __syncwarp();
      return;
    }
    const int rr = lane >> 3, ch = lane & 7;
    float4 cur[8];
    if (kResidual) {
#pragma unroll
      for (int j = 0; j < 8; ++j) cur[j] = res[j];
      if ((c + kSets) * 32 < block_n) load_res(row0, col0 + 32 * kSets);  // my next chunk's residual in flight during this one
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = bias4(i);
      float4 o = make_float4(v[4 * i] + b.x, v[4 * i + 1] + b.y, v[4 * i + 2] + b.z, v[4 * i + 3] + b.w);
      if (p.act == kActRelu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      *reinterpret_cast<float4*>(stage + lane * 128 + ((i ^ (lane & 7)) << 4)) = o;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // 4 rows x 128 bytes per instruction
      const int r = 4 * j + rr;
      float4 o = *reinterpret_cast<const float4*>(stage + r * 128 + ((ch ^ (r & 7)) << 4));
      if (kResidual) { o.x += cur[j].x; o.y += cur[j].y; o.z += cur[j].z; o.w += cur[j].w; }
      if (kHalfOut) {  // fp16 residual stream: acc + bias + residual summed in fp32, ONE rounding, 8 lanes x 8 bytes per row slice
        const __half2 h0 = __floats2half2_rn(fminf(fmaxf(o.x, -p.half_max), p.half_max), fminf(fmaxf(o.y, -p.half_max), p.half_max));
        const __half2 h1 = __floats2half2_rn(fminf(fmaxf(o.z, -p.half_max), p.half_max), fminf(fmaxf(o.w, -p.half_max), p.half_max));
        if (row0 + r < p.M)
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.C) + (row0 + r) * p.ldc + col0 + ch * 4) =
              make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        continue;
      }
      if (!kResidual && p.split_out) {  // the consumer is a three-pass split-product GEMM: emit its operand planes here
        const float4 hi = make_float4(round_tf32(o.x), round_tf32(o.y), round_tf32(o.z), round_tf32(o.w));
        const float4 lo = make_float4(round_tf32(o.x - hi.x), round_tf32(o.y - hi.y), round_tf32(o.z - hi.z), round_tf32(o.w - hi.w));
        if (row0 + r < p.M) {
          *reinterpret_cast<float4*>(p.C + (row0 + r) * p.ldc + col0 + ch * 4) = hi;
          *reinterpret_cast<float4*>(p.C + (row0 + r) * p.ldc + p.N + col0 + ch * 4) = lo;
        }
        continue;
      }
      if (p.round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
      if (row0 + r < p.M) *reinterpret_cast<float4*>(p.C + (row0 + r) * p.ldc + col0 + ch * 4) = o;
    }
    __syncwarp();
  }
};

// C[row, col] = fp16(acc + bias[col] + residual[row, col]) with fp16 residual and output rows, both moved by TMA: the token
// layer's fp16 residual stream (out-projection: x + attn(x); FFN2: x1 + ffn(x1)).
//
// The register-staged residual epilogue above is bound by the residual bytes a warp keeps in flight (one 4 KB chunk; out-proj
// ran at 0.40 of the tensor peak, profiles/r02) and halving the element size halves exactly that.  Here the residual of a
// tile arrives the way the operands do: 128-row x 64-column slabs (16 KB, 128-byte swizzle) fetched by TMA into shared
// memory a whole tile ahead and signalled on an mbarrier.  A thread reads the 64 bytes of ITS row from the slab (the swizzle
// spreads the eight rows of a quarter-warp over all 32 banks), adds accumulator and bias in fp32, rounds once, and writes
// the result back IN PLACE; when the four warps of a set have finished a slab, one thread stores it with TMA (full lines,
// rows beyond M clipped by the hardware) and, once the store has read the slab, refills it with the same slab of the NEXT
// tile.  Two warp sets; set s owns columns [128 s, 128 s + 128) of the tile = two slabs = two buffers.
struct ResidualTmaParams {
  CUtensorMap tm_res;  // residual [M, N] fp16, box 64 columns x 128 rows, SWIZZLE_128B
  CUtensorMap tm_out;  // C [M, N] fp16, same box
  const float* bias;   // [N] or nullptr
  int M, N;
  int act;
  float half_max;
};

struct ResidualTmaEpi {
  using Params = ResidualTmaParams;
  static constexpr bool kTmaIo = true;
  static constexpr int kMaxStages = 5;  // operand ring 6 -> 5 stages of 32 KB: room for the four slabs
  static constexpr int kSets = 2;
  static constexpr bool kCompactLoop = false;
  static constexpr int kSlabBytes = 128 * 128;
  static constexpr int kBarBytes = 64;  // res_full[set][slab]
  static constexpr int kSmemBytes = 1024 + 4 * kSlabBytes;  // barriers + padding to the 1024-byte swizzle period + slabs
  const Params& p;
  uint64_t* res_full;  // this set's two barriers
  uint8_t* slab;       // this set's two slabs
  int ew, lane, set;
  uint32_t tile_seq = 0;  // tiles this CTA has drained: parity of the slab barriers
  int next_m = -1, next_col0 = 0;
  static __device__ void init_barriers(uint8_t* smem) {
    for (int i = 0; i < 4; ++i) mbar_init(reinterpret_cast<uint64_t*>(smem) + i, 1);
  }
  __device__ ResidualTmaEpi(const Params& p_, uint8_t* smem, int ew_, int lane_, int, int set_)
      : p(p_), res_full(reinterpret_cast<uint64_t*>(smem) + 2 * set_), ew(ew_), lane(lane_), set(set_) {
    uint8_t* base = smem + kBarBytes;
    base += (1024u - (smem_u32(base) & 1023u)) & 1023u;
    slab = base + set_ * 2 * kSlabBytes;
  }
  __device__ __forceinline__ bool leader() const { return ew == 0 && lane == 0; }
  __device__ __forceinline__ void load_slab(int j, int m_tile, int col0) {  // leader only
    mbar_arrive_expect_tx(&res_full[j], kSlabBytes);
    tma_load_2d(&p.tm_res, &res_full[j], slab + j * kSlabBytes, col0 + (2 * set + j) * 64, m_tile * 128, kEvictFirst);
  }
  __device__ void first_unit(int m_tile, int col0) {
    if (leader()) {
      tma_prefetch_desc(&p.tm_res);
      tma_prefetch_desc(&p.tm_out);
      load_slab(0, m_tile, col0);
      load_slab(1, m_tile, col0);
    }
    __syncwarp();
  }
  __device__ void prefetch_unit(int m_tile, int col0) { next_m = m_tile; next_col0 = col0; }
  __device__ void no_next_unit() { next_m = -1; }
  __device__ void begin_unit(int, int) {}
  __device__ void end_unit(int, int) {}
  __device__ void begin_tile(int, int, int) {}
  // c runs over this set's four chunks 4 set .. 4 set + 3 in order (EpiTraits::kTmaIo); col0 = first column of chunk c
  __device__ void chunk(int m_tile, int, int c, int col0, float (&v)[32]) {
    const int i = c & 3, j = i >> 1, cc = i & 1;
    uint8_t* buf = slab + j * kSlabBytes;
    if (cc == 0) mbar_wait(&res_full[j], tile_seq & 1);
    const int r = ew * 32 + lane;
    uint8_t* rowp = buf + r * 128;
    const float4* gb = reinterpret_cast<const float4*>(p.bias + col0);  // (only dereferenced when p.bias != nullptr)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint4* slot = reinterpret_cast<uint4*>(rowp + (((cc * 4 + k) ^ (r & 7)) << 4));
      const uint4 rv = *slot;
      const float4 b0 = p.bias ? __ldg(gb + 2 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 b1 = p.bias ? __ldg(gb + 2 * k + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 rr = __half22float2(*reinterpret_cast<const __half2*>(&rw[e]));
        float x = v[8 * k + 2 * e] + bb[2 * e], y = v[8 * k + 2 * e + 1] + bb[2 * e + 1];
        if (p.act == kActRelu) { x = fmaxf(x, 0.f); y = fmaxf(y, 0.f); }
        x = fminf(fmaxf(x + rr.x, -p.half_max), p.half_max);
        y = fminf(fmaxf(y + rr.y, -p.half_max), p.half_max);
        const __half2 hh = __floats2half2_rn(x, y);
        o[e] = *reinterpret_cast<const uint32_t*>(&hh);
      }
      *slot = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (cc == 1) {
      fence_proxy_async();              // my writes to the slab -> visible to the TMA store
      named_bar_sync(1 + set, 128);     // all four warps of the set have finished this slab
      if (leader()) {
        tma_store_2d(&p.tm_out, buf, col0 - 32, m_tile * 128);  // this slab = chunks c - 1, c
        bulk_commit_group();
        if (next_m >= 0) {
          bulk_wait_group_read0();      // the store has read the slab: refill it with the next tile's residual
          load_slab(j, next_m, next_col0);
        }
      }
      __syncwarp();
      if (i == 3) ++tile_seq;
    }
  }
  __device__ void finish() {
    if (leader()) bulk_wait_group0();
    __syncwarp();
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
  // Two epilogue warp sets (even / odd 32-column chunks): the 31-shuffle butterfly per chunk made the single set the limiter
  // of GA's second layer (K = 512: 0.52 ms per 16 384 objects = 1 057 TFLOP/s, profiles/r02).
  static constexpr int kSets = 2;
  static constexpr int kSmemBytes = kSets * kEpiBiasSmem;  // one 256-float bias slice per epilogue warp
  static constexpr bool kCompactLoop = false;
  const Params& p;
  float* s_bias;
  int ew, lane, block_n;
  float side_v[8];  // this lane's side value for each 32-column chunk of the tile
  int bias_col0 = -1;
  __device__ SegMaxEpi(const Params& p_, uint8_t* smem, int ew_, int lane_, int block_n_, int set_)
      : p(p_), s_bias(reinterpret_cast<float*>(smem) + (set_ * 4 + ew_) * 256), ew(ew_), lane(lane_), block_n(block_n_) {}
  __device__ void prefetch_unit(int, int) {}
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
