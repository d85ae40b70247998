// Fused PointConv layer (models/pointcloud/pointnet2.py:31-37): gather -> edge MLP -> max, one kernel.
//
//   out[i] = max over the <=32 ball-query neighbours j of  relu(W2 . relu(Px[j] + W1p.(pos_j - pos_i) + b1) + b2)
//            (then max'ed with the re-added self-loop edge, computed separately and passed as `side`)
//
// The v1 pipeline wrote every first-layer edge activation to HBM (edge_gather -> H, 1 MB per
// object per layer) and read it back in the GEMM.  Here the A operand of the tcgen05 GEMM never
// leaves the SM: eight "gather" warps build each 128-row x 32-channel k-block of edge
// activations directly in shared memory, in the 128-byte-swizzled K-major layout the UMMA
// descriptor expects, while the B operand (W2) arrives by TMA.  A tile is 128 edge rows = 4
// centroids x 32 neighbour slots, so each epilogue warp's TMEM lane quadrant is exactly one
// centroid and the max aggregation is the SegMax epilogue's warp butterfly.
//
// Warp roles (512 threads): 0 = TMA producer (W2 k-blocks), 1 = MMA issuer, 2 = TMEM allocator,
// 4..7 = epilogue, 8..15 = two gather groups of 4 warps.  Ring slot u = tile_iter * KB + kb is filled
// by group u & 1; the ring depth is even, so every stage has exactly one producer group whose
// progress is monotonic (an mbarrier parity wait is only valid for a waiter at most one phase
// behind) and the two groups stream concurrently.  The per-edge index chasing (cnt -> nbr -> src ->
// pos) is done once by edge_records_kernel into a 16-byte record per edge row, so the gather
// warps only follow one dependent load (record -> Px row).  full[stage] counts 1 (TMA expect_tx,
// when W2 streams) + 4 (one arrive per gather warp, after its lanes' proxy fences).
#include "ops.h"
#include "umma_gemm.cuh"
#include "gemm_epilogues.cuh"

namespace t2l {

struct SaFusedParams {
  const float* Px;
  const float4* rec;  // [n_tiles * 128]: (Px row offset in float4 units as int bits, dx, dy, dz) per edge row
  const float* Wp; const float* b1;
  int n_tiles;
};

// One thread per edge row (centroid cen, slot s): which dense point feeds it and pos_j - pos_i.
__global__ void __launch_bounds__(256) edge_records_kernel(const SaFused a, float4* __restrict__ rec, long n_rows) {
  const long row = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const long cen = row >> 5;
  const int s = static_cast<int>(row & 31);
  const int o = static_cast<int>(cen / a.M), m = static_cast<int>(cen % a.M);
  long src;
  if (s < a.cnt[cen]) src = static_cast<long>(o) * a.P + a.nbr[cen * kMaxNbr + s];
  else src = static_cast<long>(a.loop_src_obj[o]) * a.P + a.loop_half[o] * a.M + m;  // empty slots replicate the self-loop edge
  const float* dp = a.dense_pos + src * a.dense_stride;
  rec[row] = make_float4(__int_as_float(static_cast<int>(src * (a.C1 / 4))), dp[0] - a.cpos[cen * 3 + 0], dp[1] - a.cpos[cen * 3 + 1],
                         dp[2] - a.cpos[cen * 3 + 2]);
}

template <int C1, int BLOCK_N>
struct SaCfg {
  static constexpr int KB = C1 / 32;            // k-blocks per tile
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = BLOCK_N * 128;
  // W2 (C2 x C1 fp32) stays resident in shared memory when it fits (SA1: 8 KB, SA2: 64 KB): the ring then
  // carries only A stages and can be deep, which is what hides the slot round trip (fill -> MMA ->
  // commit -> refill, ~4 us measured) -- profiles/r01/sa_fused_bisect.txt.  SA3's 256 KB W2 streams by TMA.
  static constexpr bool B_RESIDENT = (C1 * BLOCK_N * 4) <= 64 * 1024;
  static constexpr int B_RES_BYTES = B_RESIDENT ? KB * B_BYTES : 0;
  static constexpr int STAGE_BYTES = B_RESIDENT ? A_BYTES : A_BYTES + B_BYTES;
  static constexpr int STAGES = B_RESIDENT ? 8 : 4;
  static constexpr int TMEM_COLS = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;
  static constexpr int BAR_BYTES = 256;
  static constexpr int TABLE_BYTES = C1 * 4 * 4;  // (w1p.x, w1p.y, w1p.z, b1) per first-layer channel
  static constexpr int REC_BYTES = 2 * 128 * 16;  // per gather group: (Px row offset, dx, dy, dz) of the tile's 128 edge rows
  static constexpr uint32_t IDESC = umma_idesc(2u, 128, BLOCK_N);
  static constexpr int SMEM = 1024 + B_RES_BYTES + STAGES * STAGE_BYTES + BAR_BYTES + TABLE_BYTES + REC_BYTES + SegMaxEpi::kSmemBytes;
};

constexpr int kSaThreads = 512;

template <int C1, int BLOCK_N>
__global__ void __launch_bounds__(kSaThreads, 1)
sa_fused_kernel(const __grid_constant__ CUtensorMap tm_b, const SaFusedParams p, const SegMaxEpi::Params ep) {
  using Cfg = SaCfg<C1, BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms; offset arithmetic on the __shared__ array keeps the
  // pointer in the shared address space (integer round trips degrade every access to generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* b_res = smem;                          // resident W2 k-block tiles (B_RESIDENT)
  uint8_t* stage_base = smem + Cfg::B_RES_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_base + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* bres_bar = tmem_empty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bres_bar + 1);
  float4* table = reinterpret_cast<float4*>(stage_base + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES);
  float4* records = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(table) + Cfg::TABLE_BYTES);
  uint8_t* epi_smem = reinterpret_cast<uint8_t*>(records) + Cfg::REC_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_b);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], (Cfg::B_RESIDENT ? 0 : 1) + 4);  // one arrive per gather warp of the owning group
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
  for (int c = threadIdx.x; c < C1; c += kSaThreads)
    table[c] = make_float4(p.Wp[c * 4 + 0], p.Wp[c * 4 + 1], p.Wp[c * 4 + 2], p.b1[c]);
  static_assert(Cfg::STAGES % 2 == 0, "slot ownership by parity needs an even ring depth");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================= TMA producer: W2 k-blocks =================
    if (Cfg::B_RESIDENT) {  // all of W2 once, then this warp is done
      if (elect_one()) {
        mbar_arrive_expect_tx(bres_bar, Cfg::B_RES_BYTES);
        for (int kb = 0; kb < Cfg::KB; ++kb) tma_load_2d(&tm_b, bres_bar, b_res + kb * Cfg::B_BYTES, kb * 32, 0, kEvictLast);
      }
      __syncwarp();
    } else {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < Cfg::KB; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::B_BYTES);
            tma_load_2d(&tm_b, &full_bar[stage], stage_base + stage * Cfg::STAGE_BYTES + Cfg::A_BYTES, kb * 32, 0, kEvictLast);
          }
          __syncwarp();
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    if (Cfg::B_RESIDENT) mbar_wait(bres_bar, 0);
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_addr = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < Cfg::KB; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {  // see umma_gemm.cuh: keeps UTCHMMA/UTCBAR straight-line
          const uint8_t* sa = stage_base + stage * Cfg::STAGE_BYTES;
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(Cfg::B_RESIDENT ? b_res + kb * Cfg::B_BYTES : sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32(d_addr, adesc + 2 * k, bdesc + 2 * k, Cfg::IDESC, (kb | k) != 0);
          tc_commit(&empty_bar[stage]);
          if (kb == Cfg::KB - 1) tc_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4 && warp < 8) {
    // ================= epilogue: bias + ReLU + max over each centroid's 32 rows =================
    const int ew = warp - 4;
    SegMaxEpi epi(ep, epi_smem, ew, lane, BLOCK_N);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      epi.begin_tile(tile, 0, 0);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BLOCK_N;
      float v[2][32];
      tmem_ld_32x32(t_addr, v[0]);
#pragma unroll
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        tmem_ld_wait(v[c & 1]);
        if (c + 1 < BLOCK_N / 32) tmem_ld_32x32(t_addr + (c + 1) * 32, v[(c + 1) & 1]);
        epi.chunk(tile, 0, c, c * 32, v[c & 1]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 8) {
    // ================= gather producers: build the A operand in shared memory =================
    // Coalesced mapping: 8 lanes cover one 128-byte row slice (a warp instruction = 4 full lines);
    // thread t owns 16-byte chunk `sub` of rows rb, rb+16, ..., rb+112 of the tile.
    const int group = (warp - 8) >> 2;
    const int t = threadIdx.x & 127;
    const int sub = t & 7, rb = t >> 3;
    float4* rec_s = records + group * 128;
    const float4* __restrict__ px4 = reinterpret_cast<const float4*>(p.Px);
    const bool every_tile = (Cfg::KB % 2) == 0;  // otherwise the groups alternate tiles
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      if (!every_tile && (it & 1) != group) continue;
      const float4 my_rec = __ldg(p.rec + static_cast<long>(tile) * 128 + t);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");  // the group is done reading the previous tile's records
      rec_s[t] = my_rec;
      asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
      int off[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) off[i] = __float_as_int(rec_s[rb + 16 * i].x) + sub;
      const long u0 = static_cast<long>(it) * Cfg::KB;
      int kb = static_cast<int>((u0 & 1) != group);  // first k-block of this tile whose ring slot is ours
      float4 cur[8];
      if (kb < Cfg::KB) {
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = __ldg(px4 + off[i] + kb * 8);
      }
#pragma unroll 1
      for (; kb < Cfg::KB; kb += 2) {
        const long u = u0 + kb;
        const int stage = static_cast<int>(u % Cfg::STAGES);
        const uint32_t phase = static_cast<uint32_t>((u / Cfg::STAGES) & 1);
        float4 nxt[8];
        if (kb + 2 < Cfg::KB) {
#pragma unroll
          for (int i = 0; i < 8; ++i) nxt[i] = __ldg(px4 + off[i] + (kb + 2) * 8);  // our next k-block in flight during this one
        }
        float4 w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = table[kb * 32 + sub * 4 + q];
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* abase = stage_base + stage * Cfg::STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = rb + 16 * i;
          const float4 d = rec_s[r];
          const float in4[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
          float o4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v = in4[q] + w[q].w;
            v = fmaf(w[q].x, d.y, v);
            v = fmaf(w[q].y, d.z, v);
            v = fmaf(w[q].z, d.w, v);
            o4[q] = round_tf32(fmaxf(v, 0.f));
          }
          // 128B swizzle: 16-byte chunk `sub` of row r lives at chunk position sub ^ (r % 8)
          *reinterpret_cast<float4*>(abase + r * 128 + ((sub ^ (r & 7)) << 4)) = make_float4(o4[0], o4[1], o4[2], o4[3]);
        }
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core's async proxy
        __syncwarp();
        // 128 threads arriving on one mbarrier word serialise (~0.5 us per k-block measured): one per warp
        if ((threadIdx.x & 31) == 0) mbar_arrive(&full_bar[stage]);
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int C1, int BLOCK_N>
static cudaError_t launch_sa(const SaFused& a, cudaStream_t st) {
  using Cfg = SaCfg<C1, BLOCK_N>;
  CUtensorMap tb;
  if (make_operand_map(&tb, a.W2, false, BLOCK_N, C1, a.ldw2, BLOCK_N)) return cudaErrorInvalidValue;
  static bool configured_dev[64] = {};  // the attribute is per device: one flag per device ordinal
  bool& configured = configured_dev[current_device() & 63];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(sa_fused_kernel<C1, BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const long n_rows = static_cast<long>(a.n_obj) * a.M * 32;
  if (n_rows <= 0) return cudaSuccess;
  edge_records_kernel<<<static_cast<unsigned>((n_rows + 255) / 256), 256, 0, st>>>(a, a.rec, n_rows);
  SaFusedParams p{a.Px, a.rec, a.Wp, a.b1, a.n_obj * a.M / 4};
  SegMaxEpi::Params ep{a.out, BLOCK_N, a.b2, a.side, BLOCK_N, a.n_obj * a.M * 32, BLOCK_N, 1};
  if (p.n_tiles <= 0) return cudaSuccess;
  const int grid = p.n_tiles < tma_api().num_sms ? p.n_tiles : tma_api().num_sms;
  sa_fused_kernel<C1, BLOCK_N><<<grid, kSaThreads, Cfg::SMEM, st>>>(tb, p, ep);
  return cudaGetLastError();
}

cudaError_t sa_fused(const SaFused& a, cudaStream_t st, Launches* lc) {
  if (a.n_obj <= 0) return cudaSuccess;
  if (lc) lc->n += 2;
  if (a.C1 == 32 && a.C2 == 64) return launch_sa<32, 64>(a, st);
  if (a.C1 == 128 && a.C2 == 128) return launch_sa<128, 128>(a, st);
  if (a.C1 == 256 && a.C2 == 256) return launch_sa<256, 256>(a, st);
  return cudaErrorInvalidValue;
}

}  // namespace t2l
