// Persistent, warp-specialised tcgen05 GEMM skeleton for sm_100a.
//
//   D[M,N] (fp32, TMEM) = A[M,K] * B[N,K]^T      both operands K-major (row-major, K contiguous)
//
// One CTA per SM loops over work units; a unit is one 128-row M tile and a contiguous range
// of N tiles.  Roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread),
// warp 2 = TMEM allocator, warps 4..7 = epilogue (one TMEM lane quadrant each).  Three
// pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue, double-buffered
// accumulator so the epilogue of tile t overlaps the MMAs of tile t+1), and the static
// unit schedule.  Operand tiles are 128-byte rows (32 tf32 / 64 bf16 per k-block) with the
// 128B TMA swizzle, which is what the UMMA shared-memory descriptor in common.cuh describes.
//
// The K loop is a list of "passes": pass p reads A columns a_off[p].. and B columns b_off[p]..
// for kb_per_pass k-blocks.  One pass is an ordinary GEMM; three passes over [hi|lo] operand
// planes give the error-compensated split product hi*hi + hi*lo + lo*hi in one accumulator.
//
// The epilogue is a policy class (see gemm_epilogues.cuh, search.cu):
//   struct Epi { struct Params; static constexpr int kSmemBytes; static constexpr bool kCompactLoop;  // (see the drain loop)
//     static constexpr int kSets;   // epilogue warp sets (1, or 2 with kCompactLoop)
//     __device__ Epi(const Params&, uint8_t* smem, int epi_warp, int lane, int block_n, int set);
//     __device__ void prefetch_unit(int m_tile, int col0);   // the unit this CTA will process AFTER the current one (L2 prefetch hook)
//     __device__ void begin_unit(int m_tile, int split);
//     __device__ void begin_tile(int m_tile, int n_tile, int col0);   // BEFORE the accumulator is waited for: issue global loads here
//     __device__ void chunk(int m_tile, int n_tile, int c, int col0, float (&v)[32]);   // 32 fp32 columns of this thread's row
//     __device__ void end_unit(int m_tile, int split); };
#pragma once

#include <cuda.h>

#include <type_traits>

#include "common.cuh"

namespace t2l {

struct KSchedule {
  int n_pass;       // 1..3
  int kb_per_pass;  // k-blocks per pass
  int a_off[3];     // element offset along K of A for each pass
  int b_off[3];     // element offset along K of B for each pass
};

struct GemmShape {
  int M, N;
  int m_tiles, n_tiles;
  int n_splits;         // units per M tile
  int tiles_per_split;  // N tiles per unit
  KSchedule ks;
  const int* m_rows_dev = nullptr;  // optional device scalar: only the first *m_rows_dev rows of A are live (count produced by an earlier kernel)
};

// kCtaGroup = 2: a cluster of two CTAs computes a 256 x BLOCK_N tile with one cta_group::2 MMA stream
// issued by the leader CTA.  CTA r holds A rows [128 r, 128 r + 128) and B rows [BLOCK_N/2 r, ...) of the
// tile: 32 KB instead of 48 KB per k-block at BLOCK_N = 256, so the ring is 6 deep instead of 4 and the
// per-SM shared-memory traffic (TMA fill + MMA operand reads), which caps the 1-CTA tf32 kernel at ~68 %
// of the tensor peak, drops by a third.
enum : int { kOpTf32 = 0, kOpBf16 = 1, kOpF16 = 2 };  // operand element type (accumulation is always fp32)

template <int kBlockN, int kType, int kCtaGroup = 1>
struct GemmCfg {
  static constexpr int CTA_GROUP = kCtaGroup;
  static constexpr int BLOCK_M = 128;                 // rows per CTA
  static constexpr int BLOCK_N = kBlockN;
  static constexpr int LOAD_N = kBlockN / kCtaGroup;  // B rows each CTA loads
  static constexpr int ELEM_BYTES = kType == kOpTf32 ? 4 : 2;
  static constexpr int BLOCK_K = 128 / ELEM_BYTES;  // one 128-byte swizzle atom per row
  static constexpr int UMMA_K = 32 / ELEM_BYTES;    // 8 (tf32) / 16 (bf16)
  static constexpr int A_BYTES = BLOCK_M * 128;
  static constexpr int B_BYTES = LOAD_N * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (STAGE_BYTES > 32768) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;  // two accumulator buffers
  static constexpr int BAR_BYTES = 256;
  // instruction-descriptor operand format: kind::tf32 -> 2; kind::f16 -> 0 (f16) / 1 (bf16)
  static constexpr uint32_t IDESC = umma_idesc(kType == kOpTf32 ? 2u : (kType == kOpBf16 ? 1u : 0u), BLOCK_M * kCtaGroup, BLOCK_N);
  static constexpr bool IS_HALF = kType != kOpTf32;  // 2-byte operands: tcgen05.mma kind::f16
  static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns must be a power of two >= 32");
};

constexpr int kGemmThreads = 256;

// Optional parts of the epilogue interface, detected by the marker `static constexpr bool kTmaIo`: an epilogue that moves its
// side input and its output with TMA (ResidualTmaEpi, gemm_epilogues.cuh) owns mbarriers in its shared memory
// (init_barriers), shrinks the operand ring to make room for its slabs (kMaxStages), takes the 32-column chunks of a tile
// as one contiguous run per warp set, and is told about the first unit, the absence of a next unit, and the end of the loop.
template <class Epi, class = void>
struct EpiTraits {
  static constexpr bool kTmaIo = false;
  static constexpr int kMaxStages = 64;
};
template <class Epi>
struct EpiTraits<Epi, std::enable_if_t<Epi::kTmaIo>> {
  static constexpr bool kTmaIo = true;
  static constexpr int kMaxStages = Epi::kMaxStages;
};
template <class Cfg, class Epi>
struct GemmLayout {
  static constexpr int STAGES = Cfg::STAGES < EpiTraits<Epi>::kMaxStages ? Cfg::STAGES : EpiTraits<Epi>::kMaxStages;
  static constexpr int SMEM_BYTES = 1024 + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES + Epi::kSmemBytes;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory of the GEMM kernel exceeds 227 KB");
};

// Epi::kSets epilogue warp sets of four warps each (warps 4..7, 8..11): with two sets every SM sub-partition hosts two epilogue
// warps, which interleave the 32-column chunks of a tile between them (an epilogue that is latency-bound in a single warp --
// the top-k list maintenance -- then keeps pace with single-pass K = 256 tiles).
template <class Cfg, class Epi>
__global__ void __launch_bounds__(128 + 128 * Epi::kSets, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const GemmShape shape, const __grid_constant__ typename Epi::Params ep) {
  constexpr int STAGES = GemmLayout<Cfg, Epi>::STAGES;
  constexpr bool kTmaIo = EpiTraits<Epi>::kTmaIo;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms; offset arithmetic on the __shared__ array keeps the
  // pointer in the shared address space (integer round trips degrade every access to generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr bool kPair = Cfg::CTA_GROUP == 2;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs)
  const int cta = kPair ? blockIdx.x >> 1 : blockIdx.x;  // scheduling index (cluster index for pairs)
  const int n_cta = kPair ? gridDim.x >> 1 : gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], Cfg::CTA_GROUP);  // pairs: both producers arrive on the LEADER's barrier
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4 * Cfg::CTA_GROUP * Epi::kSets);  // one arrive per epilogue warp (of both CTAs, on the leader's barrier)
    }
    if constexpr (kTmaIo) Epi::init_barriers(epi_smem);
    fence_barrier_init();
  }
  if (kPair) cluster_sync_all();  // both CTAs resident before the paired TMEM allocation
  if (warp == 2) {
    if (kPair) tmem_alloc_pair(tmem_ptr, Cfg::TMEM_COLS);
    else tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  int n_units = shape.m_tiles * shape.n_splits;  // m_tiles counts 128*CTA_GROUP-row tiles
  if (shape.m_rows_dev) {  // every role reads the same value, so they all skip the same units
    const int live_tiles = (*shape.m_rows_dev + Cfg::BLOCK_M * Cfg::CTA_GROUP - 1) / (Cfg::BLOCK_M * Cfg::CTA_GROUP);
    n_units = min(n_units, live_tiles * shape.n_splits);
  }
  const int kb_total = shape.ks.n_pass * shape.ks.kb_per_pass;

  if (warp == 0) {
    // ================= TMA producer =================
    // the whole warp walks the schedule (converged), one elected lane issues: same reason as the MMA warp
    int stage = 0;
    uint32_t phase = 0;
    for (int u = cta; u < n_units; u += n_cta) {
      const int m_tile = u / shape.n_splits, split = u % shape.n_splits;
      const int nt0 = split * shape.tiles_per_split;
      const int nt1 = min(nt0 + shape.tiles_per_split, shape.n_tiles);
      const int a_row = (m_tile * Cfg::CTA_GROUP + static_cast<int>(rank)) * Cfg::BLOCK_M;
      for (int nt = nt0; nt < nt1; ++nt) {
        const int b_row = nt * Cfg::BLOCK_N + static_cast<int>(rank) * Cfg::LOAD_N;
        for (int p = 0; p < shape.ks.n_pass; ++p) {
          for (int kk = 0; kk < shape.ks.kb_per_pass; ++kk) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
              uint8_t* sa = stage_base + stage * Cfg::STAGE_BYTES;
              const int ka = shape.ks.a_off[p] + kk * Cfg::BLOCK_K, kb = shape.ks.b_off[p] + kk * Cfg::BLOCK_K;
              if (kPair) {
                tma_load_2d_pair(&tm_a, &full_bar[stage], sa, ka, a_row, kEvictNormal);
                tma_load_2d_pair(&tm_b, &full_bar[stage], sa + Cfg::A_BYTES, kb, b_row, kEvictLast);
                if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);  // both CTAs' bytes land on this barrier
                else mbar_arrive_remote(&full_bar[stage], 0);
              } else {
                mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                tma_load_2d(&tm_a, &full_bar[stage], sa, ka, a_row, kEvictNormal);
                tma_load_2d(&tm_b, &full_bar[stage], sa + Cfg::A_BYTES, kb, b_row, kEvictLast);
              }
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ================= MMA issuer (leader CTA only for pairs) =================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = cta; u < n_units; u += n_cta) {
      const int split = u % shape.n_splits;
      const int nt0 = split * shape.tiles_per_split;
      const int nt1 = min(nt0 + shape.tiles_per_split, shape.n_tiles);
      for (int nt = nt0; nt < nt1; ++nt) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + acc * Cfg::BLOCK_N;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          // elect.sync (not `lane == 0`): the compiler then knows a single lane runs the uniform-datapath
          // UTCHMMA/UTCBAR instructions and emits them straight-line; under a plain lane predicate it wraps
          // every one in an ELECT/BRA.U.ANY loop (~400 cycles of issue overhead per k-block, profiles/r01)
          if (elect_one()) {
            const uint8_t* sa = stage_base + stage * Cfg::STAGE_BYTES;
            const uint64_t adesc = umma_desc_sw128(sa);
            const uint64_t bdesc = umma_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
            for (int k = 0; k < Cfg::BLOCK_K / Cfg::UMMA_K; ++k) {
              // advance 32 bytes (one UMMA_K slice) inside the swizzle atom: +2 in 16-byte units
              const uint32_t accum = (kb | k) != 0;
              if (kPair) {
                if (Cfg::IS_HALF) umma_f16_pair(d_addr, adesc + 2 * k, bdesc + 2 * k, Cfg::IDESC, accum);
                else umma_tf32_pair(d_addr, adesc + 2 * k, bdesc + 2 * k, Cfg::IDESC, accum);
              } else {
                if (Cfg::IS_HALF) umma_f16(d_addr, adesc + 2 * k, bdesc + 2 * k, Cfg::IDESC, accum);
                else umma_tf32(d_addr, adesc + 2 * k, bdesc + 2 * k, Cfg::IDESC, accum);
              }
            }
            if (kPair) {  // arrive on the barriers of BOTH CTAs: each has its own producer and epilogue
              tc_commit_pair(&empty_bar[stage]);
              if (kb == kb_total - 1) tc_commit_pair(&tmem_full[acc]);
            } else {
              tc_commit(&empty_bar[stage]);                         // smem slot free once these MMAs retire
              if (kb == kb_total - 1) tc_commit(&tmem_full[acc]);  // accumulator ready
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int ew = (warp - 4) & 3;   // == warp % 4: the TMEM lane quadrant this warp may read
    const int set = (warp - 4) >> 2;  // epilogue warp set: handles chunks set, set + kSets, ...
    Epi epi(ep, epi_smem, ew, lane, Cfg::BLOCK_N, set);
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (kTmaIo) {
      if (cta < n_units)
        epi.first_unit((cta / shape.n_splits) * Cfg::CTA_GROUP + static_cast<int>(rank), (cta % shape.n_splits) * shape.tiles_per_split * Cfg::BLOCK_N);
    }
    for (int u = cta; u < n_units; u += n_cta) {
      const int m_tile = (u / shape.n_splits) * Cfg::CTA_GROUP + static_cast<int>(rank);  // this CTA's 128-row tile
      const int split = u % shape.n_splits;
      const int nt0 = split * shape.tiles_per_split;
      const int nt1 = min(nt0 + shape.tiles_per_split, shape.n_tiles);
      if (u + n_cta < n_units) {
        const int u2 = u + n_cta;
        epi.prefetch_unit((u2 / shape.n_splits) * Cfg::CTA_GROUP + static_cast<int>(rank), (u2 % shape.n_splits) * shape.tiles_per_split * Cfg::BLOCK_N);
      } else if constexpr (kTmaIo) {
        epi.no_next_unit();
      }
      epi.begin_unit(m_tile, split);
      for (int nt = nt0; nt < nt1; ++nt) {
        epi.begin_tile(m_tile, nt, nt * Cfg::BLOCK_N);  // bias / side-input loads fly while the MMAs finish
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * Cfg::BLOCK_N;
        // software-pipelined drain: the TMEM load of chunk c+1 is in flight while chunk c is processed
        float v[2][32];
        if constexpr (!Epi::kCompactLoop && Epi::kSets == 1) tmem_ld_32x32(t_addr, v[0]);
        if constexpr (Epi::kCompactLoop) {
          // two inlined copies of chunk() instead of BLOCK_N / 32: an epilogue whose chunk() is hundreds of instructions long
          // (the top-k list maintenance) would otherwise not fit the instruction cache (measured: 2.3x slower unrolled 8x)
          constexpr int NC = Cfg::BLOCK_N / 32, ST = Epi::kSets;
          static_assert(NC % (2 * ST) == 0, "chunks must split evenly over the epilogue sets, two per iteration");
          tmem_ld_32x32(t_addr + set * 32, v[0]);
#pragma unroll 1
          for (int c = set; c < NC; c += 2 * ST) {
            tmem_ld_wait(v[0]);
            tmem_ld_32x32(t_addr + (c + ST) * 32, v[1]);
            epi.chunk(m_tile, nt, c, nt * Cfg::BLOCK_N + c * 32, v[0]);
            tmem_ld_wait(v[1]);
            if (c + 2 * ST < NC) tmem_ld_32x32(t_addr + (c + 2 * ST) * 32, v[0]);
            epi.chunk(m_tile, nt, c + ST, nt * Cfg::BLOCK_N + (c + ST) * 32, v[1]);
          }
        } else {
          constexpr int NC = Cfg::BLOCK_N / 32, ST = Epi::kSets;
          static_assert(NC % ST == 0, "chunks must split evenly over the epilogue sets");
          if constexpr (kTmaIo) tmem_ld_32x32(t_addr + set * (NC / ST) * 32, v[0]);
          else if constexpr (ST > 1) tmem_ld_32x32(t_addr + set * 32, v[0]);
#pragma unroll
          for (int i = 0; i < NC / ST; ++i) {
            // compile-time for single-set epilogues (their chunk() folds it into addresses); TMA epilogues: one contiguous run per set
            const int c = ST == 1 ? i : (kTmaIo ? set * (NC / ST) + i : set + i * ST);
            tmem_ld_wait(v[i & 1]);
            if (i + 1 < NC / ST) tmem_ld_32x32(t_addr + (c + (kTmaIo ? 1 : ST)) * 32, v[(i + 1) & 1]);
            epi.chunk(m_tile, nt, c, nt * Cfg::BLOCK_N + c * 32, v[i & 1]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kPair && rank != 0) mbar_arrive_remote(&tmem_empty[acc], 0);  // the MMA issuer lives in the leader
          else mbar_arrive(&tmem_empty[acc]);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      epi.end_unit(m_tile, split);
    }
    if constexpr (kTmaIo) epi.finish();
  }

  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();  // pairs: nobody leaves while the peer may still signal its barriers
  if (warp == 2) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------
struct TmaApi {
  typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeTiled encode = nullptr;
  int num_sms = 0;
};
const TmaApi& tma_api();  // resolved once via cudaGetDriverEntryPoint (api.cu)

// 2-D K-major operand [rows, k_elems] with row pitch `ld` elements; box = 128 bytes x box_rows.
// Out-of-bounds elements (row or K tails) are zero-filled by the TMA unit.
inline int make_operand_map(CUtensorMap* tm, const void* base, int type, long rows, long k_elems, long ld, int box_rows) {
  const int esz = type == kOpTf32 ? 4 : 2;
  const CUtensorMapDataType dt = type == kOpTf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : (type == kOpBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(k_elems), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * esz};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15)) return -1;
  CUresult r = tma_api().encode(tm, dt, 2,
                                const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

template <class Cfg, class Epi>
cudaError_t launch_umma_gemm(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const GemmShape& shape,
                             const typename Epi::Params& ep, cudaStream_t stream) {
  constexpr int smem = GemmLayout<Cfg, Epi>::SMEM_BYTES;
  static bool configured_dev[64] = {};  // the attribute is per device: one flag per device ordinal
  bool& configured = configured_dev[current_device() & 63];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(umma_gemm_kernel<Cfg, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int n_units = shape.m_tiles * shape.n_splits;
  if (n_units <= 0) return cudaSuccess;
  const int slots = tma_api().num_sms / Cfg::CTA_GROUP;  // CTAs, or CTA pairs
  const int n_sched = n_units < slots ? n_units : slots;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_sched * Cfg::CTA_GROUP);
  cfg.blockDim = dim3(128 + 128 * Epi::kSets);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = Cfg::CTA_GROUP;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, umma_gemm_kernel<Cfg, Epi>, tm_a, tm_b, shape, ep);
}

}  // namespace t2l
