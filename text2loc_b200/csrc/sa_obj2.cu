// Object-resident fused PointConv layer, second generation (models/pointcloud/pointnet2.py:25-37; PyG PointConv with
// add_self_loops=True), fp16 operands, fp32 accumulation:
//
//   out[i] = max( max over the <=32 ball-query neighbours j of  relu(W2 . relu(Px[j] + W1p.(pos_j - pos_i)) + b2),
//                 the same expression for the re-added self-loop edge (dense point with the centroid's per-cell index) )
//
// The first Linear is applied once per POINT and once per CENTROID instead of once per edge:
//   Px[j] + W1p.(pos_j - pos_i) = Qx[j] - v[i],   Qx[j] = W1x x_j + b1 + W1p.(pos_j - o),   v[i] = W1p.(pos_i - o)
// with o the object's own origin (its point 0 = centroid 0 of every level, so |pos - o| stays of the size of the ball
// radii and nothing cancels).  Qx arrives as fp16 (one rounding of the fp32 sum, in the producing kernel), v is formed here in
// fp32 and rounded to fp16, and an edge element is ONE instruction: fma.rn.relu.f16x2(1, Qx, -v) -- a single rounding of the
// difference, then ReLU -- for two channels.  (Round 1 converted Px to fp32, added three packed FMAs per channel pair and
// converted back: 6 instructions per pair, ~250 per 128 x 64 item and thread; ncu showed the gather warps, not shared memory,
// as the limiter: all three levels took ~1 300 cycles per item, MMAs or not.)  CPU emulation of these rounding points against
// the reference: cell embeddings 9.8e-5 (round 1's points: 9.3e-5), tests/test_oracle.py.
//
// What changed against sa_obj.cu (profiles/r01: shared-memory-bandwidth bound -- an SS-mode 128x128x16 UMMA alone reads
// 128 B/clk of operands, the gather warps' LDS/STS come on top; tensor pipe 9 / 19 / 33 % active):
//   * W2 is the M-side operand of the transposed product D^T[channel, edge] = W2 . A^T and never changes, so it now lives
//     in TENSOR MEMORY (written once per CTA with tcgen05.st, lane = channel, one 32-bit column per pair of K elements) and
//     the MMAs run in TS mode: `tcgen05.mma [d], [a_tmem], b_desc`.  Only the gathered edge tile is read from shared memory:
//     64 B/clk instead of 128, and the 128 KB of shared memory W2 occupied for SA3 become ring slots and a second object
//     buffer.
//   * The self-loop edge of every centroid is folded in: its source rows (a contiguous half of ANOTHER object's Px block)
//     arrive with the object's own block, form one extra tile per object whose columns are centroids instead of edges,
//     and its post-ReLU result waits in a thread-private column of shared memory until the centroid's neighbour tile is
//     reduced.  The separate self-edge gather kernel, the side GEMM and the [n*M, C2] side tensor (96 KB per object written
//     and read back through HBM) are gone.
//   * SA1 (C1 = 32, C2 = 64) filled half of every 128-byte operand row, half of the accumulator lanes and two of the four
//     epilogue warps.  Its tiles are now PAIRED: one item carries the 32 channels of two different edge groups side by side
//     (K = 64) and W2 sits in tensor memory as a block-diagonal [[W2, 0], [0, W2]], so lanes 0..63 of the accumulator
//     are group A's channels and lanes 64..127 group B's: half as many tiles, barriers and fences per object.
//   * For SA3 the two 128-channel halves of a tile go to two accumulators; half 0 consumes every item as it arrives, half 1
//     follows two items behind and releases the ring slot, so its tail overlaps the drain of half 0 and no item waits in the
//     ring for a whole tile (the first form -- half outer, K inner -- held all four items until half 1 had read them).
//
// Warp roles (512 threads): 0 = bulk-copy producer (one block of six copies per object), 1 = MMA issuer, 2 = TMEM
// allocator, 4..7 = W2 upload, then epilogue (one TMEM lane quadrant each), 8..15 = gather (two groups of four warps).
// Pipelines: object buffers full/empty (producer <-> gather), A ring full/empty (gather <-> MMA), accumulators
// full/empty (MMA <-> epilogue).
#include "ops.h"
#include "umma_gemm.cuh"

namespace t2l {

struct SaObj2Params {
  const __half* Qx16;         // [n*P, C1] fp16 (W1x x_j + b1 + W1p (pos_j - o)), |.| <= 32752
  const float* cpos;          // [n*M, 3]  (centroid 0 of an object is its origin o)
  const uint8_t* nbr;         // [n*M, 32]
  const uint8_t* cnt;         // [n*M]
  const int32_t* loop_src;    // [n] object whose dense points feed this object's self loops
  const int32_t* loop_half;   // [n] which half (0 / 1) of that object's dense points
  const float* Wp;            // [C1, 4] position part of the first Linear
  const __half* W2h;          // [C2, C1] fp16
  const float* b2;            // [C2]
  float* out;                 // [n*M, ldo]
  int ldo;                    // row pitch of out (>= C2: the next level appends position columns)
  int n_obj;
  int bisect;                 // timing bisect (t2l_debug_sa_bisect; results are INVALID when != 0): 1 = no accumulator drain, 2 = no MMAs
};

template <int C1_, int C2_, int P_, int M_, bool PAIR_>
struct Sa2Cfg {
  static constexpr int C1 = C1_, C2 = C2_, P = P_, M = M_;
  static constexpr bool PAIR = PAIR_;
  static constexpr int KC = PAIR ? 2 * C1 : C1;       // K extent of a tile's MMAs (halfs): two 32-channel groups side by side when paired
  static constexpr int KB = KC / 64;                   // 128-byte-row items per tile
  static constexpr int MH = C2 > 128 ? C2 / 128 : 1;   // 128-channel halves of W2
  static constexpr int TILE_CEN = PAIR ? 8 : 4;        // centroids per tile (32 edge slots each)
  static constexpr int TPO = M / TILE_CEN;             // neighbour tiles per object (+ 1 self-loop tile)
  static constexpr int SELF_COLS = PAIR ? M / 2 : M;   // live columns of the self-loop tile (one per centroid; two lane groups when paired)
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int PX_BYTES = P * C1 * 2;
  static constexpr int CPOS_BYTES = M * 12;
  static constexpr int NBR_BYTES = M * 32;
  static constexpr int CNT_BYTES = M;
  static constexpr int SPX_BYTES = M * C1 * 2;         // self-loop source rows: half of the source object's Qx block
  static constexpr int SORG_BYTES = 16;                // origin of the source object (its centroid 0)
  // block layout: Qx | self-loop Qx rows | centroids | lists | counts | source origin, padded to a multiple of 128 bytes: the
  // row blocks then start on 128-byte lines in BOTH buffers, so a quarter-warp's 128-byte row read is one shared-memory
  // wavefront (ncu: 28 % of the gather's wavefronts were the second halves of rows straddling two lines)
  static constexpr int OBJ_RAW = PX_BYTES + SPX_BYTES + CPOS_BYTES + NBR_BYTES + CNT_BYTES + SORG_BYTES;
  static constexpr int OBJ_BYTES = (OBJ_RAW + 127) / 128 * 128;
  static constexpr int NOBJ = 2;
  static constexpr int SIDE_BYTES = M * C2 * 4;        // post-ReLU self-loop results, [centroid][channel]
  static constexpr int BAR_BYTES = 256;
  static constexpr int TABLE_BYTES = (C1 < 64 ? 64 : C1) * 12;  // W1p, 768 bytes per 64-channel slice (layout: kernel prologue)
  static constexpr int FIXED = 1024 + NOBJ * OBJ_BYTES + SIDE_BYTES + BAR_BYTES + TABLE_BYTES;
  static constexpr int FIT = (227 * 1024 - FIXED) / A_BYTES;
  static constexpr int STAGES = FIT > 8 ? 8 : FIT;
  static constexpr int W_COLS = MH * (KC / 2);         // tensor-memory columns of W2 (two K elements per 32-bit column)
  static constexpr int NACC = MH == 2 ? 2 : 3;
  static constexpr int ACC_COL0 = 512 - NACC * 128;
  static constexpr uint32_t IDESC = umma_idesc(0u, 128, 128);  // f16 x f16 -> f32, M = 128 channels (lanes), N = 128 edge rows
  static constexpr int SMEM = FIXED + STAGES * A_BYTES;
  static_assert(PX_BYTES % 16 == 0 && CPOS_BYTES % 16 == 0 && NBR_BYTES % 16 == 0 && CNT_BYTES % 16 == 0 && SPX_BYTES % 16 == 0,
                "bulk copies move multiples of 16 bytes");
  static_assert(W_COLS <= ACC_COL0, "W2 and the accumulators share the 512 tensor-memory columns");
  static_assert((2 * STAGES + 2 * NACC + 2 * NOBJ) * 8 + 4 <= BAR_BYTES, "barrier block too small");
  static_assert(STAGES >= 4 || MH == 1, "half 1 of a tile lags two items behind half 0: three items live plus one being built");
  static_assert(STAGES >= 3, "ring too shallow");
  static_assert(PAIR ? (C1 == 32 && C2 == 64) : (C1 % 64 == 0 && C2 % 128 == 0), "shape not covered");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

constexpr int kSa2Threads = 512;

T2L_DEVICE void bulk_load2(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]^T: the M-side operand is read from tensor memory (lane = row, 32-bit column = two K elements)
T2L_DEVICE void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// registers -> TMEM: this warp's 32 lanes x 32 consecutive columns (thread t = lane t of the warp's quadrant)
T2L_DEVICE void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0],"
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16,"
      " %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
T2L_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// relu(q - v) for two channels: one fused multiply-add in fp16 (1 * q + (-v), single rounding), then ReLU
T2L_DEVICE uint32_t s2_sub_relu_h2(uint32_t q, uint32_t neg_v) {
  uint32_t d;
  asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(0x3C003C00u), "r"(q), "r"(neg_v));
  return d;
}
// (lo, hi) fp32 pair -> packed fp16x2, round-to-nearest, saturating at 65504
T2L_DEVICE uint32_t s2_pack_half2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <class Cfg>
__global__ void __launch_bounds__(kSa2Threads, 1) sa_obj2_kernel(const SaObj2Params p) {
  constexpr int C1 = Cfg::C1, C2 = Cfg::C2, P = Cfg::P, M = Cfg::M;
  constexpr bool PAIR = Cfg::PAIR;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;  // A ring, 1024-aligned (128-byte swizzle atoms)
  uint8_t* obj_base = stage_base + Cfg::STAGES * Cfg::A_BYTES;
  float* side_s = reinterpret_cast<float*>(obj_base + Cfg::NOBJ * Cfg::OBJ_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(side_s) + Cfg::SIDE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + Cfg::NACC;
  uint64_t* obj_full = tmem_empty + Cfg::NACC;
  uint64_t* obj_empty = obj_full + Cfg::NOBJ;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(obj_empty + Cfg::NOBJ);
  float* table = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + Cfg::BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 4);  // one arrive per warp of the gather group that built the item
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < Cfg::NACC; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    for (int i = 0; i < Cfg::NOBJ; ++i) {
      mbar_init(&obj_full[i], 1);
      mbar_init(&obj_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, 512);
  // W1p, laid out for the gather's reads: [64-channel slice][x|y|z][channel half 0|1][8-channel group][4 channels].  The eight
  // lanes of a quarter-warp (one 8-channel group each) then read 128 contiguous bytes per load -- with channel-major rows
  // their 32-byte stride made every one of the six loads a 2-way bank conflict (ncu: 13M of the 21M conflict wavefronts).
  for (int c = threadIdx.x; c < C1; c += kSa2Threads) {
    const int slice = c >> 6, grp = (c & 63) >> 3, half = (c & 7) >> 2, e = c & 3;
#pragma unroll
    for (int a = 0; a < 3; ++a) table[((((slice * 3 + a) * 2 + half) * 8 + grp) << 2) + e] = p.Wp[c * 4 + a];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // ---- W2 -> tensor memory (epilogue warps: warp % 4 is the lane quadrant a warp may touch) ----
  if (warp >= 4 && warp < 8) {
    const int ew = warp - 4;
    const int L = ew * 32 + lane;  // TMEM lane = output channel (mod 128)
#pragma unroll 1
    for (int h = 0; h < Cfg::MH; ++h) {
#pragma unroll 1
      for (int q = 0; q < Cfg::KC / 64; ++q) {  // 32 columns = 64 K elements per store
        uint32_t r[32];
        if (PAIR) {
          // block-diagonal [[W2, 0], [0, W2]]: lanes 0..63 multiply K elements 0..31 (edge group A), lanes 64..127 elements 32..63
          const uint4* src = reinterpret_cast<const uint4*>(p.W2h + static_cast<long>(L & 63) * C1);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 w = __ldg(src + i);
            const bool lo = L < 64;
            r[4 * i + 0] = lo ? w.x : 0u; r[4 * i + 1] = lo ? w.y : 0u; r[4 * i + 2] = lo ? w.z : 0u; r[4 * i + 3] = lo ? w.w : 0u;
            r[16 + 4 * i + 0] = lo ? 0u : w.x; r[16 + 4 * i + 1] = lo ? 0u : w.y; r[16 + 4 * i + 2] = lo ? 0u : w.z; r[16 + 4 * i + 3] = lo ? 0u : w.w;
          }
        } else {
          const uint4* src = reinterpret_cast<const uint4*>(p.W2h + static_cast<long>(h * 128 + L) * C1 + q * 64);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 w = __ldg(src + i);
            r[4 * i + 0] = w.x; r[4 * i + 1] = w.y; r[4 * i + 2] = w.z; r[4 * i + 3] = w.w;
          }
        }
        tmem_st_32x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + h * (Cfg::KC / 2) + q * 32, r);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // contiguous run of whole objects per CTA
  const int o0 = static_cast<int>(static_cast<long>(p.n_obj) * blockIdx.x / gridDim.x);
  const int o1 = static_cast<int>(static_cast<long>(p.n_obj) * (blockIdx.x + 1) / gridDim.x);
  constexpr int kTilesPerObj = Cfg::TPO + 1;  // tile 0 of an object = its self-loop edges

  if (warp == 0) {
    // ================= producer: six bulk copies per object onto one barrier =================
    for (int o = o0, n = 0; o < o1; ++o, ++n) {
      const int buf = n % Cfg::NOBJ;
      const int src_obj = __ldg(p.loop_src + o);
      const long src = static_cast<long>(src_obj) * P + __ldg(p.loop_half + o) * M;  // first dense point of the self-loop sources
      mbar_wait(&obj_empty[buf], (((n / Cfg::NOBJ) & 1) ^ 1));
      if (elect_one()) {
        uint8_t* dst = obj_base + buf * Cfg::OBJ_BYTES;
        mbar_arrive_expect_tx(&obj_full[buf], Cfg::OBJ_RAW);
        bulk_load2(dst, p.Qx16 + static_cast<long>(o) * P * C1, Cfg::PX_BYTES, &obj_full[buf]);
        dst += Cfg::PX_BYTES;
        bulk_load2(dst, p.Qx16 + src * C1, Cfg::SPX_BYTES, &obj_full[buf]);
        dst += Cfg::SPX_BYTES;
        bulk_load2(dst, p.cpos + static_cast<long>(o) * M * 3, Cfg::CPOS_BYTES, &obj_full[buf]);
        dst += Cfg::CPOS_BYTES;
        bulk_load2(dst, p.nbr + static_cast<long>(o) * M * 32, Cfg::NBR_BYTES, &obj_full[buf]);
        dst += Cfg::NBR_BYTES;
        bulk_load2(dst, p.cnt + static_cast<long>(o) * M, Cfg::CNT_BYTES, &obj_full[buf]);
        dst += Cfg::CNT_BYTES;
        bulk_load2(dst, p.cpos + static_cast<long>(src_obj) * M * 3, Cfg::SORG_BYTES, &obj_full[buf]);  // centroid 0 of the source object
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // (32-bit item counters: `u % STAGES` on a 64-bit counter was a 25-instruction sequence per item in every gather thread)
    const uint32_t n_tiles = static_cast<uint32_t>(o1 - o0) * kTilesPerObj;
    if constexpr (Cfg::MH == 2) {
      // Two 128-channel halves per tile, one accumulator each (NACC == 2).  Half 0 consumes an item as soon as it is built;
      // half 1 follows kLag items behind and releases the ring slot.  (Round-2 first form: half outer, K inner -- all four items
      // of a tile stayed in the five-slot ring until half 1 had read them, and the gather warps spent a third of their time
      // waiting for slots: ncu source view, 16 % of all samples on that one wait.)  Half 1's tail overlaps the drain of half 0.
      constexpr uint32_t kLag = 2;
      const uint32_t n_items = n_tiles * Cfg::KB;
      auto issue = [&](uint32_t g, int h) {
        const uint32_t tile = g / Cfg::KB, kb = g % Cfg::KB;
        const int stage = static_cast<int>(g % Cfg::STAGES);
        if (kb == 0) {  // first MMA of this tile into accumulator h: the epilogue has drained the previous tile's
          mbar_wait(&tmem_empty[h], (tile & 1) ^ 1);
          tc_fence_after();
        }
        if (h == 0) {
          mbar_wait(&full_bar[stage], (g / Cfg::STAGES) & 1);
          tc_fence_after();
        }
        if (elect_one()) {
          const uint64_t edesc = umma_desc_sw128(stage_base + stage * Cfg::A_BYTES);
          const uint32_t d_addr = tmem_base + Cfg::ACC_COL0 + h * 128;
          const uint32_t w_addr = tmem_base + h * (Cfg::KC / 2) + kb * 32;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (!(p.bisect & 2)) umma_f16_ts(d_addr, w_addr + 8 * k, edesc + 2 * k, Cfg::IDESC, (kb | k) != 0);
          if (h == 1) tc_commit(&empty_bar[stage]);
          if (kb == Cfg::KB - 1) tc_commit(&tmem_full[h]);
        }
        __syncwarp();
      };
      for (uint32_t g = 0; g < n_items + kLag; ++g) {
        if (g >= kLag) issue(g - kLag, 1);
        if (g < n_items) issue(g, 0);
      }
    } else {
      uint32_t u0 = 0;  // first ring item of the current tile
      int acc = 0;
      uint32_t acc_phase = 0;
      for (uint32_t tile = 0; tile < n_tiles; ++tile, u0 += Cfg::KB) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + Cfg::ACC_COL0 + acc * 128;
#pragma unroll 1
        for (int kb = 0; kb < Cfg::KB; ++kb) {
          const uint32_t u = u0 + kb;
          const int stage = static_cast<int>(u % Cfg::STAGES);
          mbar_wait(&full_bar[stage], (u / Cfg::STAGES) & 1);
          tc_fence_after();
          if (elect_one()) {  // see umma_gemm.cuh: keeps UTCHMMA/UTCBAR straight-line
            const uint64_t edesc = umma_desc_sw128(stage_base + stage * Cfg::A_BYTES);  // edge tile: the N-side operand
            const uint32_t w_addr = tmem_base + kb * 32;                                // 64 K elements = 32 columns per item
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (!(p.bisect & 2)) umma_f16_ts(d_addr, w_addr + 8 * k, edesc + 2 * k, Cfg::IDESC, (kb | k) != 0);
            tc_commit(&empty_bar[stage]);
            if (kb == Cfg::KB - 1) tc_commit(&tmem_full[acc]);
          }
          __syncwarp();
        }
        if (++acc == Cfg::NACC) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ================= epilogue =================
    const int ew = warp - 4;  // TMEM lane quadrant
    const int set = PAIR ? (ew >> 1) : 0;                 // paired tiles: lanes 64..127 belong to the second edge group
    const int ch_lo = PAIR ? ((ew & 1) * 32 + lane) : (ew * 32 + lane);  // this thread's channel inside a 128-channel half
    float bias[Cfg::MH];
#pragma unroll
    for (int h = 0; h < Cfg::MH; ++h) bias[h] = __ldg(p.b2 + h * 128 + ch_lo);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int o = o0; o < o1; ++o) {
      float* out_o = p.out + static_cast<long>(o) * M * p.ldo;
#pragma unroll 1
      for (int t = 0; t < kTilesPerObj; ++t) {
#pragma unroll 1
        for (int h = 0; h < Cfg::MH; ++h) {
          const int ch = h * 128 + ch_lo;
          mbar_wait(&tmem_full[acc], acc_phase);
          tc_fence_after();
          const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + Cfg::ACC_COL0 + acc * 128;
          float v[2][32];
          if (p.bisect & 1) {  // timing bisect: release the accumulator without reading it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == Cfg::NACC) { acc = 0; acc_phase ^= 1; }
            continue;
          }
          tmem_ld_32x32(t_addr, v[0]);
          if (t == 0) {
            // self-loop tile: column e = one centroid's self-loop edge; park relu(. + b2) until that centroid's neighbour tile is reduced
            constexpr int kChunks = Cfg::SELF_COLS / 32;
#pragma unroll
            for (int q = 0; q < kChunks; ++q) {
              tmem_ld_wait(v[q & 1]);
              if (q + 1 < kChunks) tmem_ld_32x32(t_addr + (q + 1) * 32, v[(q + 1) & 1]);
              // Paired tiles: lane group `set` reduces centroids 8 m + 4 set + (0..3) in the neighbour tiles, so its self-loop
              // columns must be exactly those centroids -- the parked value is then read back by the thread that wrote it.
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int e = q * 32 + j;
                const int cen = PAIR ? 8 * (e >> 2) + 4 * set + (e & 3) : e;
                side_s[cen * C2 + ch] = fmaxf(v[q & 1][j] + bias[h], 0.f);
              }
            }
          } else {
            const int cen0 = PAIR ? (2 * (t - 1) + set) * 4 : (t - 1) * 4;  // first centroid of this lane group's tile
#pragma unroll
            for (int q = 0; q < 4; ++q) {  // chunk q = centroid cen0 + q: this thread's channel, its 32 edge slots
              tmem_ld_wait(v[q & 1]);
              if (q + 1 < 4) tmem_ld_32x32(t_addr + (q + 1) * 32, v[(q + 1) & 1]);
              const float* x = v[q & 1];
              float m[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) m[i] = fmaxf(fmaxf(x[4 * i], x[4 * i + 1]), fmaxf(x[4 * i + 2], x[4 * i + 3]));
              const float mx = fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
              // bias and ReLU commute with the max over edges (per-channel constant, monotonic rounding)
              const float keep = fmaxf(fmaxf(mx + bias[h], 0.f), side_s[(cen0 + q) * C2 + ch]);
              out_o[(cen0 + q) * p.ldo + ch] = round_tf32(keep);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          if (++acc == Cfg::NACC) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= 8) {
    // ================= gather: A items from the resident object block =================
    // Two groups of four warps.  Where a tile has <= 2 items the groups alternate TILES (one proxy fence per tile),
    // else they alternate items.  Thread t of a group owns 16-byte chunk `sub` (8 channels) of eight rows of ONE centroid:
    // rows cq*32 + (rb & 3) + 4 i, so -v (the centroid's half of the first Linear) is formed once per item and reused
    // by all eight rows.  The eight lanes of a quarter-warp cover one 128-byte operand row (two 64-byte Qx rows when
    // paired): the Qx reads and the swizzled A writes are conflict-free.
    constexpr bool kTileMode = Cfg::KB <= 2;
    const int group = (warp - 8) >> 2;
    const int t = threadIdx.x & 127;
    const int sub = t & 7, rb = t >> 3;
    const int cq = rb >> 2, r0 = cq * 32 + (rb & 3);     // my rows: r0 + 4 i
    const int set = PAIR ? (sub >> 2) : 0;
    const int ch0 = (PAIR ? (sub & 3) : sub) * 8;         // my 8 channels inside a 64-channel slice of a Qx row
    const ulonglong2* tab2 = reinterpret_cast<const ulonglong2*>(table);
    // -v for my 8 channels: W1p . (o' - pos_i), fp32, rounded once to fp16 (saturating)
    // paired tiles (SA1) have ONE 64-channel slice: the thread's 24 W1p values never change and stay in registers (the six
    // 16-byte table reads were a quarter of the gather's shared-memory wavefronts per item)
    ulonglong2 tabr[6];
    if (PAIR) {
#pragma unroll
      for (int k = 0; k < 6; ++k) tabr[k] = tab2[(sub & 3) + 8 * k];
    }
    auto neg_v8 = [&](int slice, float ex, float ey, float ez, uint32_t (&nv)[4]) {
      const ulonglong2* tb = tab2 + slice * 48 + (PAIR ? (sub & 3) : sub);  // 48 = 3 coordinates x 2 halves x 8 groups
      const ulonglong2 wx01 = PAIR ? tabr[0] : tb[0], wx23 = PAIR ? tabr[1] : tb[8], wy01 = PAIR ? tabr[2] : tb[16],
                       wy23 = PAIR ? tabr[3] : tb[24], wz01 = PAIR ? tabr[4] : tb[32], wz23 = PAIR ? tabr[5] : tb[40];
      const uint64_t wx[4] = {wx01.x, wx01.y, wx23.x, wx23.y}, wy[4] = {wy01.x, wy01.y, wy23.x, wy23.y},
                     wz[4] = {wz01.x, wz01.y, wz23.x, wz23.y};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 x2 = *reinterpret_cast<const float2*>(&wx[q]), y2 = *reinterpret_cast<const float2*>(&wy[q]),
                     z2 = *reinterpret_cast<const float2*>(&wz[q]);
        const float lo = fmaf(z2.x, ez, fmaf(y2.x, ey, x2.x * ex)), hi = fmaf(z2.y, ez, fmaf(y2.y, ey, x2.y * ex));
        nv[q] = s2_pack_half2(lo, hi);
      }
    };
    uint32_t u = 0;     // ring item counter of this CTA
    uint32_t tile = 0;  // tile counter of this CTA (self-loop tiles included)
    for (int o = o0, n = 0; o < o1; ++o, ++n) {
      const int buf = n % Cfg::NOBJ;
      const uint8_t* ob = obj_base + buf * Cfg::OBJ_BYTES;
      const uint8_t* px_s = ob;
      const uint8_t* spx_s = ob + Cfg::PX_BYTES;
      const float* cpos_s = reinterpret_cast<const float*>(spx_s + Cfg::SPX_BYTES);
      const uint8_t* nbr_s = reinterpret_cast<const uint8_t*>(cpos_s) + Cfg::CPOS_BYTES;
      const uint8_t* cnt_s = nbr_s + Cfg::NBR_BYTES;
      const float* sorg_s = reinterpret_cast<const float*>(cnt_s + Cfg::CNT_BYTES);
      mbar_wait(&obj_full[buf], (n / Cfg::NOBJ) & 1);
#pragma unroll 1
      for (int tt = 0; tt < kTilesPerObj; ++tt, ++tile) {
        if (kTileMode && static_cast<int>(tile & 1) != group) { u += Cfg::KB; continue; }
        const uint32_t u_tile = u;
        if (tt == 0) {
          // ---- self-loop tile: row e = the self-loop edge of centroid e (source row e of the other object's half block);
          // every row has its own centroid, so -v is formed per row.  Only the tile's live rows are built.
          const bool live = r0 < Cfg::SELF_COLS;  // my centroid quarter lies inside the live rows (warp-uniform per quarter-warp group)
          const float ox = sorg_s[0], oy = sorg_s[1], oz = sorg_s[2];  // origin of the SOURCE object: its Qx rows are relative to it
#pragma unroll 1
          for (int kb = 0; kb < Cfg::KB; ++kb, ++u) {
            if (!kTileMode && static_cast<int>(u & 1) != group) continue;
            const int stage = static_cast<int>(u % Cfg::STAGES);
            mbar_wait(&empty_bar[stage], ((u / Cfg::STAGES) & 1) ^ 1);
            uint8_t* abase = stage_base + stage * Cfg::A_BYTES;
            if (live) {
#pragma unroll 2
              for (int i = 0; i < 8; ++i) {
                const int r = r0 + 4 * i;
                const int cen = PAIR ? 8 * (r >> 2) + 4 * set + (r & 3) : r;  // same centroid split between the lane groups as the neighbour tiles
                const uint4 raw = *reinterpret_cast<const uint4*>(spx_s + cen * (C1 * 2) + (PAIR ? 0 : kb * 128) + ch0 * 2);
                uint32_t nv[4];
                neg_v8(PAIR ? 0 : kb, ox - cpos_s[cen * 3 + 0], oy - cpos_s[cen * 3 + 1], oz - cpos_s[cen * 3 + 2], nv);
                *reinterpret_cast<uint4*>(abase + r * 128 + ((sub ^ (r & 7)) << 4)) =
                    make_uint4(s2_sub_relu_h2(raw.x, nv[0]), s2_sub_relu_h2(raw.y, nv[1]), s2_sub_relu_h2(raw.z, nv[2]), s2_sub_relu_h2(raw.w, nv[3]));
              }
            }
            if (!kTileMode) {
              fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core's async proxy
              __syncwarp();
              if (lane == 0) mbar_arrive(&full_bar[stage]);
            }
          }
        } else {
          // ---- neighbour tile: my eight rows are edge slots (rb & 3) + 4 i of centroid cen
          const int cen = (PAIR ? (2 * (tt - 1) + set) * 4 : (tt - 1) * 4) + cq;
          const int cnt = cnt_s[cen];
          int src_off[8];
          {
            // the centroid's whole 32-byte list as two broadcast 16-byte reads; my slot (rb & 3) + 4 i is byte rb & 3 of word i
            const uint4 l0 = *reinterpret_cast<const uint4*>(nbr_s + cen * 32), l1 = *reinterpret_cast<const uint4*>(nbr_s + cen * 32 + 16);
            const uint32_t lw[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
            const int sh = (rb & 3) * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int sl = (rb & 3) + 4 * i;
              const int j = sl < cnt ? ((lw[i] >> sh) & 0xff) : (lw[0] & 0xff);  // empty slots replicate slot 0: the max is unchanged
              src_off[i] = j * (C1 * 2) + ch0 * 2;
            }
          }
          const float ex = cpos_s[0] - cpos_s[cen * 3 + 0], ey = cpos_s[1] - cpos_s[cen * 3 + 1], ez = cpos_s[2] - cpos_s[cen * 3 + 2];  // o - pos_i
#pragma unroll 1
          for (int kb = 0; kb < Cfg::KB; ++kb, ++u) {
            if (!kTileMode && static_cast<int>(u & 1) != group) continue;
            const int stage = static_cast<int>(u % Cfg::STAGES);
            uint4 raw[8];  // all Qx reads of the item in flight before anything waits
#pragma unroll
            for (int i = 0; i < 8; ++i) raw[i] = *reinterpret_cast<const uint4*>(px_s + src_off[i] + (PAIR ? 0 : kb * 128));
            uint32_t nv[4];
            neg_v8(PAIR ? 0 : kb, ex, ey, ez, nv);
            mbar_wait(&empty_bar[stage], ((u / Cfg::STAGES) & 1) ^ 1);
            uint8_t* abase = stage_base + stage * Cfg::A_BYTES;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = r0 + 4 * i;
              // 128B swizzle: 16-byte chunk `sub` of row r lives at chunk position sub ^ (r % 8)
              *reinterpret_cast<uint4*>(abase + r * 128 + ((sub ^ (r & 7)) << 4)) =
                  make_uint4(s2_sub_relu_h2(raw[i].x, nv[0]), s2_sub_relu_h2(raw[i].y, nv[1]), s2_sub_relu_h2(raw[i].z, nv[2]),
                             s2_sub_relu_h2(raw[i].w, nv[3]));
            }
            if (!kTileMode) {
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) mbar_arrive(&full_bar[stage]);
            }
          }
        }
        if (kTileMode) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int kb = 0; kb < Cfg::KB; ++kb) mbar_arrive(&full_bar[(u_tile + kb) % Cfg::STAGES]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&obj_empty[buf]);  // this warp no longer reads the object block
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <class Cfg>
static cudaError_t launch_sa_obj2(const SaObj2& a, cudaStream_t st) {
  static bool configured_dev[64] = {};  // the attribute is per device: one flag per device ordinal
  bool& configured = configured_dev[current_device() & 63];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(sa_obj2_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  SaObj2Params p{a.Qx16, a.cpos, a.nbr, a.cnt, a.loop_src_obj, a.loop_half, a.Wp, a.W2h, a.b2, a.out, a.ldo, a.n_obj, a.bisect};
  const int grid = a.n_obj < tma_api().num_sms ? a.n_obj : tma_api().num_sms;
  sa_obj2_kernel<Cfg><<<grid, kSa2Threads, Cfg::SMEM, st>>>(p);
  return cudaGetLastError();
}

cudaError_t sa_obj2(const SaObj2& a, cudaStream_t st, Launches* lc) {
  if (a.n_obj <= 0) return cudaSuccess;
  if (lc) lc->n++;
  if (a.ldo < a.C2) return cudaErrorInvalidValue;
  if (a.C1 == 32 && a.C2 == 64 && a.P == 256 && a.M == 128) return launch_sa_obj2<Sa2Cfg<32, 64, 256, 128, true>>(a, st);
  if (a.C1 == 128 && a.C2 == 128 && a.P == 128 && a.M == 64) return launch_sa_obj2<Sa2Cfg<128, 128, 128, 64, false>>(a, st);
  if (a.C1 == 256 && a.C2 == 256 && a.P == 64 && a.M == 32) return launch_sa_obj2<Sa2Cfg<256, 256, 64, 32, false>>(a, st);
  return cudaErrorInvalidValue;
}

}  // namespace t2l
