// Object-resident fused PointConv layer (models/pointcloud/pointnet2.py:25-37), fp16 operands.
//
//   out[i] = max over the <=32 ball-query neighbours j of  relu(W2 . relu(Px[j] + W1p.(pos_j - pos_i) + b1) + b2)
//            (then max'ed with the re-added self-loop edge, computed separately and passed as `side`)
//
// sa_fused.cu gathers every edge's Px row from global memory: a dependent, mostly L2-missing load
// per row that the gather warps can only cover with registers, and for SA3 a 256 KB fp32 W2 that
// has to stream through shared memory for every tile (profiles/r01/sa_fused_bisect_v2.txt).  Here
// a CTA owns a run of whole OBJECTS and everything an object's edges touch is resident in shared
// memory before the first edge is built:
//   * the object's Px block [P, C1] as fp16 (16-32 KB), its dense positions, centroid positions and
//     ball-query lists -- five 1-D bulk TMA copies per object landing on one mbarrier, issued one
//     object ahead where the buffers fit twice;
//   * W2 as fp16 [C2, C1], loaded once per CTA (4 / 32 / 128 KB).
// The gather warps then read Px rows with LDS (no global latency, no per-edge records: the edge's
// source point and pos_j - pos_i come straight from the ball-query list and the positions), add the
// exact fp32 position term, apply ReLU and write the fp16 A tile (128 edge rows x 64 channels, 128-byte
// swizzled rows) that `tcgen05.mma kind::f16` consumes.  fp16 has the 11-bit significand of the tf32
// rounding the layer used before, fp32 accumulation in TMEM is unchanged; A and B bytes per MMA halve,
// the MMA rate doubles.  Empty neighbour slots replicate slot 0 (the set always contains the centroid's
// own point, and a repeated member does not change a max), so no row ever refers to another object.
//
// The product is computed TRANSPOSED: D^T[channel, edge] = W2[channel, :] . A[edge, :], i.e. W2 is the MMA's M-side
// operand and the gathered tile its N-side operand (both are K-major 128-byte-row tiles, so nothing else changes).
// In TMEM a lane is then an output CHANNEL and the 128 columns are the tile's 4 centroids x 32 edges, so the
// 32x32b TMEM load hands every thread the 32 edges of one centroid for its channel: PointConv's max aggregation
// is 31 in-register max instructions, bias and ReLU are applied once per (centroid, channel), and a warp stores 32
// consecutive channels of one centroid (one coalesced 128-byte line).  The row-major variant needed a 31-shuffle
// select butterfly per 32 columns -- ~1 760 instructions per tile and warp, MIO-bound next to the gather warps'
// LDS/STS traffic -- and was the critical path of every sa_fused variant (profiles/r01/sa_fused_bisect_v2.txt).
//
// Warp roles (512 threads): 0 = TMA producer (W2 once, then one object block per object), 1 = MMA
// issuer, 2 = TMEM allocator, 4..7 = epilogue (SegMax: one centroid per TMEM lane quadrant), 8..15 =
// gather.  Pipelines: object buffers full/empty (TMA <-> gather), A ring full/empty (gather <-> MMA),
// TMEM accumulators full/empty (MMA <-> epilogue).
#include <cstdlib>

#include "ops.h"
#include "umma_gemm.cuh"
#include "gemm_epilogues.cuh"

namespace t2l {

struct SaObjOut {
  float* out;         // [n*M, C2]
  const float* b2;    // [C2]
  const float* side;  // [n*M, C2] post-ReLU self-loop rows (>= 0)
};

struct SaObjParams {
  const __half* Px16;
  const float* dense_pos;
  const float* cpos;
  const uint8_t* nbr;
  const uint8_t* cnt;
  const float* Wp;
  const float* b1;
  int n_obj;
  int dbg;  // timing bisect only (T2L_SA_DBG bitmask; results are then wrong): 1 no proxy fence, 2 no output stores, 4 no side loads
};

template <int C1, int C2, int P, int M, int POS_STRIDE>
struct SaObjCfg {
  static constexpr int CH = C1 < 64 ? C1 : 64;  // channels per k-block item (fp16: 64 = one 128-byte swizzle row)
  static constexpr int KB = C1 / CH;            // items per tile
  static constexpr int UMMAS = CH / 16;         // kind::f16 K = 16 per instruction
  static constexpr int CPR = CH / 8;            // 16-byte chunks per A row that carry data
  static constexpr int RPT = CPR;               // rows per gather thread per item (a group of 128 threads builds one item)
  static constexpr int RSTEP = 128 / CPR;
  static constexpr int TPO = M / 4;             // tiles (4 centroids x 32 slots) per object
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int MH = C2 > 128 ? C2 / 128 : 1;          // 128-channel halves of W2 = MMAs (M = 128) per K slice
  static constexpr int B_BYTES = MH * 128 * 128;  // one k-block of W2: MH x 128 rows x 64 halfs (rows >= C2 zero-filled by TMA)
  static constexpr int B_RES_BYTES = KB * B_BYTES;
  static constexpr int PX_BYTES = P * C1 * 2;
  static constexpr int POS_BYTES = P * POS_STRIDE * 4;
  static constexpr int CPOS_BYTES = M * 12;
  static constexpr int NBR_BYTES = M * 32;
  static constexpr int CNT_BYTES = M;
  static constexpr int OBJ_BYTES = PX_BYTES + POS_BYTES + CPOS_BYTES + NBR_BYTES + CNT_BYTES;  // every part a multiple of 16
  static constexpr int NOBJ = (B_RES_BYTES + 2 * OBJ_BYTES + 3 * A_BYTES) <= 190 * 1024 ? 2 : 1;
  // A ring: four slots where they fit (slot parity = owning gather group), else three.  With three slots the two groups
  // alternate on every slot; a parity wait is still unambiguous because the MMA warp consumes items in order: when a group
  // waits for slot s of item u (previous use: item u-3) it has already filled u-2, whose slot could only be free after
  // the MMAs of u-5 -- hence of u-6, the use before the one it waits for -- had completed.  So the barrier is never more
  // than one phase behind the waiter (and cannot be ahead: the next phase needs the item the waiter has not built yet).
  static constexpr int STAGES = (B_RES_BYTES + NOBJ * OBJ_BYTES + 4 * A_BYTES) <= 200 * 1024 ? 4 : 3;
  static constexpr int ACC_COLS = MH * 128;     // accumulator of one tile: [128 channels] x [128 edge rows] per half
  static constexpr int NACC = 512 / ACC_COLS;
  static constexpr int TMEM_COLS = NACC * ACC_COLS;
  static constexpr int BAR_BYTES = 256;
  static constexpr int TABLE_BYTES = C1 * 12;   // w1p as channel pairs: x2[C1/2] | y2[C1/2] | z2[C1/2] (float2 each)
  static constexpr uint32_t IDESC = umma_idesc(0u, 128, 128);  // f16 x f16 -> f32, M = 128 channels, N = 128 edge rows
  static constexpr int SMEM = 1024 + B_RES_BYTES + STAGES * A_BYTES + NOBJ * OBJ_BYTES + BAR_BYTES + TABLE_BYTES;
  static_assert(PX_BYTES % 16 == 0 && POS_BYTES % 16 == 0 && CPOS_BYTES % 16 == 0 && NBR_BYTES % 16 == 0 && CNT_BYTES % 16 == 0, "bulk copies");
  static_assert((2 * STAGES + 2 * NACC + 1 + 2 * NOBJ) * 8 + 4 <= BAR_BYTES, "barrier block too small");
  static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns must be a power of two");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

constexpr int kSaObjThreads = 512;

T2L_DEVICE void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Packed fp32 pairs live in 64-bit registers from the shared-memory load to the fp16 pack, so the two-lane FMAs
// (sm_100 packed fp32 pipe) need no per-operand MOVs (a float2-based wrapper cost ~150 MOVs per item and thread).
T2L_DEVICE uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
T2L_DEVICE uint64_t dup2(float x) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %1};" : "=l"(d) : "f"(x));
  return d;
}
T2L_DEVICE uint64_t half2_to_f32x2(uint32_t h2) {
  uint64_t d;
  asm("{\n\t.reg .f16 l, h;\n\t.reg .f32 a, b;\n\t"
      "mov.b32 {l, h}, %1;\n\t"
      "cvt.f32.f16 a, l;\n\tcvt.f32.f16 b, h;\n\t"
      "mov.b64 %0, {a, b};\n\t}"
      : "=l"(d) : "r"(h2));
  return d;
}
// (lo, hi) fp32 pair -> packed fp16x2 with ReLU, round-to-nearest, saturating at 65504: one F2FP instruction
T2L_DEVICE uint32_t relu_pack_half2(uint64_t v) {
  uint32_t r;
  asm("{\n\t.reg .f32 a, b;\n\tmov.b64 {a, b}, %1;\n\tcvt.rn.relu.satfinite.f16x2.f32 %0, b, a;\n\t}" : "=r"(r) : "l"(v));
  return r;
}

template <int C1, int C2, int P, int M, int POS_STRIDE>
__global__ void __launch_bounds__(kSaObjThreads, 1)
sa_obj_kernel(const __grid_constant__ CUtensorMap tm_b, const SaObjParams p, const SaObjOut ep) {
  using Cfg = SaObjCfg<C1, C2, P, M, POS_STRIDE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* b_res = smem;                                   // W2 k-block tiles, 1024-aligned (B_BYTES is a multiple of 1 KB)
  uint8_t* stage_base = b_res + Cfg::B_RES_BYTES;          // A ring, 1024-aligned
  uint8_t* obj_base = stage_base + Cfg::STAGES * Cfg::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(obj_base + Cfg::NOBJ * Cfg::OBJ_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + Cfg::NACC;
  uint64_t* bres_bar = tmem_empty + Cfg::NACC;
  uint64_t* obj_full = bres_bar + 1;
  uint64_t* obj_empty = obj_full + Cfg::NOBJ;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(obj_empty + Cfg::NOBJ);
  float* table = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + Cfg::BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_b);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 4);  // one arrive per warp of the owning gather group
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < Cfg::NACC; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    mbar_init(bres_bar, 1);
    for (int i = 0; i < Cfg::NOBJ; ++i) {
      mbar_init(&obj_full[i], 1);
      mbar_init(&obj_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
  for (int c = threadIdx.x; c < C1; c += kSaObjThreads) {  // b1 is already folded into Px
    table[c] = p.Wp[c * 4 + 0];
    table[C1 + c] = p.Wp[c * 4 + 1];
    table[2 * C1 + c] = p.Wp[c * 4 + 2];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // contiguous run of whole objects per CTA
  const int o0 = static_cast<int>(static_cast<long>(p.n_obj) * blockIdx.x / gridDim.x);
  const int o1 = static_cast<int>(static_cast<long>(p.n_obj) * (blockIdx.x + 1) / gridDim.x);
  const int n_tiles = (o1 - o0) * Cfg::TPO;
  const long tile0 = static_cast<long>(o0) * Cfg::TPO;

  if (warp == 0) {
    // ================= TMA producer: W2 once, then one block per object =================
    if (elect_one()) {
      mbar_arrive_expect_tx(bres_bar, Cfg::B_RES_BYTES);
      for (int kb = 0; kb < Cfg::KB; ++kb) tma_load_2d(&tm_b, bres_bar, b_res + kb * Cfg::B_BYTES, kb * 64, 0, kEvictLast);
    }
    __syncwarp();
    constexpr int kSideBytes = M * C2 * 4;  // one object's self-loop rows, contiguous
    if (o0 < o1 && lane < 8)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(ep.side) + static_cast<long>(o0) * kSideBytes + lane * (kSideBytes / 8)),
                   "r"(kSideBytes / 8) : "memory");
    for (int o = o0, n = 0; o < o1; ++o, ++n) {
      const int buf = n % Cfg::NOBJ;
      if (o + 1 < o1 && lane < 8)  // the epilogue reaches the next object's side rows in ~10-20 us
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(ep.side) + static_cast<long>(o + 1) * kSideBytes + lane * (kSideBytes / 8)),
                     "r"(kSideBytes / 8) : "memory");
      mbar_wait(&obj_empty[buf], (((n / Cfg::NOBJ) & 1) ^ 1));
      if (elect_one()) {
        uint8_t* dst = obj_base + buf * Cfg::OBJ_BYTES;
        mbar_arrive_expect_tx(&obj_full[buf], Cfg::OBJ_BYTES);
        bulk_load(dst, p.Px16 + static_cast<long>(o) * P * C1, Cfg::PX_BYTES, &obj_full[buf]);
        dst += Cfg::PX_BYTES;
        bulk_load(dst, p.dense_pos + static_cast<long>(o) * P * POS_STRIDE, Cfg::POS_BYTES, &obj_full[buf]);
        dst += Cfg::POS_BYTES;
        bulk_load(dst, p.cpos + static_cast<long>(o) * M * 3, Cfg::CPOS_BYTES, &obj_full[buf]);
        dst += Cfg::CPOS_BYTES;
        bulk_load(dst, p.nbr + static_cast<long>(o) * M * 32, Cfg::NBR_BYTES, &obj_full[buf]);
        dst += Cfg::NBR_BYTES;
        bulk_load(dst, p.cnt + static_cast<long>(o) * M, Cfg::CNT_BYTES, &obj_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    mbar_wait(bres_bar, 0);
    for (int tile = 0; tile < n_tiles; ++tile) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_addr = tmem_base + acc * Cfg::ACC_COLS;
      for (int kb = 0; kb < Cfg::KB; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {  // see umma_gemm.cuh: keeps UTCHMMA/UTCBAR straight-line
          const uint64_t edesc = umma_desc_sw128(stage_base + stage * Cfg::A_BYTES);  // edge tile: the N-side operand
#pragma unroll
          for (int h = 0; h < Cfg::MH; ++h) {
            const uint64_t wdesc = umma_desc_sw128(b_res + kb * Cfg::B_BYTES + h * (128 * 128));  // 128 channels of W2: the M-side operand
#pragma unroll
            for (int k = 0; k < Cfg::UMMAS; ++k) umma_f16(d_addr + h * 128, wdesc + 2 * k, edesc + 2 * k, Cfg::IDESC, (kb | k) != 0);
          }
          tc_commit(&empty_bar[stage]);
          if (kb == Cfg::KB - 1) tc_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == Cfg::NACC) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4 && warp < 8) {
    // ================= epilogue: max over each centroid's 32 edges (in registers), bias, ReLU, self-loop row =================
    const int ew = warp - 4;  // TMEM lane quadrant = channels 32 ew .. 32 ew + 31 of each 128-channel half
    float bias[Cfg::MH];
#pragma unroll
    for (int h = 0; h < Cfg::MH; ++h) bias[h] = (h * 128 + ew * 32 + lane < C2) ? __ldg(ep.b2 + h * 128 + ew * 32 + lane) : 0.f;
    int acc = 0;
    uint32_t acc_phase = 0;
    if (ew * 32 < C2 || Cfg::MH > 1) {
      // The self-loop rows (`side`) are the only global reads of the epilogue.  Loaded at the top of a tile's own iteration
      // they put one DRAM round trip (~1.4k cycles) on EVERY tile -- the whole per-tile time of SA1 (profiles/r01).  So:
      // the TMA warp L2-prefetches each object's side block one object ahead, and the values are register-prefetched
      // kSideAhead tiles ahead here.
      constexpr int kSideAhead = 2;
      float side_q[kSideAhead + 1][Cfg::MH][4];
      auto load_side = [&](int tile, float (&dst)[Cfg::MH][4]) {
        const long g0 = (tile0 + tile) * 4;
#pragma unroll
        for (int h = 0; h < Cfg::MH; ++h)
#pragma unroll
          for (int c = 0; c < 4; ++c)
            dst[h][c] = (tile < n_tiles && !(p.dbg & 4)) ? __ldg(ep.side + (g0 + c) * C2 + h * 128 + ew * 32 + lane) : 0.f;
      };
#pragma unroll
      for (int a = 0; a < kSideAhead; ++a) load_side(a, side_q[a]);
      for (int tile = 0; tile < n_tiles; ++tile) {
        const long g0 = (tile0 + tile) * 4;  // first centroid of the tile
        load_side(tile + kSideAhead, side_q[kSideAhead]);
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * Cfg::ACC_COLS;
        float v[2][32];
        tmem_ld_32x32(t_addr, v[0]);
#pragma unroll
        for (int q = 0; q < Cfg::MH * 4; ++q) {  // chunk q = (half q / 4, centroid q % 4): this thread's channel, 32 edges
          tmem_ld_wait(v[q & 1]);
          if (q + 1 < Cfg::MH * 4) tmem_ld_32x32(t_addr + (q + 1) * 32, v[(q + 1) & 1]);
          const float* x = v[q & 1];
          float m[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) m[i] = fmaxf(fmaxf(x[4 * i], x[4 * i + 1]), fmaxf(x[4 * i + 2], x[4 * i + 3]));
          const float mx = fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
          // bias and ReLU commute with the max over edges (per-channel constant, monotonic rounding)
          const float keep = fmaxf(fmaxf(mx + bias[q >> 2], 0.f), side_q[0][q >> 2][q & 3]);
          if (!(p.dbg & 2) || keep == 12345.f) ep.out[(g0 + (q & 3)) * C2 + (q >> 2) * 128 + ew * 32 + lane] = round_tf32(keep);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (++acc == Cfg::NACC) { acc = 0; acc_phase ^= 1; }
#pragma unroll
        for (int a = 0; a < kSideAhead; ++a)
#pragma unroll
          for (int h = 0; h < Cfg::MH; ++h)
#pragma unroll
            for (int c = 0; c < 4; ++c) side_q[a][h][c] = side_q[a + 1][h][c];
      }
    } else {
      // channels >= C2 (SA1: 64 real channels in a 128-lane accumulator): nothing to read, just release the buffers
      for (int tile = 0; tile < n_tiles; ++tile) {
        mbar_wait(&tmem_full[acc], acc_phase);
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (++acc == Cfg::NACC) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 8) {
    // ================= gather: A tiles from the resident object block =================
    // Two groups of four warps build alternate items (item u = tile * KB + kb goes to group u % 2, ring slot u % STAGES),
    // so the ~600-cycle generic->async proxy fence that ends every item (the largest single cost of the gather loop,
    // profiles/r01) overlaps the other group's loads and math.  Thread t of a group owns 16-byte chunk `sub`
    // (8 channels) of rows rb, rb + RSTEP, ...: 8 (or 4) lanes cover one row's 128 (64) data bytes, so the Px reads and
    // the swizzled A writes are conflict-free quarter-warp accesses.
    const int group = (warp - 8) >> 2;
    const int t = threadIdx.x & 127;
    const int sub = t % Cfg::CPR, rb = t / Cfg::CPR;
    const ulonglong2* tab2 = reinterpret_cast<const ulonglong2*>(table);  // two channel pairs per 16-byte load
    long u = 0;  // item counter of this CTA
    for (int o = o0, n = 0; o < o1; ++o, ++n) {
      const int buf = n % Cfg::NOBJ;
      const uint8_t* ob = obj_base + buf * Cfg::OBJ_BYTES;
      const uint8_t* px_s = ob;
      const float* pos_s = reinterpret_cast<const float*>(ob + Cfg::PX_BYTES);
      const float* cpos_s = reinterpret_cast<const float*>(ob + Cfg::PX_BYTES + Cfg::POS_BYTES);
      const uint8_t* nbr_s = ob + Cfg::PX_BYTES + Cfg::POS_BYTES + Cfg::CPOS_BYTES;
      const uint8_t* cnt_s = nbr_s + Cfg::NBR_BYTES;
      mbar_wait(&obj_full[buf], (n / Cfg::NOBJ) & 1);
#pragma unroll 1
      for (int tau = 0; tau < Cfg::TPO; ++tau) {
        // kTileMode (KB <= 2, four slots): the groups alternate TILES; a group builds all k-blocks of its tile and pays ONE
        // proxy fence for them (its items always land in the same two slots).  Otherwise (SA3: two slots) they alternate items.
        constexpr bool kTileMode = Cfg::KB <= 2 && Cfg::STAGES == 4;
        if (kTileMode && ((u / Cfg::KB) & 1) != group) { u += Cfg::KB; continue; }
        // this tile's edges: source point and pos_j - pos_i of my rows (exact fp32 subtraction, as the reference's message())
        int src_off[Cfg::RPT];
        uint64_t dx[Cfg::RPT], dy[Cfg::RPT], dz[Cfg::RPT];
#pragma unroll
        for (int i = 0; i < Cfg::RPT; ++i) {
          const int r = rb + Cfg::RSTEP * i;
          const int cen = tau * 4 + (r >> 5), sl = r & 31;
          const int j = nbr_s[cen * 32 + (sl < cnt_s[cen] ? sl : 0)];  // empty slots replicate slot 0: the max is unchanged
          src_off[i] = j * (C1 * 2) + sub * 16;
          dx[i] = dup2(pos_s[j * POS_STRIDE + 0] - cpos_s[cen * 3 + 0]);
          dy[i] = dup2(pos_s[j * POS_STRIDE + 1] - cpos_s[cen * 3 + 1]);
          dz[i] = dup2(pos_s[j * POS_STRIDE + 2] - cpos_s[cen * 3 + 2]);
        }
        const long u_tile = u;
#pragma unroll 1
        for (int kb = 0; kb < Cfg::KB; ++kb, ++u) {
          if (!kTileMode && (u & 1) != group) continue;
          const int stage = static_cast<int>(u % Cfg::STAGES);
          const uint32_t use = static_cast<uint32_t>(u / Cfg::STAGES);
          uint4 raw[Cfg::RPT];  // all Px reads of the item in flight before anything waits
#pragma unroll
          for (int i = 0; i < Cfg::RPT; ++i) raw[i] = *reinterpret_cast<const uint4*>(px_s + src_off[i] + kb * 128);
          // w1p of my 8 channels as pairs: (x, y, z) x 4 pairs
          const int pair0 = (kb * 64 + sub * 8) >> 2;  // index in ulonglong2 units (2 pairs each)
          const ulonglong2 wx01 = tab2[pair0], wx23 = tab2[pair0 + 1];
          const ulonglong2 wy01 = tab2[C1 / 4 + pair0], wy23 = tab2[C1 / 4 + pair0 + 1];
          const ulonglong2 wz01 = tab2[C1 / 2 + pair0], wz23 = tab2[C1 / 2 + pair0 + 1];
          const uint64_t wx[4] = {wx01.x, wx01.y, wx23.x, wx23.y}, wy[4] = {wy01.x, wy01.y, wy23.x, wy23.y},
                         wz[4] = {wz01.x, wz01.y, wz23.x, wz23.y};
          mbar_wait(&empty_bar[stage], (use & 1) ^ 1);
          uint8_t* abase = stage_base + stage * Cfg::A_BYTES;
#pragma unroll
          for (int i = 0; i < Cfg::RPT; ++i) {
            const int r = rb + Cfg::RSTEP * i;
            const uint32_t rw[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
            uint32_t packed[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {  // channels 2q, 2q+1:  Px (+ b1) + w1p . (pos_j - pos_i)
              uint64_t v = half2_to_f32x2(rw[q]);
              v = fma2(wx[q], dx[i], v);
              v = fma2(wy[q], dy[i], v);
              v = fma2(wz[q], dz[i], v);
              packed[q] = relu_pack_half2(v);
            }
            // 128B swizzle: 16-byte chunk `sub` of row r lives at chunk position sub ^ (r % 8)
            *reinterpret_cast<uint4*>(abase + r * 128 + ((sub ^ (r & 7)) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          }
          if (!kTileMode) {
            if (!(p.dbg & 1)) fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[stage]);
          }
        }
        if (kTileMode) {
          if (!(p.dbg & 1)) fence_proxy_async();
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
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int C1, int C2, int P, int M, int POS_STRIDE>
static cudaError_t launch_sa_obj(const SaObj& a, cudaStream_t st) {
  using Cfg = SaObjCfg<C1, C2, P, M, POS_STRIDE>;
  CUtensorMap tb;
  if (make_operand_map(&tb, a.W2h, kOpF16, C2, C1, C1, Cfg::MH * 128 > 256 ? 256 : Cfg::MH * 128)) return cudaErrorInvalidValue;
  static bool configured_dev[64] = {};  // the attribute is per device: one flag per device ordinal
  bool& configured = configured_dev[current_device() & 63];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(sa_obj_kernel<C1, C2, P, M, POS_STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const char* dbg_env = getenv("T2L_SA_DBG");
  SaObjParams p{a.Px16, a.dense_pos, a.cpos, a.nbr, a.cnt, a.Wp, a.b1, a.n_obj, dbg_env ? atoi(dbg_env) : 0};
  SaObjOut ep{a.out, a.b2, a.side};
  const int grid = a.n_obj < tma_api().num_sms ? a.n_obj : tma_api().num_sms;
  sa_obj_kernel<C1, C2, P, M, POS_STRIDE><<<grid, kSaObjThreads, Cfg::SMEM, st>>>(tb, p, ep);
  return cudaGetLastError();
}

cudaError_t sa_obj(const SaObj& a, cudaStream_t st, Launches* lc) {
  if (a.n_obj <= 0) return cudaSuccess;
  if (lc) lc->n++;
  if (a.C1 == 32 && a.C2 == 64 && a.P == 256 && a.M == 128 && a.dense_stride == 6) return launch_sa_obj<32, 64, 256, 128, 6>(a, st);
  if (a.C1 == 128 && a.C2 == 128 && a.P == 128 && a.M == 64 && a.dense_stride == 3) return launch_sa_obj<128, 128, 128, 64, 3>(a, st);
  if (a.C1 == 256 && a.C2 == 256 && a.P == 64 && a.M == 32 && a.dense_stride == 3) return launch_sa_obj<256, 256, 64, 32, 3>(a, st);
  return cudaErrorInvalidValue;
}

}  // namespace t2l
