// Linear layers: C = act(A * W^T + bias).  Two kernels, picked by the caller per layer:
//   linear_umma  tcgen05 kind::tf32 GEMM (umma_gemm.cuh) for the layers that carry the FLOPs
//   linear_simt  exact fp32 FMA GEMM for tiny-K layers and the precision-critical tails
#include "ops.h"
#include "umma_gemm.cuh"
#include "gemm_epilogues.cuh"

namespace t2l {

// ---------------------------------------------------------------------------------------
// tcgen05 path
// ---------------------------------------------------------------------------------------
template <int BN, int GROUP, class Epi, int TYPE = kOpTf32>
static cudaError_t run_umma(const Linear& l, const typename Epi::Params& ep, cudaStream_t st) {
  using Cfg = GemmCfg<BN, TYPE, GROUP>;
  CUtensorMap ta, tb;
  const int kw = l.passes == 3 ? 2 * l.K : l.K;  // physical operand width
  if (make_operand_map(&ta, l.A, TYPE, l.M, kw, l.lda, Cfg::BLOCK_M)) return cudaErrorInvalidValue;
  if (make_operand_map(&tb, l.W, TYPE, l.N, kw, l.ldw, Cfg::LOAD_N)) return cudaErrorInvalidValue;
  GemmShape s;
  s.M = l.M; s.N = l.N;
  s.m_tiles = (l.M + Cfg::BLOCK_M * GROUP - 1) / (Cfg::BLOCK_M * GROUP);
  s.n_tiles = (l.N + BN - 1) / BN;
  s.n_splits = s.n_tiles;
  s.tiles_per_split = 1;
  s.ks.n_pass = l.passes;
  s.ks.kb_per_pass = (l.K + Cfg::BLOCK_K - 1) / Cfg::BLOCK_K;
  s.ks.a_off[0] = 0;   s.ks.b_off[0] = 0;    // hi * hi
  s.ks.a_off[1] = l.K; s.ks.b_off[1] = 0;    // lo * hi
  s.ks.a_off[2] = 0;   s.ks.b_off[2] = l.K;  // hi * lo
  return launch_umma_gemm<Cfg, Epi>(ta, tb, s, ep, st);
}

// fp16 residual stream (out-proj, FFN2 of the token layer): residual and output rows move by TMA (ResidualTmaEpi)
static cudaError_t run_residual_tma(const Linear& l, cudaStream_t st) {
  ResidualTmaParams ep;
  if (make_operand_map(&ep.tm_res, l.residual, kOpF16, l.M, l.N, l.ldr, 128)) return cudaErrorInvalidValue;
  if (make_operand_map(&ep.tm_out, l.C, kOpF16, l.M, l.N, l.ldc, 128)) return cudaErrorInvalidValue;
  ep.bias = l.bias; ep.M = l.M; ep.N = l.N; ep.act = l.act; ep.half_max = l.half_max;
  return run_umma<256, 2, ResidualTmaEpi, kOpF16>(l, ep, st);
}

template <class Epi, int TYPE>
static cudaError_t run_store(const Linear& l, const StoreParams& ep, bool wide, bool pair, cudaStream_t st) {
  if (pair) return run_umma<256, 2, Epi, TYPE>(l, ep, st);
  return wide ? run_umma<256, 1, Epi, TYPE>(l, ep, st) : run_umma<128, 1, Epi, TYPE>(l, ep, st);
}

cudaError_t linear_umma(const Linear& l, cudaStream_t st, Launches* lc) {
  if (l.M <= 0) return cudaSuccess;
  const int row_align = l.half_ops ? 8 : 4;  // 16-byte rows for TMA
  if ((l.N % 32) || (l.lda % row_align) || (l.ldw % row_align)) return cudaErrorInvalidValue;
  if (l.passes != 1 && (l.passes != 3 || l.K % 32 || l.half_ops)) return cudaErrorInvalidValue;
  if (lc) lc->n++;
  // N % 256 == 0 and more than one 128-row tile: CTA pairs (256 x 256 tiles, cta_group::2); else single CTAs
  const bool wide = (l.N % 256) == 0;
  const bool pair = wide && l.M > 128;
  if (l.segmax) {
    if (l.M % 32 || !l.bias) return cudaErrorInvalidValue;
    SegMaxEpi::Params ep{l.C, l.ldc, l.bias, l.side, l.lds, l.M, l.N, l.round_out};
    if (l.half_ops) {  // GA's second layer: N = 1024
      if (!wide) return cudaErrorInvalidValue;
      return pair ? run_umma<256, 2, SegMaxEpi, kOpF16>(l, ep, st) : run_umma<256, 1, SegMaxEpi, kOpF16>(l, ep, st);
    }
    if (pair) return run_umma<256, 2, SegMaxEpi>(l, ep, st);
    return wide ? run_umma<256, 1, SegMaxEpi>(l, ep, st) : run_umma<128, 1, SegMaxEpi>(l, ep, st);
  }
  if (l.out_half && l.residual && !l.half_ops) return cudaErrorInvalidValue;
  if (l.split_out && (l.residual || l.out_half || l.ldc < 2L * l.N)) return cudaErrorInvalidValue;
  StoreParams ep{l.C, l.ldc, l.bias, l.residual, l.ldr, l.M, l.N, l.act, l.round_out, l.out_half, l.half_max, l.residual_half, l.split_out};
  if (l.half_ops) {  // fp16 operands (A, W are __half), kind::f16: twice the tf32 rate at the same 11-bit significand
    if (l.out_half && l.residual) {  // fp16 residual stream
      if (pair && l.residual_half && !l.reg_epilogue && !(l.ldr % 8) && !(l.ldc % 8)) return run_residual_tma(l, st);
      return run_store<StoreEpiT<true, true>, kOpF16>(l, ep, wide, pair, st);
    }
    if (l.out_half) return run_store<StoreEpiT<true, false>, kOpF16>(l, ep, wide, pair, st);
    if (l.residual) return run_store<StoreEpiT<false, true>, kOpF16>(l, ep, wide, pair, st);
    return run_store<StoreEpiT<false, false>, kOpF16>(l, ep, wide, pair, st);
  }
  if (l.out_half) return run_store<StoreEpiT<true, false>, kOpTf32>(l, ep, wide, pair, st);  // tf32 GEMM feeding an fp16 consumer (Px)
  if (l.residual) return run_store<StoreEpiT<false, true>, kOpTf32>(l, ep, wide, pair, st);
  return run_store<StoreEpiT<false, false>, kOpTf32>(l, ep, wide, pair, st);
}

// fp32 -> [hi | lo] tf32 planes: x = hi + lo up to 2^-22 |x|
__global__ void split_tf32_kernel(const float* __restrict__ x, long ldx, float* __restrict__ planes, long rows, int K4) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * K4) return;
  const long r = (i < (1L << 31)) ? static_cast<long>(static_cast<unsigned>(i) / static_cast<unsigned>(K4)) : i / K4;  // 32-bit divide when it fits
  const int c = static_cast<int>(i - r * K4);
  const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + 4 * c);
  float4 hi = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
  float4 lo = make_float4(round_tf32(v.x - hi.x), round_tf32(v.y - hi.y), round_tf32(v.z - hi.z), round_tf32(v.w - hi.w));
  float4* dst = reinterpret_cast<float4*>(planes + r * 8 * K4);
  dst[c] = hi;
  dst[K4 + c] = lo;
}

cudaError_t split_tf32_planes(const float* x, long ldx, float* planes, int rows, int K, cudaStream_t st, Launches* lc) {
  if (rows <= 0) return cudaSuccess;
  if ((K % 4) || (ldx % 4)) return cudaErrorInvalidValue;
  if (lc) lc->n++;
  const long n = static_cast<long>(rows) * (K / 4);
  split_tf32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(x, ldx, planes, rows, K / 4);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// fp32 SIMT path: 64x64 tile, 16-deep k slices, 256 threads x (4x4) outputs
// ---------------------------------------------------------------------------------------
constexpr int SB_M = 64, SB_N = 64, SB_K = 16;

__global__ void __launch_bounds__(256) linear_simt_kernel(Linear l) {
  __shared__ float As[SB_K][SB_M + 4];
  __shared__ float Ws[SB_K][SB_N + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long m0 = static_cast<long>(blockIdx.x) * SB_M;
  const int n0 = blockIdx.y * SB_N;
  float acc[4][4] = {};
  // loader mapping: 256 threads cover 64 rows x 16 k as 4 k-values per thread
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const bool vec = ((l.lda | l.ldw | l.K) & 3) == 0 && ((reinterpret_cast<uintptr_t>(l.A) | reinterpret_cast<uintptr_t>(l.W)) & 15) == 0;
  for (int k0 = 0; k0 < l.K; k0 += SB_K) {
    const long ar = m0 + lr;
    const int wr = n0 + lr;
    if (vec) {
      const int k = k0 + lk;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), w = a;
      if (ar < l.M && k < l.K) a = *reinterpret_cast<const float4*>(l.A + ar * l.lda + k);
      if (wr < l.N && k < l.K) w = *reinterpret_cast<const float4*>(l.W + static_cast<long>(wr) * l.ldw + k);
      As[lk + 0][lr] = a.x; As[lk + 1][lr] = a.y; As[lk + 2][lr] = a.z; As[lk + 3][lr] = a.w;
      Ws[lk + 0][lr] = w.x; Ws[lk + 1][lr] = w.y; Ws[lk + 2][lr] = w.z; Ws[lk + 3][lr] = w.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + lk + j;
        As[lk + j][lr] = (ar < l.M && k < l.K) ? l.A[ar * l.lda + k] : 0.f;
        Ws[lk + j][lr] = (wr < l.N && k < l.K) ? l.W[static_cast<long>(wr) * l.ldw + k] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SB_K; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long r = m0 + ty * 4 + i;
    if (r >= l.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= l.N) continue;
      float x = acc[i][j];
      if (l.bias) x += l.bias[c];
      if (l.act == 1) x = fmaxf(x, 0.f);
      if (l.residual) x += l.residual_half ? __half2float(reinterpret_cast<const __half*>(l.residual)[r * l.ldr + c]) : l.residual[r * l.ldr + c];
      if (l.out_half) { reinterpret_cast<__half*>(l.C)[r * l.ldc + c] = __float2half_rn(fminf(fmaxf(x, -l.half_max), l.half_max)); continue; }
      if (l.round_out) x = round_tf32(x);
      l.C[r * l.ldc + c] = x;
    }
  }
}

// segmax for the SIMT path (test hook / tiny shapes): out[g, c] = max_r relu(C[g*32 + r, c])
__global__ void segmax32_kernel(const float* C, long ldc, float* out, long ldo, const float* side, long lds, int groups, int N) {
  const int g = blockIdx.x;
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    float m = 0.f;
    for (int r = 0; r < 32; ++r) m = fmaxf(m, C[(static_cast<long>(g) * 32 + r) * ldc + c]);
    if (side) m = fmaxf(m, side[static_cast<long>(g) * lds + c]);
    out[static_cast<long>(g) * ldo + c] = m;
  }
}

cudaError_t linear_simt(const Linear& l, cudaStream_t st, Launches* lc) {
  if (l.M <= 0) return cudaSuccess;
  if (l.segmax) return cudaErrorInvalidValue;  // callers run segmax32 separately (see api.cu debug hook)
  if (lc) lc->n++;
  dim3 grid(static_cast<unsigned>((l.M + SB_M - 1) / SB_M), (l.N + SB_N - 1) / SB_N);
  linear_simt_kernel<<<grid, 256, 0, st>>>(l);
  return cudaGetLastError();
}

cudaError_t segmax32(const float* C, long ldc, float* out, long ldo, const float* side, long lds, int groups, int N,
                     cudaStream_t st, Launches* lc) {
  if (groups <= 0) return cudaSuccess;
  if (lc) lc->n++;
  segmax32_kernel<<<groups, 256, 0, st>>>(C, ldc, out, ldo, side, lds, groups, N);
  return cudaGetLastError();
}

}  // namespace t2l
