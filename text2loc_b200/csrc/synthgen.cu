// Counter-based synthetic cells generated on the device (BASELINE configs[3]: 1M cells x 16 objects x 256 points = 98 GB of
// points that must never exist at once).  Bench / test input tooling, not part of the
// reference's path: the distributions follow synth.make_packed_cells (object centre ~U[0,1]^2 x U[0,0.2], extents
// ~U[0.02,0.3] x (1,1,0.3), raw count ~U{30..5000}, points uniform in the box, rgb = object colour + noise clipped to
// [0,1], objects with fewer raw points than samples repeat points).  Every value is a pure function of
// (seed, global object index, slot): any chunking of the cell range produces the same bytes, and oracle/synthgen.py
// restates it bit for bit in numpy (integer hash, then fp32 multiplies and adds rounded one at a time: no FMA).
#include "ops.h"
#include "common.cuh"

namespace t2l {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// uniform in [0, 1) with 24 random bits (exactly representable in fp32)
T2L_DEVICE float synth_u(uint64_t seed, uint64_t obj, uint32_t slot) {
  const uint64_t h = splitmix64(seed ^ splitmix64(obj * 4096ull + slot));
  return static_cast<float>(static_cast<uint32_t>(h >> 40)) * 5.9604644775390625e-8f;  // 2^-24
}

constexpr uint32_t kSlotPoint0 = 16;  // slots 0..9: object parameters; 16 + 8 k + c: point k, component c; c = 6: sample index

__global__ void __launch_bounds__(256) synth_cells_kernel(uint64_t seed, long first_obj, long n_obj, float* __restrict__ pts, float* __restrict__ meta) {
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long o = idx >> 8;  // 256 points per object
  const int k = static_cast<int>(idx & 255);
  if (o >= n_obj) return;
  const uint64_t g = static_cast<uint64_t>(first_obj + o);
  float centre[3], extent[3], colour[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    centre[c] = __fmul_rn(synth_u(seed, g, c), c == 2 ? 0.2f : 1.0f);
    extent[c] = __fmul_rn(__fadd_rn(0.02f, __fmul_rn(synth_u(seed, g, 3 + c), 0.28f)), c == 2 ? 0.3f : 1.0f);
    colour[c] = synth_u(seed, g, 6 + c);
  }
  const int n_raw = 30 + static_cast<int>(__fmul_rn(synth_u(seed, g, 9), 4971.0f));  // 30 .. 5000
  // sampling with replacement from fewer raw points than samples: point k shows raw point floor(u * n_raw)
  const int s = n_raw < 256 ? static_cast<int>(__fmul_rn(synth_u(seed, g, kSlotPoint0 + 8 * k + 6), static_cast<float>(n_raw))) : k;
  float* p = pts + (o * 256 + k) * 6;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float u = __fadd_rn(synth_u(seed, g, kSlotPoint0 + 8 * s + c), -0.5f);
    p[c] = __fadd_rn(centre[c], __fmul_rn(u, extent[c]));
    const float nz = __fmul_rn(__fadd_rn(synth_u(seed, g, kSlotPoint0 + 8 * s + 3 + c), -0.5f), 0.17320508f);  // uniform, sigma 0.05
    p[3 + c] = fminf(fmaxf(__fadd_rn(colour[c], nz), 0.f), 1.f);
  }
  if (k == 0) {
    float* m = meta + o * 7;
    m[0] = colour[0]; m[1] = colour[1]; m[2] = colour[2];
    m[3] = centre[0]; m[4] = centre[1]; m[5] = centre[2];
    m[6] = static_cast<float>(n_raw);
  }
}

cudaError_t synth_cells(uint64_t seed, long first_obj, long n_obj, float* pts, float* meta, cudaStream_t st, Launches* lc) {
  if (n_obj <= 0) return cudaSuccess;
  if (lc) lc->n++;
  const long threads = n_obj * 256;
  synth_cells_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, st>>>(seed, first_obj, n_obj, pts, meta);
  return cudaGetLastError();
}

}  // namespace t2l
