// Accuracy bookkeeping of the retrieval loop on the device:
//   training/coarse.py:131-150    top-k hit (`target_cell_id in retrieved_cell_ids[0:k]`) and close-by accuracy
//                                 (`np.any(dists[0:k] <= cell_size / 2)`, dists to the retrieved cells' centres)
//   evaluation/utils.py:31-54     calc_sample_accuracies: distance of the pose to the predicted in-cell position of each
//                                 retrieved cell, +inf across scenes, `np.min(dists[0:k]) <= t` per (k, threshold)
// The reference walks the queries in a Python loop (O(nq k) numpy calls); here one thread owns one query.
// Distances are evaluated as numpy does for a 2-vector norm: sqrt(dx*dx + dy*dy) in float64, one rounding per
// operation (no FMA contraction), so the thresholded comparisons agree bit for bit.
#include <math.h>

#include "ops.h"

namespace t2l {

__global__ void __launch_bounds__(128) topk_accuracy_kernel(TopkAccuracy a) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.nq) return;
  const double qx = a.query_xy[2 * q], qy = a.query_xy[2 * q + 1];
  const long target = a.target_row ? a.target_row[q] : -2;
  const int qs = a.query_scene ? a.query_scene[q] : 0;
  int first_hit = a.k;         // position of the target row in the list
  double run_min = INFINITY;   // min distance over the list prefix
  int t = 0;                   // next top_k entry to close (top_k ascending)
  for (int j = 0; j <= a.k; ++j) {
    while (t < a.n_top && a.top_k[t] == j) {  // prefix [0, j) complete: emit the row for k = j
      if (a.hit) a.hit[static_cast<long>(q) * a.n_top + t] = first_hit < j;
      if (a.within)
        for (int h = 0; h < a.n_thr; ++h) a.within[(static_cast<long>(q) * a.n_top + t) * a.n_thr + h] = run_min <= a.threshs[h];
      ++t;
    }
    if (j == a.k) break;
    const long row = a.idx[static_cast<long>(q) * a.k + j];
    double d = INFINITY;
    if (row >= 0) {
      const double dx = __dsub_rn(qx, a.cell_xy[2 * row]), dy = __dsub_rn(qy, a.cell_xy[2 * row + 1]);
      d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
      if (a.cell_scene && a.cell_scene[row] != qs) d = INFINITY;  // evaluation/utils.py:48-50
      if (row == target && first_hit == a.k) first_hit = j;
    }
    if (a.dists) a.dists[static_cast<long>(q) * a.k + j] = d;
    run_min = fmin(run_min, d);
  }
  // top_k entries beyond the list length see the whole list (numpy slicing semantics)
  for (; t < a.n_top; ++t) {
    if (a.hit) a.hit[static_cast<long>(q) * a.n_top + t] = first_hit < a.k;
    if (a.within)
      for (int h = 0; h < a.n_thr; ++h) a.within[(static_cast<long>(q) * a.n_top + t) * a.n_thr + h] = run_min <= a.threshs[h];
  }
}

cudaError_t topk_accuracy(const TopkAccuracy& a, cudaStream_t st, Launches* lc) {
  if (a.nq <= 0) return cudaSuccess;
  if (lc) lc->n++;
  topk_accuracy_kernel<<<(a.nq + 127) / 128, 128, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace t2l
