// LibFM-style per-iteration evaluation on the device (reference src/myfm/utils/callbacks/libfm.py:57-262:
// RegressionCallback / ClassificationCallback / OrderedProbitCallback).  Every sweep the callback
// scores a held-out set with the CURRENT sample, keeps two running sums of the predictions (all
// sweeps, and all but the first five) and reports metrics of the running means and of this sweep's
// prediction.  Here the forward pass, the link, both running sums and every metric run on the
// device against a test matrix that stays in HBM; nine scalars come back.  Sums and metrics are
// accumulated in double (the reference's numpy arrays are float64).
#pragma once

#include "kernels.cuh"

namespace myfm {

constexpr int EVAL_BLOCKS = 592, EVAL_THREADS = 256, EVAL_TERMS = 9;

struct EvalArgs {
  int n, width;          // rows; columns of a prediction (1, or the number of classes)
  int task;              // MYFM_TASK_*
  int iteration;         // sweep index i of the callback
  int n_samples;         // sweeps accumulated including this one
  int burn_in;           // 5: sweeps left out of the second running sum
  double clip_min, clip_max; // regression: NaN = no clipping
  double eps;            // classification: clip of the running means; ordered: floor of the picked probability; < 0: none
  double *sum, *late;    // [n x width] running sums
  const double *y;       // [n] targets (classification: 0 / 1, ordered: class index)
  const double *cutpoints; // ordered: [width - 1]
  double *partial;       // [EVAL_BLOCKS x EVAL_TERMS]
};

__device__ __forceinline__ double eval_clip(double v, double lo, double hi) {
  if (lo == lo && v <= lo) // lo == lo: not NaN
    v = lo;
  if (hi == hi && v >= hi)
    v = hi;
  return v;
}
__device__ __forceinline__ double eval_std_cdf(double x) { return (1.0 + erf(x * 0.70710678118654757)) / 2.0; } // base.py:41-43

// Terms (sums over rows; the host divides / takes roots):
//   regression      0 se(mean) 1 se(this) 2 se(late)
//   classification  0 ll(mean) 1 ll(this) 2 ll(late) 3 hits(mean) 4 hits(this) 5 hits(late)
//   ordered         0..5 as classification, 6 se(mean) 7 se(this) 8 se(late)   (expected class vs label)
template <typename Real>
__global__ void __launch_bounds__(EVAL_THREADS) k_eval(EvalArgs a, const Real *__restrict__ score) {
  __shared__ double s_red[EVAL_THREADS / 32][EVAL_TERMS];
  double t[EVAL_TERMS];
#pragma unroll
  for (int k = 0; k < EVAL_TERMS; k++)
    t[k] = 0;
  const bool has_late = a.iteration >= a.burn_in;
  const double n_late = static_cast<double>(a.iteration + 1 - a.burn_in);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
    const double s = static_cast<double>(score[i]), y = a.y[i];
    if (a.task == MYFM_TASK_REGRESSION) {
      const double sum = a.sum[i] + s;
      a.sum[i] = sum;
      const double mean = eval_clip(sum / a.n_samples, a.clip_min, a.clip_max);
      t[0] += (y - mean) * (y - mean);
      t[1] += (y - s) * (y - s);
      if (has_late) {
        const double ls = a.late[i] + s;
        a.late[i] = ls;
        const double lm = eval_clip(ls / n_late, a.clip_min, a.clip_max);
        t[2] += (y - lm) * (y - lm);
      }
    } else if (a.task == MYFM_TASK_CLASSIFICATION) {
      const double p = eval_std_cdf(s);
      const double sum = a.sum[i] + p;
      a.sum[i] = sum;
      const double lo = a.eps >= 0 ? a.eps : nan(""), hi = a.eps >= 0 ? 1 - a.eps : nan("");
      const double mean = eval_clip(sum / a.n_samples, lo, hi);
      const bool pos = y == 1.0;
      t[0] -= pos ? log(mean) : log(1 - mean);
      t[1] -= pos ? log(p) : log(1 - p);
      t[3] += pos == (mean >= 0.5);
      t[4] += pos == (p >= 0.5);
      if (has_late) {
        const double ls = a.late[i] + p;
        a.late[i] = ls;
        const double lm = eval_clip(ls / n_late, lo, hi);
        t[2] -= pos ? log(lm) : log(1 - lm);
        t[5] += pos == (lm >= 0.5);
      }
    } else { // ordered probit: FM.hpp:137-162, then libfm.py:213-262
      const int label = static_cast<int>(y);
      double prev = 0;
      double best[3] = {-1, -1, -1}, picked[3] = {0, 0, 0}, expect[3] = {0, 0, 0};
      int arg[3] = {0, 0, 0};
      for (int c = 0; c < a.width; c++) {
        double cdf = 1;
        if (c + 1 < a.width)
          cdf = eval_std_cdf(a.cutpoints[c] - s);
        const double p = cdf - prev;
        prev = cdf;
        const size_t at = static_cast<size_t>(i) * a.width + c;
        const double sum = a.sum[at] + p;
        a.sum[at] = sum;
        double v[3] = {sum / a.n_samples, p, 0};
        if (has_late) {
          const double ls = a.late[at] + p;
          a.late[at] = ls;
          v[2] = ls / n_late;
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          if (v[k] > best[k])
            best[k] = v[k], arg[k] = c; // first maximum, like numpy argmax
          if (c == label)
            picked[k] = v[k];
          expect[k] += c * v[k];
        }
      }
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (k == 2 && !has_late)
          continue;
        const double ps = (a.eps >= 0 && picked[k] <= a.eps) ? a.eps : picked[k];
        t[k] -= log(ps);
        t[3 + k] += arg[k] == label;
        t[6 + k] += (label - expect[k]) * (label - expect[k]);
      }
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < EVAL_TERMS; k++) {
    t[k] = warp_sum(t[k]);
    if (lane == 0)
      s_red[wid][k] = t[k];
  }
  __syncthreads();
  if (threadIdx.x < EVAL_TERMS) {
    double v = 0;
    for (int w = 0; w < EVAL_THREADS / 32; w++)
      v += s_red[w][threadIdx.x];
    a.partial[static_cast<size_t>(blockIdx.x) * EVAL_TERMS + threadIdx.x] = v;
  }
}

// block partials in block order (deterministic)
__global__ void __launch_bounds__(32) k_eval_finish(int n_blocks, const double *__restrict__ partial, double *__restrict__ out) {
  if (threadIdx.x < EVAL_TERMS) {
    double v = 0;
    for (int b = 0; b < n_blocks; b++)
      v += partial[static_cast<size_t>(b) * EVAL_TERMS + threadIdx.x];
    out[threadIdx.x] = v;
  }
}

} // namespace myfm
