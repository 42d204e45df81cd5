// Latent-variable draws of the classification / ordered-probit tasks on the device
// (MYFM_RNG_PHILOX).
//
// The reference draws one truncated normal per training row per sweep from its single mt19937
// (FMTrainer.hpp:498-521, OProbitSampler.hpp:238-272, util.hpp:15-78); the number of engine words
// a draw consumes depends on the data, so that stream can only be reproduced row by row (the
// engine does exactly that on the host in MYFM_RNG_MT19937 mode).  Here every row owns a
// counter-based Philox4x32-10 stream keyed by (seed; global row, sweep, attempt), so the N draws
// of a sweep are one data-parallel kernel.  The samplers are the reference's (Robert 1995: naive
// rejection for a one-sided bound below the mean, exponential proposal above it, uniform proposal
// for the two-sided case); the chain is statistically equivalent, not seed-identical.
#pragma once

#include "kernels.cuh"

namespace myfm {

struct Philox { // Philox4x32-10 (Salmon et al. 2011), restated from the published algorithm
  uint32_t key0, key1;
  uint32_t c0, c1, c2, c3; // c0: attempt counter, c1: row, c2: sweep, c3: stream tag

  __device__ __forceinline__ void round(uint32_t &x0, uint32_t &x1, uint32_t &x2, uint32_t &x3, uint32_t k0,
                                        uint32_t k1) const {
    const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
    x0 = hi1 ^ x1 ^ k0, x1 = lo1, x2 = hi0 ^ x3 ^ k1, x3 = lo0;
  }
  // four fresh words; advances the attempt counter
  __device__ __forceinline__ uint4 next() {
    uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, k0 = key0, k1 = key1;
#pragma unroll
    for (int r = 0; r < 10; r++) {
      round(x0, x1, x2, x3, k0, k1);
      k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
    }
    c0++;
    return make_uint4(x0, x1, x2, x3);
  }
};

// uniform in (0, 1): never 0 (log) and never 1
template <typename Real> __device__ __forceinline__ Real philox_uniform(uint32_t hi, uint32_t lo);
template <> __device__ __forceinline__ float philox_uniform<float>(uint32_t hi, uint32_t) {
  return (static_cast<float>(hi >> 8) + 0.5f) * 5.9604644775390625e-8f; // 2^-24
}
template <> __device__ __forceinline__ double philox_uniform<double>(uint32_t hi, uint32_t lo) {
  const unsigned long long m = (static_cast<unsigned long long>(hi) << 21) ^ (lo >> 11); // 53 bits
  return (static_cast<double>(m) + 0.5) * 1.1102230246251565e-16; // 2^-53
}

__device__ __forceinline__ float cos_2pi(float u) { return cospif(2.0f * u); }
__device__ __forceinline__ double cos_2pi(double u) { return cospi(2.0 * u); }

template <typename Real> struct TruncatedNormal {
  Philox rng;

  __device__ __forceinline__ Real normal(const uint4 &w) { // Box-Muller on (w.x, w.y), (w.z, w.w)
    const Real u1 = philox_uniform<Real>(w.x, w.y), u2 = philox_uniform<Real>(w.z, w.w);
    return sqrt(Real(-2) * log(u1)) * cos_2pi(u2);
  }
  // Z > a  (util.hpp:15-37)
  __device__ Real left(Real a) {
    if (a < 0) {
      for (int it = 0; it < 4096; it++) {
        const Real z = normal(rng.next());
        if (z > a)
          return z;
      }
      return a < Real(-1) ? Real(0) : a + Real(1e-3); // unreachable for a < 0 (p(accept) >= 1/2)
    }
    const Real alpha_star = (a + sqrt(a * a + 4)) / 2;
    for (int it = 0; it < 4096; it++) {
      const uint4 w = rng.next();
      const Real z = -log(philox_uniform<Real>(w.x, w.y)) / alpha_star + a;
      const Real rho = exp(-(z - alpha_star) * (z - alpha_star) / 2);
      if (philox_uniform<Real>(w.z, w.w) < rho)
        return z;
    }
    return a + 1 / alpha_star; // mean of the proposal; acceptance is >= 0.76, never reached
  }
  // Z < b  (util.hpp:68-71)
  __device__ Real right(Real b) { return -left(-b); }
  // a < Z < b  (util.hpp:39-60)
  __device__ Real twoside(Real a, Real b) {
    for (int it = 0; it < 65536; it++) {
      const uint4 w = rng.next();
      const Real z = a + (b - a) * philox_uniform<Real>(w.x, w.y);
      Real rho;
      if (a <= Real(0) && b >= Real(0))
        rho = exp(-z * z / 2);
      else if (b < Real(0))
        rho = exp((b * b - z * z) / 2);
      else
        rho = exp((a * a - z * z) / 2);
      if (philox_uniform<Real>(w.z, w.w) < rho)
        return z;
    }
    return (a + b) / 2;
  }
};

template <typename Real>
__device__ __forceinline__ TruncatedNormal<Real> latent_stream(uint64_t seed, uint32_t row, uint32_t sweep) {
  TruncatedNormal<Real> tn;
  tn.rng.key0 = static_cast<uint32_t>(seed), tn.rng.key1 = static_cast<uint32_t>(seed >> 32);
  tn.rng.c0 = 0, tn.rng.c1 = row, tn.rng.c2 = sweep, tn.rng.c3 = 0x6c61746eu; // "latn"
  return tn;
}

// CLASSIFICATION, FMTrainer.hpp:498-512: e_i holds the score; e_i -= TN(score, 1) truncated to
// z > 0 when y_i > 0, to z < 0 otherwise.  orig_row: caller's index of device row i (the stream key).
template <typename Real>
__global__ void __launch_bounds__(256)
    k_latent_classification(int64_t n, Pair<Real> *__restrict__ eq, const Real *__restrict__ y,
                            const int *__restrict__ orig_row, uint64_t seed, uint32_t sweep) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n)
    return;
  TruncatedNormal<Real> tn = latent_stream<Real>(seed, static_cast<uint32_t>(orig_row[i]), sweep);
  const Real pred = eq[i].x;
  // mean + sd * tn((bound - mean) / sd) with sd = 1, bound = 0  (util.hpp:61-78)
  const Real z = y[i] > 0 ? pred + tn.left(-pred) : pred + tn.right(-pred);
  eq[i].x = pred - z;
}

// ORDERED, OProbitSampler.hpp:238-272 (sample_z_given_cutpoint): the group's rows arrive grouped by
// label (class_ptr / class_rows); z_i ~ N(score_i, 1) truncated to the label's interval, e_i = score_i - z_i.
template <typename Real>
__global__ void __launch_bounds__(256)
    k_latent_ordered(int n_class, const int *__restrict__ class_ptr, const int *__restrict__ class_rows,
                     const Real *__restrict__ gamma, Pair<Real> *__restrict__ eq,
                     const int *__restrict__ orig_row, uint64_t seed, uint32_t sweep) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= class_ptr[n_class])
    return;
  int label = 0;
  while (label + 1 < n_class && t >= class_ptr[label + 1])
    label++;
  const int i = class_rows[t];
  TruncatedNormal<Real> tn = latent_stream<Real>(seed, static_cast<uint32_t>(orig_row[i]), sweep);
  const Real pred = eq[i].x;
  Real z;
  if (label == 0)
    z = tn.right(gamma[0] - pred) + pred;
  else if (label == n_class - 1)
    z = tn.left(gamma[n_class - 2] - pred) + pred;
  else
    z = tn.twoside(gamma[label - 1] - pred, gamma[label] - pred) + pred;
  eq[i].x = pred - z;
}

} // namespace myfm
