// Host half of the RNG contract (MYFM_RNG_MT19937).
//
// "Same seed" in the reference means: one std::mt19937 consumed by libstdc++ distributions that
// are constructed fresh for almost every draw (include/myfm/FMTrainer.hpp:122-125,143,165;
// include/myfm/FM.hpp:34-45 is the one persistent object).  Two facts make this cheap to honour
// without dragging the sampler back to the host:
//   * a Gaussian draw is  first/quad + N(0,1)/sqrt(quad)  and a Gamma draw is
//     Gamma(shape, 1) * scale  bit for bit (libstdc++ multiplies by the scale last), so the
//     data-dependent part is applied on the device to a STANDARDISED variate;
//   * how many engine words a standardised variate consumes depends only on the stream and on
//     the (data-independent) Gamma shapes, so for regression the whole variate sequence of a
//     sweep can be produced ahead of the device.
// This file produces those standardised variates, in the reference's consumption order.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <random>
#include <vector>

namespace myfm {

template <typename Real> struct MtStream {
  std::mt19937 gen;
  explicit MtStream(int seed) : gen(seed) {}

  // FMTrainer.hpp:122-125 — fresh distribution object: the second polar variate is dropped.
  Real normal() { return std::normal_distribution<Real>(0, 1)(gen); }
  // FMTrainer.hpp:143,165 — fresh object; result scales exactly with the distribution's beta.
  Real gamma(Real shape) { return std::gamma_distribution<Real>(shape, Real(1))(gen); }

  // FM.hpp:34-45 — V column by column, then w, then w0 from ONE normal_distribution.
  void init_weights(Real *V_colmajor, size_t n_V, Real *w, size_t n_w, Real *w0, Real init_std) {
    std::normal_distribution<Real> nd;
    for (size_t i = 0; i < n_V; i++)
      V_colmajor[i] = nd(gen) * init_std;
    for (size_t i = 0; i < n_w; i++)
      w[i] = nd(gen) * init_std;
    *w0 = nd(gen) * init_std;
  }

  // Truncated standard normal, Z > a (Robert 1995, Prop. 2.3); include/myfm/util.hpp:15-37.
  Real tn_left(Real a) {
    if (a < 0) {
      std::normal_distribution<Real> dist(0, 1);
      for (;;) {
        Real z = dist(gen);
        if (z > a)
          return z;
      }
    }
    Real alpha_star = (a + std::sqrt(a * a + 4)) / 2;
    std::uniform_real_distribution<Real> dist(0, 1);
    for (;;) {
      Real z = -std::log(dist(gen)) / alpha_star + a;
      Real rho = std::exp(-(z - alpha_star) * (z - alpha_star) / 2);
      Real u = dist(gen);
      if (u < rho)
        return z;
    }
  }
  // util.hpp:68-71
  Real tn_right(Real b) { return -tn_left(-b); }
  // util.hpp:39-60
  Real tn_twoside(Real a, Real b) {
    std::uniform_real_distribution<Real> proposal(a, b);
    std::uniform_real_distribution<Real> acceptance(0, 1);
    for (;;) {
      Real z = proposal(gen);
      Real rho;
      if (a <= Real(0) && b >= Real(0))
        rho = std::exp(-z * z / 2);
      else if (b < Real(0))
        rho = std::exp((b * b - z * z) / 2);
      else
        rho = std::exp((a * a - z * z) / 2);
      Real u = acceptance(gen);
      if (u < rho)
        return z;
    }
  }
  // util.hpp:61-78 (scaled forms)
  Real tn_left(Real mean, Real sd, Real lower) { return mean + sd * tn_left((lower - mean) / sd); }
  Real tn_right(Real mean, Real sd, Real upper) { return mean + sd * tn_right((upper - mean) / sd); }
};

// Where each standardised variate of one update_all sweep sits in the device buffer
// (reference order: BaseFMTrainer.hpp:135-152).  -1 = not drawn.
struct SweepLayout {
  int64_t g_alpha = -1; // 1 gamma (regression only)
  int64_t z_w0 = -1;    // 1 normal (fit_w0)
  int64_t g_lw = 0;     // G gammas
  int64_t z_mw = 0;     // G normals
  int64_t z_w = -1;     // dim_all normals (fit_linear)
  int64_t g_lV = 0;     // K*G gammas, [r*G + g]
  int64_t z_mV = 0;     // K*G normals, [r*G + g]
  int64_t z_V = 0;      // K*dim_all normals, [r*dim_all + j]
  int64_t total = 0;

  static SweepLayout make(bool regression, bool fit_w0, bool fit_linear, int64_t G, int64_t K,
                          int64_t dim_all) {
    SweepLayout L;
    int64_t o = 0;
    if (regression)
      L.g_alpha = o++;
    if (fit_w0)
      L.z_w0 = o++;
    L.g_lw = o, o += G;
    L.z_mw = o, o += G;
    if (fit_linear)
      L.z_w = o, o += dim_all;
    L.g_lV = o, o += K * G;
    L.z_mV = o, o += K * G;
    L.z_V = o, o += K * dim_all;
    L.total = o;
    return L;
  }
};

} // namespace myfm
