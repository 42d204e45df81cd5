// Ordered-probit cut-point sampling (reference include/myfm/OProbitSampler.hpp).
//
// Split of the work: everything that touches all training rows — the log-posterior of the
// cut-points with its gradient and tri-diagonal Hessian (OProbitSampler.hpp:389-463 over
// safe_lcdf / safe_lccdf / safe_ldiff :111-236) — is a device reduction over rows GROUPED BY
// LABEL, so that every thread of a launch evaluates the same branch family and accumulates the
// same six sums (loss, d/dx, d/dy, d2/dx2, d2/dy2, d2/dxdy).  The (K-1)-dimensional algebra around
// it — the alpha <-> gamma reparametrisation, damped Newton mode search (:289-357), the
// multivariate-t Metropolis-Hastings proposal (:51-72, :359-387) — stays on the host, where the
// mt19937 stream lives for this task.
#pragma once

#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <functional>
#include <random>
#include <stdexcept>
#include <vector>

namespace myfm {

constexpr int OP_TERMS = 6;            // loss, dx, dy, hxx, hyy, hxy
constexpr int OP_BLOCKS_PER_CLASS = 64;
constexpr int OP_THREADS = 256;

#ifdef __CUDACC__
template <typename Real> struct OpConst {
  static constexpr double SQRT2 = 1.4142135623730951;
  static constexpr double SQRT2PI = 1.4142135623730951 * 1.7724538509055159;
  static constexpr double PI = 3.141592653589793;
};
// erfcx / erf of the reference come from Faddeeva (double in, double out, result narrowed to Real)
template <typename Real> __device__ __forceinline__ Real op_erfcx(Real x) {
  return static_cast<Real>(erfcx(static_cast<double>(x)));
}
template <typename Real> __device__ __forceinline__ Real op_erf(Real x) {
  return static_cast<Real>(erf(static_cast<double>(x)));
}

// log Phi(x) and derivatives (label 0; OProbitSampler.hpp:183-209): t[0] loss, t[1] d, t[3] d2
template <typename Real> __device__ __forceinline__ void op_lcdf(Real x, Real *t) {
  const Real s2 = static_cast<Real>(OpConst<Real>::SQRT2), s2pi = static_cast<Real>(OpConst<Real>::SQRT2PI),
             pi = static_cast<Real>(OpConst<Real>::PI);
  if (x > 1) {
    const Real ef = exp(-x * x / 2);
    const Real den = 1 + op_erf<Real>(x / s2);
    t[1] += (2 / s2pi) * ef / den;
    t[0] += log(den / 2);
    t[3] += -(s2pi * x * den * ef + 2 * ef * ef) / pi / den / den;
  } else {
    const Real den = op_erfcx<Real>(-x / s2);
    t[1] += (2 / s2pi) / den;
    t[0] -= x * x / 2;
    t[0] += log(den / 2);
    t[3] += -(s2pi * x * den + 2) / pi / den / den;
  }
}
// log (1 - Phi(x)) (top label; :211-236): t[0] loss, t[1] d (enters with a minus), t[3] d2
template <typename Real> __device__ __forceinline__ void op_lccdf(Real x, Real *t) {
  const Real s2 = static_cast<Real>(OpConst<Real>::SQRT2), s2pi = static_cast<Real>(OpConst<Real>::SQRT2PI),
             pi = static_cast<Real>(OpConst<Real>::PI);
  if (x > -1) {
    const Real den = op_erfcx<Real>(x / s2);
    t[1] -= (2 / s2pi) / den;
    t[0] += log(den / 2);
    t[0] -= x * x / 2;
    t[3] += (s2pi * x * den - 2) / den / den / pi;
  } else {
    const Real den = 1 - op_erf<Real>(x / s2);
    const Real ef = exp(-(x * x) / 2);
    t[1] -= (2 / s2pi) * exp(-x * x / 2) / den;
    t[0] += log(den / 2);
    t[3] += -(-s2pi * x * den * ef + 2 * ef * ef) / pi / den / den;
  }
}
// log (Phi(x) - Phi(y)), x > y (inner labels; :111-181): overflow-safe in three sign cases
template <typename Real> __device__ __forceinline__ void op_ldiff(Real x, Real y, Real *t) {
  const Real s2 = static_cast<Real>(OpConst<Real>::SQRT2), s2pi = static_cast<Real>(OpConst<Real>::SQRT2PI),
             pi = static_cast<Real>(OpConst<Real>::PI);
  if (y > 0) {
    const Real ef = exp((y * y - x * x) / 2);
    const Real den = op_erfcx<Real>(y / s2) - ef * op_erfcx<Real>(x / s2);
    t[0] -= y * y / 2;
    t[0] += log(den / 2);
    t[1] += (2 / s2pi) * ef / den;
    t[2] -= (2 / s2pi) / den;
    t[3] += -(s2pi * x * den * exp((y * y - x * x) / 2) + 2 * exp(y * y - x * x)) / den / den / pi;
    t[4] += (s2pi * y * den - 2) / den / den / pi;
    t[5] += 2 * exp((y * y - x * x) / 2) / pi / den / den;
  } else if (x < 0) {
    t[0] -= x * x / 2;
    const Real ef = exp((x * x - y * y) / 2);
    const Real den = op_erfcx<Real>(-x / s2) - ef * op_erfcx<Real>(-y / s2);
    t[0] += log(den / 2);
    t[1] += (2 / s2pi) / den;
    t[2] -= (2 / s2pi) * ef / den;
    t[3] += -(s2pi * x * den + 2) / pi / den / den;
    t[4] += (s2pi * y * ef * den - 2 * (ef * ef)) / pi / den / den;
    t[5] += 2 * ef / pi / den / den;
  } else {
    const Real den = op_erf<Real>(x / s2) - op_erf<Real>(y / s2);
    const Real exx = exp(-x * x / 2), eyy = exp(-y * y / 2);
    t[1] += 2 * exx / den / s2pi;
    t[2] -= 2 * eyy / den / s2pi;
    t[0] += log(den / 2);
    t[3] += -(s2pi * x * den * exx + 2 * exx * exx) / pi / den / den;
    t[4] += -(-s2pi * y * den * eyy + 2 * eyy * eyy) / pi / den / den;
    t[5] += 2 * exx * eyy / pi / den / den;
  }
}

// Block (c, b) of the grid reduces a strided share of the rows with label c.
// score[row * stride] is the current FM score of a training row (device row order).
template <typename Real>
__global__ void __launch_bounds__(OP_THREADS)
    k_oprobit_terms(int n_class, const int *__restrict__ class_ptr, const int *__restrict__ class_rows,
                    const Real *__restrict__ score, int stride, const Real *__restrict__ gamma,
                    Real *__restrict__ partial) {
  __shared__ Real scratch[32];
  const int c = blockIdx.x / OP_BLOCKS_PER_CLASS, b = blockIdx.x % OP_BLOCKS_PER_CLASS;
  const int lo = class_ptr[c], hi = class_ptr[c + 1];
  Real t[OP_TERMS] = {0, 0, 0, 0, 0, 0};
  const Real g_hi = c < n_class - 1 ? gamma[c] : Real(0), g_lo = c > 0 ? gamma[c - 1] : Real(0);
  for (int p = lo + b * OP_THREADS + threadIdx.x; p < hi; p += OP_BLOCKS_PER_CLASS * OP_THREADS) {
    const Real s = score[static_cast<size_t>(class_rows[p]) * stride];
    if (c == 0)
      op_lcdf<Real>(g_hi - s, t);
    else if (c == n_class - 1)
      op_lccdf<Real>(g_lo - s, t);
    else
      op_ldiff<Real>(g_hi - s, g_lo - s, t);
  }
#pragma unroll
  for (int k = 0; k < OP_TERMS; k++) {
    Real v = block_sum(t[k], scratch);
    if (threadIdx.x == 0)
      partial[static_cast<size_t>(blockIdx.x) * OP_TERMS + k] = v;
  }
}
// sums[c * 6 + k] = sum over the class' blocks, in block order
template <typename Real>
__global__ void k_oprobit_finish(int n_class, const Real *__restrict__ partial, Real *__restrict__ sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_class * OP_TERMS)
    return;
  const int c = i / OP_TERMS, k = i % OP_TERMS;
  Real acc = 0;
  for (int b = 0; b < OP_BLOCKS_PER_CLASS; b++)
    acc += partial[(static_cast<size_t>(c) * OP_BLOCKS_PER_CLASS + b) * OP_TERMS + k];
  sums[i] = acc;
}
#endif // __CUDACC__

// ------------------------------------------------------------------------------------------------
// Host side: the sampler of one cut-point group.
// ------------------------------------------------------------------------------------------------
template <typename Real> struct SmallMat { // square, column-major
  int n = 0;
  std::vector<Real> a;
  explicit SmallMat(int n_ = 0) : n(n_), a(static_cast<size_t>(n_) * n_, Real(0)) {}
  Real &operator()(int i, int j) { return a[i + static_cast<size_t>(n) * j]; }
  Real operator()(int i, int j) const { return a[i + static_cast<size_t>(n) * j]; }
};

template <typename Real> struct CutpointSampler {
  using Vec = std::vector<Real>;
  using Mat = SmallMat<Real>;
  // class_sums(gamma) -> [n_class * 6] label-wise sums of the row terms at these cut-points
  using RowSums = std::function<void(const Vec &gamma, Vec &sums)>;

  int K;       // number of classes
  Real reg, nu;
  RowSums row_sums;
  Vec alpha_now, gamma_now;
  Mat H;
  int64_t accept_count = 0;

  CutpointSampler(int n_class, Real reg_, Real nu_, RowSums f)
      : K(n_class), reg(reg_), nu(nu_), row_sums(std::move(f)), alpha_now(n_class - 1, Real(0)),
        gamma_now(n_class - 1, Real(0)), H(n_class - 1) {
    to_gamma(alpha_now, gamma_now);
  }

  // gamma_0 = alpha_0, gamma_i = gamma_{i-1} + exp(alpha_i)   (OProbitSampler.hpp:95-101)
  static void to_gamma(const Vec &alpha, Vec &gamma) {
    if (alpha.empty())
      return;
    gamma[0] = alpha[0];
    for (size_t i = 1; i < alpha.size(); i++)
      gamma[i] = gamma[i - 1] + std::exp(alpha[i]);
  }

  struct NanError : std::runtime_error { // "H has NaN" / "dalpha has NaN" of the reference
    using std::runtime_error::runtime_error;
  };

  // Negative log posterior of alpha, its gradient, optionally its Hessian (:389-463).
  Real objective(const Vec &alpha, Vec &grad, Mat *Hout) {
    const int n = K - 1;
    Vec gamma(n), sums;
    to_gamma(alpha, gamma);
    row_sums(gamma, sums);
    // assemble d ll / d gamma and the tri-diagonal d2 ll / d gamma2 from the label-wise sums
    Real ll = 0;
    Vec dg(n, Real(0));
    Mat Hg(n);
    for (int c = 0; c < K; c++) {
      const Real *t = sums.data() + static_cast<size_t>(c) * OP_TERMS;
      ll += t[0];
      if (c == 0) {
        dg[0] += t[1];
        Hg(0, 0) += t[3];
      } else if (c == K - 1) {
        dg[K - 2] += t[1];
        Hg(K - 2, K - 2) += t[3];
      } else {
        dg[c] += t[1], dg[c - 1] += t[2];
        Hg(c, c) += t[3], Hg(c - 1, c - 1) += t[4];
        Hg(c, c - 1) += t[5], Hg(c - 1, c) += t[5];
      }
    }
    // J(i, j) = d gamma_j / d alpha_i   (:74-93)
    Mat J(n);
    Vec ea(n);
    for (int i = 0; i < n; i++)
      ea[i] = std::exp(alpha[i]);
    for (int j = 0; j < n; j++)
      J(0, j) = 1;
    for (int i = 1; i < n; i++)
      for (int j = i; j < n; j++)
        J(i, j) = ea[i];
    if (Hout) {
      Mat T(n), R(n);
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
          Real acc = 0;
          for (int k = 0; k < n; k++)
            acc += J(i, k) * Hg(k, j);
          T(i, j) = acc;
        }
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
          Real acc = 0;
          for (int k = 0; k < n; k++)
            acc += T(i, k) * J(j, k);
          R(i, j) = acc;
        }
      for (int m = 1; m < n; m++) // second derivative of gamma wrt alpha (:419-431)
        for (int j = 1; j <= m; j++)
          R(j, j) += dg[m] * ea[j];
      for (int m = 0; m < n; m++)
        R(m, m) -= reg;
      for (auto &v : R.a) {
        v = -v;
        if (std::isnan(v))
          throw NanError("H has NaN");
      }
      *Hout = R;
    }
    for (int i = 0; i < n; i++) {
      Real acc = 0;
      for (int k = 0; k < n; k++)
        acc += (-J(i, k)) * dg[k];
      if (std::isnan(acc))
        throw NanError("dalpha has NaN");
      grad[i] = acc;
    }
    for (int m = 0; m < n; m++) {
      grad[m] += reg * alpha[m];
      ll = static_cast<Real>(ll - 0.5 * reg * alpha[m] * alpha[m]); // 0.5 is a double in the reference
    }
    return -ll;
  }

  static Real norm2(const Vec &v) {
    Real s = 0;
    for (Real x : v)
      s += x * x;
    return std::sqrt(s);
  }
  static Mat cholesky(const Mat &A) { // lower factor
    const int n = A.n;
    Mat L(n);
    for (int k = 0; k < n; k++) {
      Real d = A(k, k);
      for (int p = 0; p < k; p++)
        d -= L(k, p) * L(k, p);
      d = std::sqrt(d);
      L(k, k) = d;
      for (int i = k + 1; i < n; i++) {
        Real s = A(i, k);
        for (int p = 0; p < k; p++)
          s -= L(i, p) * L(k, p);
        L(i, k) = s / d;
      }
    }
    return L;
  }
  static void forward(const Mat &L, Vec &b) {
    for (int i = 0; i < L.n; i++) {
      Real s = b[i];
      for (int p = 0; p < i; p++)
        s -= L(i, p) * b[p];
      b[i] = s / L(i, i);
    }
  }
  static void backward(const Mat &L, Vec &b) { // L^T x = b
    for (int i = L.n - 1; i >= 0; i--) {
      Real s = b[i];
      for (int p = i + 1; p < L.n; p++)
        s -= L(p, i) * b[p];
      b[i] = s / L(i, i);
    }
  }

  // Damped Newton with step halving and the reference's three stopping rules (:289-357).
  // Leaves the Hessian at the mode in H.
  void find_mode(Vec &alpha) {
    const int max_iter = 10000, past = 3;
    const Real eps = static_cast<Real>(1e-5);
    Vec trial(alpha), grad(alpha), dir(alpha), history(past);
    Real f = 0;
    bool fresh = true;
    int it = 0;
    for (;;) {
      if (fresh)
        f = objective(alpha, grad, &H);
      const Real gn = norm2(grad);
      if (gn < eps || gn < eps * norm2(alpha))
        break;
      Mat L = cholesky(H);
      dir = grad;
      forward(L, dir);
      backward(L, dir);
      Real step = 1;
      for (int halvings = 0;;) {
        for (size_t k = 0; k < alpha.size(); k++)
          trial[k] = alpha[k] + step * (-dir[k]);
        Real f_new;
        try {
          f_new = objective(trial, grad, &H);
        } catch (NanError &) {
          step /= 2;
          continue;
        }
        if (f_new >= f * (1 + eps)) {
          step /= 2;
        } else {
          alpha = trial, f = f_new;
          break;
        }
        if (++halvings > 1000)
          break;
      }
      fresh = false;
      if (it >= past) {
        const Real old = history[it % past];
        if (std::abs(old - f) <= eps * std::max(std::max(std::abs(f), std::abs(old)), Real(1)))
          break;
      }
      history[it % past] = f;
      if (++it >= max_iter)
        throw std::runtime_error("Failed to converge. See fail-log.txt");
    }
  }

  void start() { // :274-279
    Vec a(K - 1, Real(0));
    find_mode(a);
    alpha_now = a;
    to_gamma(alpha_now, gamma_now);
  }

  Real log_t_density(const Vec &mode, const Vec &x) const { // :51-55, unnormalised
    const int n = H.n;
    Vec d(n), t(n);
    for (int i = 0; i < n; i++)
      d[i] = x[i] - mode[i];
    for (int j = 0; j < n; j++) {
      Real acc = 0;
      for (int i = 0; i < n; i++)
        acc += d[i] * H(i, j);
      t[j] = acc;
    }
    Real quad = 0;
    for (int j = 0; j < n; j++)
      quad += t[j] * d[j];
    return std::log(1 + quad / nu) * (-nu - n) / 2;
  }

  // One Metropolis-Hastings move with a t_nu(mode, H^-1) proposal (:57-72, :359-387).
  bool step(std::mt19937 &gen) {
    const int n = K - 1;
    Vec mode = alpha_now;
    find_mode(mode);
    Vec cand(n);
    {
      std::normal_distribution<Real> base(0, 1);
      std::gamma_distribution<Real> chi(nu / 2);
      for (int i = 0; i < n; i++)
        cand[i] = base(gen);
      Mat L = cholesky(H);
      backward(L, cand);
      const Real denom = std::sqrt(chi(gen) * 2 / nu);
      for (int i = 0; i < n; i++)
        cand[i] = cand[i] / denom + mode[i];
    }
    Vec scratch(n);
    Real ll_cand, ll_old;
    try {
      ll_cand = -objective(cand, scratch, nullptr);
      ll_old = -objective(alpha_now, scratch, nullptr);
    } catch (NanError &) {
      return false;
    }
    const Real lp_cand = log_t_density(mode, cand), lp_old = log_t_density(mode, alpha_now);
    const Real ratio = std::exp(ll_cand - lp_cand - ll_old + lp_old);
    const Real u = std::uniform_real_distribution<Real>{0, 1}(gen);
    if (u < ratio) {
      alpha_now = cand;
      to_gamma(alpha_now, gamma_now);
      accept_count++;
      return true;
    }
    return false;
  }
};

} // namespace myfm
