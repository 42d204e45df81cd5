// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (C++17, no Eigen) of the Gibbs sampler of tohtsky/myFM, used as the parity
// checker for the CUDA engine and as the single-thread CPU baseline.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it; the
// product path (myfm_b200/) never does.
//
// Pinning status: PINNED against every fixture and known-answer test the reference's own
// test-suite holds for this path — tests/test_oracle.py re-runs them on the oracle:
// block == flat at rtol 1e-7 incl. the n_workers / pickle round trip
// (reference tests/regression/test_block.py:80-149), predictor == running mean of per-iteration
// predict_score and planted-parameter recovery (tests/regression/test_fit.py:20-72,
// tests/classification/test_classification.py:14-70), ordered-probit cut-points and
// predict_proba == manual Phi differencing (tests/oprobit/test_oprobit_1dim.py:9-61), the fixtures
// of tests/conftest.py:15-45 — plus libstdc++ <random> known-answer values for the RNG contract.
// The reference ships no golden vectors and cannot itself be run here (it needs Eigen 3.4.0,
// fetched from the network by setup.py:21-50; Eigen is not installed, there is no network), so
// nothing tighter exists to pin against: what those tests do not constrain — the order of Eigen's
// vectorised dense reductions (restated from Eigen 3.4.0's Redux.h for an SSE2 build, see
// eigen_dense_sum) and the column-major fill order of FM.hpp:34-45 — is an assumption of the
// restatement, stated in every parity report.
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).
// RNG: std::mt19937 + libstdc++ distributions constructed exactly where the reference constructs
// them (fresh vs. persistent objects), since that is what "same seed" means.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

// Provided by the reference's vendored cpp_source/Faddeeva.cc, compiled where it lies into
// oracle/_ref/ (see oracle/Makefile).  Only the two real-argument entry points are used
// (include/myfm/OProbitSampler.hpp:122-225, include/myfm/util.hpp:97,101).
namespace Faddeeva {
extern double erfcx(double x);
extern double erf(double x);
} // namespace Faddeeva

namespace oracle {

enum class Task : int { REGRESSION = 0, CLASSIFICATION = 1, ORDERED = 2 };

// Sum of f(0) .. f(n-1) in the order of Eigen's dense vectorised reduction, which is what the
// reference's two dense `.sum()` calls run (FMTrainer.hpp:138, :223).  Eigen 3.4.0 (the version
// setup.py:21-23 pins; not available here) implements it in Eigen/src/Core/Redux.h as
// redux_impl<Func, Evaluator, LinearVectorizedTraversal, NoUnrolling>, restated from the published
// source: two packet accumulators walk the vector in strides of two packets, are added, a last
// odd packet is added, the packet is reduced horizontally, and the scalar tail follows.  A default
// `pip install` of the reference compiles without -march, i.e. SSE2: 16-byte packets (4 floats,
// 2 doubles), predux = (a0 + a2) + (a1 + a3) for floats and a0 + a1 for doubles
// (Eigen/src/Core/arch/SSE/PacketMath.h).  Both expressions are coefficient-wise ops without
// direct access, so the aligned start is 0.  For double this differs from a serial loop at the
// 1e-16 level; for float over 10^7 rows a serial accumulator is off by 1e-3, so the order matters.
template <typename Real, typename F> Real eigen_dense_sum(size_t n, F f) {
  constexpr size_t P = 16 / sizeof(Real);
  if (n == 0)
    return Real(0); // Eigen returns Scalar(0) for an empty sum (DenseBase::sum)
  const size_t aligned2 = (n / (2 * P)) * (2 * P), aligned = (n / P) * P;
  Real res;
  if (aligned) {
    Real p0[P], p1[P];
    for (size_t k = 0; k < P; k++)
      p0[k] = f(k);
    if (aligned > P) {
      for (size_t k = 0; k < P; k++)
        p1[k] = f(P + k);
      for (size_t i = 2 * P; i < aligned2; i += 2 * P)
        for (size_t k = 0; k < P; k++) {
          p0[k] += f(i + k);
          p1[k] += f(i + P + k);
        }
      for (size_t k = 0; k < P; k++)
        p0[k] += p1[k];
      if (aligned > aligned2)
        for (size_t k = 0; k < P; k++)
          p0[k] += f(aligned2 + k);
    }
    res = P == 4 ? (p0[0] + p0[2 % P]) + (p0[1] + p0[3 % P]) : p0[0] + p0[1];
    for (size_t i = aligned; i < n; i++)
      res += f(i);
  } else {
    res = f(0);
    for (size_t i = 1; i < n; i++)
      res += f(i);
  }
  return res;
}

// Row-compressed sparse matrix; stands in for Eigen::SparseMatrix<Real, RowMajor>
// (include/myfm/definitions.hpp:22-23).  Indices are taken as given: not sorted, duplicates kept.
template <typename Real> struct Csr {
  int64_t rows = 0, cols = 0;
  std::vector<int64_t> ptr;
  std::vector<int32_t> idx;
  std::vector<Real> val;

  // X.transpose() as a row-major matrix (BaseFMTrainer.hpp:61): every output row (= feature)
  // lists its entries by ascending source row, which fixes the summation order of all
  // per-column reductions.
  Csr transposed() const {
    Csr t;
    t.rows = cols;
    t.cols = rows;
    t.ptr.assign(cols + 1, 0);
    for (int32_t c : idx)
      t.ptr[c + 1]++;
    for (int64_t c = 0; c < cols; c++)
      t.ptr[c + 1] += t.ptr[c];
    t.idx.resize(idx.size());
    t.val.resize(val.size());
    std::vector<int64_t> cur(t.ptr.begin(), t.ptr.end() - 1);
    for (int64_t r = 0; r < rows; r++)
      for (int64_t p = ptr[r]; p < ptr[r + 1]; p++) {
        int64_t dst = cur[idx[p]]++;
        t.idx[dst] = static_cast<int32_t>(r);
        t.val[dst] = val[p];
      }
    return t;
  }

  // out = this * x[0:cols]; one sequential accumulator per row, as Eigen's row-major
  // sparse * dense kernel does.
  void spmv(const Real *x, Real *out) const {
    for (int64_t r = 0; r < rows; r++) {
      Real acc = 0;
      for (int64_t p = ptr[r]; p < ptr[r + 1]; p++)
        acc += val[p] * x[idx[p]];
      out[r] = acc;
    }
  }
  // out = cwiseAbs2(this) * (x .^ 2)
  void spmv_sq(const Real *x, Real *out) const {
    for (int64_t r = 0; r < rows; r++) {
      Real acc = 0;
      for (int64_t p = ptr[r]; p < ptr[r + 1]; p++)
        acc += (val[p] * val[p]) * (x[idx[p]] * x[idx[p]]);
      out[r] = acc;
    }
  }
};

// include/myfm/definitions.hpp:30-52
template <typename Real> struct RelationBlock {
  std::vector<size_t> original_to_block;
  Csr<Real> X;
  size_t block_size = 0, feature_size = 0;

  RelationBlock(std::vector<size_t> map, Csr<Real> data)
      : original_to_block(std::move(map)), X(std::move(data)), block_size(X.rows),
        feature_size(X.cols) {
    for (size_t c : original_to_block)
      if (c >= block_size)
        throw std::runtime_error("index mapping points to non-existing row.");
  }
};

// include/myfm/definitions.hpp:54-84
template <typename Real> struct RelationCache {
  Csr<Real> X_t;
  std::vector<Real> cardinality, q, q_S, c, c_S, e, e_q;
  explicit RelationCache(const RelationBlock<Real> &src)
      : X_t(src.X.transposed()), cardinality(src.block_size, Real(0)), q(src.block_size),
        q_S(src.block_size), c(src.block_size), c_S(src.block_size), e(src.block_size),
        e_q(src.block_size) {
    for (size_t v : src.original_to_block)
      cardinality[v]++;
  }
};

// include/myfm/FMLearningConfig.hpp:12-89 (the validated, immutable config)
struct Config {
  double alpha_0 = 1, beta_0 = 1, gamma_0 = 1, mu_0 = 1, reg_0 = 1;
  Task task = Task::REGRESSION;
  double nu_oprobit = 5;
  bool fit_w0 = true, fit_linear = true;
  int n_iter = 100, n_kept_samples = 10;
  double cutpoint_scale = 10;
  std::vector<size_t> group_index;
  std::vector<std::pair<size_t, std::vector<size_t>>> cutpoint_groups;

  size_t n_groups = 0;
  std::vector<std::vector<size_t>> group_vs_feature_index;

  // FMLearningConfig.hpp:29-56
  void finalize() {
    std::vector<size_t> sorted(group_index);
    std::sort(sorted.begin(), sorted.end());
    sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
    n_groups = sorted.size();
    for (size_t i = 0; i < n_groups; i++)
      if (sorted[i] != i) {
        std::ostringstream ss;
        ss << "No matching index for group index " << i << " found.";
        throw std::invalid_argument(ss.str());
      }
    group_vs_feature_index.assign(n_groups, {});
    for (size_t f = 0; f < group_index.size(); f++)
      group_vs_feature_index[group_index[f]].push_back(f);
    if (n_kept_samples < 0)
      throw std::invalid_argument("n_kept_samples must be non-negative,");
    if (n_iter <= 0)
      throw std::invalid_argument("n_iter must be positive.");
    if (n_iter < n_kept_samples)
      throw std::invalid_argument("n_kept_samples must not exceed n_iter.");
  }
};

// include/myfm/HyperParams.hpp:8-37; mu_V / lambda_V are (n_groups x n_factors) column-major.
template <typename Real> struct Hyper {
  size_t n_groups = 0, n_factors = 0;
  Real alpha = 0;
  std::vector<Real> mu_w, lambda_w, mu_V, lambda_V;
  Hyper(size_t rank, size_t groups)
      : n_groups(groups), n_factors(rank), mu_w(groups), lambda_w(groups), mu_V(groups * rank),
        lambda_V(groups * rank) {}
  Real &muV(size_t g, size_t r) { return mu_V[g + n_groups * r]; }
  Real &lamV(size_t g, size_t r) { return lambda_V[g + n_groups * r]; }
};

// ---------------------------------------------------------------------------------------------
// Truncated normal samplers, include/myfm/util.hpp:15-78 (Robert 1995, Prop. 2.3).
// ---------------------------------------------------------------------------------------------
template <typename Real> inline Real tn_left(std::mt19937 &gen, Real mu_minus) {
  if (mu_minus < 0) {
    std::normal_distribution<Real> dist(0, 1); // one object for the whole rejection loop
    while (true) {
      Real z = dist(gen);
      if (z > mu_minus)
        return z;
    }
  } else {
    Real alpha_star = (mu_minus + std::sqrt(mu_minus * mu_minus + 4)) / 2;
    std::uniform_real_distribution<Real> dist(0, 1);
    while (true) {
      Real z = -std::log(dist(gen)) / alpha_star + mu_minus;
      Real rho = std::exp(-(z - alpha_star) * (z - alpha_star) / 2);
      Real u = dist(gen);
      if (u < rho)
        return z;
    }
  }
}

template <typename Real> inline Real tn_twoside(std::mt19937 &gen, Real mu_minus, Real mu_plus) {
  std::uniform_real_distribution<Real> proposal(mu_minus, mu_plus);
  std::uniform_real_distribution<Real> acceptance(0, 1);
  Real rho;
  while (true) {
    Real z = proposal(gen);
    if ((mu_minus <= static_cast<Real>(0)) && (mu_plus >= static_cast<Real>(0)))
      rho = std::exp(-z * z / 2);
    else if (mu_plus < static_cast<Real>(0))
      rho = std::exp((mu_plus * mu_plus - z * z) / 2);
    else
      rho = std::exp((mu_minus * mu_minus - z * z) / 2);
    Real u = acceptance(gen);
    if (u < rho)
      return z;
  }
}

template <typename Real> inline Real tn_right(std::mt19937 &gen, Real mu_plus) {
  return -tn_left<Real>(gen, -mu_plus);
}
template <typename Real> inline Real tn_left(std::mt19937 &gen, Real mean, Real sd, Real mu_minus) {
  return mean + sd * tn_left<Real>(gen, (mu_minus - mean) / sd);
}
template <typename Real> inline Real tn_right(std::mt19937 &gen, Real mean, Real sd, Real mu_plus) {
  return mean + sd * tn_right<Real>(gen, (mu_plus - mean) / sd);
}

// ---------------------------------------------------------------------------------------------
// FM state and forward pass, include/myfm/FM.hpp.
// ---------------------------------------------------------------------------------------------
template <typename Real> struct FM {
  int n_factors = 0;
  size_t n_features = 0;
  Real w0 = 0;
  std::vector<Real> w;                      // [n_features]
  std::vector<Real> V;                      // [n_features x n_factors], column-major
  std::vector<std::vector<Real>> cutpoints; // ordered probit

  explicit FM(int rank) : n_factors(rank) {}
  Real &v(size_t j, size_t r) { return V[j + n_features * r]; }
  const Real *vcol(size_t r) const { return V.data() + n_features * r; }

  // FM.hpp:34-45: V (column by column), then w, then w0, all from ONE normal_distribution so the
  // cached second polar variate is consumed.
  void initialize_weight(size_t dim, Real init_std, std::mt19937 &gen) {
    n_features = dim;
    std::normal_distribution<Real> nd;
    V.resize(dim * static_cast<size_t>(n_factors));
    for (auto &x : V)
      x = nd(gen) * init_std;
    w.resize(dim);
    for (auto &x : w)
      x = nd(gen) * init_std;
    w0 = nd(gen) * init_std;
  }

  // FM.hpp:54-136
  void predict_score(Real *target, const Csr<Real> &X,
                     const std::vector<RelationBlock<Real>> &rels) const {
    size_t case_size = X.rows, feature_size_all = X.cols;
    for (auto const &rel : rels) {
      if (case_size != rel.original_to_block.size())
        throw std::invalid_argument("Relation blocks have inconsistent mapper size with case_size");
      feature_size_all += rel.feature_size;
    }
    if (feature_size_all != w.size()) {
      std::ostringstream ss;
      ss << "Total feature size mismatch. Should be " << w.size() << ", but got "
         << feature_size_all << ".";
      throw std::invalid_argument(ss.str());
    }
    const size_t n = case_size;
    X.spmv(w.data(), target);
    for (size_t i = 0; i < n; i++)
      target[i] = w0 + target[i];
    std::vector<Real> block;
    size_t offset = X.cols;
    for (auto const &rel : rels) {
      block.resize(rel.block_size);
      rel.X.spmv(w.data() + offset, block.data());
      for (size_t i = 0; i < n; i++)
        target[i] += block[rel.original_to_block[i]];
      offset += rel.feature_size;
    }
    std::vector<Real> q(n), v2;
    for (int r = 0; r < n_factors; r++) {
      const Real *vr = vcol(r);
      X.spmv(vr, q.data());
      offset = X.cols;
      for (auto const &rel : rels) {
        block.resize(rel.block_size);
        rel.X.spmv(vr + offset, block.data());
        offset += rel.feature_size;
        for (size_t i = 0; i < n; i++)
          q[i] += block[rel.original_to_block[i]];
      }
      for (size_t i = 0; i < n; i++)
        target[i] += (q[i] * q[i]) * static_cast<Real>(0.5);
      X.spmv_sq(vr, q.data());
      offset = X.cols;
      for (auto const &rel : rels) {
        block.resize(rel.block_size);
        rel.X.spmv_sq(vr + offset, block.data());
        offset += rel.feature_size;
        for (size_t i = 0; i < n; i++)
          q[i] += block[rel.original_to_block[i]];
      }
      for (size_t i = 0; i < n; i++)
        target[i] -= q[i] * static_cast<Real>(0.5);
    }
  }

  // FM.hpp:137-162; out is (rows x (n_cpt+1)) column-major.
  void oprobit_predict_proba(Real *out, const Csr<Real> &X,
                             const std::vector<RelationBlock<Real>> &rels,
                             size_t cutpoint_index) const {
    if (cutpoints.empty())
      throw std::runtime_error("No cutpoint available for this FM.");
    const std::vector<Real> &cp = cutpoints.at(cutpoint_index);
    const int n_cpt = static_cast<int>(cp.size());
    const size_t n = X.rows;
    std::vector<Real> score(n);
    predict_score(score.data(), X, rels);
    for (int c = 0; c < n_cpt; c++)
      for (size_t i = 0; i < n; i++)
        out[i + n * c] =
            (1 + std::erf((cp[c] - score[i]) * static_cast<Real>(std::sqrt(0.5)))) / 2;
    for (size_t i = 0; i < n; i++)
      out[i + n * n_cpt] = 1 - out[i + n * (n_cpt - 1)];
    for (int c = n_cpt - 1; c >= 1; c--)
      for (size_t i = 0; i < n; i++)
        out[i + n * c] -= out[i + n * (c - 1)];
  }
};

// ---------------------------------------------------------------------------------------------
// Ordered-probit cut-point sampler, include/myfm/OProbitSampler.hpp.
// Small dense algebra ((K-1) x (K-1), column-major) is written out by hand.
// ---------------------------------------------------------------------------------------------
template <typename Real> struct OprobitSampler {
  static constexpr Real SQRT2 = 1.4142135623730951;
  static constexpr Real SQRTPI = 1.7724538509055159;
  static constexpr Real SQRT2PI = SQRT2 * SQRTPI;
  static constexpr Real PI = 3.141592653589793;
  using Vec = std::vector<Real>;

  struct Mat { // square, column-major
    int n = 0;
    std::vector<Real> a;
    explicit Mat(int n_ = 0) : n(n_), a(static_cast<size_t>(n_) * n_, Real(0)) {}
    Real &operator()(int i, int j) { return a[i + static_cast<size_t>(n) * j]; }
    Real operator()(int i, int j) const { return a[i + static_cast<size_t>(n) * j]; }
    void zero() { std::fill(a.begin(), a.end(), Real(0)); }
    bool has_nan() const {
      for (Real x : a)
        if (std::isnan(x))
          return true;
      return false;
    }
  };

  Vec &x_;
  const Vec &y_;
  int K;
  std::vector<size_t> indices_;
  Real reg, nu;
  std::mt19937 &rng;
  Vec alpha_now, gamma_now;
  Mat H;
  Vec zmins, zmaxs;
  std::vector<size_t> histogram;
  size_t accept_count = 0;

  // OProbitSampler.hpp:25-49
  OprobitSampler(Vec &x, const Vec &y, int K_, const std::vector<size_t> &indices,
                 std::mt19937 &rng_, Real reg_, Real nu_)
      : x_(x), y_(y), K(K_), indices_(indices), reg(reg_), nu(nu_), rng(rng_),
        alpha_now(K_ - 1, Real(0)), gamma_now(K_ - 1, Real(0)), H(K_ - 1), zmins(K_), zmaxs(K_),
        histogram(K_, 0) {
    alpha_to_gamma(gamma_now, alpha_now);
    for (size_t i : indices_) {
      int y_label = static_cast<int>(y_[i]);
      if (std::abs(y_label - y_[i]) > 1e-3)
        throw std::invalid_argument("y has a floating-point element.");
      if (y_label < 0)
        throw std::invalid_argument("y has a negative element.");
      if (y_label >= K) {
        std::ostringstream ss;
        ss << "y[ " << i << "] is greater than " << (K - 1) << ".";
        throw std::invalid_argument(ss.str());
      }
      histogram[y_label]++;
    }
  }

  // :95-101
  static void alpha_to_gamma(Vec &target, const Vec &alpha) {
    target[0] = alpha[0];
    for (size_t i = 1; i < alpha.size(); i++)
      target[i] = target[i - 1] + std::exp(alpha[i]);
  }

  // :74-93 (fix_gamma0 == false); J(i,j) = d gamma_j / d alpha_i
  static void jacobian(Mat &J, const Vec &alpha) {
    const int n = static_cast<int>(alpha.size());
    J.zero();
    J(0, 0) = 1;
    for (int j = 1; j < n; j++)
      J(0, j) = 1;
    for (int i = 1; i < n; i++) {
      Real ed = std::exp(alpha[i]);
      for (int j = i; j < n; j++)
        J(i, j) = ed;
    }
  }

  // lower Cholesky factor of a symmetric positive-definite matrix (stands in for Eigen::LLT)
  static Mat cholesky_lower(const Mat &A) {
    const int n = A.n;
    Mat L(n);
    for (int k = 0; k < n; k++) {
      Real x = A(k, k);
      for (int p = 0; p < k; p++)
        x -= L(k, p) * L(k, p);
      x = std::sqrt(x);
      L(k, k) = x;
      for (int i = k + 1; i < n; i++) {
        Real s = A(i, k);
        for (int p = 0; p < k; p++)
          s -= L(i, p) * L(k, p);
        L(i, k) = s / x;
      }
    }
    return L;
  }
  static void solve_lower(const Mat &L, Vec &b) { // L y = b
    for (int i = 0; i < L.n; i++) {
      Real s = b[i];
      for (int p = 0; p < i; p++)
        s -= L(i, p) * b[p];
      b[i] = s / L(i, i);
    }
  }
  static void solve_lower_transposed(const Mat &L, Vec &b) { // L^T x = b
    for (int i = L.n - 1; i >= 0; i--) {
      Real s = b[i];
      for (int p = i + 1; p < L.n; p++)
        s -= L(p, i) * b[p];
      b[i] = s / L(i, i);
    }
  }

  // :51-55
  Real log_p_mvt(const Mat &SigmaInverse, const Vec &mu, Real nu_, const Vec &x) const {
    const int n = SigmaInverse.n;
    Vec d(n), t(n, Real(0));
    for (int i = 0; i < n; i++)
      d[i] = x[i] - mu[i];
    // (d^T * S) then * d, as the left-to-right Eigen product
    for (int j = 0; j < n; j++) {
      Real acc = 0;
      for (int i = 0; i < n; i++)
        acc += d[i] * SigmaInverse(i, j);
      t[j] = acc;
    }
    Real log_p = 0;
    for (int j = 0; j < n; j++)
      log_p += t[j] * d[j];
    return std::log(1 + log_p / nu_) * (-nu_ - n) / 2;
  }

  // :57-72
  Vec sample_mvt(const Mat &SigmaInverse, Real nu_) {
    const int n = SigmaInverse.n;
    Vec result(n);
    std::normal_distribution<Real> base_dist(0, 1);
    std::gamma_distribution<Real> chi_gen(nu_ / 2);
    for (int i = 0; i < n; i++)
      result[i] = base_dist(rng);
    Mat L = cholesky_lower(SigmaInverse); // U = L^T; solve U x = z
    solve_lower_transposed(L, result);
    Real denom = std::sqrt(chi_gen(rng) * 2 / nu_);
    for (auto &v : result)
      v /= denom;
    return result;
  }

  // :111-181
  static void safe_ldiff(Real x, Real y, Real &loss, Real &dx, Real &dy, Mat *Ht, int label) {
    Real denominator, exp_factor;
    if (y > 0) {
      exp_factor = std::exp((y * y - x * x) / 2);
      denominator = Faddeeva::erfcx(y / SQRT2) - exp_factor * Faddeeva::erfcx(x / SQRT2);
      loss -= y * y / 2;
      loss += std::log(denominator / 2);
      dx += (2 / SQRT2PI) * exp_factor / denominator;
      dy -= (2 / SQRT2PI) / denominator;
      if (Ht != nullptr) {
        (*Ht)(label, label) += -(SQRT2PI * x * denominator * std::exp((y * y - x * x) / 2) +
                                 2 * std::exp(y * y - x * x)) /
                               denominator / denominator / PI;
        (*Ht)(label - 1, label - 1) +=
            (SQRT2PI * y * denominator - 2) / denominator / denominator / PI;
        Real off_diag = 2 * std::exp((y * y - x * x) / 2) / PI / denominator / denominator;
        (*Ht)(label, label - 1) += off_diag;
        (*Ht)(label - 1, label) += off_diag;
      }
    } else if (x < 0) {
      loss -= x * x / 2;
      exp_factor = std::exp((x * x - y * y) / 2);
      denominator = Faddeeva::erfcx(-x / SQRT2) - exp_factor * Faddeeva::erfcx(-y / SQRT2);
      loss += std::log(denominator / 2);
      dx += (2 / SQRT2PI) / denominator;
      dy -= (2 / SQRT2PI) * exp_factor / denominator;
      if (Ht != nullptr) {
        (*Ht)(label, label) += -(SQRT2PI * x * denominator + 2) / PI / denominator / denominator;
        (*Ht)(label - 1, label - 1) +=
            (SQRT2PI * y * exp_factor * denominator - 2 * (exp_factor * exp_factor)) / PI /
            denominator / denominator;
        Real off_diag = 2 * exp_factor / PI / denominator / denominator;
        (*Ht)(label, label - 1) += off_diag;
        (*Ht)(label - 1, label) += off_diag;
      }
    } else {
      denominator = Faddeeva::erf(x / SQRT2) - Faddeeva::erf(y / SQRT2);
      Real expxx = std::exp(-x * x / 2);
      Real expyy = std::exp(-y * y / 2);
      dx += 2 * expxx / denominator / SQRT2PI;
      dy -= 2 * expyy / denominator / SQRT2PI;
      loss += std::log(denominator / 2);
      if (Ht != nullptr) {
        (*Ht)(label, label) += -(SQRT2PI * x * denominator * expxx + 2 * expxx * expxx) / PI /
                               denominator / denominator;
        (*Ht)(label - 1, label - 1) +=
            -(-SQRT2PI * y * denominator * expyy + 2 * expyy * expyy) / PI / denominator /
            denominator;
        Real off_diag = 2 * expxx * expyy / PI / denominator / denominator;
        (*Ht)(label, label - 1) += off_diag;
        (*Ht)(label - 1, label) += off_diag;
      }
    }
  }

  // :183-209
  static void safe_lcdf(Real x, Real &loss, Real &dx, Mat *Ht, int label) {
    Real denominator, exp_factor;
    if (x > 1) {
      exp_factor = std::exp(-x * x / 2);
      denominator = 1 + Faddeeva::erf(x / SQRT2);
      dx += (2 / SQRT2PI) * exp_factor / denominator;
      loss += std::log(denominator / 2);
      if (Ht != nullptr)
        (*Ht)(label, label) +=
            -(SQRT2PI * x * denominator * exp_factor + 2 * exp_factor * exp_factor) / PI /
            denominator / denominator;
    } else {
      denominator = Faddeeva::erfcx(-x / SQRT2);
      dx += (2 / SQRT2PI) / denominator;
      loss -= x * x / 2;
      loss += std::log(denominator / 2);
      if (Ht != nullptr)
        (*Ht)(label, label) += -(SQRT2PI * x * denominator + 2) / PI / denominator / denominator;
    }
  }

  // :211-236
  static void safe_lccdf(Real x, Real &loss, Real &dx, Mat *Ht, int label) {
    Real denominator;
    if (x > -1) {
      denominator = Faddeeva::erfcx(x / SQRT2);
      dx -= (2 / SQRT2PI) / denominator;
      loss += std::log(denominator / 2);
      loss -= x * x / 2;
      if (Ht != nullptr)
        (*Ht)(label - 1, label - 1) +=
            (SQRT2PI * x * denominator - 2) / denominator / denominator / PI;
    } else {
      denominator = 1 - Faddeeva::erf(x / SQRT2);
      dx -= (2 / SQRT2PI) * std::exp(-x * x / 2) / denominator;
      loss += std::log(denominator / 2);
      if (Ht != nullptr) {
        Real exp_factor = std::exp(-(x * x) / 2);
        (*Ht)(label - 1, label - 1) +=
            -(-SQRT2PI * x * denominator * exp_factor + 2 * exp_factor * exp_factor) / PI /
            denominator / denominator;
      }
    }
  }

  // :389-463; returns the NEGATIVE log posterior, writes its gradient to dalpha and (optionally)
  // its Hessian to *Ht.
  Real objective(const Vec &alpha, Vec &dalpha, Mat *Ht = nullptr) {
    const int n = static_cast<int>(alpha.size());
    Vec gamma(n, Real(0));
    std::fill(dalpha.begin(), dalpha.end(), Real(0));
    alpha_to_gamma(gamma, alpha);
    Mat J(n);
    jacobian(J, alpha);
    Real ll = 0;
    if (Ht != nullptr)
      Ht->zero();
    for (size_t i : indices_) {
      int label = y_[i];
      if (label == 0)
        safe_lcdf(gamma[0] - x_[i], ll, dalpha[0], Ht, label);
      else if (label == (K - 1))
        safe_lccdf(gamma[K - 2] - x_[i], ll, dalpha[K - 2], Ht, label);
      else
        safe_ldiff(gamma[label] - x_[i], gamma[label - 1] - x_[i], ll, dalpha[label],
                   dalpha[label - 1], Ht, label);
    }
    if (Ht != nullptr) {
      Mat &Hm = *Ht;
      Vec expAlpha(n);
      for (int i = 0; i < n; i++)
        expAlpha[i] = std::exp(alpha[i]);
      Mat T(n), R(n); // T = J * H ; R = T * J^T
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
          Real acc = 0;
          for (int k = 0; k < n; k++)
            acc += J(i, k) * Hm(k, j);
          T(i, j) = acc;
        }
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
          Real acc = 0;
          for (int k = 0; k < n; k++)
            acc += T(i, k) * J(j, k);
          R(i, j) = acc;
        }
      Hm = R;
      for (int m = 1; m < (K - 1); m++)
        for (int j = 1; j <= m; j++)
          Hm(j, j) += dalpha[m] * expAlpha[j];
      Hm(0, 0) -= reg;
      for (int m = 1; m < (K - 1); m++)
        Hm(m, m) -= reg;
      for (auto &v : Hm.a)
        v *= -1;
      if (Hm.has_nan())
        throw std::runtime_error("H has NaN");
    }
    { // dalpha = -J * dalpha (product evaluated into a temporary first)
      Vec tmp(n);
      for (int i = 0; i < n; i++) {
        Real acc = 0;
        for (int k = 0; k < n; k++)
          acc += (-J(i, k)) * dalpha[k];
        tmp[i] = acc;
      }
      dalpha = tmp;
    }
    for (Real v : dalpha)
      if (std::isnan(v))
        throw std::runtime_error("dalpha has NaN");
    dalpha[0] += reg * alpha[0];
    ll -= 0.5 * reg * alpha[0] * alpha[0];
    for (int m = 1; m < (K - 1); m++) {
      dalpha[m] += reg * alpha[m];
      ll -= 0.5 * reg * alpha[m] * alpha[m];
    }
    return -ll;
  }

  static Real norm2(const Vec &v) {
    Real s = 0;
    for (Real x : v)
      s += x * x;
    return std::sqrt(s);
  }

  // :289-357 (damped Newton with step halving; control flow kept as is)
  void find_minimum(Vec &alpha_hat) {
    const int max_iter = 10000;
    const Real epsilon = 1e-5, epsilon_rel = 1e-5, delta = 1e-5;
    const int past = 3;
    Vec history(past);
    Vec alpha_new(alpha_hat), dalpha(alpha_hat), direction(alpha_hat);
    Real ll_current = 0;
    bool first = true;
    int i = 0;
    while (true) {
      if (first)
        ll_current = objective(alpha_hat, dalpha, &H);
      {
        Real alpha2 = norm2(alpha_hat), dalpha2 = norm2(dalpha);
        if (dalpha2 < epsilon || dalpha2 < epsilon_rel * alpha2)
          break;
      }
      {
        Mat L = cholesky_lower(H);
        direction = dalpha;
        solve_lower(L, direction);
        solve_lower_transposed(L, direction);
        for (auto &v : direction)
          v = -v;
      }
      Real step_size = 1;
      int lsc = 0;
      while (true) {
        for (size_t k = 0; k < alpha_hat.size(); k++)
          alpha_new[k] = alpha_hat[k] + step_size * direction[k];
        Real ll_new;
        try {
          ll_new = objective(alpha_new, dalpha, &H);
        } catch (std::runtime_error &) {
          step_size /= 2;
          continue;
        }
        if (ll_new >= (ll_current * (1 + delta))) {
          step_size /= 2;
        } else {
          alpha_hat = alpha_new;
          ll_current = ll_new;
          break;
        }
        if (++lsc > 1000)
          break;
      }
      first = false;
      if (i >= past) {
        Real past_loss = history[i % past];
        if (std::abs(past_loss - ll_current) <=
            delta * std::max(std::max(std::abs(ll_current), std::abs(past_loss)), Real(1)))
          break;
      }
      history[i % past] = ll_current;
      i++;
      if (i >= max_iter)
        break;
    }
    if (i == max_iter)
      throw std::runtime_error("Failed to converge. See fail-log.txt");
  }

  // :274-279
  void start_sample() {
    Vec alpha_hat(K - 1, Real(0));
    find_minimum(alpha_hat);
    alpha_now = alpha_hat;
    alpha_to_gamma(gamma_now, alpha_now);
  }

  // :359-387
  bool step() {
    Vec alpha_hat = alpha_now;
    Vec gamma(alpha_hat);
    find_minimum(alpha_hat);
    Vec alpha_candidate = sample_mvt(H, nu);
    for (size_t k = 0; k < alpha_candidate.size(); k++)
      alpha_candidate[k] = alpha_candidate[k] + alpha_hat[k];
    Real ll_candidate, ll_old;
    try {
      ll_candidate = -objective(alpha_candidate, gamma);
      ll_old = -objective(alpha_now, gamma);
    } catch (std::runtime_error &) {
      return false;
    }
    Real lp_candidate = log_p_mvt(H, alpha_hat, nu, alpha_candidate);
    Real lp_old = log_p_mvt(H, alpha_hat, nu, alpha_now);
    Real test_ratio = std::exp(ll_candidate - lp_candidate - ll_old + lp_old);
    Real u = std::uniform_real_distribution<Real>{0, 1}(rng);
    if (u < test_ratio) {
      alpha_now = alpha_candidate;
      alpha_to_gamma(gamma_now, alpha_now);
      accept_count++;
      return true;
    }
    return false;
  }

  // :238-272
  void sample_z_given_cutpoint() {
    std::fill(zmins.begin(), zmins.end(), std::numeric_limits<Real>::max());
    std::fill(zmaxs.begin(), zmaxs.end(), std::numeric_limits<Real>::lowest());
    Real deviation = 1;
    for (int train_data_index : indices_) {
      int class_index = static_cast<int>(y_[train_data_index]);
      Real pred_score = x_[train_data_index];
      Real z_new;
      if (class_index == 0) {
        z_new = deviation * tn_right<Real>(rng, (gamma_now[class_index] - pred_score) / deviation) +
                pred_score;
        zmaxs[0] = std::max(zmaxs[0], z_new);
      } else if (class_index == (K - 1)) {
        z_new =
            deviation * tn_left<Real>(rng, (gamma_now[K - 2] - pred_score) / deviation) + pred_score;
        zmins[K - 1] = std::min(zmins[K - 1], z_new);
      } else {
        z_new = deviation * tn_twoside<Real>(
                                rng, (gamma_now[class_index - 1] - pred_score) / deviation,
                                (gamma_now[class_index] - pred_score) / deviation) +
                pred_score;
        zmins[class_index] = std::min(zmins[class_index], z_new);
        zmaxs[class_index] = std::max(zmaxs[class_index], z_new);
      }
      x_[train_data_index] -= z_new;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Gibbs trainer: include/myfm/BaseFMTrainer.hpp + include/myfm/FMTrainer.hpp.
// ---------------------------------------------------------------------------------------------
template <typename Real> struct Trainer {
  Csr<Real> X;
  std::vector<RelationBlock<Real>> relations;
  Csr<Real> X_t;
  size_t dim_all = 0;
  std::vector<Real> y;
  int n_train = 0;
  std::vector<Real> e_train, q_train;
  std::vector<RelationCache<Real>> relation_caches;
  Config cfg;
  std::mt19937 gen_;
  std::vector<OprobitSampler<Real>> cutpoint_sampler;

  // BaseFMTrainer.hpp:58-105
  Trainer(Csr<Real> X_, std::vector<RelationBlock<Real>> rels, std::vector<Real> y_, int seed,
          Config config)
      : X(std::move(X_)), relations(std::move(rels)), X_t(X.transposed()), y(std::move(y_)),
        n_train(static_cast<int>(X.rows)), e_train(X.rows), q_train(X.rows), cfg(std::move(config)),
        gen_(seed) {
    // util.hpp:147-165
    dim_all = X.cols;
    int i = 0;
    for (auto const &rel : relations) {
      if (static_cast<size_t>(X.rows) != rel.original_to_block.size()) {
        std::ostringstream ss;
        ss << "main table has size " << X.rows << " but the relation[" << i << "] has size "
           << rel.original_to_block.size();
        throw std::runtime_error(ss.str());
      }
      dim_all += rel.feature_size;
      i++;
    }
    for (auto const &rel : relations)
      relation_caches.emplace_back(rel);
    if (static_cast<size_t>(X.rows) != y.size()) {
      std::ostringstream ss;
      ss << "Shape mismatch: X has size " << X.rows << " and y has size " << y.size();
      throw std::runtime_error(ss.str());
    }
    if (cfg.task == Task::ORDERED) {
      const size_t rows = X.rows;
      std::vector<bool> existence(rows, false);
      for (auto &group : cfg.cutpoint_groups)
        for (size_t k : group.second) {
          if (k >= rows)
            throw std::invalid_argument("out of range for cutpoint group config.");
          if (existence[k]) {
            std::ostringstream ss;
            ss << "index " << k << " overlapping in cutpoint config.";
            throw std::invalid_argument(ss.str());
          }
          existence[k] = true;
        }
      for (size_t r = 0; r < rows; r++)
        if (!existence[r]) {
          std::ostringstream ss;
          ss << "cutpoint group not specified for " << r << ".";
          throw std::invalid_argument(ss.str());
        }
    }
  }

  // BaseFMTrainer.hpp:107-111
  FM<Real> create_FM(int rank, Real init_std) {
    FM<Real> fm(rank);
    fm.initialize_weight(dim_all, init_std, gen_);
    return fm;
  }
  Hyper<Real> create_Hyper(size_t rank) { return Hyper<Real>(rank, cfg.n_groups); }

  // FMTrainer.hpp:89-97
  void initialize_hyper(Hyper<Real> &hyper) {
    hyper.alpha = static_cast<Real>(1);
    std::fill(hyper.mu_w.begin(), hyper.mu_w.end(), static_cast<Real>(0));
    std::fill(hyper.lambda_w.begin(), hyper.lambda_w.end(), static_cast<Real>(1e-5));
    std::fill(hyper.mu_V.begin(), hyper.mu_V.end(), static_cast<Real>(0));
    std::fill(hyper.lambda_V.begin(), hyper.lambda_V.end(), static_cast<Real>(1e-5));
  }

  // FMTrainer.hpp:99-119
  void initialize_e(FM<Real> &fm) {
    fm.predict_score(e_train.data(), X, relations);
    if (cfg.task == Task::ORDERED) {
      cutpoint_sampler.reserve(cfg.cutpoint_groups.size()); // samplers hold references
      int i = 0;
      for (auto &group : cfg.cutpoint_groups) {
        fm.cutpoints.emplace_back(group.first - 1);
        cutpoint_sampler.emplace_back(e_train, y, static_cast<int>(group.first), group.second, gen_,
                                      static_cast<Real>(cfg.reg_0),
                                      static_cast<Real>(cfg.nu_oprobit));
        cutpoint_sampler[i].start_sample();
        OprobitSampler<Real>::alpha_to_gamma(fm.cutpoints[i], cutpoint_sampler[i].alpha_now);
        cutpoint_sampler[i].sample_z_given_cutpoint();
        i++;
      }
      return;
    }
    for (int r = 0; r < n_train; r++)
      e_train[r] -= y[r];
  }

  // FMTrainer.hpp:122-125 — a FRESH normal_distribution per draw (second polar variate dropped)
  Real sample_normal(const Real &quad, const Real &first) {
    return (first / quad) + std::normal_distribution<Real>(0, 1)(gen_) / std::sqrt(quad);
  }

  // FMTrainer.hpp:127-145
  void update_alpha(Hyper<Real> &hyper) {
    if (cfg.task == Task::CLASSIFICATION || cfg.task == Task::ORDERED) {
      hyper.alpha = static_cast<Real>(1);
      return;
    }
    Real e_all = eigen_dense_sum<Real>(e_train.size(), [&](size_t i) { return e_train[i] * e_train[i]; }); // :138
    Real exponent = (static_cast<Real>(cfg.alpha_0) + X.rows) / 2;
    Real variance = (static_cast<Real>(cfg.beta_0) + e_all) / 2;
    hyper.alpha = std::gamma_distribution<Real>(exponent, 1 / variance)(gen_);
  }

  // FMTrainer.hpp:150-169 ; weight/mu/lambda are one column (factor) or the w vector
  void update_lambda_generic(const Real *mu, Real *lambda, const Real *weight) {
    size_t g = 0;
    for (const auto &features : cfg.group_vs_feature_index) {
      Real mean = mu[g];
      Real alpha = static_cast<Real>(cfg.alpha_0) + features.size();
      Real beta = static_cast<Real>(cfg.beta_0);
      for (size_t f : features) {
        auto dev = weight[f] - mean;
        beta += dev * dev;
      }
      lambda[g] = std::gamma_distribution<Real>(alpha / 2, 2 / beta)(gen_);
      g++;
    }
  }

  // FMTrainer.hpp:174-192
  void update_mu_generic(Real *mu, const Real *lambda, const Real *weight) {
    size_t g = 0;
    for (const auto &features : cfg.group_vs_feature_index) {
      size_t n_in_group = features.size();
      Real square = lambda[g] * (static_cast<Real>(cfg.gamma_0) + n_in_group);
      Real linear = static_cast<Real>(cfg.gamma_0) * static_cast<Real>(cfg.mu_0);
      for (size_t f : features)
        linear += weight[f];
      linear *= lambda[g];
      mu[g] = sample_normal(square, linear);
      g++;
    }
  }

  // FMTrainer.hpp:218-229
  void update_w0(FM<Real> &fm, Hyper<Real> &hyper) {
    if (!cfg.fit_w0) {
      fm.w0 = 0; // NB: e_train keeps the stale contribution until update_e
      return;
    }
    Real s = eigen_dense_sum<Real>(e_train.size(), [&](size_t i) { return fm.w0 - e_train[i]; }); // :223
    Real lin = hyper.alpha * s;
    Real quad = hyper.alpha * n_train + static_cast<Real>(cfg.reg_0);
    Real w0_new = sample_normal(quad, lin);
    Real d = (w0_new - fm.w0);
    for (Real &v : e_train)
      v += d;
    fm.w0 = w0_new;
  }

  // FMTrainer.hpp:231-314
  void update_w(FM<Real> &fm, Hyper<Real> &hyper) {
    if (!cfg.fit_linear) {
      std::fill(fm.w.begin(), fm.w.end(), Real(0));
      return;
    }
    for (int64_t j = 0; j < X.cols; j++) { // :237-254
      size_t g = cfg.group_index[j];
      const Real w_old = fm.w[j];
      const int64_t b = X_t.ptr[j], en = X_t.ptr[j + 1];
      for (int64_t p = b; p < en; p++)
        e_train[X_t.idx[p]] -= X_t.val[p] * w_old;
      Real lambda = hyper.lambda_w[g], mu = hyper.mu_w[g];
      Real x2 = 0;
      for (int64_t p = b; p < en; p++)
        x2 += X_t.val[p] * X_t.val[p];
      Real square_term = lambda + hyper.alpha * x2;
      Real dot = 0; // ((-alpha) * x_j) . e
      for (int64_t p = b; p < en; p++)
        dot += ((-hyper.alpha) * X_t.val[p]) * e_train[X_t.idx[p]];
      Real linear_term = dot + lambda * mu;
      Real w_new = sample_normal(square_term, linear_term);
      for (int64_t p = b; p < en; p++)
        e_train[X_t.idx[p]] += X_t.val[p] * w_new;
      fm.w[j] = w_new;
    }
    size_t offset = X.cols; // :256-313
    for (size_t ri = 0; ri < relations.size(); ri++) {
      RelationBlock<Real> &rel = relations[ri];
      RelationCache<Real> &cache = relation_caches[ri];
      std::fill(cache.e.begin(), cache.e.end(), Real(0));
      rel.X.spmv(fm.w.data() + offset, cache.q.data());
      {
        size_t t = 0;
        for (size_t i : rel.original_to_block) {
          cache.e[i] += e_train[t];
          e_train[t++] -= cache.q[i];
        }
      }
      for (size_t l = 0; l < rel.feature_size; l++) {
        size_t g = cfg.group_index[offset + l];
        const Real w_old = fm.w[offset + l];
        Real lambda = hyper.lambda_w[g], mu = hyper.mu_w[g];
        const int64_t b = cache.X_t.ptr[l], en = cache.X_t.ptr[l + 1];
        Real square_term = 0;
        for (int64_t p = b; p < en; p++)
          square_term += (cache.X_t.val[p] * cache.X_t.val[p]) * cache.cardinality[cache.X_t.idx[p]];
        Real linear_term = 0;
        for (int64_t p = b; p < en; p++)
          linear_term += (-cache.X_t.val[p]) * cache.e[cache.X_t.idx[p]];
        linear_term += square_term * w_old;
        square_term = lambda + hyper.alpha * square_term;
        linear_term = hyper.alpha * linear_term + lambda * mu;
        Real w_new = sample_normal(square_term, linear_term);
        fm.w[offset + l] = w_new;
        for (int64_t p = b; p < en; p++) {
          int32_t s = cache.X_t.idx[p];
          cache.e[s] += (cache.X_t.val[p] * cache.cardinality[s]) * (w_new - w_old);
        }
      }
      rel.X.spmv(fm.w.data() + offset, cache.q.data());
      {
        size_t t = 0;
        for (size_t i : rel.original_to_block)
          e_train[t++] += cache.q[i];
      }
      offset += rel.feature_size;
    }
  }

  // FMTrainer.hpp:316-486
  void update_V(FM<Real> &fm, Hyper<Real> &hyper) {
    for (int r = 0; r < fm.n_factors; r++) {
      X.spmv(fm.vcol(r), q_train.data()); // :320
      {                                   // :323-340
        size_t offset = X.cols;
        for (size_t ri = 0; ri < relations.size(); ri++) {
          const RelationBlock<Real> &rel = relations[ri];
          RelationCache<Real> &cache = relation_caches[ri];
          rel.X.spmv(fm.vcol(r) + offset, cache.q.data());
          size_t t = 0;
          for (size_t i : rel.original_to_block)
            q_train[t++] += cache.q[i];
          offset += rel.feature_size;
        }
      }
      for (int64_t j = 0; j < X_t.rows; j++) { // :343-376
        size_t g = cfg.group_index[j];
        Real v_old = fm.v(j, r);
        Real square_coeff = 0, linear_coeff = 0;
        const int64_t b = X_t.ptr[j], en = X_t.ptr[j + 1];
        for (int64_t p = b; p < en; p++) {
          int32_t i = X_t.idx[p];
          Real x = X_t.val[p];
          auto h = x * (q_train[i] - x * v_old);
          square_coeff += h * h;
          linear_coeff += (-e_train[i]) * h;
        }
        linear_coeff += square_coeff * v_old;
        square_coeff *= hyper.alpha;
        linear_coeff *= hyper.alpha;
        square_coeff += hyper.lamV(g, r);
        linear_coeff += hyper.lamV(g, r) * hyper.muV(g, r);
        Real v_new = sample_normal(square_coeff, linear_coeff);
        fm.v(j, r) = v_new;
        for (int64_t p = b; p < en; p++) {
          int32_t i = X_t.idx[p];
          Real x = X_t.val[p];
          auto h = x * (q_train[i] - x * v_old);
          q_train[i] += x * (v_new - v_old);
          e_train[i] += h * (v_new - v_old);
        }
      }
      size_t offset = X.cols; // :378-482
      for (size_t ri = 0; ri < relations.size(); ri++) {
        const RelationBlock<Real> &rel = relations[ri];
        RelationCache<Real> &cache = relation_caches[ri];
        rel.X.spmv_sq(fm.vcol(r) + offset, cache.q_S.data());
        std::fill(cache.c.begin(), cache.c.end(), Real(0));
        std::fill(cache.c_S.begin(), cache.c_S.end(), Real(0));
        std::fill(cache.e.begin(), cache.e.end(), Real(0));
        std::fill(cache.e_q.begin(), cache.e_q.end(), Real(0));
        {
          size_t t = 0;
          for (size_t i : rel.original_to_block) { // :401-417
            Real temp = (q_train[t] - cache.q[i]);
            cache.c[i] += temp;
            cache.c_S[i] += temp * temp;
            cache.e[i] += e_train[t];
            cache.e_q[i] += e_train[t] * temp;
            q_train[t] -= cache.q[i];
            // the 0.5 literals promote this expression to double when Real == float
            e_train[t] -=
                (q_train[t] * cache.q[i] + 0.5 * cache.q[i] * cache.q[i] - 0.5 * cache.q_S[i]);
            t++;
          }
        }
        for (size_t l = 0; l < rel.feature_size; l++) { // :419-470
          size_t g = cfg.group_index[offset + l];
          Real v_old = fm.v(offset + l, r);
          Real square_coeff = 0, linear_coeff = 0;
          const int64_t b = cache.X_t.ptr[l], en = cache.X_t.ptr[l + 1];
          for (int64_t p = b; p < en; p++) {
            int32_t s = cache.X_t.idx[p];
            Real x_il = cache.X_t.val[p];
            auto h_B = (cache.q[s] - x_il * v_old);
            auto h_squared = h_B * h_B * cache.cardinality[s] + 2 * cache.c[s] * h_B + cache.c_S[s];
            h_squared = x_il * x_il * h_squared;
            square_coeff += h_squared;
            linear_coeff += (-cache.e[s] * h_B - cache.e_q[s]) * x_il;
          }
          linear_coeff += square_coeff * v_old;
          square_coeff *= hyper.alpha;
          linear_coeff *= hyper.alpha;
          square_coeff += hyper.lamV(g, r);
          linear_coeff += hyper.lamV(g, r) * hyper.muV(g, r);
          Real v_new = sample_normal(square_coeff, linear_coeff);
          Real delta = v_new - v_old;
          fm.v(offset + l, r) = v_new;
          for (int64_t p = b; p < en; p++) {
            int32_t s = cache.X_t.idx[p];
            const Real x_il = cache.X_t.val[p];
            auto h_B = cache.q[s] - x_il * v_old;
            cache.q[s] += delta * x_il;
            cache.q_S[s] += delta * (v_new + v_old) * x_il * x_il;
            cache.e[s] += x_il * delta * (h_B * cache.cardinality[s] + cache.c[s]);
            cache.e_q[s] += x_il * delta * (h_B * cache.c[s] + cache.c_S[s]);
          }
        }
        {
          size_t t = 0;
          for (size_t i : rel.original_to_block) { // :473-480
            e_train[t] +=
                (q_train[t] * cache.q[i] + 0.5 * cache.q[i] * cache.q[i] - 0.5 * cache.q_S[i]);
            q_train[t] += cache.q[i];
            t++;
          }
        }
        offset += rel.feature_size;
      }
    }
  }

  // FMTrainer.hpp:493-522
  void update_e(FM<Real> &fm) {
    fm.predict_score(e_train.data(), X, relations);
    if (cfg.task == Task::REGRESSION) {
      for (int r = 0; r < n_train; r++)
        e_train[r] -= y[r];
    } else if (cfg.task == Task::CLASSIFICATION) {
      Real zero = static_cast<Real>(0), sd = static_cast<Real>(1);
      for (int r = 0; r < n_train; r++) {
        Real gt = y[r], pred = e_train[r], n;
        if (gt > 0)
          n = tn_left<Real>(gen_, pred, sd, zero);
        else
          n = tn_right<Real>(gen_, pred, sd, zero);
        e_train[r] -= n;
      }
    } else {
      int i = 0;
      for (auto &sampler : cutpoint_sampler) {
        sampler.step();
        OprobitSampler<Real>::alpha_to_gamma(fm.cutpoints[i], sampler.alpha_now);
        sampler.sample_z_given_cutpoint();
        i++;
      }
    }
  }

  // BaseFMTrainer.hpp:135-152 — the fixed order of the nine steps
  void update_all(FM<Real> &fm, Hyper<Real> &hyper) {
    update_alpha(hyper);
    update_w0(fm, hyper);
    update_lambda_generic(hyper.mu_w.data(), hyper.lambda_w.data(), fm.w.data()); // :194-196
    update_mu_generic(hyper.mu_w.data(), hyper.lambda_w.data(), fm.w.data());     // :198-200
    update_w(fm, hyper);
    for (int r = 0; r < fm.n_factors; r++) // :202-208
      update_lambda_generic(&hyper.mu_V[cfg.n_groups * r], &hyper.lambda_V[cfg.n_groups * r],
                            fm.vcol(r));
    for (int r = 0; r < fm.n_factors; r++) // :210-216
      update_mu_generic(&hyper.mu_V[cfg.n_groups * r], &hyper.lambda_V[cfg.n_groups * r],
                        fm.vcol(r));
    update_V(fm, hyper);
    update_e(fm);
  }
};

} // namespace oracle
