// ORACLE — TEST INFRASTRUCTURE ONLY (see fm_oracle.hpp).  Plain C entry points so that tests/,
// smoke() and bench.py's CPU-baseline legs can drive the restatement through ctypes.  All
// floating-point arrays cross this interface as double; Real == float instances convert at the
// edge (float -> double is exact, so nothing is hidden).
#include "fm_oracle.hpp"

#include <chrono>
#include <cstring>
#include <memory>

namespace {

thread_local std::string g_error;

struct RelationDesc { // mirrors the Python-side ctypes.Structure
  const int64_t *original_to_block;
  int64_t mapper_size;
  int64_t block_size;
  int64_t feature_size;
  const int64_t *indptr;
  const int32_t *indices;
  const double *data;
};

struct ConfigDesc {
  double alpha_0, beta_0, gamma_0, mu_0, reg_0;
  int32_t task_type;
  double nu_oprobit;
  int32_t fit_w0, fit_linear;
  int32_t n_iter, n_kept_samples;
  double cutpoint_scale;
  const int64_t *group_index;
  int64_t n_group_index;
  int32_t n_cutpoint_groups;
  const int32_t *cutpoint_n_class;      // [n_cutpoint_groups]
  const int64_t *const *cutpoint_index; // [n_cutpoint_groups] -> row ids
  const int64_t *cutpoint_index_len;    // [n_cutpoint_groups]
};

template <typename Real>
oracle::Csr<Real> make_csr(int64_t rows, int64_t cols, const int64_t *indptr, const int32_t *indices,
                           const double *data) {
  oracle::Csr<Real> m;
  m.rows = rows;
  m.cols = cols;
  m.ptr.assign(indptr, indptr + rows + 1);
  const int64_t nnz = indptr[rows];
  m.idx.assign(indices, indices + nnz);
  m.val.resize(nnz);
  for (int64_t i = 0; i < nnz; i++)
    m.val[i] = static_cast<Real>(data[i]);
  return m;
}

template <typename Real>
std::vector<oracle::RelationBlock<Real>> make_relations(int n_rel, const RelationDesc *rels) {
  std::vector<oracle::RelationBlock<Real>> out;
  for (int b = 0; b < n_rel; b++) {
    const RelationDesc &d = rels[b];
    std::vector<size_t> map(d.mapper_size);
    for (int64_t i = 0; i < d.mapper_size; i++)
      map[i] = static_cast<size_t>(d.original_to_block[i]);
    out.emplace_back(std::move(map),
                     make_csr<Real>(d.block_size, d.feature_size, d.indptr, d.indices, d.data));
  }
  return out;
}

oracle::Config make_config(const ConfigDesc *c) {
  oracle::Config cfg;
  cfg.alpha_0 = c->alpha_0;
  cfg.beta_0 = c->beta_0;
  cfg.gamma_0 = c->gamma_0;
  cfg.mu_0 = c->mu_0;
  cfg.reg_0 = c->reg_0;
  cfg.task = static_cast<oracle::Task>(c->task_type);
  cfg.nu_oprobit = c->nu_oprobit;
  cfg.fit_w0 = c->fit_w0 != 0;
  cfg.fit_linear = c->fit_linear != 0;
  cfg.n_iter = c->n_iter;
  cfg.n_kept_samples = c->n_kept_samples;
  cfg.cutpoint_scale = c->cutpoint_scale;
  cfg.group_index.assign(c->group_index, c->group_index + c->n_group_index);
  for (int g = 0; g < c->n_cutpoint_groups; g++) {
    std::vector<size_t> rows(c->cutpoint_index_len[g]);
    for (int64_t i = 0; i < c->cutpoint_index_len[g]; i++)
      rows[i] = static_cast<size_t>(c->cutpoint_index[g][i]);
    cfg.cutpoint_groups.emplace_back(static_cast<size_t>(c->cutpoint_n_class[g]), std::move(rows));
  }
  cfg.finalize();
  return cfg;
}

struct ChainBase {
  virtual ~ChainBase() = default;
  virtual void step() = 0;
  virtual double timed_steps(int n) = 0;
  virtual int64_t dim_all() const = 0;
  virtual int rank() const = 0;
  virtual int n_groups() const = 0;
  virtual int64_t n_train() const = 0;
  virtual void get_fm(double *w0, double *w, double *V) const = 0;
  virtual int n_cutpoint_groups() const = 0;
  virtual int cutpoint_len(int g) const = 0;
  virtual void get_cutpoints(int g, double *out) const = 0;
  virtual void get_hyper(double *alpha, double *mu_w, double *lambda_w, double *mu_V,
                         double *lambda_V) const = 0;
  virtual void get_e(double *e) const = 0;
  virtual void get_q(double *q) const = 0;
  virtual int64_t mh_accept(int g) const = 0;
};

template <typename Real> struct Chain : ChainBase {
  oracle::Trainer<Real> trainer;
  oracle::FM<Real> fm;
  oracle::Hyper<Real> hyper;

  // cpp_source/declare_module.hpp:30-45 up to (and including) the two initialisers at the top
  // of learn_with_callback (FMTrainer.hpp:64-65).
  Chain(oracle::Csr<Real> X, std::vector<oracle::RelationBlock<Real>> rels, std::vector<Real> y,
        int seed, oracle::Config cfg, int rank_, double init_std)
      : trainer(std::move(X), std::move(rels), std::move(y), seed, std::move(cfg)),
        fm(trainer.create_FM(rank_, static_cast<Real>(init_std))),
        hyper(trainer.create_Hyper(rank_)) {
    trainer.initialize_hyper(hyper);
    trainer.initialize_e(fm);
  }
  void step() override { trainer.update_all(fm, hyper); }
  double timed_steps(int n) override {
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; i++)
      trainer.update_all(fm, hyper);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  int64_t dim_all() const override { return static_cast<int64_t>(trainer.dim_all); }
  int rank() const override { return fm.n_factors; }
  int n_groups() const override { return static_cast<int>(trainer.cfg.n_groups); }
  int64_t n_train() const override { return trainer.n_train; }
  void get_fm(double *w0, double *w, double *V) const override {
    *w0 = fm.w0;
    for (size_t i = 0; i < fm.w.size(); i++)
      w[i] = fm.w[i];
    for (size_t i = 0; i < fm.V.size(); i++)
      V[i] = fm.V[i]; // column-major (dim_all x rank)
  }
  int n_cutpoint_groups() const override { return static_cast<int>(fm.cutpoints.size()); }
  int cutpoint_len(int g) const override { return static_cast<int>(fm.cutpoints.at(g).size()); }
  void get_cutpoints(int g, double *out) const override {
    for (size_t i = 0; i < fm.cutpoints.at(g).size(); i++)
      out[i] = fm.cutpoints[g][i];
  }
  void get_hyper(double *alpha, double *mu_w, double *lambda_w, double *mu_V,
                 double *lambda_V) const override {
    *alpha = hyper.alpha;
    for (size_t i = 0; i < hyper.mu_w.size(); i++) {
      mu_w[i] = hyper.mu_w[i];
      lambda_w[i] = hyper.lambda_w[i];
    }
    for (size_t i = 0; i < hyper.mu_V.size(); i++) { // column-major (n_groups x rank)
      mu_V[i] = hyper.mu_V[i];
      lambda_V[i] = hyper.lambda_V[i];
    }
  }
  void get_e(double *e) const override {
    for (size_t i = 0; i < trainer.e_train.size(); i++)
      e[i] = trainer.e_train[i];
  }
  void get_q(double *q) const override {
    for (size_t i = 0; i < trainer.q_train.size(); i++)
      q[i] = trainer.q_train[i];
  }
  int64_t mh_accept(int g) const override {
    return static_cast<int64_t>(trainer.cutpoint_sampler.at(g).accept_count);
  }
};

template <typename Real>
ChainBase *create_chain(int rank, double init_std, int64_t rows, int64_t cols, const int64_t *indptr,
                        const int32_t *indices, const double *data, int n_rel,
                        const RelationDesc *rels, const double *y, int64_t y_len, int seed,
                        const ConfigDesc *cfg) {
  std::vector<Real> yv(y_len);
  for (int64_t i = 0; i < y_len; i++)
    yv[i] = static_cast<Real>(y[i]);
  return new Chain<Real>(make_csr<Real>(rows, cols, indptr, indices, data),
                         make_relations<Real>(n_rel, rels), std::move(yv), seed, make_config(cfg),
                         rank, init_std);
}

template <typename Real>
void predict_score_impl(double w0, const double *w, const double *V, int64_t dim_all, int rank,
                        int64_t rows, int64_t cols, const int64_t *indptr, const int32_t *indices,
                        const double *data, int n_rel, const RelationDesc *rels, double *out) {
  oracle::FM<Real> fm(rank);
  fm.n_features = dim_all;
  fm.w0 = static_cast<Real>(w0);
  fm.w.resize(dim_all);
  fm.V.resize(dim_all * static_cast<size_t>(rank));
  for (int64_t i = 0; i < dim_all; i++)
    fm.w[i] = static_cast<Real>(w[i]);
  for (size_t i = 0; i < fm.V.size(); i++)
    fm.V[i] = static_cast<Real>(V[i]);
  auto X = make_csr<Real>(rows, cols, indptr, indices, data);
  auto relations = make_relations<Real>(n_rel, rels);
  std::vector<Real> target(rows);
  fm.predict_score(target.data(), X, relations);
  for (int64_t i = 0; i < rows; i++)
    out[i] = target[i];
}

} // namespace

#define ORACLE_TRY try {
#define ORACLE_CATCH(ret)                                                                          \
  }                                                                                                \
  catch (const std::invalid_argument &ex) {                                                        \
    g_error = std::string("invalid_argument: ") + ex.what();                                       \
    return ret;                                                                                    \
  }                                                                                                \
  catch (const std::exception &ex) {                                                               \
    g_error = std::string("runtime_error: ") + ex.what();                                          \
    return ret;                                                                                    \
  }

extern "C" {

const char *oracle_last_error() { return g_error.c_str(); }

// dtype: 0 = float (the reference's bind_float.cpp instantiation), 1 = double (what it ships)
void *oracle_chain_create(int dtype, int rank, double init_std, int64_t rows, int64_t cols,
                          const int64_t *indptr, const int32_t *indices, const double *data,
                          int n_rel, const RelationDesc *rels, const double *y, int64_t y_len,
                          int seed, const ConfigDesc *cfg) {
  ORACLE_TRY
  if (dtype == 0)
    return create_chain<float>(rank, init_std, rows, cols, indptr, indices, data, n_rel, rels, y,
                               y_len, seed, cfg);
  return create_chain<double>(rank, init_std, rows, cols, indptr, indices, data, n_rel, rels, y,
                              y_len, seed, cfg);
  ORACLE_CATCH(nullptr)
}

void oracle_chain_destroy(void *h) { delete static_cast<ChainBase *>(h); }

int oracle_chain_step(void *h) {
  ORACLE_TRY
  static_cast<ChainBase *>(h)->step();
  return 0;
  ORACLE_CATCH(-1)
}

// runs n update_all sweeps and returns the wall-clock seconds spent inside them only
double oracle_chain_timed_steps(void *h, int n) {
  ORACLE_TRY
  return static_cast<ChainBase *>(h)->timed_steps(n);
  ORACLE_CATCH(-1.0)
}

int64_t oracle_chain_dim_all(void *h) { return static_cast<ChainBase *>(h)->dim_all(); }
int oracle_chain_rank(void *h) { return static_cast<ChainBase *>(h)->rank(); }
int oracle_chain_n_groups(void *h) { return static_cast<ChainBase *>(h)->n_groups(); }
int64_t oracle_chain_n_train(void *h) { return static_cast<ChainBase *>(h)->n_train(); }
void oracle_chain_get_fm(void *h, double *w0, double *w, double *V) {
  static_cast<ChainBase *>(h)->get_fm(w0, w, V);
}
int oracle_chain_n_cutpoint_groups(void *h) {
  return static_cast<ChainBase *>(h)->n_cutpoint_groups();
}
int oracle_chain_cutpoint_len(void *h, int g) { return static_cast<ChainBase *>(h)->cutpoint_len(g); }
void oracle_chain_get_cutpoints(void *h, int g, double *out) {
  static_cast<ChainBase *>(h)->get_cutpoints(g, out);
}
void oracle_chain_get_hyper(void *h, double *alpha, double *mu_w, double *lambda_w, double *mu_V,
                            double *lambda_V) {
  static_cast<ChainBase *>(h)->get_hyper(alpha, mu_w, lambda_w, mu_V, lambda_V);
}
void oracle_chain_get_e(void *h, double *e) { static_cast<ChainBase *>(h)->get_e(e); }
void oracle_chain_get_q(void *h, double *q) { static_cast<ChainBase *>(h)->get_q(q); }
int64_t oracle_chain_mh_accept(void *h, int g) { return static_cast<ChainBase *>(h)->mh_accept(g); }

int oracle_predict_score(int dtype, double w0, const double *w, const double *V, int64_t dim_all,
                         int rank, int64_t rows, int64_t cols, const int64_t *indptr,
                         const int32_t *indices, const double *data, int n_rel,
                         const RelationDesc *rels, double *out) {
  ORACLE_TRY
  if (dtype == 0)
    predict_score_impl<float>(w0, w, V, dim_all, rank, rows, cols, indptr, indices, data, n_rel,
                              rels, out);
  else
    predict_score_impl<double>(w0, w, V, dim_all, rank, rows, cols, indptr, indices, data, n_rel,
                               rels, out);
  return 0;
  ORACLE_CATCH(-1)
}

// libstdc++ known-answer probes: first n draws of a PERSISTENT
// normal_distribution over mt19937(seed).
void oracle_kat_normal(int dtype, int seed, int n, double *out) {
  std::mt19937 gen(seed);
  if (dtype == 0) {
    std::normal_distribution<float> nd;
    for (int i = 0; i < n; i++)
      out[i] = nd(gen);
  } else {
    std::normal_distribution<double> nd;
    for (int i = 0; i < n; i++)
      out[i] = nd(gen);
  }
}

// Truncated-normal samplers (util.hpp:15-78) on a fresh mt19937(seed); kind 0 = left(a),
// 1 = right(a), 2 = twoside(a, b).
void oracle_tn_draws(int dtype, int seed, int kind, double a, double b, int n, double *out) {
  std::mt19937 gen(seed);
  for (int i = 0; i < n; i++) {
    if (dtype == 0) {
      float fa = static_cast<float>(a), fb = static_cast<float>(b);
      out[i] = kind == 0   ? oracle::tn_left<float>(gen, fa)
               : kind == 1 ? oracle::tn_right<float>(gen, fa)
                           : oracle::tn_twoside<float>(gen, fa, fb);
    } else {
      out[i] = kind == 0   ? oracle::tn_left<double>(gen, a)
               : kind == 1 ? oracle::tn_right<double>(gen, a)
                           : oracle::tn_twoside<double>(gen, a, b);
    }
  }
}

double oracle_erfcx(double x) { return Faddeeva::erfcx(x); }

} // extern "C"
