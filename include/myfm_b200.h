/*
 * myfm_b200.h — C ABI of the B200-native Gibbs engine for Bayesian Factorization Machines.
 *
 * This is the drop-in boundary for the hot path of tohtsky/myFM (reference paths are relative to
 * the reference repository root).  The reference has no C ABI of its own: its boundary is the
 * pybind11 module `myfm._myfm` (cpp_source/declare_module.hpp:67-404).  Each entry point below is
 * what a binding for that module would call instead of the header-only C++ it calls today; the
 * file:line each one replaces is cited next to it, and INTEGRATION.md shows the pybind11 stub.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is HOST memory owned by the caller and may be
 *    released as soon as the call returns (the engine keeps its own device copies);
 *  - floating-point data crosses the boundary as float64, as in the reference's Python API
 *    (src/myfm/base.py:36,285-286); the engine converts to its compute dtype (f32 or f64);
 *  - sparse matrices are CSR with int64 row pointers and int32 column indices, taken as given
 *    (indices are neither sorted nor de-duplicated — same as the reference's scipy->Eigen caster);
 *  - dense matrices V, mu_V, lambda_V are ROW-major: V[j*rank + r], mu_V[g*rank + r] (the shapes
 *    numpy shows for the reference's FM.V / FMHyperParameters.mu_V);
 *  - every function returns MYFM_OK or an error code; myfm_last_error() returns the message of
 *    the last failure on the calling thread.  MYFM_ERR_INVALID_ARGUMENT corresponds to the
 *    reference's std::invalid_argument (-> ValueError), MYFM_ERR_RUNTIME to std::runtime_error
 *    (-> RuntimeError);
 *  - there is NO CPU fallback: a call that needs the GPU fails with MYFM_ERR_CUDA when no sm_100
 *    device is usable.
 */
#ifndef MYFM_B200_H
#define MYFM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MYFM_OK 0
#define MYFM_ERR_INVALID_ARGUMENT 1
#define MYFM_ERR_RUNTIME 2
#define MYFM_ERR_CUDA 3

/* FMLearningConfig::TASKTYPE, include/myfm/FMLearningConfig.hpp:14 */
#define MYFM_TASK_REGRESSION 0
#define MYFM_TASK_CLASSIFICATION 1
#define MYFM_TASK_ORDERED 2

/* compute dtype of the engine (the reference's `Real`: bind.cpp = double, bind_float.cpp = float) */
#define MYFM_DTYPE_F32 0
#define MYFM_DTYPE_F64 1

/* RNG contract.
 * MT19937: the engine consumes the libstdc++ std::mt19937(seed) stream draw for draw in the
 *          reference's order, so chains are comparable with the reference on
 *          the same seed.
 * PHILOX : the per-row latent draws of classification / ordered probit (FMTrainer.hpp:498-521,
 *          util.hpp:15-78) come from counter-based Philox4x32-10 streams on the device, keyed by
 *          (seed; global row, sweep, attempt); the sweep's Gaussian / Gamma variates still follow the
 *          mt19937 stream.  Statistically equivalent chain, no host involvement inside a sweep;
 *          regression chains are identical in both modes.  Row-sharded (world_size > 1)
 *          classification / ordered probit needs this mode. */
#define MYFM_RNG_MT19937 0
#define MYFM_RNG_PHILOX 1

typedef struct myfm_csr {
  int64_t n_rows;
  int64_t n_cols;
  const int64_t *indptr;  /* [n_rows + 1] */
  const int32_t *indices; /* [nnz] */
  const double *data;     /* [nnz] */
} myfm_csr_t;

/* relational::RelationBlock, include/myfm/definitions.hpp:30-52 */
typedef struct myfm_relation {
  const int64_t *original_to_block; /* [mapper_size], each in [0, block.n_rows) */
  int64_t mapper_size;
  myfm_csr_t block; /* block_size x feature_size */
} myfm_relation_t;

/* FMLearningConfig, include/myfm/FMLearningConfig.hpp:17-57 (built by ConfigBuilder,
 * cpp_source/declare_module.hpp:139-156).  Validation happens in myfm_config_validate / at
 * trainer creation with the reference's messages. */
typedef struct myfm_config {
  double alpha_0, beta_0, gamma_0, mu_0, reg_0;
  int32_t task_type;
  double nu_oprobit;
  int32_t fit_w0, fit_linear;
  int32_t n_iter, n_kept_samples;
  double cutpoint_scale;
  const int64_t *group_index; /* [n_group_index] group of every feature (main table then blocks) */
  int64_t n_group_index;
  int32_t n_cutpoint_groups;            /* ordered probit only */
  const int32_t *cutpoint_n_class;      /* [n_cutpoint_groups] */
  const int64_t *const *cutpoint_index; /* [n_cutpoint_groups] row ids of each group */
  const int64_t *cutpoint_index_len;    /* [n_cutpoint_groups] */
} myfm_config_t;

/* Engine options that have no counterpart in the reference. */
typedef struct myfm_engine_options {
  int32_t dtype;      /* MYFM_DTYPE_* */
  int32_t rng;        /* MYFM_RNG_* */
  int32_t device;     /* CUDA device ordinal */
  int32_t world_size; /* row-sharded data parallelism: number of ranks (1 = single GPU) */
  int32_t rank;       /* this process' rank */
  int64_t row_offset; /* global index of this shard's first training row */
  int64_t n_rows_global;
  const void *nccl_unique_id; /* 128-byte ncclUniqueId shared by all ranks, or NULL */
  /* Dependency level of every main-table column, agreed between the ranks (the conflict graph of
   * the GLOBAL matrix decides, a shard sees only part of it; myfm_b200/distributed.py runs the
   * consensus with myfm_level_schedule / myfm_level_relax).  NULL = compute from this shard. */
  const int32_t *column_level;
  int64_t n_column_level;
  /* Global index of every training row of this shard, in the order the shard's rows are given
   * (NULL: row_offset + i).  With MYFM_RNG_PHILOX the latent draw of a row is keyed by its GLOBAL
   * index, so a row-sharded classification / ordered-probit chain draws what the single-GPU chain
   * draws for the same row. */
  const int64_t *row_ids;
  int64_t n_row_ids;
} myfm_engine_options_t;

typedef struct myfm_trainer myfm_trainer_t;
typedef struct myfm_dataset myfm_dataset_t;
typedef struct myfm_sample myfm_sample_t; /* one posterior sample resident on the device */

const char *myfm_last_error(void);
/* Library / device probe: writes the number of usable CUDA devices; never fails without a GPU. */
int myfm_device_count(int32_t *count);
/* FMLearningConfig ctor checks, FMLearningConfig.hpp:29-56; writes the number of groups. */
int myfm_config_validate(const myfm_config_t *cfg, int32_t *n_groups);

/* ---- training: replaces create_train_fm, cpp_source/declare_module.hpp:30-45 ---------------- */

/* GibbsFMTrainer ctor (include/myfm/BaseFMTrainer.hpp:58-105): shape checks, X^T, relation
 * caches, mt19937(seed); uploads everything to HBM once. */
int myfm_trainer_create(myfm_trainer_t **out, const myfm_csr_t *X, int32_t n_relations,
                        const myfm_relation_t *relations, const double *y, int64_t n_y,
                        int32_t random_seed, const myfm_config_t *cfg,
                        const myfm_engine_options_t *opt);
void myfm_trainer_destroy(myfm_trainer_t *t);

/* create_FM + create_Hyper + initialize_hyper + initialize_e (BaseFMTrainer.hpp:107-115,
 * include/myfm/FM.hpp:34-45, include/myfm/FMTrainer.hpp:89-119). */
int myfm_trainer_init_fm(myfm_trainer_t *t, int32_t rank, double init_std);

/* n_sweeps x update_all (BaseFMTrainer.hpp:135-152: alpha, w0, lambda_w, mu_w, w, lambda_V,
 * mu_V, V, e).  Asynchronous for regression; returns once the work is enqueued. */
int myfm_trainer_step(myfm_trainer_t *t, int32_t n_sweeps);
int myfm_trainer_sync(myfm_trainer_t *t);
/* n_sweeps x update_all bracketed by CUDA events on the trainer's stream (synchronises on both
 * sides); writes the device-side milliseconds. */
int myfm_trainer_timed_steps(myfm_trainer_t *t, int32_t n_sweeps, double *ms);

/* Current state (synchronises).  w: [dim_all]; V: [dim_all x rank] row-major.  In get_fm any of
 * w0 / w / V may be NULL (that part is not copied). */
int myfm_trainer_dims(const myfm_trainer_t *t, int64_t *n_train, int64_t *dim_all, int32_t *rank,
                      int32_t *n_groups);
int myfm_trainer_get_fm(myfm_trainer_t *t, double *w0, double *w, double *V);
/* cutpoints of cutpoint group g (ordered probit): [n_class_g - 1] */
int myfm_trainer_get_cutpoints(myfm_trainer_t *t, int32_t g, double *out);
/* FMHyperParameters (include/myfm/HyperParams.hpp:8-37); mu_V/lambda_V [n_groups x rank] row-major */
int myfm_trainer_get_hyper(myfm_trainer_t *t, double *alpha, double *mu_w, double *lambda_w,
                           double *mu_V, double *lambda_V);
/* residual cache e_train and factor cache q_train of this shard (tests / diagnostics) */
int myfm_trainer_get_e(myfm_trainer_t *t, double *e);
int myfm_trainer_get_q(myfm_trainer_t *t, double *q);
/* Overwrite parts of the chain state (any pointer may be NULL = keep): warm start from a stored
 * sample, and the per-sweep "teacher-forced" parity tests.  Layouts as in the getters above. */
int myfm_trainer_set_state(myfm_trainer_t *t, const double *w0, const double *w, const double *V,
                           const double *alpha, const double *mu_w, const double *lambda_w,
                           const double *mu_V, const double *lambda_V, const double *e);
/* The standardised variates (N(0,1) / Gamma(shape,1)) the most recent sweep consumed, in the
 * reference's draw order (BaseFMTrainer.hpp:135-152); writes their number to *n and copies them
 * when capacity suffices.  RNG tests compare them with myfm_rng_fill. */
int myfm_trainer_get_variates(myfm_trainer_t *t, double *out, int64_t capacity, int64_t *n);
/* OprobitSampler::accept_count of cutpoint group g (FMTrainer.hpp:83-85) */
int myfm_trainer_mh_accept(myfm_trainer_t *t, int32_t g, int64_t *count);
/* which schedule the main-table column sweeps use: 0 = general dependency-level kernels,
 * 1 = field path (streaming level 0 + gather-only last level, csrc/field_sweep.cuh), 2 = field path
 * on row shards with the column statistics exchanged through peer memory (NVLink); 3 / 4 = 1 / 2
 * with rank-exclusive level-0 columns (no exchange for the streaming level).  All are the
 * reference's update_w / update_V (FMTrainer.hpp:231-486); diagnostics and tests only. */
int myfm_trainer_sweep_path(const myfm_trainer_t *t, int32_t *path);
/* number of kernels this trainer has launched so far (bench.py's gpu_launches) */
int myfm_trainer_launch_count(const myfm_trainer_t *t, int64_t *count);
/* device milliseconds spent in the dominant kernel family since the last call (CUDA events on
 * the trainer's stream); family: 0 = column sweeps, 1 = q_init, 2 = e_refresh, and inside family 0 on
 * the field path: 3 = streaming level (k_field_stream), 4 = gather-only level (k_field_stats) */
int myfm_trainer_kernel_ms(myfm_trainer_t *t, int32_t family, double *ms, int64_t *launches);
/* turn the per-family CUDA-event timing on or off (off by default; adds two event records per
 * launch group when on) */
int myfm_trainer_set_profiling(myfm_trainer_t *t, int32_t on);

/* ---- prediction: replaces FM::predict_score / Predictor::predict* ---------------------------- */

/* A design matrix (+ relation blocks) resident on the device, reusable across calls. */
int myfm_dataset_create(myfm_dataset_t **out, const myfm_csr_t *X, int32_t n_relations,
                        const myfm_relation_t *relations, int32_t dtype, int32_t device);
void myfm_dataset_destroy(myfm_dataset_t *d);

/* FM::predict_score (include/myfm/FM.hpp:47-136) for one sample given on the host.
 * out: [n_rows].  dim_all must equal the dataset's total feature size (else INVALID_ARGUMENT). */
int myfm_predict_score(const myfm_dataset_t *d, double w0, const double *w, const double *V,
                       int64_t dim_all, int32_t rank, double *out);

/* Predictor::predict / predict_parallel (include/myfm/predictor.hpp:35-76,126-147): mean over
 * n_samples of score (REGRESSION) or Phi(score) (CLASSIFICATION).  w0s [n_samples],
 * ws [n_samples x dim_all], Vs [n_samples x dim_all x rank].  out: [n_rows]. */
int myfm_predict_mean(const myfm_dataset_t *d, int32_t task_type, int32_t n_samples,
                      const double *w0s, const double *ws, const double *Vs, int64_t dim_all,
                      int32_t rank, double *out);

/* FM::oprobit_predict_proba averaged over samples (predictor.hpp:78-124, FM.hpp:137-162).
 * cutpoints: [n_samples x n_cpt]; out: [n_rows x (n_cpt + 1)] row-major. */
int myfm_predict_oprobit_mean(const myfm_dataset_t *d, int32_t n_samples, const double *w0s,
                              const double *ws, const double *Vs, const double *cutpoints,
                              int32_t n_cpt, int64_t dim_all, int32_t rank, double *out);

/* Kept samples that never leave the device.  GibbsFMTrainer::learn_with_callback copies the FM into
 * the predictor for each of the last n_kept_samples iterations (FMTrainer.hpp:74-76);
 * myfm_trainer_snapshot is that copy, device to device.  myfm_sample_get reads it back (any of
 * w0 / w / V may be NULL; V is [dim_all x rank] row-major).  myfm_predict_samples_mean is
 * Predictor::predict / predict_parallel / predict_parallel_oprobit (predictor.hpp:35-147) over such
 * samples: n_cpt < 0 for regression / classification, else cutpoints is [n_samples x n_cpt] and out
 * is [n_rows x (n_cpt + 1)]. */
int myfm_trainer_snapshot(myfm_trainer_t *t, myfm_sample_t **out);
void myfm_sample_destroy(myfm_sample_t *s);
int myfm_sample_get(const myfm_sample_t *s, double *w0, double *w, double *V);
int myfm_predict_samples_mean(const myfm_dataset_t *d, int32_t task_type, int32_t n_samples,
                              myfm_sample_t *const *samples, const double *cutpoints, int32_t n_cpt,
                              double *out);

/* FM::predict_score with the trainer's CURRENT device-resident sample (no weight upload); used
 * by per-iteration callbacks (src/myfm/utils/callbacks/libfm.py:82-113). */
int myfm_trainer_predict_score(myfm_trainer_t *t, const myfm_dataset_t *d, double *out);

/* ---- per-iteration evaluation callbacks on the device ------------------------------------------
 * The reference's LibFM-style callbacks (src/myfm/utils/callbacks/libfm.py:57-262) call
 * fm.predict_score / fm.oprobit_predict_proba on a held-out set every iteration and keep running
 * means in numpy.  An evaluator keeps the test matrix (a myfm_dataset_t, not owned), the targets,
 * both running sums (all sweeps; all but the first five) and the metric reductions on the device.
 * task_type REGRESSION: clip_min / clip_max clip the running means (NaN = no clipping);
 * CLASSIFICATION: y_test in {0, 1}, eps clips the running means to [eps, 1 - eps] (eps < 0: none);
 * ORDERED: y_test = class index in [0, n_class), eps is the floor of the picked probability.
 * myfm_evaluator_step runs one callback step with the trainer's CURRENT sample and writes 9 sums
 * over the test rows, in the order (running mean, this sweep, mean of all but the first five):
 *   REGRESSION      [0..2] squared error
 *   CLASSIFICATION  [0..2] negative log-likelihood, [3..5] correct decisions
 *   ORDERED         [0..2] negative log-likelihood, [3..5] correct arg-max, [6..8] squared error of
 *                   the expected class
 * (the third of each group is 0 while iteration < 5).  cutpoints: the sample's cut-points
 * (ORDERED, n_cpt = n_class - 1).  myfm_evaluator_get_sums copies the running sums
 * ([n_test x width], width = 1 or n_class; either pointer may be NULL). */
typedef struct myfm_evaluator myfm_evaluator_t;
int myfm_evaluator_create(myfm_evaluator_t **out, const myfm_dataset_t *d, const double *y_test,
                          int64_t n_test, int32_t task_type, int32_t n_class, double clip_min,
                          double clip_max, double eps);
void myfm_evaluator_destroy(myfm_evaluator_t *e);
int myfm_evaluator_step(myfm_evaluator_t *e, myfm_trainer_t *t, int32_t iteration,
                        const double *cutpoints, int32_t n_cpt, double *terms);
int myfm_evaluator_get_sums(myfm_evaluator_t *e, double *sum, double *late);

/* ---- host-side pieces that run without a GPU (exercised by the CPU test-suite) -------------- */

/* The engine's MT19937 variate stream for one regression sweep layout: fills `out` with the
 * standardised variates the device consumes, in stream order. */
int myfm_rng_fill(int32_t dtype, int32_t seed, int64_t n_skip_normals_persistent,
                  const int32_t *kinds, const double *shapes, int64_t n, double *out);

/* Dependency-level schedule of the columns of a CSR matrix: level[j] for
 * every column; columns of one level are pairwise row-disjoint and running levels in order
 * reproduces the reference's serial column order exactly.  Returns the number of levels. */
int myfm_level_schedule(const myfm_csr_t *X, int32_t *level, int32_t *n_levels);
/* Same recurrence with `level` as lower bounds on entry: level[j] = max(level[j], 1 + max level of
 * an earlier column sharing a row with j in X).  Row-sharded ranks alternate this with an
 * element-wise MAX all-reduce until nothing changes; the fixed point is the schedule of the
 * global matrix.  *changed = 1 when any entry grew. */
int myfm_level_relax(const myfm_csr_t *X, int32_t *level, int32_t *n_levels, int32_t *changed);
/* Test hooks for the host-side data preparation of myfm_trainer_create (csrc/host_data.hpp; no
 * GPU needed).  myfm_host_transpose: the CSC arrays of X (entries of a column in ascending row
 * order, BaseFMTrainer.hpp:61), indptr_out [n_cols + 1], indices_out / data_out [nnz].
 * myfm_set_host_threads: number of threads the preparation passes use (0 = default). */
int myfm_host_transpose(const myfm_csr_t *X, int64_t *indptr_out, int32_t *indices_out, double *data_out);
int myfm_set_host_threads(int32_t n);
/* Jump-ahead polynomial of std::mt19937 used by the parallel device generator (csrc/mt_jump.hpp,
 * csrc/mt_device.cuh: k_mt_farm): the exponents i with g_i = 1 of g = t^n mod phi, phi the
 * characteristic polynomial of MT19937, so that out[a + n + j] = XOR_i out[a + i + j] for the
 * generator's output words.  Writes their number to *n_taps and, when capacity suffices, the
 * exponents (ascending) to taps_out. */
int myfm_mt_jump_taps(uint64_t n, uint16_t *taps_out, int32_t capacity, int32_t *n_taps);
/* ncclGetUniqueId for the row-sharded trainer: rank 0 calls it and ships the 128 bytes to the
 * other ranks by any side channel (myfm_b200/distributed.py uses torch.distributed). */
int myfm_nccl_unique_id(void *out128);

#ifdef __cplusplus
}
#endif
#endif /* MYFM_B200_H */
