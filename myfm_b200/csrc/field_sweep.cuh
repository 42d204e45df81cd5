// Column sweeps of a main table that is a stack of L position-aligned fields (every row has exactly
// L entries and its k-th entry belongs to dependency level k: one-hot / categorical encodings,
// `group_shapes` in the Python API) — the shape of every MovieLens-style workload of the reference.
//
// Measured on B200 (profiles/r01b_summary.md): a random read-modify-write scatter of the {e,q}
// pairs costs 160 us per 10 M entries, a pure gather 46 us, a streaming pass with a shared-memory
// table lookup per row 52 us.  So the sweep of a factor column (FMTrainer.hpp:343-376) is
// re-associated — same arithmetic per element, different place and time:
//
//   level 0 (rows are contiguous per column after the row reordering): ONE streaming pass
//     k_field_stream that (a) applies the rank-1 update the LAST level of the previous vector left
//     pending (looked up per row in a shared-memory table {theta_old, theta_new} of that level's
//     columns), (b) forms q_i = x_i . V[:, r] on the fly (q_init, FMTrainer.hpp:320, fused — the
//     last field's factor values sit in the same shared-memory table), (c) reduces the column
//     statistics, draws, and writes e and q back.  One read and one write of {e,q} per row.
//   levels 1 .. L-2: the gather / scatter kernels of kernels.cuh (k_level_sweep).
//   level L-1: statistics by GATHER ONLY (k_field_stats); the draw is stored and the update of
//     e is deferred to the next streaming pass.  The forward pass at the end of update_all
//     (FMTrainer.hpp:493-522) overwrites e, so the update left pending by the last factor is dropped.
//
// Every element-wise expression is the one of kernels.cuh (and of the reference); only the
// place where the last level's update lands differs.  q is only defined inside a vector's sweep.
#pragma once

#include "kernels.cuh"

namespace myfm {

constexpr int FIELD_THREADS = 1024; // streaming pass: one persistent CTA per SM
constexpr int FIELD_R = 8;          // rows per lane kept in registers between reduction and update
constexpr int FIELD_WARP_MAX = 1024;  // longest level-0 column handled by one warp
constexpr int FIELD_CTA_MAX = 32768;  // longest level-0 column (one CTA); longer: general path
constexpr int STATS_THREADS = 256;
constexpr int STATS_WARP_MAX = 256;   // last level: warp per column up to here,
constexpr int STATS_CHUNK = 8192;     // one CTA per column up to here, chunks of this size beyond

enum { PEND_NONE = 0, PEND_W = 1, PEND_V = 2 };

template <typename Real> struct FieldStreamArgs {
  const int4 *item; // level-0 columns {column, first row, end row, -}: nC CTA-wide, then nW per warp
  int nC, nW;
  int *sched;       // work counter of the warp items (zero at launch)
  Pair<Real> *eq;
  int64_t n_rows;
  int n_tail;           // L - 1
  const int *tail_idx;  // [n_tail][n_rows]: column of the row's entries 1 .. L-1
  const Real *tail_val; // [n_tail][n_rows], unused when UNIT
  const Real *own_val;  // [n_rows] value of the level-0 entry, unused when UNIT
  Real *theta;          // w, or column r of V (column-major)
  Real *theta_t;        // feature-major mirror: theta_t[j * t_stride], or nullptr
  int64_t t_stride;
  const Real *z;
  const int *group;
  const Real *alpha, *lambda, *mu;
  int last_base, n_tab; // the last level's columns are [last_base, last_base + n_tab)
  const Real *pend_told, *pend_tnew; // [n_tab] draw left pending by the previous vector
};

template <typename Real, bool IS_V, bool UNIT, int PEND> struct FieldRow {
  const FieldStreamArgs<Real> &a;
  const Real *s_told, *s_tnew, *s_tnext;

  // Row i as the level-0 column sees it: e with the pending update applied, q = x_i . V[:, r]
  // (IS_V) and the row's level-0 value x0.
  __device__ __forceinline__ void operator()(int64_t i, Real theta_old, Real &e, Real &q, Real &x0) const {
    const Pair<Real> v = __ldcg(a.eq + i);
    const int nt = a.n_tail;
    int jl = 0;
    Real xl = 1;
    if (IS_V || PEND != PEND_NONE) {
      jl = __ldcs(a.tail_idx + static_cast<int64_t>(nt - 1) * a.n_rows + i) - a.last_base;
      if (!UNIT)
        xl = __ldcs(a.tail_val + static_cast<int64_t>(nt - 1) * a.n_rows + i);
    }
    x0 = UNIT ? Real(1) : __ldcs(a.own_val + i);
    e = v.x, q = v.y;
    if (PEND == PEND_V) { // FMTrainer.hpp:366-374 of the previous factor's last-level column
      const Real told = s_told[jl], tnew = s_tnew[jl];
      const Real h = xl * (q - xl * told);
      e = e + h * (tnew - told);
    } else if (PEND == PEND_W) { // FMTrainer.hpp:240,251
      const Real told = s_told[jl], tnew = s_tnew[jl];
      e = (e - xl * told) + xl * tnew;
    }
    if (IS_V) { // q_init in CSR order (FMTrainer.hpp:320)
      Real acc = x0 * theta_old;
      for (int k = 0; k + 1 < nt; k++) {
        const int j = __ldcs(a.tail_idx + static_cast<int64_t>(k) * a.n_rows + i);
        const Real x = UNIT ? Real(1) : __ldcs(a.tail_val + static_cast<int64_t>(k) * a.n_rows + i);
        acc += x * a.theta[j];
      }
      acc += xl * s_tnext[jl];
      q = acc;
    }
  }
};

template <typename Real, bool IS_V>
__device__ __forceinline__ void field_stats(Real e, Real q, Real x0, Real theta_old, Real alpha, Real &sq,
                                            Real &lin) {
  if (IS_V) {
    const Real h = x0 * (q - x0 * theta_old);
    sq += h * h;
    lin += (-e) * h;
  } else {
    const Real e1 = e - x0 * theta_old;
    sq += x0 * x0;
    lin += ((-alpha) * x0) * e1;
  }
}

template <typename Real, bool IS_V>
__device__ __forceinline__ Pair<Real> field_update(Real e, Real q, Real x0, Real theta_old, Real theta_new) {
  Pair<Real> v;
  if (IS_V) {
    const Real h = x0 * (q - x0 * theta_old);
    v.y = q + x0 * (theta_new - theta_old);
    v.x = e + h * (theta_new - theta_old);
  } else {
    const Real e1 = e - x0 * theta_old;
    v.x = e1 + x0 * theta_new;
    v.y = q;
  }
  return v;
}

template <typename Real, bool IS_V, bool UNIT, int PEND>
__global__ void __launch_bounds__(FIELD_THREADS, 1) k_field_stream(FieldStreamArgs<Real> a) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ Real scratch[32];
  Real *s_told = reinterpret_cast<Real *>(s_raw), *s_tnew = s_told + a.n_tab, *s_tnext = s_tnew + a.n_tab;
  for (int t = threadIdx.x; t < a.n_tab; t += FIELD_THREADS) {
    if (PEND != PEND_NONE)
      s_told[t] = a.pend_told[t], s_tnew[t] = a.pend_tnew[t];
    if (IS_V)
      s_tnext[t] = a.theta[a.last_base + t];
  }
  __syncthreads();
  const FieldRow<Real, IS_V, UNIT, PEND> row{a, s_told, s_tnew, s_tnext};
  const Real alpha = *a.alpha;
  const int lane = threadIdx.x & 31;

  // long columns: the whole CTA, two passes over the rows (the second re-reads through L2)
  for (int c = blockIdx.x; c < a.nC; c += gridDim.x) {
    const int4 it = __ldg(a.item + c);
    const int j = it.x;
    const Real theta_old = a.theta[j];
    const int g = a.group[j];
    const Real lam = a.lambda[g], mu = a.mu[g], z = a.z[j];
    Real sq = 0, lin = 0;
    for (int64_t i = it.y + threadIdx.x; i < it.z; i += FIELD_THREADS) {
      Real e, q, x0;
      row(i, theta_old, e, q, x0);
      field_stats<Real, IS_V>(e, q, x0, theta_old, alpha, sq, lin);
    }
    sq = block_sum(sq, scratch);
    lin = block_sum(lin, scratch);
    const Real theta_new = column_draw<Real, IS_V>(sq, lin, theta_old, alpha, lam, mu, z);
    for (int64_t i = it.y + threadIdx.x; i < it.z; i += FIELD_THREADS) {
      Real e, q, x0;
      row(i, theta_old, e, q, x0);
      __stcg(a.eq + i, field_update<Real, IS_V>(e, q, x0, theta_old, theta_new));
    }
    __syncthreads(); // every thread has read theta[j]
    if (threadIdx.x == 0) {
      a.theta[j] = theta_new;
      if (a.theta_t)
        a.theta_t[static_cast<int64_t>(j) * a.t_stride] = theta_new;
    }
  }

  // the rest: one warp per column, longest first, handed out through a counter
  int k = 0;
  if (lane == 0)
    k = atomicAdd(a.sched, 1);
  k = __shfl_sync(FULL_MASK, k, 0);
  while (k < a.nW) {
    int k_next = 0;
    if (lane == 0)
      k_next = atomicAdd(a.sched, 1);
    const int4 it = __ldg(a.item + a.nC + k);
    const int j = it.x, n = it.z - it.y;
    const Real theta_old = a.theta[j];
    const int g = a.group[j];
    const Real lam = a.lambda[g], mu = a.mu[g], z = a.z[j];
    Real sq = 0, lin = 0;
    if (n <= 32 * FIELD_R) { // rows stay in registers between the reduction and the update
      Real e[FIELD_R], q[FIELD_R], x0[FIELD_R];
      const int n_slots = (n + 31) >> 5;
#pragma unroll
      for (int s = 0; s < FIELD_R; s++) {
        e[s] = 0, q[s] = 0, x0[s] = 0;
        if (s < n_slots) {
          const int i = it.y + lane + 32 * s;
          if (i < it.z)
            row(i, theta_old, e[s], q[s], x0[s]);
        }
      }
#pragma unroll
      for (int s = 0; s < FIELD_R; s++)
        if (s < n_slots && it.y + lane + 32 * s < it.z)
          field_stats<Real, IS_V>(e[s], q[s], x0[s], theta_old, alpha, sq, lin);
      sq = warp_sum(sq);
      lin = warp_sum(lin);
      const Real theta_new = column_draw<Real, IS_V>(sq, lin, theta_old, alpha, lam, mu, z);
#pragma unroll
      for (int s = 0; s < FIELD_R; s++)
        if (s < n_slots) {
          const int i = it.y + lane + 32 * s;
          if (i < it.z)
            __stcg(a.eq + i, field_update<Real, IS_V>(e[s], q[s], x0[s], theta_old, theta_new));
        }
      if (lane == 0) {
        a.theta[j] = theta_new;
        if (a.theta_t)
          a.theta_t[static_cast<int64_t>(j) * a.t_stride] = theta_new;
      }
    } else {
      for (int64_t i = it.y + lane; i < it.z; i += 32) {
        Real e, q, x0;
        row(i, theta_old, e, q, x0);
        field_stats<Real, IS_V>(e, q, x0, theta_old, alpha, sq, lin);
      }
      sq = warp_sum(sq);
      lin = warp_sum(lin);
      const Real theta_new = column_draw<Real, IS_V>(sq, lin, theta_old, alpha, lam, mu, z);
      for (int64_t i = it.y + lane; i < it.z; i += 32) {
        Real e, q, x0;
        row(i, theta_old, e, q, x0);
        __stcg(a.eq + i, field_update<Real, IS_V>(e, q, x0, theta_old, theta_new));
      }
      if (lane == 0) {
        a.theta[j] = theta_new;
        if (a.theta_t)
          a.theta_t[static_cast<int64_t>(j) * a.t_stride] = theta_new;
      }
    }
    k = __shfl_sync(FULL_MASK, k_next, 0);
  }
}

// ------------------------------------------------------------------------------------------------
// Last level: statistics by gather, draw, no scatter.
// ------------------------------------------------------------------------------------------------
template <typename Real> struct FieldStatsArgs {
  const int *idx; // CSC entry arrays of the main table (device row order)
  const Real *val;
  const int4 *item; // {column, lo, hi, first chunk}: nS chunks of long columns, nC per CTA, nW per warp
  const int *seg_count;
  int nS, nC, nW;
  const Pair<Real> *eq;
  Real *theta, *theta_t;
  int64_t t_stride;
  const Real *z;
  const int *group;
  const Real *alpha, *lambda, *mu;
  Real *partial; // [2 nS]
  int last_base;
  Real *pend_told, *pend_tnew;
};

template <typename Real, bool IS_V>
__device__ __forceinline__ void field_publish(const FieldStatsArgs<Real> &a, int j, Real sq, Real lin,
                                              Real theta_old, Real alpha) {
  const int g = a.group[j];
  const Real theta_new = column_draw<Real, IS_V>(sq, lin, theta_old, alpha, a.lambda[g], a.mu[g], a.z[j]);
  a.theta[j] = theta_new;
  if (a.theta_t)
    a.theta_t[static_cast<int64_t>(j) * a.t_stride] = theta_new;
  a.pend_told[j - a.last_base] = theta_old;
  a.pend_tnew[j - a.last_base] = theta_new;
}

template <typename Real, bool IS_V, bool UNIT>
__global__ void __launch_bounds__(STATS_THREADS) k_field_stats(FieldStatsArgs<Real> a) {
  __shared__ Real scratch[32];
  const int b = blockIdx.x;
  const bool cta_item = b < a.nS + a.nC;
  const int lane = threadIdx.x & 31;
  const int w = (b - a.nS - a.nC) * (STATS_THREADS / 32) + (threadIdx.x >> 5);
  if (!cta_item && w >= a.nW)
    return;
  const int4 it = __ldg(a.item + (cta_item ? b : a.nS + a.nC + w));
  const int j = it.x;
  const Real alpha = *a.alpha;
  const Real theta_old = a.theta[j];
  const int t = cta_item ? threadIdx.x : lane, nt = cta_item ? STATS_THREADS : 32;
  Real sq = 0, lin = 0;
  constexpr int U = 4; // gathers in flight per thread
  for (int p0 = it.y + t; p0 < it.z; p0 += U * nt) {
    int i[U];
    Real x[U];
    Pair<Real> v[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int p = p0 + u * nt;
      i[u] = p < it.z ? __ldcs(a.idx + p) : -1;
      x[u] = (UNIT || p >= it.z) ? Real(1) : __ldcs(a.val + p);
    }
#pragma unroll
    for (int u = 0; u < U; u++)
      if (i[u] >= 0)
        v[u] = __ldcg(a.eq + i[u]);
#pragma unroll
    for (int u = 0; u < U; u++)
      if (i[u] >= 0)
        field_stats<Real, IS_V>(v[u].x, v[u].y, x[u], theta_old, alpha, sq, lin);
  }
  if (cta_item) {
    sq = block_sum(sq, scratch);
    lin = block_sum(lin, scratch);
    if (threadIdx.x == 0) {
      if (b < a.nS) {
        a.partial[2 * b] = sq, a.partial[2 * b + 1] = lin;
      } else {
        field_publish<Real, IS_V>(a, j, sq, lin, theta_old, alpha);
      }
    }
  } else {
    sq = warp_sum(sq);
    lin = warp_sum(lin);
    if (lane == 0)
      field_publish<Real, IS_V>(a, j, sq, lin, theta_old, alpha);
  }
}

// Long columns: chunk statistics summed in chunk order, then the draw.  One thread per chunk item;
// only a column's first chunk acts.
template <typename Real, bool IS_V> __global__ void k_field_finish_long(FieldStatsArgs<Real> a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.nS)
    return;
  const int4 it = __ldg(a.item + b);
  if (it.w != b)
    return;
  Real sq = 0, lin = 0;
  for (int i = b; i < b + a.seg_count[b]; i++) {
    sq += a.partial[2 * i];
    lin += a.partial[2 * i + 1];
  }
  field_publish<Real, IS_V>(a, it.x, sq, lin, a.theta[it.x], *a.alpha);
}

} // namespace myfm
