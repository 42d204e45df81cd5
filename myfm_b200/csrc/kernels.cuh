// sm_100a kernels of the Gibbs sweep.  All of them are HBM/L2-bound gather/scatter work: there is
// no dense contraction, so no tensor cores.  Templated on Real (float =
// reference's bind_float.cpp instantiation, double = what the reference ships).
//
// Arithmetic is written in the reference's operation order and the file is compiled with
// -fmad=false, so every element-wise result is bit-identical to the CPU path; the only
// differences are the order of the per-column / per-vector sums (warp trees instead of a serial
// accumulator) and log/exp in the variates.
#pragma once

#include "common.cuh"

namespace myfm {

constexpr int MAX_REL = 8; // relation blocks per model

// The residual cache e and the factor cache q of the training rows live interleaved:
// eq[i] = {e_i, q_i} (see the column sweeps below).
template <typename Real> struct PairOf;
template <> struct PairOf<float> { using type = float2; };
template <> struct PairOf<double> { using type = double2; };
template <typename Real> using Pair = typename PairOf<Real>::type; // .x = e, .y = q

template <typename Real> struct CsView { // CSR or CSC, device pointers
  const int *ptr = nullptr;
  const int *idx = nullptr;
  const Real *val = nullptr;
};

// Per-block tables consumed by the forward pass: for block row s and factor r
// q[s*K + r] = (X_B V_B[:,r])[s], qs[s*K + r] = (X_B^2 V_B[:,r]^2)[s], lin[s] = (X_B w_B)[s].
template <typename Real> struct RelPredictView {
  const int *map = nullptr; // [n_rows] -> block row
  const Real *lin = nullptr;
  const Real *q = nullptr;
  const Real *qs = nullptr;
};
template <typename Real> struct RelPredictPack {
  int n = 0;
  RelPredictView<Real> r[MAX_REL];
};

// ----------------------------------------------------------------------------------------------
// Row-parallel SpMV: out[i] = sum_p val[p] * x[idx[p]]      (q_init; FMTrainer.hpp:320,331)
// SQUARED: out[i] = sum_p val[p]^2 * x[idx[p]]^2           (q_S;    FMTrainer.hpp:388-393)
// LPR lanes cooperate on one row.
// ----------------------------------------------------------------------------------------------
// UNIT: every stored value is 1 (the val stream is not read).  row_len > 0: every row has exactly
// row_len entries (the ptr stream is not read) — both hold for one-hot encoded tables.
template <typename Real, int LPR, bool SQUARED, bool UNIT = false>
__global__ void __launch_bounds__(256) k_spmv(int n_rows, CsView<Real> A, const Real *__restrict__ x,
                                               Real *__restrict__ out, int out_stride, int row_len = 0) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = tid / LPR, sub = tid % LPR;
  Real acc = 0;
  if (row < n_rows) {
    const int b = row_len > 0 ? row * row_len : A.ptr[row];
    const int en = row_len > 0 ? b + row_len : A.ptr[row + 1];
    for (int p = b + sub; p < en; p += LPR) {
      Real v = UNIT ? Real(1) : A.val[p], xv = x[__ldcs(A.idx + p)];
      acc += SQUARED ? (v * v) * (xv * xv) : v * xv;
    }
  }
  acc = subwarp_sum<Real, LPR>(acc);
  if (row < n_rows && sub == 0)
    out[static_cast<size_t>(row) * out_stride] = acc;
}

// Block tables for the forward pass, one launch per block: warp per block row, lanes over factors.
// Vt is feature-major [dim_all x K]; `offset` is the block's first feature.
template <typename Real>
__global__ void __launch_bounds__(256)
    k_block_tables(int n_block_rows, CsView<Real> B, const Real *__restrict__ w,
                   const Real *__restrict__ Vt, int K, int offset, Real *__restrict__ lin,
                   Real *__restrict__ q, Real *__restrict__ qs) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= n_block_rows)
    return;
  const int b = B.ptr[s], en = B.ptr[s + 1];
  Real l = 0;
  for (int p = b + lane; p < en; p += 32)
    l += B.val[p] * w[offset + B.idx[p]];
  l = warp_sum(l);
  if (lane == 0)
    lin[s] = l;
  for (int r = lane; r < K; r += 32) {
    Real a = 0, a2 = 0;
    for (int p = b; p < en; p++) {
      Real x = B.val[p], v = Vt[static_cast<size_t>(offset + B.idx[p]) * K + r];
      a += x * v;
      a2 += (x * x) * (v * v);
    }
    q[static_cast<size_t>(s) * K + r] = a;
    qs[static_cast<size_t>(s) * K + r] = a2;
  }
}

// ----------------------------------------------------------------------------------------------
// Forward pass, all K factors fused in ONE pass over the CSR (FM.hpp:54-136 does 2K SpMVs):
//   out[i] = w0 + X w + sum_b lin_b[map_b(i)] + 1/2 sum_r [ q_r^2 - s_r ]   ( - y[i] if y )
// LPR lanes per row; lane `sub` owns factors sub, sub+LPR, ...
// ----------------------------------------------------------------------------------------------
template <typename Real, int LPR>
__global__ void __launch_bounds__(256)
    k_predict(int n_rows, CsView<Real> X, const Real *__restrict__ w, const Real *__restrict__ Vt,
              int K, const Real *__restrict__ w0_ptr, RelPredictPack<Real> rels,
              const Real *__restrict__ y, Real *__restrict__ out, int out_stride) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = tid / LPR, sub = tid % LPR;
  Real lin = 0, acc = 0;
  if (row < n_rows) {
    const int b = X.ptr[row], en = X.ptr[row + 1];
    for (int p = b + sub; p < en; p += LPR)
      lin += X.val[p] * w[X.idx[p]];
    int srow[MAX_REL];
#pragma unroll
    for (int k = 0; k < MAX_REL; k++)
      if (k < rels.n) {
        srow[k] = rels.r[k].map[row];
        if (sub == 0)
          lin += rels.r[k].lin[srow[k]];
      }
    const Real half = static_cast<Real>(0.5);
    for (int r = sub; r < K; r += LPR) {
      Real qr = 0, sr = 0;
      for (int p = b; p < en; p++) {
        Real x = X.val[p], v = Vt[static_cast<size_t>(X.idx[p]) * K + r];
        qr += x * v;
        sr += (x * x) * (v * v);
      }
#pragma unroll
      for (int k = 0; k < MAX_REL; k++)
        if (k < rels.n) {
          qr += rels.r[k].q[static_cast<size_t>(srow[k]) * K + r];
          sr += rels.r[k].qs[static_cast<size_t>(srow[k]) * K + r];
        }
      acc += (qr * qr) * half;
      acc -= sr * half;
    }
  }
  lin = subwarp_sum<Real, LPR>(lin);
  acc = subwarp_sum<Real, LPR>(acc);
  if (row < n_rows && sub == 0) {
    Real t = (*w0_ptr + lin) + acc;
    out[static_cast<size_t>(row) * out_stride] = y ? t - y[row] : t;
  }
}

// Forward pass for short rows: one warp walks PREDICT_ROWS_PER_WARP consecutive rows.  The lanes
// first load the row's entries (coalesced), then every lane owns factors lane, lane+32, .. and
// reads each V row as one coalesced line (re-used from L1 when consecutive rows share a feature,
// as they do in the trainer's row order).  One shuffle reduction per row.
constexpr int PREDICT_ROWS_PER_WARP = 16;
template <typename Real, bool HAS_REL>
__global__ void __launch_bounds__(256)
    k_predict_warp(int n_rows, CsView<Real> X, const Real *__restrict__ w,
                   const Real *__restrict__ Vt, int K, const Real *__restrict__ w0_ptr,
                   RelPredictPack<Real> rels, const Real *__restrict__ y, Real *__restrict__ out,
                   int out_stride) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const Real w0 = *w0_ptr, half = static_cast<Real>(0.5);
  const int row0 = warp * PREDICT_ROWS_PER_WARP;
  if (row0 >= n_rows)
    return;
  const int n_mine = min(PREDICT_ROWS_PER_WARP, n_rows - row0);
  // row pointers of the tile: lane i holds ptr[row0 + i]
  const int my_ptr = lane <= n_mine ? X.ptr[row0 + lane] : 0;
  for (int i = 0; i < n_mine; i++) {
    const int row = row0 + i;
    const int b = __shfl_sync(FULL_MASK, my_ptr, i), en = __shfl_sync(FULL_MASK, my_ptr, i + 1);
    Real total = 0; // per-lane share of lin + 1/2 sum_r (q_r^2 - s_r)
    Real qr[2] = {0, 0}, sr[2] = {0, 0}; // factors lane and lane + 32; further ones below
    for (int c = b; c < en; c += 32) {
      const int p = c + lane;
      int j = 0;
      Real x = 0;
      if (p < en) {
        j = __ldcs(X.idx + p), x = __ldcs(X.val + p);
        total += x * w[j];
      }
      const int m = min(32, en - c);
      for (int k = 0; k < m; k++) {
        const int jk = __shfl_sync(FULL_MASK, j, k);
        const Real xk = __shfl_sync(FULL_MASK, x, k);
        const Real *vrow = Vt + static_cast<size_t>(jk) * K;
        if (lane < K) {
          Real v = vrow[lane];
          qr[0] += xk * v, sr[0] += (xk * xk) * (v * v);
        }
        if (lane + 32 < K) {
          Real v = vrow[lane + 32];
          qr[1] += xk * v, sr[1] += (xk * xk) * (v * v);
        }
      }
    }
    int srow[MAX_REL];
    if (HAS_REL) {
#pragma unroll
      for (int k = 0; k < MAX_REL; k++)
        if (k < rels.n) {
          srow[k] = rels.r[k].map[row];
          if (lane == 0)
            total += rels.r[k].lin[srow[k]];
        }
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int r = lane + 32 * u;
      if (r < K) {
        if (HAS_REL) {
#pragma unroll
          for (int k = 0; k < MAX_REL; k++)
            if (k < rels.n) {
              qr[u] += rels.r[k].q[static_cast<size_t>(srow[k]) * K + r];
              sr[u] += rels.r[k].qs[static_cast<size_t>(srow[k]) * K + r];
            }
        }
        total += (qr[u] * qr[u]) * half;
        total -= sr[u] * half;
      }
    }
    total = warp_sum(total);
    if (lane == 0) {
      Real t = w0 + total;
      out[static_cast<size_t>(row) * out_stride] = y ? t - y[row] : t;
    }
  }
}

// Forward pass for tables whose rows all hold exactly L entries (one-hot / categorical fields)
// and 16 <= K <= 64.  One warp per tile of 32 consecutive rows: the lanes own factors (KPL per
// lane), walk the tile's rows with the entries broadcast by shuffle and read every V row as one
// coalesced line; the 32 per-row sums are then reduced across the lanes by a 31-shuffle transposed
// butterfly (one shuffle per row instead of five) and lane l finishes row l.
// PAIR: out is the trainer's interleaved {e, q} array; both halves are written (q = 0) so that
// whole sectors are stored.
// FULL: K == 32 * KPL exactly (K = 32, 64): no per-lane predicate, the row stride is a compile-time constant and the
// address of a V row is one IMAD.WIDE.  The per-row term is accumulated as q^2 - s and halved once at the end
// (scaling by a power of two commutes with every rounding of the sum): 1177 -> ~550 instructions per tile of 32
// rows for K = 32, L = 2 — the kernel is issue-bound (profiles/r02g_predict_tile.md).
template <typename Real, int L, bool UNIT, int KPL, bool PAIR, bool FULL>
__global__ void __launch_bounds__(256)
    k_predict_tile(int n_rows, const int *__restrict__ idx, const Real *__restrict__ val,
                   const Real *__restrict__ w, const Real *__restrict__ Vt, int K,
                   const Real *__restrict__ w0_ptr, const Real *__restrict__ y, Real *__restrict__ out,
                   int out_stride) {
  const int lane = threadIdx.x & 31;
  const int n_tiles = (n_rows + 31) >> 5;
  const Real w0 = *w0_ptr, half = static_cast<Real>(0.5);
  const char *vbase = reinterpret_cast<const char *>(Vt + lane);
  const int row_bytes = (FULL ? 32 * KPL : K) * static_cast<int>(sizeof(Real));
  for (int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < n_tiles; tile += (gridDim.x * blockDim.x) >> 5) {
    const int row = tile * 32 + lane;
    const bool valid = row < n_rows;
    const int64_t p0 = static_cast<int64_t>(valid ? row : n_rows - 1) * L;
    int j[L];
    Real x[L];
    Real lin = 0;
#pragma unroll
    for (int k = 0; k < L; k++) {
      j[k] = __ldcs(idx + p0 + k);
      x[k] = UNIT ? Real(1) : __ldcs(val + p0 + k);
      lin += x[k] * w[j[k]];
    }
    Real part[32]; // twice the row's pair term: sum over this lane's factors of q^2 - s
#pragma unroll
    for (int rr = 0; rr < 32; rr++) {
      Real q[KPL], s[KPL];
#pragma unroll
      for (int u = 0; u < KPL; u++)
        q[u] = 0, s[u] = 0;
#pragma unroll
      for (int k = 0; k < L; k++) {
        const int jk = __shfl_sync(FULL_MASK, j[k], rr);
        const Real xk = UNIT ? Real(1) : __shfl_sync(FULL_MASK, x[k], rr);
        const Real *vrow = reinterpret_cast<const Real *>(vbase + static_cast<int64_t>(jk) * row_bytes);
#pragma unroll
        for (int u = 0; u < KPL; u++)
          if (FULL || lane + 32 * u < K) {
            const Real v = vrow[32 * u];
            const Real xv = UNIT ? v : xk * v;
            const Real xv2 = UNIT ? v * v : (xk * xk) * (v * v);
            if (k == 0) // 0 + a = a: the first field assigns
              q[u] = xv, s[u] = xv2;
            else
              q[u] += xv, s[u] += xv2;
          }
      }
      Real tot = 0;
#pragma unroll
      for (int u = 0; u < KPL; u++)
        if (FULL || lane + 32 * u < K) {
          if (u == 0 && FULL)
            tot = q[u] * q[u] - s[u];
          else
            tot += q[u] * q[u], tot -= s[u];
        }
      part[rr] = tot;
    }
    // transposed butterfly: afterwards part[0] of lane l is the sum over the lanes of part[l]
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const bool upper = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < o; i++) {
        const Real send = upper ? part[i] : part[i + o];
        const Real keep = upper ? part[i + o] : part[i];
        part[i] = keep + __shfl_xor_sync(FULL_MASK, send, o);
      }
    }
    if (valid) {
      Real t = (w0 + lin) + part[0] * half;
      if (y)
        t = t - y[row];
      if (PAIR) {
        Pair<Real> v;
        v.x = t, v.y = 0;
        __stcg(reinterpret_cast<Pair<Real> *>(out) + row, v);
      } else {
        out[static_cast<size_t>(row) * out_stride] = t;
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Scalars of one sweep live in a small device array so that no step needs the host.
// ----------------------------------------------------------------------------------------------
template <typename Real> struct HyperView {
  Real *alpha;    // [1]
  Real *w0;       // [1]
  Real *mu_w;     // [G]
  Real *lambda_w; // [G]
  Real *mu_V;     // [G x K] column-major (g + G*r), as in HyperParams.hpp
  Real *lambda_V; // [G x K]
};

// Stage 1 of a deterministic grid reduction of f(e_i): block partials.
// MODE 0: e^2 (update_alpha, FMTrainer.hpp:138)   MODE 1: (w0 - e) (update_w0, :223)
template <typename Real, int MODE>
__global__ void __launch_bounds__(512)
    k_reduce_e(int64_t n, const Pair<Real> *__restrict__ eq, const Real *__restrict__ w0_ptr,
               Real *__restrict__ partial) {
  __shared__ Real scratch[32];
  Real acc = 0;
  const Real w0 = MODE == 1 ? *w0_ptr : Real(0);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    Real v = eq[i].x;
    acc += MODE == 0 ? v * v : (w0 - v);
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0)
    partial[blockIdx.x] = acc;
}

// Both sums in one pass over e (regression with fit_w0: update_alpha is followed by update_w0 and neither
// changes e in between): partial[b] = block sum of e^2, partial[gridDim.x + b] = block sum of (w0 - e);
// per thread and per block the same additions in the same order as the two separate kernels.
template <typename Real>
__global__ void __launch_bounds__(512)
    k_reduce_e_both(int64_t n, const Pair<Real> *__restrict__ eq, const Real *__restrict__ w0_ptr,
                    Real *__restrict__ partial) {
  __shared__ Real scratch[32];
  Real acc2 = 0, acc1 = 0;
  const Real w0 = *w0_ptr;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    Real v = eq[i].x;
    acc2 += v * v;
    acc1 += (w0 - v);
  }
  acc2 = block_sum(acc2, scratch);
  acc1 = block_sum(acc1, scratch);
  if (threadIdx.x == 0)
    partial[blockIdx.x] = acc2, partial[gridDim.x + blockIdx.x] = acc1;
}

// Stage 2 of both (one block): k_finish_alpha, then k_finish_w0 with the new alpha.  part_w0: the
// partials of (w0 - e), n_partial of them.
template <typename Real>
__global__ void k_finish_alpha_w0(int n_partial, const Real *__restrict__ partial, const Real *__restrict__ part_w0,
                                  Real beta_0, const Real *__restrict__ g_std, Real *alpha_ptr, int n_train,
                                  Real reg_0, const Real *__restrict__ z, Real *w0, Real *delta) {
  __shared__ Real scratch[32];
  Real acc2 = 0, acc1 = 0;
  for (int i = threadIdx.x; i < n_partial; i += blockDim.x)
    acc2 += partial[i], acc1 += part_w0[i];
  acc2 = block_sum(acc2, scratch);
  acc1 = block_sum(acc1, scratch);
  if (threadIdx.x == 0) {
    Real variance = (beta_0 + acc2) / 2;
    const Real alpha = *g_std * (1 / variance);
    *alpha_ptr = alpha;
    Real lin = alpha * acc1;
    Real quad = alpha * n_train + reg_0;
    Real w0_new = (lin / quad) + *z / sqrt(quad);
    *delta = (w0_new - *w0);
    *w0 = w0_new;
  }
}

// Stage 2 (one block): alpha ~ Gamma((alpha_0+N)/2, 2/(beta_0+sum e^2)) from a standardised
// Gamma variate (FMTrainer.hpp:140-144).
template <typename Real>
__global__ void k_finish_alpha(int n_partial, const Real *__restrict__ partial, Real beta_0,
                               const Real *__restrict__ g_std, Real *alpha) {
  __shared__ Real scratch[32];
  Real acc = 0;
  for (int i = threadIdx.x; i < n_partial; i += blockDim.x)
    acc += partial[i];
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) {
    Real variance = (beta_0 + acc) / 2;
    *alpha = *g_std * (1 / variance);
  }
}

// Stage 2 of update_w0 (FMTrainer.hpp:223-228): draws w0 and leaves delta = w0_new - w0_old.
template <typename Real>
__global__ void k_finish_w0(int n_partial, const Real *__restrict__ partial, int n_train, Real reg_0,
                            const Real *__restrict__ alpha_ptr, const Real *__restrict__ z,
                            Real *w0, Real *delta) {
  __shared__ Real scratch[32];
  Real acc = 0;
  for (int i = threadIdx.x; i < n_partial; i += blockDim.x)
    acc += partial[i];
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) {
    Real alpha = *alpha_ptr;
    Real lin = alpha * acc;
    Real quad = alpha * n_train + reg_0;
    Real w0_new = (lin / quad) + *z / sqrt(quad);
    *delta = (w0_new - *w0);
    *w0 = w0_new;
  }
}

template <typename Real>
__global__ void __launch_bounds__(256)
    k_add_scalar(int64_t n, Pair<Real> *__restrict__ eq, const Real *__restrict__ delta) {
  const Real d = *delta;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    eq[i].x += d;
}

// ----------------------------------------------------------------------------------------------
// Group hyper-parameters (FMTrainer.hpp:150-216).  One block per (group, vector); the vector is
// w (n_vec == 1) or one factor column of V.  lambda first (with the old mu), then mu.
//   lambda_g = Gamma((alpha_0+n_g)/2, 1) * 2/(beta_0 + sum (theta-mu_g)^2)
//   mu_g     = lin/quad + z/sqrt(quad),  quad = lambda_g (gamma_0+n_g),
//              lin = lambda_g (gamma_0 mu_0 + sum theta)
// feat_ptr/feat_idx: features of each group, ascending.
// ----------------------------------------------------------------------------------------------
// The features of a group are cut into chunks of HYPER_CHUNK (host: chunk_group / chunk_begin, one entry per
// chunk, groups in order; an empty group still owns one chunk); a block sums one chunk for one vector, the
// block of a (group, vector) that finishes last adds the chunk sums in chunk order (deterministic) and draws.
// One block per (group, vector) walked 70 k features of the C4 workload in ~60 us of dependent latencies.
constexpr int HYPER_THREADS = 256;
constexpr int HYPER_UNROLL = 16;
constexpr int HYPER_CHUNK = HYPER_THREADS * HYPER_UNROLL;
template <typename Real>
__global__ void __launch_bounds__(HYPER_THREADS)
    k_group_hyper(int G, int n_chunks, const int *__restrict__ chunk_group, const int *__restrict__ chunk_begin,
                  const int *__restrict__ feat_ptr, const int *__restrict__ feat_idx,
                  const Real *__restrict__ theta, int64_t theta_stride, Real *mu, Real *lambda,
                  const Real *__restrict__ g_std, const Real *__restrict__ z, Real beta_0,
                  Real gamma_0, Real mu_0, Real *__restrict__ chunk_sums, unsigned int *__restrict__ done) {
  __shared__ Real scratch[32];
  const int c = blockIdx.x % n_chunks, v = blockIdx.x / n_chunks;
  const int g = chunk_group[c];
  const Real *th = theta + theta_stride * v;
  const int b = chunk_begin[c], en = min(feat_ptr[g + 1], b + HYPER_CHUNK);
  const Real mean = mu[g + G * v];
  Real dev2 = 0, sum = 0;
  Real t[HYPER_UNROLL];
#pragma unroll
  for (int u = 0; u < HYPER_UNROLL; u++) { // all gathers in flight before the first add
    const int p = b + threadIdx.x + u * HYPER_THREADS;
    t[u] = p < en ? th[feat_idx[p]] : mean;
  }
#pragma unroll
  for (int u = 0; u < HYPER_UNROLL; u++)
    if (b + threadIdx.x + u * HYPER_THREADS < en) {
      Real dev = t[u] - mean;
      dev2 += dev * dev;
      sum += t[u];
    }
  dev2 = block_sum(dev2, scratch);
  sum = block_sum(sum, scratch);
  if (threadIdx.x == 0) {
    const int first = feat_ptr[g], n_g = feat_ptr[g + 1] - first;
    const int chunks_g = max(1, (n_g + HYPER_CHUNK - 1) / HYPER_CHUNK);
    if (chunks_g > 1) {
      Real *mine = chunk_sums + 2 * (static_cast<size_t>(v) * n_chunks + c);
      __stcg(mine, dev2), __stcg(mine + 1, sum);
      __threadfence();
      if (atomicAdd(done + g + G * v, 1u) != static_cast<unsigned int>(chunks_g - 1))
        return;
      __threadfence();
      done[g + G * v] = 0; // ready for the next launch
      const int c0 = c - (b - first) / HYPER_CHUNK; // the group's first chunk
      dev2 = 0, sum = 0;
      for (int k = 0; k < chunks_g; k++) {
        const Real *part = chunk_sums + 2 * (static_cast<size_t>(v) * n_chunks + c0 + k);
        dev2 += __ldcg(part), sum += __ldcg(part + 1);
      }
    }
    Real beta = beta_0 + dev2;
    Real lam = g_std[g + G * v] * (2 / beta);
    lambda[g + G * v] = lam;
    Real square = lam * (gamma_0 + n_g);
    Real linear = gamma_0 * mu_0 + sum;
    linear *= lam;
    mu[g + G * v] = (linear / square) + z[g + G * v] / sqrt(square);
  }
}

// ----------------------------------------------------------------------------------------------
// Column sweeps over one dependency level of the main table.  The columns of a level are pairwise
// row-disjoint, so the concurrent read-modify-write of e / q below is the reference's serial
// loop, bit for bit, in any interleaving.
//
// V (FMTrainer.hpp:343-376):  h = x (q - x v_old);  sq = sum h^2;  lin = sum -e h + sq v_old
//   v_new = draw(alpha sq + lambda, alpha lin + lambda mu);  q += x d;  e += h d
// w (FMTrainer.hpp:237-254):  e' = e - x w_old;  sq = lambda + alpha sum x^2;
//   lin = sum (-alpha x) e' + lambda mu;  e = e' + x w_new
//
// e and q live interleaved, eq[i] = {e_i, q_i}: a column entry touches ONE 32-byte sector per
// row instead of two.  Index / value streams are read once with evict-first loads; eq goes
// through L2 only (no reuse inside a launch), where the f32 working set of the C4 workload
// (80 MB) stays resident between levels.
// Work items (host_data.hpp: SweepPlan): S = chunk of a long column (two-phase), C = whole
// column per CTA, W = whole column per warp; C and W keep the gathered entries in registers
// between the reduction and the update, so each entry is gathered once and scattered once.
// UNIT: all values of the level are 1 (val stream not read).  CONTIG: every column is a
// contiguous row range (idx stream not read; the primary level after the row reordering).
// ----------------------------------------------------------------------------------------------

constexpr int SWEEP_THREADS = 256;
constexpr int SWEEP_R = 8;                            // entries per thread held in registers
constexpr int SWEEP_WARP_MAX = 32 * SWEEP_R;          // longest column of a W item
constexpr int SWEEP_CHUNK = SWEEP_THREADS * SWEEP_R;  // longest column of a C item / chunk size
constexpr int SWEEP_WARPS = SWEEP_THREADS / 32;

template <typename Real> struct SweepArgs {
  const int *idx;       // CSC entry arrays of the main table (device row order)
  const Real *val;
  const int4 *item;     // work items of THIS level: {column, lo, hi, first chunk} (host_data.hpp: SweepItem)
  const int *seg_count; // S items: chunks of the column
  int nS, nC, nW;
  Pair<Real> *eq;
  Real *theta;          // w, or column r of V (column-major)
  Real *theta_t;        // feature-major mirror: theta_t[j * t_stride], or nullptr
  int64_t t_stride;
  const Real *z;        // standardised normals indexed by feature
  const int *group;     // group of every feature
  const Real *alpha;
  const Real *lambda;   // [G] of this vector
  const Real *mu;       // [G]
  Real *partial;        // [2 * nS] chunk statistics
  Real *theta_old_buf;  // [nS] value before the update, written by a column's first chunk
  // row-sharded training (k_level_dist): statistics of every column of the level, summed over ranks
  const int *item_slot; // column slot of every item (same numbering on every rank)
  Real *colstat;        // [2 * columns of the level]
  Real *told;           // [columns of the level] value before the update
};

template <typename Real, bool IS_V>
__device__ __forceinline__ Real column_draw(Real sq, Real lin, Real theta_old, Real alpha, Real lam,
                                            Real mu, Real z) {
  if (IS_V) {
    lin += sq * theta_old;
    sq *= alpha;
    lin *= alpha;
    sq += lam;
    lin += lam * mu;
  } else {
    sq = lam + alpha * sq;
    lin = lin + lam * mu;
  }
  return (lin / sq) + z / sqrt(sq);
}

// The entries one thread owns: t, t + NT, t + 2 NT, ... of [lo, hi).  Only the first
// ceil((hi - lo) / NT) slots are touched (uniform across the warp / CTA).
template <typename Real, bool IS_V, bool UNIT, bool CONTIG, int NT> struct ColumnEntries {
  int i[SWEEP_R];
  Real x[SWEEP_R], e[SWEEP_R], q[SWEEP_R];
  int n_slots;

  __device__ __forceinline__ void load(const SweepArgs<Real> &a, int lo, int hi, int t) {
    n_slots = (hi - lo + NT - 1) / NT;
    int first = 0;
    if (CONTIG)
      first = lo < hi ? a.idx[lo] : 0;
#pragma unroll
    for (int s = 0; s < SWEEP_R; s++) {
      i[s] = -1, x[s] = Real(UNIT ? 1 : 0);
      if (s < n_slots) {
        const int p = lo + t + s * NT;
        if (p < hi) {
          i[s] = CONTIG ? first + (p - lo) : __ldcs(a.idx + p);
          if (!UNIT)
            x[s] = __ldcs(a.val + p);
        }
      }
    }
#pragma unroll
    for (int s = 0; s < SWEEP_R; s++) {
      e[s] = 0, q[s] = 0;
      if (s < n_slots && i[s] >= 0) {
        const Pair<Real> v = __ldcg(a.eq + i[s]);
        e[s] = v.x, q[s] = v.y;
      }
    }
  }

  __device__ __forceinline__ void stats(Real theta_old, Real alpha, Real &sq, Real &lin) const {
#pragma unroll
    for (int s = 0; s < SWEEP_R; s++)
      if (s < n_slots && i[s] >= 0) {
        if (IS_V) {
          Real h = x[s] * (q[s] - x[s] * theta_old);
          sq += h * h;
          lin += (-e[s]) * h;
        } else {
          Real e1 = e[s] - x[s] * theta_old;
          sq += x[s] * x[s];
          lin += ((-alpha) * x[s]) * e1;
        }
      }
  }

  __device__ __forceinline__ void update(const SweepArgs<Real> &a, Real theta_old,
                                         Real theta_new) const {
#pragma unroll
    for (int s = 0; s < SWEEP_R; s++)
      if (s < n_slots && i[s] >= 0) {
        Pair<Real> v;
        if (IS_V) {
          Real h = x[s] * (q[s] - x[s] * theta_old);
          v.y = q[s] + x[s] * (theta_new - theta_old);
          v.x = e[s] + h * (theta_new - theta_old);
        } else {
          Real e1 = e[s] - x[s] * theta_old;
          v.x = e1 + x[s] * theta_new;
          v.y = q[s];
        }
        __stcg(a.eq + i[s], v);
      }
  }
};

template <typename Real>
__device__ __forceinline__ void store_theta(const SweepArgs<Real> &a, int j, Real theta_new) {
  a.theta[j] = theta_new;
  if (a.theta_t)
    a.theta_t[static_cast<int64_t>(j) * a.t_stride] = theta_new;
}

// One launch per level: blocks [0, nS) chunk statistics, [nS, nS+nC) one column per CTA, the rest
// eight columns per CTA (one per warp).  Everything the draw needs is loaded up front so that the
// dependent chain of a column is item -> {entries, theta, hypers} -> reduce -> draw -> scatter.
template <typename Real, bool IS_V, bool UNIT, bool CONTIG>
__global__ void __launch_bounds__(SWEEP_THREADS, sizeof(Real) == 4 ? 5 : 3) k_level_sweep(SweepArgs<Real> a) {
  __shared__ Real scratch[32];
  const int b = blockIdx.x;
  const bool cta_item = b < a.nS + a.nC;
  const int lane = threadIdx.x & 31;
  const int w = (b - a.nS - a.nC) * SWEEP_WARPS + (threadIdx.x >> 5);
  if (!cta_item && w >= a.nW)
    return;
  const int4 it = __ldg(a.item + (cta_item ? b : a.nS + a.nC + w));
  const int j = it.x;
  const Real alpha = *a.alpha;
  const Real theta_old = a.theta[j];
  const int g = a.group[j];
  const Real lam = a.lambda[g], mu = a.mu[g], z = a.z[j];
  Real sq = 0, lin = 0;
  if (cta_item) {
    ColumnEntries<Real, IS_V, UNIT, CONTIG, SWEEP_THREADS> en;
    en.load(a, it.y, it.z, threadIdx.x);
    en.stats(theta_old, alpha, sq, lin);
    sq = block_sum(sq, scratch);
    lin = block_sum(lin, scratch);
    if (b < a.nS) { // long column: the update runs in k_level_seg_update
      if (threadIdx.x == 0) {
        a.partial[2 * b] = sq;
        a.partial[2 * b + 1] = lin;
        if (it.w == b)
          a.theta_old_buf[b] = theta_old;
      }
      return;
    }
    const Real theta_new = column_draw<Real, IS_V>(sq, lin, theta_old, alpha, lam, mu, z);
    if (threadIdx.x == 0)
      store_theta(a, j, theta_new);
    en.update(a, theta_old, theta_new);
  } else {
    ColumnEntries<Real, IS_V, UNIT, CONTIG, 32> en;
    en.load(a, it.y, it.z, lane);
    en.stats(theta_old, alpha, sq, lin);
    sq = warp_sum(sq);
    lin = warp_sum(lin);
    const Real theta_new = column_draw<Real, IS_V>(sq, lin, theta_old, alpha, lam, mu, z);
    if (lane == 0)
      store_theta(a, j, theta_new);
    en.update(a, theta_old, theta_new);
  }
}

// Second phase of the long columns: every chunk sums its column's chunk statistics in chunk
// order (deterministic, identical in every CTA), draws, and updates its own entries.
template <typename Real, bool IS_V, bool UNIT, bool CONTIG>
__global__ void __launch_bounds__(SWEEP_THREADS) k_level_seg_update(SweepArgs<Real> a) {
  __shared__ Real bcast[2];
  const int b = blockIdx.x;
  const int4 it = __ldg(a.item + b);
  const int j = it.x;
  const int first = it.w, last = first + a.seg_count[b];
  const int g = a.group[j];
  const Real lam = a.lambda[g], mu = a.mu[g], z = a.z[j], alpha = *a.alpha;
  const Real theta_old = a.theta_old_buf[first];
  ColumnEntries<Real, IS_V, UNIT, CONTIG, SWEEP_THREADS> en;
  en.load(a, it.y, it.z, threadIdx.x);
  if (threadIdx.x < 32) {
    Real sq = 0, lin = 0;
    for (int i = first + threadIdx.x; i < last; i += 32) {
      sq += a.partial[2 * i];
      lin += a.partial[2 * i + 1];
    }
    sq = warp_sum(sq);
    lin = warp_sum(lin);
    if (threadIdx.x == 0)
      bcast[0] = sq, bcast[1] = lin;
  }
  __syncthreads();
  const Real theta_new = column_draw<Real, IS_V>(bcast[0], bcast[1], theta_old, alpha, lam, mu, z);
  if (b == first && threadIdx.x == 0)
    store_theta(a, j, theta_new);
  en.update(a, theta_old, theta_new);
}

// Row-sharded training: the rows of a column live on several GPUs, so every column takes the
// two-phase route — local statistics (UPDATE = false), all-reduce of colstat over the ranks,
// identical draw on every rank and local rank-1 update (UPDATE = true).
template <typename Real, bool IS_V, bool UNIT, bool CONTIG, bool UPDATE>
__global__ void __launch_bounds__(SWEEP_THREADS) k_level_dist(SweepArgs<Real> a) {
  __shared__ Real scratch[32];
  const int b = blockIdx.x;
  const bool cta_item = b < a.nS + a.nC;
  const int lane = threadIdx.x & 31;
  const int w = (b - a.nS - a.nC) * SWEEP_WARPS + (threadIdx.x >> 5);
  if (!cta_item && w >= a.nW)
    return;
  const int item = cta_item ? b : a.nS + a.nC + w;
  const int4 it = __ldg(a.item + item);
  const int j = it.x, slot = a.item_slot[item];
  const Real alpha = *a.alpha;
  const bool leader = cta_item ? (threadIdx.x == 0 && (b >= a.nS || it.w == b)) : lane == 0;
  if (!UPDATE) {
    const Real theta_old = a.theta[j];
    Real sq = 0, lin = 0;
    if (cta_item) {
      ColumnEntries<Real, IS_V, UNIT, CONTIG, SWEEP_THREADS> en;
      en.load(a, it.y, it.z, threadIdx.x);
      en.stats(theta_old, alpha, sq, lin);
      sq = block_sum(sq, scratch);
      lin = block_sum(lin, scratch);
    } else {
      ColumnEntries<Real, IS_V, UNIT, CONTIG, 32> en;
      en.load(a, it.y, it.z, lane);
      en.stats(theta_old, alpha, sq, lin);
      sq = warp_sum(sq);
      lin = warp_sum(lin);
    }
    if (b < a.nS && threadIdx.x == 0) // chunk of a long column: folded by k_level_chunk_fold
      a.partial[2 * b] = sq, a.partial[2 * b + 1] = lin;
    if (leader) {
      a.told[slot] = theta_old;
      if (b >= a.nS || !cta_item)
        a.colstat[2 * slot] = sq, a.colstat[2 * slot + 1] = lin;
    }
  } else {
    const int g = a.group[j];
    const Real theta_old = a.told[slot];
    const Real theta_new = column_draw<Real, IS_V>(a.colstat[2 * slot], a.colstat[2 * slot + 1], theta_old,
                                                   alpha, a.lambda[g], a.mu[g], a.z[j]);
    if (cta_item) {
      ColumnEntries<Real, IS_V, UNIT, CONTIG, SWEEP_THREADS> en;
      en.load(a, it.y, it.z, threadIdx.x);
      en.update(a, theta_old, theta_new);
    } else {
      ColumnEntries<Real, IS_V, UNIT, CONTIG, 32> en;
      en.load(a, it.y, it.z, lane);
      en.update(a, theta_old, theta_new);
    }
    if (leader)
      store_theta(a, j, theta_new);
  }
}

// colstat[slot] of a long column = sum of its chunk statistics, in chunk order.
template <typename Real> __global__ void k_level_chunk_fold(SweepArgs<Real> a) {
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
  const int slot = a.item_slot[b];
  a.colstat[2 * slot] = sq, a.colstat[2 * slot + 1] = lin;
}

// partial[0] = sum of n partial sums (one block), so that one scalar crosses the ranks
template <typename Real> __global__ void k_fold_partials(int n, Real *partial) {
  __shared__ Real scratch[32];
  Real acc = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    acc += partial[i];
  acc = block_sum(acc, scratch);
  __syncthreads();
  if (threadIdx.x == 0)
    partial[0] = acc;
}
// Two runs of n partials -> partial[0], partial[1] (one all-reduce of two scalars follows).
template <typename Real> __global__ void k_fold_partials2(int n, Real *partial) {
  __shared__ Real scratch[32];
  Real acc0 = 0, acc1 = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    acc0 += partial[i], acc1 += partial[n + i];
  acc0 = block_sum(acc0, scratch);
  acc1 = block_sum(acc1, scratch);
  __syncthreads();
  if (threadIdx.x == 0)
    partial[0] = acc0, partial[1] = acc1;
}

// ----------------------------------------------------------------------------------------------
// Relation blocks (FMTrainer.hpp:256-313 for w, :378-482 for V).  Rows are pre-grouped by block
// row: seg_ptr[s]..seg_ptr[s+1] indexes `seg_rows` (ascending training row), so every block-row
// aggregate is an order-deterministic segment sum, no atomics.
// ----------------------------------------------------------------------------------------------
template <typename Real> struct RelCache {
  const Real *card; // [S]
  Real *q, *q_S, *c, *c_S, *e, *e_q;
};

// out[i] += blk[map[i]]   (FMTrainer.hpp:309, :336)
template <typename Real>
__global__ void __launch_bounds__(256)
    k_rel_add_rows(int n, const int *__restrict__ map, const Real *__restrict__ blk,
                   Real *__restrict__ out) { // out: the e or the q component of eq (stride 2)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    out[2 * static_cast<size_t>(i)] += blk[map[i]];
}

// w part, FMTrainer.hpp:268-275: E[s] = sum e_i ; e_i -= qB[s].  Warp per block row.
template <typename Real>
__global__ void __launch_bounds__(256)
    k_rel_gather_w(int S, const int *__restrict__ seg_ptr, const int *__restrict__ seg_rows,
                   RelCache<Real> cache, Pair<Real> *__restrict__ eq) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= S)
    return;
  const Real qb = cache.q[s];
  Real acc = 0;
  for (int k = seg_ptr[s] + lane; k < seg_ptr[s + 1]; k += 32) {
    const int i = seg_rows[k];
    Real ei = eq[i].x;
    acc += ei;
    eq[i].x = ei - qb;
  }
  acc = warp_sum(acc);
  if (lane == 0)
    cache.e[s] = acc;
}

// V part, FMTrainer.hpp:396-417.
template <typename Real>
__global__ void __launch_bounds__(256)
    k_rel_gather_v(int S, const int *__restrict__ seg_ptr, const int *__restrict__ seg_rows,
                   RelCache<Real> cache, Pair<Real> *__restrict__ eq) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= S)
    return;
  const Real qb = cache.q[s], qs = cache.q_S[s];
  Real c = 0, c_S = 0, es = 0, e_q = 0;
  for (int k = seg_ptr[s] + lane; k < seg_ptr[s + 1]; k += 32) {
    const int i = seg_rows[k];
    Pair<Real> v = eq[i];
    Real ei = v.x;
    Real temp = (v.y - qb);
    c += temp;
    c_S += temp * temp;
    es += ei;
    e_q += ei * temp;
    v.y = temp;
    // 0.5 is a double literal in the reference: this expression is evaluated in double
    v.x = static_cast<Real>(ei - (temp * qb + 0.5 * qb * qb - 0.5 * qs));
    eq[i] = v;
  }
  c = warp_sum(c), c_S = warp_sum(c_S), es = warp_sum(es), e_q = warp_sum(e_q);
  if (lane == 0)
    cache.c[s] = c, cache.c_S[s] = c_S, cache.e[s] = es, cache.e_q[s] = e_q;
}

// FMTrainer.hpp:473-480
template <typename Real>
__global__ void __launch_bounds__(256)
    k_rel_resync_v(int n, const int *__restrict__ map, RelCache<Real> cache,
                   Pair<Real> *__restrict__ eq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const int s = map[i];
  Pair<Real> v = eq[i];
  const Real qb = cache.q[s], qs = cache.q_S[s], qi = v.y;
  v.x = static_cast<Real>(v.x + (qi * qb + 0.5 * qb * qb - 0.5 * qs));
  v.y = qi + qb;
  eq[i] = v;
}

// Sweep over the columns of one block, in dependency levels of the block rows.  The chain of
// levels is inherently serial and each level is tiny, so ONE persistent thread block walks it with
// block barriers instead of one launch per level.  Levels with several
// columns go warp-per-column; a level with a single column uses the whole block.
template <typename Real> struct RelSweepArgs {
  CsView<Real> Bt;      // CSC of the block (X_B^T)
  int n_levels;
  const int *level_ptr; // [n_levels + 1] into level_cols
  const int *level_cols;
  RelCache<Real> cache;
  Real *theta;          // w or V[:, r], already offset to the block's first feature
  Real *theta_t;        // feature-major mirror (offset likewise) or nullptr
  int64_t t_stride;
  const Real *z;        // offset likewise
  const int *group;     // offset likewise
  const Real *alpha;
  const Real *lambda;
  const Real *mu;
};

template <typename Real, bool IS_V>
__device__ __forceinline__ void rel_pass1(const RelSweepArgs<Real> &a, int p, Real theta_old,
                                          Real &sq, Real &lin) {
  const int s = a.Bt.idx[p];
  const Real x = a.Bt.val[p];
  if (IS_V) {
    Real h_B = (a.cache.q[s] - x * theta_old);
    Real h2 = h_B * h_B * a.cache.card[s] + 2 * a.cache.c[s] * h_B + a.cache.c_S[s];
    h2 = x * x * h2;
    sq += h2;
    lin += (-a.cache.e[s] * h_B - a.cache.e_q[s]) * x;
  } else {
    sq += (x * x) * a.cache.card[s];
    lin += (-x) * a.cache.e[s];
  }
}

template <typename Real, bool IS_V>
__device__ __forceinline__ void rel_pass2(const RelSweepArgs<Real> &a, int p, Real theta_old,
                                          Real theta_new) {
  const int s = a.Bt.idx[p];
  const Real x = a.Bt.val[p];
  const Real delta = theta_new - theta_old;
  if (IS_V) {
    Real h_B = a.cache.q[s] - x * theta_old;
    a.cache.q[s] += delta * x;
    a.cache.q_S[s] += delta * (theta_new + theta_old) * x * x;
    a.cache.e[s] += x * delta * (h_B * a.cache.card[s] + a.cache.c[s]);
    a.cache.e_q[s] += x * delta * (h_B * a.cache.c[s] + a.cache.c_S[s]);
  } else {
    a.cache.e[s] += (x * a.cache.card[s]) * delta;
  }
}

template <typename Real, bool IS_V>
__device__ __forceinline__ Real rel_draw(Real sq, Real lin, Real theta_old, Real alpha, Real lam,
                                         Real mu, Real z) {
  lin += sq * theta_old;
  if (IS_V) {
    sq *= alpha;
    lin *= alpha;
    sq += lam;
    lin += lam * mu;
  } else {
    sq = lam + alpha * sq;
    lin = alpha * lin + lam * mu;
  }
  return (lin / sq) + z / sqrt(sq);
}

template <typename Real, bool IS_V>
__global__ void __launch_bounds__(1024) k_rel_sweep(RelSweepArgs<Real> a) {
  __shared__ Real scratch[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const Real alpha = *a.alpha;
  for (int lv = 0; lv < a.n_levels; lv++) {
    const int cb = a.level_ptr[lv], ce = a.level_ptr[lv + 1];
    if (ce - cb == 1) { // the whole block works on the single column of this level
      const int l = a.level_cols[cb];
      const int b = a.Bt.ptr[l], en = a.Bt.ptr[l + 1];
      const Real theta_old = a.theta[l];
      Real sq = 0, lin = 0;
      for (int p = b + threadIdx.x; p < en; p += blockDim.x)
        rel_pass1<Real, IS_V>(a, p, theta_old, sq, lin);
      sq = block_sum(sq, scratch);
      lin = block_sum(lin, scratch);
      const int g = a.group[l];
      const Real theta_new =
          rel_draw<Real, IS_V>(sq, lin, theta_old, alpha, a.lambda[g], a.mu[g], a.z[l]);
      if (threadIdx.x == 0) {
        a.theta[l] = theta_new;
        if (a.theta_t)
          a.theta_t[static_cast<int64_t>(l) * a.t_stride] = theta_new;
      }
      for (int p = b + threadIdx.x; p < en; p += blockDim.x)
        rel_pass2<Real, IS_V>(a, p, theta_old, theta_new);
    } else {
      for (int ci = cb + wid; ci < ce; ci += nwarps) {
        const int l = a.level_cols[ci];
        const int b = a.Bt.ptr[l], en = a.Bt.ptr[l + 1];
        const Real theta_old = a.theta[l];
        Real sq = 0, lin = 0;
        for (int p = b + lane; p < en; p += 32)
          rel_pass1<Real, IS_V>(a, p, theta_old, sq, lin);
        sq = warp_sum(sq);
        lin = warp_sum(lin);
        const int g = a.group[l];
        const Real theta_new =
            rel_draw<Real, IS_V>(sq, lin, theta_old, alpha, a.lambda[g], a.mu[g], a.z[l]);
        if (lane == 0) {
          a.theta[l] = theta_new;
          if (a.theta_t)
            a.theta_t[static_cast<int64_t>(l) * a.t_stride] = theta_new;
        }
        for (int p = b + lane; p < en; p += 32)
          rel_pass2<Real, IS_V>(a, p, theta_old, theta_new);
      }
    }
    __syncthreads(); // next level reads the caches this one wrote
  }
}

// The same sweep with the block caches in SHARED memory (f32: up to ~8 000 block rows).  Blocks with
// SVD++-style implicit columns (examples/ml-1m-extended.ipynb: every column of the rated movies
// touches ~150-250 block rows and consecutive columns share rows) degenerate to one column per
// level: ~10^4 dependent steps per vector.  A level with one column is walked by ONE warp — no
// block barrier per column, shared-memory latency instead of L2 latency in every dependent
// step, the next column's entries fetched while the current one is reduced — and the other
// warps sleep at the barrier that ends the run of single-column levels (run_end, host-built).
// Arithmetic and order inside a column are those of rel_pass1 / rel_pass2 above.
template <typename Real> struct RelSmem {
  Real *card, *q, *q_S, *c, *c_S, *e, *e_q;
};

template <typename Real, bool IS_V>
__device__ __forceinline__ void rel_s_pass1(const RelSmem<Real> &m, int s, Real x, Real theta_old, Real &sq, Real &lin) {
  if (IS_V) {
    Real h_B = (m.q[s] - x * theta_old);
    Real h2 = h_B * h_B * m.card[s] + 2 * m.c[s] * h_B + m.c_S[s];
    h2 = x * x * h2;
    sq += h2;
    lin += (-m.e[s] * h_B - m.e_q[s]) * x;
  } else {
    sq += (x * x) * m.card[s];
    lin += (-x) * m.e[s];
  }
}
template <typename Real, bool IS_V>
__device__ __forceinline__ void rel_s_pass2(const RelSmem<Real> &m, int s, Real x, Real theta_old, Real theta_new) {
  const Real delta = theta_new - theta_old;
  if (IS_V) {
    Real h_B = m.q[s] - x * theta_old;
    m.q[s] += delta * x;
    m.q_S[s] += delta * (theta_new + theta_old) * x * x;
    m.e[s] += x * delta * (h_B * m.card[s] + m.c[s]);
    m.e_q[s] += x * delta * (h_B * m.c[s] + m.c_S[s]);
  } else {
    m.e[s] += (x * m.card[s]) * delta;
  }
}

constexpr int REL_TEAM = 4; // warps walking a run of single-column levels
constexpr int REL_REG = 3;  // entries per thread kept in registers (columns up to 32 * REL_TEAM * REL_REG entries)

template <typename Real, bool IS_V>
__global__ void __launch_bounds__(512)
    k_rel_sweep_smem(RelSweepArgs<Real> a, int S, const int *__restrict__ run_end, const int4 *__restrict__ level_rec) {
  extern __shared__ __align__(16) unsigned char rel_raw[];
  __shared__ Real s_red[2 * REL_TEAM];
  RelSmem<Real> m;
  {
    Real *base = reinterpret_cast<Real *>(rel_raw);
    m.card = base, m.q = base + S, m.q_S = base + 2 * S, m.c = base + 3 * S, m.c_S = base + 4 * S;
    m.e = base + 5 * S, m.e_q = base + 6 * S;
  }
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    m.card[s] = a.cache.card[s];
    m.e[s] = a.cache.e[s];
    if (IS_V)
      m.q[s] = a.cache.q[s], m.q_S[s] = a.cache.q_S[s], m.c[s] = a.cache.c[s], m.c_S[s] = a.cache.c_S[s],
      m.e_q[s] = a.cache.e_q[s];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const Real alpha = *a.alpha;
  for (int lv = 0; lv < a.n_levels;) {
    const int cb = a.level_ptr[lv], ce = a.level_ptr[lv + 1];
    if (ce - cb != 1) { // several columns, pairwise disjoint in their block rows: a warp each
      for (int ci = cb + wid; ci < ce; ci += nwarps) {
        const int l = a.level_cols[ci];
        const int b = a.Bt.ptr[l], en = a.Bt.ptr[l + 1];
        const Real theta_old = a.theta[l];
        Real sq = 0, lin = 0;
        for (int p = b + lane; p < en; p += 32)
          rel_s_pass1<Real, IS_V>(m, a.Bt.idx[p], a.Bt.val[p], theta_old, sq, lin);
        sq = warp_sum(sq);
        lin = warp_sum(lin);
        const int g = a.group[l];
        const Real theta_new = rel_draw<Real, IS_V>(sq, lin, theta_old, alpha, a.lambda[g], a.mu[g], a.z[l]);
        if (lane == 0) {
          a.theta[l] = theta_new;
          if (a.theta_t)
            a.theta_t[static_cast<int64_t>(l) * a.t_stride] = theta_new;
        }
        for (int p = b + lane; p < en; p += 32)
          rel_s_pass2<Real, IS_V>(m, a.Bt.idx[p], a.Bt.val[p], theta_old, theta_new);
      }
      __syncthreads();
      lv++;
      continue;
    }
    // A run of single-column levels [lv, lv_end): REL_TEAM warps (one per SM sub-partition, so the
    // column's few hundred instructions issue four wide) walk it with two named barriers per column;
    // the other warps sleep at the block barrier below.  Three columns are in flight: the record
    // {column, first entry, end entry, group} of column k + 2 is being loaded, the scalars and entries
    // of column k + 1 (addressed by its record) are being loaded, column k is being swept.
    const int lv_end = run_end[lv];
    if (wid < REL_TEAM) {
      constexpr int NT = 32 * REL_TEAM;
      const int tt = threadIdx.x; // 0 .. NT - 1
      // column k + 1 (set 1) and column k + 2 (set 2) in flight
      int idx_1[REL_REG], idx_2[REL_REG];
      Real val_1[REL_REG], val_2[REL_REG];
      Real theta_1 = 0, lam_1 = 0, mu_1 = 0, z_1 = 0, theta_2 = 0, lam_2 = 0, mu_2 = 0, z_2 = 0;
      int4 rec_1 = level_rec[lv];
      int4 rec_2 = lv + 1 < lv_end ? level_rec[lv + 1] : make_int4(0, 0, 0, 0);
      int4 rec_3 = lv + 2 < lv_end ? level_rec[lv + 2] : make_int4(0, 0, 0, 0);
      // static data and another column's theta: independent of the sweep's writes
#define MYFM_REL_FETCH(REC, IDX, VAL, TH, LA, MU, ZZ)                                             \
  {                                                                                                \
    TH = a.theta[REC.x], ZZ = a.z[REC.x], LA = a.lambda[REC.w], MU = a.mu[REC.w];                  \
    _Pragma("unroll") for (int k = 0; k < REL_REG; k++) {                                          \
      const int p = REC.y + tt + NT * k;                                                           \
      IDX[k] = p < REC.z ? a.Bt.idx[p] : -1;                                                       \
      VAL[k] = p < REC.z ? a.Bt.val[p] : Real(0);                                                  \
    }                                                                                              \
  }
      MYFM_REL_FETCH(rec_1, idx_1, val_1, theta_1, lam_1, mu_1, z_1)
      if (lv + 1 < lv_end)
        MYFM_REL_FETCH(rec_2, idx_2, val_2, theta_2, lam_2, mu_2, z_2)
      for (int k_lv = lv; k_lv < lv_end; k_lv++) {
        int idx[REL_REG];
        Real val[REL_REG];
#pragma unroll
        for (int k = 0; k < REL_REG; k++)
          idx[k] = idx_1[k], val[k] = val_1[k], idx_1[k] = idx_2[k], val_1[k] = val_2[k];
        const int4 rec = rec_1;
        const Real theta_old = theta_1, lam = lam_1, mu = mu_1, z = z_1;
        theta_1 = theta_2, lam_1 = lam_2, mu_1 = mu_2, z_1 = z_2;
        rec_1 = rec_2, rec_2 = rec_3;
        if (k_lv + 3 < lv_end)
          rec_3 = level_rec[k_lv + 3];
        if (k_lv + 2 < lv_end)
          MYFM_REL_FETCH(rec_2, idx_2, val_2, theta_2, lam_2, mu_2, z_2)
        const int l = rec.x, b = rec.y, en = rec.z;
        // A column holds a block row at most once, so its entries touch pairwise different cache
        // slots: all loads of the column are issued before the first store (written one entry at a
        // time the compiler would have to order every store before the next entry's loads).
        Real qv[REL_REG], cardv[REL_REG], cv[REL_REG], cSv[REL_REG], ev[REL_REG], eqv[REL_REG], qSv[REL_REG];
        Real sq = 0, lin = 0;
#pragma unroll
        for (int k = 0; k < REL_REG; k++)
          if (idx[k] >= 0) {
            const int sr = idx[k];
            cardv[k] = m.card[sr], ev[k] = m.e[sr];
            if (IS_V)
              qv[k] = m.q[sr], cv[k] = m.c[sr], cSv[k] = m.c_S[sr], eqv[k] = m.e_q[sr], qSv[k] = m.q_S[sr];
          }
#pragma unroll
        for (int k = 0; k < REL_REG; k++)
          if (idx[k] >= 0) { // rel_pass1
            const Real x = val[k];
            if (IS_V) {
              Real h_B = (qv[k] - x * theta_old);
              Real h2 = h_B * h_B * cardv[k] + 2 * cv[k] * h_B + cSv[k];
              h2 = x * x * h2;
              sq += h2;
              lin += (-ev[k] * h_B - eqv[k]) * x;
            } else {
              sq += (x * x) * cardv[k];
              lin += (-x) * ev[k];
            }
          }
        for (int p = b + tt + NT * REL_REG; p < en; p += NT) // the tail of a long column
          rel_s_pass1<Real, IS_V>(m, a.Bt.idx[p], a.Bt.val[p], theta_old, sq, lin);
        sq = warp_sum(sq);
        lin = warp_sum(lin);
        if (lane == 0)
          s_red[2 * wid] = sq, s_red[2 * wid + 1] = lin;
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        sq = 0, lin = 0;
#pragma unroll
        for (int w = 0; w < REL_TEAM; w++) // warp partials in warp order: the same sums in every thread
          sq += s_red[2 * w], lin += s_red[2 * w + 1];
        const Real theta_new = rel_draw<Real, IS_V>(sq, lin, theta_old, alpha, lam, mu, z);
        if (tt == 0) {
          a.theta[l] = theta_new;
          if (a.theta_t)
            a.theta_t[static_cast<int64_t>(l) * a.t_stride] = theta_new;
        }
        const Real delta = theta_new - theta_old;
#pragma unroll
        for (int k = 0; k < REL_REG; k++)
          if (idx[k] >= 0) { // rel_pass2
            const Real x = val[k];
            const int sr = idx[k];
            if (IS_V) {
              Real h_B = qv[k] - x * theta_old;
              m.q[sr] = qv[k] + delta * x;
              m.q_S[sr] = qSv[k] + delta * (theta_new + theta_old) * x * x;
              m.e[sr] = ev[k] + x * delta * (h_B * cardv[k] + cv[k]);
              m.e_q[sr] = eqv[k] + x * delta * (h_B * cv[k] + cSv[k]);
            } else {
              m.e[sr] = ev[k] + (x * cardv[k]) * delta;
            }
          }
        for (int p = b + tt + NT * REL_REG; p < en; p += NT)
          rel_s_pass2<Real, IS_V>(m, a.Bt.idx[p], a.Bt.val[p], theta_old, theta_new);
        // the next column reads the caches this one wrote (and s_red is free again)
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
      }
#undef MYFM_REL_FETCH
    }
    __syncthreads();
    lv = lv_end;
  }
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    a.cache.e[s] = m.e[s];
    if (IS_V)
      a.cache.q[s] = m.q[s], a.cache.q_S[s] = m.q_S[s], a.cache.e_q[s] = m.e_q[s];
  }
}

// ----------------------------------------------------------------------------------------------
// Small utilities
// ----------------------------------------------------------------------------------------------
// Boundary copies of one component of eq between device row order and the caller's row order:
// dense[perm[i]] = eq[i].c  /  eq[i].c = dense[perm[i]]   (comp 0 = e, 1 = q)
template <typename Real>
__global__ void __launch_bounds__(256)
    k_eq_export(int64_t n, const Real *__restrict__ eq, int comp, const int *__restrict__ perm,
                Real *__restrict__ dense) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n)
    dense[perm[i]] = eq[2 * i + comp];
}
template <typename Real>
__global__ void __launch_bounds__(256)
    k_eq_import(int64_t n, Real *__restrict__ eq, int comp, const int *__restrict__ perm,
                const Real *__restrict__ dense) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n)
    eq[2 * i + comp] = dense[perm[i]];
}

// Vt[j*K + r] = V[j + D*r]
template <typename Real>
__global__ void __launch_bounds__(256)
    k_transpose_V(int64_t D, int K, const Real *__restrict__ V, Real *__restrict__ Vt) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t < D * K) {
    const int64_t j = t / K;
    const int r = static_cast<int>(t % K);
    Vt[t] = V[j + D * r];
  }
}

// Merge of the owned columns (row shards with a rank-exclusive first field): pack = [w | V] with the entries of
// columns another rank owns zeroed, so that ONE all-reduce over the ranks leaves the complete sample; unpack
// writes it back to w, V and the feature-major mirror Vt.  t runs over D * (K + 1) entries.
template <typename Real>
__global__ void __launch_bounds__(256)
    k_merge_pack(int64_t D, int K, const int *__restrict__ owner, int my_rank, const Real *__restrict__ w,
                 const Real *__restrict__ V, Real *__restrict__ pack) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t < D * (K + 1)) {
    const bool mine = owner[t % D] == my_rank;
    pack[t] = mine ? (t < D ? w[t] : V[t - D]) : Real(0);
  }
}
template <typename Real>
__global__ void __launch_bounds__(256)
    k_merge_unpack(int64_t D, int K, const Real *__restrict__ pack, Real *__restrict__ w, Real *__restrict__ V,
                   Real *__restrict__ Vt) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t < D)
    w[t] = pack[t];
  if (t < D * K) { // t as an index of Vt: coalesced writes of Vt, strided reads of the pack
    const int64_t j = t / K;
    const int r = static_cast<int>(t % K);
    const Real v = pack[D + j + D * r];
    Vt[t] = v;
    V[j + D * r] = v;
  }
}

template <typename Real>
__global__ void __launch_bounds__(256) k_fill_strided(int64_t n, Real *p, int stride, Real v) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n)
    p[i * stride] = v;
}

template <typename Real>
__global__ void __launch_bounds__(256) k_fill(int64_t n, Real *p, Real v) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n)
    p[i] = v;
}

// out[i] (+)= link(score[i]);  LINK 0: identity, 1: Phi(score) = (erf(score*sqrt(.5))+1)/2
// (predictor.hpp:136-143)
template <typename Real, int LINK>
__global__ void __launch_bounds__(256)
    k_accumulate(int n, const Real *__restrict__ score, Real *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  Real s = score[i];
  if (LINK == 1)
    s = (erf(s * static_cast<Real>(0.70710678118654752440)) + static_cast<Real>(1)) /
        static_cast<Real>(2);
  out[i] += s;
}

// FM.hpp:137-162 accumulated over samples; out is [n x (n_cpt+1)] row-major
template <typename Real>
__global__ void __launch_bounds__(256)
    k_accumulate_oprobit(int n, const Real *__restrict__ score, const Real *__restrict__ cutpoints,
                         int n_cpt, Real *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const Real s = score[i];
  Real prev = 0;
  for (int c = 0; c < n_cpt; c++) {
    Real cdf = (1 + erf((cutpoints[c] - s) * static_cast<Real>(0.70710678118654752440))) / 2;
    out[static_cast<size_t>(i) * (n_cpt + 1) + c] += cdf - prev;
    prev = cdf;
  }
  out[static_cast<size_t>(i) * (n_cpt + 1) + n_cpt] += 1 - prev;
}

// widening copy for the getters (the boundary speaks float64)
template <typename Real>
__global__ void __launch_bounds__(256) k_to_double(int64_t n, const Real *__restrict__ in, double *__restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n)
    out[i] = static_cast<double>(in[i]);
}

template <typename Real>
__global__ void __launch_bounds__(256) k_scale(int64_t n, Real *p, Real inv) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n)
    p[i] /= inv;
}

} // namespace myfm
