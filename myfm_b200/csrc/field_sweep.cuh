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
// rows per lane kept in registers between the reduction and the update: 8 (f32), 4 (f64)
constexpr int FIELD_BATCH_MAX = 8;   // level-0 columns a warp takes per scheduling step (a.batch, at most 32)
constexpr int FIELD_CTA_MAX = 32768;  // longest level-0 column (one CTA); longer: general path
constexpr int STATS_THREADS = 256;
constexpr int STATS_WARP_MAX = 2048;  // last level: warp per column up to here (no barrier on the path),
constexpr int STATS_CHUNK = 8192;     // one CTA per column up to here, chunks of this size beyond

enum { PEND_NONE = 0, PEND_W = 1, PEND_V = 2 };

// ---- TMA (1-D bulk copy) and mbarrier primitives ------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WAIT_LOOP:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra WAIT_DONE;\n"
               "bra WAIT_LOOP;\n"
               "WAIT_DONE:\n"
               "}" ::"r"(smem_addr(bar)),
               "r"(parity)
               : "memory");
}
// global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src, uint32_t bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_store(void *dst, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// writes made through the generic proxy (ordinary st.shared) become visible to the async proxy (TMA)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }




// ------------------------------------------------------------------------------------------------
// Row shards on one NVLink / NVSwitch node: the all-reduce of the column statistics is fused into
// the kernels through peer memory instead of being a collective of its own.  Every rank owns a
// buffer that all peers map (cudaIpc): a header with the sequence number of the last collective
// whose local statistics are complete, then two statistics buffers used alternately.  The
// producing kernel writes this rank's partial sums into its own buffer and its last CTA publishes the
// sequence number, and the consuming kernel — after seeing the sequence number on every peer —
// reads the partial sums of ALL ranks straight over NVLink and adds them in rank order, so every
// rank forms bit-identical sums.  A rank cannot run two collectives ahead of a peer (it needs the
// peer's sequence number first), which makes two buffers enough.
// ------------------------------------------------------------------------------------------------
constexpr int PEER_MAX_RANKS = 16;
constexpr int PEER_HEADER_BYTES = 256;
template <typename Real> struct PeerView {
  int world = 0;                 // 0: not in use (one GPU, or NCCL all-reduce between the passes)
  // Collectives are numbered by a counter in device memory (peer_post_when_last increments it), so the
  // same launch parameters serve every sweep and the sequence can be replayed from a CUDA graph.
  unsigned long long *counter;       // collectives published by THIS rank so far
  unsigned long long *my_posted;     // where this rank publishes that number for the peers
  unsigned int *done;                // CTAs of the producing kernel that have finished (zero between launches)
  size_t elems;                      // Reals per statistics buffer; collective c uses buffer c & 1
  const Real *stat[PEER_MAX_RANKS];  // every rank's statistics buffers (buffer 0, then buffer 1)
  const unsigned long long *posted[PEER_MAX_RANKS]; // every rank's published sequence number
  int *error;                    // set when a peer never shows up
  unsigned long long timeout_ns; // how long peer_wait polls before it gives up (fatal)
  // MYFM_PEER_TRACE: %globaltimer stamps per collective, PEER_TRACE_SLOTS per record, a ring of trace_cap
  // records (nullptr: off).  The multi-GPU timeline of DESIGN.md section 5 is made from these.
  unsigned long long *trace = nullptr;
  int trace_cap = 0;
  // producer: where this rank's partial statistics of the NEXT collective go
  __device__ __forceinline__ Real *produce(Real *local_stat) const {
    return local_stat + ((*counter + 1) & 1) * elems;
  }
};

// Trace slots: 0 statistics kernel starts, 1 its last CTA publishes, 2 draw kernel starts, 3 every peer has
// published, 4 draw kernel (block 0) done, 5 next streaming pass starts, 6 its block 0 is done.
constexpr int PEER_TRACE_SLOTS = 8;
__device__ __forceinline__ unsigned long long peer_clock_ns();
template <typename Real>
__device__ __forceinline__ void peer_trace(const PeerView<Real> &pv, int slot, unsigned long long ahead = 0) {
  if (pv.world == 0 || pv.trace == nullptr)
    return;
  const unsigned long long seq = *pv.counter + ahead;
  pv.trace[(seq % static_cast<unsigned long long>(pv.trace_cap)) * PEER_TRACE_SLOTS + slot] = peer_clock_ns();
  if (slot == 0)
    pv.trace[(seq % static_cast<unsigned long long>(pv.trace_cap)) * PEER_TRACE_SLOTS + 7] = seq;
}

// MYFM_PEER_TRACE, phases of a sweep outside the column exchange: a one-thread kernel between the phases stamps
// phase[(sweep % cap) * PEER_TRACE_SLOTS + slot]; slot 0 also advances the sweep counter.
__global__ void k_trace_stamp(unsigned long long *phase, unsigned long long *sweep_counter, int cap, int slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  if (slot == 0)
    *sweep_counter += 1;
  phase[(*sweep_counter % static_cast<unsigned long long>(cap)) * PEER_TRACE_SLOTS + slot] = t;
}

// End of a producing kernel: the CTA that finishes last publishes collective *counter + 1 (no
// separate launch).  Called by every thread of every CTA after its last statistics store.
template <typename Real> __device__ __forceinline__ void peer_post_when_last(const PeerView<Real> &pv) {
  if (pv.world == 0)
    return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence(); // this CTA's statistics before the count
    if (atomicAdd(pv.done, 1u) == gridDim.x - 1) {
      *pv.done = 0;
      peer_trace(pv, 1, 1);
      const unsigned long long seq = *pv.counter + 1;
      *pv.counter = seq;
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long *>(pv.my_posted) = seq;
    }
  }
}

// Block-wide: returns once every rank has published `seq`.  A peer that does not show up within
// pv.timeout_ns (a rank that died, or one held up on the host for that long: MYFM_PEER_TIMEOUT_S,
// default 120 s) is fatal: the flag is raised and the kernel traps, so the chain can never continue
// on stale statistics — the next call on this trainer fails with a CUDA error.
__device__ __forceinline__ unsigned long long peer_clock_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
template <typename Real> __device__ __forceinline__ void peer_wait(const PeerView<Real> &pv) {
  if (pv.world == 0)
    return;
  if (threadIdx.x < pv.world) {
    const unsigned long long seq = *pv.counter; // the collective this rank published last
    const volatile unsigned long long *flag = pv.posted[threadIdx.x];
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while (*flag < seq) {
      if ((++spins & 0xfffu) == 0) { // look at the clock every 4096 polls
        const unsigned long long now = peer_clock_ns();
        if (t0 == 0)
          t0 = now;
        else if (now - t0 > pv.timeout_ns) {
          *pv.error = 3;
          __threadfence_system();
          __trap();
        }
      }
    }
    __threadfence_system();
  }
  __syncthreads();
}
// (sq, lin) of column `slot` summed over the ranks in rank order.  All remote loads are issued
// before the first add: a load-add-load-add loop would pay the NVLink latency once per rank.
template <typename Real> __device__ __forceinline__ void peer_sum(const PeerView<Real> &pv, int slot, Real &sq, Real &lin) {
  const size_t at = (*pv.counter & 1) * pv.elems + 2 * static_cast<size_t>(slot);
  Pair<Real> v[PEER_MAX_RANKS];
#pragma unroll
  for (int r = 0; r < PEER_MAX_RANKS; r++)
    if (r < pv.world)
      v[r] = __ldcv(reinterpret_cast<const Pair<Real> *>(pv.stat[r] + at));
  sq = 0, lin = 0;
#pragma unroll
  for (int r = 0; r < PEER_MAX_RANKS; r++)
    if (r < pv.world)
      sq += v[r].x, lin += v[r].y;
}

// update_alpha + update_w0 on row shards over the same exchange (instead of an NCCL all-reduce of two scalars):
// one block folds this rank's block partials (k_reduce_e_both) into slot 0 of its statistics buffer and
// publishes; the draw kernel waits for every rank and adds the pairs in rank order.
template <typename Real> __global__ void k_fold_partials2_peer(int n, const Real *__restrict__ partial, PeerView<Real> pv, Real *peer_local) {
  __shared__ Real scratch[32];
  Real acc0 = 0, acc1 = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    acc0 += partial[i], acc1 += partial[n + i];
  acc0 = block_sum(acc0, scratch);
  acc1 = block_sum(acc1, scratch);
  if (threadIdx.x == 0) {
    Real *out = pv.produce(peer_local);
    out[0] = acc0, out[1] = acc1;
  }
  peer_post_when_last(pv);
}
template <typename Real>
__global__ void k_finish_alpha_w0_peer(PeerView<Real> pv, Real beta_0, const Real *__restrict__ g_std, Real *alpha_ptr,
                                       int n_train, Real reg_0, const Real *__restrict__ z, Real *w0, Real *delta) {
  peer_wait(pv);
  if (threadIdx.x == 0) {
    Real acc2, acc1;
    peer_sum(pv, 0, acc2, acc1);
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

template <typename Real> struct FieldStreamArgs {
  const int4 *item; // level-0 columns {column, first row, end row, -}, longest first (classes: k_field_stream)
  int nCC, nCR, nG, nW;
  int batch;        // warp columns taken per scheduling step: 1 .. FIELD_BATCH_MAX (few columns: small batches)
  int *sched;       // work counter of the warp items (zero at launch)
  Pair<Real> *eq;
  int64_t n_rows;
  int n_tail;           // L - 1
  const int *tail_idx;  // [n_tail][n_rows]: column of the row's entries 1 .. L-1
  const Real *tail_val; // [n_tail][n_rows], unused when UNIT
  const int *tail_last;  // = tail_idx + (n_tail - 1) * n_rows: the last field
  const Real *tval_last; // likewise
  const Real *own_val;  // [n_rows] value of the level-0 entry, unused when UNIT
  Real *theta;          // w, or column r of V (column-major)
  Real *theta_t;        // feature-major mirror: theta_t[j * t_stride], or nullptr
  int64_t t_stride;
  const Real *z;
  const int *group;
  const Real *alpha, *lambda, *mu;
  int last_base, n_tab; // the last level's columns are [last_base, last_base + n_tab)
  const Real *pend_told, *pend_tnew; // [n_tab] draw left pending by the previous vector
  // row-sharded training (FIELD_STATS / FIELD_UPDATE): statistics of every level-0 column, summed
  // over the ranks between the two passes
  const int *item_slot; // column slot of every item (same numbering on every rank)
  Real *colstat;        // [2 * columns of level 0]
  PeerView<Real> peer;  // peer-memory exchange (world == 0: colstat holds the statistics / NCCL sums)
  Real *peer_local;     // this rank's statistics buffers (FIELD_STATS writes the next collective's)
};

// One GPU: FIELD_FUSED (statistics, draw and update in one pass).  Row shards: the rows of a column
// live on several GPUs, so the pass runs twice around an all-reduce of the column statistics —
// FIELD_STATS leaves (sum h^2, sum -e h) per column, FIELD_UPDATE draws from the summed statistics
// (identically on every rank) and writes e, q.
enum { FIELD_FUSED = 0, FIELD_STATS = 1, FIELD_UPDATE = 2 };

// Shared-memory table of the last level's columns: {theta_old, theta_new} of the update left
// pending and theta of the vector being swept.  f32: one 16-byte record per column (one LDS.128 per
// row instead of three scattered LDS.32); f64: three arrays.
// COMPACT (f32 with staged rows, below): 12 bytes per column — {theta_old, theta_new} pairs, then theta_next.
template <typename Real, bool COMPACT = false> struct FieldTab;
template <> struct FieldTab<float, false> {
  static constexpr int BYTES_PER_COLUMN = 16;
  static __device__ __forceinline__ void get(const float *base, int, int j, float &told, float &tnew, float &tnext) {
    const float4 r = reinterpret_cast<const float4 *>(base)[j];
    told = r.x, tnew = r.y, tnext = r.z;
  }
  static __device__ __forceinline__ void put(float *base, int, int j, float told, float tnew, float tnext) {
    reinterpret_cast<float4 *>(base)[j] = make_float4(told, tnew, tnext, 0.f);
  }
};
template <> struct FieldTab<float, true> {
  static constexpr int BYTES_PER_COLUMN = 12;
  static __device__ __forceinline__ void get(const float *base, int n, int j, float &told, float &tnew, float &tnext) {
    const float2 r = reinterpret_cast<const float2 *>(base)[j];
    told = r.x, tnew = r.y, tnext = base[2 * n + j];
  }
  static __device__ __forceinline__ void put(float *base, int n, int j, float told, float tnew, float tnext) {
    reinterpret_cast<float2 *>(base)[j] = make_float2(told, tnew);
    base[2 * n + j] = tnext;
  }
};
template <bool COMPACT> struct FieldTab<double, COMPACT> {
  static constexpr int BYTES_PER_COLUMN = 24;
  static __device__ __forceinline__ void get(const double *base, int n, int j, double &told, double &tnew,
                                             double &tnext) {
    told = base[j], tnew = base[n + j], tnext = base[2 * n + j];
  }
  static __device__ __forceinline__ void put(double *base, int n, int j, double told, double tnew, double tnext) {
    base[j] = told, base[n + j] = tnew, base[2 * n + j] = tnext;
  }
};

// Rows of a warp's column staged in shared memory by TMA bulk copies (k_field_stream<..., STAGED>):
// eq[i - eq_row0], idx[i - idx_row0] for the rows i of the column (the copies start on 16-byte
// boundaries, so they begin up to one / three rows early).
template <typename Real> struct FieldStage {
  const Pair<Real> *eq = nullptr;
  const int *idx = nullptr;
  int eq_row0 = 0, idx_row0 = 0;
};
constexpr int FIELD_STAGE_ROWS = 256; // the warp class: up to 32 lanes x 8 rows (f32)
constexpr int FIELD_STAGE_IDX_BYTES = (FIELD_STAGE_ROWS + 8) * 4;
template <typename Real> struct FieldStageSize {
  static constexpr int EQ_BYTES = (FIELD_STAGE_ROWS + 2) * static_cast<int>(sizeof(Pair<Real>));
  static constexpr int BYTES = EQ_BYTES + FIELD_STAGE_IDX_BYTES;
};

// U rows of the streaming pass in flight per thread: all global loads are issued first (load),
// the shared-memory lookups and the arithmetic follow (finish) — a row's table index comes out of
// its own load, so interleaving the two would serialise the rows on the memory latency.
template <typename Real, bool IS_V, bool UNIT, bool HAS_MID, int PEND, int U, bool COMPACT = false> struct FieldBatch {
  static constexpr bool NEED_LAST = IS_V || PEND != PEND_NONE;
  Pair<Real> v[U];
  int jl[U];
  Real xl[U], x0[U];

  // the same from the warp's staged copy of its column (UNIT tables only)
  __device__ __forceinline__ void load(const FieldStreamArgs<Real> &a, const FieldStage<Real> &st, int u, int i) {
    v[u] = st.eq[i - st.eq_row0];
    jl[u] = 0, xl[u] = Real(1), x0[u] = Real(1);
    if (NEED_LAST)
      jl[u] = st.idx[i - st.idx_row0] - a.last_base;
  }
  __device__ __forceinline__ void load(const FieldStreamArgs<Real> &a, int u, int i) {
    v[u] = __ldcg(a.eq + i);
    jl[u] = 0, xl[u] = Real(1), x0[u] = Real(1);
    if (NEED_LAST)
      jl[u] = __ldcs(a.tail_last + i) - a.last_base;
    if (!UNIT) {
      x0[u] = __ldcs(a.own_val + i);
      if (NEED_LAST)
        xl[u] = __ldcs(a.tval_last + i);
    }
  }

  // Row i as the level-0 column sees it: e with the pending update applied, q = x_i . V[:, r]
  // (IS_V; the stored q otherwise) and the row's level-0 value x0.
  __device__ __forceinline__ void finish(const FieldStreamArgs<Real> &a, const Real *s_tab, int u, int i, Real theta_old, Real &e, Real &q,
                                         Real &x0_out) const {
    e = v[u].x, q = v[u].y, x0_out = x0[u];
    const Real xlu = xl[u];
    Real told = 0, tnew = 0, tnext = 0;
    if (NEED_LAST)
      FieldTab<Real, COMPACT>::get(s_tab, a.n_tab, jl[u], told, tnew, tnext);
    if (PEND == PEND_V) { // FMTrainer.hpp:366-374 of the previous factor's last-level column
      const Real h = xlu * (q - xlu * told);
      e = e + h * (tnew - told);
    } else if (PEND == PEND_W) { // FMTrainer.hpp:240,251
      e = (e - xlu * told) + xlu * tnew;
    }
    if (IS_V) { // q_init in CSR order (FMTrainer.hpp:320)
      Real acc = x0[u] * theta_old;
      if (HAS_MID) {
#pragma unroll 1
        for (int k = 0; k + 1 < a.n_tail; k++) {
          const int64_t off = static_cast<int64_t>(k) * a.n_rows + i;
          const Real x = UNIT ? Real(1) : __ldcs(a.tail_val + off);
          acc += x * a.theta[__ldcs(a.tail_idx + off)];
        }
      }
      acc += xlu * tnext;
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

// One pass of NT cooperating threads (t = 0 .. NT-1) over the rows [lo, hi) of a column, U rows
// per thread in flight.  UPDATE = false: accumulates the statistics; true: writes e and q back.
template <typename Real, bool IS_V, bool UNIT, bool HAS_MID, int PEND, int NT, bool UPDATE, bool COMPACT = false>
__device__ __forceinline__ void field_pass(const FieldStreamArgs<Real> &a, const Real *s_tab, int lo, int hi, int t,
                                           Real theta_old, Real theta_new, Real alpha, Real &sq, Real &lin) {
  constexpr int U = 4;
  for (int base = lo + t; base < hi; base += U * NT) {
    FieldBatch<Real, IS_V, UNIT, HAS_MID, PEND, U, COMPACT> b;
#pragma unroll
    for (int u = 0; u < U; u++)
      b.load(a, u, min(base + u * NT, hi - 1));
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int i = base + u * NT;
      if (i < hi) {
        Real e, q, x0;
        b.finish(a, s_tab, u, i, theta_old, e, q, x0);
        if (UPDATE)
          __stcg(a.eq + i, field_update<Real, IS_V>(e, q, x0, theta_old, theta_new));
        else
          field_stats<Real, IS_V>(e, q, x0, theta_old, alpha, sq, lin);
      }
    }
  }
}

constexpr int FIELD_WARPS = FIELD_THREADS / 32;
constexpr int FIELD_GROUP_WARPS = 4; // medium columns: groups of four warps

// (sq, lin) summed over the GW warps that share a column.  GW = 1: the warp alone; FIELD_GROUP_WARPS:
// a 128-thread group with its own named barrier; FIELD_WARPS: the CTA.  Warp partials are added in
// warp order (deterministic).  `parity` alternates between a group's consecutive columns, so one
// barrier per column is enough.
template <typename Real, int GW>
__device__ __forceinline__ void field_group_sum(Real &sq, Real &lin, Real *s_part, int parity, int grp, int wig,
                                                int lane) {
  sq = warp_sum(sq);
  lin = warp_sum(lin);
  if (GW == 1)
    return;
  Real *buf = s_part + ((parity * (FIELD_WARPS / GW) + grp) * GW) * 2;
  if (lane == 0)
    buf[2 * wig] = sq, buf[2 * wig + 1] = lin;
  if (GW == FIELD_WARPS)
    __syncthreads();
  else
    asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(GW * 32) : "memory");
  Real s = 0, l = 0;
#pragma unroll
  for (int w = 0; w < GW; w++)
    s += buf[2 * w], l += buf[2 * w + 1];
  sq = s, lin = l;
}

// A column of at most 32 * GW * NS rows handled by GW warps (thread t of the group): the rows stay
// in registers between the reduction and the update, so every row is read once and written once.
// FULL: slots 0 .. NS-2 hold a row in every thread (the caller picked NS = ceil(rows / threads)),
// only the last slot is ragged; otherwise every slot is checked.
// STAGED: the rows come from the warp's shared-memory copy `st`; `after_rows()` runs once they sit in
// registers (the warp then starts the copy of its next column).
struct FieldNoHook {
  __device__ __forceinline__ void operator()() const {}
};
template <typename Real, bool IS_V, bool UNIT, bool HAS_MID, int PEND, int MODE, int GW, int NS, bool FULL,
          bool COMPACT = false, bool STAGED = false, typename Hook = FieldNoHook>
__device__ __forceinline__ Real field_column_regs(const FieldStreamArgs<Real> &a, const Real *s_tab, int4 it, int t,
                                                  Real theta_old, Real alpha, Real lam, Real mu, Real z,
                                                  Real *s_part, int parity, int grp, int wig, int lane,
                                                  Real theta_given, Real &sq, Real &lin,
                                                  const FieldStage<Real> &st = FieldStage<Real>(),
                                                  Hook after_rows = Hook()) {
  constexpr int NT = 32 * GW;
  FieldBatch<Real, IS_V, UNIT, HAS_MID, PEND, NS, COMPACT> b;
  Real e[NS], q[NS], x0[NS];
  const int i0 = it.y + t;
  bool ok[NS];
#pragma unroll
  for (int s = 0; s < NS; s++) {
    ok[s] = (FULL && s < NS - 1) || i0 + NT * s < it.z;
    if (STAGED)
      b.load(a, st, s, ok[s] ? i0 + NT * s : it.z - 1);
    else
      b.load(a, s, ok[s] ? i0 + NT * s : it.z - 1);
  }
  sq = 0, lin = 0;
#pragma unroll
  for (int s = 0; s < NS; s++) {
    b.finish(a, s_tab, s, ok[s] ? i0 + NT * s : it.z - 1, theta_old, e[s], q[s], x0[s]);
    if (MODE != FIELD_UPDATE && ok[s])
      field_stats<Real, IS_V>(e[s], q[s], x0[s], theta_old, alpha, sq, lin);
  }
  after_rows();
  if (MODE != FIELD_UPDATE)
    field_group_sum<Real, GW>(sq, lin, s_part, parity, grp, wig, lane);
  if (MODE == FIELD_STATS)
    return theta_old;
  const Real theta_new =
      MODE == FIELD_UPDATE ? theta_given : column_draw<Real, IS_V>(sq, lin, theta_old, alpha, lam, mu, z);
#pragma unroll
  for (int s = 0; s < NS; s++)
    if (ok[s])
      __stcg(a.eq + i0 + NT * s, field_update<Real, IS_V>(e[s], q[s], x0[s], theta_old, theta_new));
  return theta_new;
}

// Picks the instantiation for the column's slot count (warp-uniform switch).
template <typename Real, bool IS_V, bool UNIT, bool HAS_MID, int PEND, int MODE, int GW, bool COMPACT = false,
          bool STAGED = false, typename Hook = FieldNoHook>
__device__ __forceinline__ Real field_column_dispatch(const FieldStreamArgs<Real> &a, const Real *s_tab, int4 it, int t,
                                                      Real theta_old, Real alpha, Real lam, Real mu, Real z,
                                                      Real *s_part, int parity, int grp, int wig, int lane,
                                                      Real theta_given, Real &sq, Real &lin,
                                                      const FieldStage<Real> &st = FieldStage<Real>(),
                                                      Hook after_rows = Hook()) {
  constexpr int R = sizeof(Real) == 8 ? 4 : 8, NT = 32 * GW;
  const int n_slots = (it.z - it.y + NT - 1) / NT;
#define MYFM_SLOTS(NS)                                                                             \
  case NS:                                                                                         \
    return field_column_regs<Real, IS_V, UNIT, HAS_MID, PEND, MODE, GW, (NS <= R ? NS : R), true, COMPACT, STAGED,  \
                             Hook>(a, s_tab, it, t, theta_old, alpha, lam, mu, z, s_part, parity, grp, wig, lane,  \
                                   theta_given, sq, lin, st, after_rows);
  switch (n_slots) {
    MYFM_SLOTS(1)
    MYFM_SLOTS(2)
    MYFM_SLOTS(3)
    MYFM_SLOTS(4)
    MYFM_SLOTS(5)
    MYFM_SLOTS(6)
    MYFM_SLOTS(7)
    MYFM_SLOTS(8)
  }
#undef MYFM_SLOTS
  // no local rows (an empty column, or a column whose rows live on other ranks)
  after_rows();
  sq = 0, lin = 0;
  if (MODE == FIELD_STATS)
    return theta_old;
  return MODE == FIELD_UPDATE ? theta_given : column_draw<Real, IS_V>(Real(0), Real(0), theta_old, alpha, lam, mu, z);
}

// What the leader of a column does with the result: store the draw, or (FIELD_STATS) the statistics.
template <typename Real, int MODE>
__device__ __forceinline__ void field_finish_column(const FieldStreamArgs<Real> &a, int j, int slot, Real theta_new,
                                                    Real sq, Real lin) {
  if (MODE == FIELD_STATS) {
    Real *out = a.peer.world ? a.peer.produce(a.peer_local) : a.colstat;
    out[2 * slot] = sq, out[2 * slot + 1] = lin;
  } else {
    a.theta[j] = theta_new;
    if (a.theta_t)
      a.theta_t[static_cast<int64_t>(j) * a.t_stride] = theta_new;
  }
}

// Level-0 columns by length (items are sorted longest first; R = 8 rows per thread in f32, 4 in f64):
//   nCC  longer than 32 R FIELD_WARPS rows: the whole CTA, two passes (the second re-reads through L2)
//   nCR  up to 32 R FIELD_WARPS rows:       the whole CTA, rows in registers
//   nG   up to 32 R FIELD_GROUP_WARPS rows: a group of four warps, rows in registers
//   nW   up to 32 R rows:                   one warp, rows in registers, handed out dynamically
// STAGED (UNIT tables, FIELD_FUSED): the warp class reads its rows from a per-warp shared-memory
// buffer that a TMA bulk copy (cp.async.bulk, mbarrier-signalled) filled while the warp was busy
// with the previous column — the copy of column c + 1 is issued as soon as column c's rows sit in
// registers, so its latency hides behind c's reduction, draw and write-back.  The table then uses
// the 12-byte layout to leave room for the 32 buffers.
template <typename Real, bool IS_V, bool UNIT, bool HAS_MID, int PEND, int MODE, bool STAGED = false>
__global__ void __launch_bounds__(FIELD_THREADS, 1) k_field_stream(const __grid_constant__ FieldStreamArgs<Real> a) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ Real scratch[32];
  __shared__ Real s_part_cta[2 * FIELD_WARPS * 2], s_part_grp[2 * FIELD_WARPS * 2];
  __shared__ __align__(8) unsigned long long s_wbar[FIELD_WARPS];
  constexpr bool COMPACT = STAGED;
  using Tab = FieldTab<Real, COMPACT>;
  Real *s_tab = reinterpret_cast<Real *>(s_raw);
  if (blockIdx.x == 0 && threadIdx.x == 0 && MODE == FIELD_FUSED)
    peer_trace(a.peer, 5);
  for (int t = threadIdx.x; t < a.n_tab; t += FIELD_THREADS)
    Tab::put(s_tab, a.n_tab, t, PEND != PEND_NONE ? a.pend_told[t] : Real(0),
             PEND != PEND_NONE ? a.pend_tnew[t] : Real(0), IS_V ? a.theta[a.last_base + t] : Real(0));
  if (STAGED && (threadIdx.x & 31) == 0)
    mbar_init(&s_wbar[threadIdx.x >> 5], 1);
  __syncthreads();
  const Real alpha = *a.alpha;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // draw of column (j, slot) from the statistics summed over the ranks (FIELD_UPDATE)
  auto draw_given = [&](int slot, Real theta_old, Real lam, Real mu, Real z) {
    Real sq, lin;
    if (a.peer.world)
      peer_sum(a.peer, slot, sq, lin);
    else
      sq = a.colstat[2 * slot], lin = a.colstat[2 * slot + 1];
    return column_draw<Real, IS_V>(sq, lin, theta_old, alpha, lam, mu, z);
  };
  if (MODE == FIELD_UPDATE)
    peer_wait(a.peer);

  for (int c = blockIdx.x; c < a.nCC; c += gridDim.x) {
    const int4 it = __ldg(a.item + c);
    const int j = it.x, slot = MODE == FIELD_FUSED ? 0 : a.item_slot[c];
    const Real theta_old = a.theta[j];
    const int g = a.group[j];
    const Real lam = a.lambda[g], mu = a.mu[g], z = a.z[j];
    Real sq = 0, lin = 0, theta_new = theta_old;
    if (MODE != FIELD_UPDATE) {
      field_pass<Real, IS_V, UNIT, HAS_MID, PEND, FIELD_THREADS, false, COMPACT>(a, s_tab, it.y, it.z, threadIdx.x,
                                                                                 theta_old, theta_old, alpha, sq, lin);
      sq = block_sum(sq, scratch);
      lin = block_sum(lin, scratch);
    }
    if (MODE != FIELD_STATS) {
      if (MODE == FIELD_UPDATE) {
        if (lane == 0)
          theta_new = draw_given(slot, theta_old, lam, mu, z);
        theta_new = __shfl_sync(FULL_MASK, theta_new, 0);
      } else {
        theta_new = column_draw<Real, IS_V>(sq, lin, theta_old, alpha, lam, mu, z);
      }
      field_pass<Real, IS_V, UNIT, HAS_MID, PEND, FIELD_THREADS, true, COMPACT>(a, s_tab, it.y, it.z, threadIdx.x,
                                                                                theta_old, theta_new, alpha, sq, lin);
    }
    __syncthreads(); // every thread has read theta[j]
    if (threadIdx.x == 0)
      field_finish_column<Real, MODE>(a, j, slot, theta_new, sq, lin);
  }

  {
    int parity = 0;
    for (int c = blockIdx.x; c < a.nCR; c += gridDim.x, parity ^= 1) {
      const int4 it = __ldg(a.item + a.nCC + c);
      const int j = it.x, slot = MODE == FIELD_FUSED ? 0 : a.item_slot[a.nCC + c];
      const Real theta_old = a.theta[j];
      const int g = a.group[j];
      const Real lam = a.lambda[g], mu = a.mu[g], z = a.z[j];
      constexpr int R = sizeof(Real) == 8 ? 4 : 8;
      Real sq, lin;
      Real given = 0;
      if (MODE == FIELD_UPDATE) { // one lane per warp reads the (possibly remote) statistics
        if (lane == 0)
          given = draw_given(slot, theta_old, lam, mu, z);
        given = __shfl_sync(FULL_MASK, given, 0);
      }
      const Real theta_new = field_column_regs<Real, IS_V, UNIT, HAS_MID, PEND, MODE, FIELD_WARPS, R, false, COMPACT>(
          a, s_tab, it, threadIdx.x, theta_old, alpha, lam, mu, z, s_part_cta, parity, 0, warp, lane,
          given, sq, lin);
      if (MODE == FIELD_UPDATE)
        __syncthreads(); // every thread has read theta[j] (the other modes synchronise in the reduction)
      if (threadIdx.x == 0)
        field_finish_column<Real, MODE>(a, j, slot, theta_new, sq, lin);
    }
  }

  {
    constexpr int GROUPS = FIELD_WARPS / FIELD_GROUP_WARPS;
    const int grp = warp / FIELD_GROUP_WARPS, wig = warp % FIELD_GROUP_WARPS;
    int parity = 0;
    for (int k = blockIdx.x + gridDim.x * grp; k < a.nG; k += gridDim.x * GROUPS, parity ^= 1) {
      const int item = a.nCC + a.nCR + k;
      const int4 it = __ldg(a.item + item);
      const int j = it.x, slot = MODE == FIELD_FUSED ? 0 : a.item_slot[item];
      const Real theta_old = a.theta[j];
      const int g = a.group[j];
      const Real lam = a.lambda[g], mu = a.mu[g], z = a.z[j];
      Real sq, lin;
      Real given = 0;
      if (MODE == FIELD_UPDATE) {
        if (lane == 0)
          given = draw_given(slot, theta_old, lam, mu, z);
        given = __shfl_sync(FULL_MASK, given, 0);
      }
      const Real theta_new = field_column_dispatch<Real, IS_V, UNIT, HAS_MID, PEND, MODE, FIELD_GROUP_WARPS, COMPACT>(
          a, s_tab, it, wig * 32 + lane, theta_old, alpha, lam, mu, z, s_part_grp, parity, grp, wig,
          lane, given, sq, lin);
      if (MODE == FIELD_UPDATE) // every warp of the group has read theta[j] before it changes
        asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(FIELD_GROUP_WARPS * 32) : "memory");
      if (wig == 0 && lane == 0)
        field_finish_column<Real, MODE>(a, j, slot, theta_new, sq, lin);
    }
  }

  // One warp per column.  The first batch of FIELD_BATCH columns is static, the following ones are
  // handed out through a counter (one atomic per batch: same-address atomics retire at about one
  // per two cycles chip-wide, one per column would bound the kernel).  Lane b of the warp owns the
  // scalars of the batch's b-th column (item, theta, z, group hypers, and in FIELD_UPDATE its draw:
  // one chain of dependent loads per BATCH, not per column).
  const int total_warps = gridDim.x * FIELD_WARPS;
  const int first_w = a.nCC + a.nCR + a.nG;
  const int4 *items_w = a.item + first_w;
  const int FIELD_BATCH = a.batch;
  int kb = (blockIdx.x + gridDim.x * warp) * FIELD_BATCH; // longest columns spread over the SMs
  // STAGED: this warp's buffer (behind the table and the other warps' buffers) and its copy engine
  constexpr bool NEED_IDX = IS_V || PEND != PEND_NONE;
  unsigned char *w_buf = s_raw + ((static_cast<size_t>(a.n_tab) * Tab::BYTES_PER_COLUMN + 15) / 16) * 16 +
                         static_cast<size_t>(warp) * FieldStageSize<Real>::BYTES;
  unsigned long long *w_bar = &s_wbar[warp];
  uint32_t w_phase = 0;
  FieldStage<Real> stage;
  stage.eq = reinterpret_cast<const Pair<Real> *>(w_buf);
  stage.idx = reinterpret_cast<const int *>(w_buf + FieldStageSize<Real>::EQ_BYTES);
  // starts the copy of the rows [lo, hi) of a column into the buffer (lane 0); an empty column copies nothing
  auto stage_issue = [&](int lo, int hi) {
    if (!STAGED || lane != 0 || hi <= lo)
      return;
    constexpr int AE = 16 / static_cast<int>(sizeof(Pair<Real>));
    const int e0 = lo & ~(AE - 1), e1 = (hi + AE - 1) & ~(AE - 1), i0 = lo & ~3, i1 = (hi + 3) & ~3;
    const uint32_t eb = static_cast<uint32_t>(e1 - e0) * sizeof(Pair<Real>), ib = NEED_IDX ? (i1 - i0) * 4u : 0u;
    fence_async_smem(); // the buffer's previous contents have been read (by all lanes: __syncwarp before)
    mbar_expect_tx(w_bar, eb + ib);
    bulk_load(w_buf, a.eq + e0, eb, w_bar);
    if (NEED_IDX)
      bulk_load(w_buf + FieldStageSize<Real>::EQ_BYTES, a.tail_last + i0, ib, w_bar);
  };
  if (STAGED && kb < a.nW) {
    const int4 first = __ldg(items_w + kb);
    stage_issue(first.y, first.z);
  }
  while (kb < a.nW) {
    int kb_next = 0;
    if (lane == 0)
      kb_next = total_warps * FIELD_BATCH + atomicAdd(a.sched, FIELD_BATCH);
    const int n_batch = min(FIELD_BATCH, a.nW - kb);
    int4 my_it = make_int4(0, 0, 0, 0);
    int my_slot = 0;
    Real my_theta = 0, my_lam = 0, my_mu = 0, my_z = 0, my_given = 0;
    if (lane < n_batch) {
      my_it = __ldg(items_w + kb + lane);
      my_theta = a.theta[my_it.x];
      my_z = a.z[my_it.x];
      const int g = a.group[my_it.x];
      my_lam = a.lambda[g], my_mu = a.mu[g];
      if (MODE != FIELD_FUSED)
        my_slot = a.item_slot[first_w + kb + lane];
      if (MODE == FIELD_UPDATE)
        my_given = draw_given(my_slot, my_theta, my_lam, my_mu, my_z);
    }
    kb_next = __shfl_sync(FULL_MASK, kb_next, 0);
    int4 next_first = make_int4(0, 0, 0, 0); // first column of this warp's next batch
    if (STAGED && kb_next < a.nW)
      next_first = __ldg(items_w + kb_next);
    for (int bi = 0; bi < n_batch; bi++) {
      int4 it;
      it.x = __shfl_sync(FULL_MASK, my_it.x, bi), it.y = __shfl_sync(FULL_MASK, my_it.y, bi);
      it.z = __shfl_sync(FULL_MASK, my_it.z, bi), it.w = 0;
      const Real theta_old = __shfl_sync(FULL_MASK, my_theta, bi), lam = __shfl_sync(FULL_MASK, my_lam, bi);
      const Real mu = __shfl_sync(FULL_MASK, my_mu, bi), z = __shfl_sync(FULL_MASK, my_z, bi);
      const Real given = MODE == FIELD_UPDATE ? __shfl_sync(FULL_MASK, my_given, bi) : Real(0);
      const int slot = MODE == FIELD_FUSED ? 0 : __shfl_sync(FULL_MASK, my_slot, bi);
      Real sq, lin;
      Real theta_new;
      if (STAGED) {
        // the column that follows this one in the warp's work (next in the batch, or first of the next batch)
        const int nb = min(bi + 1, 31);
        int n_lo = __shfl_sync(FULL_MASK, my_it.y, nb), n_hi = __shfl_sync(FULL_MASK, my_it.z, nb);
        if (bi + 1 >= n_batch)
          n_lo = next_first.y, n_hi = next_first.z;
        constexpr int AE = 16 / static_cast<int>(sizeof(Pair<Real>));
        stage.eq_row0 = it.y & ~(AE - 1), stage.idx_row0 = it.y & ~3;
        if (it.z > it.y) { // (an empty column was not copied)
          mbar_wait(w_bar, w_phase);
          w_phase ^= 1;
        }
        auto hook = [&]() {
          __syncwarp();
          stage_issue(n_lo, n_hi);
        };
        theta_new = field_column_dispatch<Real, IS_V, UNIT, HAS_MID, PEND, MODE, 1, COMPACT, true>(
            a, s_tab, it, lane, theta_old, alpha, lam, mu, z, nullptr, 0, 0, 0, lane, given, sq, lin, stage, hook);
      } else {
        theta_new = field_column_dispatch<Real, IS_V, UNIT, HAS_MID, PEND, MODE, 1, COMPACT>(
            a, s_tab, it, lane, theta_old, alpha, lam, mu, z, nullptr, 0, 0, 0, lane, given, sq, lin);
      }
      if (lane == 0)
        field_finish_column<Real, MODE>(a, it.x, slot, theta_new, sq, lin);
    }
    kb = kb_next;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && MODE == FIELD_FUSED)
    peer_trace(a.peer, 6);
  if (MODE == FIELD_STATS)
    peer_post_when_last(a.peer);
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
  const int *item_slot; // row shards: column slot of every item; colstat [2 * columns of the level]
  Real *colstat;        // (nullptr on one GPU: the kernel draws itself)
  PeerView<Real> peer;  // row shards with the peer-memory exchange: statistics go to peer_local
  Real *peer_local;
  Real *partial;   // [2 nS] chunk statistics of the long columns
  int *chunk_done; // [nS] chunks finished, per long column (at its first chunk), zero between launches
  int last_base;
  Real *pend_told, *pend_tnew;
};

template <typename Real, bool IS_V>
__device__ __forceinline__ void field_publish_draw(const FieldStatsArgs<Real> &a, int j, Real sq, Real lin,
                                                   Real theta_old, Real alpha) {
  const int g = a.group[j];
  const Real theta_new = column_draw<Real, IS_V>(sq, lin, theta_old, alpha, a.lambda[g], a.mu[g], a.z[j]);
  a.theta[j] = theta_new;
  if (a.theta_t)
    a.theta_t[static_cast<int64_t>(j) * a.t_stride] = theta_new;
  a.pend_told[j - a.last_base] = theta_old;
  a.pend_tnew[j - a.last_base] = theta_new;
}
// One GPU: draw now.  Row shards: leave the local statistics for the all-reduce (k_field_draw_last draws).
template <typename Real, bool IS_V>
__device__ __forceinline__ void field_publish(const FieldStatsArgs<Real> &a, int item, int j, Real sq, Real lin,
                                              Real theta_old, Real alpha) {
  if (a.colstat || a.peer.world) {
    const int slot = a.item_slot[item];
    Real *out = a.peer.world ? a.peer.produce(a.peer_local) : a.colstat;
    out[2 * slot] = sq, out[2 * slot + 1] = lin;
  } else {
    field_publish_draw<Real, IS_V>(a, j, sq, lin, theta_old, alpha);
  }
}

// Row shards, after the all-reduce: one thread per column of the last level draws from the summed
// statistics (identically on every rank).  cols: the level's columns in slot order.
template <typename Real, bool IS_V>
__global__ void __launch_bounds__(256)
    k_field_draw_last(FieldStatsArgs<Real> a, PeerView<Real> peer, const int *__restrict__ cols, int n_cols) {
  const bool tracer = blockIdx.x == 0 && threadIdx.x == 0;
  if (tracer)
    peer_trace(peer, 2);
  peer_wait(peer);
  if (tracer)
    peer_trace(peer, 3);
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot < n_cols) {
    const int j = cols[slot];
    Real sq, lin;
    if (peer.world)
      peer_sum(peer, slot, sq, lin);
    else
      sq = a.colstat[2 * slot], lin = a.colstat[2 * slot + 1];
    field_publish_draw<Real, IS_V>(a, j, sq, lin, a.theta[j], *a.alpha);
  }
  if (tracer)
    peer_trace(peer, 4);
}

template <typename Real, bool IS_V, bool UNIT>
__global__ void __launch_bounds__(STATS_THREADS) k_field_stats(FieldStatsArgs<Real> a) {
  __shared__ Real scratch[32];
  const int b = blockIdx.x;
  if (b == 0 && threadIdx.x == 0)
    peer_trace(a.peer, 0, 1);
  const bool cta_item = b < a.nS + a.nC;
  const int lane = threadIdx.x & 31;
  const int w = (b - a.nS - a.nC) * (STATS_THREADS / 32) + (threadIdx.x >> 5);
  const bool idle = !cta_item && w >= a.nW; // a warp beyond the last column (stays for the final barrier)
  const int4 it = idle ? make_int4(0, 0, 0, 0) : __ldg(a.item + (cta_item ? b : a.nS + a.nC + w));
  const int j = it.x;
  const Real alpha = *a.alpha;
  const Real theta_old = idle ? Real(0) : a.theta[j];
  const int t = cta_item ? threadIdx.x : lane, nt = cta_item ? STATS_THREADS : 32;
  Real sq = 0, lin = 0;
  constexpr int U = 8; // gathers in flight per thread
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
        // chunk of a long column: the chunk that finishes last adds the chunk statistics in chunk
        // order (deterministic) and draws; it leaves the counter at zero for the next launch
        __stcg(a.partial + 2 * b, sq), __stcg(a.partial + 2 * b + 1, lin);
        __threadfence();
        const int n_chunks = a.seg_count[b];
        if (atomicAdd(a.chunk_done + it.w, 1) == n_chunks - 1) {
          __threadfence();
          Real s = 0, l = 0;
          for (int i = it.w; i < it.w + n_chunks; i++)
            s += __ldcg(a.partial + 2 * i), l += __ldcg(a.partial + 2 * i + 1);
          a.chunk_done[it.w] = 0;
          field_publish<Real, IS_V>(a, it.w, j, s, l, theta_old, alpha);
        }
      } else {
        field_publish<Real, IS_V>(a, b, j, sq, lin, theta_old, alpha);
      }
    }
  } else if (!idle) {
    sq = warp_sum(sq);
    lin = warp_sum(lin);
    if (lane == 0)
      field_publish<Real, IS_V>(a, a.nS + a.nC + w, j, sq, lin, theta_old, alpha);
  }
  peer_post_when_last(a.peer);
}

} // namespace myfm
