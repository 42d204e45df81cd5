// Tile path: column sweeps of a two-field table (user x item one-hot / categorical pairs: every
// MovieLens-shaped workload, BASELINE.json's headline configuration) with the residual / factor
// caches of a ROW TILE staged in shared memory by TMA bulk copies.
//
// What bounds the field path (field_sweep.cuh) on B200 is not HBM but the rate at which an SM can
// issue divergent global accesses: the gather of the last field touches one 32-byte sector per row
// (~1.6 cycles per sector and SM: 54 us per 10 M rows whatever the cache level), and the streaming
// level waits on its row loads with nothing else to do.  Here the rows are cut into tiles of
// whole first-field columns (a contiguous row range each, ~22 k rows in f32) whose {e, q} pairs fit
// the shared memory of one SM next to the last field's factor table.  One CTA per tile:
//
//   0. cp.async.bulk (TMA, 1-D) of the tile's {e, q} pairs global -> shared, mbarrier-signalled;
//      the last field's table theta[last field] is filled meanwhile;
//   1. "B pass": the tile's rows in the order of their last-field column (a per-tile permutation
//      stored as one 32-bit word {row, column} per row, so neighbouring lanes see neighbouring
//      columns and every table access coalesces): applies the rank-1 update the last field of the
//      PREVIOUS vector left pending (FMTrainer.hpp:240-251, :366-374) and leaves the last field's
//      share x . theta of q_init (FMTrainer.hpp:320) in the tile;
//   2. the first-field sweep (FMTrainer.hpp:237-254, :343-376): a warp per column, statistics ->
//      draw -> update, on shared memory only (no global access per row when all values are 1);
//   3. the tile's pairs go back with one bulk store (shared -> global) WHILE the last field's
//      statistics are formed: the rows once more, grouped by run (= the rows of one last-field
//      column inside the tile) and laid out by padded run length, so that a warp iteration reduces
//      32 / c aligned runs of c slots with log2(c) butterfly steps (no keys, no carries; long runs
//      take a warp each) — every (column, tile) sum has a fixed order.  k_tile_fold adds a column's
//      tile sums in a fixed order, draws, and leaves {theta_old, theta_new} pending.
//
// Every random access of the sweep is a shared-memory access; global memory sees only streams:
// per row and vector 16 B of {e, q} (one read, one write) and 2 x 4 B of B order.  Element-wise
// arithmetic is that of field_sweep.cuh / kernels.cuh (and of the reference).
#pragma once

#include "field_sweep.cuh"

namespace myfm {

constexpr int TILE_THREADS = 1024;
constexpr int TILE_WARPS = TILE_THREADS / 32;
constexpr int TILE_CTA_MIN = 4096;          // first-field columns longer than this: the whole CTA
constexpr int TILE_MAX_ROWS = 65535;        // local row positions are 16 bit
constexpr int TILE_MAX_TAB = 65535;         // ... and so are last-field columns (relative to last_base)
constexpr unsigned TILE_NO_KEY = 0xffffffffu;
constexpr uint32_t TILE_BULK_CHUNK = 32768; // bytes per bulk copy instruction

constexpr int TILE_VEC = 128;      // slots per warp iteration of the B passes: one 16-byte load per lane
constexpr int TILE_TEAM = 8;       // lanes per first-field column of up to TILE_TEAM_MAX rows (4 columns per warp)
constexpr int TILE_TEAM_MAX = 256;

template <typename Real> struct TileArgs {
  // tiles (static, built once per fit: engine.cu setup_tile_path)
  const int *tile_row;        // [n_tiles + 1] first row of every tile
  const int *tile_item_ptr;   // [n_tiles + 1] into item
  const int *tile_n_cta;      // [n_tiles] leading items of the tile handled by the whole CTA (> TILE_CTA_MIN rows)
  const int *tile_n_warp;     // [n_tiles] following items handled by a warp each (> TILE_TEAM_MAX rows)
  const int4 *item;           // first-field columns {column, first row, end row, group}, longest first per tile
  // B order: the tile's rows sorted by (last-field column, row), one word per row:
  // (row - tile_row[t]) << 16 | (column - last_base); every tile's range starts at a multiple of
  // TILE_VEC words and is padded to one with TILE_NO_KEY.
  const unsigned *b_ent;
  const Real *b_val;          // last-field value of every B entry (unused when UNIT)
  const int *tile_b_ptr;      // [n_tiles + 1] word offsets of the tiles in b_ent
  // The same rows once more for the statistics of the last field, grouped by RUN (the rows of one
  // last-field column inside the tile, ascending): runs of up to 16 rows are padded to 1, 2, 4, 8 or
  // 16 slots (TILE_NO_KEY) and stored class by class, every class a multiple of TILE_VEC slots.  A
  // lane loads 4 consecutive slots, so runs of up to 4 rows are summed inside a lane and longer ones
  // with one or two butterfly steps; runs above 16 rows are stored behind the classes (each padded
  // to a multiple of 4 slots) and take a warp each.
  const unsigned *c_ent;
  const Real *c_val;          // last-field value per slot (unused when UNIT)
  const int *tile_cls_ptr;    // [n_tiles][8]: slot offsets of the classes 1, 2, 4, 8, 16 ([k] .. [k+1]), k = 0 .. 4;
                              // [5] .. [6]: the long runs' slots
  const int *tile_long_ptr;   // [n_tiles + 1] into long_run
  const int4 *long_run;       // {column - last_base, first slot, slots (multiple of 4), -}, longest first per tile
  int n_tiles;
  uint32_t eq_bytes;          // shared-memory bytes reserved for the pairs (the table follows)
  // rows
  Pair<Real> *eq;
  const Real *own_val;        // [n_rows] first-field value (unused when UNIT)
  // the vector being swept
  Real *theta, *theta_t;
  int64_t t_stride;
  const Real *z;
  const int *group;
  const Real *alpha, *lambda, *mu;
  int last_base, n_tab;
  const Pair<Real> *pend; // [n_tab] {theta_old, theta_new} left pending by the previous vector
  Pair<Real> *part;       // [n_tiles][n_tab] last-field statistics of every tile
};

// L2 prefetch of a whole range with one TMA instruction (16-byte granularity).
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// Sum of two values over the CTA, broadcast (one barrier pair for both).
template <typename Real> __device__ __forceinline__ void tile_block_sum2(Real &a, Real &b, Real *scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  a = warp_sum(a), b = warp_sum(b);
  __syncthreads();
  if (lane == 0)
    scratch[2 * wid] = a, scratch[2 * wid + 1] = b;
  __syncthreads();
  Real ta = scratch[2 * lane], tb = scratch[2 * lane + 1]; // TILE_WARPS == 32
  a = warp_sum(ta), b = warp_sum(tb);
}

// Statistics pass of a first-field column over the rows lo + t, lo + t + nt, ... of the tile.
// IS_V: q_i = x0 theta_old + (x_last theta_last, left by the B pass): q_init in CSR order.
template <typename Real, bool IS_V, bool UNIT>
__device__ __forceinline__ void tile_column_stats(const TileArgs<Real> &a, Pair<Real> *s_eq, int lo, int hi, int t,
                                                  int nt, Real theta_old, Real alpha, Real &sq, Real &lin) {
#pragma unroll 4
  for (int i = lo + t; i < hi; i += nt) {
    const Real x0 = UNIT ? Real(1) : __ldg(a.own_val + i);
    Pair<Real> v = s_eq[i];
    if (IS_V) {
      Real acc = x0 * theta_old;
      acc += v.y;
      v.y = acc;
      s_eq[i].y = acc;
    }
    field_stats<Real, IS_V>(v.x, v.y, x0, theta_old, alpha, sq, lin);
  }
}

template <typename Real, bool IS_V, bool UNIT>
__device__ __forceinline__ void tile_column_update(const TileArgs<Real> &a, Pair<Real> *s_eq, int lo, int hi, int t,
                                                   int nt, Real theta_old, Real theta_new) {
#pragma unroll 4
  for (int i = lo + t; i < hi; i += nt) {
    const Real x0 = UNIT ? Real(1) : __ldg(a.own_val + i);
    const Pair<Real> v = s_eq[i];
    s_eq[i] = field_update<Real, IS_V>(v.x, v.y, x0, theta_old, theta_new);
  }
}

template <typename Real> __device__ __forceinline__ void tile_store_run(Pair<Real> *part, unsigned ent, Real sa, Real sb) {
  Pair<Real> out;
  out.x = sa, out.y = sb;
  __stcg(part + (ent & 0xffffu), out);
}

// Statistics of the runs of class C (slots per run) in the warp iterations [it_lo, it_hi) of TILE_VEC
// slots: lane l owns the slots 4 l .. 4 l + 3 of an iteration.
template <typename Real, bool IS_V, bool UNIT, int C>
__device__ __forceinline__ void tile_class_pass(const TileArgs<Real> &a, const Pair<Real> *s_eq_tile, const Real *s_tab,
                                                Pair<Real> *part, int it_lo, int it_hi, int lane, Real alpha) {
  const uint4 *src = reinterpret_cast<const uint4 *>(a.c_ent) + lane;
  uint4 nxt = __ldcs(src + static_cast<size_t>(it_lo) * (TILE_VEC / 4));
  for (int it = it_lo; it < it_hi; it++) {
    const uint4 e4 = nxt;
    if (it + 1 < it_hi) // the next iteration's slots are on their way while this one is reduced
      nxt = __ldcs(src + static_cast<size_t>(it + 1) * (TILE_VEC / 4));
    const size_t at = static_cast<size_t>(it) * TILE_VEC + 4 * lane;
    const unsigned ent[4] = {e4.x, e4.y, e4.z, e4.w};
    Real sa[4], sb[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      sa[k] = 0, sb[k] = 0;
      if (ent[k] != TILE_NO_KEY) {
        const Real x = UNIT ? Real(1) : __ldcs(a.c_val + at + k);
        const Pair<Real> v = s_eq_tile[ent[k] >> 16];
        field_stats<Real, IS_V>(v.x, v.y, x, s_tab[ent[k] & 0xffffu], alpha, sa[k], sb[k]);
      }
    }
    if (C == 1) {
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (ent[k] != TILE_NO_KEY)
          tile_store_run(part, ent[k], sa[k], sb[k]);
    } else if (C == 2) {
      if (ent[0] != TILE_NO_KEY)
        tile_store_run(part, ent[0], sa[0] + sa[1], sb[0] + sb[1]);
      if (ent[2] != TILE_NO_KEY)
        tile_store_run(part, ent[2], sa[2] + sa[3], sb[2] + sb[3]);
    } else {
      Real ta = (sa[0] + sa[1]) + (sa[2] + sa[3]), tb = (sb[0] + sb[1]) + (sb[2] + sb[3]);
#pragma unroll
      for (int o = 1; o < C / 4; o <<= 1) {
        ta += __shfl_xor_sync(FULL_MASK, ta, o);
        tb += __shfl_xor_sync(FULL_MASK, tb, o);
      }
      if ((lane & (C / 4 - 1)) == 0 && ent[0] != TILE_NO_KEY)
        tile_store_run(part, ent[0], ta, tb);
    }
  }
}

// One warp iteration of the first B pass: pending update of the last field and its share of q_init
// for the four rows of this lane.
template <typename Real, bool IS_V, bool UNIT, int PEND>
__device__ __forceinline__ void tile_pending_pass(const TileArgs<Real> &a, Pair<Real> *s_eq_tile, const Real *s_tab,
                                                  uint4 e4, size_t at) {
  const unsigned ent[4] = {e4.x, e4.y, e4.z, e4.w};
  Pair<Real> tt[4];
  Real xv[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    xv[k] = Real(1);
    tt[k].x = 0, tt[k].y = 0;
    if (ent[k] != TILE_NO_KEY) {
      if (PEND != PEND_NONE)
        tt[k] = __ldg(a.pend + (ent[k] & 0xffffu));
      if (!UNIT)
        xv[k] = __ldcs(a.b_val + at + k);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (ent[k] != TILE_NO_KEY) {
      const Real x = xv[k];
      Pair<Real> v = s_eq_tile[ent[k] >> 16];
      if (PEND == PEND_V) { // FMTrainer.hpp:366-374 of the previous factor's last-field column
        const Real h = x * (v.y - x * tt[k].x);
        v.x = v.x + h * (tt[k].y - tt[k].x);
      } else if (PEND == PEND_W) { // FMTrainer.hpp:240,251
        v.x = (v.x - x * tt[k].x) + x * tt[k].y;
      }
      if (IS_V)
        v.y = x * s_tab[ent[k] & 0xffffu];
      s_eq_tile[ent[k] >> 16] = v;
    }
}

constexpr int TILE_P1_PRE = 8; // warp iterations of the first B pass loaded before the tile has arrived

template <typename Real, bool IS_V, bool UNIT, int PEND>
__global__ void __launch_bounds__(TILE_THREADS, 1) k_tile_sweep(const __grid_constant__ TileArgs<Real> a) {
  extern __shared__ __align__(128) unsigned char tile_smem[];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ int s_counter[2];
  __shared__ Real s_scratch[2 * TILE_WARPS];
  constexpr int AR = 16 / static_cast<int>(sizeof(Pair<Real>)); // rows per 16 bytes (bulk copy granularity)
  constexpr bool HAS_P1 = IS_V || PEND != PEND_NONE;
  const int tile = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row0 = a.tile_row[tile], row1 = a.tile_row[tile + 1];
  const int first = row0 & ~(AR - 1), last = (row1 + AR - 1) & ~(AR - 1); // the buffer has a spare pair at the end
  Pair<Real> *s_eq = reinterpret_cast<Pair<Real> *>(tile_smem) - first;  // s_eq[i]: global row i
  Real *s_tab = reinterpret_cast<Real *>(tile_smem + a.eq_bytes);
  const uint32_t bytes = static_cast<uint32_t>(last - first) * sizeof(Pair<Real>);
  const int b0 = a.tile_b_ptr[tile], b1 = a.tile_b_ptr[tile + 1];
  const int *cp = a.tile_cls_ptr + tile * 8;

  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    s_counter[0] = s_counter[1] = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (bytes) {
      mbar_expect_tx(&s_bar, bytes);
      const unsigned char *src = reinterpret_cast<const unsigned char *>(a.eq + first);
      for (uint32_t off = 0; off < bytes; off += TILE_BULK_CHUNK)
        bulk_load(tile_smem + off, src + off, min(TILE_BULK_CHUNK, bytes - off), &s_bar);
    }
    // the index streams of the B passes on their way into L2 meanwhile
    if (HAS_P1 && b1 > b0) {
      bulk_prefetch_l2(a.b_ent + b0, static_cast<uint32_t>(b1 - b0) * 4u);
      if (!UNIT)
        bulk_prefetch_l2(a.b_val + b0, static_cast<uint32_t>(b1 - b0) * static_cast<uint32_t>(sizeof(Real)));
    }
    if (cp[6] > cp[0]) {
      bulk_prefetch_l2(a.c_ent + cp[0], static_cast<uint32_t>(cp[6] - cp[0]) * 4u);
      if (!UNIT)
        bulk_prefetch_l2(a.c_val + cp[0], static_cast<uint32_t>(cp[6] - cp[0]) * static_cast<uint32_t>(sizeof(Real)));
    }
  }
  // While the rows are in flight: this warp's share of the first B pass (a contiguous range of warp
  // iterations) goes into registers, and the last field's values of this vector into the table.
  const int p1_n = (b1 - b0) / TILE_VEC, p1_per = (p1_n + TILE_WARPS - 1) / TILE_WARPS;
  const int p1_lo = min(p1_n, warp * p1_per), p1_hi = min(p1_n, p1_lo + p1_per);
  const uint4 *p1_src = reinterpret_cast<const uint4 *>(a.b_ent + b0) + lane;
  uint4 p1_pre[TILE_P1_PRE];
  if (HAS_P1) {
#pragma unroll
    for (int k = 0; k < TILE_P1_PRE; k++)
      if (p1_lo + k < p1_hi)
        p1_pre[k] = __ldcs(p1_src + static_cast<size_t>(p1_lo + k) * (TILE_VEC / 4));
  }
  for (int t0 = threadIdx.x; t0 < a.n_tab; t0 += 4 * TILE_THREADS) {
    Real tv[4];
#pragma unroll
    for (int u = 0; u < 4; u++)
      tv[u] = t0 + u * TILE_THREADS < a.n_tab ? a.theta[a.last_base + t0 + u * TILE_THREADS] : Real(0);
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (t0 + u * TILE_THREADS < a.n_tab)
        s_tab[t0 + u * TILE_THREADS] = tv[u];
  }
  const Real alpha = *a.alpha;
  if (bytes)
    mbar_wait(&s_bar, 0);
  __syncthreads();

  // ---- 1. B pass: pending update of the last field; its share of q_init -------------------------
  if (HAS_P1) {
    Pair<Real> *s_eq_tile = s_eq + row0;
#pragma unroll
    for (int k = 0; k < TILE_P1_PRE; k++)
      if (p1_lo + k < p1_hi)
        tile_pending_pass<Real, IS_V, UNIT, PEND>(a, s_eq_tile, s_tab, p1_pre[k],
                                                  static_cast<size_t>(b0) + static_cast<size_t>(p1_lo + k) * TILE_VEC +
                                                      4 * lane);
    for (int it = p1_lo + TILE_P1_PRE; it < p1_hi; it++) // tiles beyond TILE_P1_PRE * 4096 rows
      tile_pending_pass<Real, IS_V, UNIT, PEND>(a, s_eq_tile, s_tab,
                                                __ldcs(p1_src + static_cast<size_t>(it) * (TILE_VEC / 4)),
                                                static_cast<size_t>(b0) + static_cast<size_t>(it) * TILE_VEC + 4 * lane);
    __syncthreads();
  }

  // ---- 2. first field: statistics, draw, update on shared memory --------------------------------
  {
    const int it0 = a.tile_item_ptr[tile], it1 = a.tile_item_ptr[tile + 1];
    const int n_cta = a.tile_n_cta[tile], n_warp = a.tile_n_warp[tile];
    for (int c = it0; c < it0 + n_cta; c++) { // very long columns: the whole CTA
      const int4 it = __ldg(a.item + c);
      const int j = it.x, g = a.group[j];
      const Real theta_old = a.theta[j];
      Real sq = 0, lin = 0;
      tile_column_stats<Real, IS_V, UNIT>(a, s_eq, it.y, it.z, threadIdx.x, TILE_THREADS, theta_old, alpha, sq, lin);
      tile_block_sum2(sq, lin, s_scratch);
      const Real theta_new = column_draw<Real, IS_V>(sq, lin, theta_old, alpha, a.lambda[g], a.mu[g], a.z[j]);
      tile_column_update<Real, IS_V, UNIT>(a, s_eq, it.y, it.z, threadIdx.x, TILE_THREADS, theta_old, theta_new);
      __syncthreads(); // every thread has read theta[j]
      if (threadIdx.x == 0) {
        a.theta[j] = theta_new;
        if (a.theta_t)
          a.theta_t[static_cast<int64_t>(j) * a.t_stride] = theta_new;
      }
    }
    // long columns: a warp each, handed out one at a time (longest first)
    for (;;) {
      int c = 0;
      if (lane == 0)
        c = atomicAdd(&s_counter[0], 1);
      c = __shfl_sync(FULL_MASK, c, 0);
      if (c >= n_warp)
        break;
      const int4 it = __ldg(a.item + it0 + n_cta + c);
      const int j = it.x, g = a.group[j];
      const Real theta_old = a.theta[j];
      const Real lam = a.lambda[g], mu = a.mu[g], z = a.z[j];
      Real sq = 0, lin = 0;
      tile_column_stats<Real, IS_V, UNIT>(a, s_eq, it.y, it.z, lane, 32, theta_old, alpha, sq, lin);
      sq = warp_sum(sq), lin = warp_sum(lin);
      const Real theta_new = column_draw<Real, IS_V>(sq, lin, theta_old, alpha, lam, mu, z);
      tile_column_update<Real, IS_V, UNIT>(a, s_eq, it.y, it.z, lane, 32, theta_old, theta_new);
      if (lane == 0) {
        a.theta[j] = theta_new;
        if (a.theta_t)
          a.theta_t[static_cast<int64_t>(j) * a.t_stride] = theta_new;
      }
    }
    // the other columns (up to TILE_TEAM_MAX rows): TILE_TEAM lanes each, four columns per warp at a time,
    // so four chains of statistics -> draw -> update are in flight per warp
    const int team0 = it0 + n_cta + n_warp, n_team = it1 - team0;
    const int team = lane / TILE_TEAM, tl = lane % TILE_TEAM;
    constexpr int TEAMS = 32 / TILE_TEAM;
    // (the item of the next round is fetched while this round's columns are swept)
    int c0 = 0;
    if (lane == 0)
      c0 = atomicAdd(&s_counter[1], TEAMS);
    c0 = __shfl_sync(FULL_MASK, c0, 0);
    int4 it = make_int4(0, 0, 0, 0);
    if (c0 + team < n_team)
      it = __ldg(a.item + team0 + c0 + team);
    while (c0 < n_team) {
      const bool live = c0 + team < n_team;
      Real theta_old = 0, lam = 0, mu = 0, z = 0;
      if (live) { // the lanes of a team read the same words; it.w = the column's group
        theta_old = a.theta[it.x];
        z = a.z[it.x];
        lam = a.lambda[it.w], mu = a.mu[it.w];
      }
      int c_next = 0;
      if (lane == 0)
        c_next = atomicAdd(&s_counter[1], TEAMS);
      c_next = __shfl_sync(FULL_MASK, c_next, 0);
      int4 it_next = make_int4(0, 0, 0, 0);
      if (c_next + team < n_team)
        it_next = __ldg(a.item + team0 + c_next + team);
      Real sq = 0, lin = 0;
      tile_column_stats<Real, IS_V, UNIT>(a, s_eq, it.y, it.z, tl, TILE_TEAM, theta_old, alpha, sq, lin);
#pragma unroll
      for (int o = TILE_TEAM / 2; o > 0; o >>= 1) {
        sq += __shfl_xor_sync(FULL_MASK, sq, o);
        lin += __shfl_xor_sync(FULL_MASK, lin, o);
      }
      const Real theta_new = column_draw<Real, IS_V>(sq, lin, theta_old, alpha, lam, mu, z);
      tile_column_update<Real, IS_V, UNIT>(a, s_eq, it.y, it.z, tl, TILE_TEAM, theta_old, theta_new);
      if (live && tl == 0) {
        a.theta[it.x] = theta_new;
        if (a.theta_t)
          a.theta_t[static_cast<int64_t>(it.x) * a.t_stride] = theta_new;
      }
      c0 = c_next, it = it_next;
    }
  }
  fence_async_smem();
  __syncthreads();

  // ---- 3. the pairs go back while the last field's statistics are reduced -----------------------
  if (threadIdx.x == 0) {
    const int in0 = (row0 + AR - 1) & ~(AR - 1), in1 = row1 & ~(AR - 1); // 16-byte aligned interior
    if (in1 > in0) {
      const uint32_t sbytes = static_cast<uint32_t>(in1 - in0) * sizeof(Pair<Real>);
      unsigned char *dst = reinterpret_cast<unsigned char *>(a.eq + in0);
      const unsigned char *src = reinterpret_cast<const unsigned char *>(s_eq + in0);
      for (uint32_t off = 0; off < sbytes; off += TILE_BULK_CHUNK)
        bulk_store(dst + off, src + off, min(TILE_BULK_CHUNK, sbytes - off));
      bulk_commit();
    }
    // rows outside the aligned interior (a tile that starts or ends on an odd row): ordinary stores
    const int head_end = in1 > in0 ? in0 : row1;
    for (int i = row0; i < head_end; i++)
      a.eq[i] = s_eq[i];
    if (in1 > in0)
      for (int i = in1; i < row1; i++)
        a.eq[i] = s_eq[i];
  }
  {
    Pair<Real> *part = a.part + static_cast<size_t>(tile) * a.n_tab;
    const Pair<Real> *s_eq_tile = s_eq + row0;
    // short runs: every warp takes a contiguous range of the warp iterations of all five classes
    const int i_begin = cp[0] / TILE_VEC, i_end = cp[5] / TILE_VEC;
    const int per = (i_end - i_begin + TILE_WARPS - 1) / TILE_WARPS;
    const int my_lo = min(i_end, i_begin + warp * per), my_hi = min(i_end, my_lo + per);
#define MYFM_CLASS(K, C)                                                                           \
  {                                                                                                \
    const int lo = max(my_lo, cp[K] / TILE_VEC), hi = min(my_hi, cp[K + 1] / TILE_VEC);             \
    if (lo < hi)                                                                                   \
      tile_class_pass<Real, IS_V, UNIT, C>(a, s_eq_tile, s_tab, part, lo, hi, lane, alpha);        \
  }
    MYFM_CLASS(0, 1)
    MYFM_CLASS(1, 2)
    MYFM_CLASS(2, 4)
    MYFM_CLASS(3, 8)
    MYFM_CLASS(4, 16)
#undef MYFM_CLASS
    // long runs: a warp each
    for (int r = a.tile_long_ptr[tile] + warp; r < a.tile_long_ptr[tile + 1]; r += TILE_WARPS) {
      const int4 run = __ldg(a.long_run + r);
      const Real theta_col = s_tab[run.x];
      Real sa = 0, sb = 0;
#pragma unroll 2
      for (int p = run.y + 4 * lane; p < run.y + run.z; p += TILE_VEC) {
        const uint4 e4 = __ldcs(reinterpret_cast<const uint4 *>(a.c_ent + p));
        const unsigned ent[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (ent[k] != TILE_NO_KEY) {
            const Real x = UNIT ? Real(1) : __ldcs(a.c_val + p + k);
            const Pair<Real> v = s_eq_tile[ent[k] >> 16];
            field_stats<Real, IS_V>(v.x, v.y, x, theta_col, alpha, sa, sb);
          }
      }
      sa = warp_sum(sa), sb = warp_sum(sb);
      if (lane == 0) {
        Pair<Real> out;
        out.x = sa, out.y = sb;
        __stcg(part + run.x, out);
      }
    }
  }
  if (threadIdx.x == 0)
    bulk_wait_all(); // the bulk store has read the tile (and is complete) before the CTA retires
}

// A last-field column's statistics = its tiles' partial sums added in a fixed order; then the draw
// (FMTrainer.hpp:244-250, :359-367), stored with theta_old for the next vector's pending update.
template <typename Real> struct TileFoldArgs {
  const int *cols; // the last field's columns
  int n_cols, n_tiles, last_base, n_tab;
  const Pair<Real> *part; // [n_tiles][n_tab]
  Real *theta, *theta_t;
  int64_t t_stride;
  const Real *z;
  const int *group;
  const Real *alpha, *lambda, *mu;
  Pair<Real> *pend;
  // row shards: this rank's sums go to the exchange buffer instead (k_field_draw_last draws)
  int to_peer;
  PeerView<Real> peer;
  Real *peer_local, *colstat;
};

// 512 threads = 32 columns x 16 tile groups: warp g adds the tiles g, g + 16, ... of 32 neighbouring
// columns (coalesced rows of `part`), warp 0 joins the groups in group order and draws.
constexpr int FOLD_GROUPS = 16;
template <typename Real, bool IS_V>
__global__ void __launch_bounds__(32 * FOLD_GROUPS) k_tile_fold(TileFoldArgs<Real> a) {
  __shared__ Real s_sum[FOLD_GROUPS][2][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int w = blockIdx.x * 32 + lane;
  const bool live = w < a.n_cols;
  const int j = live ? a.cols[w] : 0, c = j - a.last_base;
  Real sq = 0, lin = 0;
  if (live) {
#pragma unroll 8
    for (int t = grp; t < a.n_tiles; t += FOLD_GROUPS) {
      const Pair<Real> v = __ldcg(a.part + static_cast<size_t>(t) * a.n_tab + c);
      sq += v.x, lin += v.y;
    }
  }
  s_sum[grp][0][lane] = sq, s_sum[grp][1][lane] = lin;
  __syncthreads();
  if (grp == 0 && live) {
    sq = 0, lin = 0;
#pragma unroll
    for (int g = 0; g < FOLD_GROUPS; g++)
      sq += s_sum[g][0][lane], lin += s_sum[g][1][lane];
    if (a.to_peer) {
      Real *out = a.peer.world ? a.peer.produce(a.peer_local) : a.colstat;
      out[2 * w] = sq, out[2 * w + 1] = lin;
    } else {
      const int g = a.group[j];
      const Real theta_old = a.theta[j];
      const Real theta_new = column_draw<Real, IS_V>(sq, lin, theta_old, *a.alpha, a.lambda[g], a.mu[g], a.z[j]);
      a.theta[j] = theta_new;
      if (a.theta_t)
        a.theta_t[static_cast<int64_t>(j) * a.t_stride] = theta_new;
      Pair<Real> pd;
      pd.x = theta_old, pd.y = theta_new;
      a.pend[c] = pd;
    }
  }
  if (a.to_peer)
    peer_post_when_last(a.peer);
}

} // namespace myfm
