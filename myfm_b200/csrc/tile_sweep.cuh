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
//   3. the tile's pairs go back with one bulk store (shared -> global) WHILE a second B pass forms
//      the last field's statistics: one thread per row, a segmented warp scan over equal columns,
//      runs that continue across iterations carried in registers, runs that continue across warps
//      joined in warp order — every (column, tile) sum has a fixed order.  k_tile_fold adds a
//      column's tile sums in tile order, draws, and leaves {theta_old, theta_new} pending.
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

template <typename Real> struct TileArgs {
  // tiles (static, built once per fit: engine.cu setup_tile_path)
  const int *tile_row;        // [n_tiles + 1] first row of every tile
  const int *tile_item_ptr;   // [n_tiles + 1] into item
  const int *tile_n_cta;      // [n_tiles] leading items of the tile handled by the whole CTA
  const int4 *item;           // first-field columns {column, first row, end row, -}, longest first per tile
  const unsigned *b_ent;      // [n_rows] B order: entries [tile_row[t], tile_row[t+1]) are tile t's rows sorted by
                              // (last-field column, row) as (row - tile_row[t]) << 16 | (column - last_base)
  const Real *b_val;          // [n_rows] last-field value of every B entry (unused when UNIT)
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

// Inclusive sums over runs of equal keys (sorted: a run is contiguous) inside a warp.
template <typename Real>
__device__ __forceinline__ void tile_run_scan(unsigned key, Real &sa, Real &sb, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const Real ua = __shfl_up_sync(FULL_MASK, sa, d), ub = __shfl_up_sync(FULL_MASK, sb, d);
    const unsigned uk = __shfl_up_sync(FULL_MASK, key, d);
    if (lane >= d && uk == key)
      sa += ua, sb += ub;
  }
}

// What a warp leaves for the join of runs across warps: its first and its last run.
template <typename Real> struct TileRunRecord {
  unsigned first_col, last_col; // TILE_NO_KEY: none
  Real first_a, first_b, last_a, last_b;
};

constexpr int TILE_BU = 4; // B entries per thread in flight

template <typename Real, bool IS_V, bool UNIT, int PEND>
__global__ void __launch_bounds__(TILE_THREADS, 1) k_tile_sweep(const __grid_constant__ TileArgs<Real> a) {
  extern __shared__ __align__(128) unsigned char tile_smem[];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ int s_counter;
  __shared__ Real s_scratch[2 * TILE_WARPS];
  __shared__ TileRunRecord<Real> s_rec[TILE_WARPS];
  constexpr int AR = 16 / static_cast<int>(sizeof(Pair<Real>)); // rows per 16 bytes (bulk copy granularity)
  const int tile = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row0 = a.tile_row[tile], row1 = a.tile_row[tile + 1];
  const int first = row0 & ~(AR - 1), last = (row1 + AR - 1) & ~(AR - 1); // the buffer has a spare pair at the end
  Pair<Real> *s_eq = reinterpret_cast<Pair<Real> *>(tile_smem) - first;  // s_eq[i]: global row i
  Real *s_tab = reinterpret_cast<Real *>(tile_smem + a.eq_bytes);
  const uint32_t bytes = static_cast<uint32_t>(last - first) * sizeof(Pair<Real>);

  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    s_counter = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0 && bytes) {
    mbar_expect_tx(&s_bar, bytes);
    const unsigned char *src = reinterpret_cast<const unsigned char *>(a.eq + first);
    for (uint32_t off = 0; off < bytes; off += TILE_BULK_CHUNK)
      bulk_load(tile_smem + off, src + off, min(TILE_BULK_CHUNK, bytes - off), &s_bar);
  }
  // the last field's values of this vector while the rows are in flight
  for (int t = threadIdx.x; t < a.n_tab; t += TILE_THREADS)
    s_tab[t] = a.theta[a.last_base + t];
  const Real alpha = *a.alpha;
  // this warp's share of the B order: a contiguous range, a multiple of 32 entries
  const int per_warp = ((row1 - row0 + TILE_THREADS - 1) / TILE_THREADS) * 32;
  const int b_lo = min(row1, row0 + warp * per_warp), b_hi = min(row1, b_lo + per_warp);
  if (bytes)
    mbar_wait(&s_bar, 0);
  __syncthreads();

  // ---- 1. B pass: pending update of the last field; its share of q_init -------------------------
  if (IS_V || PEND != PEND_NONE) {
    for (int base = b_lo + lane; base < b_hi; base += 32 * TILE_BU) {
      unsigned ent[TILE_BU];
      Real xv[TILE_BU];
      Pair<Real> tt[TILE_BU];
#pragma unroll
      for (int u = 0; u < TILE_BU; u++) {
        const int p = base + 32 * u;
        ent[u] = p < b_hi ? __ldcs(a.b_ent + p) : TILE_NO_KEY;
        xv[u] = (UNIT || p >= b_hi) ? Real(1) : __ldcs(a.b_val + p);
      }
      if (PEND != PEND_NONE) {
#pragma unroll
        for (int u = 0; u < TILE_BU; u++)
          if (ent[u] != TILE_NO_KEY)
            tt[u] = __ldg(a.pend + (ent[u] & 0xffffu));
      }
#pragma unroll
      for (int u = 0; u < TILE_BU; u++)
        if (ent[u] != TILE_NO_KEY) {
          const int i = row0 + static_cast<int>(ent[u] >> 16);
          const Real x = xv[u];
          Pair<Real> v = s_eq[i];
          if (PEND == PEND_V) { // FMTrainer.hpp:366-374 of the previous factor's last-field column
            const Real h = x * (v.y - x * tt[u].x);
            v.x = v.x + h * (tt[u].y - tt[u].x);
          } else if (PEND == PEND_W) { // FMTrainer.hpp:240,251
            v.x = (v.x - x * tt[u].x) + x * tt[u].y;
          }
          if (IS_V)
            v.y = x * s_tab[ent[u] & 0xffffu];
          s_eq[i] = v;
        }
    }
    __syncthreads();
  }

  // ---- 2. first field: statistics, draw, update on shared memory --------------------------------
  {
    const int it0 = a.tile_item_ptr[tile], it1 = a.tile_item_ptr[tile + 1];
    const int n_cta = a.tile_n_cta[tile];
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
    // One warp per column, longest first, handed out in batches; lane b of the warp owns the
    // scalars of the batch's b-th column (one chain of dependent loads per batch, not per column).
    const int n_warp_items = it1 - it0 - n_cta;
    const int batch = max(1, min(32, n_warp_items / (2 * TILE_WARPS)));
    for (;;) {
      int c0 = 0;
      if (lane == 0)
        c0 = atomicAdd(&s_counter, batch);
      c0 = __shfl_sync(FULL_MASK, c0, 0);
      if (c0 >= n_warp_items)
        break;
      const int n_b = min(batch, n_warp_items - c0);
      int4 my_it = make_int4(0, 0, 0, 0);
      Real my_theta = 0, my_lam = 0, my_mu = 0, my_z = 0;
      if (lane < n_b) {
        my_it = __ldg(a.item + it0 + n_cta + c0 + lane);
        my_theta = a.theta[my_it.x];
        my_z = a.z[my_it.x];
        const int g = a.group[my_it.x];
        my_lam = a.lambda[g], my_mu = a.mu[g];
      }
      Real my_new = 0;
      for (int bi = 0; bi < n_b; bi++) {
        const int lo = __shfl_sync(FULL_MASK, my_it.y, bi), hi = __shfl_sync(FULL_MASK, my_it.z, bi);
        const Real theta_old = __shfl_sync(FULL_MASK, my_theta, bi), lam = __shfl_sync(FULL_MASK, my_lam, bi);
        const Real mu = __shfl_sync(FULL_MASK, my_mu, bi), z = __shfl_sync(FULL_MASK, my_z, bi);
        Real sq = 0, lin = 0;
        tile_column_stats<Real, IS_V, UNIT>(a, s_eq, lo, hi, lane, 32, theta_old, alpha, sq, lin);
        sq = warp_sum(sq), lin = warp_sum(lin);
        const Real theta_new = column_draw<Real, IS_V>(sq, lin, theta_old, alpha, lam, mu, z);
        tile_column_update<Real, IS_V, UNIT>(a, s_eq, lo, hi, lane, 32, theta_old, theta_new);
        if (lane == bi)
          my_new = theta_new;
      }
      if (lane < n_b) {
        a.theta[my_it.x] = my_new;
        if (a.theta_t)
          a.theta_t[static_cast<int64_t>(my_it.x) * a.t_stride] = my_new;
      }
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
    // the run that reaches the end of the entries seen so far (all lanes hold the same copy)
    unsigned carry_col = TILE_NO_KEY;
    Real carry_a = 0, carry_b = 0;
    bool first_done = false;
    TileRunRecord<Real> rec;
    rec.first_col = rec.last_col = TILE_NO_KEY;
    rec.first_a = rec.first_b = rec.last_a = rec.last_b = 0;
    auto store_run = [&](unsigned col, Real sa, Real sb) {
      Pair<Real> out;
      out.x = sa, out.y = sb;
      __stcg(part + col, out);
    };
    for (int base = b_lo; base < b_hi; base += 32 * TILE_BU) {
      unsigned ent[TILE_BU];
      Real xv[TILE_BU];
#pragma unroll
      for (int u = 0; u < TILE_BU; u++) {
        const int p = base + 32 * u + lane;
        ent[u] = p < b_hi ? __ldcs(a.b_ent + p) : TILE_NO_KEY;
        xv[u] = (UNIT || p >= b_hi) ? Real(1) : __ldcs(a.b_val + p);
      }
#pragma unroll
      for (int u = 0; u < TILE_BU; u++) {
        if (base + 32 * u >= b_hi) // warp-uniform
          break;
        const unsigned key = ent[u] == TILE_NO_KEY ? TILE_NO_KEY : (ent[u] & 0xffffu);
        Real sa = 0, sb = 0;
        if (key != TILE_NO_KEY) {
          const Pair<Real> v = s_eq[row0 + static_cast<int>(ent[u] >> 16)];
          field_stats<Real, IS_V>(v.x, v.y, xv[u], s_tab[key], alpha, sa, sb);
        }
        tile_run_scan(key, sa, sb, lane);
        const unsigned next_key = __shfl_down_sync(FULL_MASK, key, 1);
        const unsigned head_key = __shfl_sync(FULL_MASK, key, 0);
        const int n_valid = __popc(__ballot_sync(FULL_MASK, key != TILE_NO_KEY)); // valid lanes: 0 .. n_valid - 1
        // Every branch below is warp-uniform (carry_*, first_done and `tails` are the same in all lanes).
        // The carried run either continues into this iteration's head run or is finished; a finished
        // run is final unless it is the warp's first one, which waits for the join across warps.
        if (carry_col != TILE_NO_KEY) {
          if (carry_col == head_key) {
            if (key == head_key)
              sa = carry_a + sa, sb = carry_b + sb;
          } else {
            if (!first_done)
              rec.first_col = carry_col, rec.first_a = carry_a, rec.first_b = carry_b;
            else if (lane == 0)
              store_run(carry_col, carry_a, carry_b);
            first_done = true;
          }
        }
        const bool is_tail = key != TILE_NO_KEY && (lane == n_valid - 1 || next_key != key);
        const bool finished = is_tail && lane != n_valid - 1; // the run of the last valid lane is carried on
        const unsigned tails = __ballot_sync(FULL_MASK, finished);
        if (tails) {
          int src = -1;
          if (!first_done) {
            src = __ffs(tails) - 1;
            rec.first_col = __shfl_sync(FULL_MASK, key, src);
            rec.first_a = __shfl_sync(FULL_MASK, sa, src), rec.first_b = __shfl_sync(FULL_MASK, sb, src);
            first_done = true;
          }
          if (finished && lane != src)
            store_run(key, sa, sb);
        }
        carry_col = __shfl_sync(FULL_MASK, key, n_valid - 1);
        carry_a = __shfl_sync(FULL_MASK, sa, n_valid - 1);
        carry_b = __shfl_sync(FULL_MASK, sb, n_valid - 1);
      }
    }
    if (carry_col != TILE_NO_KEY) {
      if (!first_done)
        rec.first_col = carry_col, rec.first_a = carry_a, rec.first_b = carry_b; // one run in the whole range
      else
        rec.last_col = carry_col, rec.last_a = carry_a, rec.last_b = carry_b;
    }
    if (lane == 0)
      s_rec[warp] = rec;
    __syncthreads();
    if (threadIdx.x == 0) { // runs that continue across warps: joined in warp order
      unsigned col = TILE_NO_KEY;
      Real sa = 0, sb = 0;
      auto flush = [&]() {
        if (col != TILE_NO_KEY) {
          Pair<Real> out;
          out.x = sa, out.y = sb;
          __stcg(part + col, out);
        }
      };
      for (int w = 0; w < TILE_WARPS; w++) {
        const TileRunRecord<Real> r = s_rec[w];
        if (r.first_col != TILE_NO_KEY) {
          if (r.first_col == col) {
            sa += r.first_a, sb += r.first_b;
          } else {
            flush();
            col = r.first_col, sa = r.first_a, sb = r.first_b;
          }
        }
        if (r.last_col != TILE_NO_KEY) {
          flush();
          col = r.last_col, sa = r.last_a, sb = r.last_b;
        }
      }
      flush();
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

// 256 threads = 32 columns x 8 tile groups: warp g adds the tiles g, g + 8, ... of 32 neighbouring
// columns (coalesced rows of `part`), warp 0 joins the 8 groups in group order and draws.
constexpr int FOLD_GROUPS = 8;
template <typename Real, bool IS_V> __global__ void __launch_bounds__(256) k_tile_fold(TileFoldArgs<Real> a) {
  __shared__ Real s_sum[FOLD_GROUPS][2][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int w = blockIdx.x * 32 + lane;
  const bool live = w < a.n_cols;
  const int j = live ? a.cols[w] : 0, c = j - a.last_base;
  Real sq = 0, lin = 0;
  if (live) {
#pragma unroll 4
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
