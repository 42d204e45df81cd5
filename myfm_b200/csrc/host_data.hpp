// Host-side data preparation shared by the trainer and the prediction datasets: validation of the
// C-ABI inputs (with the reference's error behaviour), int32 re-indexing, transposition,
// dependency-level schedules and the validated learning config.  Pure C++ — runs without a GPU.
#pragma once

#include "../../include/myfm_b200.h"

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <exception>
#include <thread>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace myfm {

// Host-side data preparation is a handful of O(nnz) passes; they run on a few threads
// (MYFM_HOST_THREADS, default min(16, hardware threads)).  fn(part, n_parts) is called once per part.
inline std::atomic<int> &host_threads_override() {
  static std::atomic<int> v{0};
  return v;
}
inline int host_threads() {
  static const int n = [] {
    const char *env = std::getenv("MYFM_HOST_THREADS");
    int v = env ? std::atoi(env) : 0;
    if (v <= 0)
      v = static_cast<int>(std::min(16u, std::max(1u, std::thread::hardware_concurrency())));
    return v;
  }();
  const int o = host_threads_override().load();
  return o > 0 ? o : n;
}
template <typename Fn> void parallel_parts(int n_parts, Fn fn) {
  if (n_parts <= 1) {
    fn(0, 1);
    return;
  }
  std::vector<std::thread> pool;
  std::vector<std::exception_ptr> errors(n_parts);
  for (int t = 0; t < n_parts; t++)
    pool.emplace_back([&, t] {
      try {
        fn(t, n_parts);
      } catch (...) {
        errors[t] = std::current_exception();
      }
    });
  for (auto &th : pool)
    th.join();
  for (auto &e : errors)
    if (e)
      std::rethrow_exception(e);
}
// [begin, end) of part t when n items are cut into n_parts contiguous ranges
inline std::pair<int64_t, int64_t> part_range(int64_t n, int t, int n_parts) {
  return {n * t / n_parts, n * (t + 1) / n_parts};
}
inline int parts_for(int64_t work) { return (work < (1 << 18) && host_threads_override().load() <= 0) ? 1 : host_threads(); }

template <typename Real> struct HostCs { // compressed sparse, major x minor
  int64_t n_major = 0, n_minor = 0;
  std::vector<int> ptr;
  std::vector<int> idx;
  std::vector<Real> val;
  int64_t nnz() const { return static_cast<int64_t>(idx.size()); }
};

template <typename Real> HostCs<Real> host_from_api(const myfm_csr_t &m, const char *what) {
  if (m.n_rows < 0 || m.n_cols < 0 || (m.n_rows > 0 && m.indptr == nullptr))
    throw std::invalid_argument(std::string(what) + ": malformed CSR matrix.");
  HostCs<Real> out;
  out.n_major = m.n_rows;
  out.n_minor = m.n_cols;
  if (m.n_rows > 0 && m.indptr[0] != 0)
    throw std::invalid_argument(std::string(what) + ": indptr[0] must be 0 (pass a canonical CSR, not a slice view).");
  const int64_t nnz = m.n_rows ? m.indptr[m.n_rows] : 0;
  if (nnz > 0 && (m.indices == nullptr || m.data == nullptr))
    throw std::invalid_argument(std::string(what) + ": indices / data must not be NULL.");
  if (nnz < 0 || nnz >= std::numeric_limits<int>::max() ||
      m.n_rows >= std::numeric_limits<int>::max() || m.n_cols >= std::numeric_limits<int>::max())
    throw std::invalid_argument(std::string(what) +
                                ": a shard must have fewer than 2^31 rows, columns and non-zeros.");
  out.ptr.resize(m.n_rows + 1);
  out.ptr[0] = 0;
  out.idx.resize(nnz);
  out.val.resize(nnz);
  parallel_parts(parts_for(nnz), [&](int t, int n_parts) {
    auto [r0, r1] = part_range(m.n_rows, t, n_parts);
    for (int64_t r = r0; r < r1; r++) {
      if (m.indptr[r + 1] < m.indptr[r])
        throw std::invalid_argument(std::string(what) + ": indptr is not monotone.");
      out.ptr[r + 1] = static_cast<int>(m.indptr[r + 1]);
    }
    auto [p0, p1] = part_range(nnz, t, n_parts);
    for (int64_t p = p0; p < p1; p++) {
      const int j = m.indices[p];
      if (j < 0 || j >= m.n_cols)
        throw std::invalid_argument(std::string(what) + ": column index out of range.");
      out.idx[p] = j;
      out.val[p] = static_cast<Real>(m.data[p]);
    }
  });
  return out;
}

// Minor-major copy; entries of each output row keep ascending source-row order
// (BaseFMTrainer.hpp:61, definitions.hpp:59).
template <typename Real> HostCs<Real> host_transpose(const HostCs<Real> &a) {
  HostCs<Real> t;
  t.n_major = a.n_minor;
  t.n_minor = a.n_major;
  t.ptr.assign(a.n_minor + 1, 0);
  t.idx.resize(a.idx.size());
  t.val.resize(a.val.size());
  // counting sort by column; every part owns a contiguous row range, so inside a column the
  // entries stay in ascending row order whatever the number of parts
  const int n_parts = a.n_minor > (1 << 22) ? 1 : parts_for(a.nnz());
  std::vector<std::vector<int>> count(n_parts);
  parallel_parts(n_parts, [&](int p, int np) {
    auto [r0, r1] = part_range(a.n_major, p, np);
    std::vector<int> &c = count[p];
    c.assign(a.n_minor, 0);
    for (int q = a.ptr[r0]; q < a.ptr[r1]; q++)
      c[a.idx[q]]++;
  });
  for (int64_t c = 0; c < a.n_minor; c++) { // count[p][c] becomes the first slot of part p in column c
    int at = t.ptr[c];
    for (int p = 0; p < n_parts; p++) {
      const int n = count[p][c];
      count[p][c] = at;
      at += n;
    }
    t.ptr[c + 1] = at;
  }
  parallel_parts(n_parts, [&](int p, int np) {
    auto [r0, r1] = part_range(a.n_major, p, np);
    std::vector<int> &cur = count[p];
    for (int64_t r = r0; r < r1; r++)
      for (int q = a.ptr[r]; q < a.ptr[r + 1]; q++) {
        const int dst = cur[a.idx[q]]++;
        t.idx[dst] = static_cast<int>(r);
        t.val[dst] = a.val[q];
      }
  });
  return t;
}

// A row that lists the same column twice makes the reference's serial pass-2 read its own
// partial update; no parallel schedule can reproduce that, so such input is rejected (the Python
// layer sums duplicates before calling).
template <typename Real> bool has_duplicate_entries(const HostCs<Real> &csr) {
  std::atomic<bool> found{false};
  parallel_parts(csr.n_minor > (1 << 22) ? 1 : parts_for(csr.nnz()), [&](int t, int np) {
    auto [r0, r1] = part_range(csr.n_major, t, np);
    std::vector<int> seen(csr.n_minor, -1); // last row that listed the column, per part
    for (int64_t r = r0; r < r1 && !found.load(std::memory_order_relaxed); r++)
      for (int p = csr.ptr[r]; p < csr.ptr[r + 1]; p++) {
        if (seen[csr.idx[p]] == static_cast<int>(r)) {
          found.store(true);
          break;
        }
        seen[csr.idx[p]] = static_cast<int>(r);
      }
  });
  return found.load();
}

// Dependency levels of the columns (given as the major axis of `csc`): the reference updates
// columns strictly in index order and column j sees every change made by an earlier column that
// shares a row with it.  level(j) = 1 + max level of such earlier columns; columns of one level are
// pairwise row-disjoint, so updating them concurrently and running the levels in order is the
// serial sweep exactly.  One pass over the non-zeros.
// `lower` (optional): lower bounds per column (level consensus between row shards).
template <typename Real>
std::vector<int> compute_levels(const HostCs<Real> &csc, int *n_levels_out,
                                const int *lower = nullptr) {
  std::vector<int> next_level(csc.n_minor, 0); // per row: first level still free
  std::vector<int> level(csc.n_major, 0);
  int n_levels = csc.n_major ? 1 : 0;
  for (int64_t j = 0; j < csc.n_major; j++) {
    int lv = lower ? lower[j] : 0;
    for (int p = csc.ptr[j]; p < csc.ptr[j + 1]; p++)
      lv = std::max(lv, next_level[csc.idx[p]]);
    for (int p = csc.ptr[j]; p < csc.ptr[j + 1]; p++)
      next_level[csc.idx[p]] = lv + 1;
    level[j] = lv;
    n_levels = std::max(n_levels, lv + 1);
  }
  *n_levels_out = n_levels;
  return level;
}

// The same schedule from the ROW side, on several threads.  Inside a row the entries in ascending
// column order form a chain (each conflicts with its predecessor), and the level of a column is
// one more than the highest level among its predecessors over all its rows — which is exactly the
// serial recurrence above, because the last earlier column of a row already dominates the ones
// before it.  The least fixed point is reached by relaxing all rows repeatedly (monotone, so the
// order of the concurrent updates does not matter); the number of passes is the number of levels.
// Needs every row's columns in ascending order: returns false (and leaves the output untouched)
// when some row is not, the caller then takes compute_levels.
template <typename Real>
bool compute_levels_by_rows(const HostCs<Real> &csr, std::vector<int> &level_out, int *n_levels_out,
                            const int *lower = nullptr) {
  const int64_t n_rows = csr.n_major, n_cols = csr.n_minor;
  const int n_parts = parts_for(csr.nnz());
  std::atomic<bool> sorted{true};
  parallel_parts(n_parts, [&](int t, int np) {
    auto [r0, r1] = part_range(n_rows, t, np);
    for (int64_t r = r0; r < r1 && sorted.load(std::memory_order_relaxed); r++)
      for (int p = csr.ptr[r] + 1; p < csr.ptr[r + 1]; p++)
        if (csr.idx[p] <= csr.idx[p - 1]) {
          sorted.store(false);
          break;
        }
  });
  if (!sorted.load())
    return false;
  std::vector<std::atomic<int>> level(n_cols);
  for (int64_t j = 0; j < n_cols; j++)
    level[j].store(lower ? lower[j] : 0, std::memory_order_relaxed);
  for (int pass = 0; pass <= n_cols; pass++) {
    std::atomic<bool> changed{false};
    parallel_parts(n_parts, [&](int t, int np) {
      auto [r0, r1] = part_range(n_rows, t, np);
      bool mine = false;
      for (int64_t r = r0; r < r1; r++) {
        int need = 0; // level the next entry of this row must reach at least
        for (int p = csr.ptr[r]; p < csr.ptr[r + 1]; p++) {
          std::atomic<int> &lv = level[csr.idx[p]];
          int cur = lv.load(std::memory_order_relaxed);
          while (cur < need && !lv.compare_exchange_weak(cur, need, std::memory_order_relaxed))
            ;
          if (cur < need)
            mine = true, cur = need;
          need = cur + 1;
        }
      }
      if (mine)
        changed.store(true);
    });
    if (!changed.load())
      break;
  }
  level_out.resize(n_cols);
  int n_levels = n_cols ? 1 : 0;
  for (int64_t j = 0; j < n_cols; j++) {
    level_out[j] = level[j].load(std::memory_order_relaxed);
    n_levels = std::max(n_levels, level_out[j] + 1);
  }
  *n_levels_out = n_levels;
  return true;
}

// Columns grouped by level.  Inside a level, columns longer than `long_threshold` come first
// (block-per-column kernel), then the rest by descending length (warp-per-column kernel; similar
// lengths share a thread block).
struct LevelPlan {
  int n_levels = 0;
  std::vector<int> level_ptr;  // [n_levels + 1] into cols
  std::vector<int> n_long;     // [n_levels]
  std::vector<int> cols;
  // long columns cut into segments of at most seg_nnz entries (k_seg_stats / k_seg_update)
  std::vector<int> seg_level_ptr;  // [n_levels + 1] into seg_*
  std::vector<int> seg_col, seg_lo, seg_hi, seg_slot;
  std::vector<int> slot_level_ptr; // [n_levels + 1] into slot_ptr (one extra entry per level)
  std::vector<int> slot_ptr;       // per level: [n_long + 1] segment offsets RELATIVE to the level
  int max_slots = 0, max_segs = 0;
};

template <typename Real>
LevelPlan make_level_plan(const HostCs<Real> &csc, int long_threshold, int seg_nnz = 1024) {
  LevelPlan plan;
  std::vector<int> level = compute_levels(csc, &plan.n_levels);
  plan.level_ptr.assign(plan.n_levels + 1, 0);
  for (int lv : level)
    plan.level_ptr[lv + 1]++;
  for (int l = 0; l < plan.n_levels; l++)
    plan.level_ptr[l + 1] += plan.level_ptr[l];
  plan.cols.resize(csc.n_major);
  std::vector<int> cur(plan.level_ptr.begin(), plan.level_ptr.end() - 1);
  for (int64_t j = 0; j < csc.n_major; j++)
    plan.cols[cur[level[j]]++] = static_cast<int>(j);
  plan.n_long.assign(plan.n_levels, 0);
  auto len = [&](int j) { return csc.ptr[j + 1] - csc.ptr[j]; };
  for (int l = 0; l < plan.n_levels; l++) {
    auto b = plan.cols.begin() + plan.level_ptr[l], e = plan.cols.begin() + plan.level_ptr[l + 1];
    std::stable_sort(b, e, [&](int x, int y) { return len(x) > len(y); });
    plan.n_long[l] = static_cast<int>(
        std::find_if(b, e, [&](int j) { return len(j) <= long_threshold; }) - b);
  }
  plan.seg_level_ptr.assign(plan.n_levels + 1, 0);
  plan.slot_level_ptr.assign(plan.n_levels + 1, 0);
  for (int l = 0; l < plan.n_levels; l++) {
    int segs_in_level = 0;
    for (int k = 0; k < plan.n_long[l]; k++) {
      const int j = plan.cols[plan.level_ptr[l] + k];
      plan.slot_ptr.push_back(segs_in_level);
      for (int lo = csc.ptr[j]; lo < csc.ptr[j + 1]; lo += seg_nnz) {
        plan.seg_col.push_back(j);
        plan.seg_lo.push_back(lo);
        plan.seg_hi.push_back(std::min(lo + seg_nnz, csc.ptr[j + 1]));
        plan.seg_slot.push_back(k);
        segs_in_level++;
      }
    }
    plan.slot_ptr.push_back(segs_in_level);
    plan.seg_level_ptr[l + 1] = plan.seg_level_ptr[l] + segs_in_level;
    plan.slot_level_ptr[l + 1] = static_cast<int>(plan.slot_ptr.size());
    plan.max_slots = std::max(plan.max_slots, plan.n_long[l]);
    plan.max_segs = std::max(plan.max_segs, segs_in_level);
  }
  return plan;
}

// ------------------------------------------------------------------------------------------------
// Main-table sweep plan (engine.cu: Trainer).
//
// Row order.  The sampler's result does not depend on the order of the training rows (only the
// order of summation inside a column does, at rounding level), so the engine stores rows in its
// own order: rows are grouped by their column in the PRIMARY level (the dependency level that
// holds the most entries; for one-hot fields the first field), columns ascending, original row
// order inside a column, rows without an entry in that level last.  In that order every column of
// the primary level is a contiguous row range: its sweep streams instead of gathering.
//
// Work items.  Per level, columns are split by length into
//   S: columns longer than `chunk` entries, cut into chunks of `chunk` entries; two-phase
//      (chunk statistics -> ordered reduction + draw + update), one CTA per chunk;
//   C: columns of (warp_max, chunk] entries, one CTA each, single pass with the gathered values
//      kept in registers between the reduction and the update;
//   W: columns of at most warp_max entries, one warp each, likewise.
// Items of a level are stored S first, then C, then W, each by descending length.
// ------------------------------------------------------------------------------------------------
struct SweepLevel {
  int s0 = 0, c0 = 0, w0 = 0, end = 0; // item ranges [s0,c0) S, [c0,w0) C, [w0,end) W
  bool unit = true;                    // every value of the level is exactly 1
  bool contig = true;                  // every column of the level is a contiguous row range
  int64_t nnz = 0;
};

struct alignas(16) SweepItem {
  int col, lo, hi; // column and its entry range [lo, hi) in the CSC arrays
  int first;       // S items: first chunk item of the column (level-relative)
};

struct SweepPlan {
  std::vector<SweepLevel> levels;
  std::vector<SweepItem> items;
  std::vector<int> seg_count; // S items: number of chunks of the column
  std::vector<int> item_slot; // rank of the item's column among the level's columns (by index)
  int max_level_cols = 0;
  int max_seg_items = 0;
  int primary_level = -1;
};

// perm[i'] = original row stored at device row i'.
template <typename Real>
std::vector<int> primary_row_order(const HostCs<Real> &csc, const std::vector<int> &level,
                                   int n_levels, int *primary_level) {
  const int64_t n_rows = csc.n_minor;
  std::vector<int64_t> level_nnz(std::max(n_levels, 1), 0);
  for (int64_t j = 0; j < csc.n_major; j++)
    level_nnz[level[j]] += csc.ptr[j + 1] - csc.ptr[j];
  int best = -1;
  for (int l = 0; l < n_levels; l++)
    if (level_nnz[l] > 0 && (best < 0 || level_nnz[l] > level_nnz[best]))
      best = l;
  *primary_level = best;
  std::vector<int> perm(n_rows);
  std::vector<char> taken(n_rows, 0);
  int64_t n_primary = 0;
  if (best >= 0) {
    std::vector<int64_t> off(csc.n_major + 1, 0);
    for (int64_t j = 0; j < csc.n_major; j++)
      off[j + 1] = off[j] + (level[j] == best ? csc.ptr[j + 1] - csc.ptr[j] : 0);
    n_primary = off[csc.n_major];
    parallel_parts(parts_for(n_primary), [&](int t, int n_parts) {
      auto [j0, j1] = part_range(csc.n_major, t, n_parts);
      for (int64_t j = j0; j < j1; j++)
        if (level[j] == best) {
          int64_t at = off[j];
          for (int p = csc.ptr[j]; p < csc.ptr[j + 1]; p++) {
            perm[at++] = csc.idx[p];
            taken[csc.idx[p]] = 1; // columns of a level are row-disjoint: no two parts touch one row
          }
        }
    });
  }
  int64_t at = n_primary;
  for (int64_t i = 0; i < n_rows; i++)
    if (!taken[i])
      perm[at++] = static_cast<int>(i);
  return perm;
}

template <typename Real>
HostCs<Real> permute_rows(const HostCs<Real> &csr, const std::vector<int> &perm) {
  HostCs<Real> out;
  out.n_major = csr.n_major, out.n_minor = csr.n_minor;
  out.ptr.resize(csr.n_major + 1);
  out.idx.resize(csr.idx.size());
  out.val.resize(csr.val.size());
  out.ptr[0] = 0;
  for (int64_t i = 0; i < csr.n_major; i++)
    out.ptr[i + 1] = out.ptr[i] + (csr.ptr[perm[i] + 1] - csr.ptr[perm[i]]);
  parallel_parts(parts_for(csr.nnz()), [&](int t, int n_parts) {
    auto [i0, i1] = part_range(csr.n_major, t, n_parts);
    for (int64_t i = i0; i < i1; i++) {
      const int src = perm[i], b = csr.ptr[src], n = csr.ptr[src + 1] - b, dst = out.ptr[i];
      std::copy(csr.idx.begin() + b, csr.idx.begin() + b + n, out.idx.begin() + dst);
      std::copy(csr.val.begin() + b, csr.val.begin() + b + n, out.val.begin() + dst);
    }
  });
  return out;
}

// What make_sweep_plan finds out per level by looking at every entry, when the caller knows it already
// (device-side preparation, csrc/prep_device.cuh: the CSC entries never come to the host).
struct LevelFlags {
  std::vector<char> unit, contig;
};

template <typename Real>
SweepPlan make_sweep_plan(const HostCs<Real> &csc, const std::vector<int> &level, int n_levels,
                          int warp_max, int chunk, int only_level = -1, const LevelFlags *known = nullptr) {
  SweepPlan plan;
  plan.levels.resize(n_levels);
  std::vector<std::vector<int>> cols(n_levels);
  for (int64_t j = 0; j < csc.n_major; j++)
    if (only_level < 0 || level[j] == only_level)
      cols[level[j]].push_back(static_cast<int>(j));
  auto len = [&](int j) { return csc.ptr[j + 1] - csc.ptr[j]; };
  for (int l = 0; l < n_levels; l++) {
    SweepLevel &L = plan.levels[l];
    if (only_level >= 0 && l != only_level) {
      L.s0 = L.c0 = L.w0 = L.end = static_cast<int>(plan.items.size());
      continue;
    }
    std::vector<int> &c = cols[l];
    std::stable_sort(c.begin(), c.end(), [&](int x, int y) { return len(x) > len(y); });
    for (int j : c) {
      L.nnz += len(j);
      if (known)
        continue;
      for (int p = csc.ptr[j]; p < csc.ptr[j + 1]; p++) {
        if (csc.val[p] != Real(1))
          L.unit = false;
        if (csc.idx[p] != csc.idx[csc.ptr[j]] + (p - csc.ptr[j]))
          L.contig = false;
      }
    }
    if (known)
      L.unit = known->unit[l] != 0, L.contig = known->contig[l] != 0;
    L.s0 = static_cast<int>(plan.items.size());
    std::vector<int> by_index(c);
    std::sort(by_index.begin(), by_index.end());
    plan.max_level_cols = std::max(plan.max_level_cols, static_cast<int>(c.size()));
    auto push = [&](int j, int lo, int hi, int first, int count) {
      plan.items.push_back(SweepItem{j, lo, hi, first});
      plan.seg_count.push_back(count);
      plan.item_slot.push_back(
          static_cast<int>(std::lower_bound(by_index.begin(), by_index.end(), j) - by_index.begin()));
    };
    size_t k = 0;
    for (; k < c.size() && len(c[k]) > chunk; k++) {
      const int j = c[k], n = (len(j) + chunk - 1) / chunk;
      const int first = static_cast<int>(plan.items.size()) - L.s0; // level-relative
      for (int s = 0; s < n; s++)
        push(j, csc.ptr[j] + s * chunk, std::min(csc.ptr[j] + (s + 1) * chunk, csc.ptr[j + 1]), first, n);
    }
    L.c0 = static_cast<int>(plan.items.size());
    for (; k < c.size() && len(c[k]) > warp_max; k++)
      push(c[k], csc.ptr[c[k]], csc.ptr[c[k] + 1], 0, 0);
    L.w0 = static_cast<int>(plan.items.size());
    for (; k < c.size(); k++)
      push(c[k], csc.ptr[c[k]], csc.ptr[c[k] + 1], 0, 0);
    L.end = static_cast<int>(plan.items.size());
    plan.max_seg_items = std::max(plan.max_seg_items, L.c0 - L.s0);
  }
  return plan;
}

// FMLearningConfig (FMLearningConfig.hpp:17-57) after validation.
struct Config {
  double alpha_0, beta_0, gamma_0, mu_0, reg_0;
  int task_type;
  double nu_oprobit;
  bool fit_w0, fit_linear;
  int n_iter, n_kept_samples;
  double cutpoint_scale;
  std::vector<int> group_index;
  int n_groups = 0;
  std::vector<int> feat_ptr, feat_idx; // features of each group, ascending
  std::vector<std::pair<int, std::vector<int64_t>>> cutpoint_groups;

  explicit Config(const myfm_config_t &c) {
    alpha_0 = c.alpha_0, beta_0 = c.beta_0, gamma_0 = c.gamma_0, mu_0 = c.mu_0, reg_0 = c.reg_0;
    task_type = c.task_type;
    nu_oprobit = c.nu_oprobit;
    fit_w0 = c.fit_w0 != 0, fit_linear = c.fit_linear != 0;
    n_iter = c.n_iter, n_kept_samples = c.n_kept_samples;
    cutpoint_scale = c.cutpoint_scale;
    if (task_type < MYFM_TASK_REGRESSION || task_type > MYFM_TASK_ORDERED)
      throw std::invalid_argument("unknown task type.");
    if (c.n_group_index < 0 || (c.n_group_index > 0 && c.group_index == nullptr))
      throw std::invalid_argument("malformed group_index.");
    group_index.resize(c.n_group_index);
    std::vector<int64_t> sorted(c.group_index, c.group_index + c.n_group_index);
    std::sort(sorted.begin(), sorted.end());
    sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
    n_groups = static_cast<int>(sorted.size());
    for (int i = 0; i < n_groups; i++) // FMLearningConfig.hpp:29-40
      if (sorted[i] != i) {
        std::ostringstream ss;
        ss << "No matching index for group index " << i << " found.";
        throw std::invalid_argument(ss.str());
      }
    feat_ptr.assign(n_groups + 1, 0);
    for (int64_t f = 0; f < c.n_group_index; f++) {
      group_index[f] = static_cast<int>(c.group_index[f]);
      feat_ptr[group_index[f] + 1]++;
    }
    for (int g = 0; g < n_groups; g++)
      feat_ptr[g + 1] += feat_ptr[g];
    feat_idx.resize(c.n_group_index);
    std::vector<int> cur(feat_ptr.begin(), feat_ptr.end() - 1);
    for (int64_t f = 0; f < c.n_group_index; f++)
      feat_idx[cur[group_index[f]]++] = static_cast<int>(f);
    if (n_kept_samples < 0) // :48-56
      throw std::invalid_argument("n_kept_samples must be non-negative,");
    if (n_iter <= 0)
      throw std::invalid_argument("n_iter must be positive.");
    if (n_iter < n_kept_samples)
      throw std::invalid_argument("n_kept_samples must not exceed n_iter.");
    for (int g = 0; g < c.n_cutpoint_groups; g++)
      cutpoint_groups.emplace_back(
          c.cutpoint_n_class[g],
          std::vector<int64_t>(c.cutpoint_index[g], c.cutpoint_index[g] + c.cutpoint_index_len[g]));
  }
};

} // namespace myfm
