// Host-side data preparation shared by the trainer and the prediction datasets: validation of the
// C-ABI inputs (with the reference's error behaviour), int32 re-indexing, transposition,
// dependency-level schedules and the validated learning config.  Pure C++ — runs without a GPU.
#pragma once

#include "../../include/myfm_b200.h"

#include <algorithm>
#include <cstdint>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace myfm {

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
  const int64_t nnz = m.n_rows ? m.indptr[m.n_rows] : 0;
  if (nnz < 0 || nnz >= std::numeric_limits<int>::max() ||
      m.n_rows >= std::numeric_limits<int>::max() || m.n_cols >= std::numeric_limits<int>::max())
    throw std::invalid_argument(std::string(what) +
                                ": a shard must have fewer than 2^31 rows, columns and non-zeros.");
  out.ptr.resize(m.n_rows + 1);
  out.ptr[0] = 0;
  for (int64_t r = 0; r < m.n_rows; r++) {
    if (m.indptr[r + 1] < m.indptr[r])
      throw std::invalid_argument(std::string(what) + ": indptr is not monotone.");
    out.ptr[r + 1] = static_cast<int>(m.indptr[r + 1]);
  }
  out.idx.assign(m.indices, m.indices + nnz);
  out.val.resize(nnz);
  for (int64_t p = 0; p < nnz; p++) {
    if (out.idx[p] < 0 || out.idx[p] >= m.n_cols)
      throw std::invalid_argument(std::string(what) + ": column index out of range.");
    out.val[p] = static_cast<Real>(m.data[p]);
  }
  return out;
}

// Minor-major copy; entries of each output row keep ascending source-row order
// (BaseFMTrainer.hpp:61, definitions.hpp:59).
template <typename Real> HostCs<Real> host_transpose(const HostCs<Real> &a) {
  HostCs<Real> t;
  t.n_major = a.n_minor;
  t.n_minor = a.n_major;
  t.ptr.assign(a.n_minor + 1, 0);
  for (int c : a.idx)
    t.ptr[c + 1]++;
  for (int64_t c = 0; c < a.n_minor; c++)
    t.ptr[c + 1] += t.ptr[c];
  t.idx.resize(a.idx.size());
  t.val.resize(a.val.size());
  std::vector<int> cur(t.ptr.begin(), t.ptr.end() - 1);
  for (int64_t r = 0; r < a.n_major; r++)
    for (int p = a.ptr[r]; p < a.ptr[r + 1]; p++) {
      int dst = cur[a.idx[p]]++;
      t.idx[dst] = static_cast<int>(r);
      t.val[dst] = a.val[p];
    }
  return t;
}

// A row that lists the same column twice makes the reference's serial pass-2 read its own
// partial update; no parallel schedule can reproduce that, so such input is rejected (the Python
// layer sums duplicates before calling).
template <typename Real> bool has_duplicate_entries(const HostCs<Real> &csr) {
  std::vector<int> seen(csr.n_minor, -1);
  for (int64_t r = 0; r < csr.n_major; r++)
    for (int p = csr.ptr[r]; p < csr.ptr[r + 1]; p++) {
      if (seen[csr.idx[p]] == static_cast<int>(r))
        return true;
      seen[csr.idx[p]] = static_cast<int>(r);
    }
  return false;
}

// Dependency levels of the columns (given as the major axis of `csc`): the reference updates
// columns strictly in index order and column j sees every change made by an earlier column that
// shares a row with it.  level(j) = 1 + max level of such earlier columns; columns of one level are
// pairwise row-disjoint, so updating them concurrently and running the levels in order is the
// serial sweep exactly.  One pass over the non-zeros.
template <typename Real>
std::vector<int> compute_levels(const HostCs<Real> &csc, int *n_levels_out) {
  std::vector<int> next_level(csc.n_minor, 0); // per row: first level still free
  std::vector<int> level(csc.n_major, 0);
  int n_levels = csc.n_major ? 1 : 0;
  for (int64_t j = 0; j < csc.n_major; j++) {
    int lv = 0;
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
