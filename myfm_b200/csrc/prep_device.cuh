// Data preparation on the device for main tables that are stacks of position-aligned categorical
// fields (every row has exactly L entries, the k-th inside the k-th field's column range — what
// `group_shapes` describes and every MovieLens-shaped workload is).  Replaces, for that shape, the
// host passes of BaseFMTrainer's constructor (BaseFMTrainer.hpp:58-105: X_t = X.transpose()) and
// of csrc/host_data.hpp (row order, permuted CSR, CSC, field arrays): the caller's index array is
// uploaded once, then
//   k_prep_scan      column range of every position, index bounds, agreement with given levels
//   radix sort       rows by their first-field column (stable: ascending row inside a column)
//   k_prep_permute   permuted CSR, the field arrays of the streaming pass, targets in device order
//   histogram + scan column pointers of the CSC
//   radix sorts      one per further field: the CSC entries of that field (rows ascending per column)
// The host only sees O(columns) data (column pointers -> work items) and the row permutation.
#pragma once

#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace myfm {

constexpr int PREP_MAX_FIELDS = 64;

// stat: [0 .. L) min column per position, [64 .. 64 + L) max column, [128] error flags
// (1: index out of range, 2: a column's given level differs from its position)
__global__ void __launch_bounds__(256) k_prep_scan(int64_t n_rows, int L, int n_cols, const int *__restrict__ idx,
                                                    const int *__restrict__ given_level, int *__restrict__ stat) {
  __shared__ int s_min[PREP_MAX_FIELDS], s_max[PREP_MAX_FIELDS], s_err;
  if (threadIdx.x < PREP_MAX_FIELDS)
    s_min[threadIdx.x] = 0x7fffffff, s_max[threadIdx.x] = -1;
  if (threadIdx.x == 0)
    s_err = 0;
  __syncthreads();
  int err = 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_rows;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    for (int k = 0; k < L; k++) {
      const int c = idx[i * L + k];
      if (c < 0 || c >= n_cols) {
        err |= 1;
        continue;
      }
      if (given_level && given_level[c] != k)
        err |= 2;
      if (c < s_min[k])
        atomicMin(&s_min[k], c);
      if (c > s_max[k])
        atomicMax(&s_max[k], c);
    }
  if (err)
    atomicOr(&s_err, err);
  __syncthreads();
  if (threadIdx.x < L) {
    atomicMin(&stat[threadIdx.x], s_min[threadIdx.x]);
    atomicMax(&stat[64 + threadIdx.x], s_max[threadIdx.x]);
  }
  if (threadIdx.x == 0 && s_err)
    atomicOr(&stat[128], s_err);
}

// keys[i] = column of row i at `position`; rows[i] = i
__global__ void __launch_bounds__(256) k_prep_keys(int64_t n_rows, int L, int position, const int *__restrict__ idx,
                                                    int *__restrict__ keys, int *__restrict__ rows) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n_rows) {
    keys[i] = idx[i * L + position];
    rows[i] = static_cast<int>(i);
  }
}

// Device row i' = caller's row perm[i']: CSR in device order, the field arrays tail_idx / tail_val
// ([L-1][n]: entries 1 .. L-1 of every row) and own_val ([n]: entry 0), column counts.
template <typename Real>
__global__ void __launch_bounds__(256)
    k_prep_permute(int64_t n_rows, int L, const int *__restrict__ perm, const int *__restrict__ idx_in,
                   const double *__restrict__ val_in, int *__restrict__ csr_ptr, int *__restrict__ csr_idx,
                   Real *__restrict__ csr_val, int *__restrict__ tail_idx, Real *__restrict__ tail_val,
                   Real *__restrict__ own_val, int *__restrict__ col_count) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i > n_rows)
    return;
  if (i == n_rows) {
    csr_ptr[i] = static_cast<int>(i * L);
    return;
  }
  csr_ptr[i] = static_cast<int>(i * L);
  const int64_t src = static_cast<int64_t>(perm[i]) * L;
  for (int k = 0; k < L; k++) {
    const int c = idx_in[src + k];
    const Real v = val_in ? static_cast<Real>(val_in[src + k]) : Real(1);
    csr_idx[i * L + k] = c;
    csr_val[i * L + k] = v;
    atomicAdd(col_count + c, 1);
    if (k == 0) {
      if (own_val)
        own_val[i] = v;
    } else {
      tail_idx[static_cast<int64_t>(k - 1) * n_rows + i] = c;
      if (tail_val)
        tail_val[static_cast<int64_t>(k - 1) * n_rows + i] = v;
    }
  }
}

// CSC entries of the first field: the rows are sorted by it, so entry p is row p.
template <typename Real>
__global__ void __launch_bounds__(256) k_prep_csc_first(int64_t n_rows, const Real *__restrict__ own_val,
                                                         int *__restrict__ csc_idx, Real *__restrict__ csc_val) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n_rows) {
    csc_idx[i] = static_cast<int>(i);
    csc_val[i] = own_val ? own_val[i] : Real(1);
  }
}
// ... of a further field: rows sorted by that field's column; the value follows its row
template <typename Real>
__global__ void __launch_bounds__(256) k_prep_csc_val(int64_t n_rows, const int *__restrict__ rows_sorted,
                                                       const Real *__restrict__ field_val, Real *__restrict__ csc_val) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n_rows)
    csc_val[i] = field_val ? field_val[rows_sorted[i]] : Real(1);
}

// Stable sort of (key, row) pairs by key with cub; `bits`: significant key bits.
struct PrepSorter {
  DevBuf<unsigned char> temp;
  void sort(const int *keys_in, int *keys_out, const int *vals_in, int *vals_out, int64_t n, int bits,
            cudaStream_t stream) {
    size_t bytes = 0;
    MYFM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, vals_in, vals_out, static_cast<int>(n), 0,
                                              bits, stream));
    if (bytes > temp.n)
      temp.alloc(bytes);
    MYFM_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, bytes, keys_in, keys_out, vals_in, vals_out, static_cast<int>(n), 0,
                                              bits, stream));
  }
  void exclusive_sum(const int *in, int *out, int64_t n, cudaStream_t stream) {
    size_t bytes = 0;
    MYFM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, static_cast<int>(n), stream));
    if (bytes > temp.n)
      temp.alloc(bytes);
    MYFM_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, bytes, in, out, static_cast<int>(n), stream));
  }
};

} // namespace myfm
