// Micro-benchmark (not part of the product): achievable random gather / scatter rate of 8-byte
// {e,q} pairs on B200, to set the floor for the non-contiguous column sweeps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <numeric>
#include <random>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

template <int U, int MODE> // MODE 0 gather-sum, 1 gather+scatter (rmw), 2 scatter only
__global__ void __launch_bounds__(256) k_gather(int n, const int *__restrict__ idx, float2 *__restrict__ eq, float *out) {
  const int base = blockIdx.x * (256 * U) + threadIdx.x;
  int i[U];
  float2 v[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const int p = base + u * 256;
    i[u] = p < n ? __ldcs(idx + p) : -1;
  }
  float acc = 0;
  if (MODE != 2) {
#pragma unroll
    for (int u = 0; u < U; u++)
      v[u] = i[u] >= 0 ? __ldcg(eq + i[u]) : make_float2(0, 0);
#pragma unroll
    for (int u = 0; u < U; u++)
      acc += v[u].x * v[u].y;
  }
  if (MODE == 0) {
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    if (acc == 123.456f) out[blockIdx.x] = acc;
  } else {
#pragma unroll
    for (int u = 0; u < U; u++)
      if (i[u] >= 0) {
        float2 w = MODE == 2 ? make_float2(1.f, 2.f) : make_float2(v[u].x + 1.f, v[u].y + 0.5f);
        __stcg(eq + i[u], w);
      }
  }
}

// streaming pass with a shared-memory table lookup per row: eq[i] read+write, idx2[i] read
__global__ void __launch_bounds__(1024) k_stream_tab(int n, const int2 *__restrict__ idx2, float2 *__restrict__ eq,
                                                      const float4 *__restrict__ tab, int n_tab) {
  extern __shared__ float4 s_tab[];
  for (int t = threadIdx.x; t < n_tab; t += blockDim.x) s_tab[t] = tab[t];
  __syncthreads();
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int2 j = __ldcs(idx2 + i);
    float2 v = __ldcg(eq + i);
    float4 t = s_tab[j.y];
    v.x += (v.y - t.x) * (t.y - t.x);
    v.y = t.z + 1.0f;
    __stcg(eq + i, v);
  }
}

template <typename F> float time_it(F f, int reps = 5) {
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int r = 0; r < reps; r++) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps * 1000.f; // us
}

int main() {
  const int n = 10000054, n_tab = 10677;
  std::vector<int> perm(n);
  std::iota(perm.begin(), perm.end(), 0);
  std::mt19937 g(1);
  std::shuffle(perm.begin(), perm.end(), g);
  // movie-like pattern: segments sorted ascending (CSC columns list rows in ascending order)
  std::vector<int> sorted_seg(perm);
  for (int s = 0; s + 937 <= n; s += 937) std::sort(sorted_seg.begin() + s, sorted_seg.begin() + s + 937);
  std::vector<int2> idx2(n);
  for (int i = 0; i < n; i++) idx2[i] = make_int2(i / 143, perm[i] % n_tab);
  int *d_idx, *d_idx_sorted; int2 *d_idx2; float2 *d_eq; float *d_out; float4 *d_tab;
  CK(cudaMalloc(&d_idx, n * 4)); CK(cudaMalloc(&d_idx_sorted, n * 4)); CK(cudaMalloc(&d_idx2, n * 8));
  CK(cudaMalloc(&d_eq, n * 8)); CK(cudaMalloc(&d_out, 1 << 20)); CK(cudaMalloc(&d_tab, n_tab * 16));
  CK(cudaMemcpy(d_idx, perm.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_idx_sorted, sorted_seg.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_idx2, idx2.data(), n * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_eq, 0, n * 8)); CK(cudaMemset(d_tab, 0, n_tab * 16));
#define RUN(U, MODE, IDX, NAME) { int grid = (n + 256 * U - 1) / (256 * U); \
    float us = time_it([&] { k_gather<U, MODE><<<grid, 256>>>(n, IDX, d_eq, d_out); }); \
    printf("%-28s U=%d  %8.1f us  %6.2f Gentries/s\n", NAME, U, us, n / us / 1e3); }
  RUN(1, 0, d_idx, "gather random") RUN(2, 0, d_idx, "gather random") RUN(4, 0, d_idx, "gather random") RUN(8, 0, d_idx, "gather random")
  RUN(4, 0, d_idx_sorted, "gather col-sorted") RUN(8, 0, d_idx_sorted, "gather col-sorted")
  RUN(1, 1, d_idx, "gather+scatter random") RUN(2, 1, d_idx, "gather+scatter random") RUN(4, 1, d_idx, "gather+scatter random") RUN(8, 1, d_idx, "gather+scatter random")
  RUN(4, 1, d_idx_sorted, "gather+scatter col-sorted")
  RUN(4, 2, d_idx, "scatter only random") RUN(8, 2, d_idx, "scatter only random")
  CK(cudaGetLastError());
  CK(cudaFuncSetAttribute(k_stream_tab, cudaFuncAttributeMaxDynamicSharedMemorySize, n_tab * 16));
  for (int blocks_per_sm = 1; blocks_per_sm <= 1; blocks_per_sm++) {
    float us = time_it([&] { k_stream_tab<<<148 * blocks_per_sm, 1024, n_tab * 16>>>(n, d_idx2, d_eq, d_tab, n_tab); });
    printf("stream + smem table (read 16B, write 8B per row)  %8.1f us  %6.1f GB/s\n", us, n * 24.0 / us / 1e3);
  }
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  return 0;
}
