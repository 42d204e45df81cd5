// Micro-benchmark: random 8-byte gathers confined to a window of W MB of an 80 MB table (is the
// gather-only level bound by L2 capacity?).  Each launch gathers n/parts entries from one window.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <numeric>
#include <random>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

template <int U>
__global__ void __launch_bounds__(256) k_gather(int n, const int *__restrict__ idx, const float2 *__restrict__ eq, float *out) {
  const int base = blockIdx.x * (256 * U) + threadIdx.x;
  int i[U];
  float2 v[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const int p = base + u * 256;
    i[u] = p < n ? __ldcs(idx + p) : -1;
  }
  float acc = 0;
#pragma unroll
  for (int u = 0; u < U; u++)
    v[u] = i[u] >= 0 ? __ldcg(eq + i[u]) : make_float2(0, 0);
#pragma unroll
  for (int u = 0; u < U; u++)
    acc += v[u].x * v[u].y;
  acc += __shfl_xor_sync(0xffffffffu, acc, 16);
  if (acc == 123.456f) out[blockIdx.x] = acc;
}
__global__ void k_stream(int n, float2 *eq, const int *tail) { // the streaming pass between gathers: rewrites eq
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float2 v = __ldcg(eq + i);
    v.x += (float)__ldcs(tail + i) * 1e-9f;
    __stcg(eq + i, v);
  }
}

int main() {
  const int n = 10000054;
  int *d_idx, *d_tail; float2 *d_eq; float *d_out;
  CK(cudaMalloc(&d_idx, n * 4)); CK(cudaMalloc(&d_tail, n * 4)); CK(cudaMalloc(&d_eq, n * 8)); CK(cudaMalloc(&d_out, 1 << 20));
  CK(cudaMemset(d_eq, 0, n * 8)); CK(cudaMemset(d_tail, 0, n * 4));
  std::mt19937 g(1);
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  for (int parts : {1, 2, 4, 8}) {
    // entries grouped by window: part k holds a random permutation of the rows of window k
    std::vector<int> perm(n);
    std::iota(perm.begin(), perm.end(), 0);
    const int per = (n + parts - 1) / parts;
    for (int k = 0; k < parts; k++) {
      int lo = k * per, hi = std::min(n, lo + per);
      std::shuffle(perm.begin() + lo, perm.begin() + hi, g);
    }
    CK(cudaMemcpy(d_idx, perm.data(), n * 4, cudaMemcpyHostToDevice));
    float tot = 0, tot_stream = 0;
    const int reps = 5;
    for (int r = 0; r < reps + 1; r++) {
      // streaming pass first, as in the sampler (evicts / rewrites eq), then the gathers window by window
      cudaEventRecord(a);
      k_stream<<<148 * 4, 512>>>(n, d_eq, d_tail);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (r) tot_stream += ms;
      cudaEventRecord(a);
      for (int k = 0; k < parts; k++) {
        int lo = k * per, cnt = std::min(n, lo + per) - lo;
        k_gather<8><<<(cnt + 2047) / 2048, 256>>>(cnt, d_idx + lo, d_eq, d_out);
      }
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      cudaEventElapsedTime(&ms, a, b);
      if (r) tot += ms;
    }
    printf("windows=%d (%5.1f MB each): gather %7.1f us per 10M entries (after a streaming pass of %6.1f us)\n", parts,
           80.0 / parts, tot / reps * 1000, tot_stream / reps * 1000);
  }
  CK(cudaGetLastError());
  return 0;
}
