// Micro-benchmark: does the width of the per-lane loads matter for a streaming read-modify-write
// when only ONE 1024-thread CTA fits per SM (as in k_field_stream, whose table takes the shared memory)?
// Same bytes per row (8 B {e,q} read + 4 B index read + 8 B write), rows per lane per access: 1, 2, 4.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

template <int W, int U> // W rows per lane per access (1: 8B+4B loads, 2: 16B+8B, 4: 2x16B+16B), U accesses in flight
__global__ void __launch_bounds__(1024, 1) k(int n, float2 *__restrict__ eq, const int *__restrict__ tail, int pad) {
  extern __shared__ float tab[];
  for (int t = threadIdx.x; t < pad; t += 1024) tab[t] = 1e-9f * t;
  __syncthreads();
  const int stride = gridDim.x * 1024 * W;
  for (int base = (blockIdx.x * 1024 + threadIdx.x) * W; base < n; base += stride * U) {
    float2 v[U][W];
    int j[U][W];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int i = base + u * stride;
      if (i + W <= n) {
        if (W == 1) { v[u][0] = __ldcg(eq + i); j[u][0] = __ldcs(tail + i); }
        if (W == 2) { float4 x = __ldcg(reinterpret_cast<const float4 *>(eq + i)); int2 y = __ldcs(reinterpret_cast<const int2 *>(tail + i));
                      v[u][0] = make_float2(x.x, x.y); v[u][1] = make_float2(x.z, x.w); j[u][0] = y.x; j[u][1] = y.y; }
        if (W == 4) { float4 x0 = __ldcg(reinterpret_cast<const float4 *>(eq + i)), x1 = __ldcg(reinterpret_cast<const float4 *>(eq + i + 2));
                      int4 y = __ldcs(reinterpret_cast<const int4 *>(tail + i));
                      v[u][0] = make_float2(x0.x, x0.y); v[u][1] = make_float2(x0.z, x0.w); v[u][2] = make_float2(x1.x, x1.y); v[u][3] = make_float2(x1.z, x1.w);
                      j[u][0] = y.x; j[u][1] = y.y; j[u][2] = y.z; j[u][3] = y.w; }
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int i = base + u * stride;
      if (i + W <= n) {
#pragma unroll
        for (int w = 0; w < W; w++) { v[u][w].x += (v[u][w].y - tab[j[u][w]]) * 0.5f; v[u][w].y = tab[j[u][w]] + 1.f; }
        if (W == 1) __stcg(eq + i, v[u][0]);
        if (W == 2) __stcg(reinterpret_cast<float4 *>(eq + i), make_float4(v[u][0].x, v[u][0].y, v[u][1].x, v[u][1].y));
        if (W == 4) { __stcg(reinterpret_cast<float4 *>(eq + i), make_float4(v[u][0].x, v[u][0].y, v[u][1].x, v[u][1].y));
                      __stcg(reinterpret_cast<float4 *>(eq + i + 2), make_float4(v[u][2].x, v[u][2].y, v[u][3].x, v[u][3].y)); }
      }
    }
  }
}

template <int W, int U> void run(int n, float2 *eq, int *tail, int pad) {
  auto kern = k<W, U>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pad * 4));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  kern<<<148, 1024, pad * 4>>>(n, eq, tail, pad);
  cudaEventRecord(a);
  for (int r = 0; r < 10; r++) kern<<<148, 1024, pad * 4>>>(n, eq, tail, pad);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  CK(cudaGetLastError());
  printf("rows/lane/access %d, accesses in flight %d: %6.1f us  %6.0f GB/s\n", W, U, ms * 100, n * 20.0 / (ms / 10 * 1e-3) / 1e9);
}

int main() {
  const int n = 10000048, pad = 32768; // 128 KB of shared memory: one CTA per SM
  float2 *eq; int *tail;
  CK(cudaMalloc(&eq, (size_t)n * 8)); CK(cudaMalloc(&tail, (size_t)n * 4));
  CK(cudaMemset(eq, 0, (size_t)n * 8)); CK(cudaMemset(tail, 0, (size_t)n * 4));
  run<1, 1>(n, eq, tail, pad); run<1, 2>(n, eq, tail, pad); run<1, 4>(n, eq, tail, pad); run<1, 8>(n, eq, tail, pad);
  run<2, 1>(n, eq, tail, pad); run<2, 2>(n, eq, tail, pad); run<2, 4>(n, eq, tail, pad);
  run<4, 1>(n, eq, tail, pad); run<4, 2>(n, eq, tail, pad);
  return 0;
}
