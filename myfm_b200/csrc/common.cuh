// Shared host/device helpers of the myfm_b200 engine: error plumbing, device buffers, reductions.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace myfm {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define MYFM_CUDA(expr)                                                                            \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      std::ostringstream _ss;                                                                      \
      _ss << "CUDA error " << cudaGetErrorName(_e) << " (" << cudaGetErrorString(_e) << ") at "    \
          << __FILE__ << ":" << __LINE__ << " in " << #expr;                                       \
      throw ::myfm::CudaError(_ss.str());                                                          \
    }                                                                                              \
  } while (0)

// Owning device allocation (plain cudaMalloc; the engine allocates once at setup).
template <typename T> struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr, o.n = 0; }
  DevBuf &operator=(DevBuf &&o) noexcept {
    if (this != &o) {
      release();
      p = o.p, n = o.n;
      o.p = nullptr, o.n = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p)
      cudaFree(p);
    p = nullptr, n = 0;
  }
  void alloc(size_t count) {
    release();
    n = count;
    if (count)
      MYFM_CUDA(cudaMalloc(&p, count * sizeof(T)));
  }
  void upload(const T *src, size_t count, cudaStream_t s = nullptr) {
    if (count > n)
      alloc(count);
    if (count)
      MYFM_CUDA(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T> &v, cudaStream_t s = nullptr) { upload(v.data(), v.size(), s); }
  void download(T *dst, size_t count, cudaStream_t s = nullptr) const {
    if (count)
      MYFM_CUDA(cudaMemcpyAsync(dst, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
  void zero(cudaStream_t s = nullptr) {
    if (n)
      MYFM_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
};

// Pinned host staging buffer.
template <typename T> struct PinnedBuf {
  T *p = nullptr;
  size_t n = 0;
  PinnedBuf() = default;
  PinnedBuf(const PinnedBuf &) = delete;
  PinnedBuf &operator=(const PinnedBuf &) = delete;
  ~PinnedBuf() {
    if (p)
      cudaFreeHost(p);
  }
  void alloc(size_t count) {
    if (p)
      cudaFreeHost(p);
    p = nullptr;
    n = count;
    if (count)
      MYFM_CUDA(cudaMallocHost(&p, count * sizeof(T)));
  }
};

inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

#ifdef __CUDACC__
constexpr unsigned FULL_MASK = 0xffffffffu;

// Butterfly sum: every lane ends with the same, order-deterministic total.
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}

// Sum over a power-of-two sub-group of LANES consecutive lanes.
template <typename T, int LANES> __device__ __forceinline__ T subwarp_sum(T v) {
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1)
    v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}

// Block-wide sum broadcast to all threads; `scratch` holds >= 32 elements of T.  Two barriers.
template <typename T> __device__ __forceinline__ T block_sum(T v, T *scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads(); // scratch may still be read by a previous call
  if (lane == 0)
    scratch[wid] = v;
  __syncthreads();
  T t = (lane < nwarps) ? scratch[lane] : T(0);
  return warp_sum(t);
}
#endif

} // namespace myfm
