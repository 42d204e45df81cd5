// Device half of the RNG contract (MYFM_RNG_MT19937, regression): the libstdc++ std::mt19937
// stream and the libstdc++ normal / gamma distributions, reproduced draw for draw on the GPU so
// that a whole sweep's standardised variates are produced without the host.
//
// Why this is possible (rng.hpp has the contract): for regression the reference consumes the
// stream in a data-independent pattern — every Gaussian draw is a FRESH
// std::normal_distribution (FMTrainer.hpp:122-125: one Marsaglia-polar round, second variate
// dropped) and every Gamma draw a fresh std::gamma_distribution whose shape does not depend on
// the data (FMTrainer.hpp:140-143,157-165).  The bulk of a sweep is (1 + K) * dim_all fresh
// normals in a row; each polar attempt consumes a fixed number of engine words (2 for float, 4
// for double; bits/random.tcc generate_canonical), so "the k-th normal" is "the k-th accepted
// attempt": an acceptance flag per attempt, a prefix sum, and a compaction.
//
// One CTA runs the whole program: it regenerates MT19937 blocks of 624 words in shared memory
// (3 barriers per block, tempering fused), lets thread 0 interpret the few scalar draws of a
// sweep, and compacts the accepted attempts of the bulk normals cooperatively; the log/sqrt of
// the polar method then runs as an ordinary data-parallel kernel.  Both run on their own stream
// one sweep ahead of the sampler.
//
// Arithmetic follows bits/random.tcc (normal_distribution::operator() :1811-1846,
// gamma_distribution::operator() :2355-2393, generate_canonical :3349-3385) including the
// float/double promotions caused by the double literals there.  Acceptance decisions of the
// normals involve no transcendental function, so the stream position is exact; log() differs from
// glibc's by at most an ulp in the VALUE of a variate (float: computed in double and rounded).
#pragma once

#include "common.cuh"

#include <cstdint>

namespace myfm {

constexpr int MT_N = 624, MT_M = 397;
constexpr int MT_THREADS = 256;
constexpr int MT_BUF_BLOCKS = 10;                 // shared word buffer capacity, in MT blocks
constexpr int MT_BUF_WORDS = MT_BUF_BLOCKS * MT_N;

// Persistent generator state in global memory: x[624] untempered words + p (next word index,
// 624 = regenerate first) — the layout std::mt19937 streams out with operator<<.
struct MtDeviceState {
  uint32_t x[MT_N];
  uint32_t p;
};

// Where the standardised variates of one regression sweep go (rng.hpp: SweepLayout) and the
// shape-dependent gamma constants, computed on the host exactly as libstdc++ does.
struct MtProgram {
  long long g_alpha, z_w0, g_lw, z_mw, z_w, g_lV, z_mV, z_V; // offsets, -1 = not drawn
  int G, K;
  long long dim_all;
  const void *a1;  // [G + 1] Real: _M_malpha - 1/3 per group, then alpha's
  const void *a2;  // [G + 1] Real: 1 / sqrt(9 a1)
};

template <typename Real> struct MtTraits;
template <> struct MtTraits<float> { static constexpr int WPA = 2; };  // words per polar attempt
template <> struct MtTraits<double> { static constexpr int WPA = 4; };

__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v) {
  const uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
  return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t z) {
  z ^= (z >> 11);
  z ^= (z << 7) & 0x9d2c5680u;
  z ^= (z << 15) & 0xefc60000u;
  z ^= (z >> 18);
  return z;
}

// generate_canonical<float, 24>: one word (bits/random.tcc:3349-3385)
__device__ __forceinline__ float mt_canonical(const uint32_t *w, float) {
  float r = __uint2float_rn(w[0]) * 2.3283064365386963e-10f; // / 2^32, exact
  return r >= 1.0f ? 0.99999994f : r; // nextafter(1, 0)
}
// generate_canonical<double, 53>: two words, low word first
__device__ __forceinline__ double mt_canonical(const uint32_t *w, double) {
  double sum = static_cast<double>(w[0]);
  sum += static_cast<double>(w[1]) * 4294967296.0;
  double r = sum * 5.42101086242752217e-20; // / 2^64, exact
  return r >= 1.0 ? 0.99999999999999988898 : r;
}
__device__ __forceinline__ float mt_log(float x) { return static_cast<float>(log(static_cast<double>(x))); }
__device__ __forceinline__ double mt_log(double x) { return log(x); }

// One polar attempt of normal_distribution::operator() starting at words w[0 .. WPA).
// Returns true when accepted; x, y, r2 as in the reference.
template <typename Real>
__device__ __forceinline__ bool mt_polar(const uint32_t *w, Real &x, Real &y, Real &r2) {
  constexpr int H = MtTraits<Real>::WPA / 2;
  x = Real(2.0) * mt_canonical(w, Real()) - Real(1.0);
  y = Real(2.0) * mt_canonical(w + H, Real()) - Real(1.0);
  r2 = x * x + y * y;
  return !(r2 > Real(1.0) || r2 == Real(0.0));
}
template <typename Real> __device__ __forceinline__ Real mt_polar_mult(Real r2) {
  return sqrt(-2 * mt_log(r2) / r2);
}

template <typename Real> struct MtCta {
  uint32_t *cur, *prev; // [624] each: untempered state of the last two generated blocks
  uint32_t *buf;        // [MT_BUF_WORDS] tempered words not yet consumed: buf[head .. head+avail)
  int *ctl;             // [0] head, [1] avail (scalar-phase hand-off), [3] error, [4..5] round scratch
  int *warp_tot;        // [32]
  // Uniform per-thread copies: every thread tracks the buffer window itself, so appending a block
  // needs no bookkeeping barrier.  Thread 0 alone moves them during a scalar draw and publishes
  // the result through ctl[0..1].
  int head = 0, avail = 0, last = 0; // last = size of the most recently appended block

  // Moves the unconsumed words to the front of the buffer.  All threads; ends with a barrier.
  __device__ void compact() {
    if (head == 0)
      return;
    const int h = head, n = avail;
    // forward copy in chunks of blockDim: a chunk is read completely before it is written
    for (int base = 0; base < n; base += blockDim.x) {
      const int i = base + threadIdx.x;
      uint32_t v = 0;
      if (i < n)
        v = buf[h + i];
      __syncthreads();
      if (i < n)
        buf[i] = v;
    }
    head = 0;
    __syncthreads();
  }

  // Appends the rest of the state block the kernel starts with.
  __device__ void append_initial(uint32_t p) {
    const int n = MT_N - static_cast<int>(p);
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      buf[head + avail + i] = mt_temper(cur[p + i]);
    avail += n, last = n;
    __syncthreads();
  }
  // Regenerates the next block of 624 words (three barriers) and appends its tempered words.
  __device__ void append_block() {
    uint32_t *x = cur, *y = prev; // y becomes the new block; the roles swap at the end
    const int t = threadIdx.x;
    uint32_t *out = buf + head + avail;
    if (t < MT_N - MT_M) { // [0, 227)
      uint32_t v = x[t + MT_M] ^ mt_twist(x[t], x[t + 1]);
      y[t] = v, out[t] = mt_temper(v);
    }
    __syncthreads();
    if (t < MT_N - MT_M) { // [227, 454)
      const int i = t + (MT_N - MT_M);
      uint32_t v = y[i - (MT_N - MT_M)] ^ mt_twist(x[i], x[i + 1]);
      y[i] = v, out[i] = mt_temper(v);
    }
    __syncthreads();
    {
      const int i = t + 2 * (MT_N - MT_M); // [454, 624)
      if (i < MT_N - 1) {
        uint32_t v = y[i - (MT_N - MT_M)] ^ mt_twist(x[i], x[i + 1]);
        y[i] = v, out[i] = mt_temper(v);
      } else if (i == MT_N - 1) {
        uint32_t v = y[MT_M - 1] ^ mt_twist(x[MT_N - 1], y[0]);
        y[i] = v, out[i] = mt_temper(v);
      }
    }
    __syncthreads();
    avail += MT_N, last = MT_N;
    cur = y, prev = x;
  }
  // Makes at least `need` words available (need <= 624).  Uniform across the CTA.
  __device__ void ensure(int need) {
    if (avail >= need)
      return;
    if (head + avail + MT_N > MT_BUF_WORDS)
      compact();
    append_block();
  }
  // Fills the buffer to capacity (bulk phase).
  __device__ void fill() {
    compact();
    while (avail + MT_N <= MT_BUF_WORDS)
      append_block();
  }

  // ---- scalar draws (thread 0 only; at least 624 words available) ------------------------------
  __device__ bool take(int n, const uint32_t *&w) {
    if (avail < n) {
      ctl[3] = 1; // ran dry inside one draw: p < 1e-90, reported to the host
      w = buf;
      return false;
    }
    w = buf + head;
    head += n, avail -= n;
    return true;
  }
  struct Normal { // std::normal_distribution<Real>(0, 1)
    bool saved_available = false;
    Real saved = 0;
  };
  __device__ Real normal(Normal &nd) {
    if (nd.saved_available) {
      nd.saved_available = false;
      return nd.saved;
    }
    Real x, y, r2;
    const uint32_t *w;
    do {
      if (!take(MtTraits<Real>::WPA, w))
        return 0;
    } while (!mt_polar<Real>(w, x, y, r2));
    const Real mult = mt_polar_mult(r2);
    nd.saved = x * mult, nd.saved_available = true;
    return y * mult;
  }
  // std::gamma_distribution<Real>(shape >= 1, 1): Marsaglia-Tsang (bits/random.tcc:2355-2393)
  __device__ Real gamma(Real a1, Real a2) {
    Normal nd;
    Real u, v, n;
    const uint32_t *w;
    for (;;) {
      do {
        n = normal(nd);
        v = Real(1.0) + a2 * n;
      } while (v <= 0.0 && !ctl[3]);
      v = v * v * v;
      if (!take(MtTraits<Real>::WPA / 2, w))
        return 0;
      u = mt_canonical(w, Real());
      // the literals below are double in the reference: these expressions are double arithmetic
      const double nn = static_cast<double>(n), vv = static_cast<double>(v);
      if (!(static_cast<double>(u) > static_cast<double>(Real(1.0)) - 0.0331 * nn * nn * nn * nn))
        break;
      const double rhs = 0.5 * nn * nn + static_cast<double>(a1) * (1.0 - vv + static_cast<double>(mt_log(v)));
      if (!(static_cast<double>(mt_log(u)) > rhs))
        break;
    }
    return a1 * v;
  }
  // One scalar draw by thread 0 (kind 0 = fresh normal, 1 = gamma); all threads call.
  __device__ void scalar(int kind, Real a1, Real a2, Real *dst) {
    ensure(MT_N);
    if (threadIdx.x == 0) {
      Normal nd;
      *dst = kind == 0 ? normal(nd) : gamma(a1, a2);
      ctl[0] = head, ctl[1] = avail;
    }
    __syncthreads();
    head = ctl[0], avail = ctl[1];
    __syncthreads(); // ctl is rewritten by the next draw
  }

  // ---- bulk: `count` fresh normals in a row -----------------------------------------------------
  // Writes, for the k-th accepted polar attempt, the pair (y, r2) to raw[k]; the transform
  // z = y * sqrt(-2 log(r2) / r2) runs afterwards over the whole array on all SMs
  // (k_mt_finish_normals) — log() is the expensive part and has no business in a one-CTA kernel.
  // Each warp owns a contiguous range of the buffered attempts, its lanes stride over it; one
  // ballot per 32 attempts gives the compaction offsets, one barrier per round the warp bases.
  __device__ void bulk_normals(Real *raw, long long count) {
    constexpr int WPA = MtTraits<Real>::WPA;
    constexpr int NWARPS = MT_THREADS / 32;
    constexpr int MAX_ATTEMPTS = MT_BUF_WORDS / WPA;
    constexpr int KMAX = ((MAX_ATTEMPTS + NWARPS - 1) / NWARPS + 31) / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    int round = 0;
    while (count > 0) {
      if (count >= MAX_ATTEMPTS)
        fill(); // every attempt of the buffer will be consumed
      else
        ensure(WPA); // tail: one block at a time, so that the state can be handed over exactly
      int *used_attempts = ctl + 4 + (round++ & 1);
      const int A = avail / WPA;
      const int h = head;
      const int S = (A + NWARPS - 1) / NWARPS;
      const int wa0 = min(A, wid * S), wa1 = min(A, wa0 + S);
      unsigned masks[KMAX];
      Real ys[KMAX], r2s[KMAX];
      int cnt = 0;
#pragma unroll
      for (int k = 0; k < KMAX; k++) {
        const int a = wa0 + k * 32 + lane;
        Real x;
        ys[k] = 0, r2s[k] = 0;
        bool ok = false;
        if (a < wa1)
          ok = mt_polar<Real>(buf + h + a * WPA, x, ys[k], r2s[k]);
        masks[k] = __ballot_sync(FULL_MASK, ok);
        cnt += __popc(masks[k]);
      }
      if (lane == 0)
        warp_tot[wid] = cnt;
      if (threadIdx.x == 0)
        *used_attempts = A; // lowered below when the segment ends inside this round
      __syncthreads();
      int off = 0, total = 0;
#pragma unroll
      for (int k = 0; k < NWARPS; k++) {
        const int v = warp_tot[k];
        if (k < wid)
          off += v;
        total += v;
      }
      const long long want = count < total ? count : total; // normals taken from this round
      const bool completes = count <= total;                // the segment ends inside this round
#pragma unroll
      for (int k = 0; k < KMAX; k++) {
        const unsigned m = masks[k];
        if ((m >> lane) & 1u) {
          const int idx = off + __popc(m & lt_mask);
          if (idx < want) {
            raw[2 * static_cast<long long>(idx)] = ys[k];
            raw[2 * static_cast<long long>(idx) + 1] = r2s[k];
          }
          if (completes && idx + 1 == want)
            *used_attempts = wa0 + k * 32 + lane + 1; // produced the segment's last normal
        }
        off += __popc(m);
      }
      __syncthreads();
      const int used = *used_attempts * WPA;
      head += used, avail -= used;
      raw += 2 * want, count -= want;
    }
  }
};

// z[i] = y * sqrt(-2 log(r2) / r2) for the bulk segments (normal_distribution::operator(),
// bits/random.tcc:1836-1839); raw holds (y, r2) per variate, indexed like z.
template <typename Real>
__global__ void __launch_bounds__(256)
    k_mt_finish_normals(const Real *__restrict__ raw, Real *__restrict__ z, long long begin0,
                        long long n0, long long begin1, long long n1) {
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n0 + n1)
    return;
  i = i < n0 ? begin0 + i : begin1 + (i - n0);
  const Real y = raw[2 * i], r2 = raw[2 * i + 1];
  z[i] = y * mt_polar_mult(r2);
}

// The variates of one regression sweep in the reference's consumption order
// (BaseFMTrainer.hpp:135-152; rng.hpp: SweepLayout).
template <typename Real>
__global__ void __launch_bounds__(MT_THREADS)
    k_mt_sweep_variates(MtDeviceState *state, MtProgram prog, Real *__restrict__ out,
                        Real *__restrict__ raw, int *error) {
  __shared__ uint32_t s_state[2][MT_N];
  __shared__ uint32_t s_buf[MT_BUF_WORDS];
  __shared__ int s_ctl[8];
  __shared__ int s_warp[32];
  MtCta<Real> c;
  c.cur = s_state[0], c.prev = s_state[1], c.buf = s_buf, c.ctl = s_ctl, c.warp_tot = s_warp;
  for (int i = threadIdx.x; i < MT_N; i += blockDim.x)
    c.cur[i] = state->x[i];
  if (threadIdx.x < 8)
    s_ctl[threadIdx.x] = 0;
  __syncthreads();
  c.append_initial(state->p);

  const Real *a1 = static_cast<const Real *>(prog.a1), *a2 = static_cast<const Real *>(prog.a2);
  const int G = prog.G, K = prog.K;
  if (prog.g_alpha >= 0)
    c.scalar(1, a1[G], a2[G], out + prog.g_alpha);
  if (prog.z_w0 >= 0)
    c.scalar(0, 0, 0, out + prog.z_w0);
  for (int g = 0; g < G; g++)
    c.scalar(1, a1[g], a2[g], out + prog.g_lw + g);
  for (int g = 0; g < G; g++)
    c.scalar(0, 0, 0, out + prog.z_mw + g);
  if (prog.z_w >= 0)
    c.bulk_normals(raw + 2 * prog.z_w, prog.dim_all);
  for (int r = 0; r < K; r++)
    for (int g = 0; g < G; g++)
      c.scalar(1, a1[g], a2[g], out + prog.g_lV + static_cast<long long>(r) * G + g);
  for (long long i = 0; i < static_cast<long long>(K) * G; i++)
    c.scalar(0, 0, 0, out + prog.z_mV + i);
  c.bulk_normals(raw + 2 * prog.z_V, static_cast<long long>(K) * prog.dim_all);

  // hand the generator over: the unconsumed words are the tail of [prev block][cur block]
  __syncthreads();
  const int left = c.avail, last = c.last;
  const uint32_t *src = left <= last ? c.cur : c.prev;
  const uint32_t p = left <= last ? MT_N - left : 2 * MT_N - left;
  for (int i = threadIdx.x; i < MT_N; i += blockDim.x)
    state->x[i] = src[i];
  if (threadIdx.x == 0) {
    state->p = p;
    if (s_ctl[3] || left > last + MT_N)
      *error = 1;
  }
}

} // namespace myfm
