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
// Split in two so that only the inherently serial part is serial:
//   k_mt_generate   ONE CTA extends the raw word stream into a ring buffer in global memory.
//                   x[n] = x[n-227] ^ twist(x[n-624], x[n-623]): 227 words per round are
//                   independent, one barrier per round, tempering fused into the store.
//   consumers       read the ring at a stream position kept on the device:
//                   k_mt_scalars (the few scalar draws of a sweep, one thread, a shared-memory
//                   window of the ring) and k_mt_bulk_count / _scan / _emit (the bulk normals,
//                   data parallel over all SMs, log/sqrt fused into the compaction).
// All of it runs on its own stream one sweep ahead of the sampler.
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
constexpr int MT_LAG = MT_N - MT_M;   // 227 words per round are independent of each other
constexpr int MT_WINDOW = 1024;       // circular window of the untempered recurrence (power of two)
constexpr int MT_GEN_THREADS = 256;
constexpr int MT_MAX_STAGES = 8;

// Stream positions count words from the start of the device stream ("window coordinates": the
// 624 state words handed over by the host are words 0 .. 623, of which the first `p` were already
// consumed there).  ring[n & mask] = tempered word n.
struct MtControl {
  uint32_t window[MT_WINDOW];       // untempered word n at window[n % MT_WINDOW], last 624 valid
  unsigned long long produced;      // words generated so far
  unsigned long long pos[MT_MAX_STAGES]; // consumer position after each stage of a sweep
  int error;                        // 1: the ring ran dry, 2: not enough accepted attempts
};

template <typename Real> struct MtTraits;
template <> struct MtTraits<float> { static constexpr int WPA = 2; };  // words per polar attempt
template <> struct MtTraits<double> { static constexpr int WPA = 4; };

__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v) {
  const uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
  return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__host__ __device__ __forceinline__ uint32_t mt_temper(uint32_t z) {
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

// One polar attempt of normal_distribution::operator() on the words w[0 .. WPA).
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

// ------------------------------------------------------------------------------------------------
// The serial part: extend the word stream until `ahead` words lie beyond the consumer position.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MT_GEN_THREADS)
    k_mt_generate(MtControl *ctl, uint32_t *__restrict__ ring, unsigned long long mask,
                  unsigned long long ahead, int final_stage) {
  __shared__ uint32_t W[MT_WINDOW];
  const int t = threadIdx.x;
  for (int i = t; i < MT_WINDOW; i += MT_GEN_THREADS)
    W[i] = ctl->window[i];
  const unsigned long long consumed = ctl->pos[final_stage];
  unsigned long long m = ctl->produced;
  __syncthreads();
  if (t == 0)
    ctl->pos[0] = consumed; // the next sweep starts where the last one stopped
  const unsigned long long goal = consumed + ahead;
  long long rounds = goal > m ? static_cast<long long>((goal - m + MT_LAG - 1) / MT_LAG) : 0;
  // thread t of round r produces word m + t; x[n-227] is the word this thread produced one round
  // earlier, the other two operands were produced at least two rounds earlier by other threads,
  // so one barrier per round orders everything.
  uint32_t prev = 0;
  if (t < MT_LAG)
    prev = W[(m + t - MT_LAG) & (MT_WINDOW - 1)];
  for (long long r = 0; r < rounds; r++) {
    if (t < MT_LAG) {
      const unsigned long long n = m + t;
      const uint32_t v = prev ^ mt_twist(W[(n - MT_N) & (MT_WINDOW - 1)], W[(n - MT_N + 1) & (MT_WINDOW - 1)]);
      W[n & (MT_WINDOW - 1)] = v;
      ring[n & mask] = mt_temper(v);
      prev = v;
    }
    m += MT_LAG;
    __syncthreads();
  }
  for (int i = t; i < MT_WINDOW; i += MT_GEN_THREADS)
    ctl->window[i] = W[i];
  if (t == 0)
    ctl->produced = m;
}

// ------------------------------------------------------------------------------------------------
// The same stream from P lanes at once (one CTA each).  The stream is cut into blocks of P * J
// words; lane k produces words [b P J + k J, b P J + (k + 1) J) of block b.  A lane reaches its next
// segment by JUMPING over the (P - 1) J words the other lanes produce: MT19937 is linear over
// GF(2), so with g = t^(Jtot) mod phi (mt_jump.hpp; Jtot = (P - 1) J + MT_JUMP_SPAN) the first 624
// words of the new segment are
//     x[start + j] = XOR over the taps i of g of x[start - Jtot + i + j],      j = 0 .. 623,
// a data-parallel XOR reduction over the last 19937 + 623 words of the lane's previous segment,
// which still lie in the ring (tempering is a linear bijection per word, so the tempered ring words
// are combined and the result is untempered for the recurrence).  The other J - 624 words of the
// segment follow from the recurrence as in k_mt_generate.  Block 0 comes from k_mt_generate (the
// host hands over one state, not P), every later block from this kernel; no lane ever waits for
// another one.  Bit-identical to the serial stream by construction; the parity tests compare it
// with libstdc++ draw for draw.
// ------------------------------------------------------------------------------------------------
constexpr int MT_FARM_THREADS = 640;   // 624 outputs of a jump, 227 words of a round
constexpr int MT_JUMP_SPAN_DEV = 19937 + 623;

struct MtFarm {
  int lanes;                        // P
  unsigned long long seg;           // J, at least MT_JUMP_SPAN_DEV + 1
  int n_taps;
  const unsigned short *taps;       // exponents of g, ascending
  unsigned long long *lane_blocks;  // [P] blocks every lane has completed (block 0: k_mt_generate)
};

__host__ __device__ __forceinline__ uint32_t mt_untemper(uint32_t z) {
  z ^= z >> 18;
  z ^= (z << 15) & 0xefc60000u;
  uint32_t t = z; // z = t ^ ((t << 7) & mask): recover 7 bits per step
  for (int i = 0; i < 4; i++)
    t = z ^ ((t << 7) & 0x9d2c5680u);
  z = t;
  t = z; // z = t ^ (t >> 11)
  t = z ^ (t >> 11);
  t = z ^ (t >> 11);
  return t;
}

__global__ void __launch_bounds__(MT_FARM_THREADS)
    k_mt_farm(MtControl *ctl, uint32_t *ring, unsigned long long mask, unsigned long long ahead, int final_stage,
              MtFarm f) {
  extern __shared__ __align__(8) unsigned short s_taps[];
  __shared__ uint32_t W[MT_WINDOW];
  const int t = threadIdx.x, lane_id = blockIdx.x;
  const unsigned long long consumed = ctl->pos[final_stage];
  const unsigned long long block_words = f.seg * static_cast<unsigned long long>(f.lanes);
  const unsigned long long target = (consumed + ahead + block_words - 1) / block_words; // blocks that must exist
  unsigned long long b = f.lane_blocks[lane_id];
  if (b < target)
    for (int i = t; i < f.n_taps; i += MT_FARM_THREADS)
      s_taps[i] = f.taps[i];
  __syncthreads();
  for (; b < target; b++) {
    const unsigned long long start = b * block_words + lane_id * f.seg;
    const unsigned long long src = start - (block_words - f.seg) - MT_JUMP_SPAN_DEV;
    if (t < MT_N) {
      uint32_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
      const unsigned long long base = src + t;
      int i = 0;
      for (; i + 4 <= f.n_taps; i += 4) { // four independent loads in flight
        const ushort4 tp = *reinterpret_cast<const ushort4 *>(s_taps + i);
        acc0 ^= ring[(base + tp.x) & mask];
        acc1 ^= ring[(base + tp.y) & mask];
        acc2 ^= ring[(base + tp.z) & mask];
        acc3 ^= ring[(base + tp.w) & mask];
      }
      for (; i < f.n_taps; i++)
        acc0 ^= ring[(base + s_taps[i]) & mask];
      const uint32_t tempered = (acc0 ^ acc1) ^ (acc2 ^ acc3);
      ring[(start + t) & mask] = tempered;
      W[(start + t) & (MT_WINDOW - 1)] = mt_untemper(tempered);
    }
    __syncthreads();
    unsigned long long m = start + MT_N;
    const unsigned long long end = start + f.seg;
    uint32_t prev = 0;
    if (t < MT_LAG)
      prev = W[(m + t - MT_LAG) & (MT_WINDOW - 1)];
    while (m < end) { // as k_mt_generate; the last round of a segment may be partial
      if (t < MT_LAG && m + t < end) {
        const unsigned long long n = m + t;
        const uint32_t v = prev ^ mt_twist(W[(n - MT_N) & (MT_WINDOW - 1)], W[(n - MT_N + 1) & (MT_WINDOW - 1)]);
        W[n & (MT_WINDOW - 1)] = v;
        ring[n & mask] = mt_temper(v);
        prev = v;
      }
      m += MT_LAG;
      __syncthreads();
    }
  }
  if (t == 0) {
    f.lane_blocks[lane_id] = b;
    if (lane_id == 0) {
      ctl->pos[0] = consumed; // the next sweep starts where the last one stopped
      ctl->produced = target * block_words;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Scalar draws.  One CTA; thread 0 interprets the stream out of a shared-memory window of the
// ring that all threads refill between draws (words beyond the window are read from the ring
// directly: a draw that needs more than MT_SCALAR_RESERVE words has probability < 1e-30).
// ------------------------------------------------------------------------------------------------
constexpr int MT_SCALAR_THREADS = 128;
constexpr int MT_SCALAR_WINDOW = 4096;
constexpr int MT_SCALAR_RESERVE = 256;

struct MtScalarSegment {
  int kind;          // 0 = fresh normal, 1 = gamma
  int count;
  long long out0;    // variate i of the segment goes to out[out0 + i]
  int shape_base, shape_mod; // gamma constants index: shape_base + (shape_mod ? i % shape_mod : 0)
};
struct MtScalarProgram {
  int n_segments;
  MtScalarSegment seg[4];
  const void *a1; // Real: _M_malpha - 1/3 per shape
  const void *a2; // Real: 1 / sqrt(9 a1)
};

template <typename Real> struct MtReader {
  const uint32_t *ring;
  unsigned long long mask, produced;
  const uint32_t *win;
  unsigned long long win_lo, win_hi;
  unsigned long long pos;
  int error = 0;

  __device__ uint32_t word(unsigned long long n) const {
    return (n >= win_lo && n < win_hi) ? win[n - win_lo] : ring[n & mask];
  }
  template <int NW> __device__ void take(uint32_t (&w)[NW]) {
    if (pos + NW > produced)
      error = 1;
#pragma unroll
    for (int i = 0; i < NW; i++)
      w[i] = word(pos + i);
    pos += NW;
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
    uint32_t w[MtTraits<Real>::WPA];
    do {
      take(w);
    } while (!mt_polar<Real>(w, x, y, r2) && !error);
    const Real mult = mt_polar_mult(r2);
    nd.saved = x * mult, nd.saved_available = true;
    return y * mult;
  }
  // std::gamma_distribution<Real>(shape >= 1, 1): Marsaglia-Tsang (bits/random.tcc:2355-2393)
  __device__ Real gamma(Real a1, Real a2) {
    Normal nd;
    Real u, v, n;
    uint32_t w[MtTraits<Real>::WPA / 2];
    for (;;) {
      do {
        n = normal(nd);
        v = Real(1.0) + a2 * n;
      } while (v <= 0.0 && !error);
      v = v * v * v;
      take(w);
      u = mt_canonical(w, Real());
      if (error)
        break;
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
};

template <typename Real>
__global__ void __launch_bounds__(MT_SCALAR_THREADS)
    k_mt_scalars(MtControl *ctl, const uint32_t *__restrict__ ring, unsigned long long mask, int stage_in,
                 int stage_out, MtScalarProgram prog, Real *__restrict__ out) {
  __shared__ uint32_t win[MT_SCALAR_WINDOW];
  __shared__ unsigned long long s_pos;
  __shared__ int s_seg, s_i, s_done;
  if (threadIdx.x == 0)
    s_pos = ctl->pos[stage_in], s_seg = 0, s_i = 0, s_done = 0;
  __syncthreads();
  const Real *a1 = static_cast<const Real *>(prog.a1), *a2 = static_cast<const Real *>(prog.a2);
  while (!s_done) {
    const unsigned long long lo = s_pos;
    for (int i = threadIdx.x; i < MT_SCALAR_WINDOW; i += MT_SCALAR_THREADS)
      win[i] = ring[(lo + i) & mask];
    __syncthreads();
    if (threadIdx.x == 0) {
      MtReader<Real> rd;
      rd.ring = ring, rd.mask = mask, rd.produced = ctl->produced;
      rd.win = win, rd.win_lo = lo, rd.win_hi = lo + MT_SCALAR_WINDOW, rd.pos = lo;
      int seg = s_seg, i = s_i;
      while (seg < prog.n_segments && rd.pos + MT_SCALAR_RESERVE <= rd.win_hi) {
        const MtScalarSegment &sg = prog.seg[seg];
        if (i >= sg.count) {
          seg++, i = 0;
          continue;
        }
        Real v;
        if (sg.kind == 0) {
          typename MtReader<Real>::Normal nd;
          v = rd.normal(nd);
        } else {
          const int sh = sg.shape_base + (sg.shape_mod ? i % sg.shape_mod : 0);
          v = rd.gamma(a1[sh], a2[sh]);
        }
        out[sg.out0 + i] = v;
        i++;
      }
      while (seg < prog.n_segments && i >= prog.seg[seg].count)
        seg++, i = 0;
      s_seg = seg, s_i = i, s_pos = rd.pos;
      if (rd.error)
        ctl->error = 1, s_done = 1;
      if (seg >= prog.n_segments) {
        s_done = 1;
        ctl->pos[stage_out] = rd.pos;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Bulk: `count` fresh normals in a row = the first `count` accepted polar attempts from the
// position of stage_in.  Thread t of CTA c owns MT_BULK_PER_THREAD consecutive attempts.
// ------------------------------------------------------------------------------------------------
constexpr int MT_BULK_THREADS = 256;
constexpr int MT_BULK_PER_THREAD = 8;
constexpr int MT_BULK_TILE = MT_BULK_THREADS * MT_BULK_PER_THREAD;

template <typename Real> struct MtBulkTile {
  Real y[MT_BULK_PER_THREAD], r2[MT_BULK_PER_THREAD];
  unsigned accept = 0; // bit u: attempt u of this thread accepted
  int n = 0;

  __device__ __forceinline__ void evaluate(const uint32_t *__restrict__ ring, unsigned long long mask,
                                           unsigned long long pos, long long first_attempt, long long n_attempts) {
    constexpr int WPA = MtTraits<Real>::WPA;
#pragma unroll
    for (int u = 0; u < MT_BULK_PER_THREAD; u++) {
      const long long a = first_attempt + u;
      y[u] = 0, r2[u] = 1;
      if (a < n_attempts) {
        uint32_t w[WPA];
        const unsigned long long p = pos + static_cast<unsigned long long>(a) * WPA;
#pragma unroll
        for (int k = 0; k < WPA; k++)
          w[k] = ring[(p + k) & mask];
        Real x;
        if (mt_polar<Real>(w, x, y[u], r2[u]))
          accept |= 1u << u, n++;
      }
    }
  }
};

template <typename Real>
__global__ void __launch_bounds__(MT_BULK_THREADS)
    k_mt_bulk_count(MtControl *ctl, const uint32_t *__restrict__ ring, unsigned long long mask, int stage_in,
                    long long n_attempts, int *__restrict__ tile_count) {
  __shared__ int scratch[32];
  const unsigned long long pos = ctl->pos[stage_in];
  if (blockIdx.x == 0 && threadIdx.x == 0 &&
      pos + static_cast<unsigned long long>(n_attempts) * MtTraits<Real>::WPA > ctl->produced)
    ctl->error = 1;
  MtBulkTile<Real> tile;
  tile.evaluate(ring, mask, pos,
                static_cast<long long>(blockIdx.x) * MT_BULK_TILE + threadIdx.x * MT_BULK_PER_THREAD, n_attempts);
  const int total = block_sum(tile.n, scratch);
  if (threadIdx.x == 0)
    tile_count[blockIdx.x] = total;
}

// Exclusive prefix sum of the tile counts (one CTA; a few thousand tiles at most per launch).
__global__ void __launch_bounds__(1024) k_mt_bulk_scan(int n_tiles, const int *__restrict__ tile_count,
                                                       long long *__restrict__ tile_offset) {
  __shared__ long long warp_tot[32];
  __shared__ long long carry;
  if (threadIdx.x == 0)
    carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < n_tiles; base += 1024) {
    const int i = base + threadIdx.x;
    const long long v = i < n_tiles ? tile_count[i] : 0;
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long up = __shfl_up_sync(FULL_MASK, incl, o);
      if (lane >= o)
        incl += up;
    }
    if (lane == 31)
      warp_tot[wid] = incl;
    __syncthreads();
    long long before = carry;
    for (int w = 0; w < wid; w++)
      before += warp_tot[w];
    if (i < n_tiles)
      tile_offset[i] = before + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023)
      carry = before + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0)
    tile_offset[n_tiles] = carry;
}

// z[out0 + k] = y * sqrt(-2 log(r2) / r2) of the k-th accepted attempt, k < count
// (normal_distribution::operator(), bits/random.tcc:1836-1839); the attempt that yields the last
// one fixes the stream position of stage_out.
template <typename Real>
__global__ void __launch_bounds__(MT_BULK_THREADS)
    k_mt_bulk_emit(MtControl *ctl, const uint32_t *__restrict__ ring, unsigned long long mask, int stage_in,
                   int stage_out, long long n_attempts, long long count, const long long *__restrict__ tile_offset,
                   int n_tiles, Real *__restrict__ out, long long out0) {
  __shared__ int warp_tot[32];
  const unsigned long long pos = ctl->pos[stage_in];
  const long long first = static_cast<long long>(blockIdx.x) * MT_BULK_TILE + threadIdx.x * MT_BULK_PER_THREAD;
  if (blockIdx.x == 0 && threadIdx.x == 0 && tile_offset[n_tiles] < count)
    ctl->error = 2;
  if (tile_offset[blockIdx.x] >= count)
    return;
  MtBulkTile<Real> tile;
  tile.evaluate(ring, mask, pos, first, n_attempts);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = tile.n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(FULL_MASK, incl, o);
    if (lane >= o)
      incl += up;
  }
  if (lane == 31)
    warp_tot[wid] = incl;
  __syncthreads();
  long long k = tile_offset[blockIdx.x] + incl - tile.n;
  for (int w = 0; w < wid; w++)
    k += warp_tot[w];
#pragma unroll
  for (int u = 0; u < MT_BULK_PER_THREAD; u++)
    if ((tile.accept >> u) & 1u) {
      if (k < count) {
        out[out0 + k] = tile.y[u] * mt_polar_mult(tile.r2[u]);
        if (k == count - 1)
          ctl->pos[stage_out] = pos + static_cast<unsigned long long>(first + u + 1) * MtTraits<Real>::WPA;
      }
      k++;
    }
}

} // namespace myfm
