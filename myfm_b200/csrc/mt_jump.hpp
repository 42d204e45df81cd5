// Jump-ahead for std::mt19937 (host side): the polynomial that lets a lane of the parallel word
// generator (mt_device.cuh: k_mt_farm) skip over the words the other lanes produce.
//
// MT19937's transition is linear over GF(2) with a primitive characteristic polynomial phi of
// degree 19937, so every bit position of the output word sequence x[1], x[2], ... satisfies the
// linear recurrence phi (tempering is a linear bijection per word, so the tempered outputs satisfy
// it as well).  Hence for any N, with g(t) = t^N mod phi(t) = sum_i g_i t^i (deg g < 19937),
//
//     x[a + N + j] = XOR over { i : g_i = 1 } of x[a + i + j]        for all a + j >= 1,
//
// i.e. a window of the stream N words ahead is a GF(2) correlation of g with 19937 + 623
// consecutive words the generator has already produced (Haramoto, Matsumoto, Nishimura, Panneton,
// L'Ecuyer 2008 evaluate g(F) on the state by Horner; here the states F^i s are simply the words
// already lying in the ring buffer, so the evaluation is a data-parallel XOR reduction).
//
// phi is not hard-coded: it is recovered with Berlekamp-Massey from 2 * 19937 output bits of
// std::mt19937 itself, then t^N mod phi by square-and-multiply.  Results are cached per N.
#pragma once

#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <random>
#include <stdexcept>
#include <vector>

namespace myfm {

constexpr int MT_DEGREE = 19937;
constexpr int MT_JUMP_SPAN = MT_DEGREE + 623;         // words a jump reads: x[a .. a + 20559]
constexpr int MT_POLY_WORDS = (MT_DEGREE + 63) / 64;  // 312 x 64 bits hold a residue mod phi

namespace mtjump {

using Poly = std::vector<uint64_t>; // bit i of word i / 64 = coefficient of t^i

inline bool get_bit(const Poly &p, int i) { return (p[i >> 6] >> (i & 63)) & 1u; }
inline void flip_bit(Poly &p, int i) { p[i >> 6] ^= uint64_t(1) << (i & 63); }

// dst ^= src << shift (bit shift), both with n_words words; bits shifted beyond the end are dropped
inline void xor_shifted(uint64_t *dst, const uint64_t *src, int n_words, int shift) {
  const int ws = shift >> 6, bs = shift & 63;
  if (bs == 0) {
    for (int i = n_words - 1; i >= ws; i--)
      dst[i] ^= src[i - ws];
    return;
  }
  for (int i = n_words - 1; i > ws; i--)
    dst[i] ^= (src[i - ws] << bs) | (src[i - ws - 1] >> (64 - bs));
  dst[ws] ^= src[0] << bs;
}

// Berlekamp-Massey over GF(2): the shortest LFSR (connection polynomial C, C_0 = 1, length L) with
// s[n] = XOR_{i=1..L} C_i s[n-i].  Word-parallel: the discrepancy is the parity of C AND the
// reversed history window.
inline Poly berlekamp_massey(const std::vector<uint8_t> &s, int *length) {
  const int n = static_cast<int>(s.size());
  const int W = n / 64 + 2;
  Poly C(W, 0), B(W, 0), T(W), hist(W, 0); // hist bit i = s[k - i] while processing s[k]
  C[0] = B[0] = 1;
  int L = 0, m = 1;
  for (int k = 0; k < n; k++) {
    // shift the history by one and insert s[k] at bit 0
    for (int i = W - 1; i > 0; i--)
      hist[i] = (hist[i] << 1) | (hist[i - 1] >> 63);
    hist[0] = (hist[0] << 1) | (s[k] & 1u);
    uint64_t acc = 0;
    const int used = L / 64 + 1;
    for (int i = 0; i < used && i < W; i++)
      acc ^= C[i] & hist[i];
    if (__builtin_parityll(acc)) {
      if (2 * L <= k) {
        T = C;
        xor_shifted(C.data(), B.data(), W, m);
        L = k + 1 - L;
        B = T;
        m = 1;
      } else {
        xor_shifted(C.data(), B.data(), W, m);
        m++;
      }
    } else {
      m++;
    }
  }
  *length = L;
  return C;
}

// phi(t) = t^19937 * C(1/t): characteristic polynomial of the recurrence, as 19938 coefficients.
inline const Poly &characteristic_polynomial() {
  static Poly phi;
  static std::once_flag once;
  std::call_once(once, [] {
    std::mt19937 gen(4357u);
    std::vector<uint8_t> bits(2 * MT_DEGREE + 64);
    gen(); // the identity holds from the second output on (the low 31 bits of x[0] are not state)
    for (auto &b : bits)
      b = static_cast<uint8_t>(gen() & 1u);
    int L = 0;
    Poly C = berlekamp_massey(bits, &L);
    if (L != MT_DEGREE)
      throw std::logic_error("mt19937 jump: Berlekamp-Massey did not find a degree-19937 recurrence.");
    phi.assign(MT_POLY_WORDS + 1, 0);
    for (int i = 0; i <= MT_DEGREE; i++)
      if (get_bit(C, i))
        flip_bit(phi, MT_DEGREE - i);
  });
  return phi;
}

// r <- r mod phi for a polynomial of degree < 2 * 19937 held in 2 * MT_POLY_WORDS + 1 words
inline void reduce(std::vector<uint64_t> &r, const Poly &phi) {
  const int n_words = static_cast<int>(r.size());
  for (int k = 2 * MT_DEGREE - 2; k >= MT_DEGREE; k--)
    if ((r[k >> 6] >> (k & 63)) & 1u)
      xor_shifted(r.data(), phi.data(), std::min(n_words, (k >> 6) + 1), k - MT_DEGREE);
}

inline uint64_t spread_bits(uint32_t v) { // bit i -> bit 2 i
  uint64_t x = v;
  x = (x | (x << 16)) & 0x0000ffff0000ffffull;
  x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
  x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
  x = (x | (x << 2)) & 0x3333333333333333ull;
  x = (x | (x << 1)) & 0x5555555555555555ull;
  return x;
}

// t^n mod phi
inline Poly power_of_t(unsigned long long n) {
  const Poly &phi = characteristic_polynomial();
  std::vector<uint64_t> r(2 * MT_POLY_WORDS + 1, 0), sq(2 * MT_POLY_WORDS + 1);
  r[0] = 1;
  int top = 63;
  while (top > 0 && !((n >> top) & 1ull))
    top--;
  for (int b = top; b >= 0; b--) {
    // square: over GF(2) the cross terms vanish, the coefficients spread to the even positions
    std::fill(sq.begin(), sq.end(), 0);
    for (int i = 0; i < MT_POLY_WORDS; i++) {
      sq[2 * i] = spread_bits(static_cast<uint32_t>(r[i]));
      sq[2 * i + 1] = spread_bits(static_cast<uint32_t>(r[i] >> 32));
    }
    reduce(sq, phi);
    r.swap(sq);
    if ((n >> b) & 1ull) { // times t
      for (int i = MT_POLY_WORDS; i > 0; i--)
        r[i] = (r[i] << 1) | (r[i - 1] >> 63);
      r[0] <<= 1;
      if (get_bit(r, MT_DEGREE))
        xor_shifted(r.data(), phi.data(), MT_POLY_WORDS + 1, 0);
    }
  }
  r.resize(MT_POLY_WORDS);
  return r;
}

} // namespace mtjump

// The exponents i with g_i = 1 of g = t^n mod phi, ascending (about 10 000 of them); cached per n.
inline const std::vector<uint16_t> &mt_jump_taps(unsigned long long n) {
  static std::map<unsigned long long, std::vector<uint16_t>> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(n);
  if (it != cache.end())
    return it->second;
  mtjump::Poly g = mtjump::power_of_t(n);
  std::vector<uint16_t> taps;
  for (int i = 0; i < MT_DEGREE; i++)
    if (mtjump::get_bit(g, i))
      taps.push_back(static_cast<uint16_t>(i));
  return cache.emplace(n, std::move(taps)).first->second;
}

} // namespace myfm
