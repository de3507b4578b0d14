// "%.18e" of a float32 value, byte for byte what C printf / Python's '%.18e' % x print for the
// value widened to double — the text np.savetxt writes for the float32 polynomial matrix
// (scripts/drones_pols_generator.py:79-81; numpy's default fmt is '%.18e').  __host__ __device__ so
// tests/hostcheck can hold the very same code to Python's formatting on the CPU.
//
// A float32 is m * 2^e with an integer m < 2^24, so the 19 significant decimal digits come from
// exact integer arithmetic: D = round_half_even(m * 2^e * 10^p) with p chosen so that
// 10^18 <= D < 10^19.  m * 5^p needs at most 24 + 147 bits (p <= 63 for the smallest denormal):
// eight 32-bit limbs.
#pragma once
#include <stdint.h>

namespace mst {

struct Big256 { uint32_t w[8]; };

__host__ __device__ __forceinline__ void big_set(Big256& a, uint32_t v) {
  a.w[0] = v;
  for (int i = 1; i < 8; ++i) a.w[i] = 0u;
}
__host__ __device__ __forceinline__ void big_mul_small(Big256& a, uint32_t f) {
  uint64_t carry = 0;
  for (int i = 0; i < 8; ++i) {
    const uint64_t t = (uint64_t)a.w[i] * f + carry;
    a.w[i] = (uint32_t)t;
    carry = t >> 32;
  }
}
// a /= f (f < 2^32), returns the remainder
__host__ __device__ __forceinline__ uint32_t big_div_small(Big256& a, uint32_t f) {
  uint64_t rem = 0;
  for (int i = 7; i >= 0; --i) {
    const uint64_t t = (rem << 32) | a.w[i];
    a.w[i] = (uint32_t)(t / f);
    rem = t % f;
  }
  return (uint32_t)rem;
}
__host__ __device__ __forceinline__ void big_shl(Big256& a, int bits) {
  const int ws = bits >> 5, bs = bits & 31;
  for (int i = 7; i >= 0; --i) {
    uint32_t v = i - ws >= 0 ? a.w[i - ws] << bs : 0u;
    if (bs && i - ws - 1 >= 0) v |= a.w[i - ws - 1] >> (32 - bs);
    a.w[i] = v;
  }
}
__host__ __device__ __forceinline__ bool big_bit(const Big256& a, int bit) {
  return bit < 256 && ((a.w[bit >> 5] >> (bit & 31)) & 1u);
}
__host__ __device__ __forceinline__ bool big_any_below(const Big256& a, int bit) {  // any set bit below `bit`
  for (int i = 0; i < 8; ++i) {
    const int lo = 32 * i;
    if (lo >= bit) break;
    const uint32_t mask = bit - lo >= 32 ? 0xffffffffu : ((1u << (bit - lo)) - 1u);
    if (a.w[i] & mask) return true;
  }
  return false;
}
__host__ __device__ __forceinline__ void big_shr(Big256& a, int bits) {
  const int ws = bits >> 5, bs = bits & 31;
  for (int i = 0; i < 8; ++i) {
    uint32_t v = i + ws < 8 ? a.w[i + ws] >> bs : 0u;
    if (bs && i + ws + 1 < 8) v |= a.w[i + ws + 1] << (32 - bs);
    a.w[i] = v;
  }
}
__host__ __device__ __forceinline__ bool big_fits64(const Big256& a) {
  for (int i = 2; i < 8; ++i) if (a.w[i]) return false;
  return true;
}
__host__ __device__ __forceinline__ uint64_t big_low64(const Big256& a) { return ((uint64_t)a.w[1] << 32) | a.w[0]; }

// D = round_half_even(m * 2^e2 * 10^p); false when D does not fit 64 bits (p too large)
__host__ __device__ inline bool scaled_digits(uint32_t m, int e2, int p, uint64_t* D) {
  Big256 a;
  big_set(a, m);
  bool sticky = false, half = false;
  if (p >= 0) {
    int q = p;
    while (q >= 13) { big_mul_small(a, 1220703125u); q -= 13; }   // 5^13
    uint32_t f = 1u;
    for (int i = 0; i < q; ++i) f *= 5u;
    big_mul_small(a, f);
    const int sh = e2 + p;   // remaining power of two
    if (sh >= 0) {
      if (sh > 200) return false;
      big_shl(a, sh);
    } else {
      const int r = -sh;
      half = big_bit(a, r - 1);
      sticky = big_any_below(a, r - 1);
      if (r >= 256) { big_set(a, 0u); } else big_shr(a, r);
    }
  } else {
    // m * 2^e2 / 10^(-p): e2 > 0 here (the value is >= 10^19)
    if (e2 > 200) return false;
    big_shl(a, e2);
    int q = -p;
    uint32_t last = 0;
    while (q > 0) {   // divide by 10 one digit at a time, remembering what was cut off
      const uint32_t r = big_div_small(a, 10u);
      sticky = sticky || half || last != 0;
      // the digit cut off last is the most significant one of the remainder
      half = false;
      last = r;
      --q;
      if (q == 0) {
        // remainder as a fraction of 10: compare with one half using the last digit and what lies below it
        if (r > 5u || (r == 5u && sticky)) { half = true; sticky = true; }
        else if (r == 5u) { half = true; sticky = false; }
        else { half = false; sticky = sticky || r != 0; }
      }
    }
  }
  if (!big_fits64(a)) return false;
  uint64_t d = big_low64(a);
  if (half && (sticky || (d & 1ull))) {
    if (d == 0xffffffffffffffffull) return false;
    ++d;
  }
  *D = d;
  return true;
}

// writes the text of `v` at out (no terminator), returns its length (24 or 25; 3 / 4 for nan / inf)
__host__ __device__ inline int format_e18(float v, char* out) {
  union { float f; uint32_t u; } bits;
  bits.f = v;
  const uint32_t u = bits.u;
  int len = 0;
  const uint32_t ex = (u >> 23) & 0xffu, frac = u & 0x7fffffu;
  if (ex == 0xffu) {
    if (frac) { out[0] = 'n'; out[1] = 'a'; out[2] = 'n'; return 3; }   // Python prints 'nan' without a sign
    if (u >> 31) out[len++] = '-';
    out[len++] = 'i'; out[len++] = 'n'; out[len++] = 'f';
    return len;
  }
  if (u >> 31) out[len++] = '-';
  uint64_t D = 0;
  int k = 0;   // decimal exponent
  if (ex != 0u || frac != 0u) {
    const uint32_t m = ex ? (frac | 0x800000u) : frac;
    const int e2 = ex ? (int)ex - 150 : -149;
    // first guess of floor(log10 x) from the position of the leading bit, then corrected
    int top = 31;
    while (!((m >> top) & 1u)) --top;
    const int l2 = top + e2;   // floor(log2 x)
    k = (int)((l2 >= 0 ? (long long)l2 * 30103 : (long long)l2 * 30103 - 99999) / 100000);
    const uint64_t lo = 1000000000000000000ull, hi = 10000000000000000000ull;
    for (int it = 0; it < 4; ++it) {
      if (!scaled_digits(m, e2, 18 - k, &D) || D >= hi) {
        if (D == hi && scaled_digits(m, e2, 18 - k, &D) && D == hi) { D = lo; ++k; break; }   // rounded up to 10^19
        ++k;
        continue;
      }
      if (D < lo) { --k; continue; }
      break;
    }
  }
  // digits of D (19 of them; all zero for a zero value)
  char dg[19];
  uint64_t t = D;
  for (int i = 18; i >= 0; --i) { dg[i] = (char)('0' + (int)(t % 10ull)); t /= 10ull; }
  out[len++] = dg[0];
  out[len++] = '.';
  for (int i = 1; i < 19; ++i) out[len++] = dg[i];
  out[len++] = 'e';
  int ke = k;
  if (ke < 0) { out[len++] = '-'; ke = -ke; } else out[len++] = '+';
  out[len++] = (char)('0' + ke / 10);
  out[len++] = (char)('0' + ke % 10);
  return len;
}

}  // namespace mst
