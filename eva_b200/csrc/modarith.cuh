// modarith.cuh -- 64-bit modular arithmetic for RNS residues (primes < 2^60).
//
// All routines are __host__ __device__ so that the kernel phases built on top
// of them can be replayed on the CPU by tests/emu (index/bounds validation in
// a container without a GPU).  On the device everything lowers to 32-bit
// IMAD/IMAD.WIDE + IADD3 chains: B200 has no native 64x64 multiplier.
#pragma once
#include <stdint.h>

typedef uint64_t u64;
typedef uint32_t u32;

#if defined(__CUDACC__)
#define EVAB_HD __host__ __device__ __forceinline__
#else
#define EVAB_HD inline
#endif

#if !defined(__CUDACC__) && !defined(__align__)
#define __align__(n) __attribute__((aligned(n)))
#endif
struct __align__(16) u64x2 { u64 x, y; };

EVAB_HD u64 mulhi64(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}

// Per-prime constants resident in device memory (one entry per key-level prime)
struct PrimeDev {
  u64 p;
  u64 ratio_lo, ratio_hi;  // floor(2^128 / p)   (Barrett, 128-bit inputs)
  u64 ratio64;             // floor(2^64 / p)    (Barrett, 64-bit inputs)
  u64 ninv, ninv_s;        // N^-1 mod p and its Shoup companion
  const u64x2 *tw;         // forward twiddles  {psi^bitrev(i), shoup}   [N]
  const u64x2 *itw;        // inverse twiddles  {psi^-bitrev(i), shoup}  [N]
};

// x >= c ? x - c : x
EVAB_HD u64 csub(u64 x, u64 c) { return x >= c ? x - c : x; }

// Shoup multiplication by a constant w (ws = floor(w*2^64/p)); any u64 y.
// Result is congruent to w*y mod p and lies in [0, 2p).
EVAB_HD u64 shoup_lazy(u64 y, u64 w, u64 ws, u64 p) {
  u64 q = mulhi64(ws, y);
  return w * y - q * p;
}
EVAB_HD u64 shoup_mul(u64 y, u64 w, u64 ws, u64 p) { return csub(shoup_lazy(y, w, ws, p), p); }
// same with the caller supplying np = 2^64 - p: w*y + q*np (mod 2^64) saves the
// negation of q in the multiply-accumulate chain
EVAB_HD u64 shoup_lazy_n(u64 y, u64 w, u64 ws, u64 np) {
  u64 q = mulhi64(ws, y);
  return w * y + q * np;
}

// canonical add / sub / neg for operands already in [0,p)
EVAB_HD u64 addmod(u64 a, u64 b, u64 p) { return csub(a + b, p); }
EVAB_HD u64 submod(u64 a, u64 b, u64 p) { return a >= b ? a - b : a + p - b; }
EVAB_HD u64 negmod(u64 a, u64 p) { return a ? p - a : 0; }

// Barrett reduction of a 64-bit value to [0,p)
EVAB_HD u64 barrett64(u64 x, u64 p, u64 ratio64) {
  u64 q = mulhi64(x, ratio64);
  return csub(x - q * p, p);
}

// Barrett reduction of a 128-bit value (hi:lo) to [0,p); valid for
// hi:lo < 2^128 and p < 2^61 (one conditional subtraction suffices).
EVAB_HD u64 barrett128(u64 lo, u64 hi, u64 p, u64 rlo, u64 rhi) {
  u64 carry = mulhi64(lo, rlo);
  u64 t0 = lo * rhi, t1 = mulhi64(lo, rhi);
  u64 tmp1 = t0 + carry;
  u64 tmp3 = t1 + (tmp1 < carry);
  u64 u0 = hi * rlo, u1 = mulhi64(hi, rlo);
  u64 tmp1b = tmp1 + u0;
  carry = u1 + (tmp1b < tmp1);
  u64 q = hi * rhi + tmp3 + carry;
  return csub(lo - q * p, p);
}

// same for any 128-bit input (e.g. sums of products of lazily reduced operands): the
// quotient estimate is short by at most 3, two conditional subtractions finish
EVAB_HD u64 barrett128_wide(u64 lo, u64 hi, u64 p, u64 rlo, u64 rhi) {
  u64 carry = mulhi64(lo, rlo);
  u64 t0 = lo * rhi, t1 = mulhi64(lo, rhi);
  u64 tmp1 = t0 + carry;
  u64 tmp3 = t1 + (tmp1 < carry);
  u64 u0 = hi * rlo, u1 = mulhi64(hi, rlo);
  u64 tmp1b = tmp1 + u0;
  carry = u1 + (tmp1b < tmp1);
  u64 q = hi * rhi + tmp3 + carry;
  return csub(csub(lo - q * p, 2 * p), p);
}

// canonical product of two values in [0,p)
EVAB_HD u64 mulmod(u64 a, u64 b, u64 p, u64 rlo, u64 rhi) {
  return barrett128(a * b, mulhi64(a, b), p, rlo, rhi);
}

// 128-bit accumulate acc += a*b
EVAB_HD void mac128(u64 &lo, u64 &hi, u64 a, u64 b) {
  u64 pl = a * b, ph = mulhi64(a, b);
  lo += pl;
  hi += ph + (lo < pl);
}
