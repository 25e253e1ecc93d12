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
  u64 itw1n, itw1n_s;      // itw[1] * N^-1 mod p (+Shoup): twiddle of the last inverse stage with the scaling folded in
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

// ---- hand-scheduled lazy Shoup product for the NTT butterflies --------------------------------
// B200 has no 64x64 multiplier; IMAD.WIDE.U32 / IMAD.HI.U32 hold the fma pipe for two cycles and
// that pipe bounds the transforms.  The quotient estimate therefore uses THREE partial products:
// q~ = a1*b1 + hi32(a1*b0) + hi32(a0*b1) drops a0*b0 and the low halves of the cross terms, so
// q - 2 <= q~ <= q = floor(ws*y / 2^64), and  w*y - q~*p  is congruent to w*y and lies in [0, 4p)
// for ANY u64 y (primes < 2^60 leave room for 16p).  shoup_mad4 returns  c + w*y + q~*np  (mod 2^64,
// np = 2^64 - p) with the addend c riding in the multiply-add chain: 3 wide + 2 high + 4 low
// multiplies and one 64-bit add.  Host and device evaluate the same integers.
EVAB_HD u64 shoup_mad4(u64 y, u64 w, u64 ws, u64 np, u64 c) {
  const u32 y0 = (u32)y, y1 = (u32)(y >> 32), w0 = (u32)w, w1 = (u32)(w >> 32);
  const u32 s0 = (u32)ws, s1 = (u32)(ws >> 32), n0 = (u32)np, n1 = (u32)(np >> 32);
#if defined(__CUDA_ARCH__)
  u32 c0, c1; u64 q, m;
  asm("mul.hi.u32 %0, %1, %2;" : "=r"(c0) : "r"(s1), "r"(y0));
  asm("mul.hi.u32 %0, %1, %2;" : "=r"(c1) : "r"(s0), "r"(y1));
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(q) : "r"(s1), "r"(y1), "l"((u64)c0));
  q += c1;
  const u32 q0 = (u32)q, q1 = (u32)(q >> 32);
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(m) : "r"(w0), "r"(y0), "l"(c));
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(m) : "r"(q0), "r"(n0), "l"(m));
  u32 h = (u32)(m >> 32);
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(h) : "r"(w0), "r"(y1));
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(h) : "r"(w1), "r"(y0));
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(h) : "r"(q0), "r"(n1));
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(h) : "r"(q1), "r"(n0));
  return ((u64)h << 32) | (u32)m;
#else
  const u64 q = (u64)s1 * y1 + (((u64)s1 * y0) >> 32) + (((u64)s0 * y1) >> 32);
  return c + w * y + q * np;
#endif
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
