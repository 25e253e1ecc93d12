// ntt_core.cuh -- register-resident negacyclic NTT / iNTT over one RNS residue.
//
// One CTA of T = N/32 threads transforms one residue polynomial (N = 2^n,
// 10 <= n <= 14; n = 15 is handled by the callers as two n = 14 sub-transforms
// with a twiddle "root prefix").  Every thread keeps 32 coefficients in
// registers for the whole transform -- the residue never lives in shared
// memory, which is used only as the exchange medium between three register
// passes:
//
//   pass A  stages 0..4     thread owns idx = (k << (n-5)) | tid        (strided)
//   pass B  stages 5..9     thread owns idx = (hi << (n-5)) | (k << (n-10)) | lo
//   pass C  stages 10..n-1  thread owns idx = (tid << 5) | k            (contiguous)
//
// Forward = Cooley-Tukey (natural in, bit-reversed out), inverse =
// Gentleman-Sande; semantics match seal::util::ntt_negacyclic_harvey /
// inverse_ntt_negacyclic_harvey as called below reference
// eva/seal/seal_executor.h:162-213,242: out[i] =
// a(psi^(2*bitrev(i)+1)), canonical in [0,p).
//
// Butterflies are lazy: primes are < 2^60, so a u64 holds values up to 16p.
// shoup_lazy() accepts any u64 and returns [0,2p); only the "X" input of a
// butterfly ever needs a conditional subtraction, and the unrolled code tracks
// a compile-time bound (in units of p) to place those subtractions sparsely.
//
// All phases are __host__ __device__: tests replay them on the CPU (threads
// serialised, shared memory as a plain array) to validate indexing, swizzles
// and bounds without a GPU.
#pragma once
#include "modarith.cuh"

#if defined(__CUDA_ARCH__)
#define EVAB_LDG(p) __ldg(p)
#else
#define EVAB_LDG(p) (*(p))
#endif

EVAB_HD u64x2 ldg_tw(const u64x2 *p) {
#if defined(__CUDA_ARCH__)
  ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
  u64x2 r; r.x = v.x; r.y = v.y; return r;
#else
  return *p;
#endif
}

template <int LOGN> struct NttGeom {
  static_assert(LOGN >= 10 && LOGN <= 14, "register NTT core supports 2^10..2^14");
  static constexpr int N = 1 << LOGN;
  static constexpr int T = N / 32;          // threads per residue
  static constexpr int LO = LOGN - 10;      // low index bits owned by tid in pass B
  static constexpr int SA = LOGN - 5;       // shift of k in layout A
  static constexpr int NC = LOGN - 10;      // stages in pass C
  static constexpr u32 MASK_AB = (16u >> LO) - 1u;
};

// ---------------------------------------------------------------------------
// shared-memory exchange layouts (element = 8 bytes)
// ---------------------------------------------------------------------------
// A<->B: linear with an XOR that spreads the rows a half-warp of pass-B readers
// touches over distinct banks when tid carries fewer than 4 "lo" bits.
template <int LOGN> EVAB_HD u32 swz_ab(u32 idx) {
  typedef NttGeom<LOGN> G;
  return idx ^ (((idx >> G::SA) & G::MASK_AB) << G::LO);
}
// B<->C: 256-byte rows (one per pass-C thread), 16-byte chunks XOR-swizzled by
// the row number so that 128-bit row reads of 8 consecutive threads hit 8
// distinct bank groups.
EVAB_HD u32 swz_bc(u32 idx) {
  u32 row = idx >> 5, e = idx & 31u;
  return (row << 5) | ((((e >> 1) ^ (row & 15u)) << 1) | (e & 1u));
}

template <int LOGN> EVAB_HD u32 idx_a(u32 tid, u32 k) { return (k << NttGeom<LOGN>::SA) | tid; }
template <int LOGN> EVAB_HD u32 idx_b(u32 tid, u32 k) {
  typedef NttGeom<LOGN> G;
  u32 hi = tid >> G::LO, lo = tid & ((1u << G::LO) - 1u);
  return (hi << G::SA) | (k << G::LO) | lo;
}
EVAB_HD u32 idx_c(u32 tid, u32 k) { return (tid << 5) | k; }

template <int LOGN> EVAB_HD void xchg_write_a(const u64 (&x)[32], u64 *sm, u32 tid) {
#pragma unroll
  for (int k = 0; k < 32; k++) sm[swz_ab<LOGN>(idx_a<LOGN>(tid, k))] = x[k];
}
template <int LOGN> EVAB_HD void xchg_read_a(u64 (&x)[32], const u64 *sm, u32 tid) {
#pragma unroll
  for (int k = 0; k < 32; k++) x[k] = sm[swz_ab<LOGN>(idx_a<LOGN>(tid, k))];
}
template <int LOGN> EVAB_HD void xchg_write_b_ab(const u64 (&x)[32], u64 *sm, u32 tid) {
#pragma unroll
  for (int k = 0; k < 32; k++) sm[swz_ab<LOGN>(idx_b<LOGN>(tid, k))] = x[k];
}
template <int LOGN> EVAB_HD void xchg_read_b_ab(u64 (&x)[32], const u64 *sm, u32 tid) {
#pragma unroll
  for (int k = 0; k < 32; k++) x[k] = sm[swz_ab<LOGN>(idx_b<LOGN>(tid, k))];
}
template <int LOGN> EVAB_HD void xchg_write_b_bc(const u64 (&x)[32], u64 *sm, u32 tid) {
#pragma unroll
  for (int k = 0; k < 32; k++) sm[swz_bc(idx_b<LOGN>(tid, k))] = x[k];
}
template <int LOGN> EVAB_HD void xchg_read_b_bc(u64 (&x)[32], const u64 *sm, u32 tid) {
#pragma unroll
  for (int k = 0; k < 32; k++) x[k] = sm[swz_bc(idx_b<LOGN>(tid, k))];
}
EVAB_HD void xchg_read_c(u64 (&x)[32], const u64 *sm, u32 tid) {
  const u64x2 *row = reinterpret_cast<const u64x2 *>(sm + ((size_t)tid << 5));
#pragma unroll
  for (int c = 0; c < 16; c++) {
    u64x2 v = row[c ^ (tid & 15u)];
    x[2 * c] = v.x; x[2 * c + 1] = v.y;
  }
}
EVAB_HD void xchg_write_c(const u64 (&x)[32], u64 *sm, u32 tid) {
  u64x2 *row = reinterpret_cast<u64x2 *>(sm + ((size_t)tid << 5));
#pragma unroll
  for (int c = 0; c < 16; c++) {
    u64x2 v; v.x = x[2 * c]; v.y = x[2 * c + 1];
    row[c ^ (tid & 15u)] = v;
  }
}

// ---------------------------------------------------------------------------
// forward (Cooley-Tukey) register passes.  `b` is the compile-time tracked
// upper bound of every live value in units of p (values < b*p <= 16p < 2^64).
// `root` is the twiddle root prefix: 1 for a full transform; 2+h for the h-th
// half of an n+1 transform (twiddle index = (root << s) + group).
// ---------------------------------------------------------------------------
EVAB_HD void ct_bfly(u64 &X, u64 &Y, const u64x2 w, u64 np, u64 two_p, bool fix, u64 eight_p) {
  u64 x = X;
  if (fix) x = csub(x, eight_p);
  u64 t = shoup_lazy_n(Y, w.x, w.y, np);
  X = x + t;
  Y = x - t + two_p;
}
// one forward stage over the 32 registers: pair distance d (in k), 16/d groups
// of twiddles starting at table index tw0 (consecutive).
template <int D> EVAB_HD void fwd_stage(u64 (&x)[32], const u64x2 *tw, u32 tw0, u64 p, int &b) {
  const u64 two_p = 2 * p, eight_p = 8 * p, np = 0 - p;
  const bool fix = b > 14;
  if (fix) b = 8;
#pragma unroll
  for (int g = 0; g < 16 / D; g++) {
    const u64x2 w = ldg_tw(tw + tw0 + g);
#pragma unroll
    for (int j = 0; j < D; j++) {
      const int k = g * 2 * D + j;
      ct_bfly(x[k], x[k + D], w, np, two_p, fix, eight_p);
    }
  }
  b += 2;
}
template <int LOGN> EVAB_HD void fwd_pass_a(u64 (&x)[32], const u64x2 *tw, u32 root, u64 p, int &b) {
  fwd_stage<16>(x, tw, (root << 0), p, b);
  fwd_stage<8>(x, tw, (root << 1), p, b);
  fwd_stage<4>(x, tw, (root << 2), p, b);
  fwd_stage<2>(x, tw, (root << 3), p, b);
  fwd_stage<1>(x, tw, (root << 4), p, b);
}
template <int LOGN> EVAB_HD void fwd_pass_b(u64 (&x)[32], const u64x2 *tw, u32 root, u64 p, u32 tid, int &b) {
  const u32 hi = tid >> NttGeom<LOGN>::LO;
  fwd_stage<16>(x, tw, (root << 5) + (hi << 0), p, b);
  fwd_stage<8>(x, tw, (root << 6) + (hi << 1), p, b);
  fwd_stage<4>(x, tw, (root << 7) + (hi << 2), p, b);
  fwd_stage<2>(x, tw, (root << 8) + (hi << 3), p, b);
  fwd_stage<1>(x, tw, (root << 9) + (hi << 4), p, b);
}
// pass C: thread owns 32 contiguous coefficients; stage s = 10+u pairs at
// distance 2^(NC-1-u) and uses 2^(5-NC+u+... ) consecutive twiddles.
template <int LOGN> EVAB_HD void fwd_pass_c(u64 (&x)[32], const u64x2 *tw, u32 root, u64 p, u32 tid, int &b) {
  constexpr int NC = NttGeom<LOGN>::NC;
  // stage 10+u: group index = ((tid<<5)|k) >> (NC-u)  =>  first = tid << (5-NC+u)
  if (NC >= 4) fwd_stage<8>(x, tw, (root << (LOGN - 4)) + (tid << 1), p, b);
  if (NC >= 3) fwd_stage<4>(x, tw, (root << (LOGN - 3)) + (tid << 2), p, b);
  if (NC >= 2) fwd_stage<2>(x, tw, (root << (LOGN - 2)) + (tid << 3), p, b);
  if (NC >= 1) fwd_stage<1>(x, tw, (root << (LOGN - 1)) + (tid << 4), p, b);
}
// reduce every register from < b*p to canonical [0,p)
EVAB_HD void canon(u64 (&x)[32], u64 p, int b) {
#pragma unroll
  for (int k = 0; k < 32; k++) {
    u64 v = x[k];
    if (b > 8) v = csub(v, 8 * p);
    if (b > 4) v = csub(v, 4 * p);
    if (b > 2) v = csub(v, 2 * p);
    if (b > 1) v = csub(v, p);
    x[k] = v;
  }
}

// ---------------------------------------------------------------------------
// inverse (Gentleman-Sande) register passes.  Values stay < 8p on stage entry.
// ---------------------------------------------------------------------------
template <int D> EVAB_HD void inv_stage(u64 (&x)[32], const u64x2 *tw, u32 tw0, u64 p, int &b) {
  // entry bound b in {1,2,4,8}; sums are reduced by 8p only once they could
  // reach 16p.
  const u64 eight_p = 8 * p, np = 0 - p;
  const u64 bias = (u64)b * p;
  const bool fix = b > 4;
#pragma unroll
  for (int g = 0; g < 16 / D; g++) {
    const u64x2 w = ldg_tw(tw + tw0 + g);
#pragma unroll
    for (int j = 0; j < D; j++) {
      const int k = g * 2 * D + j;
      u64 X = x[k], Y = x[k + D];
      u64 s = X + Y;
      if (fix) s = csub(s, eight_p);
      x[k] = s;
      x[k + D] = shoup_lazy_n(X - Y + bias, w.x, w.y, np);
    }
  }
  b = fix ? 8 : 2 * b;
  if (b < 2) b = 2;
}
template <int LOGN> EVAB_HD void inv_pass_c(u64 (&x)[32], const u64x2 *tw, u32 root, u64 p, u32 tid, int &b) {
  constexpr int NC = NttGeom<LOGN>::NC;
  if (NC >= 1) inv_stage<1>(x, tw, (root << (LOGN - 1)) + (tid << 4), p, b);
  if (NC >= 2) inv_stage<2>(x, tw, (root << (LOGN - 2)) + (tid << 3), p, b);
  if (NC >= 3) inv_stage<4>(x, tw, (root << (LOGN - 3)) + (tid << 2), p, b);
  if (NC >= 4) inv_stage<8>(x, tw, (root << (LOGN - 4)) + (tid << 1), p, b);
}
template <int LOGN> EVAB_HD void inv_pass_b(u64 (&x)[32], const u64x2 *tw, u32 root, u64 p, u32 tid, int &b) {
  const u32 hi = tid >> NttGeom<LOGN>::LO;
  inv_stage<1>(x, tw, (root << 9) + (hi << 4), p, b);
  inv_stage<2>(x, tw, (root << 8) + (hi << 3), p, b);
  inv_stage<4>(x, tw, (root << 7) + (hi << 2), p, b);
  inv_stage<8>(x, tw, (root << 6) + (hi << 1), p, b);
  inv_stage<16>(x, tw, (root << 5) + (hi << 0), p, b);
}
template <int LOGN> EVAB_HD void inv_pass_a(u64 (&x)[32], const u64x2 *tw, u32 root, u64 p, int &b) {
  inv_stage<1>(x, tw, (root << 4), p, b);
  inv_stage<2>(x, tw, (root << 3), p, b);
  inv_stage<4>(x, tw, (root << 2), p, b);
  inv_stage<8>(x, tw, (root << 1), p, b);
  inv_stage<16>(x, tw, (root << 0), p, b);
}
// multiply by a Shoup constant (e.g. N^-1) and canonicalise
EVAB_HD void scale_canon(u64 (&x)[32], u64 c, u64 cs, u64 p) {
#pragma unroll
  for (int k = 0; k < 32; k++) x[k] = csub(shoup_lazy(x[k], c, cs, p), p);
}
