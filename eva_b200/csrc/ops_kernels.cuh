// ops_kernels.cuh -- element-wise (dyadic) ciphertext kernels, fused sums of
// products and the key-switch inner product, as __host__ __device__ bodies
// (two coefficients per call, 128-bit accesses) shared by the CUDA kernels and
// the CPU emulator.  Semantics: SURVEY.md Appendix A.4 / A.5 / A.8.
#pragma once
#include "ntt_core.cuh"

enum { DY_ADD = 0, DY_SUB = 1, DY_NEG = 2, DY_MULPT = 3, DY_COPY = 4 };

struct DyArgs {
  u64 *out; const u64 *a; const u64 *b;
  const PrimeDev *primes;
  int ell, N;
  int sa, sb, sout;  // polys in a / b / out
  int b_is_plain;    // b = plaintext [ell][N]: add/sub touch poly 0 only, mulpt every poly
  int a_ell;         // residues per polynomial of `a` (>= ell; DY_COPY with a_ell = ell+1 drops the last residue)
};

EVAB_HD u64x2 ld2(const u64 *p) { return *reinterpret_cast<const u64x2 *>(p); }
EVAB_HD void st2(u64 *p, u64x2 v) { *reinterpret_cast<u64x2 *>(p) = v; }
// 4 contiguous coefficients = one 256-bit global access
EVAB_HD void load4g(u64 (&a)[4], const u64 *p) {
#if defined(__CUDA_ARCH__)
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a[0]), "=l"(a[1]), "=l"(a[2]), "=l"(a[3]) : "l"(p));
#else
  for (int k = 0; k < 4; k++) a[k] = p[k];
#endif
}
EVAB_HD void store4g(u64 *p, const u64 (&a)[4]) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a[0]), "l"(a[1]), "l"(a[2]), "l"(a[3]) : "memory");
#else
  for (int k = 0; k < 4; k++) p[k] = a[k];
#endif
}

// coefficients j .. j+3 of output residue `res` (= s*ell + i): one 256-bit access per operand (a full 32-byte
// sector per lane, twice the bytes in flight of the 128-bit version: multiply_plain went from 54 % to HBM-bound)
// boff: batch instance offset (words) applied to every ciphertext / plaintext pointer
template <int OP> EVAB_HD void dyadic_elem(const DyArgs &A, int res, int j, long long boff = 0) {
  const int s = res / A.ell, i = res % A.ell;
  const PrimeDev P = A.primes[i];
  const u64 p = P.p;
  const size_t off = (size_t)res * A.N;
  const size_t aoff = ((size_t)s * A.a_ell + i) * A.N;
  const bool has_a = s < A.sa;
  const bool has_b = A.b_is_plain ? (OP == DY_MULPT || s == 0) : (s < A.sb);
  const size_t poff = A.b_is_plain ? (size_t)i * A.N : off;
  u64 va[4] = {0, 0, 0, 0}, vb[4] = {0, 0, 0, 0}, r[4];
  if (has_a) load4g(va, A.a + boff + aoff + j);
  if (OP != DY_NEG && OP != DY_COPY && has_b) load4g(vb, A.b + boff + poff + j);
#pragma unroll
  for (int e = 0; e < 4; e++) {
    if (OP == DY_ADD) r[e] = addmod(va[e], vb[e], p);
    else if (OP == DY_SUB) r[e] = submod(va[e], vb[e], p);
    else if (OP == DY_NEG) r[e] = negmod(va[e], p);
    else if (OP == DY_COPY) r[e] = va[e];
    else r[e] = mulmod_p(va[e], vb[e], P);
  }
  store4g(A.out + boff + off + j, r);
}

// Evaluator::multiply_plain: coefficients j .. j+3 of residue i of EVERY polynomial of the ciphertext -- the plaintext
// is loaded once per thread instead of once per polynomial (algorithmic traffic (2s + 1) ell R)
EVAB_HD void mulpt_elem(const DyArgs &A, int i, int j, long long boff = 0) {
  const PrimeDev P = A.primes[i];
  u64 w[4];
  load4g(w, A.b + boff + (size_t)i * A.N + j);
  for (int s = 0; s < A.sa; s++) {
    u64 v[4];
    load4g(v, A.a + boff + ((size_t)s * A.a_ell + i) * A.N + j);
#pragma unroll
    for (int e = 0; e < 4; e++) v[e] = mulmod_p(v[e], w[e], P);
    store4g(A.out + boff + ((size_t)s * A.ell + i) * A.N + j, v);
  }
}

// fused sum of terms; term t is  ct[t]                      (kind 0: Evaluator::add operand),
//                               ct[t] * pt[t]              (kind 1: Evaluator::multiply_plain result), or
//                               ct[t] (x) ct2[t], 2x2 -> 3 (kind 2: Evaluator::multiply / square result)
// i.e. a chain of multiply_plain / multiply / add calls (reference seal_executor.h:124,162-168)
// evaluated in one pass.  Products are accumulated in 128 bits and reduced once; the canonical
// result equals the sequence of canonical products and sums (exact modular arithmetic).
#define SUM_MAX_TERMS 32
struct SumArgs {
  u64 *out;
  const u64 *ct[SUM_MAX_TERMS];
  const u64 *pt[SUM_MAX_TERMS];          // kind 1: plaintext [ell][N]; kind 2: second ciphertext [2][ell][N]
  unsigned char size[SUM_MAX_TERMS];    // polynomials of the term's value (kind 2: 3)
  unsigned char kind[SUM_MAX_TERMS];
  const PrimeDev *primes;
  int n, ell, N, sout;
};
// coefficients j, j+1 of output residue `res` (= s*ell + i); off = batch instance offset (words)
EVAB_HD void sum_terms_elem(const SumArgs &A, int res, int j, long long off) {
  const int s = res / A.ell, i = res % A.ell;
  const PrimeDev P = A.primes[i];
  const size_t poly = (size_t)A.ell * A.N;
  const size_t roff = (size_t)i * A.N + j;          // residue i of polynomial 0
  u64 lx = 0, hx = 0, ly = 0, hy = 0;
  for (int t = 0; t < A.n; t++) {
    if (s >= (int)A.size[t]) continue;
    const u64 *a = A.ct[t] + off + roff;
    if (A.kind[t] == 2) {
      const u64 *b = A.pt[t] + off + roff;
      // tensor product component s: (a0 b0, a0 b1 + a1 b0, a1 b1)
      const u64x2 a_lo = ld2(a + (s == 2 ? poly : 0)), b_hi = ld2(b + (s == 0 ? 0 : poly));
      mac128(lx, hx, a_lo.x, b_hi.x); mac128(ly, hy, a_lo.y, b_hi.y);
      if (s == 1) {
        const u64x2 a1 = ld2(a + poly), b0 = ld2(b);
        mac128(lx, hx, a1.x, b0.x); mac128(ly, hy, a1.y, b0.y);
      }
    } else {
      const u64x2 v = ld2(a + (size_t)s * poly);
      if (A.kind[t] == 1) {
        const u64x2 w = ld2(A.pt[t] + off + roff);
        mac128(lx, hx, v.x, w.x); mac128(ly, hy, v.y, w.y);
      } else {
        lx += v.x; hx += (lx < v.x);
        ly += v.y; hy += (ly < v.y);
      }
    }
  }
  u64x2 r;
  r.x = reduce128(lx, hx, P);
  r.y = reduce128(ly, hy, P);
  st2(A.out + off + (size_t)res * A.N + j, r);
}

struct MulArgs { u64 *out; const u64 *a; const u64 *b; const PrimeDev *primes; int ell, N; };

// 2x2 -> 3 tensor product (Evaluator::multiply) or square, residue i, coeffs j,j+1
template <bool SQUARE> EVAB_HD void mulct_elem(const MulArgs &A, int i, int j, long long boff = 0) {
  const PrimeDev P = A.primes[i];
  const u64 p = P.p;
  const size_t poly = (size_t)A.ell * A.N, off = (size_t)i * A.N + j;
  const u64 *pa = A.a + boff, *pb = SQUARE ? nullptr : A.b + boff;
  u64 *po = A.out + boff;
  const u64x2 a0 = ld2(pa + off), a1 = ld2(pa + poly + off);
  u64x2 d0, d1, d2;
  if (SQUARE) {
    d0.x = mulmod_p(a0.x, a0.x, P); d0.y = mulmod_p(a0.y, a0.y, P);
    u64 x0 = mulmod_p(a0.x, a1.x, P), x1 = mulmod_p(a0.y, a1.y, P);
    d1.x = addmod(x0, x0, p); d1.y = addmod(x1, x1, p);
    d2.x = mulmod_p(a1.x, a1.x, P); d2.y = mulmod_p(a1.y, a1.y, P);
  } else {
    const u64x2 b0 = ld2(pb + off), b1 = ld2(pb + poly + off);
    d0.x = mulmod_p(a0.x, b0.x, P); d0.y = mulmod_p(a0.y, b0.y, P);
    // a0*b1 + a1*b0 accumulated in 128 bits, then one reduction: same canonical value
    u64 lo = 0, hi = 0;
    mac128(lo, hi, a0.x, b1.x); mac128(lo, hi, a1.x, b0.x); d1.x = reduce128(lo, hi, P);
    lo = hi = 0;
    mac128(lo, hi, a0.y, b1.y); mac128(lo, hi, a1.y, b0.y); d1.y = reduce128(lo, hi, P);
    d2.x = mulmod_p(a1.x, b1.x, P); d2.y = mulmod_p(a1.y, b1.y, P);
  }
  st2(po + off, d0); st2(po + poly + off, d1); st2(po + 2 * poly + off, d2);
}

// key-switch inner product over digits (Appendix A.5 step 2):
// acc[c][m][j] = sum_J opnd(m,J)[j] * key[J][c][row(m)][j] mod m, with
// opnd(m,J) = t[J] when row(m)==J (the digit's own modulus) else ext[m][J].
struct IpArgs {
  const u64 *t, *ext, *key; u64 *acc; const PrimeDev *primes;
  const u32 *tperm;   // optional: t is read through this NTT-domain permutation (rotation without a permuted copy)
  const u32 *eperm;   // optional: the extended digits are read through the same permutation (they were computed once, unrotated,
                      // for all rotations of the ciphertext: hoisted_modup), and `cadd` [2][ell+1][N] is added to the result
  const u64 *cadd;
  int ell, k, N;
};
// off: batch instance offset (words) of t / ext / acc (the key is shared by all instances)
EVAB_HD void ks_inner_elem(const IpArgs &A, int mi, int j, long long off = 0) {
  const int row = (mi == A.ell) ? A.k - 1 : mi;
  const PrimeDev P = A.primes[row];
  const size_t N = A.N;
  u64 l0x = 0, h0x = 0, l0y = 0, h0y = 0, l1x = 0, h1x = 0, l1y = 0, h1y = 0;
  for (int J = 0; J < A.ell; J++) {
    u64x2 v;
    if (row == J) {
      const u64 *tp = A.t + off + (size_t)J * N;
      if (A.tperm) { v.x = EVAB_LDG(tp + EVAB_LDG(A.tperm + j)); v.y = EVAB_LDG(tp + EVAB_LDG(A.tperm + j + 1)); }
      else v = ld2(tp + j);
    } else {
      const u64 *ep = A.ext + off + ((size_t)mi * A.ell + J) * N;
      if (A.eperm) { v.x = EVAB_LDG(ep + EVAB_LDG(A.eperm + j)); v.y = EVAB_LDG(ep + EVAB_LDG(A.eperm + j + 1)); }
      else v = ld2(ep + j);
    }
    const u64x2 k0 = ld2(A.key + (((size_t)J * 2 + 0) * A.k + row) * N + j);
    const u64x2 k1 = ld2(A.key + (((size_t)J * 2 + 1) * A.k + row) * N + j);
    mac128(l0x, h0x, v.x, k0.x); mac128(l0y, h0y, v.y, k0.y);
    mac128(l1x, h1x, v.x, k1.x); mac128(l1y, h1y, v.y, k1.y);
  }
  u64x2 r0, r1;
  // ext operands may be lazily reduced (< 16p, see EPI_STORE_LAZY): wide reduction
  r0.x = reduce128(l0x, h0x, P); r0.y = reduce128(l0y, h0y, P);
  r1.x = reduce128(l1x, h1x, P); r1.y = reduce128(l1y, h1y, P);
  if (A.cadd) {
    const u64x2 c0 = ld2(A.cadd + ((size_t)0 * (A.ell + 1) + mi) * N + j), c1 = ld2(A.cadd + ((size_t)1 * (A.ell + 1) + mi) * N + j);
    r0.x = addmod(r0.x, c0.x, P.p); r0.y = addmod(r0.y, c0.y, P.p);
    r1.x = addmod(r1.x, c1.x, P.p); r1.y = addmod(r1.y, c1.y, P.p);
  }
  st2(A.acc + off + ((size_t)0 * (A.ell + 1) + mi) * N + j, r0);
  st2(A.acc + off + ((size_t)1 * (A.ell + 1) + mi) * N + j, r1);
}

// The shared extended digits ext[m][J] have no entry for m == q_J (the digit itself is used): rotate_modup_prepare_impl stores
// P * c0 mod q_m there, for the batched rotation kernels below.
EVAB_HD void scale_c0_elem(const u64 *c0, u64 *ext, const PrimeDev *primes, int ell, int k, int N, int mi, int j, long long off = 0) {
  const PrimeDev P = primes[mi];
  const u64 pm = primes[k - 1].p % P.p;
  ext[off + ((size_t)mi * ell + mi) * N + j] = mulmod_p(c0[off + (size_t)mi * N + j], pm, P);
}
// ---- all rotations of one ciphertext in one launch (exact; ops_impl.hpp rotate_modup_many_impl): rotation i's inner product through
// its permutation + cadd_i, and P * perm_i(c0) added to polynomial 0 -- the division by P that follows returns exactly
// perm_i(c0) + the key-switched part, the same residues as adding the permuted c0 afterwards ((x + P c - corr) P^-1 = (x - corr) P^-1 + c).
#define ROTMANY_MAX 16
struct RotManyArgs {
  const u64 *t, *ext;             // source c1 [ell][N], shared extended digits [ell+1][ell][N] (diagonal: P * c0, see scale_c0_elem)
  const u32 *perm[ROTMANY_MAX];
  const u64 *key[ROTMANY_MAX], *cadd[ROTMANY_MAX];
  u64 *acc;                       // [n][2][ell+1][N]
  const PrimeDev *primes;
  int n, ell, k, N;
};
EVAB_HD void rot_many_elem(const RotManyArgs &A, int i, int mi, int j, long long off = 0) {
  const int row = (mi == A.ell) ? A.k - 1 : mi;
  const PrimeDev P = A.primes[row];
  const size_t N = A.N;
  const u32 pj0 = EVAB_LDG(A.perm[i] + j), pj1 = EVAB_LDG(A.perm[i] + j + 1);
  u64 l0x = 0, h0x = 0, l0y = 0, h0y = 0, l1x = 0, h1x = 0, l1y = 0, h1y = 0;
  for (int J = 0; J < A.ell; J++) {
    const u64 *src = (row == J) ? A.t + off + (size_t)J * N : A.ext + off + ((size_t)mi * A.ell + J) * N;
    const u64 vx = EVAB_LDG(src + pj0), vy = EVAB_LDG(src + pj1);
    const u64x2 k0 = ld2(A.key[i] + (((size_t)J * 2 + 0) * A.k + row) * N + j);
    const u64x2 k1 = ld2(A.key[i] + (((size_t)J * 2 + 1) * A.k + row) * N + j);
    mac128(l0x, h0x, vx, k0.x); mac128(l0y, h0y, vy, k0.y);
    mac128(l1x, h1x, vx, k1.x); mac128(l1y, h1y, vy, k1.y);
  }
  const u64x2 c0 = ld2(A.cadd[i] + ((size_t)0 * (A.ell + 1) + mi) * N + j), c1 = ld2(A.cadd[i] + ((size_t)1 * (A.ell + 1) + mi) * N + j);
  u64x2 r0, r1;
  r0.x = addmod(reduce128(l0x, h0x, P), c0.x, P.p); r0.y = addmod(reduce128(l0y, h0y, P), c0.y, P.p);
  r1.x = addmod(reduce128(l1x, h1x, P), c1.x, P.p); r1.y = addmod(reduce128(l1y, h1y, P), c1.y, P.p);
  if (mi < A.ell) {   // + perm_i(P * c0), which rotate_modup_prepare_impl left on the diagonal of ext
    const u64 *c0p = A.ext + off + ((size_t)mi * A.ell + mi) * N;
    r0.x = addmod(r0.x, EVAB_LDG(c0p + pj0), P.p);
    r0.y = addmod(r0.y, EVAB_LDG(c0p + pj1), P.p);
  }
  st2(A.acc + off + (((size_t)i * 2 + 0) * (A.ell + 1) + mi) * N + j, r0);
  st2(A.acc + off + (((size_t)i * 2 + 1) * (A.ell + 1) + mi) * N + j, r1);
}

// ---- shared mod-up of a rotation group (exact): per Galois key and level, the constant
//   cadd[c][m] = NTT_m(I_g) (.) sum_{J < ell} (q_J mod m) * key[J][c][row(m)]          (mod m)
// where I_g is the indicator polynomial of the coefficients the automorphism negates (ops_impl.hpp: hoisted_modup).
struct HoistConstArgs {
  const u64 *ind;     // [ell+1][N]  NTT_m(I_g), row order q_0 .. q_{ell-1}, P
  const u64 *key;     // [k-1][2][k][N]
  u64 *out;           // [2][ell+1][N]
  const PrimeDev *primes;
  int ell, k, N;
};
EVAB_HD void hoist_const_elem(const HoistConstArgs &A, int mi, int j) {
  const int row = (mi == A.ell) ? A.k - 1 : mi;
  const PrimeDev P = A.primes[row];
  const size_t N = A.N;
  const u64x2 ni = ld2(A.ind + (size_t)mi * N + j);
  for (int c = 0; c < 2; c++) {
    u64 lx = 0, hx = 0, ly = 0, hy = 0;
    for (int J = 0; J < A.ell; J++) {
      const u64 qj = A.primes[J].p % P.p;
      const u64x2 kv = ld2(A.key + (((size_t)J * 2 + c) * A.k + row) * N + j);
      mac128(lx, hx, qj, kv.x); mac128(ly, hy, qj, kv.y);
    }
    u64x2 r;
    r.x = mulmod_p(reduce128(lx, hx, P), ni.x, P);
    r.y = mulmod_p(reduce128(ly, hy, P), ni.y, P);
    st2(A.out + ((size_t)c * (A.ell + 1) + mi) * N + j, r);
  }
}
// I_g as residues: row mi, coefficient j = sign bit of the coefficient-domain gather table
EVAB_HD void hoist_indicator_elem(u64 *out, const u32 *ctab, int N, int mi, int j) { out[(size_t)mi * N + j] = EVAB_LDG(ctab + j) & 1u; }

// ---- lazy_rotsum (OPT-IN, NOT bit-exact: SURVEY 8f-4) ------------------------------------------------------------------
// sum_i w_i (.) rotate(x, g_i) for several rotations of ONE ciphertext with plaintext weights, with a single mod-down for the
// whole sum: T[c][m] = sum_i w_i[m] (.) acc_i[c][m] over the extended basis (q_0 .. q_{ell-1}, P), acc_i the key-switch
// accumulator of rotation i (shared extended digits through the permutation + cadd, as in ks_inner_elem).  The reference
// rounds every rotation's accumulator down by P and multiplies afterwards; rounding the weighted sum once differs from that in
// the last bits (and carries less rounding noise) -- results are graded by the reference's MSE criterion only.
#define LRS_MAX 16
#define LRS_OUT 4
struct LazyRotSumArgs {
  const u64 *t, *ext;             // source c1 [ell][N], shared extended digits [ell+1][ell][N] (diagonal: P * c0, see scale_c0_elem)
  const u32 *perm[LRS_MAX];
  const u64 *key[LRS_MAX], *cadd[LRS_MAX];
  const u64 *wt[LRS_OUT][LRS_MAX]; // plaintext weights [ell+1][N], last row mod P (batch instance offset applies); null: rotation not in that sum
  u64 *acc;                       // [nout][2][ell+1][N]; poly 0 already carries P * sum_i w_i (.) perm_i(c0) (which the division by P gives back)
  const PrimeDev *primes;
  int n, nout, ell, k, N;
};
// One rotation's share at elements (j, j+1) of row mi, for every output sum it takes part in.  The key-switch inner product of
// the rotation is computed once and weighted per output; every value is reduced below p, so that up to LRS_MAX of them add up
// inside 64 bits.  part[o] = { w.(ip0+cadd0) + P.w.perm(c0), w.(ip1+cadd1) } for x and y.
EVAB_HD void lazy_rotsum_part(const LazyRotSumArgs &A, int mi, int j, int i, long long off, u64 part[LRS_OUT][4]) {
  const int row = (mi == A.ell) ? A.k - 1 : mi;
  const PrimeDev P = A.primes[row];
  const size_t N = A.N;
  const u32 pj0 = EVAB_LDG(A.perm[i] + j), pj1 = EVAB_LDG(A.perm[i] + j + 1);
  u64 l0x = 0, h0x = 0, l0y = 0, h0y = 0, l1x = 0, h1x = 0, l1y = 0, h1y = 0;
#pragma unroll 4
  for (int J = 0; J < A.ell; J++) {
    const u64 *src = (row == J) ? A.t + off + (size_t)J * N : A.ext + off + ((size_t)mi * A.ell + J) * N;
    const u64 vx = EVAB_LDG(src + pj0), vy = EVAB_LDG(src + pj1);
    const u64x2 k0 = ld2(A.key[i] + (((size_t)J * 2 + 0) * A.k + row) * N + j);
    const u64x2 k1 = ld2(A.key[i] + (((size_t)J * 2 + 1) * A.k + row) * N + j);
    mac128(l0x, h0x, vx, k0.x); mac128(l0y, h0y, vy, k0.y);
    mac128(l1x, h1x, vx, k1.x); mac128(l1y, h1y, vy, k1.y);
  }
  const u64x2 c0 = ld2(A.cadd[i] + ((size_t)0 * (A.ell + 1) + mi) * N + j), c1 = ld2(A.cadd[i] + ((size_t)1 * (A.ell + 1) + mi) * N + j);
  u64 a0x = addmod(reduce128(l0x, h0x, P), c0.x, P.p), a0y = addmod(reduce128(l0y, h0y, P), c0.y, P.p);
  const u64 a1x = addmod(reduce128(l1x, h1x, P), c1.x, P.p), a1y = addmod(reduce128(l1y, h1y, P), c1.y, P.p);
  if (mi < A.ell) {   // + perm(P * c0) (the diagonal of ext): comes back as perm(c0) from the division by P
    const u64 *c0p = A.ext + off + ((size_t)mi * A.ell + mi) * N;
    a0x = addmod(a0x, EVAB_LDG(c0p + pj0), P.p);
    a0y = addmod(a0y, EVAB_LDG(c0p + pj1), P.p);
  }
#pragma unroll
  for (int o = 0; o < LRS_OUT; o++) {
    if (o >= A.nout || !A.wt[o][i]) { part[o][0] = part[o][1] = part[o][2] = part[o][3] = 0; continue; }
    const u64x2 w = ld2(A.wt[o][i] + off + (size_t)mi * N + j);
    part[o][0] = mulmod_p(a0x, w.x, P); part[o][1] = mulmod_p(a0y, w.y, P);
    part[o][2] = mulmod_p(a1x, w.x, P); part[o][3] = mulmod_p(a1y, w.y, P);
  }
}
// sum[4]: the parts of all rotations of output o added up as plain integers (< LRS_MAX * p < 2^64)
EVAB_HD void lazy_rotsum_store(const LazyRotSumArgs &A, int mi, int j, int o, long long off, const u64 sum[4]) {
  const int row = (mi == A.ell) ? A.k - 1 : mi;
  const PrimeDev P = A.primes[row];
  const size_t N = A.N;
  u64x2 r0, r1;
  r0.x = reduce128(sum[0], 0, P); r0.y = reduce128(sum[1], 0, P);
  r1.x = reduce128(sum[2], 0, P); r1.y = reduce128(sum[3], 0, P);
  st2(A.acc + off + (((size_t)o * 2 + 0) * (A.ell + 1) + mi) * N + j, r0);
  st2(A.acc + off + (((size_t)o * 2 + 1) * (A.ell + 1) + mi) * N + j, r1);
}
