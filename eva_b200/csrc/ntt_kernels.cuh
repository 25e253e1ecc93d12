// ntt_kernels.cuh -- batched NTT/iNTT kernel bodies with fused prologues and
// epilogues for the CKKS evaluator ops (rescale, key-switch mod-up/mod-down).
//
// A launch is described by one NttLaunch passed by value: CTA j works on job
// (q, r) = (j / inner, j % inner); every pointer is base + q*stride_q +
// r*stride_r (element strides), the prime is pmap[q] or pmap[r].
//
// The bodies are split into phases separated by block-wide barriers so that
// the CPU emulator (tests) can replay them thread by thread.
#pragma once
#include "ntt_core.cuh"

enum : int {
  PRO_PLAIN = 0,   // x = src[idx]
  PRO_MODRED = 1,  // x = (src[idx] mod p) - sub_r   (sub_r = subtab[pmap]: e.g. floor(q_last/2) mod p)
  PRO_GATHER = 2,  // x = src[perm[idx]]              (Galois automorphism, NTT domain)
  PRO_MODRED_SG = 3,  // PRO_MODRED on the signed gather +-src[perm[idx] >> 1] (negate when perm[idx] & 1):
                      // Galois automorphism in the coefficient domain, values canonical mod the source prime
};
enum : int {
  EPI_STORE = 0,     // dst[idx] = x
  EPI_ADDHALF = 1,   // dst[idx] = (x + half) mod p        (half = floor(p/2))
  EPI_DIVROUND = 2,  // dst[idx] = (aux0[idx] - x) * c mod p  [+ aux1[idx] mod p]
  EPI_STORE_LAZY = 3,  // dst[idx] = x without canonicalisation (x < 16p; consumer reduces)
  EPI_STORE_ZFLAG = 4, // inverse: EPI_STORE, and *zflag is set when any stored coefficient is zero (shared mod-up of
                       // rotation groups: see hoisted_modup in ops_impl.hpp)
};

struct NttLaunch {
  const u64 *src; u64 *dst; const u64 *aux0; const u64 *aux1;
  const u32 *perm;             // PRO_GATHER / PRO_MODRED_SG
  const u32 *aux1_perm;        // EPI_DIVROUND: aux1 is read through this permutation (NTT-domain automorphism)
  u64 *zflag;                  // EPI_STORE_ZFLAG: one word (per batch instance), OR-ed with 1 when a zero coefficient is stored
  const u64 *cflags;           // optional [q]: 0 = polynomial q is constant (only coefficient 0 set): its
                               // transform is that value everywhere, written without running the NTT
  const PrimeDev *primes;
  const u64x2 *consts;         // EPI_DIVROUND: {c, shoup(c)} per prime index; PRO_MODRED: .x of subtab
  const u64x2 *consts_f;       // EPI_DIVROUND, fold arithmetic: {c, c * 2^32 mod p}
  const u64 *subtab;           // PRO_MODRED: value to subtract per prime index (canonical mod that prime)
  long long src_sq, src_sr, dst_sq, dst_sr, aux0_sq, aux0_sr, aux1_sq, aux1_sr;
  int inner;                   // jobs per q
  int prime_on_q;              // 1: prime = pmap[q], 0: prime = pmap[r]
  int pro, epi;
  int skip_diag;               // key-switch mod-up: CTA exits when pmap[q] == pmap2[r]
  int aux1_polys;              // aux1 applies to q < aux1_polys only (rotate: c0 has a base, c1 none)
  unsigned char pmap[32];
  unsigned char pmap2[32];
  FoldPrime fp[NTT_MAX_PRIMES];   // indexed by prime index (filled by base_launch from CtxView::fold_host)
};

// fold arithmetic applies when every prime the launch computes in is fold-friendly (bit i of foldmask: prime i)
inline bool ntt_launch_folds(const NttLaunch &L, size_t jobs, unsigned foldmask) {
  const size_t np = L.prime_on_q ? (jobs + (size_t)L.inner - 1) / (size_t)L.inner : (size_t)L.inner;
  if (np > 32) return false;
  for (size_t i = 0; i < np; i++)
    if (!((foldmask >> L.pmap[i]) & 1u)) return false;
  return true;
}

struct NttState { u64 x[NTT_E]; int b; int bb[NTT_E]; };   // b: Shoup path bound (units of p); bb: fold path, per register, units of p/16

EVAB_HD void flag_or(u64 *f) {
#if defined(__CUDA_ARCH__)
  atomicOr(reinterpret_cast<unsigned long long *>(f), 1ull);
#else
  *f |= 1ull;
#endif
}
struct NttJob {
  const u64 *src; u64 *dst; const u64 *aux0; const u64 *aux1; u64 *zflag;
  u32 pi; u32 h; bool skip;
  bool bcast;  // constant polynomial (cflags)
  u32 spi;   // prime index of the values stored in src (PRO_MODRED)
};

// boff: element offset of this batch instance (evab_set_batch), applied to every data pointer
// job (q, r), CTA h of its cluster.  The CUDA grid is (inner * CL, q, batch), so that the prime index is a plain
// table lookup on block indices (warp-uniform: the per-prime constants land in uniform registers)
EVAB_HD NttJob ntt_job_qr(const NttLaunch &L, u32 q, u32 r, u32 h, long long boff = 0) {
  NttJob J;
  J.h = h;
  J.pi = L.prime_on_q ? L.pmap[q] : L.pmap[r];
  J.skip = L.skip_diag && (L.pmap[q] == L.pmap2[r]);
  J.spi = L.pmap2[r];
  J.bcast = L.cflags && (L.cflags + boff)[q] == 0;
  J.src = L.src + q * L.src_sq + r * L.src_sr + boff;
  J.dst = L.dst + q * L.dst_sq + r * L.dst_sr + boff;
  J.aux0 = L.aux0 ? L.aux0 + q * L.aux0_sq + r * L.aux0_sr + boff : nullptr;
  J.aux1 = (L.aux1 && (int)q < L.aux1_polys) ? L.aux1 + q * L.aux1_sq + r * L.aux1_sr + boff : nullptr;
  J.zflag = L.zflag ? L.zflag + boff : nullptr;
  return J;
}
EVAB_HD NttJob ntt_job(const NttLaunch &L, u32 cta, int ctas_per_job, long long boff = 0) {
  const u32 job = cta / ctas_per_job;
  return ntt_job_qr(L, job / L.inner, job % L.inner, cta % ctas_per_job, boff);
}

// PRO_MODRED: the source holds canonical residues of prime `srcp`; when srcp <= 2p one
// conditional subtraction reduces them mod p, otherwise a 64-bit Barrett reduction
template <int PRO, bool CHEAP> EVAB_HD u64 pro_load(const NttLaunch &L, const NttJob &J, const PrimeDev &P, u32 idx, u64 sub) {
  if (PRO == PRO_GATHER) return EVAB_LDG(J.src + EVAB_LDG(L.perm + idx));
  u64 v;
  if (PRO == PRO_MODRED_SG) {
    const u32 e = EVAB_LDG(L.perm + idx);
    v = EVAB_LDG(J.src + (e >> 1));
    if ((e & 1u) && v) v = L.primes[J.spi].p - v;   // canonical negation mod the source prime
  } else {
    v = EVAB_LDG(J.src + idx);
  }
  if (PRO == PRO_MODRED || PRO == PRO_MODRED_SG) {
    v = CHEAP ? csub(v, P.p) : barrett64(v, P.p, P.ratio64);
    v = submod(v, sub, P.p);
  }
  return v;
}

// 4 contiguous coefficients = one 256-bit access
EVAB_HD void load4(u64 (&a)[4], const u64 *p) {
#if defined(__CUDA_ARCH__)
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a[0]), "=l"(a[1]), "=l"(a[2]), "=l"(a[3]) : "l"(p));
#else
  for (int k = 0; k < 4; k++) a[k] = p[k];
#endif
}
EVAB_HD void store4(u64 *p, const u64 (&a)[4]) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a[0]), "l"(a[1]), "l"(a[2]), "l"(a[3]) : "memory");
#else
  for (int k = 0; k < 4; k++) p[k] = a[k];
#endif
}
// 16 contiguous coefficients with 256-bit accesses (full 32-byte sectors per lane)
EVAB_HD void load16(u64 (&a)[NTT_E], const u64 *p) {
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int c = 0; c < NTT_E / 4; c++) {
    u64 v0, v1, v2, v3;
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v0), "=l"(v1), "=l"(v2), "=l"(v3) : "l"(p + 4 * c));
    a[4 * c] = v0; a[4 * c + 1] = v1; a[4 * c + 2] = v2; a[4 * c + 3] = v3;
  }
#else
  for (int k = 0; k < NTT_E; k++) a[k] = p[k];
#endif
}
EVAB_HD void store16(u64 *p, const u64 (&a)[NTT_E]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int c = 0; c < NTT_E / 4; c++)
    asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p + 4 * c), "l"(a[4 * c]), "l"(a[4 * c + 1]), "l"(a[4 * c + 2]), "l"(a[4 * c + 3]) : "memory");
#else
  for (int k = 0; k < NTT_E; k++) p[k] = a[k];
#endif
}
// forward transform of a constant polynomial c (canonical): c at every evaluation point.
// Virtual thread tid writes its 16 contiguous outputs (in place: dst[0] keeps its value).
template <int LOGN> EVAB_HD void fwd_const_poly(const NttJob &J, u32 tid) {
  const u64 v = EVAB_LDG(J.src);
  u64 x[NTT_E];
#pragma unroll
  for (int k = 0; k < NTT_E; k++) x[k] = v;
  store16(J.dst + (size_t)tid * NTT_E, x);
}


// ------------------------------ forward ------------------------------------
// Phases (separated by block barriers), P = NttGeom::P register passes:
//   0          : load (layout of pass 0) + pass 0 + exchange write
//   2j-1, 2j   : exchange read + pass j   |   exchange write        (1 <= j <= P-2)
//   2(P-1)-1   : exchange read + last (contiguous) pass
//   phE        : fused epilogue + store
// AR = 0: lazy Shoup butterflies (any prime < 2^60); AR = 1: two-row fold arithmetic (every prime of the launch fold-friendly)
template <int LOGN, int PRO = PRO_PLAIN, int EPI = EPI_STORE, int CL = 1, int AR = 0> struct FwdBody {
  typedef NttGeom<LOGN> G;
  typedef ClGeom<LOGN, CL> C;
  static constexpr int NPH = G::NPH;
  // barrier after phase PH: 0 = block, 1 = cluster (the exchange written in phase 0 crosses CTAs)
  static EVAB_HD constexpr int sync_kind(int ph) { return (CL > 1 && ph == 0) ? 1 : 0; }
  // virtual thread id: rank * Tc + tid
  static EVAB_HD u32 vtid(const NttJob &J, u32 tid) { return CL > 1 ? J.h * (u32)C::Tc + tid : tid; }

  // load of pass 0 (strided, coalesced) with the fused prologue
  template <bool CHEAP> static EVAB_HD void load0(NttState &S, const NttLaunch &L, const NttJob &J, const PrimeDev &P, u32 tid, u64 sub) {
#pragma unroll
    for (int k = 0; k < NTT_E; k++) S.x[k] = pro_load<PRO, CHEAP>(L, J, P, idx_s<LOGN, 0>(tid, k), sub);
    S.b = 1;
  }
  // fold arithmetic: the multiplicand of a product may be any u64, so the reduction of the source residues to
  // this prime disappears: x = v + (p - sub) < 2^60 + p
  static EVAB_HD void load0_fold(NttState &S, const NttLaunch &L, const NttJob &J, const PrimeDev &P, u32 tid, u64 sub) {
    constexpr bool MR = (PRO == PRO_MODRED || PRO == PRO_MODRED_SG);
    const u64 add = L.fp[J.pi].p - sub;
#pragma unroll
    for (int k = 0; k < NTT_E; k++) {
      const u32 idx = idx_s<LOGN, 0>(tid, k);
      u64 v;
      if (PRO == PRO_MODRED_SG) {
        const u32 e = EVAB_LDG(L.perm + idx);
        v = EVAB_LDG(J.src + (e >> 1));
        if ((e & 1u) && v) v = L.fp[J.spi].p - v;   // canonical negation mod the source prime
      } else if (PRO == PRO_GATHER) {
        v = EVAB_LDG(J.src + EVAB_LDG(L.perm + idx));
      } else {
        v = EVAB_LDG(J.src + idx);
      }
      S.x[k] = MR ? v + add : v;
      S.bb[k] = MR ? 2 * FB_CANON + 1 : FB_CANON;
    }
  }
  // ltid: thread index inside the CTA; the passes run on the virtual thread id
  template <int PH, class SMV, class Hooks> static EVAB_HD void phase(NttState &S, const NttLaunch &L, const NttJob &J, u32 ltid, const SMV &smv, Hooks hk) {
    const PrimeDev P = L.primes[J.pi];
    const u32 root = 1;   // twiddle root prefix of a full transform
    const u32 tid = vtid(J, ltid);
    u64 *sm = smv.local;
    if constexpr (PH == 0) {
      constexpr bool MR = (PRO == PRO_MODRED || PRO == PRO_MODRED_SG);
      const u64 sub = MR ? EVAB_LDG(L.subtab + J.pi) : 0;
      const bool cheap = MR && (L.primes[J.spi].p <= 2 * P.p);   // CTA-uniform
      if constexpr (AR == 1) {
        load0_fold(S, L, J, P, tid, sub);
        ffwd_pass_s<LOGN, 0>(S.x, S.bb, L.fp[J.pi].ftw, fold_params(L.fp[J.pi]), tid);
      } else {
        if (cheap) load0<true>(S, L, J, P, tid, sub); else load0<false>(S, L, J, P, tid, sub);
        fwd_pass_s<LOGN, 0>(S.x, P.tw, root, P.p, tid, S.b);
      }
      if constexpr (CL > 1) { hk.ready(); xchg_write_dist_fwd<LOGN, CL>(S.x, smv, tid); }   // peers resident, barriers initialised: waited for here, behind pass 0
      else xchg_write_s<LOGN, 0, 1>(S.x, sm, tid);
    } else if constexpr (PH == NPH - 1) {
      if constexpr (CL > 1) xchg_read_cl<LOGN, CL>(S.x, sm, tid); else xchg_read_c<LOGN>(S.x, sm, tid);
      if constexpr (AR == 1) ffwd_pass_c<LOGN>(S.x, S.bb, L.fp[J.pi].ftw, fold_params(L.fp[J.pi]), tid);
      else fwd_pass_c<LOGN>(S.x, P.tw, root, P.p, tid, S.b);
    } else if constexpr (PH % 2 == 1) {
      constexpr int j = (PH + 1) / 2;
      if constexpr (CL > 1) xchg_read_sl<LOGN, j, j, CL>(S.x, sm, tid); else xchg_read_s<LOGN, j, j>(S.x, sm, tid);
      if constexpr (AR == 1) ffwd_pass_s<LOGN, j>(S.x, S.bb, L.fp[J.pi].ftw, fold_params(L.fp[J.pi]), tid);
      else fwd_pass_s<LOGN, j>(S.x, P.tw, root, P.p, tid, S.b);
    } else {
      constexpr int j = PH / 2;
      if constexpr (CL > 1) {
        if constexpr (j + 1 == G::P - 1) xchg_write_scl<LOGN, j, CL>(S.x, sm, tid);
        else xchg_write_sl<LOGN, j, j + 1, CL>(S.x, sm, tid);
      } else {
        if constexpr (j + 1 == G::P - 1) xchg_write_sc<LOGN, j>(S.x, sm, tid);
        else xchg_write_s<LOGN, j, j + 1>(S.x, sm, tid);
      }
    }
  }
  // final phase: fused epilogue + store of 16 contiguous coefficients.
  static EVAB_HD void phE(NttState &S, const NttLaunch &L, const NttJob &J, u32 ltid) {
    const PrimeDev P = L.primes[J.pi];
    const u32 tid = vtid(J, ltid);
    const size_t base = (size_t)tid << NTT_EL;
    if constexpr (AR == 1) { phE_fold(S, L, J, P, base); return; }
    if (EPI == EPI_DIVROUND) {
      // (aux0 - x) * c [+ aux1] without canonicalising x first: x < b*p, so
      // aux0 + b*p - x is positive and < 16p; the Shoup product lands in [0,2p).
      // Processed four coefficients at a time (one 256-bit access per operand) to
      // keep the register footprint at 64.
      const u64x2 c = ldg_tw(L.consts + J.pi);
      const bool fix = S.b > 14;
      const u64 bias = (u64)(fix ? 8 : S.b) * P.p, eight_p = 8 * P.p, two_p = 2 * P.p;
#pragma unroll
      for (int q = 0; q < NTT_E / 4; q++) {
        u64 a[4];
        load4(a, J.aux0 + base + 4 * q);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          u64 x = S.x[4 * q + k];
          if (fix) x = csub(x, eight_p);
          a[k] = shoup_lazy(a[k] + bias - x, c.x, c.y, P.p);
        }
        if (J.aux1) {
          u64 d[4];
          if (L.aux1_perm) {
#pragma unroll
            for (int k = 0; k < 4; k++) d[k] = EVAB_LDG(J.aux1 + EVAB_LDG(L.aux1_perm + base + 4 * q + k));
          } else {
            load4(d, J.aux1 + base + 4 * q);
          }
#pragma unroll
          for (int k = 0; k < 4; k++) a[k] = csub(csub(a[k] + d[k], two_p), P.p);
        } else {
#pragma unroll
          for (int k = 0; k < 4; k++) a[k] = csub(a[k], P.p);
        }
        store4(J.dst + base + 4 * q, a);
      }
      return;
    }
    if (EPI != EPI_STORE_LAZY) canon(S.x, P.p, S.b);
    store16(J.dst + base, S.x);
  }
  // the same epilogues in fold arithmetic
  static EVAB_HD void phE_fold(NttState &S, const NttLaunch &L, const NttJob &J, const PrimeDev &P, size_t base) {
    const FoldP F = fold_params(L.fp[J.pi]);
    if (EPI == EPI_DIVROUND) {
      // (aux0 - x) * c [+ aux1]: aux0 + C p - x with C p >= x is a valid multiplicand as it stands
      const u64x2 c = ldg_tw(L.consts_f + J.pi);
#pragma unroll
      for (int q = 0; q < NTT_E / 4; q++) {
        u64 a[4];
        load4(a, J.aux0 + base + 4 * q);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          u64 x = S.x[4 * q + k];
          int bx = S.bb[4 * q + k];
          if (bx > 128) { x = fold61(x, F.eps); bx = FB_FOLD; }
          a[k] = fold_mul(a[k] + (bx > 48 ? F.p8 : F.p3) - x, c.x, c.y, F.eps);   // < 2.2501 p
        }
        if (J.aux1) {
          u64 d[4];
          if (L.aux1_perm) {
#pragma unroll
            for (int k = 0; k < 4; k++) d[k] = EVAB_LDG(J.aux1 + EVAB_LDG(L.aux1_perm + base + 4 * q + k));
          } else {
            load4(d, J.aux1 + base + 4 * q);
          }
#pragma unroll
          for (int k = 0; k < 4; k++) a[k] += d[k];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) a[k] = fold_canon(a[k], F.eps, F.p);
        store4(J.dst + base + 4 * q, a);
      }
      return;
    }
    if (EPI != EPI_STORE_LAZY) fcanon(S.x, F);
    store16(J.dst + base, S.x);
  }
};

// ------------------------------ inverse ------------------------------------
// Phases: 0: contiguous load + last-pass stages + exchange write;
//         then for j = P-2 .. 1: exchange read + pass j | exchange write;
//         last: exchange read + pass 0 + scale by N^-1 + epilogue store (strided, coalesced).
template <int LOGN, int PRO = PRO_PLAIN, int EPI = EPI_STORE, int CL = 1, int AR = 0> struct InvBody {
  typedef NttGeom<LOGN> G;
  typedef ClGeom<LOGN, CL> C;
  static constexpr int NPH = G::NPH;
  // barrier after phase PH: 0 = block, 1 = cluster, 2 = none.  CL > 1: the exchange into pass 0
  // (written in phase NPH-2) crosses CTAs; every CTA signals "done reading my slice" right after the
  // exchange read of phase NPH-3 (hk.arrive) and waits for all peers before writing (hk.wait), so the
  // barrier latency hides behind the butterflies of that pass.
  static EVAB_HD constexpr int sync_kind(int ph) { return CL == 1 ? 0 : (ph == NPH - 2 ? 1 : (ph == NPH - 3 ? 2 : 0)); }
  static EVAB_HD u32 vtid(const NttJob &J, u32 tid) { return CL > 1 ? J.h * (u32)C::Tc + tid : tid; }

  template <int PH, class SMV, class Hooks> static EVAB_HD void phase(NttState &S, const NttLaunch &L, const NttJob &J, u32 ltid, const SMV &smv, Hooks hk) {
    const PrimeDev P = L.primes[J.pi];
    const u32 root = 1;   // twiddle root prefix of a full transform
    const u32 tid = vtid(J, ltid);
    u64 *sm = smv.local;
    if constexpr (PH == 0) {
      const size_t base = (size_t)tid << NTT_EL;
      if (PRO == PRO_GATHER) {
#pragma unroll
        for (int k = 0; k < NTT_E; k++) S.x[k] = EVAB_LDG(J.src + EVAB_LDG(L.perm + base + k));
      } else {
        load16(S.x, J.src + base);
      }
      S.b = 1;
      if constexpr (AR == 1) {
#pragma unroll
        for (int k = 0; k < NTT_E; k++) S.bb[k] = FB_CANON;
        finv_pass_c<LOGN>(S.x, S.bb, L.fp[J.pi].fitw, fold_params(L.fp[J.pi]), tid);
      } else {
        inv_pass_c<LOGN>(S.x, P.itw, root, P.p, tid, S.b);
      }
      if constexpr (CL > 1) xchg_write_cl<LOGN, CL>(S.x, sm, tid); else xchg_write_c<LOGN>(S.x, sm, tid);
    } else if constexpr (PH == NPH - 1) {
      if constexpr (CL > 1) xchg_read_dist_inv<LOGN, CL>(S.x, sm, ltid); else xchg_read_s<LOGN, 0, 1>(S.x, sm, tid);
      const u64 half = (AR == 1 ? L.fp[J.pi].p : P.p) >> 1;
      if constexpr (AR == 1) {
        finv_pass0_scaled<LOGN>(S.x, S.bb, L.fp[J.pi].fitw, fold_params(L.fp[J.pi]), tid, L.fp[J.pi], EPI == EPI_ADDHALF ? half : 0);   // canonical, half added
        bool zero = false;
#pragma unroll
        for (int k = 0; k < NTT_E; k++) { J.dst[idx_s<LOGN, 0>(tid, k)] = S.x[k]; zero = zero || S.x[k] == 0; }
        if (EPI == EPI_STORE_ZFLAG && zero) flag_or(J.zflag);
      } else {
        inv_pass0_scaled<LOGN>(S.x, P.itw, P.p, tid, S.b, P.ninv, P.ninv_s, P.itw1n, P.itw1n_s);   // * N^-1 folded in, canonical
#pragma unroll
        bool zero = false;
        for (int k = 0; k < NTT_E; k++) {
          u64 v = S.x[k];
          if (EPI == EPI_ADDHALF) v = addmod(v, half, P.p);
          J.dst[idx_s<LOGN, 0>(tid, k)] = v;
          zero = zero || v == 0;
        }
        if (EPI == EPI_STORE_ZFLAG && zero) flag_or(J.zflag);
      }
    } else if constexpr (PH % 2 == 1) {
      constexpr int j = G::P - 2 - (PH - 1) / 2;
      if constexpr (CL > 1) {
        if constexpr (j == G::P - 2) xchg_read_scl<LOGN, j, CL>(S.x, sm, tid);
        else xchg_read_sl<LOGN, j, j + 1, CL>(S.x, sm, tid);
      } else {
        if constexpr (j == G::P - 2) xchg_read_sc<LOGN, j>(S.x, sm, tid);
        else xchg_read_s<LOGN, j, j + 1>(S.x, sm, tid);
      }
      if constexpr (AR == 1) finv_pass_s<LOGN, j>(S.x, S.bb, L.fp[J.pi].fitw, fold_params(L.fp[J.pi]), tid);
      else inv_pass_s<LOGN, j>(S.x, P.itw, root, P.p, tid, S.b);
      // the butterflies above consumed every value read from this CTA's slice: from here on the peers may overwrite it
      if constexpr (CL > 1 && j == 1) { hk.ready(); hk.released(); }   // (the start handshake is waited for here, two passes into the kernel)
    } else {
      constexpr int j = G::P - 2 - (PH - 2) / 2;   // pass that just ran
      if constexpr (CL > 1) {
        if constexpr (j == 1) { hk.acquire_free(); xchg_write_dist_inv<LOGN, CL>(S.x, smv, tid); }
        else xchg_write_sl<LOGN, j, j, CL>(S.x, sm, tid);
      } else {
        xchg_write_s<LOGN, j, j>(S.x, sm, tid);
      }
    }
  }
};

// sync(kind): barrier after a phase (B::sync_kind: 0 block, 1 "the distributed exchange has landed", 2 none).
// hk: cluster protocol around the one exchange that crosses CTAs --
//   start()         kernel start: barriers initialised, CTA announced to its cluster
//   ready()         every peer is resident and initialised (before the first store into a peer)
//   released()      (inverse) this thread has read its slice for the last time before the peers overwrite it
//   acquire_free()  (inverse) every CTA of the cluster has released its slice
struct NoHooks { EVAB_HD void start() const {} EVAB_HD void ready() const {} EVAB_HD void released() const {} EVAB_HD void acquire_free() const {} };
template <class B, int PH, int NPH> struct PhaseLoop {
  template <class SM, class Sync, class Hooks> static EVAB_HD void run(NttState &S, const NttLaunch &L, const NttJob &J, u32 tid, const SM &sm, Sync sync, Hooks hk) {
    B::template phase<PH>(S, L, J, tid, sm, hk);
    if (PH + 1 < NPH) sync(B::sync_kind(PH));
    PhaseLoop<B, PH + 1, NPH>::run(S, L, J, tid, sm, sync, hk);
  }
};
template <class B, int NPH> struct PhaseLoop<B, NPH, NPH> {
  template <class SM, class Sync, class Hooks> static EVAB_HD void run(NttState &, const NttLaunch &, const NttJob &, u32, const SM &, Sync, Hooks) {}
};
