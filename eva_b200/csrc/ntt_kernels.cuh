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
};
enum : int {
  EPI_STORE = 0,     // dst[idx] = x
  EPI_ADDHALF = 1,   // dst[idx] = (x + half) mod p        (half = floor(p/2))
  EPI_DIVROUND = 2,  // dst[idx] = (aux0[idx] - x) * c mod p  [+ aux1[idx] mod p]
};

struct NttLaunch {
  const u64 *src; u64 *dst; const u64 *aux0; const u64 *aux1;
  const u32 *perm;             // PRO_GATHER
  const PrimeDev *primes;
  const u64x2 *consts;         // EPI_DIVROUND: {c, shoup(c)} per prime index; PRO_MODRED: .x of subtab
  const u64 *subtab;           // PRO_MODRED: value to subtract per prime index (canonical mod that prime)
  long long src_sq, src_sr, dst_sq, dst_sr, aux0_sq, aux0_sr, aux1_sq, aux1_sr;
  int inner;                   // jobs per q
  int prime_on_q;              // 1: prime = pmap[q], 0: prime = pmap[r]
  int pro, epi;
  int skip_diag;               // key-switch mod-up: CTA exits when pmap[q] == pmap2[r]
  unsigned char pmap[32];
  unsigned char pmap2[32];
};

struct NttState { u64 x[32]; int b; u32 pi; u64 p; };

template <int LOGN, bool SPLIT> struct NttJobGeom {
  typedef NttGeom<LOGN> G;
  static constexpr int CTAS_PER_JOB = SPLIT ? 2 : 1;
  static constexpr int NFULL = SPLIT ? 2 * G::N : G::N;
};

struct NttJob {
  const u64 *src; u64 *dst; const u64 *aux0; const u64 *aux1;
  u32 pi; u32 h; bool skip;
};

EVAB_HD NttJob ntt_job(const NttLaunch &L, u32 cta, int ctas_per_job) {
  NttJob J;
  u32 job = cta / ctas_per_job;
  J.h = cta % ctas_per_job;
  u32 q = job / L.inner, r = job % L.inner;
  J.pi = L.prime_on_q ? L.pmap[q] : L.pmap[r];
  J.skip = L.skip_diag && (L.pmap[q] == L.pmap2[r]);
  J.src = L.src + q * L.src_sq + r * L.src_sr;
  J.dst = L.dst + q * L.dst_sq + r * L.dst_sr;
  J.aux0 = L.aux0 ? L.aux0 + q * L.aux0_sq + r * L.aux0_sr : nullptr;
  J.aux1 = L.aux1 ? L.aux1 + q * L.aux1_sq + r * L.aux1_sr : nullptr;
  return J;
}

template <int PRO> EVAB_HD u64 pro_load(const NttLaunch &L, const NttJob &J, const PrimeDev &P, u32 idx, u64 sub) {
  if (PRO == PRO_GATHER) return EVAB_LDG(J.src + EVAB_LDG(L.perm + idx));
  u64 v = EVAB_LDG(J.src + idx);
  if (PRO == PRO_MODRED) v = submod(barrett64(v, P.p, P.ratio64), sub, P.p);
  return v;
}

// ------------------------------ forward ------------------------------------
template <int LOGN, bool SPLIT, int PRO = PRO_PLAIN, int EPI = EPI_STORE> struct FwdBody {
  typedef NttGeom<LOGN> G;
  static constexpr int NPH = (G::NC > 0) ? 4 : 2;

  // phase 0: load (layout A) + optional split stage + pass A + exchange write
  static EVAB_HD void ph0(NttState &S, const NttLaunch &L, const NttJob &J, u32 tid, u64 *sm) {
    const PrimeDev P = L.primes[J.pi];
    S.p = P.p; S.pi = J.pi;
    const u64 sub = (PRO == PRO_MODRED) ? EVAB_LDG(L.subtab + J.pi) : 0;
    u32 root = 1;
    if (!SPLIT) {
#pragma unroll
      for (int k = 0; k < 32; k++) S.x[k] = pro_load<PRO>(L, J, P, idx_a<LOGN>(tid, k), sub);
      S.b = 1;
    } else {
      // first stage of the 2N transform: pairs (i, i + N) with twiddle tw[1];
      // this CTA keeps the h-th output half and continues with root prefix 2+h
      const u64x2 w = ldg_tw(P.tw + 1);
      const u64 two_p = 2 * P.p;
#pragma unroll
      for (int k = 0; k < 32; k++) {
        u32 i = idx_a<LOGN>(tid, k);
        u64 X = pro_load<PRO>(L, J, P, i, sub), Y = pro_load<PRO>(L, J, P, i + G::N, sub);
        u64 t = shoup_lazy(Y, w.x, w.y, P.p);
        S.x[k] = J.h ? X - t + two_p : X + t;
      }
      S.b = 3;
      root = 2 + J.h;
    }
    fwd_pass_a<LOGN>(S.x, P.tw, root, P.p, S.b);
    xchg_write_a<LOGN>(S.x, sm, tid);
  }
  // phase 1: exchange read + pass B (+ final epilogue when there is no pass C)
  static EVAB_HD void ph1(NttState &S, const NttLaunch &L, const NttJob &J, u32 tid, u64 *sm) {
    const PrimeDev P = L.primes[J.pi];
    xchg_read_b_ab<LOGN>(S.x, sm, tid);
    fwd_pass_b<LOGN>(S.x, P.tw, SPLIT ? 2 + J.h : 1, P.p, tid, S.b);
  }
  static EVAB_HD void ph2(NttState &S, const NttLaunch &, const NttJob &, u32 tid, u64 *sm) {
    xchg_write_b_bc<LOGN>(S.x, sm, tid);
  }
  static EVAB_HD void ph3(NttState &S, const NttLaunch &L, const NttJob &J, u32 tid, u64 *sm) {
    const PrimeDev P = L.primes[J.pi];
    xchg_read_c(S.x, sm, tid);
    fwd_pass_c<LOGN>(S.x, P.tw, SPLIT ? 2 + J.h : 1, P.p, tid, S.b);
  }
  // final phase: fused epilogue + store of 32 contiguous coefficients (layout C).
  // SPLIT: the two CTAs of a job both read the whole input, so for in-place
  // transforms the caller must barrier the CTA pair (cluster) before this phase.
  static EVAB_HD void phE(NttState &S, const NttLaunch &L, const NttJob &J, u32 tid) {
    const PrimeDev P = L.primes[J.pi];
    canon(S.x, P.p, S.b);
    const size_t base = (size_t)(SPLIT ? J.h * G::N : 0) + ((size_t)tid << 5);
    if (EPI == EPI_DIVROUND) {
      const u64x2 c = ldg_tw(L.consts + J.pi);
      u64 a[32];
      load32(a, J.aux0 + base);
#pragma unroll
      for (int k = 0; k < 32; k++) S.x[k] = shoup_mul(submod(a[k], S.x[k], P.p), c.x, c.y, P.p);
      if (J.aux1) {
        load32(a, J.aux1 + base);
#pragma unroll
        for (int k = 0; k < 32; k++) S.x[k] = addmod(S.x[k], a[k], P.p);
      }
    }
    store32(J.dst + base, S.x);
  }
  static EVAB_HD void load32(u64 (&a)[32], const u64 *p) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int c = 0; c < 8; c++) {
      u64 v0, v1, v2, v3;
      asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v0), "=l"(v1), "=l"(v2), "=l"(v3) : "l"(p + 4 * c));
      a[4 * c] = v0; a[4 * c + 1] = v1; a[4 * c + 2] = v2; a[4 * c + 3] = v3;
    }
#else
    for (int k = 0; k < 32; k++) a[k] = p[k];
#endif
  }
  static EVAB_HD void store32(u64 *p, const u64 (&a)[32]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int c = 0; c < 8; c++)
      asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p + 4 * c), "l"(a[4 * c]), "l"(a[4 * c + 1]), "l"(a[4 * c + 2]), "l"(a[4 * c + 3]) : "memory");
#else
    for (int k = 0; k < 32; k++) p[k] = a[k];
#endif
  }
};

// ------------------------------ inverse ------------------------------------
// SPLIT: each CTA runs the LOGN-stage inverse on one half (root prefix 2+h) and
// stores lazily-reduced values; k_inv_last_stage then finishes the transform.
template <int LOGN, bool SPLIT, int PRO = PRO_PLAIN, int EPI = EPI_STORE> struct InvBody {
  typedef NttGeom<LOGN> G;
  typedef FwdBody<LOGN, SPLIT> F;
  static constexpr int NPH = (G::NC > 0) ? 4 : 2;

  static EVAB_HD void load_c(NttState &S, const NttLaunch &L, const NttJob &J, u32 tid) {
    const size_t base = (size_t)(SPLIT ? J.h * G::N : 0) + ((size_t)tid << 5);
    if (PRO == PRO_GATHER) {
#pragma unroll
      for (int k = 0; k < 32; k++) S.x[k] = EVAB_LDG(J.src + EVAB_LDG(L.perm + base + k));
    } else {
      F::load32(S.x, J.src + base);
    }
    S.b = 1;
  }
  static EVAB_HD void ph0(NttState &S, const NttLaunch &L, const NttJob &J, u32 tid, u64 *sm) {
    const PrimeDev P = L.primes[J.pi];
    S.p = P.p; S.pi = J.pi;
    load_c(S, L, J, tid);
    if (G::NC > 0) {
      inv_pass_c<LOGN>(S.x, P.itw, SPLIT ? 2 + J.h : 1, P.p, tid, S.b);
      xchg_write_c(S.x, sm, tid);
    } else {
      inv_pass_b<LOGN>(S.x, P.itw, SPLIT ? 2 + J.h : 1, P.p, tid, S.b);
      xchg_write_b_ab<LOGN>(S.x, sm, tid);
    }
  }
  static EVAB_HD void ph1(NttState &S, const NttLaunch &L, const NttJob &J, u32 tid, u64 *sm) {
    const PrimeDev P = L.primes[J.pi];
    if (G::NC > 0) {
      xchg_read_b_bc<LOGN>(S.x, sm, tid);
      inv_pass_b<LOGN>(S.x, P.itw, SPLIT ? 2 + J.h : 1, P.p, tid, S.b);
    } else {
      finish(S, L, J, P, tid, sm);
    }
  }
  static EVAB_HD void ph2(NttState &S, const NttLaunch &, const NttJob &, u32 tid, u64 *sm) {
    xchg_write_b_ab<LOGN>(S.x, sm, tid);
  }
  static EVAB_HD void ph3(NttState &S, const NttLaunch &L, const NttJob &J, u32 tid, u64 *sm) {
    const PrimeDev P = L.primes[J.pi];
    finish(S, L, J, P, tid, sm);
  }
  static EVAB_HD void finish(NttState &S, const NttLaunch &L, const NttJob &J, const PrimeDev &P, u32 tid, u64 *sm) {
    xchg_read_a<LOGN>(S.x, sm, tid);
    inv_pass_a<LOGN>(S.x, P.itw, SPLIT ? 2 + J.h : 1, P.p, S.b);
    if (SPLIT) {  // leave < 8p values for the last-stage kernel
#pragma unroll
      for (int k = 0; k < 32; k++) J.dst[(size_t)J.h * G::N + idx_a<LOGN>(tid, k)] = S.x[k];
      return;
    }
    scale_canon(S.x, P.ninv, P.ninv_s, P.p);
    const u64 half = P.p >> 1;
#pragma unroll
    for (int k = 0; k < 32; k++) {
      u64 v = S.x[k];
      if (EPI == EPI_ADDHALF) v = addmod(v, half, P.p);
      J.dst[idx_a<LOGN>(tid, k)] = v;
    }
  }
};

// last stage of a split inverse transform of length 2N2 = 2^(LOGN+1):
// X' = (X + Y) * ninv, Y' = (X - Y) * itw[1] * ninv, inputs < 8p.
EVAB_HD void inv_last_stage_elem(const NttLaunch &L, const NttJob &J, u32 i, u32 half_n) {
  const PrimeDev P = L.primes[J.pi];
  const u64x2 w = ldg_tw(P.itw + 1);
  u64 X = J.dst[i], Y = J.dst[i + half_n];
  u64 s = X + Y;                 // < 16p
  u64 d = X - Y + 8 * P.p;       // < 16p
  u64 a = shoup_mul(s, P.ninv, P.ninv_s, P.p);
  u64 b = shoup_mul(shoup_lazy(d, w.x, w.y, P.p), P.ninv, P.ninv_s, P.p);
  if (L.epi == EPI_ADDHALF) { const u64 h = P.p >> 1; a = addmod(a, h, P.p); b = addmod(b, h, P.p); }
  J.dst[i] = a; J.dst[i + half_n] = b;
}
