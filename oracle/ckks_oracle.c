/*
 * ckks_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 * See ckks_oracle.h for status ("parity unpinned" at the SEAL boundary) and
 * for who may link this.  Every function cites the reference call site in
 * /root/reference (EVA) whose SEAL 3.6 behaviour it restates, plus the
 * SURVEY.md Appendix A item describing that behaviour.
 */
#include "ckks_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

#define ORA_MAXK 24

/* ------------------------------------------------------------------ */
/* scalar modular arithmetic                                          */
/* ------------------------------------------------------------------ */
u64 ora_mulmod(u64 a, u64 b, u64 p) { return (u64)(((u128)a * b) % p); }

u64 ora_powmod(u64 a, u64 e, u64 p) {
    u64 r = 1 % p;
    a %= p;
    while (e) {
        if (e & 1) r = ora_mulmod(r, a, p);
        a = ora_mulmod(a, a, p);
        e >>= 1;
    }
    return r;
}

u64 ora_invmod(u64 a, u64 p) { return ora_powmod(a, p - 2, p); } /* p prime */

/* deterministic Miller-Rabin for u64 (SEAL uses randomized MR; same answer) */
int ora_is_prime(u64 n) {
    static const u64 bases[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return 0;
    for (int i = 0; i < 12; i++) {
        if (n == bases[i]) return 1;
        if (n % bases[i] == 0) return 0;
    }
    u64 d = n - 1;
    int r = 0;
    while ((d & 1) == 0) { d >>= 1; r++; }
    for (int i = 0; i < 12; i++) {
        u64 x = ora_powmod(bases[i], d, n);
        if (x == 1 || x == n - 1) continue;
        int comp = 1;
        for (int j = 1; j < r; j++) {
            x = ora_mulmod(x, x, n);
            if (x == n - 1) { comp = 0; break; }
        }
        if (comp) return 0;
    }
    return 1;
}

/* SEAL 3.6 CoeffModulus::Create, called at reference eva/seal/seal.cpp:181-182.
 * Appendix A.1: per distinct bit size b (appearing c times) scan
 * v = 2^b - 2N + 1, v -= 2N while v > 2^(b-1) collecting c primes (descending);
 * then each entry of bit_sizes, in order, takes the smallest unused prime of
 * its size. */
int ora_gen_primes(u64 N, const int *bit_sizes, int k, u64 *out) {
    u64 factor = 2 * N;
    int used_bits[ORA_MAXK], nb = 0;
    for (int i = 0; i < k; i++) {
        int b = bit_sizes[i], seen = 0;
        if (b < 2 || b > 60) return -1;
        for (int j = 0; j < nb; j++) if (used_bits[j] == b) seen = 1;
        if (!seen) used_bits[nb++] = b;
    }
    for (int bi = 0; bi < nb; bi++) {
        int b = used_bits[bi], count = 0;
        for (int i = 0; i < k; i++) if (bit_sizes[i] == b) count++;
        u64 found[ORA_MAXK];
        int nf = 0;
        u64 value = (((u64)1 << b) - 1) / factor * factor + 1;
        u64 lower = (u64)1 << (b - 1);
        while (nf < count && value > lower) {
            if (ora_is_prime(value)) found[nf++] = value;
            value -= factor;
        }
        if (nf < count) return -2;
        /* hand out from the back (smallest first) in bit_sizes order */
        int next = nf - 1;
        for (int i = 0; i < k; i++)
            if (bit_sizes[i] == b) out[i] = found[next--];
    }
    return 0;
}

/* SEAL try_minimal_primitive_root(2N, q): smallest primitive 2N-th root.
 * Appendix A.3.  Observable through every NTT-form value. */
u64 ora_min_primitive_root(u64 N, u64 p) {
    u64 two_n = 2 * N;
    if ((p - 1) % two_n) return 0;
    u64 e = (p - 1) / two_n, root = 0;
    for (u64 g = 2; g < p; g++) {
        u64 r = ora_powmod(g, e, p);
        if (ora_powmod(r, N, p) == p - 1) { root = r; break; }
    }
    u64 sq = ora_mulmod(root, root, p), cur = root, best = root;
    for (u64 i = 0; i < N; i++) { /* all odd powers = all primitive roots */
        if (cur < best) best = cur;
        cur = ora_mulmod(cur, sq, p);
    }
    return best;
}

/* ------------------------------------------------------------------ */
/* context                                                            */
/* ------------------------------------------------------------------ */
typedef struct {
    u64 p;
    u64 psi;
    u64 ratio_hi, ratio_lo; /* floor(2^128 / p) */
    u64 *w, *ws;            /* psi^bitrev(i), Shoup companion */
    u64 *iw, *iws;          /* psi^-bitrev(i) */
    u64 ninv, ninv_s;
} ora_prime;

struct ora_ctx {
    u64 N;
    int logN, k;
    ora_prime pr[ORA_MAXK];
    u64 inv[ORA_MAXK][ORA_MAXK]; /* inv[j][i] = q_j^-1 mod q_i (i != j) */
    /* FP64 encoder tables (CKKSEncoder, Appendix A.9) */
    double *root_re, *root_im;  /* root_powers_[i] = zeta^bitrev(i) */
    uint32_t *slot_index;       /* matrix_reps_index_map_, size N */
};

static inline u64 bitrev(u64 x, int bits) {
    u64 r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
static inline u64 shoup_pre(u64 w, u64 p) { return (u64)(((u128)w << 64) / p); }
/* lazy Shoup: result in [0,2p) for any y */
static inline u64 shoup_mul_lazy(u64 y, u64 w, u64 ws, u64 p) {
    u64 q = (u64)(((u128)ws * y) >> 64);
    return w * y - q * p;
}
/* Barrett reduction of a 128-bit value (SEAL barrett_reduce_128) */
static inline u64 barrett128(u128 z, const ora_prime *m) {
    u64 z0 = (u64)z, z1 = (u64)(z >> 64);
    /* q = floor(z * ratio / 2^128), computed like SEAL */
    u64 carry = (u64)(((u128)z0 * m->ratio_lo) >> 64);
    u128 t = (u128)z0 * m->ratio_hi;
    u64 tmp1 = (u64)t + carry;
    u64 tmp3 = (u64)(t >> 64) + (tmp1 < carry);
    t = (u128)z1 * m->ratio_lo;
    u64 tmp1b = tmp1 + (u64)t;
    carry = (u64)(t >> 64) + (tmp1b < tmp1);
    u64 q = z1 * m->ratio_hi + tmp3 + carry;
    u64 r = z0 - q * m->p;
    return r >= m->p ? r - m->p : r;
}
static inline u64 mm(u64 a, u64 b, const ora_prime *m) { return barrett128((u128)a * b, m); }
static inline u64 addm(u64 a, u64 b, u64 p) { u64 s = a + b; return s >= p ? s - p : s; }
static inline u64 subm(u64 a, u64 b, u64 p) { return a >= b ? a - b : a + p - b; }
static inline u64 negm(u64 a, u64 p) { return a ? p - a : 0; }

static int ilog2(u64 n) { int l = 0; while (((u64)1 << l) < n) l++; return l; }

static void complex_root(u64 index, u64 degree, double *re, double *im) {
    /* SEAL ComplexRoots::get_root: table over the first octant + symmetries */
    const double PI = 3.1415926535897932384626433832795028842;
    index &= degree - 1;
    if (index <= degree / 8) {
        double ang = 2.0 * PI * (double)index / (double)degree;
        *re = cos(ang); *im = sin(ang);
    } else if (index <= degree / 4) {
        double a, b; complex_root(degree / 4 - index, degree, &a, &b); *re = b; *im = a;
    } else if (index <= degree / 2) {
        double a, b; complex_root(degree / 2 - index, degree, &a, &b); *re = -a; *im = b;
    } else if (index <= 3 * degree / 4) {
        double a, b; complex_root(index - degree / 2, degree, &a, &b); *re = -a; *im = -b;
    } else {
        double a, b; complex_root(degree - index, degree, &a, &b); *re = a; *im = -b;
    }
}

ora_ctx *ora_ctx_create_from_primes(u64 N, const u64 *primes, int k) {
    if (k < 1 || k > ORA_MAXK || N < 2 || (N & (N - 1))) return NULL;
    ora_ctx *c = (ora_ctx *)calloc(1, sizeof(ora_ctx));
    c->N = N; c->logN = ilog2(N); c->k = k;
    for (int i = 0; i < k; i++) {
        ora_prime *m = &c->pr[i];
        m->p = primes[i];
        u128 num = ~(u128)0; /* floor(2^128/p) == floor((2^128-1)/p) for p not a power of 2 */
        u128 ratio = num / m->p;
        m->ratio_lo = (u64)ratio; m->ratio_hi = (u64)(ratio >> 64);
        m->psi = ora_min_primitive_root(N, m->p);
        if (!m->psi) { ora_ctx_destroy(c); return NULL; }
        m->w = (u64 *)malloc(4 * N * sizeof(u64));
        m->ws = m->w + N; m->iw = m->w + 2 * N; m->iws = m->w + 3 * N;
        u64 ipsi = ora_invmod(m->psi, m->p), pw = 1, ipw = 1;
        for (u64 j = 0; j < N; j++) {
            u64 r = bitrev(j, c->logN);
            m->w[r] = pw; m->ws[r] = shoup_pre(pw, m->p);
            m->iw[r] = ipw; m->iws[r] = shoup_pre(ipw, m->p);
            pw = ora_mulmod(pw, m->psi, m->p);
            ipw = ora_mulmod(ipw, ipsi, m->p);
        }
        m->ninv = ora_invmod(N % m->p, m->p);
        m->ninv_s = shoup_pre(m->ninv, m->p);
    }
    for (int j = 0; j < k; j++)
        for (int i = 0; i < k; i++)
            if (i != j) c->inv[j][i] = ora_invmod(c->pr[j].p % c->pr[i].p, c->pr[i].p);
    /* encoder tables */
    c->root_re = (double *)malloc(2 * N * sizeof(double));
    c->root_im = c->root_re + N;
    for (u64 i = 0; i < N; i++)
        complex_root(bitrev(i, c->logN), 2 * N, &c->root_re[i], &c->root_im[i]);
    c->slot_index = (uint32_t *)malloc(N * sizeof(uint32_t));
    u64 slots = N / 2, m2 = 2 * N, pos = 1;
    for (u64 i = 0; i < slots; i++) {
        u64 i1 = (pos - 1) >> 1, i2 = (m2 - pos - 1) >> 1;
        c->slot_index[i] = (uint32_t)bitrev(i1, c->logN);
        c->slot_index[slots | i] = (uint32_t)bitrev(i2, c->logN);
        pos = (pos * 3) & (m2 - 1);
    }
    return c;
}

ora_ctx *ora_ctx_create(u64 N, const int *bit_sizes, int k) {
    u64 primes[ORA_MAXK];
    if (k < 1 || k > ORA_MAXK) return NULL;
    if (ora_gen_primes(N, bit_sizes, k, primes)) return NULL;
    return ora_ctx_create_from_primes(N, primes, k);
}

void ora_ctx_destroy(ora_ctx *c) {
    if (!c) return;
    for (int i = 0; i < c->k; i++) free(c->pr[i].w);
    free(c->root_re); free(c->slot_index); free(c);
}
u64 ora_ctx_N(const ora_ctx *c) { return c->N; }
int ora_ctx_k(const ora_ctx *c) { return c->k; }
u64 ora_ctx_prime(const ora_ctx *c, int i) { return c->pr[i].p; }
u64 ora_ctx_psi(const ora_ctx *c, int i) { return c->pr[i].psi; }

/* ------------------------------------------------------------------ */
/* NTT (Appendix A.3).  Harvey lazy butterflies; canonical output.     */
/* ------------------------------------------------------------------ */
void ora_ntt_fwd(const ora_ctx *c, int pi, u64 *a) {
    const ora_prime *m = &c->pr[pi];
    const u64 p = m->p, two_p = 2 * p, N = c->N;
    u64 t = N;
    for (u64 mm_ = 1; mm_ < N; mm_ <<= 1) {
        t >>= 1;
        for (u64 i = 0; i < mm_; i++) {
            const u64 w = m->w[mm_ + i], ws = m->ws[mm_ + i];
            u64 *x = a + 2 * i * t, *y = x + t;
            for (u64 j = 0; j < t; j++) {
                u64 X = x[j]; if (X >= two_p) X -= two_p;
                u64 T = shoup_mul_lazy(y[j], w, ws, p);
                x[j] = X + T;
                y[j] = X - T + two_p;
            }
        }
    }
    for (u64 j = 0; j < N; j++) {
        u64 v = a[j];
        if (v >= two_p) v -= two_p;
        if (v >= p) v -= p;
        a[j] = v;
    }
}

void ora_ntt_inv(const ora_ctx *c, int pi, u64 *a) {
    const ora_prime *m = &c->pr[pi];
    const u64 p = m->p, two_p = 2 * p, N = c->N;
    u64 t = 1;
    for (u64 mm_ = N >> 1; mm_ >= 1; mm_ >>= 1) {
        for (u64 i = 0; i < mm_; i++) {
            const u64 w = m->iw[mm_ + i], ws = m->iws[mm_ + i];
            u64 *x = a + 2 * i * t, *y = x + t;
            for (u64 j = 0; j < t; j++) {
                u64 X = x[j], Y = y[j];
                u64 S = X + Y; if (S >= two_p) S -= two_p;
                u64 D = X - Y + two_p;
                x[j] = S;
                y[j] = shoup_mul_lazy(D, w, ws, p);
            }
        }
        t <<= 1;
    }
    for (u64 j = 0; j < N; j++) {
        u64 v = shoup_mul_lazy(a[j], m->ninv, m->ninv_s, p);
        if (v >= p) v -= p;
        a[j] = v;
    }
}

/* ------------------------------------------------------------------ */
/* dyadic evaluator ops (Appendix A.4)                                 */
/* ------------------------------------------------------------------ */
/* Evaluator::add  -- reference eva/seal/seal_executor.h:124 */
int ora_add(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa, const u64 *b, int sb) {
    const u64 N = c->N;
    int smax = sa > sb ? sa : sb, smin = sa < sb ? sa : sb;
    for (int s = 0; s < smax; s++)
        for (int i = 0; i < ell; i++) {
            const u64 p = c->pr[i].p;
            size_t o = ((size_t)s * ell + i) * N;
            if (s < smin) for (u64 j = 0; j < N; j++) out[o + j] = addm(a[o + j], b[o + j], p);
            else { const u64 *src = sa > sb ? a : b; for (u64 j = 0; j < N; j++) out[o + j] = src[o + j]; }
        }
    return 0;
}
/* Evaluator::sub  -- reference eva/seal/seal_executor.h:140 */
int ora_sub(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa, const u64 *b, int sb) {
    const u64 N = c->N;
    int smax = sa > sb ? sa : sb, smin = sa < sb ? sa : sb;
    for (int s = 0; s < smax; s++)
        for (int i = 0; i < ell; i++) {
            const u64 p = c->pr[i].p;
            size_t o = ((size_t)s * ell + i) * N;
            if (s < smin) for (u64 j = 0; j < N; j++) out[o + j] = subm(a[o + j], b[o + j], p);
            else if (sa > sb) for (u64 j = 0; j < N; j++) out[o + j] = a[o + j];
            else for (u64 j = 0; j < N; j++) out[o + j] = negm(b[o + j], p);
        }
    return 0;
}
/* Evaluator::add_plain -- reference eva/seal/seal_executor.h:127 */
int ora_add_plain(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa, const u64 *pt) {
    const u64 N = c->N;
    if (out != a) memcpy(out, a, (size_t)sa * ell * N * sizeof(u64));
    for (int i = 0; i < ell; i++) {
        const u64 p = c->pr[i].p;
        for (u64 j = 0; j < N; j++) out[i * N + j] = addm(a[i * N + j], pt[i * N + j], p);
    }
    return 0;
}
/* Evaluator::sub_plain -- reference eva/seal/seal_executor.h:143 */
int ora_sub_plain(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa, const u64 *pt) {
    const u64 N = c->N;
    if (out != a) memcpy(out, a, (size_t)sa * ell * N * sizeof(u64));
    for (int i = 0; i < ell; i++) {
        const u64 p = c->pr[i].p;
        for (u64 j = 0; j < N; j++) out[i * N + j] = subm(a[i * N + j], pt[i * N + j], p);
    }
    return 0;
}
/* Evaluator::negate -- reference eva/seal/seal_executor.h:194 */
int ora_negate(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa) {
    const u64 N = c->N;
    for (int s = 0; s < sa; s++)
        for (int i = 0; i < ell; i++) {
            const u64 p = c->pr[i].p;
            size_t o = ((size_t)s * ell + i) * N;
            for (u64 j = 0; j < N; j++) out[o + j] = negm(a[o + j], p);
        }
    return 0;
}
/* Evaluator::multiply_plain -- reference eva/seal/seal_executor.h:168 */
int ora_mul_plain(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa, const u64 *pt) {
    const u64 N = c->N;
    for (int s = 0; s < sa; s++)
        for (int i = 0; i < ell; i++) {
            const ora_prime *m = &c->pr[i];
            size_t o = ((size_t)s * ell + i) * N;
            for (u64 j = 0; j < N; j++) out[o + j] = mm(a[o + j], pt[i * N + j], m);
        }
    return 0;
}
/* Evaluator::multiply (2x2->3) -- reference eva/seal/seal_executor.h:164 */
int ora_mul(const ora_ctx *c, int ell, u64 *out, const u64 *a, const u64 *b) {
    const u64 N = c->N;
    size_t P = (size_t)ell * N;
    for (int i = 0; i < ell; i++) {
        const ora_prime *m = &c->pr[i];
        for (u64 j = 0; j < N; j++) {
            size_t o = (size_t)i * N + j;
            u64 a0 = a[o], a1 = a[P + o], b0 = b[o], b1 = b[P + o];
            u64 d0 = mm(a0, b0, m);
            u64 d1 = addm(mm(a0, b1, m), mm(a1, b0, m), m->p);
            u64 d2 = mm(a1, b1, m);
            out[o] = d0; out[P + o] = d1; out[2 * P + o] = d2;
        }
    }
    return 0;
}
/* Evaluator::square -- reference eva/seal/seal_executor.h:162 */
int ora_square(const ora_ctx *c, int ell, u64 *out, const u64 *a) {
    const u64 N = c->N;
    size_t P = (size_t)ell * N;
    for (int i = 0; i < ell; i++) {
        const ora_prime *m = &c->pr[i];
        for (u64 j = 0; j < N; j++) {
            size_t o = (size_t)i * N + j;
            u64 a0 = a[o], a1 = a[P + o];
            u64 x = mm(a0, a1, m);
            out[o] = mm(a0, a0, m); out[P + o] = addm(x, x, m->p); out[2 * P + o] = mm(a1, a1, m);
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* RNS divide-and-round by the last prime (shared by rescale, the      */
/* key-switch mod-down and encryption).  poly holds nres residues in   */
/* NTT form, the divisor is prime index `last`; the other residues use */
/* prime indices idx[0..nres-2].  Result overwrites the first nres-1.  */
/* SEAL RNSTool::divide_and_round_q_last_ntt_inplace.  Appendix A.6.   */
/* ------------------------------------------------------------------ */
static void divround_last_ntt(const ora_ctx *c, u64 *poly, int nres, const int *idx, int last, u64 *tmp) {
    const u64 N = c->N;
    const u64 ql = c->pr[last].p, half = ql >> 1;
    u64 *r = poly + (size_t)(nres - 1) * N;
    ora_ntt_inv(c, last, r);
    for (u64 j = 0; j < N; j++) r[j] = addm(r[j], half, ql); /* (x + floor(q/2)) mod q */
    for (int t = 0; t < nres - 1; t++) {
        const int i = idx[t];
        const ora_prime *m = &c->pr[i];
        const u64 p = m->p, half_i = half % p, qinv = c->inv[last][i];
        for (u64 j = 0; j < N; j++) tmp[j] = subm(r[j] % p, half_i, p);
        ora_ntt_fwd(c, i, tmp);
        u64 *x = poly + (size_t)t * N;
        for (u64 j = 0; j < N; j++) x[j] = mm(subm(x[j], tmp[j], p), qinv, m);
    }
}

/* Evaluator::rescale_to_next -- reference eva/seal/seal_executor.h:213 */
int ora_rescale(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa) {
    const u64 N = c->N;
    if (ell < 2) return -1;
    int idx[ORA_MAXK];
    for (int i = 0; i < ell; i++) idx[i] = i;
    u64 *buf = (u64 *)malloc(((size_t)ell + 1) * N * sizeof(u64));
    for (int s = 0; s < sa; s++) {
        memcpy(buf, a + (size_t)s * ell * N, (size_t)ell * N * sizeof(u64));
        divround_last_ntt(c, buf, ell, idx, ell - 1, buf + (size_t)ell * N);
        memcpy(out + (size_t)s * (ell - 1) * N, buf, (size_t)(ell - 1) * N * sizeof(u64));
    }
    free(buf);
    return 0;
}

/* Evaluator::mod_switch_to_next -- reference eva/seal/seal_executor.h:206 */
int ora_mod_switch(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa) {
    const u64 N = c->N;
    if (ell < 2) return -1;
    for (int s = 0; s < sa; s++)
        memmove(out + (size_t)s * (ell - 1) * N, a + (size_t)s * ell * N, (size_t)(ell - 1) * N * sizeof(u64));
    return 0;
}

/* ------------------------------------------------------------------ */
/* key switching (Appendix A.5).  Evaluator::switch_key_inplace, used  */
/* by relinearize (seal_executor.h:200) and rotate_vector (:181,:188). */
/* ------------------------------------------------------------------ */
int ora_keyswitch(const ora_ctx *c, int ell, u64 *ks, const u64 *t, const u64 *key) {
    const u64 N = c->N;
    const int k = c->k, sp = k - 1; /* special prime index */
    if (ell < 1 || ell > k - 1) return -1;
    const int nm = ell + 1;          /* output moduli q_0..q_{ell-1}, P */
    u64 *that = (u64 *)malloc((size_t)ell * N * sizeof(u64));           /* step 1 */
    u64 *acc = (u64 *)calloc((size_t)2 * nm * N, sizeof(u64));          /* [c][m][N] */
    u64 *tmp = (u64 *)malloc((size_t)N * sizeof(u64));
    memcpy(that, t, (size_t)ell * N * sizeof(u64));
    for (int J = 0; J < ell; J++) ora_ntt_inv(c, J, that + (size_t)J * N);
    for (int mi = 0; mi < nm; mi++) {
        const int ki = (mi == ell) ? sp : mi; /* key row / prime index */
        const ora_prime *m = &c->pr[ki];
        for (int J = 0; J < ell; J++) {
            const u64 *opnd;
            if (ki == J) opnd = t + (size_t)J * N;
            else {
                const u64 *src = that + (size_t)J * N;
                for (u64 j = 0; j < N; j++) tmp[j] = src[j] % m->p;
                ora_ntt_fwd(c, ki, tmp);
                opnd = tmp;
            }
            for (int cc = 0; cc < 2; cc++) {
                const u64 *krow = key + (((size_t)J * 2 + cc) * k + ki) * N;
                u64 *a = acc + ((size_t)cc * nm + mi) * N;
                for (u64 j = 0; j < N; j++) a[j] = addm(a[j], mm(opnd[j], krow[j], m), m->p);
            }
        }
    }
    /* step 3: mod-down by P with rounding */
    int idx[ORA_MAXK];
    for (int i = 0; i < ell; i++) idx[i] = i;
    for (int cc = 0; cc < 2; cc++) {
        u64 *poly = acc + (size_t)cc * nm * N;
        divround_last_ntt(c, poly, nm, idx, sp, tmp);
        memcpy(ks + (size_t)cc * ell * N, poly, (size_t)ell * N * sizeof(u64));
    }
    free(that); free(acc); free(tmp);
    return 0;
}

/* Evaluator::relinearize (size 3 -> 2) -- reference seal_executor.h:200 */
int ora_relinearize(const ora_ctx *c, int ell, u64 *out, const u64 *a, const u64 *rk) {
    const u64 N = c->N;
    size_t P = (size_t)ell * N;
    u64 *ks = (u64 *)malloc(2 * P * sizeof(u64));
    int rc = ora_keyswitch(c, ell, ks, a + 2 * P, rk);
    if (!rc)
        for (int cc = 0; cc < 2; cc++)
            for (int i = 0; i < ell; i++) {
                const u64 p = c->pr[i].p;
                size_t o = (size_t)cc * P + (size_t)i * N;
                for (u64 j = 0; j < N; j++) out[o + j] = addm(a[o + j], ks[o + j], p);
            }
    free(ks);
    return rc;
}

/* GaloisTool::get_elt_from_step (Appendix A.8) */
u64 ora_galois_elt_from_step(u64 N, int steps) {
    u64 m = 2 * N;
    if (steps == 0) return m - 1;
    u64 pos = steps < 0 ? (u64)(-(long long)steps) : (u64)steps;
    if (pos >= N / 2) return 0;
    u64 s = steps < 0 ? N / 2 - pos : pos, g = 1;
    for (u64 i = 0; i < s; i++) g = (g * 3) & (m - 1);
    return g;
}
/* GaloisTool::generate_table_ntt (Appendix A.8) */
void ora_galois_table(u64 N, u64 elt, uint32_t *table) {
    int logN = ilog2(N);
    for (u64 i = 0; i < N; i++) {
        u64 r = bitrev(i, logN);
        u64 raw = ((elt * (2 * r + 1)) >> 1) & (N - 1);
        table[i] = (uint32_t)bitrev(raw, logN);
    }
}
int ora_apply_galois(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa, u64 elt) {
    const u64 N = c->N;
    uint32_t *tab = (uint32_t *)malloc(N * sizeof(uint32_t));
    ora_galois_table(N, elt, tab);
    for (int s = 0; s < sa; s++)
        for (int i = 0; i < ell; i++) {
            size_t o = ((size_t)s * ell + i) * N;
            for (u64 j = 0; j < N; j++) out[o + j] = a[o + tab[j]];
        }
    free(tab);
    return 0;
}
/* Evaluator::rotate_vector -> apply_galois_inplace -- seal_executor.h:181,188 */
int ora_rotate(const ora_ctx *c, int ell, u64 *out, const u64 *a, u64 elt, const u64 *gk) {
    const u64 N = c->N;
    size_t P = (size_t)ell * N;
    u64 *perm = (u64 *)malloc(2 * P * sizeof(u64));
    u64 *ks = (u64 *)malloc(2 * P * sizeof(u64));
    ora_apply_galois(c, ell, perm, a, 2, elt);
    int rc = ora_keyswitch(c, ell, ks, perm + P, gk);
    if (!rc)
        for (int i = 0; i < ell; i++) {
            const u64 p = c->pr[i].p;
            size_t o = (size_t)i * N;
            for (u64 j = 0; j < N; j++) {
                out[o + j] = addm(perm[o + j], ks[o + j], p);
                out[P + o + j] = ks[P + o + j];
            }
        }
    free(perm); free(ks);
    return rc;
}

/* ------------------------------------------------------------------ */
/* CKKS encoder (Appendix A.9) -- CKKSEncoder::encode at               */
/* seal_executor.h:242 and seal.cpp:68,80; decode at seal.cpp:133,136. */
/* FP64; compile with -ffp-contract=off so results are reproducible.   */
/* ------------------------------------------------------------------ */
int ora_encode(const ora_ctx *c, int ell, const double *values, size_t nvalues, double scale, u64 *pt) {
    const u64 N = c->N, slots = N / 2;
    if (nvalues > slots) return -1;
    double *re = (double *)calloc(2 * N, sizeof(double)), *im = re + N;
    for (size_t i = 0; i < nvalues; i++) {
        re[c->slot_index[i]] = values[i];           im[c->slot_index[i]] = 0.0;
        re[c->slot_index[slots + i]] = values[i];   im[c->slot_index[slots + i]] = -0.0;
    }
    /* inverse FFT, Gentleman-Sande from bit-reversed order, conj twiddles */
    u64 gap = 1;
    for (u64 m = N >> 1; m >= 1; m >>= 1) {
        for (u64 i = 0; i < m; i++) {
            const double wr = c->root_re[m + i], wi = -c->root_im[m + i];
            u64 base = 2 * i * gap;
            for (u64 j = base; j < base + gap; j++) {
                double ur = re[j], ui = im[j], vr = re[j + gap], vi = im[j + gap];
                re[j] = ur + vr; im[j] = ui + vi;
                double dr = ur - vr, di = ui - vi;
                re[j + gap] = dr * wr - di * wi;
                im[j + gap] = dr * wi + di * wr;
            }
        }
        gap <<= 1;
    }
    const double fix = scale / (double)N;
    for (u64 j = 0; j < N; j++) {
        double cd = round(re[j] * fix);
        int neg = signbit(cd);
        double mag = fabs(cd);
        int ex;
        double fr = frexp(mag, &ex);              /* mag = fr * 2^ex, fr in [0.5,1) */
        u64 mant; int sh;
        if (mag < 18446744073709551616.0) { mant = (u64)mag; sh = 0; }
        else { mant = (u64)ldexp(fr, 64); sh = ex - 64; } /* exact: mag = mant * 2^sh */
        for (int i = 0; i < ell; i++) {
            const u64 p = c->pr[i].p;
            u64 v = mant % p;
            if (sh) v = ora_mulmod(v, ora_powmod(2, (u64)sh, p), p);
            pt[(size_t)i * N + j] = neg ? negm(v, p) : v;
        }
    }
    for (int i = 0; i < ell; i++) ora_ntt_fwd(c, i, pt + (size_t)i * N);
    free(re);
    return 0;
}

/* small fixed-width big integers for CRT composition in decode */
#define BW 32
static void big_mul_u64(u64 *a, int n, u64 m) { /* a *= m */
    u64 carry = 0;
    for (int i = 0; i < n; i++) { u128 t = (u128)a[i] * m + carry; a[i] = (u64)t; carry = (u64)(t >> 64); }
}
static void big_addmul(u64 *acc, const u64 *a, int n, u64 m) { /* acc += a*m */
    u64 carry = 0;
    for (int i = 0; i < n; i++) { u128 t = (u128)a[i] * m + acc[i] + carry; acc[i] = (u64)t; carry = (u64)(t >> 64); }
}
static int big_cmp(const u64 *a, const u64 *b, int n) {
    for (int i = n - 1; i >= 0; i--) { if (a[i] != b[i]) return a[i] > b[i] ? 1 : -1; }
    return 0;
}
static void big_sub(u64 *a, const u64 *b, int n) { /* a -= b */
    u64 br = 0;
    for (int i = 0; i < n; i++) { u128 t = (u128)a[i] - b[i] - br; a[i] = (u64)t; br = (u64)(t >> 64) & 1; }
}
static u64 big_mod_u64(const u64 *a, int n, u64 p) {
    u64 r = 0;
    for (int i = n - 1; i >= 0; i--) r = (u64)((((u128)r << 64) | a[i]) % p);
    return r;
}

int ora_decode(const ora_ctx *c, int ell, const u64 *pt, double scale, double *values) {
    const u64 N = c->N, slots = N / 2;
    const int nw = ell + 1;
    if (nw > BW) return -1;
    u64 *coef = (u64 *)malloc((size_t)ell * N * sizeof(u64));
    memcpy(coef, pt, (size_t)ell * N * sizeof(u64));
    for (int i = 0; i < ell; i++) ora_ntt_inv(c, i, coef + (size_t)i * N);
    /* CRT constants */
    u64 Q[BW] = {0}, punct[ORA_MAXK][BW], ipunct[ORA_MAXK], halfQ[BW];
    Q[0] = 1;
    for (int i = 0; i < ell; i++) big_mul_u64(Q, nw, c->pr[i].p);
    for (int i = 0; i < ell; i++) {
        memset(punct[i], 0, sizeof(punct[i])); punct[i][0] = 1;
        for (int j = 0; j < ell; j++) if (j != i) big_mul_u64(punct[i], nw, c->pr[j].p);
        ipunct[i] = ora_invmod(big_mod_u64(punct[i], nw, c->pr[i].p), c->pr[i].p);
    }
    /* halfQ = (Q+1)/2 = upper_half_threshold */
    memcpy(halfQ, Q, sizeof(halfQ));
    { u64 carry = 1; for (int i = 0; i < nw && carry; i++) { halfQ[i] += carry; carry = (halfQ[i] == 0); } }
    for (int i = 0; i < nw; i++) halfQ[i] = (halfQ[i] >> 1) | (i + 1 < nw ? halfQ[i + 1] << 63 : 0);
    double *re = (double *)calloc(2 * N, sizeof(double)), *im = re + N;
    const double inv_scale = 1.0 / scale, two64 = 18446744073709551616.0;
    for (u64 j = 0; j < N; j++) {
        u64 X[BW] = {0};
        for (int i = 0; i < ell; i++) {
            u64 v = ora_mulmod(coef[(size_t)i * N + j], ipunct[i], c->pr[i].p);
            big_addmul(X, punct[i], nw, v);
        }
        while (big_cmp(X, Q, nw) >= 0) big_sub(X, Q, nw);
        int neg = big_cmp(X, halfQ, nw) >= 0;
        if (neg) { u64 T[BW]; memcpy(T, Q, sizeof(T)); big_sub(T, X, nw); memcpy(X, T, sizeof(T)); }
        double acc = 0.0, f = inv_scale;
        for (int i = 0; i < nw; i++) { if (X[i]) acc += (double)X[i] * f; f *= two64; }
        re[j] = neg ? -acc : acc;
    }
    /* forward FFT (Cooley-Tukey, natural in -> bit-reversed out) */
    u64 t = N;
    for (u64 m = 1; m < N; m <<= 1) {
        t >>= 1;
        for (u64 i = 0; i < m; i++) {
            const double wr = c->root_re[m + i], wi = c->root_im[m + i];
            u64 base = 2 * i * t;
            for (u64 j = base; j < base + t; j++) {
                double vr = re[j + t] * wr - im[j + t] * wi;
                double vi = re[j + t] * wi + im[j + t] * wr;
                double ur = re[j], ui = im[j];
                re[j] = ur + vr; im[j] = ui + vi;
                re[j + t] = ur - vr; im[j + t] = ui - vi;
            }
        }
    }
    for (u64 i = 0; i < slots; i++) values[i] = re[c->slot_index[i]];
    free(re); free(coef);
    return 0;
}

/* ------------------------------------------------------------------ */
/* client side: keygen / encrypt / decrypt (Appendix A.11).            */
/* KeyGenerator at reference eva/seal/seal.cpp:186-200; Encryptor at   */
/* seal.cpp:85; Decryptor at seal.cpp:132.  Randomness is this file's  */
/* own seeded PRNG: nothing in the reference pins it.                  */
/* ------------------------------------------------------------------ */
typedef struct { u64 s; } ora_rng;
static u64 rng_next(ora_rng *r) { /* SplitMix64 */
    u64 z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static u64 rng_uniform(ora_rng *r, u64 p) { /* rejection sampling in [0,p) */
    u64 lim = UINT64_MAX - (UINT64_MAX % p) - 1, v;
    do { v = rng_next(r); } while (v > lim);
    return v % p;
}
/* signed small poly -> residues (first nres primes), coefficient form */
static void small_to_rns(const ora_ctx *c, const int *v, int nres, u64 *out) {
    const u64 N = c->N;
    for (int i = 0; i < nres; i++) {
        const u64 p = c->pr[i].p;
        for (u64 j = 0; j < N; j++) out[(size_t)i * N + j] = v[j] >= 0 ? (u64)v[j] : p - (u64)(-v[j]);
    }
}
static void sample_ternary(ora_rng *r, u64 N, int *v) {
    for (u64 j = 0; j < N; j++) v[j] = (int)rng_uniform(r, 3) - 1;
}
static void sample_cbd(ora_rng *r, u64 N, int *v) { /* SEAL sample_poly_cbd: 21 - 21 bits */
    for (u64 j = 0; j < N; j++) {
        u64 x = rng_next(r);
        v[j] = __builtin_popcountll(x & 0x1FFFFF) - __builtin_popcountll((x >> 21) & 0x1FFFFF);
    }
}

typedef struct gk_node { u64 elt; u64 *key; struct gk_node *next; } gk_node;
struct ora_keys {
    const ora_ctx *c;
    ora_rng rng;
    u64 *sk;   /* [k][N] NTT */
    u64 *pk;   /* [2][k][N] NTT */
    u64 *rk;   /* [k-1][2][k][N] */
    gk_node *gks;
};

/* (c0,c1) = (-(a*s+e), a) at key level, NTT form: encrypt_zero_symmetric */
static void enc_zero_sym(ora_keys *K, u64 *c0, u64 *c1) {
    const ora_ctx *c = K->c; const u64 N = c->N; const int k = c->k;
    int *e = (int *)malloc(N * sizeof(int));
    sample_cbd(&K->rng, N, e);
    small_to_rns(c, e, k, c0);
    for (int i = 0; i < k; i++) {
        const ora_prime *m = &c->pr[i];
        u64 *e_i = c0 + (size_t)i * N, *a_i = c1 + (size_t)i * N;
        ora_ntt_fwd(c, i, e_i);
        for (u64 j = 0; j < N; j++) a_i[j] = rng_uniform(&K->rng, m->p);
        for (u64 j = 0; j < N; j++)
            e_i[j] = negm(addm(mm(a_i[j], K->sk[(size_t)i * N + j], m), e_i[j], m->p), m->p);
    }
    free(e);
}
/* KeyGenerator::generate_one_kswitch_key: key[J] = enc_zero + P*newkey on residue J */
static u64 *make_kswitch_key(ora_keys *K, const u64 *newkey /* [k][N] NTT */) {
    const ora_ctx *c = K->c; const u64 N = c->N; const int k = c->k;
    u64 *key = (u64 *)malloc((size_t)(k - 1) * 2 * k * N * sizeof(u64));
    const u64 P = c->pr[k - 1].p;
    for (int J = 0; J < k - 1; J++) {
        u64 *c0 = key + ((size_t)J * 2 + 0) * k * N, *c1 = key + ((size_t)J * 2 + 1) * k * N;
        enc_zero_sym(K, c0, c1);
        const ora_prime *m = &c->pr[J];
        u64 f = P % m->p;
        for (u64 j = 0; j < N; j++)
            c0[(size_t)J * N + j] = addm(c0[(size_t)J * N + j], mm(newkey[(size_t)J * N + j], f, m), m->p);
    }
    return key;
}

ora_keys *ora_keygen(const ora_ctx *c, u64 seed) {
    const u64 N = c->N; const int k = c->k;
    if (k < 2) return NULL;
    ora_keys *K = (ora_keys *)calloc(1, sizeof(ora_keys));
    K->c = c; K->rng.s = seed;
    int *s = (int *)malloc(N * sizeof(int));
    sample_ternary(&K->rng, N, s);
    K->sk = (u64 *)malloc((size_t)k * N * sizeof(u64));
    small_to_rns(c, s, k, K->sk);
    for (int i = 0; i < k; i++) ora_ntt_fwd(c, i, K->sk + (size_t)i * N);
    free(s);
    K->pk = (u64 *)malloc((size_t)2 * k * N * sizeof(u64));
    enc_zero_sym(K, K->pk, K->pk + (size_t)k * N);
    u64 *s2 = (u64 *)malloc((size_t)k * N * sizeof(u64));
    for (int i = 0; i < k; i++)
        for (u64 j = 0; j < N; j++) {
            u64 v = K->sk[(size_t)i * N + j];
            s2[(size_t)i * N + j] = mm(v, v, &c->pr[i]);
        }
    K->rk = make_kswitch_key(K, s2);
    free(s2);
    return K;
}
void ora_keys_destroy(ora_keys *K) {
    if (!K) return;
    free(K->sk); free(K->pk); free(K->rk);
    for (gk_node *g = K->gks; g;) { gk_node *n = g->next; free(g->key); free(g); g = n; }
    free(K);
}
const u64 *ora_keys_secret(const ora_keys *K) { return K->sk; }
const u64 *ora_keys_public(const ora_keys *K) { return K->pk; }
const u64 *ora_keys_relin(const ora_keys *K) { return K->rk; }
const u64 *ora_keys_galois(ora_keys *K, u64 elt) {
    for (gk_node *g = K->gks; g; g = g->next) if (g->elt == elt) return g->key;
    const ora_ctx *c = K->c; const u64 N = c->N; const int k = c->k;
    u64 *rot = (u64 *)malloc((size_t)k * N * sizeof(u64));
    ora_apply_galois(c, k, rot, K->sk, 1, elt);
    gk_node *g = (gk_node *)malloc(sizeof(gk_node));
    g->elt = elt; g->key = make_kswitch_key(K, rot); g->next = K->gks; K->gks = g;
    free(rot);
    return g->key;
}

/* Encryptor::encrypt (public key) -- encrypt zero one level up, divide and
 * round by that level's last prime, add the plaintext to c0. */
/* Encryptor::encrypt with the public key and EXPLICIT randomness (u ternary, e0 / e1 small): the deterministic core
 * of ora_encrypt, so that the product's encryptor can be compared bit for bit on the same (u, e0, e1). */
int ora_encrypt_with(const ora_keys *K, int ell, const u64 *pt, const int *u, const int *e0, const int *e1, u64 *ct) {
    const ora_ctx *c = K->c; const u64 N = c->N; const int k = c->k;
    if (ell < 1 || ell > k - 1) return -1;
    const int nres = ell + 1;
    u64 *ur = (u64 *)malloc((size_t)nres * N * sizeof(u64));
    u64 *buf = (u64 *)malloc(((size_t)nres + 1) * N * sizeof(u64));
    small_to_rns(c, u, nres, ur);
    for (int i = 0; i < nres; i++) ora_ntt_fwd(c, i, ur + (size_t)i * N);
    int idx[ORA_MAXK];
    for (int i = 0; i < nres; i++) idx[i] = i;
    for (int cc = 0; cc < 2; cc++) {
        small_to_rns(c, cc ? e1 : e0, nres, buf);
        for (int i = 0; i < nres; i++) {
            const ora_prime *m = &c->pr[i];
            u64 *x = buf + (size_t)i * N;
            ora_ntt_fwd(c, i, x);
            const u64 *pk = K->pk + ((size_t)cc * k + i) * N;
            for (u64 j = 0; j < N; j++) x[j] = addm(mm(pk[j], ur[(size_t)i * N + j], m), x[j], m->p);
        }
        divround_last_ntt(c, buf, nres, idx, nres - 1, buf + (size_t)nres * N);
        memcpy(ct + (size_t)cc * ell * N, buf, (size_t)ell * N * sizeof(u64));
    }
    for (int i = 0; i < ell; i++) {
        const u64 p = c->pr[i].p;
        for (u64 j = 0; j < N; j++) ct[(size_t)i * N + j] = addm(ct[(size_t)i * N + j], pt[(size_t)i * N + j], p);
    }
    free(ur); free(buf);
    return 0;
}
int ora_encrypt(const ora_keys *K, int ell, const u64 *pt, u64 seed, u64 *ct) {
    const ora_ctx *c = K->c; const u64 N = c->N;
    ora_rng rng; rng.s = seed ^ 0xC0FFEE1234ull;
    int *u = (int *)malloc(N * sizeof(int)), *e0 = (int *)malloc(N * sizeof(int)), *e1 = (int *)malloc(N * sizeof(int));
    sample_ternary(&rng, N, u);
    sample_cbd(&rng, N, e0);
    sample_cbd(&rng, N, e1);
    const int rc = ora_encrypt_with(K, ell, pt, u, e0, e1, ct);
    free(u); free(e0); free(e1);
    return rc;
}

/* Decryptor::decrypt (CKKS): pt = c0 + c1*s + c2*s^2 (NTT domain) */
int ora_decrypt(const ora_keys *K, int ell, const u64 *ct, int size, u64 *pt) {
    const ora_ctx *c = K->c; const u64 N = c->N;
    size_t P = (size_t)ell * N;
    for (int i = 0; i < ell; i++) {
        const ora_prime *m = &c->pr[i];
        for (u64 j = 0; j < N; j++) {
            size_t o = (size_t)i * N + j;
            u64 s = K->sk[o], acc = ct[o], sp = s;
            for (int t = 1; t < size; t++) {
                acc = addm(acc, mm(ct[(size_t)t * P + o], sp, m), m->p);
                sp = mm(sp, s, m);
            }
            pt[o] = acc;
        }
    }
    return 0;
}
