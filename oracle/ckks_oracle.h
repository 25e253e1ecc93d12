/*
 * ckks_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the arithmetic that microsoft/EVA's hot path
 * (SEALPublic::execute -> SEALExecutor::operator() -> seal::Evaluator /
 * seal::CKKSEncoder, reference eva/seal/seal.cpp:104-122 and
 * eva/seal/seal_executor.h:114-243) delegates to Microsoft SEAL 3.6.
 *
 * PARITY STATUS: "parity unpinned" at the SEAL boundary.  SEAL 3.6 is an
 * un-vendored dependency (reference CMakeLists.txt:24, README.md:28-36 pins
 * v3.6.4) and is absent from this environment; the reference holds no golden
 * ciphertext vectors.  This file restates SEAL 3.6's
 * published algorithms and is pinned instead by
 *   - mathematical ground truth (schoolbook negacyclic products, big-integer
 *     CRT rounding, direct polynomial evaluation at psi^(2*bitrev(i)+1)),
 *   - decrypt-consistency of every evaluator op,
 *   - the reference's own acceptance criterion (tests/common.py:34, MSE<0.01).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library.
 *
 * Data layouts (all little-endian u64, NTT form unless stated):
 *   ciphertext  [size][ell][N]        (SEAL Ciphertext::data())
 *   plaintext   [ell][N]              (SEAL Plaintext, CKKS)
 *   kswitch key [k-1 digits][2][k][N] (SEAL KSwitchKeys::data()[idx], T3)
 * where k = number of primes at key level, level j has ell = k-1-j residues.
 */
#ifndef CKKS_ORACLE_H
#define CKKS_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t u64;
typedef struct ora_ctx ora_ctx;

/* ---- scalar helpers (exported for tests) ---- */
u64 ora_mulmod(u64 a, u64 b, u64 p);
u64 ora_powmod(u64 a, u64 e, u64 p);
u64 ora_invmod(u64 a, u64 p);
int ora_is_prime(u64 n);
/* SEAL CoeffModulus::Create(N, bit_sizes) order (Appendix A.1). returns 0 ok */
int ora_gen_primes(u64 N, const int *bit_sizes, int k, u64 *out_primes);
/* numerically smallest primitive 2N-th root of unity mod p (A.3) */
u64 ora_min_primitive_root(u64 N, u64 p);

/* ---- context: primes, NTT tables, rescale/key-switch constants ---- */
ora_ctx *ora_ctx_create(u64 N, const int *bit_sizes, int k);
ora_ctx *ora_ctx_create_from_primes(u64 N, const u64 *primes, int k);
void ora_ctx_destroy(ora_ctx *c);
u64 ora_ctx_N(const ora_ctx *c);
int ora_ctx_k(const ora_ctx *c);
u64 ora_ctx_prime(const ora_ctx *c, int i);
u64 ora_ctx_psi(const ora_ctx *c, int i);

/* negacyclic NTT of one residue polynomial mod prime i; natural order in,
 * out[j] = a(psi^(2*bitrev(j)+1)), fully reduced (A.3). */
void ora_ntt_fwd(const ora_ctx *c, int prime_idx, u64 *a);
void ora_ntt_inv(const ora_ctx *c, int prime_idx, u64 *a);

/* ---- evaluator ops (Appendix A.4-A.8). ell = residues at the input level.
 * Outputs may alias inputs only where stated.  Return 0 on success. ---- */
/* out size = max(sa,sb); extra polys copied (negated for sub from b). A.4 */
int ora_add(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa, const u64 *b, int sb);
int ora_sub(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa, const u64 *b, int sb);
int ora_add_plain(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa, const u64 *pt);
int ora_sub_plain(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa, const u64 *pt);
int ora_negate(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa);
int ora_mul_plain(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa, const u64 *pt);
/* 2x2 -> 3 */
int ora_mul(const ora_ctx *c, int ell, u64 *out3, const u64 *a2, const u64 *b2);
int ora_square(const ora_ctx *c, int ell, u64 *out3, const u64 *a2);
/* key-switch core: target t[ell][N] (NTT form) with key[k-1][2][k][N];
 * writes ks[2][ell][N] (A.5 steps 1-3). */
int ora_keyswitch(const ora_ctx *c, int ell, u64 *ks2, const u64 *t, const u64 *key);
/* out2 = (a0+ks0, a1+ks1) where ks = keyswitch(a2, relin key).  A.5 step 4 */
int ora_relinearize(const ora_ctx *c, int ell, u64 *out2, const u64 *a3, const u64 *relin_key);
/* Galois element for rotate_vector(steps) (A.8); steps==0 -> 2N-1 */
u64 ora_galois_elt_from_step(u64 N, int steps);
void ora_galois_table(u64 N, u64 elt, uint32_t *table); /* out[i]=in[table[i]] */
int ora_apply_galois(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa, u64 elt);
/* rotate_vector with the direct key for elt (steps!=0). A.8 */
int ora_rotate(const ora_ctx *c, int ell, u64 *out2, const u64 *a2, u64 elt, const u64 *galois_key);
/* rescale_to_next: [s][ell][N] -> [s][ell-1][N] (A.6) */
int ora_rescale(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa);
/* mod_switch_to_next: drop last residue (A.7) */
int ora_mod_switch(const ora_ctx *c, int ell, u64 *out, const u64 *a, int sa);

/* ---- CKKS encoder (A.9, FP64; not bit-pinned by the reference) ---- */
/* values[slots=N/2] real -> pt[ell][N] NTT form; scale = absolute (2^bits) */
int ora_encode(const ora_ctx *c, int ell, const double *values, size_t nvalues, double scale, u64 *pt);
/* pt[ell][N] NTT form -> values[N/2] (real parts) */
int ora_decode(const ora_ctx *c, int ell, const u64 *pt, double scale, double *values);

/* ---- client side: keygen / encrypt / decrypt (A.11; seeded, unpinned) ---- */
typedef struct ora_keys ora_keys;
ora_keys *ora_keygen(const ora_ctx *c, u64 seed);
void ora_keys_destroy(ora_keys *k);
const u64 *ora_keys_secret(const ora_keys *k);            /* [k][N] NTT */
const u64 *ora_keys_public(const ora_keys *k);            /* [2][k][N] NTT */
const u64 *ora_keys_relin(const ora_keys *k);             /* [k-1][2][k][N] */
/* creates (or returns cached) galois key for elt: [k-1][2][k][N] */
const u64 *ora_keys_galois(ora_keys *k, u64 elt);
/* public-key encryption of pt[ell][N] at data level with ell residues:
 * encrypt zero at key level, divide-and-round by P, drop to ell residues by
 * SEAL's modulus switching of the *plaintext level* -- here ell must be k-1
 * or lower; lower levels are produced by encrypting at that level directly
 * (SEAL Encryptor::encrypt at plain.parms_id()).  out ct[2][ell][N]. */
int ora_encrypt(const ora_keys *k, int ell, const u64 *pt, u64 seed, u64 *ct);
int ora_encrypt_with(const ora_keys *k, int ell, const u64 *pt, const int *u, const int *e0, const int *e1, u64 *ct);
/* pt[ell][N] = sum_i ct[i] * s^i */
int ora_decrypt(const ora_keys *k, int ell, const u64 *ct, int size, u64 *pt);

#ifdef __cplusplus
}
#endif
#endif
