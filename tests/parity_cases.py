"""Bit-exact parity cases: implementation under test (CPU kernel replay or the
CUDA library through the C-ABI) versus the CPU oracle on identical seeded
inputs.  Integer pipeline => the bar is equality of every u64 word."""
import numpy as np

from oracle import oracle as o

_ORACLES = {}


def get_oracle(N, bits, seed=1):
    key = (N, tuple(bits))
    if key not in _ORACLES:
        _ORACLES[key] = o.Oracle(N, list(bits)).keygen(seed)
    return _ORACLES[key]


def rand_ct(orc, size, ell, seed):
    """uniform residues in [0,q_i): valid evaluator input (ops are data-independent)"""
    return np.stack([np.stack([o.splitmix64_fill(seed * 1000 + s * 37 + i, orc.N, orc.primes[i]) for i in range(ell)])
                     for s in range(size)])


def eq(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError("mismatch at %d/%d words, first %s: got %d want %d" % (
            len(bad), a.size, bad[0], a[tuple(bad[0])], b[tuple(bad[0])]))


def case_ntt(be, orc, L=None):
    """residues seeded SplitMix64(0x5EA10000 + N + L)."""
    N = orc.N
    L = L or orc.k
    data = np.stack([np.stack([o.splitmix64_fill(0x5EA10000 + N + L + 131 * b + r, N, orc.primes[r]) for r in range(L)])
                     for b in range(2)])
    want = np.stack([np.stack([orc.ntt_fwd(data[b, r], r) for r in range(L)]) for b in range(2)])
    got = be.ntt(data, list(range(L)))
    eq(got, want)
    back = be.ntt(got, list(range(L)), inverse=True)
    eq(back, data)
    # edge values: 0, 1, p-1
    edge = np.zeros((L, N), dtype=np.uint64)
    for r in range(L):
        edge[r, 0::3] = orc.primes[r] - 1
        edge[r, 1::3] = 1
    eq(be.ntt(edge, list(range(L))), np.stack([orc.ntt_fwd(edge[r], r) for r in range(L)]))
    eq(be.ntt(edge, list(range(L)), inverse=True), np.stack([orc.ntt_inv(edge[r], r) for r in range(L)]))


def case_dyadic(be, orc, ell):
    a2, b2 = rand_ct(orc, 2, ell, 1), rand_ct(orc, 2, ell, 2)
    a3, b3 = rand_ct(orc, 3, ell, 3), rand_ct(orc, 3, ell, 4)
    pt = rand_ct(orc, 1, ell, 5)[0]
    eq(be.add(a2, b2), orc.add(a2, b2))
    eq(be.add(a3, b2), orc.add(a3, b2))
    eq(be.add(a2, b3), orc.add(a2, b3))
    eq(be.add(a3, b3), orc.add(a3, b3))
    eq(be.sub(a2, b2), orc.sub(a2, b2))
    eq(be.sub(a3, b2), orc.sub(a3, b2))
    eq(be.sub(a2, b3), orc.sub(a2, b3))
    eq(be.negate(a3), orc.negate(a3))
    eq(be.add_plain(a2, pt), orc.add_plain(a2, pt))
    eq(be.add_plain(a3, pt), orc.add_plain(a3, pt))
    eq(be.sub_plain(a3, pt), orc.sub_plain(a3, pt))
    eq(be.mul_plain(a2, pt), orc.mul_plain(a2, pt))
    eq(be.mul_plain(a3, pt), orc.mul_plain(a3, pt))
    eq(be.mul(a2, b2), orc.mul(a2, b2))
    eq(be.square(a2), orc.square(a2))
    z = np.zeros_like(a2)
    eq(be.negate(z), z)
    eq(be.sub(z, z), z)
    if ell >= 2:
        eq(be.mod_switch(a3), orc.mod_switch(a3))


def case_sum_terms(be, orc, ell, nterms=(2, 9, 32)):
    """fused multiply_plain / add chains == the same calls one by one on the oracle"""
    for n in nterms:
        cts = [rand_ct(orc, 3 if (t % 5 == 4) else 2, ell, 300 + t) for t in range(n)]
        pts = [None if (t % 3 == 2) else rand_ct(orc, 1, ell, 400 + t)[0] for t in range(n)]
        want = None
        for c, p in zip(cts, pts):
            term = c if p is None else orc.mul_plain(c, p)
            want = term if want is None else orc.add(want, term)
        eq(be.sum_terms(cts, pts), want)
    # worst-case magnitudes: every residue p-1
    c = np.stack([np.stack([np.full(orc.N, orc.primes[i] - 1, dtype=np.uint64) for i in range(ell)])] * 2)
    cts, pts = [c] * 32, [c[0]] * 32
    want = None
    for t in range(32):
        term = orc.mul_plain(cts[t], pts[t])
        want = term if want is None else orc.add(want, term)
    eq(be.sum_terms(cts, pts), want)


def case_encode_uniform(be, orc):
    """scalar constants through the one-pass encoder == the full FFT encoder of the oracle on the replicated vector"""
    vals = [0.0, -0.0, 1.0, -1.0, 2.0, -2.0, 0.5, 0.17254603006834726, -1.0984324107372518, 2.213787482387662, 1e-9, -3e-7, 123456.789, 1e12, -7e15]
    for scale_bits, ell in ((25, min(2, orc.k)), (40, orc.k - 1), (60, 1), (90, orc.k), (120, orc.k)):
        got = be.encode_uniform(vals, 2.0 ** scale_bits, ell)
        for e, v in enumerate(vals):
            eq(got[e], orc.encode(np.array([v]), 2.0 ** scale_bits, ell))
            if e % 5 == 0:   # replicated explicitly over more than one element
                eq(got[e], orc.encode(np.full(8, v), 2.0 ** scale_bits, ell))


def case_sum_products(be, orc, ell, nterms=(2, 7, 32)):
    """sums with ciphertext products (multiply / square) among the terms == the oracle calls one by one"""
    for n in nterms:
        cts, seconds, kinds, want = [], [], [], None
        for t in range(n):
            kind = (2, 2, 1, 0, 2)[t % 5]
            a = rand_ct(orc, 3 if kind == 0 and t % 2 else 2, ell, 500 + t)
            if kind == 2:
                b = a if t % 7 == 3 else rand_ct(orc, 2, ell, 600 + t)       # a (x) a: square
                term = orc.square(a) if b is a else orc.mul(a, b)
            elif kind == 1:
                b = rand_ct(orc, 1, ell, 700 + t)[0]
                term = orc.mul_plain(a, b)
            else:
                b, term = None, a
            cts.append(a); seconds.append(b); kinds.append(kind)
            want = term if want is None else orc.add(want, term)
        eq(be.sum_products(cts, seconds, kinds), want)
    # worst case for the 128-bit accumulators: 32 products of residues p-1 (two per middle coefficient)
    c = np.stack([np.stack([np.full(orc.N, orc.primes[i] - 1, dtype=np.uint64) for i in range(ell)])] * 2)
    want = None
    for t in range(32):
        term = orc.square(c)
        want = term if want is None else orc.add(want, term)
    eq(be.sum_products([c] * 32, [c] * 32, [2] * 32), want)


def case_rescale(be, orc, ell):
    for size in (2, 3):
        a = rand_ct(orc, size, ell, 10 + size)
        eq(be.rescale(a), orc.rescale(a))


def case_keyswitch(be, orc, ell, steps=(1, -2)):
    a3 = rand_ct(orc, 3, ell, 21)
    rk = orc.relin_key()
    eq(be.relinearize(a3, rk), orc.relinearize(a3, rk))
    a2 = rand_ct(orc, 2, ell, 22)
    gks = []
    for s in steps:
        gk = orc.galois_key(o.galois_elt_from_step(orc.N, s))
        gks.append(gk)
        eq(be.rotate(a2, s, gk), orc.rotate(a2, s, gk))
    # rotations sharing the inverse NTT of the input (hoisted, exact)
    if hasattr(be, "rotate_many"):
        for s, gk, got in zip(steps, gks, be.rotate_many(a2, list(steps), gks)):
            eq(got, orc.rotate(a2, s, gk))
    # rotations sharing the inverse NTT and the mod-up of the input (exact: ops_impl.hpp hoisted_modup)
    if hasattr(be, "rotate_many_modup"):
        outs, flag = be.rotate_many_modup(a2, list(steps), gks)
        assert flag == 0
        for s, gk, got, got_many in zip(steps, gks, outs, be.many):
            want = orc.rotate(a2, s, gk)
            eq(got, want)
            eq(got_many, want)      # evab_rotate_modup_many: all of them in three launches
        # a digit with a zero coefficient: negate(0) = 0 carries no q_J, the shared mod-up must say so (the caller falls back)
        coeff = orc.ntt_inv(a2[1].copy(), list(range(ell))) if False else None
        z = a2.copy()
        c1 = np.stack([orc.ntt_inv(z[1, i], i) for i in range(ell)])
        c1[0, 5] = 0
        z[1] = np.stack([orc.ntt_fwd(c1[i], i) for i in range(ell)])
        _, flag = be.rotate_many_modup(z, list(steps[:1]), gks[:1])
        assert flag == 1


def case_decode(be, orc):
    """CKKSEncoder::decode on the device: identical doubles to the oracle's decoder at every level, for plaintexts from the
    encoder and for raw residues (values far beyond the scale: the multi-word path of the CRT composition)"""
    rng = np.random.default_rng(orc.N + 17)
    for ell in range(1, min(orc.k, 8) + 1):
        for scale_bits in (30, 50):
            pt = orc.encode(rng.uniform(-4, 4, orc.N // 2), 2.0 ** scale_bits, ell)
            got = be.decode(pt, 2.0 ** scale_bits, orc.primes)
            want = orc.decode(pt, 2.0 ** scale_bits)
            assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (ell, scale_bits)
        raw = np.stack([rng.integers(0, orc.primes[i], orc.N, dtype=np.uint64) for i in range(ell)])
        got, want = be.decode(raw, 2.0 ** 40, orc.primes), orc.decode(raw, 2.0 ** 40)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), ("raw", ell)


def case_lazy_rotsum(be, orc, ell, steps=(1, -1, 7, 64)):
    """evab_lazy_rotsum (opt-in approx_hoist): out_o = sum_i pt_oi (.) rotate(ct, s_i), one rounding by P per sum, the key-switch
    inner product of a rotation shared by the sums.  Not the reference's rounding (one per rotation) -- so no bit-exact oracle:
    the plaintexts it consumes ARE bit-exact (rows mod q equal the oracle encoder's, the extra row is the same integer polynomial
    mod P), and every sum must decrypt to the same slots as the reference sequence rotate -> multiply_plain -> add within CKKS noise."""
    N = orc.N
    xs, wsc = (2.0 ** 40, 2.0 ** 30) if ell >= 2 else (2.0 ** 25, 2.0 ** 20)   # the product has to fit the level's modulus
    rng = np.random.default_rng(N + ell)
    x = rng.uniform(-1, 1, N // 2)
    ct = orc.encrypt(orc.encode(x, xs, ell), seed=11)
    n = len(steps)
    sets = [[rng.uniform(-1, 1, N // 2) for _ in steps],                                   # every rotation
            [rng.uniform(-1, 1, N // 2) if i != 1 else None for i in range(n)],            # one rotation missing
            [np.full(N // 2, 0.5 - i) if i % 2 == 0 else None for i in range(n)]]          # scalar weights, every other rotation
    gks = [orc.galois_key(o.galois_elt_from_step(N, s)) for s in steps]
    outs, pts, flag = be.lazy_rotsum(ct, list(steps), gks, sets, wsc)
    assert flag == 0
    rots = [orc.rotate(ct, s, gk) for s, gk in zip(steps, gks)]
    errs = []
    for oi, ws in enumerate(sets):
        want, true_v = None, 0
        for i, w in enumerate(ws):
            if w is None:
                continue
            eq(pts[(oi, i)][:ell], orc.encode(w, wsc, ell))
            term = orc.mul_plain(rots[i], pts[(oi, i)][:ell])
            want = term if want is None else orc.add(want, term)
            true_v = true_v + w * np.roll(x, -steps[i])
        got_v, want_v = orc.decode(orc.decrypt(outs[oi]), xs * wsc), orc.decode(orc.decrypt(want), xs * wsc)
        e_want, e_got = np.abs(want_v - true_v).max(), np.abs(got_v - true_v).max()
        bound = 64 / wsc + N * 1024 / xs     # encoding error of the weights + key-switching noise at this scale
        assert e_want < bound and e_got < bound and np.abs(got_v - want_v).max() < bound, (oi, e_want, e_got, bound)
        errs.append((e_want, e_got))
    return errs


def check_special_row(orc, pt_ext, ell):
    """rows [0, ell) and the extra row of an encode_ext(with_p=1) plaintext hold the same small integer polynomial"""
    q0, P = int(orc.primes[0]), int(orc.primes[orc.k - 1])
    c = orc.ntt_inv(pt_ext[0].copy(), 0).astype(object)
    c = np.where(c > q0 // 2, c - q0, c)
    assert max(abs(int(v)) for v in c) < 2 ** 58
    cp = np.array([int(v) % P for v in c], dtype=np.uint64)
    eq(pt_ext[ell], orc.ntt_fwd(cp, orc.k - 1))
