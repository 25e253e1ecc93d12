"""Pins the CPU oracle (oracle/ckks_oracle.c) to mathematical ground truth and
to SEAL known answers.  The reference holds no golden ciphertexts for this path
(SURVEY.md 8c), so these are the oracle's anchors:
  * SEAL's hard-coded default moduli (seal/util/globals.cpp, [SEAL-KNOWLEDGE])
    which the CoeffModulus::Create scan rule must reproduce,
  * direct polynomial evaluation / schoolbook negacyclic products,
  * big-integer CRT rounding,
  * decrypt consistency and the reference's MSE criterion (tests/common.py:34).
"""
import numpy as np
import pytest

from oracle import oracle as o


def bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2)


def test_prime_generation_matches_seal_default_tables():
    # SEAL 3.6 util/globals.cpp default_coeff_modulus_128 entries (descending
    # generation order there; Create() hands out smallest-first).
    assert o.gen_primes(4096, [36, 36]) == [0xffffc4001, 0xffffee001]
    assert o.gen_primes(4096, [37]) == [0x1ffffe0001]
    assert o.gen_primes(8192, [43, 43, 44, 44, 44]) == [
        0x7fffffc8001, 0x7fffffd8001, 0xfffffebc001, 0xffffff6c001, 0xfffffffc001]
    assert o.gen_primes(16384, [48, 48, 48, 49, 49, 49, 49, 49, 49]) == [
        0xfffffff00001, 0xfffffffa0001, 0xfffffffd8001, 0x1ffffffe48001, 0x1ffffffe88001,
        0x1ffffffea0001, 0x1ffffffee8001, 0x1fffffff50001, 0x1fffffff68001]


def test_prime_order_mixed_sizes():
    # [60,20,60,60] (reference tests/bug_fixes.py:68): each entry takes the
    # smallest unused prime of its size; last entry (key prime) is the largest.
    p = o.gen_primes(32768, [60, 20, 60, 60])
    assert p[0] < p[2] < p[3] and p[1] < 2 ** 20
    for q in p:
        assert q % 65536 == 1 and o.lib.ora_is_prime(q)
    assert p[3] == 0xffffffffffc0001


def test_minimal_primitive_root_bruteforce():
    for N, p in ((16, 97), (32, 193), (64, 257), (128, 7681), (256, 12289)):
        roots = [r for r in range(2, p) if pow(r, N, p) == p - 1]
        assert int(o.lib.ora_min_primitive_root(N, p)) == roots[0]


@pytest.mark.parametrize("N", [16, 64, 256])
def test_ntt_definition(N):
    bits = [30, 31, 60]
    orc = o.Oracle(N, bits)
    logN = N.bit_length() - 1
    rng = np.random.default_rng(N)
    for pi, p in enumerate(orc.primes):
        psi = orc.psi[pi]
        assert pow(psi, N, p) == p - 1
        assert psi == int(o.lib.ora_min_primitive_root(N, p))
        a = rng.integers(0, p, size=N, dtype=np.uint64)
        A = orc.ntt_fwd(a, pi)
        for i in range(N):
            x = pow(psi, 2 * bitrev(i, logN) + 1, p)
            acc = 0
            for c in reversed([int(v) for v in a]):
                acc = (acc * x + c) % p
            assert acc == int(A[i])
        assert np.array_equal(orc.ntt_inv(A, pi), a)


def test_ntt_negacyclic_convolution():
    N = 128
    orc = o.Oracle(N, [60, 60])
    p = orc.primes[0]
    rng = np.random.default_rng(3)
    a = rng.integers(0, p, size=N, dtype=np.uint64)
    b = rng.integers(0, p, size=N, dtype=np.uint64)
    ct = np.stack([orc.ntt_fwd(a, 0)])[None]  # [1,1,N]
    pt = np.stack([orc.ntt_fwd(b, 0)])
    prod = orc.ntt_inv(orc.mul_plain(ct, pt)[0, 0], 0)
    ref = [0] * N
    ai = [int(x) for x in a]; bi = [int(x) for x in b]
    for i in range(N):
        for j in range(N):
            k = i + j
            if k < N:
                ref[k] = (ref[k] + ai[i] * bi[j]) % p
            else:
                ref[k - N] = (ref[k - N] - ai[i] * bi[j]) % p
    assert [int(x) for x in prod] == ref


def test_ntt_large_roundtrip_and_60bit():
    for N in (4096, 16384):
        orc = o.Oracle(N, [60, 60, 60])
        for pi, p in enumerate(orc.primes):
            a = o.splitmix64_fill(0x5EA10000 + N + pi, N, p)
            assert np.array_equal(orc.ntt_inv(orc.ntt_fwd(a, pi), pi), a)


def test_dyadic_ops_vs_python_ints():
    N = 64
    orc = o.Oracle(N, [60, 59, 30, 60])
    ell = 3
    rng = np.random.default_rng(5)
    P = orc.primes[:ell]

    def rnd(s):
        return np.stack([np.stack([rng.integers(0, p, size=N, dtype=np.uint64) for p in P]) for _ in range(s)])
    a2, b2, a3, b3 = rnd(2), rnd(2), rnd(3), rnd(3)
    pt = rnd(1)[0]
    I = lambda x: [[[int(v) for v in r] for r in pl] for pl in x]
    A2, B2, A3, PT = I(a2), I(b2), I(a3), I(pt[None])[0]

    def chk(out, fn):
        out = I(out)
        for s in range(len(out)):
            for i, p in enumerate(P):
                for j in range(N):
                    assert out[s][i][j] == fn(s, i, j) % p, (s, i, j)
    chk(orc.add(a2, b2), lambda s, i, j: A2[s][i][j] + B2[s][i][j])
    chk(orc.sub(a2, b2), lambda s, i, j: A2[s][i][j] - B2[s][i][j])
    chk(orc.add(a2, a3), lambda s, i, j: (A2[s][i][j] if s < 2 else 0) + A3[s][i][j])
    chk(orc.sub(a2, a3), lambda s, i, j: (A2[s][i][j] if s < 2 else 0) - A3[s][i][j])
    chk(orc.sub(a3, a2), lambda s, i, j: A3[s][i][j] - (A2[s][i][j] if s < 2 else 0))
    chk(orc.negate(a3), lambda s, i, j: -A3[s][i][j])
    chk(orc.add_plain(a3, pt), lambda s, i, j: A3[s][i][j] + (PT[i][j] if s == 0 else 0))
    chk(orc.sub_plain(a2, pt), lambda s, i, j: A2[s][i][j] - (PT[i][j] if s == 0 else 0))
    chk(orc.mul_plain(a3, pt), lambda s, i, j: A3[s][i][j] * PT[i][j])
    chk(orc.mul(a2, b2), lambda s, i, j: [A2[0][i][j] * B2[0][i][j],
                                          A2[0][i][j] * B2[1][i][j] + A2[1][i][j] * B2[0][i][j],
                                          A2[1][i][j] * B2[1][i][j]][s])
    chk(orc.square(a2), lambda s, i, j: [A2[0][i][j] ** 2, 2 * A2[0][i][j] * A2[1][i][j], A2[1][i][j] ** 2][s])
    assert np.array_equal(orc.mod_switch(a3), a3[:, :2])
    z = np.zeros_like(a2)
    assert np.array_equal(orc.negate(z), z)


def crt(res, primes):
    Q = 1
    for p in primes:
        Q *= p
    x = 0
    for r, p in zip(res, primes):
        m = Q // p
        x += int(r) * m * pow(m, -1, p)
    return x % Q


def test_rescale_is_rounded_division_bigint():
    N = 32
    orc = o.Oracle(N, [40, 45, 50, 60])
    ell = 3
    P = orc.primes[:ell]
    rng = np.random.default_rng(11)
    coef = np.stack([rng.integers(0, p, size=N, dtype=np.uint64) for p in P])
    ct = np.stack([np.stack([orc.ntt_fwd(coef[i], i) for i in range(ell)])] * 2)
    out = orc.rescale(ct)
    assert out.shape == (2, 2, N)
    half = P[-1] >> 1
    for s in range(2):
        oc = [orc.ntt_inv(out[s, i], i) for i in range(ell - 1)]
        for j in range(N):
            x = crt([coef[i][j] for i in range(ell)], P)
            y = (x + half) // P[-1]
            for i in range(ell - 1):
                assert int(oc[i][j]) == y % P[i]


def _setup(N=1024, bits=(60, 60, 60, 60), seed=1):
    orc = o.Oracle(N, list(bits)).keygen(seed)
    return orc


def _enc(orc, vals, scale_bits, ell, seed=7):
    return orc.encrypt(orc.encode(vals, 2.0 ** scale_bits, ell), seed)


def _dec(orc, ct, scale_bits):
    return orc.decode(orc.decrypt(ct), 2.0 ** scale_bits)


def test_encode_decode_roundtrip():
    orc = o.Oracle(2048, [60, 60, 60])
    rng = np.random.default_rng(0)
    v = rng.uniform(-2, 2, 1024)
    for sb in (30, 45, 70):   # 70 exercises the > 2^64 coefficient path
        pt = orc.encode(v, 2.0 ** sb, 2)
        assert np.abs(orc.decode(pt, 2.0 ** sb) - v).max() < 1e-6
    # replication of short vectors (reference seal.cpp:71-79)
    pt = orc.encode(v[:8], 2.0 ** 30, 2)
    assert np.abs(orc.decode(pt, 2.0 ** 30) - np.tile(v[:8], 128)).max() < 1e-6


def test_encrypt_ops_decrypt_pipeline():
    orc = _setup()
    N, k = orc.N, orc.k
    rng = np.random.default_rng(1)
    x = rng.uniform(-2, 2, N // 2); y = rng.uniform(-2, 2, N // 2)
    cx, cy = _enc(orc, x, 30, 3, 7), _enc(orc, y, 30, 3, 8)
    cx45, cy45 = _enc(orc, x, 45, 3, 7), _enc(orc, y, 45, 3, 8)
    mse = lambda a, b: float(np.mean((a - b) ** 2))
    assert mse(_dec(orc, cx, 30), x) < 1e-8
    assert mse(_dec(orc, orc.add(cx, cy), 30), x + y) < 1e-8
    assert mse(_dec(orc, orc.sub(cx, cy), 30), x - y) < 1e-8
    assert mse(_dec(orc, orc.negate(cx), 30), -x) < 1e-8
    ptc = orc.encode(y, 2.0 ** 30, 3)
    assert mse(_dec(orc, orc.mul_plain(cx, ptc), 60), x * y) < 1e-8
    assert mse(_dec(orc, orc.add_plain(cx, ptc), 30), x + y) < 1e-8
    c3 = orc.mul(cx, cy)
    assert mse(_dec(orc, c3, 60), x * y) < 1e-8
    assert mse(_dec(orc, orc.square(cx), 60), x * x) < 1e-8
    rk = orc.relin_key()
    c2 = orc.relinearize(c3, rk)
    assert mse(_dec(orc, c2, 60), x * y) < 1e-8
    # relinearisation noise is tiny relative to q: decryptions agree to ~2^-30 of scale
    d = _dec(orc, c2, 60) - _dec(orc, c3, 60)
    assert np.abs(d).max() < 1e-6
    r = orc.rescale(orc.mul(cx45, cy45))     # size-3 rescale (lazy relinearisation)
    assert r.shape == (3, 2, N)
    q_last = orc.primes[2]
    assert mse(orc.decode(orc.decrypt(r), 2.0 ** 90 / q_last), x * y) < 1e-8
    r2 = orc.relinearize(r, rk)              # relinearise at a lower level
    assert mse(orc.decode(orc.decrypt(r2), 2.0 ** 90 / q_last), x * y) < 1e-8
    ms = orc.mod_switch(cx)
    assert mse(_dec(orc, ms, 30), x) < 1e-8
    # encryption directly at a lower level
    cl = _enc(orc, x, 30, 2, 9)
    assert mse(_dec(orc, cl, 30), x) < 1e-8


@pytest.mark.parametrize("steps", [1, 2, 5, -1, -3, 64])
def test_rotation(steps):
    orc = _setup(N=1024)
    N = orc.N
    x = np.random.default_rng(2).uniform(-2, 2, N // 2)
    cx = _enc(orc, x, 30, 3)
    elt = o.galois_elt_from_step(N, steps)
    gk = orc.galois_key(elt)
    out = orc.rotate(cx, steps, gk)
    assert np.mean((_dec(orc, out, 30) - np.roll(x, -steps)) ** 2) < 1e-8
    # also at a lower level (only key rows of live primes + P are used)
    cl = orc.mod_switch(cx)
    out = orc.rotate(cl, steps, gk)
    assert np.mean((_dec(orc, out, 30) - np.roll(x, -steps)) ** 2) < 1e-8


def test_galois_elements():
    N = 16384
    assert o.galois_elt_from_step(N, 0) == 2 * N - 1
    assert o.galois_elt_from_step(N, 1) == 3
    assert o.galois_elt_from_step(N, 2) == 9
    assert o.galois_elt_from_step(N, -1) == pow(3, N // 2 - 1, 2 * N)
    t = o.galois_table(16, 3)
    assert sorted(t.tolist()) == list(range(16))
