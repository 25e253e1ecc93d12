"""GPU parity proper: the CUDA library through the C-ABI (include/evab200.h)
versus the CPU oracle, bit-exact on identical seeded inputs."""
import numpy as np
import pytest

import parity_cases as pc
from oracle import oracle as o

pytestmark = pytest.mark.gpu


def _be(N, primes):
    from backends import GpuBackend
    return GpuBackend(N, primes)


@pytest.mark.parametrize("N,bits", [
    (1024, [30, 30, 40]), (2048, [54, 55, 60]), (4096, [60, 20, 60, 60]), (8192, [60, 60, 60]),
    (16384, [60, 60, 60, 60, 60]), (32768, [60, 20, 60, 60]),
])
def test_gpu_ntt(N, bits):
    orc = pc.get_oracle(N, bits)
    pc.case_ntt(_be(N, orc.primes), orc)


@pytest.mark.parametrize("N,bits", [(1024, [40, 50, 60, 60]), (4096, [60, 20, 60, 60]), (8192, [60, 60, 60])])
def test_gpu_ops_small(N, bits):
    orc = pc.get_oracle(N, bits)
    be = _be(N, orc.primes)
    for ell in range(1, orc.k):
        pc.case_dyadic(be, orc, ell)
        pc.case_sum_terms(be, orc, ell)
        pc.case_sum_products(be, orc, ell)
        pc.case_keyswitch(be, orc, ell)
        if ell >= 2:
            pc.case_rescale(be, orc, ell)


def test_gpu_ops_sobel_shape():
    """N=16384, 5x60-bit primes: the Sobel/Harris parameter set."""
    orc = pc.get_oracle(16384, [60] * 5)
    be = _be(16384, orc.primes)
    for ell in (4, 3, 2, 1):
        pc.case_dyadic(be, orc, ell)
        pc.case_sum_terms(be, orc, ell, nterms=(9,))
        pc.case_keyswitch(be, orc, ell, steps=(1, 64, 130))
        if ell >= 2:
            pc.case_rescale(be, orc, ell)


def test_gpu_ops_n32768():
    orc = pc.get_oracle(32768, [60, 20, 60, 60, 60, 60])
    be = _be(32768, orc.primes)
    pc.case_dyadic(be, orc, 5)
    pc.case_rescale(be, orc, 5)
    pc.case_keyswitch(be, orc, 5, steps=(-1,))
    pc.case_keyswitch(be, orc, 2, steps=(3,))


def test_gpu_ntt_large_batch_roundtrip():
    """Full-size property check (oracle only on the first polynomials): batch
    larger than L2, inverse(forward(x)) == x for every word."""
    N, L, B = 16384, 4, 256
    orc = pc.get_oracle(16384, [60] * 5)
    be = _be(N, orc.primes)
    data = np.stack([o.splitmix64_fill(0x5EA10000 + N + L + r, B * N, orc.primes[r]).reshape(B, N) for r in range(L)], axis=1)
    data = np.ascontiguousarray(data)            # [B][L][N]
    fwd = be.ntt(data, list(range(L)))
    for b in range(2):
        for r in range(L):
            pc.eq(fwd[b, r], orc.ntt_fwd(data[b, r], r))
    pc.eq(be.ntt(fwd, list(range(L)), inverse=True), data)


def test_gpu_error_paths():
    from eva_b200 import cabi
    import ctypes as C
    lib = cabi.load()
    h = C.c_void_p()
    bad = np.array([97, 193], dtype=np.uint64)
    assert lib.evab_ctx_create(C.c_uint64(4096), bad.ctypes.data_as(cabi.u64p), 2, 0, C.byref(h)) != 0
    assert b"2N" in lib.evab_last_error()
    assert lib.evab_ctx_create(C.c_uint64(3000), bad.ctypes.data_as(cabi.u64p), 2, 0, C.byref(h)) != 0


@pytest.mark.parametrize("N,bits", [(4096, [60, 40, 60]), (16384, [60] * 5), (32768, [60, 20, 60, 60])])
def test_gpu_encoder_bit_exact(N, bits):
    pc.case_encode_uniform(_be(N, pc.get_oracle(N, bits).primes), pc.get_oracle(N, bits))
    """SURVEY 8a row E: device encoder (FP64 FFT + rounding + NTT) == host encoder == oracle, bit for bit;
    decode(encode(x)) ~ x."""
    from eva_b200 import b200
    orc = pc.get_oracle(N, bits)
    k = len(orc.primes)
    z = np.zeros((k - 1, 2, k, N), dtype=np.uint64)
    pub = b200.context_from_raw_keys(N, orc.primes, z, {})
    rng = np.random.default_rng(N)
    for vals, sb in ((rng.uniform(-2, 2, N // 2), 30), (np.array([0.17254603006834726]), 60), (rng.uniform(-1, 1, 64), 90),
                     (np.zeros(8), 40)):
        for ell in (k - 1, 1):
            want = orc.encode(vals, 2.0 ** sb, ell)
            dev = pub.encode(list(vals), 2.0 ** sb, ell)
            host = pub.encode(list(vals), 2.0 ** sb, ell, True)
            pc.eq(dev, want)
            pc.eq(host, want)
    x = rng.uniform(-2, 2, N // 2)
    dec = np.array(pub.decode(pub.encode(list(x), 2.0 ** 40, k - 1), 2.0 ** 40))
    assert np.abs(dec - x).max() < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("N,bits,cl", [(4096, [60, 20, 60, 60], 1), (8192, [60, 60, 60], 1), (8192, [60, 60, 60], 2), (16384, [60] * 5, 1), (16384, [60] * 5, 4)])
def test_gpu_cluster_distributed_ntt(N, bits, cl):
    """evab_set_ntt_cluster: one residue over a cluster of 2 / 4 CTAs (distributed shared memory) --
    identical results for the transforms and every fused variant (key switch, rescale, encoder)."""
    import numpy as np
    from eva_b200 import cabi
    lib = cabi.load()
    orc = pc.get_oracle(N, bits)
    be = _be(N, orc.primes)
    assert lib.evab_set_ntt_cluster(cl) == 0
    try:
        pc.case_ntt(be, orc)
        for ell in range(len(bits) - 1, 0, -1):
            pc.case_keyswitch(be, orc, ell, steps=(1, -3))
            if ell >= 2:
                pc.case_rescale(be, orc, ell)
    finally:
        assert lib.evab_set_ntt_cluster(0) == 0   # library default (automatic)
    assert lib.evab_set_ntt_cluster(3) != 0


@pytest.mark.parametrize("N,bits", [(4096, [60, 20, 60, 60]), (16384, [60] * 5)])
def test_encrypt_decrypt_decode_bit_exact_vs_oracle(N, bits):
    """the client side on the device.  Encryptor::encrypt with the oracle's public key and the SAME
    randomness (u, e0, e1) gives the oracle's ciphertext bit for bit at every level; Decryptor::decrypt with the
    oracle's secret key gives the oracle's plaintext polynomial (sizes 2 and 3); decoding it gives the oracle's doubles."""
    from eva_b200 import b200
    orc = pc.get_oracle(N, bits).keygen(5)
    pub = b200.public_from_raw(N, orc.primes, orc.public_key(), None, {})
    sec = b200.secret_from_raw(N, orc.primes, orc.secret_key())
    rng = np.random.default_rng(N)
    for ell in range(1, orc.k):
        vals = rng.uniform(-1, 1, N // 2)
        pt = orc.encode(vals, 2.0 ** 30, ell)
        u = rng.integers(-1, 2, N).astype(np.int32)
        e0, e1 = rng.integers(-20, 21, N).astype(np.int32), rng.integers(-20, 21, N).astype(np.int32)
        want = orc.encrypt_with(pt, u, e0, e1)
        got = pub.encrypt_poly(pt, [int(x) for x in u], [int(x) for x in e0], [int(x) for x in e1])
        assert np.array_equal(got, want), "encrypt differs at ell=%d" % ell
        dp = sec.decrypt_poly(want)
        assert np.array_equal(dp, orc.decrypt(want)), "decrypt differs at ell=%d" % ell
        ct3 = orc.mul(want, want)        # a size-3 ciphertext: c0 + c1 s + c2 s^2
        assert np.array_equal(sec.decrypt_poly(ct3), orc.decrypt(ct3))
        dec = np.asarray(sec.decode(dp, 2.0 ** 30))
        ref = orc.decode(orc.decrypt(want), 2.0 ** 30)
        assert np.array_equal(dec[: N // 2], ref), "decode differs at ell=%d" % ell
        assert np.max(np.abs(dec[: N // 2] - vals)) < 1e-4


@pytest.mark.parametrize("N,bits", [(4096, [60, 20, 60, 60]), (16384, [60] * 5)])
def test_gpu_decode_bit_exact(N, bits):
    """evab_decode: identical doubles to the oracle's decoder"""
    orc = pc.get_oracle(N, bits)
    pc.case_decode(_be(N, orc.primes), orc)


@pytest.mark.parametrize("N,bits", [(1024, [60, 60, 60, 60]), (16384, [60] * 5), (4096, [60, 40, 40, 60])])
def test_gpu_lazy_rotsum(N, bits):
    """opt-in approx_hoist kernels (evab_lazy_rotsum, evab_encode_ext): the plaintexts are bit-exact (rows mod q) and carry
    the same integers mod P; the sum decrypts to the reference sequence's slots within CKKS noise (it is NOT bit-exact)"""
    orc = pc.get_oracle(N, bits)
    be = _be(N, orc.primes)
    for ell in range(orc.k - 1, 0, -1):
        pc.case_lazy_rotsum(be, orc, ell)
    rng = np.random.default_rng(N)
    ell = orc.k - 1
    ct = orc.encrypt(orc.encode(rng.uniform(-1, 1, N // 2), 2.0 ** 40, ell))
    steps = list(range(1, 17))
    gks = [orc.galois_key(o.galois_elt_from_step(N, s)) for s in steps]
    _, pts, flag = be.lazy_rotsum(ct, steps, gks, [[rng.uniform(-8, 8, N // 2) for _ in steps]] * 4, 2.0 ** 45)   # 4 sums of 16: the most one call takes
    assert flag == 0
    for i in range(3):
        pc.check_special_row(orc, pts[(0, i)], ell)
