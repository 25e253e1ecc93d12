"""CPU replay of the CUDA kernel bodies vs the oracle, bit-exact (no GPU).
Validates indexing / swizzles / strides / lazy-reduction bounds of the exact
code the sm_100a kernels execute, for every supported N."""
import pytest

import numpy as np

from oracle import oracle as o
import parity_cases as pc
from backends import EmuBackend


@pytest.mark.parametrize("N,bits", [
    (1024, [30, 30, 40]), (2048, [54, 55, 60]), (4096, [60, 20, 60, 60]), (8192, [60, 60, 60]),
    (16384, [60, 60, 60, 60, 60]), (32768, [60, 20, 60, 60]),
])
def test_emu_ntt(N, bits):
    orc = pc.get_oracle(N, bits)
    pc.case_ntt(EmuBackend(N, orc.primes), orc)


@pytest.mark.parametrize("N,bits", [(1024, [40, 50, 60, 60]), (4096, [60, 20, 60, 60])])
def test_emu_ops_small(N, bits):
    orc = pc.get_oracle(N, bits)
    be = EmuBackend(N, orc.primes)
    for ell in (1, 2, 3):
        pc.case_dyadic(be, orc, ell)
        pc.case_sum_terms(be, orc, ell)
        pc.case_sum_products(be, orc, ell)
        pc.case_keyswitch(be, orc, ell)
        if ell >= 2:
            pc.case_rescale(be, orc, ell)


def test_emu_ops_sobel_shape():
    orc = pc.get_oracle(16384, [60] * 5)
    be = EmuBackend(16384, orc.primes)
    pc.case_rescale(be, orc, 4)
    pc.case_keyswitch(be, orc, 4, steps=(64,))
    pc.case_keyswitch(be, orc, 2, steps=(1,))


def test_emu_ops_n32768():
    orc = pc.get_oracle(32768, [60, 20, 60, 60])
    be = EmuBackend(32768, orc.primes)
    pc.case_rescale(be, orc, 3)
    pc.case_keyswitch(be, orc, 3, steps=(-1,))


@pytest.mark.parametrize("N,bits", [(1024, [40, 50, 60, 60]), (4096, [60, 20, 60, 60]), (32768, [60, 20, 60, 60])])
def test_emu_encode(N, bits):
    """Device encoder bodies (FP64 FFT, rounding, residues, NTT) vs the oracle, including scalar
    constants, which take the constant-polynomial path of the forward NTT."""
    import numpy as np
    orc = pc.get_oracle(N, bits)
    be = EmuBackend(N, orc.primes)
    rng = np.random.default_rng(N)
    vecs = [np.array([1.0]), np.array([-2.0]), np.array([0.0]), rng.uniform(-3, 3, 8), np.array([0.17254603006834726]),
            rng.uniform(-1, 1, N // 2), np.array([2.5, 2.5]), np.array([1.0, -1.0])]
    pc.case_encode_uniform(be, orc)
    for scale_bits, ell in ((25, 2), (40, len(bits) - 1), (60, 1), (90, len(bits))):
        got = be.encode(vecs, 2.0 ** scale_bits, ell)
        for e, v in enumerate(vecs):
            want = orc.encode(v, 2.0 ** scale_bits, ell)
            assert np.array_equal(got[e], want), (scale_bits, ell, e)


@pytest.mark.parametrize("N,bits,cl", [(4096, [60, 20, 60, 60], 2), (4096, [60, 20, 60, 60], 4), (8192, [60, 60, 60], 4),
                                        (16384, [60, 60, 60, 60, 60], 2), (16384, [60, 60, 60, 60, 60], 4), (8192, [60, 60, 60], 8), (16384, [60, 60, 60, 60, 60], 8),
                                        (32768, [60, 20, 60, 60], 4), (32768, [60, 20, 60, 60], 8)])
def test_emu_cluster_distributed(N, bits, cl):
    """one residue spread over a cluster of 2 / 4 CTAs (distributed shared-memory exchange)"""
    import numpy as np
    orc = pc.get_oracle(N, bits)
    be = EmuBackend(N, orc.primes, cluster=cl)
    pc.case_ntt(be, orc)
    ell = len(bits) - 1
    pc.case_rescale(be, orc, ell)
    pc.case_keyswitch(be, orc, ell, steps=(3,))
    v = np.array([0.5, -1.25, 3.0, 2.0])
    assert np.array_equal(be.encode(v, 2.0 ** 30, ell), orc.encode(v, 2.0 ** 30, ell))


@pytest.mark.parametrize("N,bits", [(1024, [40, 50, 60, 60]), (4096, [60, 20, 60, 60])])
def test_emu_decode(N, bits):
    """device decoder bodies (CRT composition, FP64 forward FFT) vs the oracle: identical doubles"""
    orc = pc.get_oracle(N, bits)
    pc.case_decode(EmuBackend(N, orc.primes), orc)


def test_lazy_rotsum_emulated():
    """opt-in approx_hoist kernels on the CPU emulator: plaintexts with the extra residue row, one mod-down per rotation sum"""
    N = 1024
    orc = pc.get_oracle(N, [60, 60, 60, 60])
    be = EmuBackend(N, orc.primes)
    for ell in (3, 1):
        pc.case_lazy_rotsum(be, orc, ell)
    orc2 = pc.get_oracle(N, [60, 40, 40, 60])
    be2 = EmuBackend(N, orc2.primes)
    pc.case_lazy_rotsum(be2, orc2, 2, steps=(2, 5))
    rng = np.random.default_rng(3)
    _, pts, _ = be2.lazy_rotsum(orc2.encrypt(orc2.encode(rng.uniform(-1, 1, N // 2), 2.0 ** 40, 2)), [1, 2],
                                [orc2.galois_key(o.galois_elt_from_step(N, s)) for s in (1, 2)], [[rng.uniform(-8, 8, N // 2), np.full(N // 2, 0.37)]], 2.0 ** 45)
    for pt in pts.values():
        pc.check_special_row(orc2, pt, 2)
