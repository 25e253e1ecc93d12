"""CPU replay of the CUDA kernel bodies vs the oracle, bit-exact (no GPU).
Validates indexing / swizzles / strides / lazy-reduction bounds of the exact
code the sm_100a kernels execute, for every supported N."""
import pytest

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
