"""An INDEPENDENT second derivation of the two operations whose RNS formulas carry the most hidden choices --
key switching and rescaling (A.6) -- in exact big-integer arithmetic, compared word for word
with oracle/ckks_oracle.c on small rings (N = 16 ... 64).

The oracle works residue by residue (inverse NTT of each digit, unsigned representatives, +floor(P/2) before the
reduction, subtraction of floor(P/2) mod q_J afterwards, multiplication by P^-1 mod q_J).  Nothing of that appears
here.  Instead every polynomial is lifted to integer coefficients with the Chinese remainder theorem and

    rescale:      y = floor((X + floor(q_last/2)) / q_last)                       over the integers, then mod q_i
    key switch:   A_c = sum_J  D_J * K_c^(J)   (negacyclic product over Z, reduced mod Q_l * P, D_J = the digit
                  t mod q_J taken in [0, q_J)),   ks_c = floor((A_c + floor(P/2)) / P)   then mod q_J

with the NTT itself replaced by its definition, evaluation at psi^(2 bitrev(i) + 1) for the smallest primitive 2N-th
root psi found by exhaustive search.  Agreement pins the oracle's digit convention, rounding, root choice and output
ordering to one closed-form statement each.  (This narrows what could differ from SEAL; it does not replace SEAL:
parity with the library itself stays unpinned)
"""
import numpy as np
import pytest

from oracle import oracle as o


def bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


def is_prime(n):
    if n < 2:
        return False
    i = 2
    while i * i <= n and i < 1 << 16:      # trial division first, then Miller-Rabin with many bases
        if n % i == 0:
            return n == i
        i += 1
    d, r = n - 1, 0
    while d % 2 == 0:
        d //= 2; r += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41):
        if a % n == 0:
            continue
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(r - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def create_primes(N, bits):
    """A.1 restated: per bit size scan 2^b - 2N + 1 downwards; entries take the smallest unused prime of their size"""
    found = {}
    for b in set(bits):
        v, lst = (1 << b) - 2 * N + 1, []
        while len(lst) < bits.count(b):
            if is_prime(v):
                lst.append(v)
            v -= 2 * N
        found[b] = lst
    return [found[b].pop() for b in bits]


def smallest_root(N, q):
    """numerically smallest primitive 2N-th root of unity mod q, by exhaustive search over the group"""
    best = None
    g = 2
    while True:
        r = pow(g, (q - 1) // (2 * N), q)
        if pow(r, N, q) == q - 1:
            break
        g += 1
    x = r
    for _ in range(N):            # all primitive roots are the odd powers of one of them
        best = x if best is None or x < best else best
        x = x * r * r % q
    return best


class Ring:
    def __init__(self, N, primes):
        self.N, self.n, self.primes = N, N.bit_length() - 1, primes
        self.psi = [smallest_root(N, q) for q in primes]

    def ntt(self, coeffs, i):      # definition: evaluation at psi^(2 bitrev(j) + 1)
        q, psi = self.primes[i], self.psi[i]
        return [sum(c * pow(psi, (2 * bitrev(j, self.n) + 1) * e, q) for e, c in enumerate(coeffs)) % q for j in range(self.N)]

    def intt(self, vals, i):       # interpolation: c_e = N^-1 sum_j v_j x_j^-e
        q, psi = self.primes[i], self.psi[i]
        ninv = pow(self.N, q - 2, q)
        xs = [pow(psi, 2 * bitrev(j, self.n) + 1, q) for j in range(self.N)]
        return [ninv * sum(v * pow(x, q - 1 - e, q) for v, x in zip(vals, xs)) % q for e in range(self.N)]

    def lift(self, residues, idx):
        """CRT: coefficient vectors mod primes[idx[r]] -> integer coefficients in [0, prod)"""
        mods = [self.primes[i] for i in idx]
        M = 1
        for m in mods:
            M *= m
        out = [0] * self.N
        for r, m in enumerate(mods):
            Mi = M // m
            w = Mi * pow(Mi, -1, m)
            for e in range(self.N):
                out[e] = (out[e] + residues[r][e] * w) % M
        return out, M


def negacyclic(a, b, M):
    N = len(a)
    out = [0] * N
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                k = i + j
                if k < N:
                    out[k] = (out[k] + x * y) % M
                else:
                    out[k - N] = (out[k - N] - x * y) % M
    return out


@pytest.mark.parametrize("N,bits", [(16, [30, 30, 40]), (32, [60, 60, 60, 60]), (64, [60, 50, 60]), (16, [60, 20, 60, 60])])
def test_primes_roots_and_ntt_agree_with_the_definition(N, bits):
    orc = o.Oracle(N, bits)
    assert orc.primes == create_primes(N, bits)
    ring = Ring(N, orc.primes)
    assert orc.psi == ring.psi
    rng = np.random.default_rng(N)
    for i, q in enumerate(orc.primes):
        a = [int(v) for v in rng.integers(0, q, N, dtype=np.uint64)]
        assert [int(v) for v in orc.ntt_fwd(np.array(a, dtype=np.uint64), i)] == ring.ntt(a, i)
        assert [int(v) for v in orc.ntt_inv(np.array(a, dtype=np.uint64), i)] == ring.intt(a, i)


@pytest.mark.parametrize("N,bits,size", [(16, [30, 30, 40, 40], 2), (32, [60, 60, 60, 60], 3), (64, [60, 50, 60], 2)])
def test_rescale_is_rounded_integer_division(N, bits, size):
    orc = o.Oracle(N, bits)
    ring = Ring(N, orc.primes)
    rng = np.random.default_rng(7 * N)
    for ell in range(2, orc.k):             # data levels with at least two residues
        ct = np.stack([np.stack([rng.integers(0, orc.primes[i], N, dtype=np.uint64) for i in range(ell)]) for _ in range(size)])
        got = orc.rescale(ct)
        ql = orc.primes[ell - 1]
        for s in range(size):
            coeff = [ring.intt([int(v) for v in ct[s, i]], i) for i in range(ell)]
            X, _ = ring.lift(coeff, list(range(ell)))
            Y = [(x + ql // 2) // ql for x in X]
            for i in range(ell - 1):
                want = ring.ntt([y % orc.primes[i] for y in Y], i)
                assert [int(v) for v in got[s, i]] == want, (ell, s, i)


@pytest.mark.parametrize("N,bits", [(16, [30, 30, 40, 40]), (32, [60, 60, 60, 60]), (16, [60, 20, 60, 60]), (64, [60, 50, 60])])
def test_key_switch_is_digit_product_and_rounded_division_by_P(N, bits):
    orc = o.Oracle(N, bits).keygen(3)
    ring = Ring(N, orc.primes)
    k, P = orc.k, orc.primes[-1]
    key = orc.relin_key()                    # [k-1][2][k][N], NTT form at key level
    rng = np.random.default_rng(11 * N)
    for ell in range(1, k):
        t = np.stack([rng.integers(0, orc.primes[i], N, dtype=np.uint64) for i in range(ell)])
        got = orc.keyswitch(t, key)           # [2][ell][N]
        live = list(range(ell)) + [k - 1]     # output moduli: q_0 .. q_{ell-1}, P
        digits = [ring.intt([int(v) for v in t[J]], J) for J in range(ell)]     # integers in [0, q_J): the unsigned digit
        for c in range(2):
            A, M = [0] * N, None
            for J in range(ell):
                Kc = [ring.intt([int(v) for v in key[J, c, i]], i) for i in live]
                Klift, M = ring.lift(Kc, live)
                prod = negacyclic(digits[J], Klift, M)
                A = [(a + b) % M for a, b in zip(A, prod)]
            ks = [(a + P // 2) // P for a in A]
            for J in range(ell):
                want = ring.ntt([v % orc.primes[J] for v in ks], J)
                assert [int(v) for v in got[c, J]] == want, (ell, c, J)
