"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg may import this module (see oracle/ckks_oracle.h).
The product package ``eva_b200`` never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libckks_oracle.so")
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


def build(force=False):
    if force or not os.path.exists(_LIB):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


def _load():
    build()
    lib = C.CDLL(_LIB)
    lib.ora_ctx_create.restype = C.c_void_p
    lib.ora_ctx_create.argtypes = [C.c_uint64, C.POINTER(C.c_int), C.c_int]
    lib.ora_ctx_create_from_primes.restype = C.c_void_p
    lib.ora_ctx_create_from_primes.argtypes = [C.c_uint64, u64p, C.c_int]
    lib.ora_ctx_destroy.argtypes = [C.c_void_p]
    for n in ("ora_ctx_N", "ora_ctx_prime", "ora_ctx_psi", "ora_mulmod", "ora_powmod", "ora_invmod",
              "ora_min_primitive_root", "ora_galois_elt_from_step"):
        getattr(lib, n).restype = C.c_uint64
    lib.ora_ctx_N.argtypes = [C.c_void_p]
    lib.ora_ctx_k.argtypes = [C.c_void_p]
    lib.ora_ctx_prime.argtypes = [C.c_void_p, C.c_int]
    lib.ora_ctx_psi.argtypes = [C.c_void_p, C.c_int]
    lib.ora_mulmod.argtypes = [C.c_uint64] * 3
    lib.ora_powmod.argtypes = [C.c_uint64] * 3
    lib.ora_invmod.argtypes = [C.c_uint64] * 2
    lib.ora_is_prime.argtypes = [C.c_uint64]
    lib.ora_min_primitive_root.argtypes = [C.c_uint64] * 2
    lib.ora_gen_primes.argtypes = [C.c_uint64, C.POINTER(C.c_int), C.c_int, u64p]
    lib.ora_galois_elt_from_step.argtypes = [C.c_uint64, C.c_int]
    lib.ora_galois_table.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]
    lib.ora_ntt_fwd.argtypes = [C.c_void_p, C.c_int, u64p]
    lib.ora_ntt_inv.argtypes = [C.c_void_p, C.c_int, u64p]
    lib.ora_keygen.restype = C.c_void_p
    lib.ora_keygen.argtypes = [C.c_void_p, C.c_uint64]
    lib.ora_keys_destroy.argtypes = [C.c_void_p]
    for n in ("ora_keys_secret", "ora_keys_public", "ora_keys_relin"):
        getattr(lib, n).restype = u64p
        getattr(lib, n).argtypes = [C.c_void_p]
    lib.ora_keys_galois.restype = u64p
    lib.ora_keys_galois.argtypes = [C.c_void_p, C.c_uint64]
    return lib


lib = _load()


def _p(a):
    assert a.dtype == np.uint64 and a.flags.c_contiguous
    return a.ctypes.data_as(u64p)


def gen_primes(N, bits):
    arr = (C.c_int * len(bits))(*bits)
    out = np.zeros(len(bits), dtype=np.uint64)
    rc = lib.ora_gen_primes(N, arr, len(bits), _p(out))
    if rc:
        raise RuntimeError("failed to find enough qualifying primes")
    return [int(x) for x in out]


def galois_elt_from_step(N, steps):
    return int(lib.ora_galois_elt_from_step(N, steps))


def galois_table(N, elt):
    t = np.zeros(N, dtype=np.uint32)
    lib.ora_galois_table(N, elt, t.ctypes.data_as(C.POINTER(C.c_uint32)))
    return t


class Oracle:
    """One CKKS context (N, primes in SEAL order; last prime = special prime).

    Level j has ell = k-1-j residues (reference eva/seal/seal_executor.h:221-224).
    Ciphertexts are numpy uint64 arrays [size, ell, N]; plaintexts [ell, N];
    key-switch keys [k-1, 2, k, N].
    """

    def __init__(self, N, prime_bits=None, primes=None):
        if primes is None:
            arr = (C.c_int * len(prime_bits))(*prime_bits)
            self.h = lib.ora_ctx_create(N, arr, len(prime_bits))
        else:
            pa = np.array(primes, dtype=np.uint64)
            self.h = lib.ora_ctx_create_from_primes(N, _p(pa), len(primes))
        if not self.h:
            raise RuntimeError("oracle context creation failed")
        self.N = N
        self.k = lib.ora_ctx_k(self.h)
        self.primes = [int(lib.ora_ctx_prime(self.h, i)) for i in range(self.k)]
        self.psi = [int(lib.ora_ctx_psi(self.h, i)) for i in range(self.k)]
        self.keys = None

    def __del__(self):
        try:
            if self.keys:
                lib.ora_keys_destroy(self.keys)
            if self.h:
                lib.ora_ctx_destroy(self.h)
        except Exception:
            pass

    def ell(self, level):
        return self.k - 1 - level

    # ---- NTT on a [*, N] stack using prime indices ----
    def ntt_fwd(self, a, prime_idx):
        a = np.ascontiguousarray(a, dtype=np.uint64).copy()
        lib.ora_ntt_fwd(self.h, prime_idx, _p(a))
        return a

    def ntt_inv(self, a, prime_idx):
        a = np.ascontiguousarray(a, dtype=np.uint64).copy()
        lib.ora_ntt_inv(self.h, prime_idx, _p(a))
        return a

    # ---- evaluator ----
    def _bin(self, fn, a, b):
        ell = a.shape[1]
        out = np.empty((max(a.shape[0], b.shape[0]), ell, self.N), dtype=np.uint64)
        rc = getattr(lib, fn)(C.c_void_p(self.h), ell, _p(out), _p(a), a.shape[0], _p(b), b.shape[0])
        assert rc == 0
        return out

    def add(self, a, b):
        return self._bin("ora_add", a, b)

    def sub(self, a, b):
        return self._bin("ora_sub", a, b)

    def _plain(self, fn, a, pt):
        out = np.empty_like(a)
        rc = getattr(lib, fn)(C.c_void_p(self.h), a.shape[1], _p(out), _p(a), a.shape[0], _p(pt))
        assert rc == 0
        return out

    def add_plain(self, a, pt):
        return self._plain("ora_add_plain", a, pt)

    def sub_plain(self, a, pt):
        return self._plain("ora_sub_plain", a, pt)

    def mul_plain(self, a, pt):
        return self._plain("ora_mul_plain", a, pt)

    def negate(self, a):
        out = np.empty_like(a)
        assert lib.ora_negate(C.c_void_p(self.h), a.shape[1], _p(out), _p(a), a.shape[0]) == 0
        return out

    def mul(self, a, b):
        out = np.empty((3, a.shape[1], self.N), dtype=np.uint64)
        assert lib.ora_mul(C.c_void_p(self.h), a.shape[1], _p(out), _p(a), _p(b)) == 0
        return out

    def square(self, a):
        out = np.empty((3, a.shape[1], self.N), dtype=np.uint64)
        assert lib.ora_square(C.c_void_p(self.h), a.shape[1], _p(out), _p(a)) == 0
        return out

    def keyswitch(self, t, key):
        out = np.empty((2, t.shape[0], self.N), dtype=np.uint64)
        assert lib.ora_keyswitch(C.c_void_p(self.h), t.shape[0], _p(out), _p(t), _p(key)) == 0
        return out

    def relinearize(self, a, rk):
        out = np.empty((2, a.shape[1], self.N), dtype=np.uint64)
        assert lib.ora_relinearize(C.c_void_p(self.h), a.shape[1], _p(out), _p(a), _p(rk)) == 0
        return out

    def apply_galois(self, a, elt):
        out = np.empty_like(a)
        assert lib.ora_apply_galois(C.c_void_p(self.h), a.shape[1], _p(out), _p(a), a.shape[0], C.c_uint64(elt)) == 0
        return out

    def rotate(self, a, steps, gk):
        """rotate_vector(a, steps) with the direct Galois key (steps != 0)."""
        elt = galois_elt_from_step(self.N, steps)
        out = np.empty((2, a.shape[1], self.N), dtype=np.uint64)
        assert lib.ora_rotate(C.c_void_p(self.h), a.shape[1], _p(out), _p(a), C.c_uint64(elt), _p(gk)) == 0
        return out

    def rescale(self, a):
        out = np.empty((a.shape[0], a.shape[1] - 1, self.N), dtype=np.uint64)
        assert lib.ora_rescale(C.c_void_p(self.h), a.shape[1], _p(out), _p(a), a.shape[0]) == 0
        return out

    def mod_switch(self, a):
        out = np.empty((a.shape[0], a.shape[1] - 1, self.N), dtype=np.uint64)
        assert lib.ora_mod_switch(C.c_void_p(self.h), a.shape[1], _p(out), _p(a), a.shape[0]) == 0
        return out

    # ---- encoder ----
    def encode(self, values, scale, ell):
        v = np.ascontiguousarray(values, dtype=np.float64)
        if v.size < self.N // 2:  # replicate (reference seal.cpp:71-79, seal_executor.h:229-240)
            assert (self.N // 2) % v.size == 0
            v = np.tile(v, (self.N // 2) // v.size)
        pt = np.empty((ell, self.N), dtype=np.uint64)
        rc = lib.ora_encode(C.c_void_p(self.h), ell, v.ctypes.data_as(f64p), C.c_size_t(v.size),
                            C.c_double(scale), _p(pt))
        assert rc == 0
        return pt

    def decode(self, pt, scale):
        out = np.empty(self.N // 2, dtype=np.float64)
        rc = lib.ora_decode(C.c_void_p(self.h), pt.shape[0], _p(pt), C.c_double(scale), out.ctypes.data_as(f64p))
        assert rc == 0
        return out

    # ---- client side ----
    def keygen(self, seed=1):
        if self.keys:
            lib.ora_keys_destroy(self.keys)
        self.keys = lib.ora_keygen(self.h, seed)
        assert self.keys
        return self

    def _view(self, ptr, shape):
        n = int(np.prod(shape))
        return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape).copy()

    def secret_key(self):
        return self._view(lib.ora_keys_secret(self.keys), (self.k, self.N))

    def relin_key(self):
        return self._view(lib.ora_keys_relin(self.keys), (self.k - 1, 2, self.k, self.N))

    def galois_key(self, elt):
        return self._view(lib.ora_keys_galois(self.keys, elt), (self.k - 1, 2, self.k, self.N))

    def encrypt(self, pt, seed=7):
        ell = pt.shape[0]
        ct = np.empty((2, ell, self.N), dtype=np.uint64)
        assert lib.ora_encrypt(C.c_void_p(self.keys), ell, _p(pt), C.c_uint64(seed), _p(ct)) == 0
        return ct

    def encrypt_with(self, pt, u, e0, e1):
        """encrypt with explicit randomness (ternary u, small e0 / e1: int32 arrays of N)"""
        ell = pt.shape[0]
        ct = np.empty((2, ell, self.N), dtype=np.uint64)
        ip = C.POINTER(C.c_int)
        arrs = [np.ascontiguousarray(x, dtype=np.int32) for x in (u, e0, e1)]
        assert lib.ora_encrypt_with(C.c_void_p(self.keys), ell, _p(pt), *[a.ctypes.data_as(ip) for a in arrs], _p(ct)) == 0
        return ct

    def public_key(self):
        return self._view(lib.ora_keys_public(self.keys), (2, self.k, self.N))

    def decrypt(self, ct):
        pt = np.empty((ct.shape[1], self.N), dtype=np.uint64)
        assert lib.ora_decrypt(C.c_void_p(self.keys), ct.shape[1], _p(ct), ct.shape[0], _p(pt)) == 0
        return pt


def splitmix64_fill(seed, n, moduli=None):
    """Deterministic uniform u64 stream (SplitMix64), optionally reduced mod q.

    Used to make synthetic residues for NTT tests/benchmarks.
    """
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    if moduli is not None:
        z = z % np.uint64(moduli)
    return z
