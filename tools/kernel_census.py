"""One launch of every kernel of the evaluator at the Sobel shape (N = 16384, level-0 ell = 4, 64 instances per launch),
inside a cudaProfiler range, for the per-kernel ncu table of profiles/ (run under gpurun):

    ncu --profile-from-start off --metrics <see tools/kernel_census_report.py> --csv --log-file gpurun_out/census.csv python tools/kernel_census.py

Prints one JSON line per op with the algorithmic bytes per launch so that the report can set measured DRAM
traffic against them."""
import ctypes as C, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from eva_b200 import cabi
from tools.microbench import gen_primes
lib = cabi.load(); torch.cuda.set_device(0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
N, k, ell, B = 16384, 5, 4, 64
R = 8 * N
pa = np.array(gen_primes(N, [60] * k), dtype=np.uint64)
h = C.c_void_p(); assert lib.evab_ctx_create(N, pa.ctypes.data_as(cabi.u64p), k, 0, C.byref(h)) == 0
lib.evab_galois_prepare(h, 3)
key = torch.randint(0, 1 << 59, (k - 1, 2, k, N), dtype=torch.int64, device="cuda")
P = lambda t: C.c_void_p(t.data_ptr())
ws = lib.evab_keyswitch_work_bytes(h, ell) // 8
stride = (3 * ell * N * 4 + ws + 64 + 63) // 64 * 64          # a3 | a2 | b2 | out | work per instance
buf = torch.randint(0, 1 << 59, (B, stride), dtype=torch.int64, device="cuda")
a3, a2 = buf[0, :3 * ell * N], buf[0, 3 * ell * N:5 * ell * N]
b2, hoist = buf[0, 5 * ell * N:7 * ell * N], buf[0, 7 * ell * N:8 * ell * N]
out, work = buf[0, 9 * ell * N:12 * ell * N], buf[0, 12 * ell * N:12 * ell * N + ws]
pt = torch.randint(0, 1 << 59, (ell, N), dtype=torch.int64, device="cuda")   # plaintext shared by the instances (stride 0 would need a second batch stride; per-instance copy below)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def sum_products():
    n = 9
    cts = (C.c_void_p * n)(*[a2.data_ptr()] * n)
    pts = (C.c_void_p * n)(*[b2.data_ptr()] * n)
    sizes = (C.c_int * n)(*[2] * n)
    kinds = (C.c_int * n)(*[1] * n)
    return lib.evab_sum_products(h, ell, P(out), n, cts, sizes, pts, kinds, st)


ops = [
    ("add", lambda: lib.evab_add(h, ell, P(out), P(a2), 2, P(b2), 2, st), 3 * 2 * ell * R),
    ("negate", lambda: lib.evab_negate(h, ell, P(out), P(a2), 2, st), 2 * 2 * ell * R),
    ("multiply_plain", lambda: lib.evab_mul_plain(h, ell, P(out), P(a2), 2, P(b2), st), 5 * ell * R),
    ("multiply", lambda: lib.evab_mul(h, ell, P(out), P(a2), P(b2), st), 7 * ell * R),
    ("square", lambda: lib.evab_square(h, ell, P(out), P(a2), st), 5 * ell * R),
    ("sum of 9 multiply_plain (fused)", sum_products, (9 * 3 + 2) * ell * R),
    ("mod_switch", lambda: lib.evab_mod_switch(h, ell, P(out), P(a2), 2, st), 2 * 2 * (ell - 1) * R),
    ("rescale (size 3)", lambda: lib.evab_rescale(h, ell, P(out), P(a3), 3, P(work), st), 3 * (2 * ell - 1) * R),
    ("relinearize", lambda: lib.evab_relinearize(h, ell, P(out), P(a3), P(key), P(work), st), (2 * ell * ell + 7 * ell) * R),
    ("rotate", lambda: lib.evab_rotate(h, ell, P(out), P(a2), 3, P(key), P(work), st), (2 * ell * ell + 6 * ell) * R),
    ("rotate_prepare", lambda: lib.evab_rotate_prepare(h, ell, P(hoist), P(a2), st), 2 * ell * R),
    ("rotate_prepared", lambda: lib.evab_rotate_prepared(h, ell, P(out), P(a2), P(hoist), 3, P(key), P(work), st), (2 * ell * ell + 5 * ell) * R),
]
# rotations sharing inverse NTT and mod-up (evab_rotate_modup_*): the prepare once per ciphertext, then per rotation
ext = torch.empty((B, lib.evab_rotate_modup_ext_bytes(h, ell) // 8), dtype=torch.int64, device="cuda")
# the per-instance buffers must follow the batch stride: carve them out of a second strided block
stride2 = (ell * N + lib.evab_rotate_modup_ext_bytes(h, ell) // 8 + 64 + 63) // 64 * 64
assert stride2 <= stride, "census layout: the hoist buffers must fit the instance stride"
that2, ext2, zf2 = buf[0, 7 * ell * N:8 * ell * N], hoist, None
cadd = torch.empty(lib.evab_hoist_const_bytes(h, ell) // 8, dtype=torch.int64, device="cuda")
ctmp = torch.empty((ell + 1) * N, dtype=torch.int64, device="cuda")
assert lib.evab_rotate_hoist_const(h, ell, 3, P(key), P(cadd), P(ctmp), st) == 0
big = torch.randint(0, 1 << 59, (B, stride), dtype=torch.int64, device="cuda")     # that | ext | flag of every instance, same stride
bthat, bext, bflag = big[0, :ell * N], big[0, ell * N:ell * N + (ell + 1) * ell * N], big[0, ell * N + (ell + 1) * ell * N:ell * N + (ell + 1) * ell * N + 8]
ops += [
    ("rotate_modup_prepare (once per ciphertext)", lambda: lib.evab_rotate_modup_prepare(h, ell, P(bthat), P(bext), P(a2), P(bflag), st), (ell + ell + (ell + 1) * ell) * R),
    ("rotate_modup_prepared (per rotation)", lambda: lib.evab_rotate_modup_prepared(h, ell, P(out), P(a2), P(bext), 3, P(key), P(cadd), P(work), st), (ell * (ell + 1) + 2 * ell * (ell + 1) + 4 * ell) * R),
]
for name, fn, _ in ops:       # warm-up outside the profiled range
    lib.evab_set_batch(B, stride, 0); assert fn() == 0; lib.evab_set_batch(1, 0, 0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for name, fn, algo in ops:
    flush.zero_()
    lib.evab_set_batch(B, stride, 0)
    l0 = lib.evab_launch_count(h)
    assert fn() == 0
    n = lib.evab_launch_count(h) - l0
    lib.evab_set_batch(1, 0, 0)
    torch.cuda.synchronize()
    print(json.dumps({"op": name, "launches": int(n), "instances": B, "algorithmic_bytes_per_instance": algo}), flush=True)
# ---- 8 rotations of one ciphertext in one call (evab_rotate_modup_many) and two weighted sums over them (evab_lazy_rotsum, opt-in,
# not bit-exact): their own instance stride (a | that | ext | flag | 2 weights | out | work), 16 instances per launch
nrot, B2 = 8, 16
elts = [pow(3, i + 1, 2 * N) for i in range(nrot)]
for e_ in elts:
    assert lib.evab_galois_prepare(h, C.c_uint64(e_)) == 0
wm = max(lib.evab_rotate_modup_many_work_bytes(h, ell, nrot), lib.evab_lazy_rotsum_work_bytes(h, ell, 2)) // 8
o_a, o_that = 0, 2 * ell * N
o_ext = o_that + ell * N
o_flag = o_ext + (ell + 1) * ell * N
o_w = o_flag + 8
o_out = o_w + 2 * (ell + 1) * N
o_work = o_out + nrot * 2 * ell * N
stride3 = (o_work + wm + 63) // 64 * 64
big3 = torch.randint(0, 1 << 59, (B2, stride3), dtype=torch.int64, device="cuda")
sl = lambda off: C.c_void_p(big3.data_ptr() + 8 * off)
keys = (C.c_void_p * nrot)(*[key.data_ptr()] * nrot)
cadds = (C.c_void_p * nrot)(*[cadd.data_ptr()] * nrot)
eltsA = (C.c_uint64 * nrot)(*elts)
wts = (C.c_void_p * (2 * nrot))(*([big3.data_ptr() + 8 * o_w] * nrot + [big3.data_ptr() + 8 * (o_w + (ell + 1) * N)] * nrot))
ops2 = [
    ("rotate_modup_prepare + scale_c0 (once per ciphertext)", lambda: lib.evab_rotate_modup_prepare(h, ell, sl(o_that), sl(o_ext), sl(o_a), sl(o_flag), st)
     or lib.evab_rotate_modup_scale_c0(h, ell, sl(o_ext), sl(o_a), st), (ell + ell + (ell + 1) * ell + 2 * ell) * R),
    ("rotate_modup_many (8 rotations of one ciphertext)", lambda: lib.evab_rotate_modup_many(h, ell, nrot, sl(o_out), sl(o_a), sl(o_ext), eltsA, keys, cadds, sl(o_work), st),
     (ell * (ell + 1) + nrot * (2 * ell * (ell + 1) + 2 * 2 * (ell + 1) + 2 * ell)) * R),
    ("lazy_rotsum (2 weighted sums over 8 rotations; opt-in, not bit-exact)", lambda: lib.evab_lazy_rotsum(h, ell, 2, sl(o_out), sl(o_a), sl(o_ext), nrot, eltsA, keys, cadds, wts, sl(o_work), st),
     (ell * (ell + 1) + nrot * 2 * ell * (ell + 1) + 2 * (ell + 1) + 2 * (2 * 2 * (ell + 1) + 2 * ell)) * R),
]
torch.cuda.synchronize()
torch.cuda.profiler.stop()            # warm-up outside the profiled range
for name, fn, _ in ops2:
    lib.evab_set_batch(B2, stride3, 0); assert fn() == 0; lib.evab_set_batch(1, 0, 0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for name, fn, algo in ops2:
    flush.zero_()
    lib.evab_set_batch(B2, stride3, 0)
    l0 = lib.evab_launch_count(h)
    assert fn() == 0
    n = lib.evab_launch_count(h) - l0
    lib.evab_set_batch(1, 0, 0)
    torch.cuda.synchronize()
    print(json.dumps({"op": name, "launches": int(n), "instances": B2, "algorithmic_bytes_per_instance": algo}), flush=True)
# the encoder (E row): one batch of 23 vectors as in a Sobel execute (dense 4096-slot vectors) + replicated scalars
cnt = 8
vals = torch.rand((cnt, N // 2), dtype=torch.float64, device="cuda")
ptrs = (C.c_void_p * cnt)(*[vals[i].data_ptr() for i in range(cnt)])
vec = (C.c_uint32 * cnt)(*[N // 2] * cnt)
sc = (C.c_double * cnt)(*[2.0 ** 30] * cnt)
wbytes = lib.evab_encode_work_bytes(h, cnt)
ework = torch.empty(wbytes, dtype=torch.uint8, device="cuda")
eout = torch.empty((cnt, ell, N), dtype=torch.int64, device="cuda")
l0 = lib.evab_launch_count(h)
assert lib.evab_encode(h, cnt, ptrs, vec, sc, ell, P(eout), P(ework), st) == 0
print(json.dumps({"op": "encode (8 dense vectors)", "launches": int(lib.evab_launch_count(h) - l0), "instances": 1, "algorithmic_bytes_per_instance": cnt * (N // 2 * 8 + ell * R)}), flush=True)
uv = (C.c_double * cnt)(*[0.5] * cnt)
l0 = lib.evab_launch_count(h)
assert lib.evab_encode_uniform(h, cnt, uv, sc, ell, P(eout), st) == 0
print(json.dumps({"op": "encode_uniform (8 scalars)", "launches": int(lib.evab_launch_count(h) - l0), "instances": 1, "algorithmic_bytes_per_instance": cnt * ell * R}), flush=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
lib.evab_ctx_destroy(h)
