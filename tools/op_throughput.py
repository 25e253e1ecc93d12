"""Steady-state throughput of the evaluator ops at the Sobel shape with fat (batched) launches:
how much SM-time one residue-NTT costs inside each fused op (run under gpurun)."""
import ctypes as C, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from eva_b200 import cabi
from tools.microbench import gen_primes, timeit
lib = cabi.load(); torch.cuda.set_device(0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
N, k = 16384, 5
pa = np.array(gen_primes(N, [60] * k), dtype=np.uint64)
h = C.c_void_p(); assert lib.evab_ctx_create(N, pa.ctypes.data_as(cabi.u64p), k, 0, C.byref(h)) == 0
lib.evab_galois_prepare(h, 3)
B = 64
key = torch.randint(0, 1 << 59, (k - 1, 2, k, N), dtype=torch.int64, device="cuda")
P = lambda t: C.c_void_p(t.data_ptr())
for ell in (4, 3):
    ws = lib.evab_keyswitch_work_bytes(h, ell) // 8
    stride = (3 * ell * N * 3 + ws + 64 + 63) // 64 * 64          # a3 | a2 | out | work per instance
    buf = torch.randint(0, 1 << 59, (B, stride), dtype=torch.int64, device="cuda")
    a3, a2, out, work = buf[0, :3*ell*N], buf[0, 3*ell*N:5*ell*N], buf[0, 6*ell*N:9*ell*N], buf[0, 9*ell*N:9*ell*N+ws]
    hoist = torch.randint(0, 1 << 59, (B, stride), dtype=torch.int64, device="cuda")   # same stride: instance b at b*stride
    def rot_prepared():
        return lib.evab_rotate_prepared(h, ell, P(out), P(a2), P(buf[0, 5*ell*N:6*ell*N]), 3, P(key), P(work), st)
    ops = {"rotate_prepared": (rot_prepared, ell*ell + 2 + 2*ell),
           "rotate_prepare": (lambda: lib.evab_rotate_prepare(h, ell, P(buf[0, 5*ell*N:6*ell*N]), P(a2), st), ell),
           "relinearize": (lambda: lib.evab_relinearize(h, ell, P(out), P(a3), P(key), P(work), st), ell + ell*ell + 2 + 2*ell),
           "rotate": (lambda: lib.evab_rotate(h, ell, P(out), P(a2), 3, P(key), P(work), st), ell + ell*ell + 2 + 2*ell),
           "rescale3": (lambda: lib.evab_rescale(h, ell, P(out), P(a3), 3, P(work), st), 3 + 3*(ell-1)),
           "mul_plain": (lambda: lib.evab_mul_plain(h, ell, P(out), P(a2), 2, P(a3), st), 0),
           "add": (lambda: lib.evab_add(h, ell, P(out), P(a2), 2, P(a3), 2, st), 0)}
    for name, (fn, ntts) in ops.items():
        lib.evab_set_batch(B, stride, 0)
        ms = timeit(lambda: fn(), iters=10, warmup=3)
        lib.evab_set_batch(1, 0, 0)
        d = {"op": name, "ell": ell, "batch": B, "ms": ms, "us_per_op": ms * 1e3 / B, "residue_ntts_per_op": ntts}
        if ntts: d["sm_us_per_residue_ntt"] = ms * 1e3 * 148 / (B * ntts)
        print(json.dumps(d))
