"""One wide program sharded across the GPUs of a box (BASELINE config "synthetic wide DAG, >= 4096
parallel ciphertext multiplications").  Run with

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 tools/shard_bench.py [--products 4096]

Every rank holds the same keys and inputs (synthetic uniform residues), runs its part of the DAG, rank 0
gathers the partial sums over NCCL (the only exchange) and finishes the program.  Prints ONE JSON line on
rank 0: latency of the sharded execute (host buffers in and out, max over ranks) next to the single-GPU
execute of the same program on rank 0, and whether the two results are bit-identical."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_wide(products, vec=8192):
    """tests/large_programs.py style: sum_i rot(x, i % 64) * rot(y, (i / 64) % 64), one relinearize"""
    from eva_b200 import EvaProgram, Input, Output
    from eva_b200.ckks import CKKSCompiler
    prog = EvaProgram("wide%d" % products, vec_size=vec)
    with prog:
        x, y = Input("x"), Input("y")
        xs = [x << i for i in range(min(64, products))]
        ys = [y << j for j in range(min(64, (products + 63) // 64))]
        terms = [xs[i % 64] * ys[(i // 64) % 64] for i in range(products)]
        while len(terms) > 1:
            terms = [terms[i] + terms[i + 1] if i + 1 < len(terms) else terms[i] for i in range(0, len(terms), 2)]
        Output("z", terms[0])
    prog.set_input_scales(40)
    prog.set_output_ranges(30)
    return CKKSCompiler(config={"warn_vec_size": "false"}).compile(prog)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--products", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from eva_b200 import b200, multi, shard
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    prog, params, sig = build_wide(args.products)
    N = params.poly_modulus_degree
    primes = b200.create_coeff_modulus(N, list(params.prime_bits))
    k = len(primes)
    rng = np.random.default_rng(7)   # same keys and inputs on every rank (replicated)

    def uni(prefix, nres):
        a = np.empty(tuple(prefix) + (nres, N), dtype=np.uint64)
        for j in range(nres):
            a[..., j, :] = rng.integers(0, primes[j], size=tuple(prefix) + (N,), dtype=np.uint64)
        return a
    relin = uni((k - 1, 2), k)
    galois = {}
    for s in sorted(params.rotations):
        if s != 0:
            galois[pow(3, s if s > 0 else N // 2 + s, 2 * N)] = uni((k - 1, 2), k)
    pub = b200.context_from_raw_keys(N, primes, relin, galois, local)
    val = b200.B200Valuation()
    for name, info in sig.inputs.items():
        val.set_cipher(name, uni((2,), k - 1 - info.level), 2.0 ** info.scale)
    nops = pub.cipher_op_count(prog)
    plan = shard.split_program(prog, world) if world > 1 else None

    def timed(fn):
        for _ in range(args.warmup):
            out = fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        return out, dt

    t_single = out_single = None
    if rank == 0:
        for _ in range(args.warmup):
            out_single = pub.execute(prog, val)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out_single = pub.execute(prog, val)
        t_single = (time.perf_counter() - t0) / args.steps
        pub.drop_plan(prog, 1)     # free the single-GPU arena before the sharded runs
    if world == 1:
        print(json.dumps({"workload": "wide DAG: %d ciphertext products" % args.products, "cipher_ops": nops, "n_gpus": 1,
                          "single_gpu_ms": t_single * 1e3, "ops_per_s_single": nops / t_single}))
        return
    out, t_shard = timed(lambda: shard.execute_sharded(pub, prog, val, rank, world, plan, device="cuda"))
    if world > 1:
        (t_shard,) = multi.max_over_ranks([t_shard], world, device="cuda")
    if rank == 0:
        same = bool(np.array_equal(out.get("z")[1], out_single.get("z")[1]))
        print(json.dumps({"workload": "wide DAG: %d ciphertext products of rotated inputs, tree sum, relinearize; N=%d prime_bits=%s" % (args.products, N, list(params.prime_bits)),
                          "cipher_ops": nops, "n_gpus": world, "steps": args.steps,
                          "single_gpu_ms": t_single * 1e3, "sharded_ms": t_shard * 1e3, "speedup": t_single / t_shard,
                          "ops_per_s_single": nops / t_single, "ops_per_s_sharded": nops / t_shard,
                          "bit_identical_to_single_gpu": same,
                          "leaves_per_part": plan.leaves_per_part if plan else None,
                          "exchange": "NCCL gather of %d partial ciphertexts of size %d (the only collective)" % (world, plan.partial_size) if plan else None,
                          "note": "host buffers in and out on every call; shared rotations are recomputed per part (no cross-part edges)"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
