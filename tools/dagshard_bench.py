"""DAG-sharded execution of one compiled program across the GPUs of a box (BASELINE configs 4 and 5): run under

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 tools/dagshard_bench.py [--workload wide4096|harris|sobel]

Every rank holds the same keys and inputs (synthetic uniform residues: timing does not depend on the values),
runs its stage programs, partial sums travel over NCCL from arena to arena.  Rank 0 prints one JSON line:
single-GPU latency of the same program (as compiled, and with duplicate terms aliased), sharded latency
(max over ranks, host buffers in and out), speed-up, and whether the two results are bit-identical."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def setup(workload, local):
    from eva_b200 import b200, program_io
    d = program_io.load_json(workload)
    prog, params, sig, _ = program_io.build_program(d)
    N = d["poly_modulus_degree"]
    primes = b200.create_coeff_modulus(N, list(d["prime_bits"]))
    k = len(primes)
    rng = np.random.default_rng(7)   # same keys and inputs on every rank (replicated)

    def uni(prefix, nres):
        a = np.empty(tuple(prefix) + (nres, N), dtype=np.uint64)
        for j in range(nres):
            a[..., j, :] = rng.integers(0, primes[j], size=tuple(prefix) + (N,), dtype=np.uint64)
        return a
    relin = uni((k - 1, 2), k)
    galois = {}
    for s in sorted(d["rotations"]):
        if s != 0:
            galois[pow(3, s if s > 0 else N // 2 + s, 2 * N)] = uni((k - 1, 2), k)
    pub = b200.context_from_raw_keys(N, primes, relin, galois, local)
    val = b200.B200Valuation()
    for name, info in d["signature"].items():
        val.set_cipher(name, uni((2,), k - 1 - info["level"]), 2.0 ** info["scale"])
    return d, prog, pub, val


def measure(workload, rank, world, local, steps, warmup, quiet=False):
    import torch
    import torch.distributed as dist
    from eva_b200 import dagshard, multi
    d, prog, pub, val = setup(workload, local)
    nops = pub.cipher_op_count(prog)
    res = {"workload": "%s compiled by the reference compiler: N=%d prime_bits=%s, %d ciphertext ops" % (workload, d["poly_modulus_degree"], d["prime_bits"], nops),
           "cipher_ops": nops, "n_gpus": world}

    def timed(fn, collective=False):
        for _ in range(warmup):
            out = fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(steps):
            if collective:       # every rank is here (the single-GPU runs above are rank 0's alone: no barrier there)
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return out, float(np.median(ts))
    single = None
    if rank == 0:
        if nops <= 2000:     # the program exactly as compiled, every term evaluated (skipped for the 16k-term DAG: seconds per run)
            pub.set_options(dedup_terms=False)
            _, t_raw = timed(lambda: pub.execute(prog, val))
            pub.drop_plan(prog, 1)
            res["ms_single_gpu_no_dedup"] = t_raw * 1e3
        pub.set_options(dedup_terms=True)
        single, t_single = timed(lambda: pub.execute(prog, val))
        pub.drop_plan(prog, 1)
        res["ms_single_gpu"] = t_single * 1e3
    if world > 1:
        plan = dagshard.plan_stages(prog, world)
        if plan is None:
            if rank == 0:
                res["note"] = "the cost model finds no cut that pays"
                if not quiet:
                    print(json.dumps(res))
            return res
        runner = dagshard.ShardedRunner(pub, plan, rank, world)
        dist.barrier()
        out, t_shard = timed(lambda: runner.run(val), collective=True)
        (t_shard,) = multi.max_over_ranks([t_shard], world, device="cuda")
        if rank == 0:
            same = all(np.array_equal(out.get(o)[1], single.get(o)[1]) for o in d["outputs"])
            sp = res["ms_single_gpu"] / (t_shard * 1e3)
            res.update({"ms_sharded": t_shard * 1e3, "speedup": sp, "bit_identical": bool(same),
                        "auto_choice": "sharded" if sp > 1.03 else "single GPU (dagshard.AutoShardedRunner measures both and keeps the faster: this plan does not pay at this GPU count)",
                        "stages": plan.describe(), "exchange": "NCCL all_gather_into_tensor between stages, gather on rank 0 for the last cut; device pointers, no host staging"})
    if rank == 0 and not quiet:
        print(json.dumps(res), flush=True)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="wide4096")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for w in args.workload.split(","):
        measure(w, rank, world, local, args.steps, args.warmup)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
