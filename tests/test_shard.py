"""Sharding one compiled program across GPUs (eva_b200/shard.py): graph surgery checked with the plaintext
reference semantics and over gloo on CPU; bit-exactness of the sharded execution on the GPU."""
import os
import socket
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from eva_b200 import evaluate, program_io, shard  # noqa: E402

_FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "programs")
if _FIX not in program_io.SEARCH:      # also in the spawned gloo workers, which never see conftest.py
    program_io.SEARCH.append(_FIX)


def _plain_sharded(prog, plan, x):
    partials = {"partial_%d" % r: evaluate(p, {k: x[k] for k in p.inputs})["partial"] for r, p in enumerate(plan.parts)}
    tin = dict(partials)
    tin.update({k: x[k] for k in plan.tail.inputs if k in x})
    return evaluate(plan.tail, tin)


@pytest.mark.parametrize("name,parts", [("wide64", 2), ("wide64", 4), ("wide64", 8), ("sobel", 2), ("sobel", 4), ("harris", 2), ("harris", 8),
                                        ("polynomial", 2), ("feat_hsum", 2)])
def test_split_preserves_semantics(name, parts):
    d = program_io.load_json(name)
    prog = program_io.build_program(d)[0]
    plan = shard.split_program(prog, parts)
    if plan is None:
        pytest.skip("no sum with %d leaves" % parts)
    assert len(plan.parts) == parts and sum(plan.leaves_per_part) >= parts and min(plan.leaves_per_part) >= 1
    assert set(plan.tail.outputs) == set(prog.outputs)
    for p in plan.parts:
        assert set(p.outputs) == {"partial"} and set(p.inputs) <= set(prog.inputs)
    rng = np.random.default_rng(3)
    x = {k: list(rng.uniform(-1, 1, prog.vec_size)) for k in prog.inputs}
    want, got = evaluate(prog, x), _plain_sharded(prog, plan, x)
    for k in want:
        assert np.allclose(got[k], want[k], rtol=1e-9, atol=1e-9)


def test_wide_dag_splits_evenly_without_a_tail():
    d = program_io.load_json("wide64")
    prog = program_io.build_program(d)[0]
    plan = shard.split_program(prog, 4)
    assert plan.leaves_per_part == [16, 16, 16, 16] and plan.partial_size == 3
    assert max(len(p.terms()) for p in plan.parts) <= len(prog.terms()) // 4 + 4     # no recomputation to speak of
    assert len(plan.tail.terms()) <= 10                                               # inputs + 3 adds + relinearize + output
    assert shard.split_program(prog, 1) is None and shard.split_program(program_io.build_program(program_io.load_json("feat_unary"))[0], 2) is None


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from eva_b200 import multi
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = program_io.load_json("wide64")
        prog = program_io.build_program(d)[0]
        plan = shard.split_program(prog, world)
        rng = np.random.default_rng(5)
        x = {k: list(rng.uniform(-1, 1, prog.vec_size)) for k in sorted(prog.inputs)}
        part = np.asarray(evaluate(plan.parts[rank], {k: x[k] for k in plan.parts[rank].inputs})["partial"], dtype=np.float64)
        gathered = multi.gather_outputs(part.view(np.uint64), rank, world)      # the one exchange of the sharded DAG
        if rank == 0:
            tin = {"partial_%d" % r: list(g.view(np.float64)) for r, g in enumerate(gathered)}
            got = evaluate(plan.tail, tin)
            want = evaluate(prog, x)
            q.put(all(np.allclose(got[k], want[k], rtol=1e-9, atol=1e-9) for k in want))
        else:
            q.put(gathered is None)
    finally:
        dist.destroy_process_group()


def test_sharded_flow_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [True, True]


@pytest.mark.gpu
@pytest.mark.parametrize("name,parts", [("wide64", 2), ("wide64", 4), ("sobel", 2), ("harris", 4)])
def test_sharded_execution_bit_exact(name, parts):
    """the P parts (run here one after the other on one GPU) + tail give the very bits of execute(prog)"""
    from eva_b200 import b200
    from oracle import oracle as o
    from oracle_exec import OracleProgram
    d = program_io.load_json(name)
    prog = program_io.build_program(d)[0]
    N = d["poly_modulus_degree"]
    orc = o.Oracle(N, d["prime_bits"]).keygen(2)
    op = OracleProgram(d, orc)
    op.prepare_keys()
    pub = b200.context_from_raw_keys(N, orc.primes, op.rk, {int(e): k for e, k in op.gks.items()})
    rng = np.random.default_rng(8)
    val = b200.B200Valuation()
    for name_, info in d["signature"].items():
        x = rng.uniform(0, 0.2, d["vec_size"])
        val.set_cipher(name_, orc.encrypt(orc.encode(x, 2.0 ** info["scale"], orc.k - 1 - info["level"]), seed=31), 2.0 ** info["scale"])
    want = pub.execute(prog, val)
    plan = shard.split_program(prog, parts)
    partials, scale = [], None
    for r in range(parts):
        p, scale = shard.run_part(pub, plan, r, val)
        assert p.shape[0] == plan.partial_size
        partials.append(p)
    got = shard.run_tail(pub, plan, partials, scale, val)
    for oname in d["outputs"]:
        assert np.array_equal(got.get(oname)[1], want.get(oname)[1]) and got.get(oname)[2] == want.get(oname)[2]
    # world = 1 degenerates to execute()
    assert np.array_equal(shard.execute_sharded(pub, prog, val).get(list(d["outputs"])[0])[1], want.get(list(d["outputs"])[0])[1])


@pytest.mark.parametrize("seed", range(24))
def test_split_random_programs(seed):
    """random DSL programs (sums of products of rotated / scaled inputs with shared sub-expressions): cutting at
    the widest sum and summing the parts in the tail reproduces the plaintext semantics, for 2..5 parts"""
    from eva_b200 import EvaProgram, Input, Output
    rng = np.random.default_rng(1000 + seed)
    vec = 16
    prog = EvaProgram("rand%d" % seed, vec_size=vec)
    with prog:
        ins = [Input("x"), Input("y")]
        pool = list(ins)
        for _ in range(int(rng.integers(3, 9))):       # shared sub-expressions
            a, b = pool[int(rng.integers(len(pool)))], pool[int(rng.integers(len(pool)))]
            kind = int(rng.integers(5))
            pool.append(a << int(rng.integers(1, vec)) if kind == 0 else a * float(rng.uniform(-2, 2)) if kind == 1 else
                        a * b if kind == 2 else a - b if kind == 3 else -a)
        leaves = []
        for _ in range(int(rng.integers(4, 14))):
            a = pool[int(rng.integers(len(pool)))]
            leaves.append(a * pool[int(rng.integers(len(pool)))] if rng.random() < 0.5 else a << int(rng.integers(vec)))
        total = leaves[0]
        for leaf in leaves[1:]:
            total = total + leaf
        Output("out", total * 0.5 + ins[0])
        if rng.random() < 0.5:
            Output("aux", pool[-1])
    x = {"x": list(rng.uniform(-1, 1, vec)), "y": list(rng.uniform(-1, 1, vec))}
    want = evaluate(prog, x)
    for parts in (2, 3, 5):
        plan = shard.split_program(prog, parts)
        if plan is None:
            continue
        got = _plain_sharded(prog, plan, x)
        assert set(got) == set(want)
        for k in want:
            assert np.allclose(got[k], want[k], rtol=1e-9, atol=1e-9), (seed, parts, k)
