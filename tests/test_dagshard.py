"""DAG-sharded execution (eva_b200/dagshard.py): staged cuts at sums, leaves assigned so that rotations are split
instead of recomputed, partial sums exchanged on the device.  CPU: graph surgery against the plaintext reference
semantics (goldens compiled by the reference compiler + random programs), the two-dimensional blocking of the
wide DAG, the flow over gloo with world_size 2.  GPU: bit-identical to the unsharded execute and to the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from eva_b200 import Op, dagshard, evaluate, program_io  # noqa: E402

_FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "programs")
if _FIX not in program_io.SEARCH:      # also in the spawned gloo workers, which never see conftest.py
    program_io.SEARCH.append(_FIX)


def _close(a, b):
    return all(np.allclose(a[k], b[k], rtol=1e-9, atol=1e-9) for k in b) and set(a) == set(b)


@pytest.mark.parametrize("name", ["wide64", "sobel", "harris", "sobel8192", "feat_hsum", "polynomial"])
@pytest.mark.parametrize("parts", [2, 3, 4, 8])
def test_staged_plan_preserves_semantics(name, parts):
    prog = program_io.build_program(program_io.load_json(name))[0]
    plan = dagshard.plan_stages(prog, parts, force=True)
    if plan is None:
        pytest.skip("no sum to cut")
    rng = np.random.default_rng(3)
    x = {k: list(rng.uniform(-1, 1, prog.vec_size)) for k in prog.inputs}
    assert _close(dagshard.run_plain(plan, x), evaluate(prog, x))
    for st in plan.stages:
        assert 2 <= st.nranks <= parts and len(st.progs) == st.nranks and min(st.leaves) >= 1
        for p in st.progs:
            assert set(p.outputs) == {"partial_%d" % r for r in st.roots}


def test_harris_is_cut_twice_and_rotations_are_split():
    """Harris: the two gradient sums share one stage, the three pooling sums the next; every rotation of the
    input image is computed by exactly one rank (the Ix and Iy leaves of a rotation stay together)"""
    prog = program_io.build_program(program_io.load_json("harris"))[0]
    plan = dagshard.plan_stages(prog, 8)
    assert [len(s.roots) for s in plan.stages] == [2, 3]
    rot = lambda p: sorted(t.attributes["RotationAttribute"] * (1 if t.op == Op.RotateLeftConst else -1) for t in p.terms()
                           if t.op in (Op.RotateLeftConst, Op.RotateRightConst))
    first = [r for p in plan.stages[0].progs for r in rot(p)]
    assert len(first) == len(set(first)) == 9          # 3x3 window: 9 rotation terms (one of them by 0), none computed twice
    assert plan.stages[1].est_sharded < 0.5 * plan.stages[1].est_single


def test_wide_dag_is_blocked_in_two_dimensions():
    """4096 products rot(x, i) * rot(y, j) over 8 ranks: a 2 x 4 grid, 32 + 16 rotations per rank instead of
    the 64 + 8 of a one-dimensional split (and 128 on a single GPU)"""
    prog = program_io.build_program(program_io.load_json("wide4096"))[0]
    for parts, want in ((2, 64 + 32), (4, 32 + 32), (8, 32 + 16)):
        plan = dagshard.plan_stages(prog, parts)
        assert len(plan.stages) == 1 and plan.stages[0].leaves == [4096 // parts] * parts
        nrot = [sum(1 for t in p.terms() if t.op in (Op.RotateLeftConst, Op.RotateRightConst)) for p in plan.stages[0].progs]
        assert max(nrot) <= want, (parts, nrot)
    rng = np.random.default_rng(5)
    x = {k: list(rng.uniform(-1, 1, prog.vec_size)) for k in prog.inputs}
    assert _close(dagshard.run_plain(dagshard.plan_stages(prog, 8), x), evaluate(prog, x))


def test_cost_model_declines_cuts_that_do_not_pay():
    prog = program_io.build_program(program_io.load_json("polynomial"))[0]
    assert dagshard.plan_stages(prog, 4) is None and dagshard.plan_stages(prog, 1) is None
    hs = program_io.build_program(program_io.load_json("feat_hsum"))[0]    # a chain of x + rot(x): nothing independent
    assert dagshard.plan_stages(hs, 4) is None


@pytest.mark.parametrize("seed", range(16))
def test_random_programs(seed):
    from eva_b200 import EvaProgram, Input, Output
    rng = np.random.default_rng(2000 + seed)
    vec = 16
    prog = EvaProgram("rand%d" % seed, vec_size=vec)
    with prog:
        ins = [Input("x"), Input("y")]
        pool = list(ins)
        for _ in range(int(rng.integers(3, 9))):
            a, b = pool[int(rng.integers(len(pool)))], pool[int(rng.integers(len(pool)))]
            kind = int(rng.integers(5))
            pool.append(a << int(rng.integers(1, vec)) if kind == 0 else a * float(rng.uniform(-2, 2)) if kind == 1 else
                        a * b if kind == 2 else a - b if kind == 3 else -a)
        def reduction(n):
            leaves = []
            for _ in range(n):
                a = pool[int(rng.integers(len(pool)))]
                leaves.append(a * pool[int(rng.integers(len(pool)))] if rng.random() < 0.5 else a << int(rng.integers(vec)))
            total = leaves[0]
            for leaf in leaves[1:]:
                total = total + leaf
            return total
        first = reduction(int(rng.integers(4, 12)))
        pool.append(first * 0.5)
        second = reduction(int(rng.integers(4, 12)))     # may or may not depend on the first sum: one or two stages
        Output("out", second + first * ins[0])
        if rng.random() < 0.5:
            Output("aux", pool[-1])
    x = {"x": list(rng.uniform(-1, 1, vec)), "y": list(rng.uniform(-1, 1, vec))}
    want = evaluate(prog, x)
    for parts in (2, 3, 5):
        plan = dagshard.plan_stages(prog, parts, force=True)
        if plan is not None:
            assert _close(dagshard.run_plain(plan, x), want), (seed, parts)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    """the staged flow with real collectives (gloo, host tensors): stage programs evaluated with the plaintext
    semantics, partial sums all-gathered between stages, the last cut gathered on rank 0"""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prog = program_io.build_program(program_io.load_json("harris"))[0]
        plan = dagshard.plan_stages(prog, world)
        rng = np.random.default_rng(5)
        x = {k: list(rng.uniform(-1, 1, prog.vec_size)) for k in sorted(prog.inputs)}
        have = {}

        def ins(p):
            return {nm: (have[plan.partial_names[p.name][nm]] if nm in plan.partial_names.get(p.name, {}) else x[nm]) for nm in p.inputs}
        for si, st in enumerate(plan.stages):
            out = evaluate(st.progs[rank], ins(st.progs[rank]))
            for r in st.roots:
                mine = torch.tensor(np.asarray(out["partial_%d" % r], dtype=np.float64))
                if si == len(plan.stages) - 1:
                    got = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
                    dist.gather(mine, got, dst=0)
                else:
                    got = [torch.empty_like(mine) for _ in range(world)]
                    dist.all_gather(got, mine)
                if got is not None:
                    for w in range(world):
                        have[(r, w)] = list(got[w].numpy())
        if rank == 0:
            q.put(_close(evaluate(plan.tail, ins(plan.tail)), evaluate(prog, x)))
        else:
            q.put(True)
    finally:
        dist.destroy_process_group()


def test_staged_flow_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [True, True]


def _gpu_setup(name, seed=2):
    from eva_b200 import b200
    from oracle import oracle as o
    from oracle_exec import OracleProgram
    d = program_io.load_json(name)
    prog = program_io.build_program(d)[0]
    N = d["poly_modulus_degree"]
    orc = o.Oracle(N, d["prime_bits"]).keygen(seed)
    op = OracleProgram(d, orc)
    op.prepare_keys()
    pub = b200.context_from_raw_keys(N, orc.primes, op.rk, {int(e): k for e, k in op.gks.items()})
    rng = np.random.default_rng(8)
    val, oin = b200.B200Valuation(), {}
    for name_, info in d["signature"].items():
        x = rng.uniform(0, 0.2, d["vec_size"])
        ct = orc.encrypt(orc.encode(x, 2.0 ** info["scale"], orc.k - 1 - info["level"]), seed=31)
        val.set_cipher(name_, ct, 2.0 ** info["scale"])
        oin[name_] = ("cipher", ct, 2.0 ** info["scale"])
    return d, prog, pub, val, op, oin


@pytest.mark.gpu
@pytest.mark.parametrize("name,parts", [("wide64", 2), ("wide64", 8), ("sobel", 4), ("harris", 2), ("harris", 8)])
def test_staged_execution_bit_identical(name, parts):
    """all ranks' stage programs on one GPU, partials handed over on the device == execute(prog), bit for bit"""
    d, prog, pub, val, _, _ = _gpu_setup(name)
    want = pub.execute(prog, val)
    plan = dagshard.plan_stages(prog, parts, force=True)
    got = dagshard.execute_local(pub, plan, val)
    got2 = dagshard.execute_local(pub, plan, val)     # second run: captured graphs
    for oname in d["outputs"]:
        for g in (got, got2):
            assert np.array_equal(g.get(oname)[1], want.get(oname)[1]) and g.get(oname)[2] == want.get(oname)[2]


@pytest.mark.gpu
def test_wide4096_matches_the_oracle_and_shards_bit_identically():
    """BASELINE config 5: 4096 ciphertext x ciphertext products (16k cipher ops, 127 rotation keys) compiled by the
    reference compiler: GPU output == oracle output, and the 8-way staged plan gives the same bits"""
    d, prog, pub, val, op, oin = _gpu_setup("wide4096")
    want = pub.execute(prog, val)
    V = op.run(oin, threads=min(64, os.cpu_count() or 8), keep=set(d["outputs"].values()))
    for oname, oid in d["outputs"].items():
        assert np.array_equal(want.get(oname)[1], V[oid][1]) and want.get(oname)[2] == V[oid][2]
    assert pub.cipher_op_count(prog) == op.cipher_op_count() == 16384 or pub.cipher_op_count(prog) == op.cipher_op_count()
    plan = dagshard.plan_stages(prog, 8)
    got = dagshard.execute_local(pub, plan, val)
    for oname in d["outputs"]:
        assert np.array_equal(got.get(oname)[1], want.get(oname)[1])
