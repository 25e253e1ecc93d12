"""N>1 path of the bench (replica sharding + final gather + max-over-ranks) on CPU:
two processes over gloo (127.0.0.1)."""
import os
import socket

import numpy as np


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from eva_b200 import multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seeds = [multi.instance_seed(rank, i) for i in range(4)]
        rng = np.random.default_rng(seeds[0])
        out = rng.integers(0, 1 << 60, size=(2, 1, 64), dtype=np.uint64)   # this rank's output ciphertext
        gathered = multi.gather_outputs(out, rank, world)
        times = multi.max_over_ranks([0.5 + rank, 2.0 - rank], world)
        q.put((rank, seeds, out, gathered, times))
    finally:
        dist.destroy_process_group()


def test_replica_sharding_world2():
    import torch.multiprocessing as mp
    from eva_b200 import multi
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        r = q.get(timeout=120)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # instances are disjoint across ranks
    assert not set(res[0][1]) & set(res[1][1])
    # rank 0 holds every rank's output, bit-exact; other ranks hold nothing
    assert res[1][3] is None
    for r in range(2):
        assert np.array_equal(res[0][3][r], res[r][2])
    # the job is as slow as its slowest rank, on every rank
    assert res[0][4] == [1.5, 2.0] and res[1][4] == [1.5, 2.0]
    assert multi.aggregate_ops(61, 32, 50, 2) == 61 * 32 * 50 * 2


def test_single_rank_passthrough():
    from eva_b200 import multi
    out = np.arange(8, dtype=np.uint64)
    assert multi.gather_outputs(out, 0, 1)[0] is out
    assert multi.max_over_ranks([1.0, 2.0], 1) == [1.0, 2.0]
