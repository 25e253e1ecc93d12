"""Multi-GPU plumbing of the replica-sharded hot path (one process per GPU).

The path shards by independent program instances: every rank owns
its own instances end to end, there is no data-path collective.  The only
exchange is the final gather of the outputs on rank 0 and the max-over-ranks of
the timed region.  Works over NCCL (device tensors) and gloo (host tensors; the
`-m "not gpu"` tests run it with world_size 2).
"""
import numpy as np


def instance_seed(rank, index, base=77):
    """seed of synthetic instance `index` of rank `rank`: distinct for every (rank, index)"""
    return base * (rank + 1) * 1000003 + index


def gather_outputs(out_host, rank, world, device=None):
    """final gather of one output ciphertext per rank on rank 0.
    out_host: uint64 numpy array (same shape on every rank).  Returns the list of
    world arrays on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [out_host]
    # ranks may hold a different number of polynomials (a partial sum of relinearized leaves has 2, one with a raw
    # product 3): agree on the largest leading dimension, pad with zero polynomials, trim after the gather
    lead = torch.tensor([out_host.shape[0]], dtype=torch.int64, device=device if device is not None else "cpu")
    leads = [torch.empty_like(lead) for _ in range(world)]
    dist.all_gather(leads, lead)
    leads = [int(x.item()) for x in leads]
    padded = np.zeros((max(leads),) + tuple(out_host.shape[1:]), dtype=np.uint64)
    padded[: out_host.shape[0]] = out_host
    t = torch.from_numpy(padded.view(np.int64))
    if device is not None:
        t = t.to(device)
    gathered = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, gathered, dst=0)
    if rank != 0:
        return None
    return [np.ascontiguousarray(g.cpu().numpy().view(np.uint64)[: leads[r]]) for r, g in enumerate(gathered)]


def max_over_ranks(times, world, device=None):
    """element-wise max of per-rank timings (seconds): the job takes as long as its slowest rank"""
    import torch
    import torch.distributed as dist
    if world == 1:
        return list(times)
    tt = torch.tensor(list(times), dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return tt.tolist()


def aggregate_ops(ops_per_instance, instances_per_rank, steps, world):
    """whole-job unit count of a weak-scaling run"""
    return ops_per_instance * instances_per_rank * steps * world
