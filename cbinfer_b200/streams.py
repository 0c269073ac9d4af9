"""Multi-GPU partitioning of the path: independent video streams sharded over ranks.

Every CB layer's state is per stream and the weights are read-only, so the path shards by stream
with no data-path collective.  One process per GPU; torch.distributed is used
for the rendezvous, the barrier and the max-over-ranks timing only.
"""
import torch
import torch.distributed as dist


def shard_streams(n_streams, world_size, rank, policy="block"):
    """Global stream ids owned by `rank`.  'block': contiguous ranges (rank r gets
    [r*n/w, (r+1)*n/w)), 'cyclic': stream i -> rank i % world_size."""
    if policy == "cyclic":
        return list(range(rank, n_streams, world_size))
    lo = (n_streams * rank) // world_size
    hi = (n_streams * (rank + 1)) // world_size
    return list(range(lo, hi))


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (identity without a process group)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def whole_job_rate(units_this_rank, elapsed_ms_this_rank, device=None):
    """Whole-job throughput: all ranks' units / slowest rank's time (units per second)."""
    total = sum_over_ranks(units_this_rank, device)
    worst = max_over_ranks(elapsed_ms_this_rank, device)
    return total / (worst * 1e-3), worst
