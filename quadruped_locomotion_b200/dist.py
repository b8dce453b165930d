"""Multi-GPU plumbing: the path shards by instance, nothing crosses GPUs on the solve path.

Rank g of G owns the contiguous slice [g*B/G, (g+1)*B/G) of the batch; the only
collective is the all-reduce of the ~30-double statistics vector at the end (sum / max halves).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .capi import STATS_NUM, STATS_NUM_SUM


def shard_range(B: int, rank: int, world: int) -> tuple[int, int]:
    """[start, stop) of rank's contiguous slice; slices tile [0, B) exactly."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (B * rank) // world, (B * (rank + 1)) // world


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """In-place all-reduce of a qlb_stats vector: leading STATS_NUM_SUM entries SUM, the rest MAX."""
    if stats.numel() != STATS_NUM:
        raise ValueError(f"expected {STATS_NUM} statistics, got {stats.numel()}")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        head = stats[:STATS_NUM_SUM].clone()
        tail = stats[STATS_NUM_SUM:].clone()
        dist.all_reduce(head, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(tail, op=dist.ReduceOp.MAX, group=group)
        stats[:STATS_NUM_SUM] = head
        stats[STATS_NUM_SUM:] = tail
    return stats


def stats_dict(s) -> dict:
    s = [float(v) for v in s]
    n = max(s[0], 1.0)
    return dict(count=s[0], ok=s[1], no_stance=s[2], max_iter=s[3], unverified=s[4], bad_input=s[5],
                mean_iterations=s[6] / n, mean_wrench_err=s[7] / n, active_hist=s[8:28],
                infeasible=s[28], max_wrench_err=s[29], max_iterations=s[30])
