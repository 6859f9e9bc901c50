"""Chain sharding + the only exchange step of the hot path (torch.distributed plumbing).

Chains are independent (exe_flow_matching.py:303,312-313); the FM loss is a SUM over chains (:178),
so ranks all-reduce the flat gradient with SUM.  Works with NCCL (GPU) and gloo (CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as tdist


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of rank; the first n_total % world ranks hold one extra chain."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def world_info(group=None):
    if tdist.is_available() and tdist.is_initialized():
        return tdist.get_rank(group), tdist.get_world_size(group)
    return 0, 1


def allreduce_sum_(tensors, group=None):
    """In-place SUM all-reduce of the FM gradient buffer(s) and loss scalar."""
    _, world = world_info(group)
    if world > 1:
        for t in tensors:
            tdist.all_reduce(t, op=tdist.ReduceOp.SUM, group=group)
    return tensors


def allgather_chains(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Gather per-chain values (e.g. log-likelihoods for the tempering ESS) from every shard, in
    global chain order.  Shards may differ in size by one."""
    rank, world = world_info(group)
    if world == 1:
        return local
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    tdist.all_gather(out, buf, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)
