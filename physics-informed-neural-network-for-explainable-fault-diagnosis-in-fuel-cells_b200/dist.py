"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the
box, gloo in CPU tests).  The path shards by SAMPLES:

* training: each rank owns a contiguous row range, kernels return un-normalised sums, one
  ``all_reduce`` of the flat gradient bucket per step (driven from ``PhysicsInformedNN``);
* MC sweep: sample-sharded needs no collective (outputs are gathered); pass-sharded sweeps
  merge per-rank Welford partials ``(count, mean, M2, sum logvar)`` with Chan's update.

Everything here is tensor-type agnostic host logic (works on CPU tensors under gloo).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced row range ``[lo, hi)`` of rank ``rank`` (first ``n % world`` ranks get one extra)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def chan_merge(count_a, mean_a, m2_a, slv_a, count_b, mean_b, m2_b, slv_b):
    """Merge two Welford partials over disjoint pass sets (Chan, Golub & LeVeque 1979)."""
    if count_b == 0:
        return count_a, mean_a, m2_a, slv_a
    if count_a == 0:
        return count_b, mean_b, m2_b, slv_b
    n = count_a + count_b
    d = mean_b - mean_a
    mean = mean_a + d * (count_b / n)
    m2 = m2_a + m2_b + d * d * (count_a * count_b / n)
    return n, mean, m2, slv_a + slv_b


def finalize(count, m2, slv):
    """``a_u = sqrt(exp(mean_t logvar))``, ``e_u = sqrt(var_t u)`` with ddof 0 (01:1483-1486)."""
    return torch.sqrt(torch.exp(slv / count)), torch.sqrt(torch.clamp(m2, min=0) / count)


def merge_pass_shards(count, mean, m2, slv, group=None):
    """All-gather every rank's raw partial and fold them in rank order (deterministic)."""
    world = dist.get_world_size(group)
    packed = torch.stack([mean, m2, slv])
    bufs = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(bufs, packed, group=group)
    cnt = torch.tensor([count], dtype=torch.int64, device=mean.device)
    cnts = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    acc = (0, None, None, None)
    for c, b in zip(cnts, bufs):
        acc = chan_merge(acc[0], acc[1], acc[2], acc[3], int(c.item()), b[0], b[1], b[2])
    return acc


def gather_rows(local: torch.Tensor, n_total: int, group=None):
    """Concatenate sample-sharded per-row outputs (ragged shards allowed) on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    assert sizes[rank][1] - sizes[rank][0] == local.shape[0]
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


def allreduce_bucket(bucket: torch.Tensor, group=None):
    """The one collective of a data-parallel train step: sum the flat fp32 gradient bucket."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(bucket, group=group)
    return bucket
