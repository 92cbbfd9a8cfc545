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


class SymmetricBucket:
    """Gradient bucket in torch symmetric memory (NVLink peer-mapped): ``[64 flag words | slot 0 | slot 1]`` per rank.
    ``slot(i)`` is the view K2 writes the local gradients of a step into; ``ptrs`` the device array of every rank's
    buffer address for ``kernels.adam_step_p2p``.  ``SymmetricBucket.create`` returns None when symmetric memory is not
    available (single process, gloo, no peer access): the caller then keeps the NCCL all-reduce."""

    def __init__(self, buf, handle, n, device):
        self.buf, self.handle, self.n = buf, handle, n
        self.ptrs = torch.tensor([int(p) for p in handle.buffer_ptrs], dtype=torch.int64, device=device)
        self.rank, self.world = handle.rank, handle.world_size
        self.tag = 0

    def slot(self, i):
        from .kernels import P2P_FLAG_WORDS
        return self.buf[P2P_FLAG_WORDS + i * self.n: P2P_FLAG_WORDS + (i + 1) * self.n]

    @staticmethod
    def create(n, device, group=None, words=None):
        """``words``: total 32-bit words of the buffer when it is not the ``[64 | n | n]`` layout (``train_dnn_steps_dp``
        uses ``kernels.dp_bucket_words``)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
            return None
        if dist.get_backend(group) != "nccl":
            return None
        try:
            import torch.distributed._symmetric_memory as symm
            from .kernels import P2P_FLAG_WORDS
            buf = symm.empty(int(words) if words is not None else P2P_FLAG_WORDS + 2 * n, dtype=torch.float32, device=device)
            buf.zero_()
            handle = symm.rendezvous(buf, group if group is not None else dist.group.WORLD)
            torch.cuda.synchronize(device)
            handle.barrier()
            ok = torch.ones(1, device=device)
        except Exception:                       # noqa: BLE001 -- any failure means "not available here"
            ok, buf, handle = torch.zeros(1, device=device), None, None
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)      # all ranks take the same path
        if ok.item() < 1:
            return None
        return SymmetricBucket(buf, handle, n, device)
