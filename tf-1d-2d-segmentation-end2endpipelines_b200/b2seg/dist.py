"""Data-parallel plumbing: one process per GPU, identical weights, one exchange step per iteration =
sum-all-reduce of the flat gradient arena; Adam then scales by 1/world_size (mean of per-replica mean losses, the
MirroredStrategy convention).  BatchNorm statistics stay per replica like Keras BN.  Inference shards by batch with no
collective.  torch.distributed only moves bytes here; on GPUs the backend is NCCL over NVLink, the CPU tests use gloo."""
from __future__ import annotations

from typing import List, Tuple


def allreduce_flat_(flat, group=None, bucket_elems: int = 0):
    """In-place sum-all-reduce of a flat tensor, optionally in buckets (so later buckets can overlap other work).
    Returns the list of async work handles (empty when the process group is not initialised / world == 1)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return []
    if bucket_elems <= 0 or bucket_elems >= flat.numel():
        return [dist.all_reduce(flat, group=group, async_op=True)]
    works = []
    for s in range(0, flat.numel(), bucket_elems):
        works.append(dist.all_reduce(flat[s:s + bucket_elems], group=group, async_op=True))
    return works


def wait_all(works):
    for w in works:
        w.wait()


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the samples rank `rank` handles when n samples are split as evenly as possible"""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_sizes(n: int, world: int) -> List[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
