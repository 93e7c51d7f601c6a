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


def reduce_scatter_bucket_(flat, lo: int, hi: int, rank: int, world: int, group=None):
    """sum-reduce-scatter of flat[lo:hi] in place: afterwards this rank's 1/world slice of the bucket holds the sum over ranks (the
    other slices are scratch).  NCCL: one reduce_scatter_tensor whose output aliases the rank's slice of the input (in-place form);
    gloo (CPU tests) has no reduce-scatter: an all-reduce of the bucket gives the same slice.  Returns the async work handle."""
    import torch.distributed as dist
    n = (hi - lo) // world
    if dist.get_backend(group) == "gloo":
        return dist.all_reduce(flat[lo:hi], group=group, async_op=True)
    return dist.reduce_scatter_tensor(flat[lo + rank * n:lo + (rank + 1) * n], flat[lo:hi], group=group, async_op=True)


def all_gather_bucket_(flat, lo: int, hi: int, rank: int, world: int, group=None):
    """all-gather of the ranks' 1/world slices of flat[lo:hi] in place (every rank ends with the whole bucket)"""
    import torch.distributed as dist
    n = (hi - lo) // world
    if dist.get_backend(group) == "gloo":
        return dist.all_gather([flat[lo + r * n:lo + (r + 1) * n] for r in range(world)], flat[lo + rank * n:lo + (rank + 1) * n].clone(),
                               group=group, async_op=True)
    return dist.all_gather_into_tensor(flat[lo:hi], flat[lo + rank * n:lo + (rank + 1) * n], group=group, async_op=True)


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the samples rank `rank` handles when n samples are split as evenly as possible"""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_sizes(n: int, world: int) -> List[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def ops_overlapping_exchange(op_ms, schedule, ms_per_byte, slowdown, launch_ms=0.03, margin_ms=0.05):
    """Which backward ops run while the gradient exchange is on the wire (host-side replay used by
    Engine.reserve_sms_for_exchange).  op_ms[i]: device time of backward op i; schedule: [(n_ops, lo, hi)] from
    Planner.exchange_schedule (bucket [lo, hi) of fp32 gradients may start once the first n_ops ops have run);
    ms_per_byte: measured all-reduce speed; slowdown: factor by which an op that leaves SMs to the collective gets slower.
    Bucket j starts when its last producer op has finished and the wire is free; an op is reserved if its interval
    touches any bucket's interval (+- margin).  Reserved ops run slower, which moves the windows: iterate to a fixed point.
    Returns (sorted op indices, [(start_ms, end_ms)] per bucket)."""
    reserved, windows = set(), []
    for _ in range(4):
        end, t = [], 0.0
        for i, d in enumerate(op_ms):
            t += d * (slowdown if i in reserved else 1.0)
            end.append(t)
        busy, windows = 0.0, []
        for (n_ops, lo, hi) in schedule:
            start = max(end[n_ops - 1] if n_ops > 0 else 0.0, busy)
            busy = start + launch_ms + (hi - lo) * 4 * ms_per_byte
            windows.append((start, busy))
        new = set()
        for i in range(len(op_ms)):
            s_i, e_i = (end[i - 1] if i else 0.0), end[i]
            if any(s_i < w1 + margin_ms and e_i > w0 - margin_ms for (w0, w1) in windows):
                new.add(i)
        if new == reserved:
            break
        reserved = new
    return sorted(reserved), windows
