"""Data-parallel sharding of clips / swing pairs over the GPUs of one box.

Every clip and every pair is independent, so each rank runs the hot path on a
contiguous shard with no inter-GPU traffic; the ONLY exchange is the final gather
of per-frame logits (or u8 labels) and of the fixed-size padded alignment paths.
One process per GPU; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is
plumbing.  The reference has no multi-GPU code (collectives "None").
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `n` units for `rank`; the first n % world ranks
    get one extra unit, so shards differ by at most one and concatenate in rank order."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError(f"bad shard request n={n} rank={rank} world={world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n: int, world: int) -> List[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_shards(local, n_total: int, dist=None, group=None):
    """All-gather per-rank result shards (dim 0 = units) into the full result in global
    unit order.  Ragged shards (n % world != 0) are padded to the largest shard for the
    collective and trimmed afterwards.  Without an initialised process group returns `local`."""
    import torch

    if dist is None:
        import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    if local.dtype == torch.int16:
        # the NCCL process group has no 16-bit integer type: an all-gather only moves bytes, so int16 payloads
        # (alignment paths) travel as float16 bit patterns
        return gather_shards(local.view(torch.float16), n_total, dist, group).view(torch.int16)
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_total, world)
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError(f"rank holds {local.shape[0]} units, shard plan says {sizes[dist.get_rank(group)]}")
    biggest = max(sizes)
    if local.shape[0] < biggest:
        pad = torch.zeros((biggest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], 0)
    out = torch.empty((world * biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    if all(s == biggest for s in sizes):
        return out
    return torch.cat([out[r * biggest:r * biggest + sizes[r]] for r in range(world)], 0)


class OverlappedGather:
    """`gather_shards` on a side stream, for a loop of independent batches: the collective of batch n runs under the
    kernels of batch n+1 instead of between them (the gather is ~0.1 ms of launch and synchronisation latency per
    batch whatever its size, 3 % of a 4 ms step on 8 GPUs).  `submit` returns the gathered tensor at once; it is valid
    on the CURRENT stream only after `wait()`.  Without an initialised process group, or for host tensors (the gloo
    tests), it degrades to the synchronous `gather_shards`."""

    def __init__(self, dist=None, group=None):
        self.dist, self.group, self.stream = dist, group, None

    def submit(self, local, n_total: int):
        import torch

        dist = self.dist
        if dist is None:
            import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or not local.is_cuda:
            return gather_shards(local, n_total, dist, self.group)
        if self.stream is None:
            self.stream = torch.cuda.Stream(local.device)
        self.stream.wait_stream(torch.cuda.current_stream(local.device))      # behind the kernels that produced `local`
        with torch.cuda.stream(self.stream):
            out = gather_shards(local, n_total, dist, self.group)
        local.record_stream(self.stream)          # the allocator must not hand `local` out again while the collective reads it
        return out

    def wait(self):
        import torch

        if self.stream is not None:
            torch.cuda.current_stream(self.stream.device).wait_stream(self.stream)


def run_sharded(fn: Callable, inputs: Sequence, n_total: Optional[int] = None, dist=None, group=None):
    """Apply `fn(*shards)` to this rank's contiguous shard of every input (dim 0 = units) and
    gather every output.  `fn` returns a tensor or a tuple of tensors whose dim 0 is the shard."""
    if dist is None:
        import torch.distributed as dist
    n = int(inputs[0].shape[0]) if n_total is None else int(n_total)
    if dist.is_available() and dist.is_initialized():
        lo, hi = shard_range(n, dist.get_rank(group), dist.get_world_size(group))
    else:
        lo, hi = 0, n
    outs = fn(*[x[lo:hi] for x in inputs])
    single = not isinstance(outs, (tuple, list))
    outs = [outs] if single else list(outs)
    gathered = [gather_shards(o, n, dist, group) for o in outs]
    return gathered[0] if single else tuple(gathered)
