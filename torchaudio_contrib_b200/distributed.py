"""Multi-GPU plumbing for the batch-sharded path (BASELINE config 4): one process per GPU, the batch
split contiguously across ranks, no data-path collective; an optional all-gather of the output.

Every (batch, channel) sequence is independent (reference functional.py:89-91 flattens all leading
dims into one batch), so rank r of W simply owns a contiguous range of batch items.  The only
collective is the optional `all_gather_output`, NCCL over NVLink on GPUs (gloo in the CPU tests).
"""
import torch
import torch.distributed as dist

__all__ = ["shard_range", "shard_batch", "all_gather_output"]


def shard_range(n_items, rank, world):
    """Contiguous [lo, hi) of `n_items` owned by `rank`; the first `n_items % world` ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    base, extra = divmod(int(n_items), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x, rank=None, world=None):
    """This rank's slice of a batch-first tensor (a view, no copy)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(x.size(0), rank, world)
    return x[lo:hi]


def all_gather_output(local_out, n_items, group=None):
    """Gather per-rank outputs (batch-first, sharded by `shard_range(n_items, ...)`) into the full
    `(n_items, ...)` tensor on every rank.  Equal shards use one `all_gather_into_tensor`; ragged
    shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_out
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [hi - lo for lo, hi in (shard_range(n_items, r, world) for r in range(world))]
    if local_out.size(0) != sizes[rank]:
        raise ValueError("rank %d holds %d items, expected %d" % (rank, local_out.size(0), sizes[rank]))
    local_out = local_out.contiguous()
    tail = tuple(local_out.shape[1:])
    if len(set(sizes)) == 1:
        full = torch.empty((n_items,) + tail, dtype=local_out.dtype, device=local_out.device)
        dist.all_gather_into_tensor(full, local_out, group=group)
        return full
    biggest = max(sizes)
    padded = torch.zeros((biggest,) + tail, dtype=local_out.dtype, device=local_out.device)
    padded[:sizes[rank]] = local_out
    buf = torch.empty((world * biggest,) + tail, dtype=local_out.dtype, device=local_out.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * biggest:r * biggest + sizes[r]] for r in range(world)], dim=0)
