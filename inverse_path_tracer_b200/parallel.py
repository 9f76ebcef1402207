"""Multi-GPU plumbing: one process per GPU, scene replicated, pixel rows sharded.

The forward path has no data-path collective: every pixel is owned by exactly one rank and the RNG is
keyed by (global pixel, sample, bounce), so the gathered image is bit-identical for any world size.
The only exchange steps are (a) the all-gather of row blocks and (b) the all-reduce of the fused
[loss, dLoss/dKd] buffer (or of the raw edge sums) in the inverse pass.  Both go through
torch.distributed (NCCL on GPUs; gloo in the CPU tests, where the compute backend is injected).
"""
import torch
import torch.distributed as dist


def shard_rows(height, rank, world):
    """contiguous row block [begin, end) of `rank`; the first height % world ranks get one extra row"""
    base, extra = divmod(height, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_rows_interleaved(height, rank, world):
    """(row_begin, row_end, row_step) of `rank`: rows rank, rank + world, ...  Every rank sees the same mix of image
    regions (ceiling, walls, floor have different path lengths), so the shards take the same time; contiguous
    blocks measured 94 % efficiency at 4 GPUs because the slowest block sets the step."""
    return rank, height, world


def gather_rows_interleaved(local, height, group=None):
    """all-gather interleaved shards and put the rows back in image order (height must be a multiple of world)"""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    if height % world:
        raise ValueError("interleaved sharding needs height % world == 0")
    tail = tuple(local.shape[1:])
    gathered = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered.view(-1, *tail), local.contiguous(), group=group)
    return gathered.transpose(0, 1).reshape((height,) + tail)


def scenes_for_rank(n_scenes, rank, world):
    """BASELINE config 4: scene i is optimised on GPU i mod world"""
    return list(range(rank, n_scenes, world))


def gather_rows(local, height, group=None, out=None):
    """all-gather the row blocks into the full (height, W, 3) image on every rank.

    One all_gather_into_tensor call; when height is not a multiple of the world size the blocks are
    padded to the largest block (collectives need equal sizes) and the padding rows are dropped."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    rank = dist.get_rank(group)
    base, extra = divmod(height, world)
    tail = tuple(local.shape[1:])
    if extra == 0:
        if out is None:
            out = torch.empty((height,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    mx = base + 1
    padded = torch.zeros((mx,) + tail, dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    gathered = torch.empty((world * mx,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    parts = []
    for r in range(world):
        b, e = shard_rows(height, r, world)
        parts.append(gathered[r * mx:r * mx + (e - b)])
    assert parts[rank].shape[0] == local.shape[0]
    return torch.cat(parts, 0)


def render_sharded(render_rows, height, group=None):
    """render_rows(begin, end) -> tensor of this rank's rows; returns the gathered full image"""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b, e = shard_rows(height, rank, world)
    return gather_rows(render_rows(b, e), height, group)


def allreduce_sum(t, group=None):
    if dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
