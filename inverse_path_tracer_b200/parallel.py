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


def scenes_for_rank(n_scenes, rank, world):
    """BASELINE config 4: scene i is optimised on GPU i mod world"""
    return list(range(rank, n_scenes, world))


def gather_rows(local, height, group=None):
    """all-gather row blocks of unequal size into the full (height, W, 3) image on every rank"""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    blocks = []
    for r in range(world):
        b, e = shard_rows(height, r, world)
        blocks.append(torch.empty((e - b,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device))
    dist.all_gather(blocks, local.contiguous(), group=group)
    return torch.cat(blocks, 0)


def render_sharded(render_rows, height, group=None):
    """render_rows(begin, end) -> tensor of this rank's rows; returns the gathered full image"""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b, e = shard_rows(height, rank, world)
    return gather_rows(render_rows(b, e), height, group)


def allreduce_sum(t, group=None):
    if dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
