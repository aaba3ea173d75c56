"""Frame-batch data parallelism: one process per GPU, torch.distributed (NCCL over NVLink) as plumbing.

The stage-1 path shards on frames: every op is per-frame except the optimiser step, so
  * training needs exactly one sum all-reduce per optimiser per step, on the flat gradient buffer of that
    optimiser's ParamGroup (G: 6.5 M floats, D: 44.7 M floats); the 1/world scaling is folded into the Adam kernel;
    batch-norm statistics stay per replica (the reference normalises per call at the same per-replica batch);
  * pseudo-label / evaluation passes shard the frame list contiguously and use NO collective.
"""
import torch
import torch.distributed as dist


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank(group=None):
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def allreduce_sum_(buf, group=None):
    """In-place sum all-reduce of a flat buffer (no-op for a single process)."""
    if world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def shard_range(n_items, rank_, world):
    """Contiguous shard [lo, hi) of n_items for rank_ of world (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank_ * base + min(rank_, rem)
    return lo, lo + base + (1 if rank_ < rem else 0)


def broadcast_parameters(ctx, src=0, group=None):
    """Make every replica start from rank `src`'s parameters, moving statistics and constants."""
    if world_size(group) > 1:
        for g in (ctx.G, ctx.D, ctx.S, ctx.V):
            if g.data is not None and g.data.numel():
                dist.broadcast(g.data, src=src, group=group)
        ctx.params_changed()
