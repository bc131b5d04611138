"""Multi-GPU partitioning of the scoring path.

Three nested levels, all independent except one sum:
  regions    -> round-robin over ranks, no communication (`assign_regions`);
  events     -> one region's events split into contiguous blocks, one per rank; every rank scores
                all mutations against its block and the per-mutation partial sums are combined
                with ONE all_reduce(sum, float64, n_mutations) (`score_mutations_event_sharded`);
  mutations  -> not needed once events are split.

On the GPUs the combine step lives INSIDE the library (`score_mutations_sharded` -> ps_score_mutations_sharded,
csrc/ps_comm.cu: NCCL on the library's stream, ordered = bit-identical to one GPU, or one all-reduce).  The
`torch.distributed` form below (`score_mutations_event_sharded`) is the same host logic over any backend: gloo in the
CPU test.  The sum over
a rank's own events stays the ordered FP64 sum of the kernel; across ranks the order of the (few)
partials is fixed by the reduction, so scores agree with the single-GPU path to ~1e-16 relative
(not bit for bit) while accept/reject decisions on clearly signed scores are unchanged.
"""
import numpy as np


def assign_regions(n_regions, rank, world):
    """Indices of the regions rank `rank` of `world` processes (poreseq's own scaling model:
    one worker per region file, README.md:48-54)."""
    return list(range(rank, n_regions, world))


def event_block(n_events, rank, world):
    """Contiguous [lo, hi) block of a region's events for this rank (keeps event order inside the shard)."""
    per = (n_events + world - 1) // world
    lo = min(rank * per, n_events)
    return lo, min(lo + per, n_events)


class RegionShard(object):
    """A PSAlign-like view of one region holding only this rank's block of events."""

    def __init__(self, region, rank, world):
        lo, hi = event_block(len(region.events), rank, world)
        self.sequence = region.sequence
        self.events = region.events[lo:hi]
        self.params = region.params
        self.block = (lo, hi)


def score_mutations_event_sharded(region, starts, origs, muts, partial_fn, rank, world, device="cpu", group=None):
    """Scores `muts` against a region whose events are split over `world` ranks.

    partial_fn(shard, starts, origs, muts) -> float64 array of per-mutation sums over the shard's
    events starting at 0 (on the product path: NativeRegion.score_mutations_partial).
    Returns the full scores (-1e-6 + sum over all events), identical on every rank."""
    import torch
    import torch.distributed as dist
    shard = RegionShard(region, rank, world)
    local = np.zeros(len(starts))
    if len(shard.events):
        local = np.asarray(partial_fn(shard, starts, origs, muts), dtype="f8")
    t = torch.from_numpy(local.copy()).to(device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return -1e-6 + t.cpu().numpy()


def score_mutations_sharded(ctx, region, starts, origs, muts, rank, world, width_key=None):
    """The product path: this rank's block of the region's events goes into a native region on this rank's GPU and
    ps_score_mutations_sharded combines the per-mutation sums over NCCL inside the library (ctx.comm_init first).
    Returns the complete scores (identical on every rank)."""
    from . import poreseqcpp
    shard = RegionShard(region, rank, world)
    nr = poreseqcpp.NativeRegion(ctx, shard.sequence, shard.events, shard.params, width_key)
    try:
        return nr.score_mutations_sharded(starts, origs, muts)
    finally:
        nr.close()


def cuda_partial(ctx, width_key=None):
    """partial_fn for the product path: the shard's events go through the C-ABI on this rank's GPU."""
    from . import poreseqcpp

    def fn(shard, starts, origs, muts):
        nr = poreseqcpp.NativeRegion(ctx, shard.sequence, shard.events, shard.params, width_key)
        try:
            return nr.score_mutations_partial(starts, origs, muts)
        finally:
            nr.close()
    return fn
