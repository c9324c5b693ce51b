"""Multi-GPU driver: one process per GPU (torch.distributed is the plumbing, NCCL on GPUs / gloo in the CPU tests).

Sharding unit = the frame (a contiguous slice of acquisition time, the reference's epoch, gPET.cu:260-427): every frame is
sampled, transported and digitized on one GPU, so the data path needs no collective; rank r owns frames f with
f % world == r (gpet_set_shard).  Philox counters are global photon ids, hence the union of the shards' singles is
byte-identical to a single-GPU run.  Only the tallies (counters, energy spectrum) cross GPUs: one all-reduce."""
from __future__ import annotations

import numpy as np

TALLY_FIELDS = ("pairs", "photons_phantom_out", "photons_on_panel", "hits", "events_adder", "events_threshold",
                "events_deadtime", "singles", "coincidences", "trues", "scatters", "randoms", "overflow_hits", "overflow_events",
                "overflow_adder", "frames")


def owned_frames(nframes: int, rank: int, world: int):
    """Frame indices rank `rank` of `world` runs (same rule as run_impl in csrc/abi.cu)."""
    return [f for f in range(nframes) if f % world == rank]


def stats_vector(st) -> np.ndarray:
    return np.array([int(getattr(st, k)) for k in TALLY_FIELDS], np.int64)


def allreduce_tallies(vec, device=None, group=None):
    """Sum an int64 tally vector (stats_vector, spectrum bins, ...) over all ranks; returns a numpy array."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(vec, np.int64), device=device if device is not None else "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def plan_shard(ctx, max_pairs=0, group=None):
    """Plan the acquisition (identical on every rank: the planner is a pure function of seed and inputs) and return
    (owned frame indices, pairs in the owned frames, pairs in all frames)."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = ctx.plan_frames(max_pairs)
    ctx.set_shard(rank, world)
    mine = owned_frames(n, rank, world)
    pairs = [ctx.frame_pairs(f) for f in range(n)]
    return mine, int(sum(pairs[f] for f in mine)), int(sum(pairs))


def run_sharded(ctx, resident=False, device=None, group=None, spectrum_bins=0):
    """This rank's share of the acquisition + the all-reduced tallies.  Returns (local stats, dict of global totals)."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    ctx.set_shard(rank, world)
    st = ctx.run_resident() if resident else ctx.run(None)
    tot = allreduce_tallies(stats_vector(st), device=device, group=group)
    out = dict(zip(TALLY_FIELDS, (int(x) for x in tot)))
    if spectrum_bins:
        out["spectrum"] = allreduce_tallies(ctx.spectrum(spectrum_bins), device=device, group=group)
    return st, out
