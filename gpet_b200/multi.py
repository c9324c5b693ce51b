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


# ------------------------------------------------------------------------------------------------ exchange by time slice
# Second way to shard (BASELINE.json north_star): by DECAY INDEX.  Every rank transports its share of the decays
# of the whole acquisition window (disjoint Philox subsequences: gpet_set_first_pair), so its post-readout events are spread
# over the whole window.  The digitizer couples events of all ranks again -- dead time per site, coincidence windows over
# the global time order (the reference digitizes a whole epoch = time slice as ONE list, gPET.cu:385-424) -- so the events
# are exchanged by time slice: rank j receives every event with edges[j] - halo_back <= t < edges[j+1] + halo_fwd, digitizes
# the list with the emit window [edges[j], edges[j+1]) (gpet_set_emit_window) and keeps the singles of its slice and the
# coincidences opened in it.  The union over the ranks equals the digitization of all events as one list, record for record
# (tests/test_multi_gloo.py with the oracle as digitizer on CPU / gloo; tests/test_gpu_parity.py with the CUDA digitizer).
# The only collectives: one all-to-all of the counts, one all-to-all-v of the 48-byte records (NCCL over NVLink on GPUs).
EVENT_BYTES = 48
T_WORD = 3            # the fp64 time is the 4th 8-byte word of a 48-byte record (Event, gPET.h:87-92)
DEAD_T = 1.0e19       # records at t >= MAXT / 10 are dead (constants.h:18)


def slice_edges(t_lo_us, t_hi_us, world):
    """world + 1 edges of equal time slices; the outermost slices are open-ended (flight times push a few events past the window)"""
    e = np.linspace(float(t_lo_us), float(t_hi_us), world + 1)
    e[0], e[-1] = -np.inf, np.inf
    return e


def halo_for(dead_time_us, coinc_window_us, chains=64.0):
    """(halo_back, halo_fwd) in us.  Backwards a dead-time or window chain has to find its certain start (an event no
    predecessor can touch), so the halo is `chains` times the longer of the two; forwards one coincidence window suffices.
    A halo that turns out too short is reported by the digitizer (emit_counts()[2]), never silently wrong."""
    reach = max(float(dead_time_us), float(coinc_window_us))
    return chains * reach + 1e-3, float(coinc_window_us) * 1.0001 + 1e-6


def route_by_time_slice(t, edges, halo_back, halo_fwd):
    """Send lists: for every destination rank the indices (ascending) of the events it needs.  `t`: torch fp64 tensor."""
    import torch
    alive = t < DEAD_T
    out = []
    for j in range(len(edges) - 1):
        lo, hi = float(edges[j]) - halo_back, float(edges[j + 1]) + halo_fwd
        out.append(torch.nonzero(alive & (t >= lo) & (t < hi), as_tuple=False).flatten())
    return out


def exchange_events(events_u8, edges, halo_back, halo_fwd, group=None):
    """events_u8: torch.uint8 tensor [n, 48] of this rank's post-readout events (CUDA tensor under NCCL, CPU under gloo).
    Returns the [m, 48] tensor of the events this rank has to digitize (its slice + halos, from all ranks, in rank order)
    and the bytes this rank sent."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    ev = events_u8.contiguous().view(-1, EVENT_BYTES)
    t = ev.view(torch.float64).view(-1, EVENT_BYTES // 8)[:, T_WORD]
    lists = route_by_time_slice(t, edges, halo_back, halo_fwd)
    if world == 1:
        return ev[lists[0]], 0
    send_counts = torch.tensor([int(ix.numel()) for ix in lists], dtype=torch.int64, device=ev.device)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    send = ev[torch.cat(lists)] if ev.shape[0] else ev
    rc, sc = [int(x) for x in recv_counts.tolist()], [int(x) for x in send_counts.tolist()]
    recv = torch.empty((sum(rc), EVENT_BYTES), dtype=torch.uint8, device=ev.device)
    dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=sc, group=group)
    return recv, int(sum(sc)) * EVENT_BYTES


def digitize_slice(ctx, recv_u8, edges, rank, halo_back):
    """The CUDA digitizer over one exchanged list (a CUDA uint8 tensor [m, 48]): returns (singles of the slice as a numpy
    record array, coincidences opened in the slice, their classes, halo-too-short flag)."""
    lo, hi = float(edges[rank]), float(edges[rank + 1])
    ctx.set_emit_window(lo, hi, float("-inf") if rank == 0 else lo - halo_back)
    try:
        ctx.put_events_device(recv_u8.data_ptr() if recv_u8.numel() else 0, recv_u8.shape[0])
        ctx.stage_digitize()
        before, inside, flag, _ = ctx.emit_counts()
        singles = ctx.fetch_singles()[before:before + inside]
        co = ctx.fetch_coincidences()
    finally:
        ctx.clear_emit_window()
    return singles, co, flag


def pair_block(npairs, rank, world):
    """Contiguous block of the pairs 0..npairs-1 that rank `rank` transports (decay-index sharding)."""
    lo = npairs * rank // world
    return lo, npairs * (rank + 1) // world - lo


def exchange_local(per_rank_events_u8, edges, halo_back, halo_fwd):
    """The same routing without a process group: `per_rank_events_u8[r]` are rank r's events; returns what each rank would
    receive (concatenated in source-rank order, as all_to_all_single delivers it).  Used by the single-process tests."""
    import torch
    world, ndest = len(per_rank_events_u8), len(edges) - 1
    parts = [[None] * world for _ in range(ndest)]
    for r, ev in enumerate(per_rank_events_u8):
        ev = ev.contiguous().view(-1, EVENT_BYTES)
        t = ev.view(torch.float64).view(-1, EVENT_BYTES // 8)[:, T_WORD]
        for j, ix in enumerate(route_by_time_slice(t, edges, halo_back, halo_fwd)):
            parts[j][r] = ev[ix]
    return [torch.cat(p) for p in parts]


def run_exchange(ctx, atoms, device, group=None, frame_pairs=0):
    """ONE acquisition sharded by decays over the ranks, digitized by time slice (the north_star data path).

    `atoms`: the acquisition's atoms per source.  Rank r transports the decays of atoms // world of them (independent
    binomial thinning of disjoint atom sets = the decays of the whole source, shared out; disjoint Philox subsequences
    through gpet_set_first_pair) over the WHOLE time window, frame by frame; per frame the post-readout events are
    exchanged by time slice (all-to-all-v over NCCL, device to device) and every rank digitizes its slice of the frame.
    Returns a dict of this rank's tallies (pairs, events sent / received, singles, coincidences, halo flag, bytes sent)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    for i, n in enumerate(atoms):
        ctx.set_source_atoms(i, int(n) // world)      # equal shares: every rank plans the very same time slices
    ctx.set_first_pair(rank << 40)
    ctx.set_shard(0, 1)
    # the frame plan must be the same time slices on every rank: plan with the capacity a rank's share needs
    nframes = ctx.plan_frames(frame_pairs)
    dig = ctx.get_digitizer()
    hb, hf = halo_for(dig.dead_time_us, dig.coinc_window_us)
    out = dict(pairs=0, events=0, received=0, singles=0, coincidences=0, halo_flag=0, bytes_sent=0, frames=nframes)
    buf = torch.empty((1 << 22, EVENT_BYTES), dtype=torch.uint8, device=device)
    for f in range(nframes):
        fr = ctx.frame(f)
        ctx.stage_front(f)
        ctx.stage_panel_transport()
        n = ctx.copy_events_to_device(buf.data_ptr(), buf.shape[0])
        edges = slice_edges(fr["t0_s"] * 1e6, (fr["t0_s"] + fr["dt_s"]) * 1e6, world)
        recv, sent = exchange_events(buf[:n], edges, hb, hf, group=group)
        lo, hi = float(edges[rank]), float(edges[rank + 1])
        ctx.set_emit_window(lo, hi, float("-inf") if rank == 0 else lo - hb)
        ctx.put_events_device(recv.data_ptr() if recv.numel() else 0, recv.shape[0])
        ctx.stage_digitize()
        before, inside, flag, nco = ctx.emit_counts()
        out["pairs"] += ctx.frame_pairs(f); out["events"] += int(n); out["received"] += int(recv.shape[0])
        out["singles"] += inside; out["coincidences"] += nco
        out["halo_flag"] |= flag; out["bytes_sent"] += sent
    ctx.clear_emit_window()
    ctx.set_first_pair(0)
    return out
