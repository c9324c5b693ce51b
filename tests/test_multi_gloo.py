"""World-size-2 tests of the multi-GPU host logic on CPU (gloo): frame sharding and tally reduction.
The compute itself needs a GPU; what is checked here is that every rank plans the same frames, that the shards are a
disjoint cover, and that the all-reduced tallies are the sums -- the logic `bench.py --gpus N` and gpet_run rely on."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
EXAMPLE = ROOT / "examples" / "small_animal"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def stats_ok(tot, world, fields):
    tri = world * (world + 1) // 2
    return (all(tot[name] == tri * (k + 1) for k, name in enumerate(fields))
            and {"trues", "scatters", "randoms", "coincidences", "singles", "pairs"} <= set(fields))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from gpet_b200 import api, multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        with api.Context(-1) as c:                       # host-only context: planning works, compute must refuse
            c.set_seed(1234)
            c.load_isotopes(EXAMPLE / "data" / "isotopes.txt")
            c.load_source(EXAMPLE / "input" / "source.txt")
            c.set_time_window(0, 30)
            mine, my_pairs, all_pairs = multi.plan_shard(c, max_pairs=40000)
            n = c.plan_frames(40000)
            frames = [(c.frame(f)["t0_s"], c.frame(f)["dt_s"], int(c.frame(f)["first_pair"]), c.frame_pairs(f)) for f in range(n)]
            refused = False
            try:
                c.run_resident()
            except api.GpetError as e:
                refused = e.code == -4                   # GPET_ERR_NO_DEVICE: no CPU fallback
        tot = multi.allreduce_tallies([my_pairs, len(mine), 1])
        # the whole tally vector of a run (gpet_stats), coincidence classes included
        st = api.Stats()
        for k, name in enumerate(multi.TALLY_FIELDS):
            setattr(st, name, (rank + 1) * (k + 1))
        stats_tot = dict(zip(multi.TALLY_FIELDS, multi.allreduce_tallies(multi.stats_vector(st)).tolist()))
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine, frames))
        q.put((rank, mine, my_pairs, all_pairs, tot.tolist(), gathered, refused and stats_ok(stats_tot, world, multi.TALLY_FIELDS)))
    finally:
        dist.destroy_process_group()


def test_two_ranks_shard_frames_and_reduce_tallies():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, mine0, pairs0, all0, tot0, gath0, ref0), (r1, mine1, pairs1, all1, tot1, gath1, ref1) = res
    assert ref0 and ref1
    # both ranks planned the very same frames (the planner is a pure function of seed and inputs)
    assert gath0[0][1] == gath0[1][1] and all0 == all1
    nframes = len(gath0[0][1])
    assert nframes >= 5
    # disjoint cover, round robin
    assert sorted(mine0 + mine1) == list(range(nframes)) and not set(mine0) & set(mine1)
    assert mine0 == list(range(0, nframes, 2)) and mine1 == list(range(1, nframes, 2))
    # all-reduced tallies are the sums, identical on both ranks
    assert tot0 == tot1 == [all0, nframes, 2]
    assert pairs0 + pairs1 == all0
    # frames tile the acquisition window and global pair indices are contiguous
    fr = gath0[0][1]
    for a, b in zip(fr[:-1], fr[1:]):
        assert abs(a[0] + a[1] - b[0]) < 1e-9 and a[2] + a[3] == b[2]
    assert abs(fr[-1][0] + fr[-1][1] - 30.0) < 1e-6


def test_owned_frames_rule():
    from gpet_b200 import multi
    for world in (1, 2, 3, 8):
        cover = sorted(f for r in range(world) for f in multi.owned_frames(37, r, world))
        assert cover == list(range(37))


# ------------------------------------------------------------------------------------------------ exchange by time slice
def _slice_of(singles, co, lo, hi):
    """what a rank keeps of the digitization of its list: the singles of its slice, the coincidences opened in it"""
    s = singles[(singles["t"] >= lo) & (singles["t"] < hi)]
    c = co[(co["a"]["t"] >= lo) & (co["a"]["t"] < hi)]
    return s, c


def _exchange_worker(rank, world, port, q, n_events, dead_type, dead_level):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT / "tests"))
    import parity
    from gpet_b200 import api, multi
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the same list on every rank; rank r "transported" the photons with parn % world == r
        rng = np.random.default_rng(99)
        T = 3.0e4
        ev = parity.random_events(n_events, rng, tmax=T, dead_fraction=0.01)
        p, d = parity.make_digi_params(dead_type=dead_type, dead_level=dead_level, dead_time_us=2.2, coinc_window_us=0.5, coinc_policy=1,
                                       blur_Rref=0.05)
        mine = np.ascontiguousarray(ev[ev["parn"] % world == rank])
        edges = multi.slice_edges(0.0, T, world)
        hb, hf = multi.halo_for(d["dead_time_us"], d["coinc_window_us"])
        recv, sent = multi.exchange_events(torch.from_numpy(mine.view(np.uint8).reshape(-1, 48)), edges, hb, hf)
        got = np.ascontiguousarray(recv.numpy()).view(api.EVENT_DTYPE).reshape(-1)
        singles, counts, co = orc.digitize(got, p)
        s, c = _slice_of(singles, co, edges[rank], edges[rank + 1])
        gathered = [None] * world
        dist.all_gather_object(gathered, (s.tobytes(), c.tobytes(), got.size, sent))
        if rank == 0:
            want_s, _, want_c = orc.digitize(ev, p)
            s_all = b"".join(g[0] for g in gathered)
            c_all = b"".join(g[1] for g in gathered)
            received = [g[2] for g in gathered]
            q.put((s_all == want_s.tobytes(), c_all == want_c.tobytes(), want_s.size, want_c.size, received, int((ev["t"] < 1e19).sum()),
                   [g[3] for g in gathered]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dead_type,dead_level", [(0, 3), (1, 2)])
def test_time_slice_exchange_with_halo_reproduces_the_single_list_digitization(dead_type, dead_level):
    """Decay-index sharding: the ranks' events are exchanged by time slice with a dead-time / coincidence-window
    halo over the process group (gloo here, NCCL on GPUs); each rank digitizes its list and keeps its slice.  With the
    oracle as the digitizer, the union over the ranks must be the digitization of ALL events as one list, byte for byte:
    singles (dead time per site across the cuts) and coincidences (windows across the cuts)."""
    import torch.multiprocessing as mp
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, q, 60000, dead_type, dead_level)) for r in range(world)]
    for p in procs:
        p.start()
    same_s, same_c, ns, nc, received, alive, sent = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ns > 15000 and nc > 100
    assert same_s and same_c
    # every live event reaches its slice's owner once, plus the halo copies: a few per cent of traffic, not a broadcast
    assert alive <= sum(received) < 1.1 * alive and all(s > 0 for s in sent)
