#!/usr/bin/env python
"""Strong scaling of ONE acquisition over the GPUs of a node by frame sharding (gpet_b200/multi.py): every rank plans the
same frames, runs those with frame % world == rank, and the tallies are all-reduced over NCCL.  The workload is the
shipped example's source.txt scaled to many frames (--scale multiplies the atoms).  Prints one JSON line on rank 0.
Usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/strong_scaling.py [--scale 16]"""
import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=16.0)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from gpet_b200 import api, multi
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    with tempfile.TemporaryDirectory() as tmp:
        ex = bench.make_workdir(tmp, source="source.txt")
        c = api.Context(local)
        c.load_config_file(ex / "input_PET.in", base_dir=ex)          # same seed on every rank: one acquisition
        c.set_digitizer(coinc_window_us=0.01)
        c.set_coincidence_format(api.Context.COINC_PAIRS)
        for i, s in enumerate(c.sources()):
            c.set_source_atoms(i, int(s["natom"] * a.scale))
        c.set_spectrum(128, 0.0, 1.0e6)
        mine, my_pairs, all_pairs = multi.plan_shard(c, 0)
        times = []
        tot = None
        for rep in range(a.reps + 1):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st, tot = multi.run_sharded(c, resident=True, device=dev)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            if rep:
                times.append(time.perf_counter() - t0)
        t = torch.tensor([min(times)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"n_gpus": world, "frames": int(tot["frames"]), "pairs": int(tot["pairs"]), "singles": int(tot["singles"]),
                              "coincidences": int(tot["coincidences"]), "trues": int(tot["trues"]), "scatters": int(tot["scatters"]),
                              "randoms": int(tot["randoms"]), "seconds": float(t.item()),
                              "pairs_per_s": tot["pairs"] / float(t.item()), "scaling": "strong (one acquisition, frames sharded)"}))
        c.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
