#!/usr/bin/env python
"""BASELINE.json configs 4 and 5 at the scale they are named at: the 256^3 mouse phantom at 1e9 decays, the 32-panel ring
with the 20 cm water cylinder at 1e10 decays, ONE acquisition each, its frames sharded over the GPUs of one box
(gpet_set_shard: rank r runs frames f with f % world == r; no data-path collective, tallies all-reduced once over NCCL).
History numbers are 64 bits wide (2e10 photons do not fit the reference's 32-bit ids, gPET.h:50).

Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
             tools/scale_runs.py --config config5_ring --decays 1e10
Rank 0 prints one JSON line: totals of the acquisition (pairs, singles, coincidences, true / scatter / random classes),
device time (max over ranks, CUDA events around the resident run), pairs/s."""
import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="config4_mouse", choices=["config4_mouse", "config5_ring", "config1_water10"])
    ap.add_argument("--decays", type=float, default=1e9)
    ap.add_argument("--frame-pairs", type=int, default=1_240_000)
    a = ap.parse_args()
    # fd 1 is reserved for the ONE JSON line (NCCL prints its version banner there)
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import torch.distributed as dist
    from gpet_b200 import api, multi
    from tools import gen_inputs
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    decays = int(a.decays)
    with tempfile.TemporaryDirectory() as tmp:
        cfg = gen_inputs.stats_config(a.config, decays)
        ex = gen_inputs.write_workdir(Path(tmp) / "ex", cfg, ROOT / "examples" / "small_animal", ROOT / "gpet_b200" / "_data" / "input4gPET.gpettab")
        stream = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(stream)
        with api.Context(local) as c:
            c.set_stream(stream.cuda_stream)
            c.load_config_file(ex / "input_PET.in", base_dir=ex)
            c.set_digitizer(coinc_window_us=0.01, coinc_min_panel_diff=4 if a.config == "config5_ring" else 0)
            c.set_coincidence_format(api.Context.COINC_PAIRS)
            c.set_spectrum(128, 0.0, 1.0e6)
            t0 = time.perf_counter()
            nf = c.plan_frames(a.frame_pairs)
            plan_s = time.perf_counter() - t0
            c.set_shard(rank, world)
            last = c.frame(nf - 1)
            # warm-up outside the timed region: the first compute call uploads the phantom (67 MB), the tables and the panels
            c.stage_front(rank % nf); c.stage_panel_transport(); c.stage_digitize()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            w0 = time.perf_counter()
            e0.record(stream)
            st = c.run_resident()
            tally = torch.from_numpy(multi.stats_vector(st)).to(dev)
            spec = torch.from_numpy(c.spectrum(128).astype(np.int64)).to(dev)
            if world > 1:
                dist.all_reduce(tally); dist.all_reduce(spec)
            e1.record(stream)
            torch.cuda.synchronize()
            wall = time.perf_counter() - w0
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            tot = dict(zip(multi.TALLY_FIELDS, (int(x) for x in tally.tolist())))
            if rank == 0:
                os.write(json_fd, (json.dumps({"config": a.config, "decays_requested": decays, "n_gpus": world, "frames": nf, "planning_s": plan_s,
                                  "device_ms_max_over_ranks": float(ms.item()), "wall_s_rank0": wall,
                                  "pairs_per_s": tot["pairs"] / (float(ms.item()) * 1e-3),
                                  "first_pair_of_last_frame": int(last["first_pair"]), "photon_index_bits": int(2 * (int(last["first_pair"]) + 1)).bit_length(),
                                  "totals": tot, "scatter_fraction": tot["scatters"] / max(tot["trues"] + tot["scatters"], 1),
                                  "randoms_fraction": tot["randoms"] / max(tot["coincidences"], 1),
                                  "singles_per_pair": tot["singles"] / max(tot["pairs"], 1), "coincidences_per_pair": tot["coincidences"] / max(tot["pairs"], 1),
                                  "singles_spectrum_128_bins_0_1MeV": [int(x) for x in spec.tolist()]}) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
