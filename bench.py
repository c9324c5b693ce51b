#!/usr/bin/env python
"""bench.py -- annihilation pairs/s of the gPET hot path (source -> phantom -> detector -> digitizer) on B200.

One "step" = one complete run of the shipped small-animal example as BASELINE.json configs[0] names it (input_PET.in,
8-panel config8.geo, 1 cm water cylinder in a 200^3 phantom, the shipped source.txt = 7 F-18 cylinders, 0-120 s
acquisition = ~1.118 M annihilation pairs), i.e. every frame of the acquisition through all four stages.  (The shipped
input_PET.in points at input/pointsource.txt, ~179 k pairs; `--source pointsource.txt` runs that variant, and its
numbers are reported under "extra" in the default line.)  `value` is measured with the inputs resident in HBM
(gpet_run_resident: only counters leave the device); `e2e` goes through the public C-ABI call a user makes (gpet_run:
frame planning + descriptor upload + all stages + singles/coincidences copied back to host memory).

N > 1 (torchrun): every rank runs the same acquisition with a disjoint Philox key (independent decay histories:
weak scaling, per-GPU work fixed); tallies are all-reduced over NCCL inside the timed region.

--impl reference: times the reference's own implementation of the path (the CUDA-12-patched reference binary under
oracle/_ref when it was built and a GPU is present; otherwise the single-thread CPU oracle port) on the same config.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "annihilation_pairs_per_s"
UNIT = "pairs/s"
DEFAULT_SOURCE = "source.txt"


def workload_name(source):
    return f"shipped small-animal example: input_PET.in, config8.geo (8 panels), 1 cm water cylinder 200^3, {source}, 0-120 s"


WORKLOAD = workload_name(DEFAULT_SOURCE)


def make_workdir(tmp, n=200, source=DEFAULT_SOURCE):
    from tools import gen_inputs
    ex_src = ROOT / "examples" / "small_animal"
    ex = Path(tmp) / "ex"
    (ex / "input").mkdir(parents=True)
    (ex / "data").mkdir()
    (ex / "output").mkdir()
    text = (ex_src / "input_PET.in").read_text().replace("input/pointsource.txt", f"input/{source}")
    (ex / "input_PET.in").write_text(text)
    for f in ("config8.geo", "pointsource.txt", "source.txt"):
        (ex / "input" / f).write_text((ex_src / "input" / f).read_text())
    (ex / "data" / "isotopes.txt").write_text((ex_src / "data" / "isotopes.txt").read_text())
    packed = ROOT / "gpet_b200" / "_data" / "input4gPET.gpettab"
    if not packed.exists():
        raise SystemExit("gpet_b200/_data/input4gPET.gpettab missing: run __graft_entry__.build() where /root/reference exists")
    os.symlink(packed, ex / "data" / "input4gPET.gpettab")
    mat, den = gen_inputs.cylinder_phantom(n=n)
    gen_inputs.write_phantom(mat, den, ex / "input" / "cylinder_phantom_mat.dat", ex / "input" / "cylinder_phantom_den.dat")
    return ex


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the measurement (B200_PROFILING.md recipe) through NVML, every 5 ms."""

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.sm, self.power, self.reasons = [], [], set()
        self.max_sm = None
        self._stop = threading.Event()
        self._t = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML indices follow the physical order; honour CUDA_VISIBLE_DEVICES if it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.idx
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                idx = int(vis.split(",")[self.idx])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            return
        names = {"hw_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4)}

        def loop():
            while not self._stop.is_set():
                try:
                    self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                    try:
                        mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:  # noqa: BLE001
                        mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for k, bit in names.items():
                        if mask & bit:
                            self.reasons.add(k)
                except Exception:  # noqa: BLE001
                    pass
                time.sleep(0.005)

        self._t = threading.Thread(target=loop, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=1.0)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None}


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU (one process per GPU: the pinned result
    arenas are then first-touched on that NUMA node and the D2H copies do not cross the socket interconnect)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = gpu_index
        if vis and all(x.strip().isdigit() for x in vis.split(",")):
            idx = int(vis.split(",")[gpu_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:  # noqa: BLE001
        return None


def cpu_baseline(ex, source=DEFAULT_SOURCE, npairs=6_000_000, chunk=1_000_000, seed=12345):
    """Oracle port on one host core over a bounded sample of the same workload: `npairs` pairs of the workload's source
    distribution through source -> phantom -> detector -> digitizer, in chunks of `chunk` pairs (one digitizer pass each)."""
    sys.path.insert(0, str(ROOT / "tests"))
    import parity
    from oracle import oracle as orc
    from gpet_b200 import refio
    s = parity.Setup(device=-1, phantom=parity.gen_inputs.cylinder_phantom(n=200), size=1.0)
    src = refio.parse_sources(ex / "input" / source)
    iso = refio.parse_isotopes(ex / "data" / "isotopes.txt")
    tau = np.array([np.float64(iso[x["type"]]["halftime"]) * 1.442695 for x in src])
    frac = -np.expm1(-120.0 / tau)
    weight = np.array([x["natom"] * frac[i] * iso[x["type"]]["ratio"] for i, x in enumerate(src)], np.float64)
    p, _ = parity.make_digi_params(blur_Rref=0.05, coinc_window_us=0.01)
    t0 = time.perf_counter()
    done = 0
    while done < npairs:
        n = min(chunk, npairs - done)
        per_src = np.floor(weight / weight.sum() * n).astype(np.uint64)
        per_src[-1] += np.uint64(n - int(per_src.sum()))
        ph = orc.source(np.cumsum(per_src), [x["shape"] for x in src], np.concatenate([x["coeff"] for x in src]),
                        tau, frac, 0.0, done, 0.0037056, n, seed)
        ph = orc.phantom(ph, s.mat, s.den, s.offset, s.size, s.tab_ph, s.eabs, seed)
        res = orc.detector(ph, s.panels, s.counts4, s.pmat, s.pdens, s.surfaces, s.tab_det, s.eabs, 2, 1, seed)
        orc.digitize(res["events"], p)
        done += n
    dt = time.perf_counter() - t0
    s.close()
    return {"value": npairs / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{npairs} pairs of the same workload ({source}, 200^3 phantom, config8.geo) through the "
                      f"single-thread C oracle (source+phantom+detector+digitizer) in chunks of {chunk}, {dt:.1f} s"}


def have_gpu():
    return shutil.which("nvidia-smi") is not None and subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0


def bench_config(source):
    """The `config` object of BOTH arms (identical keys and values: the driver compares them)."""
    return {"workload": workload_name(source),
            "step": "our arm: one step = the acquisition with its activity scaled 160 x (~179 M pairs per GPU, >= 40 ms of kernels) and, on N GPUs, "
                    "its duration scaled N x (0 - 120 N s: the same event rate at every N), cut into frames of ~8 M pairs (~4 M in the end-to-end steps; the frame is this "
                    "library's batch: per-frame fixed costs amortise with it), sharded by frames over the GPUs; reference arm: one step = the "
                    "acquisition at the shipped activity (1.118 M pairs, a bounded sample of the same workload -- the metric is per pair)",
            "l2": "working set of a frame (~2 GB of queues, hit / event / sort buffers) exceeds the 126 MB L2; 256 MiB flush between timed steps",
            "coincidence_window_us": 0.01, "rng": "Philox4x32-10, key 0x67504554, 64-bit history numbers", "time_path": "fp64",
            "multi_gpu": "one acquisition, frames sharded round robin over the ranks (gpet_set_shard), no data-path collective; "
                         "tallies all-reduced once over NCCL after the last step, inside the timed region"}


def reference_line(ex, args, source, steps, warmup):
    """The reference's own implementation of the path, timed on this box: its CUDA build (oracle/_ref, texture-object
    patch only) when it travelled with the snapshot and a GPU is present -- the reference has no CPU path at all
    (README.md:31) -- else the single-thread CPU oracle port."""
    from oracle import run_ref
    line = None
    if run_ref.available() and have_gpu():
        try:
            line = run_ref.bench_reference(ex, steps=steps, warmup=warmup, metric=METRIC, unit=UNIT, workload=workload_name(source))
            line["details"] = {k: v for k, v in line["config"].items() if k != "workload"}
        except Exception as e:  # noqa: BLE001
            line = None
            print(f"reference binary unusable ({type(e).__name__}: {e}); falling back to the CPU oracle port", file=sys.stderr)
    if line is None:
        base = cpu_baseline(ex, source, npairs=4_000_000)
        v = base["value"]
        line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
                "ms_per_step": 1e3 * 4_000_000 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "cpu_baseline": base,
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line["config"] = bench_config(source)
    line["n_gpus"] = args.gpus
    line["impl"] = "reference"
    return line


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    with tempfile.TemporaryDirectory() as tmp:
        ex = make_workdir(tmp, source=args.source)
        # --steps / --warmup as given: every step is one complete run of the reference binary (~2 s of process time each)
        print(json.dumps(reference_line(ex, args, args.source, steps=max(1, args.steps), warmup=max(0, args.warmup))))


# algorithmic bytes one launch of each kernel must move; c = per-frame counters
ALG_BYTES = {
    "k_front": lambda c: 48 * c["on_panel"],          # fused source+phantom+panel entry: only the entered photons are written
    "k_source": lambda c: 48 * c["photons"],
    "k_phantom": lambda c: 48 * c["photons"] + 48 * c["q1"],
    "k_panel_entry": lambda c: 48 * c["q1"] + 48 * c["on_panel"],
    "k_detector": lambda c: 48 * c["on_panel"] + 48 * c["hits"] + 44 * c["events"],
    # digitizer: 48-byte records in, per-event side arrays (u64 time key, site, arrival rank | window flag)
    "k_prep": lambda c: 48 * c["events"] + 8 * c["events"] + 8 * c["alive"],
    "k_bucket_scan": lambda c: 8 * 2 ** max(6, math.ceil(math.log2(max(c["events"], 2))) - 3),
    "k_bucket_scatter": lambda c: 8 * c["events"] + 8 * c["alive"] + 16 * c["alive"],
    "k_bucket_rank": lambda c: 16 * c["alive"] + 16 * c["alive"],
    "k_deadtime_chain": lambda c: 12 * c["alive"] + c["alive"],
    # singles: 48-byte record gathered and written, side arrays time (8) + panel, photon number, annihilation number (4 each)
    "k_emit_singles": lambda c: 16 * c["alive"] + 48 * c["singles"] + (48 + 20) * c["singles"],
    # the side arrays of every single, its scatter tag (1 byte), index pair + class byte per coincidence
    "k_coinc": lambda c: (20 + 1) * c["singles"] + (8 + 1) * c["coinc"],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--source", default=DEFAULT_SOURCE, help="source file of the example (source.txt | pointsource.txt)")
    ap.add_argument("--frames-per-step", type=int, default=160,
                    help="activity scale per GPU of a resident step: the shipped acquisition (1.118 M pairs) x this = ~179 M pairs, ~50 ms of kernels")
    ap.add_argument("--e2e-frames-per-step", type=int, default=64, help="activity scale per GPU of an end-to-end step (35 MB of results per 1.118 M pairs)")
    ap.add_argument("--frame-pairs", type=int, default=9_200_000,
                    help="frame capacity in pairs of the resident steps (the planner fills ~0.9 of it): per-frame fixed costs amortise with "
                         "the frame, tools/bigframes_sweep.py -- 304 us per M pairs at 1.1 M pairs a frame, 250 at 4.5 M, 242 at 8.9 M")
    ap.add_argument("--e2e-frame-pairs", type=int, default=4_600_000,
                    help="frame capacity of the end-to-end steps: smaller frames, more of them per step, so that the copy of a frame "
                         "overlaps the kernels of the next and the pipeline's fill and drain stay short")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.warmup < 3:
        args.warmup = 3

    # fd 1 is reserved for the ONE JSON line: anything a library prints there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from gpet_b200 import api, multi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)   # before any pinned allocation: result arenas land on the GPU's own NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    tmp = tempfile.TemporaryDirectory()
    # a real (non-NULL) stream: gpet_set_stream(NULL) means "library-owned stream", and the CUDA events below must be
    # recorded on the stream the kernels are launched on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def make_ctx(source, sub):
        ex = make_workdir(Path(tmp.name) / sub, source=source)
        ctx = api.Context(local)
        ctx.set_stream(stream.cuda_stream)
        # buffers for one frame: 2 photons per pair; ~1.01 hits and ~0.66 events per pair on this example (headroom on both)
        ctx.set_capacity(2 * args.frame_pairs, int(1.3 * args.frame_pairs), int(0.9 * args.frame_pairs))
        ctx.load_config_file(ex / "input_PET.in", base_dir=ex)
        ctx.set_digitizer(coinc_window_us=0.01)
        # coincidences reach the host as index pairs into the singles list (8 B each instead of two copied records)
        ctx.set_coincidence_format(api.Context.COINC_PAIRS)
        ctx.set_spectrum(128, 0.0, 1.0e6)
        ctx.base_atoms = [int(x["natom"]) for x in ctx.sources()]
        from gpet_b200 import refio
        cfg = refio.parse_config(ex / "input_PET.in")
        ctx.base_window = (float(cfg["tstart"]), float(cfg["tend"]))
        return ex, ctx

    def plan(ctx, frames_per_gpu, nranks=None, frame_pairs=None):
        """ONE acquisition: the shipped one with `frames_per_gpu` times its activity and, on N GPUs, N times its duration
        (0 .. 120 N s).  The event RATE -- what dead time, coincidence windows and the time sort see -- is then the same at every
        N (F-18, T1/2 6586 s: 5 % lower on average over 960 s), so every GPU does the same work per pair; scaling the activity
        with N instead changes the workload itself (measured at N = 8: 7.5 x the random coincidences, 1.5 % fewer singles per
        pair, +6 % digitizer time per pair on EVERY rank, profiles/r02r_bench_n8_activity_scaled.json).  The planner cuts the
        acquisition into frames of ~8 M pairs (~4 M in the end-to-end steps; same seed on every rank: same plan), rank r runs frames f with f % world == r."""
        nranks = world if nranks is None else nranks
        scale = max(1, frames_per_gpu)
        for i, n in enumerate(ctx.base_atoms):
            ctx.set_source_atoms(i, n * scale)
        t0, t1 = ctx.base_window
        ctx.set_time_window(t0, t0 + (t1 - t0) * nranks)
        cap = int(frame_pairs or args.frame_pairs)
        nf = ctx.plan_frames(cap)
        # every rank the same number of frames: a slightly smaller frame capacity until the count divides by the world size
        # (346 frames over 8 ranks would leave two ranks with 44 frames and six with 43: the step is the slowest rank's)
        for _ in range(12):
            if nf % nranks == 0:
                break
            cap = int(cap * nf / (nranks * ((nf + nranks - 1) // nranks))) - 1
            nf = ctx.plan_frames(cap)
        ctx.set_shard(rank if nranks == world else 0, nranks)
        return nf

    tally_dev = torch.zeros(len(multi.TALLY_FIELDS), dtype=torch.int64, device=dev)

    def timed(ctx, nsteps, resident=True, reduce_at_end=False):
        """`nsteps` steps, each bracketed by CUDA events on the launching stream (resident) or by the wall clock (end to
        end: host results), L2 flushed in between.  The tallies of all steps are all-reduced ONCE, after the last step,
        inside the timed region (the reduction's own device time is added)."""
        times, stats = [], []
        for k in range(nsteps):
            flush.fill_(1)                      # L2 flush between timed iterations
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            w0 = time.perf_counter()
            st = ctx.run_resident() if resident else ctx.run(None)
            if reduce_at_end and k == nsteps - 1 and world > 1:
                acc = np.sum([multi.stats_vector(x) for x in stats + [st]], axis=0)
                tally_dev.copy_(torch.from_numpy(acc), non_blocking=True)
                dist.all_reduce(tally_dev)      # NCCL on this stream: inside the last step's event bracket
            e1.record(stream)
            torch.cuda.synchronize()
            w1 = time.perf_counter()
            times.append(max(e0.elapsed_time(e1), 0.0) if resident else (w1 - w0) * 1e3)
            stats.append(st)
        return times, stats

    def reduce_max_sum(ms, *counts):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        c = torch.tensor(list(counts), dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(c)
        return float(t.item()), [int(x) for x in c.tolist()]

    def measure(ctx, steps, warmup, frames_per_step, e2e_frames_per_step):
        nf = plan(ctx, frames_per_step)
        timed(ctx, warmup)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        times, stats = timed(ctx, steps, reduce_at_end=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        total_ms, (pairs, coinc) = reduce_max_sum(sum(times), sum(s.pairs for s in stats), sum(s.coincidences for s in stats))
        rank_ms = torch.tensor([sum(times) / max(len(times), 1)], dtype=torch.float64, device=dev)   # every rank's mean step, for the record
        all_ms = [torch.zeros_like(rank_ms) for _ in range(world)]
        if world > 1:
            dist.all_gather(all_ms, rank_ms)
        else:
            all_ms = [rank_ms]
        # e2e through gpet_run: planning + all stages + results in pinned host memory, wall clock.  Singles travel as 32-byte
        # gpet_single_compact records (the run is bound by their copy; gpet_result_singles rebuilds the 48-byte Event byte for
        # byte, tests/test_gpu_parity.py); the same run with 48-byte records is measured beside it
        nf_e = plan(ctx, e2e_frames_per_step, frame_pairs=args.e2e_frame_pairs)
        e2e = {}
        for name, fmt in (("records48", api.Context.SINGLES_RECORDS), ("compact32", api.Context.SINGLES_COMPACT)):
            ctx.set_singles_format(fmt)
            timed(ctx, 2, resident=False)
            if world > 1:
                dist.barrier()
            e_times, e_stats = timed(ctx, max(3, min(steps, 10)), resident=False)
            e_ms, (e_pairs,) = reduce_max_sum(sum(e_times), sum(s.pairs for s in e_stats))
            e2e[name] = (e_ms, e_pairs, e_stats)
        ctx.set_singles_format(api.Context.SINGLES_RECORDS)
        e_ms, e_pairs, e_stats = e2e["compact32"]
        step_ms = [round(x, 3) for x in times]
        return {"total_ms": total_ms, "pairs": pairs, "coinc": coinc, "stats": stats, "frames": nf, "e2e_ms": e_ms, "e2e_pairs": e_pairs,
                "e2e_stats": e_stats, "e2e_frames": nf_e, "e2e_records48": e2e["records48"][1] / (e2e["records48"][0] * 1e-3),
                "step_ms": step_ms, "rank_ms": [round(float(x.item()), 3) for x in all_ms], "tallies": dict(zip(multi.TALLY_FIELDS, (int(x) for x in tally_dev.tolist())))}

    ex, ctx = make_ctx(args.source, "main")
    m = measure(ctx, args.steps, args.warmup, args.frames_per_step, args.e2e_frames_per_step)
    value = m["pairs"] / (m["total_ms"] * 1e-3)
    st = m["e2e_stats"][-1]
    h2d = int(st.frames) * 4096 + 64
    d2h = int(st.singles * 32 + st.coincidences * (8 + 1) + 32 * 4 * st.frames)   # compact singles, index pairs + class bytes, counters

    # ---- same deliverable as the reference's timed region: adder.dat + singles.dat appended frame by frame (gPET.cu:383, 424;
    # its hit dumps off as in gPET_nodump), wall clock of gpet_run(output_dir).  One GPU only (one host, one file system).
    e2e_files = None
    if world == 1 and not args.no_extra:
        plan(ctx, 8, frame_pairs=args.e2e_frame_pairs)
        ctx.set_transport(record_hits=0)
        ctx.set_digitizer(coinc_window_us=0.0)
        od = Path(tmp.name) / "files_out"
        f_pairs, f_ms, f_bytes = 0, 0.0, 0
        for k in range(4):
            if od.exists():
                shutil.rmtree(od)
            od.mkdir()
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            stf = ctx.run(od)
            w1 = time.perf_counter()
            if k:                                # first pass = warm-up (pinned arenas grow)
                f_pairs += int(stf.pairs); f_ms += (w1 - w0) * 1e3
                f_bytes += (od / "adder.dat").stat().st_size + (od / "singles.dat").stat().st_size
        e2e_files = {"value": f_pairs / (f_ms * 1e-3), "unit": UNIT, "bytes_written_per_step": f_bytes // 3, "steps": 3,
                     "files": "adder.dat + singles.dat, appended frame by frame by a pool of writer threads while the next frames compute",
                     "timing": "wall clock around gpet_run(output_dir), files complete on return; same directory tree (tmp) as the reference arm's runs"}
        ctx.set_transport(record_hits=1)
        ctx.set_digitizer(coinc_window_us=0.01)

    extra = None
    if not args.no_extra and world == 1 and args.source == DEFAULT_SOURCE:
        ex2, ctx2 = make_ctx("pointsource.txt", "extra")
        m2 = measure(ctx2, max(5, min(args.steps, 10)), 3, args.frames_per_step, args.e2e_frames_per_step)
        extra = {"workload": workload_name("pointsource.txt"), "value": m2["pairs"] / (m2["total_ms"] * 1e-3),
                 "e2e": m2["e2e_pairs"] / (m2["e2e_ms"] * 1e-3), "pairs_per_step": m2["pairs"] / len(m2["stats"]), "unit": UNIT}
        ctx2.close()

    # ---- the exchange path (decay-index sharding + singles exchanged by time slice with halo over NCCL), measured beside the
    # frame-sharded headline: one acquisition of `world` x the shipped activity
    exchange = None
    if world > 1 and not args.no_extra:
        try:
            atoms = [n * world for n in ctx.base_atoms]
            ctx.set_time_window(*ctx.base_window)
            multi.run_exchange(ctx, atoms, dev, frame_pairs=args.frame_pairs)          # warm-up
            dist.barrier(); torch.cuda.synchronize()
            w0 = time.perf_counter()
            reps = 5
            acc = None
            for _ in range(reps):
                r = multi.run_exchange(ctx, atoms, dev, frame_pairs=args.frame_pairs)
                acc = r if acc is None else {k: (acc[k] + r[k] if k != "halo_flag" else acc[k] | r[k]) for k in r}
            torch.cuda.synchronize(); dist.barrier()
            x_ms, (x_pairs, x_sing, x_co, x_bytes, x_flag) = reduce_max_sum((time.perf_counter() - w0) * 1e3, acc["pairs"], acc["singles"],
                                                                           acc["coincidences"], acc["bytes_sent"], acc["halo_flag"])
            exchange = {"value": x_pairs / (x_ms * 1e-3), "unit": UNIT, "pairs": x_pairs // reps, "singles": x_sing // reps,
                        "coincidences": x_co // reps, "nccl_bytes_per_acquisition": x_bytes // reps, "halo_too_short": int(x_flag > 0),
                        "timing": "wall clock, stage-level calls from Python (unfused: a reference point for the exchange, not the headline)"}
        except Exception as e:  # noqa: BLE001
            exchange = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        # ---- per-kernel device times (CUDA events around every launch, on the launching stream) -> roofline
        plan(ctx, 8, nranks=1)
        ctx.profile(True)
        nprof = 3
        for _ in range(nprof):
            flush.fill_(1)
            sp = ctx.run_resident()
        kt = ctx.kernel_times()
        ctx.profile(False)
        fr = max(int(sp.frames), 1)
        cnt = {"photons": 2 * sp.pairs / fr, "q1": sp.photons_phantom_out / fr, "on_panel": sp.photons_on_panel / fr,
               "hits": sp.hits / fr, "events": sp.events_adder / fr, "alive": sp.events_threshold / fr,
               "singles": sp.singles / fr, "coinc": sp.coincidences / fr}
        kernels = {}
        for name, (ms, n) in kt.items():
            alg = float(ALG_BYTES.get(name, lambda c: 0)(cnt))
            kernels[name] = {"launches_per_frame": n / nprof / fr, "us_per_launch": 1e3 * ms / max(n, 1), "us_per_frame": 1e3 * ms / nprof / fr,
                             "algorithmic_bytes_per_launch": alg, "achieved_gbs": alg / (ms / max(n, 1) * 1e-3) / 1e9 if ms > 0 else 0.0}
        top = max(kernels, key=lambda k: kernels[k]["us_per_frame"])
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None
        tf = ROOT / "profiles" / "traffic.json"   # dram bytes per launch from the committed `ncu --set full` captures
        if tf.exists():
            tj = json.loads(tf.read_text())
            traffic = tj.get(args.source, {}).get(top)
            if traffic is not None:   # the capture is of a 1.118 M-pair frame: per launch of THIS frame size (DRAM traffic follows the photons)
                traffic = float(traffic) * (sp.pairs / fr) / float(tj.get("_pairs_per_frame", sp.pairs / fr))
        roofline = {"bound": "hbm", "kernel": top, "achieved": kernels[top]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[top]["achieved_gbs"] / peak, "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if pk.exists() else "fallback 6650 GB/s (B200_PROFILING.md)",
                    "us_per_launch": kernels[top]["us_per_launch"], "share_of_step": kernels[top]["us_per_frame"] / sum(k["us_per_frame"] for k in kernels.values()),
                    "timing": "CUDA events bracketing every launch on the launching stream (gpet_profile_enable), 3 runs of the acquisition at 8 x the shipped activity (8.9 M pairs in %d frame(s)), L2 flushed between runs; traffic: ncu dram bytes of a 1.118 M-pair frame (profiles/traffic.json) scaled to this frame's pairs" % fr,
                    "note": "Monte-Carlo transport is latency/issue bound: the algorithmic bytes are tiny against HBM (DESIGN.md section 4); see profiles/ for issue-slot, SIMT-efficiency and pipe numbers",
                    "kernels": kernels}
        base = None
        if not args.no_cpu_baseline:
            try:
                ref = reference_line(ex, args, args.source, steps=5, warmup=1)   # median of five runs (two were too noisy)
                base = ref["cpu_baseline"]
                base["value"] = ref["value"]
                if base.get("kind") == "reference":
                    base["reference_counters"] = ref.get("reference_counters")
                    base["oracle_port"] = cpu_baseline(ex, args.source, npairs=3_000_000)
            except Exception as e:  # noqa: BLE001
                base = cpu_baseline(ex, args.source)
                base["note"] = f"reference binary failed: {e}"
        clocks = sampler.stop()
        nsteps = len(m["stats"])
        s0 = m["stats"][-1]
        tl = m["tallies"] if world > 1 else dict(zip(multi.TALLY_FIELDS, (int(x) for x in np.sum([multi.stats_vector(x) for x in m["stats"]], axis=0))))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": m["total_ms"] / nsteps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": bench_config(args.source),
                "details": {"pairs_per_step_per_gpu": m["pairs"] / nsteps / world, "frames_in_the_acquisition": int(m["frames"]),
                            "frames_per_step_per_gpu": int(s0.frames), "activity_scale": args.frames_per_step, "duration_scale": world,
                            "e2e_frames_per_step_per_gpu": int(st.frames), "e2e_activity_scale": args.e2e_frames_per_step,
                            "frame_capacity_pairs": args.frame_pairs, "step_ms_rank0": m["step_ms"], "mean_step_ms_by_rank": m["rank_ms"]},
                "clocks": clocks,
                "e2e": {"value": m["e2e_pairs"] / (m["e2e_ms"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "singles_format": "gpet_single_compact, 32 B (GPET_SINGLES_COMPACT: t, E, x, y, z, eventid, packed panel / module / crystal / photon bit; gpet_result_singles expands to the 48-byte Event on demand, byte-identical)",
                        "with_48_byte_records": {"value": m["e2e_records48"], "unit": UNIT, "d2h_bytes_per_step": int(st.singles * 48 + st.coincidences * 9 + 128 * st.frames)},
                        "timing": "wall clock around gpet_run: all stages, singles and coincidences (index pairs into the singles plus one class byte each) delivered to pinned host memory, the copy of a frame overlapping the kernels of the next; max over ranks"},
                "e2e_files": e2e_files,
                "gpu_launches": int(sum(s.kernel_launches for s in m["stats"])),
                "roofline": roofline, "cpu_baseline": base,
                "counters": {"pairs": tl["pairs"], "hits": tl["hits"], "events_adder": tl["events_adder"], "singles": tl["singles"],
                             "coincidences": tl["coincidences"], "trues": tl["trues"], "scatters": tl["scatters"], "randoms": tl["randoms"],
                             "of": "all timed steps, all ranks (one NCCL all-reduce after the last step)" if world > 1 else "all timed steps",
                             "coincidences_per_s": float(m["coinc"] / (m["total_ms"] * 1e-3))},
                "exchange": exchange, "extra": extra}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    tmp.cleanup()


if __name__ == "__main__":
    main()
