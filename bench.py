#!/usr/bin/env python
"""bench.py -- annihilation pairs/s of the gPET hot path (source -> phantom -> detector -> digitizer) on B200.

One "step" = one complete run of the shipped small-animal example (BASELINE.json configs[0]: input_PET.in, 8-panel
config8.geo, 1 cm water cylinder in a 200^3 phantom, pointsource.txt, 0-120 s acquisition = ~178.7k pairs), i.e. every
frame of the acquisition through all four stages.  `value` is measured with the inputs resident in HBM
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
WORKLOAD = "shipped small-animal example: input_PET.in, config8.geo (8 panels), 1 cm water cylinder 200^3, pointsource.txt, 0-120 s"


def make_workdir(tmp, n=200, source="pointsource.txt"):
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
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.p:
            self.p.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(ex, npairs=8_000_000, chunk=1_000_000, seed=12345):
    """Oracle port on one host core over a bounded sample of the same workload: `npairs` pairs of the shipped point source
    through source -> phantom -> detector -> digitizer, in chunks of `chunk` pairs (one digitizer pass per chunk)."""
    sys.path.insert(0, str(ROOT / "tests"))
    import parity
    from oracle import oracle as orc
    from gpet_b200 import refio
    s = parity.Setup(device=-1, phantom=parity.gen_inputs.cylinder_phantom(n=200), size=1.0)
    src = refio.parse_sources(ex / "input" / "pointsource.txt")
    iso = refio.parse_isotopes(ex / "data" / "isotopes.txt")
    tau = np.array([np.float64(iso[x["type"]]["halftime"]) * 1.442695 for x in src])
    frac = -np.expm1(-120.0 / tau)
    p, _ = parity.make_digi_params(blur_Rref=0.05, coinc_window_us=0.01)
    t0 = time.perf_counter()
    done = 0
    while done < npairs:
        n = min(chunk, npairs - done)
        ph = orc.source(np.array([n], np.uint64), [x["shape"] for x in src], np.concatenate([x["coeff"] for x in src]),
                        tau, frac, 0.0, done, 0.0037056, n, seed)
        ph = orc.phantom(ph, s.mat, s.den, s.offset, s.size, s.tab_ph, s.eabs, seed)
        res = orc.detector(ph, s.panels, s.counts4, s.pmat, s.pdens, s.surfaces, s.tab_det, s.eabs, 2, 1, seed)
        orc.digitize(res["events"], p)
        done += n
    dt = time.perf_counter() - t0
    s.close()
    return {"value": npairs / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{npairs} pairs of the same workload (shipped point source, 200^3 phantom, config8.geo) through the "
                      f"single-thread C oracle (source+phantom+detector+digitizer) in chunks of {chunk}, {dt:.1f} s"}


def run_reference(args):
    """Reference arm: the patched reference CUDA binary when available on a GPU box, else the CPU oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    with tempfile.TemporaryDirectory() as tmp:
        ex = make_workdir(tmp)
        line = None
        from oracle import run_ref
        have_gpu = shutil.which("nvidia-smi") is not None and subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0
        if run_ref.available() and have_gpu:
            try:
                line = run_ref.bench_reference(ex, steps=max(1, min(args.steps, 5)), warmup=max(1, min(args.warmup, 1)),
                                               metric=METRIC, unit=UNIT, workload=WORKLOAD)
                line["n_gpus"] = args.gpus
            except Exception as e:  # noqa: BLE001
                line = None
                print(f"reference binary unusable ({type(e).__name__}: {e}); falling back to the CPU oracle port", file=sys.stderr)
        if line is None:
            vals = []
            base = None
            for _ in range(max(1, min(args.steps, 3))):
                base = cpu_baseline(ex, npairs=4_000_000)
                vals.append(base["value"])
            v = float(np.mean(vals))
            base["value"] = v
            line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals), "warmup": 0,
                    "ms_per_step": 1e3 * 4_000_000 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD},
                    "cpu_baseline": base,
                    "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        line["impl"] = "reference"
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--activity-scale", type=float, default=1.0, help="multiply the source atoms (extra line only; default = shipped file)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from gpet_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    tmp = tempfile.TemporaryDirectory()
    ex = make_workdir(tmp.name)
    ctx = api.Context(local)
    # a real (non-NULL) stream: gpet_set_stream(NULL) means "library-owned stream", and the CUDA events below must be
    # recorded on the stream the kernels are launched on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_seed(0x67504554 + 1000003 * rank)   # disjoint Philox keys per rank
    ctx.load_config_file(ex / "input_PET.in", base_dir=ex)
    ctx.set_digitizer(coinc_window_us=0.01)
    if args.activity_scale != 1.0:
        ctx.set_source_atoms(0, int(1762974000 * args.activity_scale))
    ctx.set_spectrum(128, 0.0, 1.0e6)
    nframes = ctx.plan_frames(0)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    tally = torch.zeros(16, dtype=torch.int64, device=dev)

    def one_step(resident=True):
        st = ctx.run_resident() if resident else ctx.run(None)
        if world > 1:
            tally[:9] = torch.tensor([st.pairs, st.photons_phantom_out, st.photons_on_panel, st.hits, st.events_adder,
                                      st.events_threshold, st.events_deadtime, st.singles, st.coincidences],
                                     dtype=torch.int64, device=dev)
            dist.all_reduce(tally)
        return st

    def timed(nsteps, resident=True):
        times, stats = [], []
        for _ in range(nsteps):
            flush.fill_(1)                      # L2 flush between timed iterations
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            if not resident:
                ctx.plan_frames(0)              # e2e: planning + descriptor upload are part of the user's call
            e0.record(stream)
            w0 = time.perf_counter()
            st = one_step(resident)
            e1.record(stream)
            torch.cuda.synchronize()
            w1 = time.perf_counter()
            times.append(max(e0.elapsed_time(e1), 0.0) if resident else (w1 - w0) * 1e3)
            stats.append(st)
        return times, stats

    timed(args.warmup)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    times, stats = timed(args.steps)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = torch.tensor([sum(times)], dtype=torch.float64, device=dev)
    pairs = torch.tensor([sum(s.pairs for s in stats)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(pairs)
    total_ms = float(total_ms.item())
    total_pairs = int(pairs.item())
    value = total_pairs / (total_ms * 1e-3)

    # ---- e2e through gpet_run (host results), wall clock around the call incl. planning/upload and D2H
    timed(2, resident=False)
    e2e_times, e2e_stats = timed(max(3, min(args.steps, 10)), resident=False)
    e2e_ms = torch.tensor([sum(e2e_times)], dtype=torch.float64, device=dev)
    e2e_pairs = torch.tensor([sum(s.pairs for s in e2e_stats)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_pairs)
    st = e2e_stats[-1]
    h2d = nframes * 4096 + 64
    d2h = int(st.singles * 48 + st.coincidences * 96 + 21 * 4 * st.frames)

    if rank == 0:
        # ---- per-stage device times (CUDA events on the launching stream) for the roofline of the dominant kernel
        stage_ms = {"source": [], "phantom": [], "detector": [], "digitizer": []}
        for _ in range(max(5, args.warmup)):
            flush.fill_(1)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            ev[0].record(stream); ctx.stage_source(0)
            ev[1].record(stream); ctx.stage_phantom()
            ev[2].record(stream); ctx.stage_detector()
            ev[3].record(stream); ctx.stage_digitize()
            ev[4].record(stream)
            torch.cuda.synchronize()
            for k, name in enumerate(stage_ms):
                stage_ms[name].append(ev[k].elapsed_time(ev[k + 1]))
        stage_med = {k: statistics.median(v[1:]) for k, v in stage_ms.items()}
        s0 = stats[-1]
        fp = ctx.frame_pairs(0)
        n_q0 = 2 * fp
        # algorithmic bytes per launch (DESIGN.md "kernels"): 48 B photon records, 48 B hit rows, 44 B event columns
        alg = {"source": 48 * n_q0,
               "phantom": 48 * n_q0 + 48 * s0.photons_phantom_out / max(s0.frames, 1),
               "detector": 48 * s0.photons_phantom_out / max(s0.frames, 1) + 48 * s0.hits / max(s0.frames, 1) + 44 * s0.events_adder / max(s0.frames, 1),
               "digitizer": 480 * s0.events_adder / max(s0.frames, 1)}
        top = max(("source", "phantom", "detector", "digitizer"), key=lambda k: stage_med[k])
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = alg[top] / (stage_med[top] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": {"source": "k_source", "phantom": "k_phantom", "detector": "k_detector",
                                                "digitizer": "digitizer chain (k_prep + radix passes + k_deadtime + compaction)"}[top],
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json (measured)" if pk.exists() else "fallback",
                    "stage_ms": stage_med, "algorithmic_bytes": {k: float(v) for k, v in alg.items()},
                    "note": "latency/issue bound Monte-Carlo: algorithmic bytes are tiny, see DESIGN.md and profiles/"}
        base = None if args.no_cpu_baseline else cpu_baseline(ex)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": total_pairs / args.steps / world,
                           "frames_per_step": int(stats[-1].frames), "l2": "flushed between timed steps (256 MiB write)",
                           "activity_scale": args.activity_scale, "coincidence_window_us": 0.01, "rng": "Philox4x32-10"},
                "clocks": clocks,
                "e2e": {"value": int(e2e_pairs.item()) / (float(e2e_ms.item()) * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "timing": "wall clock around gpet_plan_frames + gpet_run, max over ranks"},
                "gpu_launches": int(sum(s.kernel_launches for s in stats)),
                "roofline": roofline, "cpu_baseline": base,
                "counters": {"pairs": int(s0.pairs), "hits": int(s0.hits), "events_adder": int(s0.events_adder),
                             "singles": int(s0.singles), "coincidences": int(s0.coincidences),
                             "coincidences_per_s": float(sum(s.coincidences for s in stats) * world / (total_ms * 1e-3))}}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    tmp.cleanup()


if __name__ == "__main__":
    main()
