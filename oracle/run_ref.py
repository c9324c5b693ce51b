#!/usr/bin/env python
"""Run the CUDA-12-patched reference binary (oracle/_ref/gPET*, built by oracle/build_ref.py) in a scratch directory
laid out the way the reference expects (./data, ./input, ./output relative to cwd; main.cu:50, initialize.cu:13-672)
and parse the counters it prints.  TEST / BASELINE INFRASTRUCTURE ONLY -- nothing under gpet_b200/ imports this."""
from __future__ import annotations

import json
import os
import re
import shutil
import subprocess
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REFDIR = ROOT / "oracle" / "_ref"
TABLE_EXTS = ("matter", "lamph", "compt", "cmpsf", "phote", "rayle", "rayff")


def available(binname="gPET_nodump"):
    return (REFDIR / binname).exists() and all((REFDIR / "data" / f"input4gPET.{e}").exists() for e in TABLE_EXTS)


def prepare_workdir(ex: Path):
    """`ex` already holds input_PET.in, input/*, data/isotopes.txt and the phantom volumes (bench.make_workdir);
    add the reference's ASCII table set and an empty output directory."""
    ex = Path(ex)
    (ex / "data").mkdir(exist_ok=True)
    for e in TABLE_EXTS:
        dst = ex / "data" / f"input4gPET.{e}"
        if not dst.exists():
            os.symlink(REFDIR / "data" / f"input4gPET.{e}", dst)
    out = ex / "output"
    if out.exists():
        shutil.rmtree(out)   # the reference appends to its output files (gPET.cu:371-383)
    out.mkdir()


def run_once(ex: Path, binname="gPET_nodump", input_file="input_PET.in", timeout=600):
    prepare_workdir(ex)
    t0 = time.perf_counter()
    r = subprocess.run([str(REFDIR / binname), input_file], cwd=ex, capture_output=True, text=True, timeout=timeout)
    wall = time.perf_counter() - t0
    out = r.stdout
    res = {"returncode": r.returncode, "process_wall_s": wall, "stdout_tail": out[-1500:], "stderr_tail": r.stderr[-500:]}
    emitted = [int(x) for x in re.findall(r"currently emitted photons (\d+)", out)]
    res["pairs"] = sum(emitted) // 2
    res["epochs"] = len(emitted)
    for key, pat in (("sim_wall_s", r"Simulation wall time: ([\d.eE+-]+) s"), ("sim_cpu_s", r"Simulation time: ([\d.eE+-]+) s"),
                     ("total_wall_s", r"Total wall time: ([\d.eE+-]+) s"), ("init_cpu_s", r"Initialize time: ([\d.eE+-]+) s")):
        m = re.search(pat, out)
        res[key] = float(m.group(1)) if m else None
    for key, pat in (("hits", r"there are (\d+) Hits in this batch"), ("events_adder", r"counts of events after adder is (\d+)"),
                     ("events_threshold", r"counts of events after thresholder is (\d+)"),
                     ("events_deadtime", r"counts of events after deadtime is (\d+)"), ("singles", r"counts of singles is (\d+)")):
        res[key] = sum(int(x) for x in re.findall(pat, out))
    return res


def bench_reference(ex: Path, steps=3, warmup=1, binname="gPET_nodump", metric="annihilation_pairs_per_s", unit="pairs/s",
                    workload=""):
    """`steps` complete runs of the reference on the work directory; the timed region is the one the reference itself
    reports as "Simulation time" (sampleParticle / simulateParticle body, gPET.cu:245-247 -> 432-435), by wall clock."""
    runs = []
    for k in range(warmup + steps):
        r = run_once(ex, binname)
        if r["returncode"] != 0 or not r["sim_wall_s"] or r["pairs"] <= 0:
            raise RuntimeError(f"reference run failed (rc={r['returncode']}): {r['stdout_tail'][-400:]} {r['stderr_tail']}")
        if k >= warmup:
            runs.append(r)
    # the median run (the reference's epochs wait on host sorts and file appends: a run now and then takes 3x as long
    # on a busy box, and a mean would flatter the comparison); every run's time is kept in reference_times
    med = sorted(runs, key=lambda r: r["sim_wall_s"])[len(runs) // 2]
    sim = med["sim_wall_s"] * len(runs)
    v = med["pairs"] / med["sim_wall_s"]
    ncores = os.cpu_count()
    coinc = reference_coincidences(ex, med)
    # one run of the binary as shipped (OUTPUTHIT = 1: Hits.dat / HitsID.dat written per epoch), so that the comparison
    # above is visibly not an I/O contest
    dumps = None
    if (REFDIR / "gPET").exists():
        try:
            r = run_once(ex, "gPET")
            if r["returncode"] == 0 and r["sim_wall_s"]:
                dumps = {"binary": "gPET", "sim_wall_s": r["sim_wall_s"], "pairs": r["pairs"], "value": r["pairs"] / r["sim_wall_s"]}
        except Exception as e:  # noqa: BLE001
            dumps = {"note": f"unavailable: {type(e).__name__}: {e}"}
    return {"metric": metric, "value": v, "unit": unit, "n_gpus": 1, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * sim / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": workload, "binary": binname,
                                            "statistic": "median run", "timed_region": "sampleParticle body by wall clock (the reference's own 'Simulation time' region); process start-up, table parsing and curand_init excluded"},
            "cpu_baseline": {"value": v, "unit": unit, "cores": 1, "kind": "reference",
                             "sample": f"{steps} full runs of the shipped example by the reference's own CUDA build (texture-object patch only) on the same GPU; its host side (3 std::sort + orderevents per epoch, file appends) is single-threaded; node has {ncores} cores"},
            "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "reference_counters": {k: runs[-1][k] for k in ("pairs", "epochs", "hits", "events_adder", "events_threshold", "events_deadtime", "singles")},
            "reference_coincidences": coinc, "reference_with_dumps": dumps,
            "reference_times": {"sim_wall_s": [r["sim_wall_s"] for r in runs], "total_wall_s": [r["total_wall_s"] for r in runs],
                                "process_wall_s": [r["process_wall_s"] for r in runs]}}


def reference_coincidences(ex: Path, run, window_us=0.01):
    """The reference has no coincidence sorter: its singles.dat of the last run goes through the oracle's
    sorter (thresholds and dead time switched off, so that the singles pass unchanged), timed separately."""
    try:
        import numpy as np
        from oracle import oracle as orc
        sing = np.fromfile(Path(ex) / "output" / "singles.dat", dtype=orc.EVENT_DTYPE)
        if sing.size == 0:
            return None
        p = orc.DigiParams()
        for k, v in dict(readout_depth=2, readout_policy=1, threshold_eV=0.0, blur_policy=1, blur_Eref=662000.0, blur_Rref=0.0, blur_slope=0.0,
                         blur_space=0.0, dead_level=3, dead_type=0, dead_time_us=0.0, ewin_min=0.0, ewin_max=2.0e6, time_blur_sigma_us=0.0,
                         coinc_window_us=window_us, coinc_policy=0, coinc_min_panel_diff=0, npanels=8, moduleN=117, crystalN=64, seed=1).items():
            setattr(p, k, v)
        t0 = time.perf_counter()
        out, counts, co = orc.digitize(sing, p)
        dt = time.perf_counter() - t0
        if out.size != sing.size:
            return {"note": f"pass-through changed the singles ({sing.size} -> {out.size})"}
        return {"singles_in_file": int(sing.size), "coincidences": int(co.size), "window_us": window_us,
                "per_s_of_simulation_region": co.size / run["sim_wall_s"], "oracle_sorter_s": dt,
                "note": "output/ is cleaned before every run: singles.dat holds the last run only"}
    except Exception as e:  # noqa: BLE001
        return {"note": f"unavailable: {type(e).__name__}: {e}"}


if __name__ == "__main__":
    import sys
    print(json.dumps(run_once(Path(sys.argv[1]), *(sys.argv[2:3])), indent=1))
