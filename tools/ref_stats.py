#!/usr/bin/env python
"""Statistical transport pins from the REAL reference beyond the shipped example (run on a GPU box:
`gpurun -- python tools/ref_stats.py`).

For each named configuration of tools/gen_inputs.py (config1_water10: shipped geometry with a 10 cm water phantom,
config2_psf: psf.dat photon-pair input, config4_mouse: 256^3 water/bone phantom, config5_ring: 32-panel ring with a
20 cm water cylinder) the CUDA-12-patched reference binary (oracle/_ref, built by oracle/build_ref.py) is run at
`--decays` annihilation pairs; its printed counters and its own adder.dat / singles.dat (and, at `--hit-decays`, the
Hits.dat / HitsID.dat of the binary as shipped) are reduced to rates and histograms:

  rates per pair      hits, post-readout events, after thresholder, after dead time, singles, coincidences
                      (the reference has no sorter: its singles.dat through the oracle's sorter, window 10 ns)
  spectra             post-readout event energy (adder.dat, before blur), singles energy (after blur), hit energy
  scatter sensitive   share of post-readout events below 400 keV, per-panel and per-module occupancy, hit-type shares

The same inputs then go through this library (gpet_run) and both sets are written to gpurun_out/ref_stats/<name>.json.
tests/golden/ref_stats/ keeps the reference halves; tests/test_reference_stats.py holds the CUDA path to them
(chi-square / 1 %) on the GPU box, where /root/reference does not exist.
"""
from __future__ import annotations

import argparse
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from tools import gen_inputs  # noqa: E402
from gpet_b200 import refio  # noqa: E402

OUT = ROOT / "gpurun_out" / "ref_stats"
EXAMPLE = ROOT / "examples" / "small_animal"
PACKED = ROOT / "gpet_b200" / "_data" / "input4gPET.gpettab"

E_BINS = np.linspace(0.0, 720000.0, 73)     # 10 keV bins
COINC_WINDOW_US = 0.01


def reduce_events(adder, singles, npanels, moduleN):
    """histograms of one run's post-readout events (adder.dat; None in PSF mode, where the reference does not write it,
    gPET.cu:131) and singles (singles.dat); occupancies from the singles"""
    out = {"n_singles": int(singles.size), "hist_singles_E": np.histogram(singles["E"], E_BINS)[0].tolist(),
           "singles_below_400keV": int((singles["E"] < 400000.0).sum()),
           "panel_occupancy": np.bincount(singles["pann"], minlength=npanels)[:npanels].tolist(),
           "module_occupancy": np.bincount(singles["modn"], minlength=moduleN)[:moduleN].tolist()}
    if adder is not None:
        out.update({"n_adder": int(adder.size), "hist_adder_E": np.histogram(adder["E"], E_BINS)[0].tolist(),
                    "adder_below_400keV": int((adder["E"] < 400000.0).sum())})
    return out


def reduce_hits(ids, f):
    return {"n_hits": int(ids.shape[0]), "hist_hit_E": np.histogram(f[:, 0], E_BINS)[0].tolist(),
            "hit_types": np.bincount(ids[:, 4], minlength=5)[:5].tolist(),
            "hit_x_hist": np.histogram(f[:, 2], np.linspace(-2.0, 0.0, 21))[0].tolist()}


def oracle_coincidences(singles, npanels, moduleN, min_panel_diff=0):
    """the oracle's sorter over a singles list that passes through unchanged (thresholds and dead time off)"""
    from oracle import oracle as orc
    import parity
    p, _ = parity.make_digi_params(threshold_eV=0.0, blur_Rref=0.0, dead_time_us=0.0, ewin_min=0.0, ewin_max=2.0e6,
                                   coinc_window_us=COINC_WINDOW_US, coinc_min_panel_diff=min_panel_diff, npanels=npanels, moduleN=moduleN)
    out, counts, co = orc.digitize(singles, p)
    assert out.size == singles.size
    return int(co.size)


def run_reference(name, cfg, decays, hit_decays, tmp):
    from oracle import run_ref
    ex = gen_inputs.write_workdir(Path(tmp) / f"{name}_ref", cfg, EXAMPLE)
    t0 = time.perf_counter()
    r = run_ref.run_once(ex, "gPET_nodump", timeout=1500)
    if cfg["psf"] is not None:
        r["pairs"] = cfg["psf"].shape[0] // 2   # simulateParticle prints no emission counts (gPET.cu:13-199); every record is run
        r["epochs"] = -(-cfg["psf"].shape[0] // 524288)
    if r["returncode"] != 0 or r["pairs"] <= 0 or not (ex / "output" / "singles.dat").exists():
        return {"error": (r["stdout_tail"][-600:] + r["stderr_tail"])}
    adder = refio.read_events(ex / "output" / "adder.dat") if (ex / "output" / "adder.dat").exists() else None
    singles = refio.read_events(ex / "output" / "singles.dat")
    rep = {"binary": "oracle/_ref/gPET_nodump", "pairs": r["pairs"], "epochs": r["epochs"], "sim_wall_s": r["sim_wall_s"],
           "process_wall_s": time.perf_counter() - t0,
           "counters": {k: r[k] for k in ("hits", "events_adder", "events_threshold", "events_deadtime", "singles")}}
    rep.update(reduce_events(adder, singles, cfg["npanels"], cfg["moduleN"]))
    rep["coincidences"] = oracle_coincidences(singles, cfg["npanels"], cfg["moduleN"], 4 if cfg["npanels"] == 32 else 0)
    del adder, singles
    # hit-level observables from the binary as shipped (OUTPUTHIT = 1), at a smaller count: it writes every hit per epoch
    if hit_decays > 0:
        cfg2 = gen_inputs.stats_config(name, hit_decays)
        ex2 = gen_inputs.write_workdir(Path(tmp) / f"{name}_refhits", cfg2, EXAMPLE)
        r2 = run_ref.run_once(ex2, "gPET", timeout=1500)
        if r2["returncode"] == 0 and (ex2 / "output" / "Hits.dat").exists():
            ids, f = refio.read_hits(ex2 / "output" / "HitsID.dat", ex2 / "output" / "Hits.dat")
            rep["hits_run"] = {"binary": "oracle/_ref/gPET", "pairs": r2["pairs"], "counter_hits": r2["hits"]}
            rep["hits_run"].update(reduce_hits(ids, f))
        else:
            rep["hits_run"] = {"error": r2["stdout_tail"][-300:] + r2["stderr_tail"]}
    return rep


def run_ours(name, cfg, tmp, seed=20260101, min_panel_diff=0):
    """the same files through gpet_run(output_dir): same observables from our adder.dat / singles.dat / Hits.dat"""
    from gpet_b200 import api
    ex = gen_inputs.write_workdir(Path(tmp) / f"{name}_ours", cfg, EXAMPLE, PACKED)
    with api.Context(0) as c:
        c.set_seed(seed)
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        c.set_digitizer(coinc_window_us=COINC_WINDOW_US, coinc_min_panel_diff=min_panel_diff)
        st = c.run(ex / "output")
    adder = refio.read_events(ex / "output" / "adder.dat")
    singles = refio.read_events(ex / "output" / "singles.dat")
    rep = {"pairs": int(st.pairs), "frames": int(st.frames),
           "counters": {"hits": int(st.hits), "events_adder": int(st.events_adder), "events_threshold": int(st.events_threshold),
                        "events_deadtime": int(st.events_deadtime), "singles": int(st.singles)},
           "coincidences": int(st.coincidences), "classes": [int(st.trues), int(st.scatters), int(st.randoms)]}
    rep.update(reduce_events(adder, singles, cfg["npanels"], cfg["moduleN"]))
    ids, f = refio.read_hits(ex / "output" / "HitsID.dat", ex / "output" / "Hits.dat")
    rep["hits_run"] = {"pairs": int(st.pairs)}
    rep["hits_run"].update(reduce_hits(ids, f))
    return rep


def chi2(ha, hb, min_count=20):
    ha, hb = np.asarray(ha, np.float64), np.asarray(hb, np.float64)
    m = (ha + hb) > min_count
    na, nb = ha.sum(), hb.sum()
    if m.sum() < 2 or na == 0 or nb == 0:
        return 0.0, 0
    k1, k2 = np.sqrt(nb / na), np.sqrt(na / nb)
    return float((((k1 * ha[m] - k2 * hb[m]) ** 2) / (ha[m] + hb[m])).sum()), int(m.sum() - 1)


def compare(ref, ours):
    """rates per pair (relative difference, statistical sigma of the difference) and shape chi-squares"""
    out = {"rates": {}, "chi2": {}}
    for k in ("hits", "events_adder", "events_threshold", "events_deadtime", "singles"):
        if not ref["counters"][k]:
            continue
        a, b = ref["counters"][k] / ref["pairs"], ours["counters"][k] / ours["pairs"]
        sig = np.sqrt(1.0 / max(ref["counters"][k], 1) + 1.0 / max(ours["counters"][k], 1))
        out["rates"][k] = {"reference": a, "ours": b, "rel_diff": (b - a) / a if a else None, "stat_sigma_rel": float(sig)}
    a, b = ref["coincidences"] / ref["pairs"], ours["coincidences"] / ours["pairs"]
    out["rates"]["coincidences"] = {"reference": a, "ours": b, "rel_diff": (b - a) / a if a else None,
                                    "stat_sigma_rel": float(np.sqrt(1.0 / max(ref["coincidences"], 1) + 1.0 / max(ours["coincidences"], 1)))}
    for what in ("adder", "singles"):
        if f"n_{what}" in ref and f"n_{what}" in ours:
            a, b = ref[f"{what}_below_400keV"] / ref[f"n_{what}"], ours[f"{what}_below_400keV"] / ours[f"n_{what}"]
            out["rates"][f"{what}_share_below_400keV"] = {"reference": a, "ours": b, "rel_diff": (b - a) / a}
    for k in ("hist_adder_E", "hist_singles_E", "panel_occupancy", "module_occupancy"):
        if k in ref and k in ours:
            c2, ndf = chi2(ref[k], ours[k])
            out["chi2"][k] = {"chi2": c2, "ndf": ndf}
    if "hits_run" in ref and "hist_hit_E" in ref["hits_run"]:
        for k in ("hist_hit_E", "hit_types", "hit_x_hist"):
            c2, ndf = chi2(ref["hits_run"][k], ours["hits_run"][k])
            out["chi2"][k] = {"chi2": c2, "ndf": ndf}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--decays", type=int, default=10_000_000)
    ap.add_argument("--hit-decays", type=int, default=1_000_000)
    ap.add_argument("--ours-decays", type=int, default=4_000_000)
    ap.add_argument("--configs", default=",".join(gen_inputs.STATS_CONFIGS))
    a = ap.parse_args()
    from oracle import run_ref
    OUT.mkdir(parents=True, exist_ok=True)
    if not run_ref.available("gPET_nodump"):
        print("reference binary not available")
        return 1
    for name in a.configs.split(","):
        decays = min(a.decays, 1_000_000) if name == "config2_psf" else a.decays   # config 2 is 1e6 pairs by definition
        with tempfile.TemporaryDirectory() as tmp:
            t0 = time.perf_counter()
            ref = run_reference(name, gen_inputs.stats_config(name, decays), decays, min(a.hit_decays, decays), tmp)
            t1 = time.perf_counter()
            if "error" in ref:
                print(name, "REFERENCE FAILED", ref["error"])
                (OUT / f"{name}.json").write_text(json.dumps({"config": name, "reference": ref}, indent=1))
                continue
            od = min(a.ours_decays, decays)
            ours = run_ours(name, gen_inputs.stats_config(name, od), tmp, min_panel_diff=4 if name == "config5_ring" else 0)
            t2 = time.perf_counter()
        cmp_ = compare(ref, ours)
        doc = {"config": name, "decays_requested": decays, "coinc_window_us": COINC_WINDOW_US, "e_bins_eV": [float(E_BINS[0]), float(E_BINS[-1]), len(E_BINS) - 1],
               "generated_by": "tools/ref_stats.py on a B200 box (reference = oracle/_ref, the unmodified algorithm + CUDA-12 texture-object patch)",
               "reference": ref, "ours_at_generation": ours, "comparison_at_generation": cmp_,
               "seconds": {"reference": t1 - t0, "ours": t2 - t1}}
        (OUT / f"{name}.json").write_text(json.dumps(doc, indent=1))
        print(f"== {name}: reference {ref['pairs']} pairs in {ref['epochs']} epochs ({t1 - t0:.1f} s), ours {ours['pairs']} pairs ({t2 - t1:.1f} s)")
        for k, v in cmp_["rates"].items():
            print(f"   {k:28s} ref {v['reference']:.6f} ours {v['ours']:.6f} rel {100 * (v['rel_diff'] or 0):+.3f} %" +
                  (f" (sigma {100 * v['stat_sigma_rel']:.3f} %)" if "stat_sigma_rel" in v else ""))
        for k, v in cmp_["chi2"].items():
            print(f"   chi2 {k:24s} {v['chi2']:.1f} / {v['ndf']}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
