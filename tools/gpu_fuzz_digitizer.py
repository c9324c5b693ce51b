#!/usr/bin/env python
"""Differential fuzz of the CUDA digitizer (through the C ABI) against the CPU oracle on thousands of tiny adversarial event
lists -- the GPU twin of tests/test_oracle_and_host.py::test_digitizer_oracle_equals_the_literal_walk_on_small_adversarial_lists
(times on a half-microsecond grid: ties, exact window and dead-time boundaries, tau = 0; early and at 1e8 us; energies on the
window bounds; negative and repeated site numbers; every dead-time level and type; both sorter policies; panel distance).
Written when no GPU time was left in round 1: run it on a B200 first (`gpurun -- python tools/gpu_fuzz_digitizer.py`) and turn
it into a `-m gpu` test once it is green.  Prints the first mismatches and a summary; exit code 1 on any mismatch."""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=3000)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    import parity
    from gpet_b200 import api
    from oracle import oracle as orc
    rng = np.random.default_rng(a.seed)
    bad = 0
    with api.Context(0) as c:
        c.load_geometry(parity.EXAMPLE / "input" / "config8.geo")      # moduleN = 117, crystalN = 64 of setSitenum
        for trial in range(a.trials):
            n = int(rng.integers(0, 14))
            ev = np.zeros(n, api.EVENT_DTYPE)
            ev["parn"] = rng.permutation(n); ev["eventid"] = ev["parn"] // 2
            ev["pann"] = rng.integers(0, 3, n); ev["modn"] = rng.integers(0, 2, n); ev["cryn"] = rng.integers(0, 2, n)
            ev["siten"] = rng.integers(-1, 3, n)
            ev["t"] = rng.integers(0, 12, n) * 0.5 + (1e8 if trial % 5 == 0 else 1.0)
            ev["E"] = rng.choice([40e3, 50e3, 60e3, 300e3, 700e3, 700e3 + 1, 2e6, 2.1e6], n)
            p, d = parity.make_digi_params(dead_level=int(rng.integers(0, 4)), dead_type=int(rng.integers(0, 2)),
                                           dead_time_us=float(rng.choice([0.0, 0.5, 1.0, 2.2])),
                                           coinc_window_us=float(rng.choice([0.25, 0.5, 1.0])), coinc_policy=int(rng.integers(0, 2)),
                                           coinc_min_panel_diff=int(rng.integers(0, 3)), threshold_eV=50e3, ewin_min=55e3, ewin_max=700e3)
            parity.apply_digi_params(c, d)
            got, counts = c.digitize(ev)
            co = c.fetch_coincidences()
            want, wcounts, wco = orc.digitize(ev, p)
            ok = (list(counts) == list(wcounts) and got.tobytes() == want.astype(api.EVENT_DTYPE).tobytes()
                  and co.tobytes() == wco.astype(api.COINC_DTYPE).tobytes())
            if not ok:
                bad += 1
                if bad <= 5:
                    print("MISMATCH trial", trial, {k: d[k] for k in ("dead_level", "dead_type", "dead_time_us", "coinc_window_us", "coinc_policy", "coinc_min_panel_diff")})
                    print("  counts", list(counts), list(wcounts), "coincidences", co.size, wco.size)
                    print("  events", ev[["parn", "pann", "siten", "t", "E"]].tolist())
    print(f"{a.trials} lists, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
