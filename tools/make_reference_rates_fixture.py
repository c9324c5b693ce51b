#!/usr/bin/env python
"""tests/golden/reference_transport_rates.json from the pin report of the REAL reference (profiles/r01_reference_pin_report.json,
written by tools/ref_pin.py on the GPU box: the CUDA-12-patched reference binary run on the shipped example, blur off).
Per-pair rates of the reference's own counters, summed over its repeat runs -- the known answers the CPU oracle's transport
is held to in tests/test_oracle_units.py.  Usage: python tools/make_reference_rates_fixture.py"""
import json
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def main():
    rep = json.loads((ROOT / "profiles" / "r01_reference_pin_report.json").read_text())
    out = {"_source": "profiles/r01_reference_pin_report.json (tools/ref_pin.py, reference binary oracle/_ref/gPET on a B200, blur off)", "cases": []}
    for t in rep["transport"]:
        runs = t["reference_runs"]
        pairs = sum(r["pairs"] for r in runs)
        case = {"source": t["source"], "window_s": t["window_s"], "reference_pairs": pairs}
        for k in ("hits", "events_adder", "events_threshold", "singles"):
            case[k + "_per_pair"] = sum(r[k] for r in runs) / pairs
        out["cases"].append(case)
    dst = ROOT / "tests" / "golden" / "reference_transport_rates.json"
    dst.write_text(json.dumps(out, indent=1) + "\n")
    print(dst.read_text())


if __name__ == "__main__":
    main()
