#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: kernels of one step (between two k_source
launches) with their device times and shares.  Usage: tools/launch_summary.py gpurun_out/launches.csv [step_index]"""
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row["Metric Name"] != "gpu__time_duration.sum":   # lists taken with DRAM byte counts carry three rows per launch
            continue
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Unit"] in ("ns", "nsecond"):
            v /= 1000.0
        elif row["Metric Unit"] in ("ms", "msecond"):
            v *= 1000.0
        rows.append((row["Kernel Name"], v))
    return rows


def short(name):
    name = name.replace("gpet::<unnamed>::", "").replace("unnamed>::", "").replace("gpet::rsort::", "").replace("rsort::", "").replace("void ", "")
    return name.split("(")[0][:60]


def main():
    rows = load(sys.argv[1])
    marker = sys.argv[3] if len(sys.argv) > 3 else "k_front"
    idx = [i for i, x in enumerate(rows) if marker in x[0]]
    k = int(sys.argv[2]) if len(sys.argv) > 2 else max(0, len(idx) - 3)
    a, b = idx[k], idx[k + 1] if k + 1 < len(idx) else len(rows)
    step = [r for r in rows[a:b] if "at::" not in r[0]]
    tot = sum(v for _, v in step)
    print(f"# step {k}: {len(step)} launches, {tot:.1f} us of kernel time (cold-cache, serialised under ncu)")
    for n, v in step:
        print(f"{v:9.2f} us {100 * v / tot:5.1f}%  {short(n)}")


if __name__ == "__main__":
    main()
