#!/usr/bin/env python
"""Summarise an `ncu --set full` report: one line per kernel launch with duration, DRAM traffic, SIMT efficiency, issue
utilisation, occupancy, L2 hit rate and the top warp-stall reasons.  Usage: tools/ncu_summary.py report.ncu-rep [--json out.json]"""
import csv
import io
import json
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rd_MB"), ("dram__bytes_write.sum", "wr_MB"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"), ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("lts__t_sector_hit_rate.pct", "L2hit%"),
        ("l1tex__t_sector_hit_rate.pct", "L1hit%"), ("launch__registers_per_thread", "regs"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        # pipe utilisation: the double-precision time path (t += s / c, the fp64 divide of the panel entry, decay times) runs
        # on the FP64 pipe; north_star asks for its utilisation by name
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%")]


def short(name):
    name = name.split("(")[0].replace("void ", "").replace("unnamed>::", "").replace("gpet::", "").replace("<unnamed>::", "")
    name = name.replace("rsort::", "").replace("<unsigned long long>", "<u64>").replace("<unsigned int>", "<u32>")
    return name[:40]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    out = []
    print("%-40s " % "kernel" + " ".join("%9s" % c[1] for c in COLS) + "  top stalls (warps per issue)")
    for r in data:
        rec = {"kernel": short(r[idx["Kernel Name"]]), "grid": r[idx["Grid Size"]], "block": r[idx["Block Size"]]}
        line = "%-40s " % rec["kernel"]
        for name, label in COLS:
            if name not in idx:
                line += "%9s " % "-"
                continue
            v = float(r[idx[name]].replace(",", "") or 0)
            u = units[idx[name]]
            if label == "us":
                v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
            elif label.endswith("_MB"):
                v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            rec[label] = v
            line += "%9.2f " % v
        stalls = []
        for h in stall_cols:
            try:
                stalls.append((float(r[idx[h]].replace(",", "")), h.split("issue_stalled_")[1].split("_per_")[0]))
            except ValueError:
                pass
        stalls.sort(reverse=True)
        rec["stalls"] = {k: v for v, k in stalls[:4]}
        line += " " + ", ".join("%s %.1f" % (k, v) for v, k in stalls[:4])
        print(line)
        out.append(rec)
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
