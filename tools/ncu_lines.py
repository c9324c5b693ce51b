#!/usr/bin/env python
"""Join the SASS page of an ncu report (instructions executed / thread-instructions per SASS instruction) with the line
table of the shipped cubin (nvdisasm -g): warp-instructions, lane occupancy and stall samples per source line.
Usage: tools/ncu_lines.py report.ncu-rep kernel_substring [cubin-name-substring] [top N]"""
import csv
import io
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def sass_rows(rep, kernel):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kernel, "-c", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ix = {n: i for i, n in enumerate(hdr)}
    out = []
    for r in rows[h + 1:]:
        if len(r) < len(hdr) or r[0] == "Address":
            break
        out.append((r[ix["Source"]], float(r[ix["Instructions Executed"]] or 0), float(r[ix["Thread Instructions Executed"]] or 0),
                    float(r[ix["# Samples"]] or 0)))
    return out


def line_table(kernel, cubin_sub, nsass):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "gpet_b200" / "libgpet_b200.so")], cwd=td, capture_output=True)
        cub = [p for p in Path(td).glob("*.cubin") if cub_match(p.name, cubin_sub)][0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", str(cub)], capture_output=True, text=True).stdout
    # one line list per function whose name contains `kernel` (template instances are separate functions)
    fns, lines, cur, infn = [], None, None, False
    for ln in txt.splitlines():
        if ln.startswith(".text."):
            infn = kernel in ln
            if infn:
                lines = []
                fns.append(lines)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            inl = re.search(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            cur = (Path(m.group(1)).name, int(m.group(2)), int(inl.group(2)) if inl else None)
            continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            lines.append(cur)
    if not fns:
        raise SystemExit(f"no function matching {kernel} in the cubin")
    return min(fns, key=lambda f: abs(len(f) - nsass))


def cub_match(name, sub):
    return sub in name


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    cubin_sub = sys.argv[3] if len(sys.argv) > 3 else "transport"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
    sass = sass_rows(rep, kernel)
    lt = line_table(kernel, cubin_sub, len(sass))
    if len(lt) != len(sass):
        print(f"# warning: {len(sass)} SASS rows in the report vs {len(lt)} in the cubin (rebuilt since the capture?)")
    n = min(len(lt), len(sass))
    tot_i = sum(s[1] for s in sass); tot_t = sum(s[2] for s in sass); tot_s = sum(s[3] for s in sass)
    print(f"# {kernel}: {tot_i:.3g} warp-instructions, {tot_t / tot_i:.1f} threads/instruction, {tot_s:.0f} samples")
    agg = {}
    for k in range(n):
        key = lt[k]
        a = agg.setdefault(key, [0.0, 0.0, 0.0, 0])
        a[0] += sass[k][1]; a[1] += sass[k][2]; a[2] += sass[k][3]; a[3] += 1
    src_cache = {}
    print("%-22s %7s %7s %7s %5s  %s" % ("file:line(inlined at)", "inst%", "thr/in", "smpl%", "sass", "source"))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if key is None:
            continue
        f, l, inl = key
        path = next(iter(ROOT.glob(f"gpet_b200/csrc/{f}")), None)
        text = ""
        if path:
            src_cache.setdefault(path, path.read_text().splitlines())
            text = src_cache[path][l - 1].strip()[:90] if l - 1 < len(src_cache[path]) else ""
        tag = f"{f}:{l}" + (f"({inl})" if inl else "")
        print("%-22s %7.2f %7.1f %7.2f %5d  %s" % (tag, 100 * a[0] / tot_i, a[1] / max(a[0], 1), 100 * a[2] / max(tot_s, 1), a[3], text))


if __name__ == "__main__":
    main()
