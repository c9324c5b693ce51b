#!/usr/bin/env python
"""Kernel time per million pairs against the FRAME SIZE: the shipped example with its activity scaled so that one frame holds
1.1 M x scale pairs (one frame per run, capacities sized for it).  Per-frame fixed costs (tails of the persistent transport
kernels, latency chains of the digitizer kernels) amortise with the frame.  Usage (GPU box): python tools/bigframes_sweep.py"""
import argparse
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scales", default="1,2,4,8,16")
    ap.add_argument("--reps", type=int, default=6)
    a = ap.parse_args()
    import torch
    from gpet_b200 import api
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    with tempfile.TemporaryDirectory() as tmp:
        ex = bench.make_workdir(tmp, source="source.txt")
        for scale in [int(x) for x in a.scales.split(",")]:
            pairs = int(1.25e6 * scale)
            c = api.Context(0)
            stream = torch.cuda.Stream()
            torch.cuda.set_stream(stream)
            c.set_stream(stream.cuda_stream)
            c.set_capacity(2 * pairs, int(1.3 * pairs), int(0.9 * pairs))
            c.load_config_file(ex / "input_PET.in", base_dir=ex)
            c.set_digitizer(coinc_window_us=0.01)
            c.set_coincidence_format(api.Context.COINC_PAIRS)
            c.set_spectrum(128, 0.0, 1.0e6)
            for i, s in enumerate(c.sources()):
                c.set_source_atoms(i, int(s["natom"]) * scale)
            nf = c.plan_frames(pairs)
            for _ in range(2):
                st = c.run_resident()
            torch.cuda.synchronize()
            c.profile(True)
            for _ in range(a.reps):
                flush.fill_(1)
                st = c.run_resident()
            kt = c.kernel_times()
            c.profile(False)
            tot = sum(v[0] for v in kt.values()) / a.reps
            mp = st.pairs / 1e6
            print(f"# scale {scale}: {nf} frame(s), {st.pairs} pairs, {st.singles} singles, {st.coincidences} coincidences: "
                  f"{tot * 1e3:.1f} us of kernels per run = {tot * 1e3 / mp:.1f} us per M pairs = {st.pairs / tot / 1e6:.2f} G pairs/s")
            print("   " + "  ".join(f"{k} {ms / a.reps * 1e3 / mp:.1f}" for k, (ms, n) in kt.items()))
            c.close()


if __name__ == "__main__":
    main()
