#!/usr/bin/env python
"""Per-kernel CUDA-event times of one frame of the bench workload (gpet_profile_enable), warm, averaged over --reps
passes.  Usage (GPU box): python tools/kprof.py [--source source.txt] [--reps 20] [--flush]"""
import argparse
import json
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--source", default="pointsource.txt")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--flush", action="store_true")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--staged", action="store_true", help="staged source/phantom/entry kernels instead of the fused front end")
    a = ap.parse_args()
    import torch
    from gpet_b200 import api
    with tempfile.TemporaryDirectory() as tmp:
        ex = bench.make_workdir(tmp, source=a.source)
        c = api.Context(0)
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        c.set_stream(stream.cuda_stream)
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        c.set_digitizer(coinc_window_us=0.01)
        c.set_coincidence_format(api.Context.COINC_PAIRS)   # as bench.py does
        if a.scale != 1.0:
            src = c.sources()
            for i, s in enumerate(src):
                c.set_source_atoms(i, int(s["natom"] * a.scale))
        c.set_spectrum(128, 0.0, 1.0e6)
        nf = c.plan_frames(0)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

        def frame():
            if a.staged:
                c.stage_source(0); c.stage_phantom(); c.stage_detector()
                c.stage_digitize()
            else:
                c.run_resident()   # the product path: fused front end, digitizer sliced by the frame's time range

        for _ in range(3):
            frame()
        torch.cuda.synchronize()
        c.profile(True)
        for _ in range(a.reps):
            if a.flush:
                flush.fill_(1)
            frame()
        kt = c.kernel_times()
        c.profile(False)
        tot = sum(v[0] for v in kt.values()) / a.reps
        print(f"# {a.source} frames={nf} pairs(frame0)={c.frame_pairs(0)} reps={a.reps} flush={a.flush}: {tot * 1e3:.1f} us of kernel time per frame")
        for k, (ms, n) in kt.items():
            print(f"{ms / a.reps * 1e3:9.2f} us/frame  {n // a.reps:3d} launches  {ms / n * 1e3:8.2f} us each  {k}")
        print(json.dumps({"counts": [int(x) for x in c.last_counts()]}))
        c.close()


if __name__ == "__main__":
    main()
