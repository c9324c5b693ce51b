#!/usr/bin/env python
"""Resident and end-to-end time of the bench workload against the number of frames it is split into, with the sum of the
kernel times beside it (per-frame fixed costs and pipeline gaps).  Usage (GPU box): python tools/frames_sweep.py"""
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    import torch
    from gpet_b200 import api
    with tempfile.TemporaryDirectory() as tmp:
        ex = bench.make_workdir(tmp, source="source.txt")
        c = api.Context(0)
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        c.set_stream(stream.cuda_stream)
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        c.set_digitizer(coinc_window_us=0.01)
        c.set_coincidence_format(api.Context.COINC_PAIRS)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        for mp in (0, 1200000, 900000, 800000, 700000, 620000, 590000, 560000, 500000, 450000, 400000):
            nf = c.plan_frames(mp)
            res = {}
            for mode in ("resident", "e2e"):
                ts = []
                for it in range(12):
                    flush.fill_(1)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    st = c.run_resident() if mode == "resident" else c.run(None)
                    torch.cuda.synchronize()
                    ts.append(time.perf_counter() - t0)
                res[mode] = sorted(ts[2:])[len(ts[2:]) // 2]
            c.profile(True)
            for _ in range(5):
                c.run_resident()
            kt = c.kernel_times()
            c.profile(False)
            ksum = sum(v[0] for v in kt.values()) / 5
            print(f"max_pairs={mp:7d} frames={nf:3d} resident {res['resident'] * 1e3:7.3f} ms  e2e {res['e2e'] * 1e3:7.3f} ms  "
                  f"kernel-time sum {ksum:7.3f} ms  ({ksum / nf * 1e3:6.1f} us/frame)")
        c.close()


if __name__ == "__main__":
    main()
