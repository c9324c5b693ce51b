#!/usr/bin/env python
"""End-to-end (gpet_plan_frames + gpet_run, results in pinned host memory) time of the bench workload against the frame
size: smaller frames overlap the D2H copy of frame k with the kernels of frame k+1.  Usage (GPU box): python tools/e2e_sweep.py"""
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
        for mp in (0, 600000, 400000, 300000, 200000, 150000, 100000, 60000):
            ts = []
            for it in range(13):
                flush.fill_(1)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                nf = c.plan_frames(mp)
                st = c.run(None)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            ts = sorted(ts[3:])
            med = ts[len(ts) // 2]
            print(f"max_pairs={mp:8d} frames={nf:3d} e2e median {med * 1e3:7.3f} ms  {st.pairs / med / 1e9:6.3f} G pairs/s  singles={st.singles} coinc={st.coincidences}")
        c.close()


if __name__ == "__main__":
    main()
