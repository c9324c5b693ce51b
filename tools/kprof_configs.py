#!/usr/bin/env python
"""Per-kernel CUDA-event times of BASELINE.json configs 4 (256^3 mouse phantom) and 5 (32-panel ring, 20 cm water) at a
given number of decays, resident frames.  Usage (GPU box): python tools/kprof_configs.py [--decays 4000000]"""
import argparse
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--decays", type=int, default=4_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--case", default="", help="substring of the case name (config4 | config5)")
    a = ap.parse_args()
    import torch
    import test_configs as tc
    from gpet_b200 import api
    from tools import gen_inputs
    cases = {}
    n = 256
    natom = gen_inputs.atoms_for_decays(a.decays, 6586.26, 120.0, 0.97)
    size4 = (3.2, 3.2, 6.4)
    cases["config4 mouse 256^3"] = dict(
        phantom=gen_inputs.mouse_phantom(n, size4), geo_text=None,
        text=gen_inputs.input_file(dims=(n, n, n), offset=tuple(-x / 2 for x in size4), extent=size4, mat="input/phantom_mat.dat",
                                   den="input/phantom_den.dat", source="input/src.txt", blur=(1, 662000, 0.05, 0, 0)),
        src=[(natom, 0, 1, 0, 0, 0, 1.2, 5.0, 0)], dig={})
    cases["config5 ring32 20cm water"] = dict(
        phantom=gen_inputs.water_cylinder_phantom(n, 0.1, 20.0, 20.0), geo_text=gen_inputs.ring_geo(32, 40.0),
        text=gen_inputs.input_file(dims=(n, n, n), offset=(-12.8,) * 3, extent=(25.6,) * 3, mat="input/phantom_mat.dat",
                                   den="input/phantom_den.dat", source="input/src.txt", geo="input/ring.geo", blur=(1, 662000, 0.05, 0, 0)),
        src=[(natom, 0, 1, 0, 0, 0, 0.5, 18.0, 0)], dig=dict(coinc_min_panel_diff=4))
    for name, cs in cases.items():
        if a.case and a.case not in name:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            ex = tc.workdir(Path(tmp), cs["text"], phantom=cs["phantom"], geo_text=cs["geo_text"],
                            extra={"src.txt": gen_inputs.source_file(cs["src"])})
            c = api.Context(0)
            stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); c.set_stream(stream.cuda_stream)
            c.load_config_file(ex / "input_PET.in", base_dir=ex)
            c.set_digitizer(coinc_window_us=0.01, **cs["dig"])
            c.set_coincidence_format(api.Context.COINC_PAIRS)
            nf = c.plan_frames(0)
            for _ in range(2):
                st = c.run_resident()
            torch.cuda.synchronize()
            c.profile(True)
            for _ in range(a.reps):
                st = c.run_resident()
            kt = c.kernel_times()
            c.profile(False)
            tot = sum(v[0] for v in kt.values()) / a.reps
            print(f"# {name}: {st.pairs} pairs in {nf} frames, {tot * 1e3:.1f} us of kernel time per run = {st.pairs / tot / 1e6:.2f} G pairs/s; "
                  f"phantom out {st.photons_phantom_out / (2 * st.pairs):.3f}, on panel {st.photons_on_panel / (2 * st.pairs):.3f}, "
                  f"singles {st.singles}, coincidences {st.coincidences} (trues {st.trues}, scatters {st.scatters}, randoms {st.randoms}; "
                  f"scatter fraction {st.scatters / max(st.trues + st.scatters, 1):.3f})")
            for k, (ms, nl) in kt.items():
                print(f"{ms / a.reps * 1e3:9.2f} us/run  {nl // a.reps:3d} launches  {ms / nl * 1e3:8.2f} us each  {k}")
            c.close()


if __name__ == "__main__":
    main()
