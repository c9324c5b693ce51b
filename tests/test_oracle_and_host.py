"""CPU suite (-m "not gpu"): pins the oracle on known-answer vectors, checks the host loaders against the independent
numpy parsers, and checks that the C-ABI library loads and exports every declared symbol.  No compute calls."""
import ctypes
import os
from pathlib import Path
import re
import subprocess

import numpy as np
import pytest

import parity
from gpet_b200 import api, refio
from oracle import oracle as orc

REF = parity.ROOT.parent / "reference"


# ------------------------------------------------------------------------------------------------ Philox pin
def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11)
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kats:
        assert [int(x) for x in orc.philox(ctr, key)] == want


# ------------------------------------------------------------------------------------------------ digitizer pins
@pytest.mark.parametrize("case,d,ev", list(parity.kat_cases()), ids=lambda v: v["name"] if isinstance(v, dict) and "name" in v else None)
def test_digitizer_known_answers(case, d, ev):
    p, _ = parity.make_digi_params(**d)
    singles, counts, coinc = orc.digitize(ev, p)
    parity.check_kat(case, singles, counts, coinc)


def test_digitizer_oracle_invariants():
    rng = np.random.default_rng(7)
    ev = parity.random_events(5000, rng, tmax=3.0e4, nsites=50, tie_fraction=0.05, dead_fraction=0.02)
    for dead_type in (0, 1):
        p, _ = parity.make_digi_params(dead_type=dead_type, coinc_window_us=0.5)
        s, counts, co = orc.digitize(ev, p)
        assert counts[0] == ev.size and counts[0] >= counts[1] >= counts[2] >= counts[3] == s.size
        assert np.all(np.diff(s["t"]) >= 0)
        assert np.all((s["E"] >= 30000) & (s["E"] <= 700000))
        # survivors of one site are at least tau apart (early times: fp32 == fp64 here)
        for site in np.unique(s["siten"])[:20]:
            t = s["t"][s["siten"] == site]
            assert np.all(np.diff(t) >= 2.2 - 1e-2)
        assert np.all(co["b"]["t"] - co["a"]["t"] < 0.5) and np.all(co["b"]["t"] >= co["a"]["t"])
        # idempotence: singles are a fixed point of the chain
        s2, c2, _ = orc.digitize(s, p)
        assert parity.events_equal(s, s2)


@pytest.mark.parametrize("dead_type", [0, 1])
@pytest.mark.parametrize("dead_level", [0, 1, 2, 3])
def test_digitizer_oracle_equals_a_literal_walk_of_the_reference_kernels(dead_level, dead_type):
    """The C oracle (closed-form dead time of SURVEY 8a D7, merge sorts) against tests/pyref_digitizer.py: the reference's
    host sequence and kernels followed statement by statement in plain Python, dead-time chains walked literally (every
    walker on the list as it was before any kill -- at early times the reference's own outcome does not depend on how its
    threads interleave; the last case is at t = 1e8 us, where `tdead` in fp32 is coarser than the dead time)."""
    import pyref_digitizer as pyref
    rng = np.random.default_rng(100 + 2 * dead_level + dead_type)
    kills, npairs = 0, [0, 0]
    for n, nsites, tmax, tau, policy, mindiff in [(0, 8, 1e3, 2.2, 0, 0), (1, 8, 1e3, 2.2, 0, 0), (2, 1, 3.0, 2.2, 1, 0),
                                                  (60, 3, 100.0, 2.2, 0, 0), (1200, 40, 900.0, 2.2, 0, 2),
                                                  (1200, 936, 300.0, 0.9, 1, 0), (1200, 936, 400.0, 0.9, 0, 2),
                                                  (1500, 12, 4000.0, 7.5, 1, 3), (1501, 30, 300.0, 2.2, 1, 0)]:
        ev = parity.random_events(n, rng, tmax=tmax, nsites=nsites, tie_fraction=0.05 if n > 10 else 0.0)
        if n == 1501:
            ev["t"] += 1.0e8      # late times: the fp32 `tdead` of the kernel is 8 us coarse, wider than the dead time
        p, d = parity.make_digi_params(dead_level=dead_level, dead_type=dead_type, dead_time_us=tau, coinc_window_us=0.3,
                                       coinc_policy=policy, coinc_min_panel_diff=mindiff)
        s, counts, co = orc.digitize(ev, p)
        ws, wcounts, wpairs = pyref.digitize(ev, d)
        assert [int(c) for c in counts] == wcounts, (n, nsites, counts, wcounts)
        assert s.tobytes() == ws.astype(orc.EVENT_DTYPE).tobytes()
        assert co.size == len(wpairs)
        if wpairs:
            ia, ib = np.array(wpairs).T
            assert co["a"].tobytes() == s[ia].tobytes() and co["b"].tobytes() == s[ib].tobytes()
        kills += wcounts[1] - wcounts[2]
        npairs[policy] += len(wpairs)
    # times on a grid of exactly the window and half the dead time: every comparison of the chain sits on its boundary
    # (t < tdead + tau, t > t_prev + tau, t_b < t_a + W are all strict) and ties abound
    ev = parity.random_events(900, rng, tmax=100.0, nsites=936)
    ev["t"] = 10.0 + 0.25 * rng.integers(0, 500, ev.size)
    ev["E"] = rng.uniform(100e3, 600e3, ev.size).astype(np.float32)
    ev["siten"] = rng.integers(0, 6, ev.size); ev["pann"] = ev["siten"]; ev["modn"] = 0
    for policy in (0, 1):
        p, d = parity.make_digi_params(dead_level=dead_level, dead_type=dead_type, dead_time_us=0.5, coinc_window_us=0.25, coinc_policy=policy)
        s, counts, co = orc.digitize(ev, p)
        ws, wcounts, wpairs = pyref.digitize(ev, d)
        assert [int(c) for c in counts] == wcounts and s.tobytes() == ws.astype(orc.EVENT_DTYPE).tobytes() and co.size == len(wpairs)
        if wpairs:
            ia, ib = np.array(wpairs).T
            assert co["a"].tobytes() == s[ia].tobytes() and co["b"].tobytes() == s[ib].tobytes()
        assert wcounts[1] - wcounts[2] > 50
    # dead time and both sorter policies had work to do (one site for the whole detector leaves no two singles in a window)
    assert kills > 100 and (dead_level == 0 or (npairs[0] > 30 and npairs[1] > 80))


def test_digitizer_oracle_equals_the_literal_walk_on_small_adversarial_lists():
    """Differential fuzz of the same two restatements on thousands of tiny lists built to sit on every edge at once: times
    on a half-microsecond grid (ties, exact window and dead-time boundaries, tau = 0), early and at 1e8 us, energies on the
    window bounds, negative and repeated site numbers, every dead-time level and type, both sorter policies, panel distance."""
    import pyref_digitizer as pyref
    rng = np.random.default_rng(0)
    with_pairs = with_kills = 0
    for trial in range(1500):
        n = int(rng.integers(0, 14))
        ev = np.zeros(n, orc.EVENT_DTYPE)
        ev["parn"] = rng.permutation(n); ev["eventid"] = ev["parn"] // 2
        ev["pann"] = rng.integers(0, 3, n); ev["modn"] = rng.integers(0, 2, n); ev["cryn"] = rng.integers(0, 2, n)
        ev["siten"] = rng.integers(-1, 3, n)
        ev["t"] = rng.integers(0, 12, n) * 0.5 + (1e8 if trial % 5 == 0 else 1.0)
        ev["E"] = rng.choice([40e3, 50e3, 60e3, 300e3, 700e3, 700e3 + 1, 2e6, 2.1e6], n)
        p, d = parity.make_digi_params(dead_level=int(rng.integers(0, 4)), dead_type=int(rng.integers(0, 2)),
                                       dead_time_us=float(rng.choice([0.0, 0.5, 1.0, 2.2])), coinc_window_us=float(rng.choice([0.25, 0.5, 1.0])),
                                       coinc_policy=int(rng.integers(0, 2)), coinc_min_panel_diff=int(rng.integers(0, 3)), npanels=3,
                                       moduleN=2, crystalN=2, threshold_eV=50e3, ewin_min=55e3, ewin_max=700e3)
        s, counts, co = orc.digitize(ev, p)
        ws, wcounts, wpairs = pyref.digitize(ev, d)
        assert [int(c) for c in counts] == wcounts and s.tobytes() == ws.astype(orc.EVENT_DTYPE).tobytes() and co.size == len(wpairs), (trial, d)
        if wpairs:
            ia, ib = np.array(wpairs).T
            assert co["a"].tobytes() == s[ia].tobytes() and co["b"].tobytes() == s[ib].tobytes(), (trial, d)
        with_pairs += bool(wpairs); with_kills += wcounts[1] > wcounts[2]
    assert with_pairs > 100 and with_kills > 100


# ------------------------------------------------------------------------------------------------ C ABI surface
def test_noise_oracle_is_a_poisson_process_over_the_detector():
    # addnoise (gPET_kernals.cu:699-735): mean gap 2 us over 0.1 s -> 50 000 +- 224 arrivals, uniform sites, E ~ N(300 keV, 20 keV)
    ev = orc.noise(5.0e4, 1.5e5, 2.0, 3.0e5, 2.0e4, 1000.0, 8, 117, 64, 99)
    assert abs(ev.size - 50000) < 5 * np.sqrt(50000)
    assert ev["t"].min() >= 5.0e4 and ev["t"].max() < 1.5e5 and np.all(np.diff(ev["t"]) > 0)
    gaps = np.diff(ev["t"])
    assert abs(gaps.mean() - 2.0) < 0.05 and abs(gaps.std() - 2.0) < 0.1          # exponential gaps
    assert abs(ev["E"].mean() - 3.0e5) < 5 * 2.0e4 / np.sqrt(ev.size) and abs(ev["E"].std() - 2.0e4) < 500
    assert np.all(ev["parn"] == -1) and np.all(ev["eventid"] < 0) and np.unique(ev["eventid"]).size == ev.size
    assert ev["pann"].min() == 0 and ev["pann"].max() == 7 and ev["modn"].max() == 116 and ev["cryn"].max() == 63
    assert np.array_equal(ev["siten"], (ev["pann"] * 117 + ev["modn"]) * 64 + ev["cryn"])
    assert abs(np.bincount(ev["pann"], minlength=8) - ev.size / 8).max() < 5 * np.sqrt(ev.size / 8)
    assert ev["x"].min() > 0 and ev["x"].max() <= 1
    # windows tile: the events of [a, c) are those of [a, b) followed by those of [b, c)
    a = orc.noise(5.0e4, 9.99e4, 2.0, 3.0e5, 2.0e4, 1000.0, 8, 117, 64, 99)
    b = orc.noise(9.99e4, 1.5e5, 2.0, 3.0e5, 2.0e4, 1000.0, 8, 117, 64, 99)
    assert np.concatenate([a, b]).tobytes() == ev.tobytes()
    assert orc.noise(0, 1e5, 0.0, 3e5, 2e4, 1000.0, 8, 117, 64, 99).size == 0      # disabled


def test_coincidence_class_oracle_known_answers():
    """orc_classify by hand: random = different annihilation or a noise single, scatter = same annihilation with a
    photon of the scattered list, true = the rest; the pair shift groups a photon PSF's records 2k, 2k+1."""
    def co(rows):
        c = np.zeros(len(rows), orc.COINC_DTYPE)
        for i, (pa, ea, pb, eb) in enumerate(rows):
            c[i]["a"]["parn"], c[i]["a"]["eventid"], c[i]["b"]["parn"], c[i]["b"]["eventid"] = pa, ea, pb, eb
        return c
    rows = [(10, 5, 11, 5), (10, 5, 13, 6), (20, 10, 21, 10), (21, 10, 20, 10), (-1, -5, 30, 15), (30, 15, 31, 15),
            (40, 20, -1, 20), (-2147483648, 7, 2147483647, 7)]
    cls, tot = orc.classify(co(rows), [21, 99, -2147483648], 0)
    assert cls.tolist() == [0, 2, 1, 1, 2, 0, 2, 1] and tot.tolist() == [2, 3, 3]
    cls, tot = orc.classify(co(rows), [], 0)
    assert cls.tolist() == [0, 2, 0, 0, 2, 0, 2, 0] and tot.tolist() == [5, 0, 3]
    # photon PSF: eventid = record index, records 2k and 2k+1 are one pair
    rows = [(4, 4, 5, 5), (5, 5, 6, 6), (6, 6, 7, 7)]
    cls, tot = orc.classify(co(rows), [7], 1)
    assert cls.tolist() == [0, 2, 1] and tot.tolist() == [1, 1, 1]
    assert orc.classify(co([]), [1, 2], 0)[1].tolist() == [0, 0, 0]
    # against plain numpy on a random list
    rng = np.random.default_rng(3)
    c = np.zeros(5000, orc.COINC_DTYPE)
    for side in "ab":
        c[side]["parn"] = rng.integers(-1, 400, c.size)
        c[side]["eventid"] = c[side]["parn"] >> 1
    scattered = rng.choice(400, 60, replace=False)
    cls, tot = orc.classify(c, scattered, 0)
    rnd = (c["a"]["parn"] == -1) | (c["b"]["parn"] == -1) | (c["a"]["eventid"] != c["b"]["eventid"])
    sc = np.isin(c["a"]["parn"], scattered) | np.isin(c["b"]["parn"], scattered)
    assert np.array_equal(cls, np.where(rnd, 2, np.where(sc, 1, 0))) and tot.tolist() == np.bincount(cls, minlength=3).tolist()


def test_library_exports_every_declared_symbol():
    header = (parity.ROOT / "include" / "gpet_b200.h").read_text()
    declared = set(re.findall(r"\b(gpet_[a-z_0-9]+)\s*\(", header))
    declared -= {"gpet_ctx"}
    assert len(declared) > 40
    l = ctypes.CDLL(str(api.LIB_PATH))
    for name in declared:
        assert hasattr(l, name), f"{name} declared in gpet_b200.h but not exported"
    assert set(api._SIGS) == declared, declared ^ set(api._SIGS)
    assert api.lib().gpet_abi_version() == api.ABI_VERSION == 5


def test_missing_or_stale_library_fails_loudly(monkeypatch, tmp_path):
    """No CUDA library, no product: the Python mirror raises instead of computing anything itself, and refuses a library
    built from another header version."""
    monkeypatch.setattr(api, "_lib", None)
    monkeypatch.setattr(api, "LIB_PATH", tmp_path / "libgpet_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        api.Context(0)
    monkeypatch.undo()
    monkeypatch.setattr(api, "_lib", None)
    monkeypatch.setattr(api, "ABI_VERSION", api.ABI_VERSION + 1)
    with pytest.raises(RuntimeError, match="ABI version"):
        api.lib()


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under gpet_b200/ (Python or C++/CUDA) names it, the library does not link
    it, and bench.py reaches it only inside its baseline legs."""
    import ast
    for f in (parity.ROOT / "gpet_b200" / "csrc").glob("*"):
        includes = [l for l in f.read_text().splitlines() if l.lstrip().startswith("#include")]
        assert not any("oracle" in l for l in includes), f
    rules = [l for l in (parity.ROOT / "Makefile").read_text().splitlines() if not l.lstrip().startswith("#")]
    assert not any("oracle" in l for l in rules)                      # the product's build never reaches into oracle/
    for f in (parity.ROOT / "gpet_b200").glob("*.py"):
        for node in ast.walk(ast.parse(f.read_text())):
            names = [a.name for a in node.names] if isinstance(node, ast.Import) else [node.module or ""] if isinstance(node, ast.ImportFrom) else []
            assert not any(n.split(".")[0] == "oracle" for n in names), f
    deps = subprocess.run(["ldd", str(api.LIB_PATH)], capture_output=True, text=True).stdout
    assert "oracle" not in deps
    tree = ast.parse((parity.ROOT / "bench.py").read_text())
    importing = set()
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            if isinstance(node, ast.ImportFrom) and (node.module or "").split(".")[0] == "oracle":
                importing.add(fn.name)
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    assert not any((getattr(n, "module", "") or "").startswith("oracle") for n in top)
    assert importing <= {"cpu_baseline", "reference_line"}, importing


def test_record_layouts():
    assert api.EVENT_DTYPE.itemsize == 48 and api.EVENT_DTYPE.fields["t"][1] == 24 and api.EVENT_DTYPE.fields["E"][1] == 32
    assert api.COINC_DTYPE.itemsize == 96 and api.HIT_DTYPE.itemsize == 48 and api.PHOTON_DTYPE.itemsize == 48


def test_ctypes_mirror_has_the_headers_struct_layouts(tmp_path):
    """Every struct that crosses the boundary, field by field: sizeof / offsetof from the C header (a C99 program built here)
    against the ctypes structures and numpy record types of gpet_b200/api.py."""
    structs = {"gpet_digitizer_params": api.DigitizerParams, "gpet_transport_params": api.TransportParams, "gpet_stats": api.Stats}
    records = {"gpet_event": api.EVENT_DTYPE, "gpet_coincidence": api.COINC_DTYPE, "gpet_hit": api.HIT_DTYPE,
               "gpet_photon": api.PHOTON_DTYPE, "gpet_panel": api.PANEL_DTYPE, "gpet_single_compact": api.COMPACT_DTYPE}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "gpet_b200.h"', "int main(void) {"]
    for name, st in structs.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        lines += [f'printf("{name}.{f} %zu\\n", offsetof({name}, {f}));' for f, _ in st._fields_]
    for name, dt in records.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        lines += [f'printf("{name}.{f} %zu\\n", offsetof({name}, {f}));' for f in dt.names]
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", f"-I{parity.ROOT / 'include'}", str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, st in structs.items():
        assert int(got[name]) == ctypes.sizeof(st), name
        for f, _ in st._fields_:
            assert int(got[f"{name}.{f}"]) == getattr(st, f).offset, (name, f)
    for name, dt in records.items():
        assert int(got[name]) == dt.itemsize, name
        for f in dt.names:
            assert int(got[f"{name}.{f}"]) == dt.fields[f][1], (name, f)
    # and nothing in the header's structs is missing from the mirror
    header = (parity.ROOT / "include" / "gpet_b200.h").read_text()
    for name, st in structs.items():
        body = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + ";", header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        declared = [re.sub(r"\[.*", "", t).strip() for decl in body.split(";") for t in decl.split(",") if decl.strip()]
        declared = [d.split()[-1] for d in declared if d]
        assert declared == [f for f, _ in st._fields_], (name, declared)


def test_host_only_context_refuses_compute():
    with api.Context(device=-1) as c:
        with pytest.raises(api.GpetError) as e:
            c.stage_digitize()
        assert e.value.code == -4
        with pytest.raises(api.GpetError):
            c.digitize(np.zeros(4, api.EVENT_DTYPE))
        with pytest.raises(api.GpetError):
            c.run()
        # every compute entry point, not only the big ones: no CPU fallback anywhere
        calls = [lambda: c.stage_source(0), lambda: c.stage_psf(0, 0), lambda: c.stage_phantom(), lambda: c.stage_detector(),
                 lambda: c.stage_front(-1), lambda: c.stage_panel_transport(), lambda: c.stage_noise(0.0, 1.0),
                 lambda: c.put_photons(0, np.zeros(1, api.PHOTON_DTYPE)), lambda: c.fetch_photons(0), lambda: c.put_events(np.zeros(1, api.EVENT_DTYPE)),
                 lambda: c.fetch_events(4), lambda: c.fetch_hits(4), lambda: c.fetch_singles(4), lambda: c.fetch_coincidences(4),
                 lambda: c.fetch_coincidence_classes(4), lambda: c.mark_scattered([1, 2]), lambda: c.last_counts(), lambda: c.run_resident(),
                 lambda: c.queue_size(0)]
        for k, f in enumerate(calls):
            with pytest.raises(api.GpetError) as e:
                f()
            assert e.value.code == -4, k
        # results of a run that never happened are empty, not an error
        assert c.result_singles().size == 0 and c.result_coincidences().size == 0 and c.result_coincidence_classes().size == 0
        assert c.get_digitizer().coinc_pair_shift == 0
        c.set_digitizer(coinc_pair_shift=1)
        assert c.get_digitizer().coinc_pair_shift == 1


@pytest.mark.parametrize("dead_level,rdepth,rpolicy", [(0, 2, 1), (1, 2, 1), (2, 2, 1), (3, 2, 1), (3, 3, 0), (3, 1, 0), (3, 2, 0)])
def test_compact_singles_expand_on_the_host(dead_level, rdepth, rpolicy):
    """gpet_expand_singles (host only): 32-byte compact singles -> the 48-byte Event.  The packing is restated here in numpy
    (include/gpet_b200.h: ids = pann | modn << 8 | cryn << 20 | (parn & 1) << 31), siten follows from the dead-time level --
    or, at level 3, from the readout level the detector kernel used -- and the ids; parn from eventid and the photon bit."""
    rng = np.random.default_rng(11)
    n = 5000
    with api.Context(device=-1) as c:
        c.load_geometry(parity.EXAMPLE / "input" / "config8.geo")
        cnt, _, _ = c.geometry_counts()
        module_n, crystal_n = int(cnt[2]), int(cnt[3])
        assert (module_n, crystal_n) == (117, 64)
        c.set_digitizer(dead_level=dead_level, readout_depth=rdepth, readout_policy=rpolicy)
        ev = np.zeros(n, api.EVENT_DTYPE)
        ev["pann"] = rng.integers(0, 8, n); ev["modn"] = rng.integers(0, module_n, n); ev["cryn"] = rng.integers(0, crystal_n, n)
        pair = rng.integers(0, 1 << 40, n, dtype=np.int64)                 # 64-bit history numbers: the record keeps 31 bits
        which = rng.integers(0, 2, n)
        ev["eventid"] = (pair & 0x7fffffff).astype(np.int32)
        ev["parn"] = ((2 * pair + which) & 0x7fffffff).astype(np.int32)
        depth = 2 if (rdepth != 3 and rpolicy == 1) else rdepth
        level = depth if dead_level == 3 else dead_level
        site = {0: np.zeros(n, np.int64), 1: ev["pann"].astype(np.int64), 2: ev["pann"].astype(np.int64) * module_n + ev["modn"],
                3: (ev["pann"].astype(np.int64) * module_n + ev["modn"]) * crystal_n + ev["cryn"]}[level]
        ev["siten"] = site
        ev["t"] = np.sort(rng.uniform(0, 1e8, n)); ev["E"] = rng.uniform(3e4, 7e5, n)
        ev["x"], ev["y"], ev["z"] = rng.uniform(-2, 0, n), rng.uniform(-8, 8, n), rng.uniform(-11, 11, n)
        cs = np.zeros(n, api.COMPACT_DTYPE)
        for f in ("t", "E", "x", "y", "z", "eventid"):
            cs[f] = ev[f]
        cs["ids"] = (ev["pann"].astype(np.uint32) | (ev["modn"].astype(np.uint32) << 8) | (ev["cryn"].astype(np.uint32) << 20)
                     | ((ev["parn"].astype(np.uint32) & 1) << 31))
        out = c.expand_singles(cs)
        assert out.tobytes() == ev.tobytes()
        assert c.expand_singles(cs[:0]).size == 0
    with api.Context(device=-1) as c2:   # no geometry: refused, not guessed
        with pytest.raises(api.GpetError):
            c2.expand_singles(cs[:4])
        with pytest.raises(api.GpetError):
            c2.set_singles_format(7)
        with pytest.raises(api.GpetError):
            c2.result_singles_compact()


# ------------------------------------------------------------------------------------------------ loaders
def test_config_parser_matches_numpy_parser(tmp_path):
    cfg = refio.parse_config(parity.EXAMPLE / "input_PET.in")
    assert cfg["pdim"] == [200, 200, 200] and cfg["rdepth"] == 2 and cfg["rpolicy"] == 1
    assert cfg["dlevel"] == 3 and cfg["dtype"] == 0 and abs(cfg["dtime"] - 2.2) < 1e-6
    assert cfg["nsurface"] == 1 and cfg["surface"][-1] == 1 and cfg["sourcefile"] == "input/pointsource.txt"
    # product loader on the same file (phantom written small to keep it quick)
    ex = tmp_path / "ex"
    (ex / "input").mkdir(parents=True)
    (ex / "data").mkdir()
    text = (parity.EXAMPLE / "input_PET.in").read_text().replace("200 200 200", "16 16 16")
    (ex / "input_PET.in").write_text(text)
    for f in ("config8.geo", "pointsource.txt", "source.txt"):
        (ex / "input" / f).write_text((parity.EXAMPLE / "input" / f).read_text())
    (ex / "data" / "isotopes.txt").write_text((parity.EXAMPLE / "data" / "isotopes.txt").read_text())
    mat, den = parity.gen_inputs.cylinder_phantom(n=16)
    parity.gen_inputs.write_phantom(mat, den, ex / "input" / "cylinder_phantom_mat.dat", ex / "input" / "cylinder_phantom_den.dat")
    if not parity.have_tables():
        pytest.skip("packed tables not built")
    (ex / "data" / "input4gPET.gpettab").write_bytes(parity.PACKED.read_bytes())
    with api.Context(device=-1) as c:
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        d = c.get_digitizer()
        assert (d.readout_depth, d.readout_policy, d.dead_level, d.dead_type) == (2, 1, 3, 0)
        assert d.threshold_eV == 50000 and d.ewin_min == 30000 and d.ewin_max == 700000
        assert abs(d.blur_Rref - 0.05) < 1e-7 and d.blur_policy == 1 and d.blur_Eref == 662000
        t = c.get_transport()
        assert abs(t.noncollinearity_rad - 0.0037056) < 1e-9 and t.eabs_eV == 1000 and t.nsurface == 1
        assert len(c.sources()) == 1 and c.sources()[0]["natom"] == 1762974000 and c.sources()[0]["type"] == 3
        assert len(c.isotopes()) == 4 and abs(c.isotopes()[0]["halftime"] - 6586.26) < 1e-2
        assert c.plan_frames(0) >= 1


def test_geometry_loader_matches_numpy_parser():
    geo = parity.EXAMPLE / "input" / "config8.geo"
    panels, mat, dens, counts = refio.parse_geometry(geo)
    assert list(counts) == [9, 8, 117, 64]  # moduleN=117, crystalN=64
    with api.Context(device=-1) as c:
        c.load_geometry(geo)
        got = c.panels()
        c4, m2, d2 = c.geometry_counts()
    assert list(c4) == list(counts) and list(m2) == list(mat) and np.allclose(d2, dens)
    assert got.size == 8
    for f in api.PANEL_FIELDS:
        assert np.allclose(got[f], panels[f], atol=2e-6), f
    # panel 2 = panel 0 rotated by 90 degrees about z: offset (0,-22.5,0) -> (22.5, 0, 0)
    assert np.allclose([got["offsetx"][2], got["offsety"][2]], [22.5, 0.0], atol=1e-4)


def test_source_and_isotope_loaders():
    with api.Context(device=-1) as c:
        c.load_isotopes(parity.EXAMPLE / "data" / "isotopes.txt")
        c.load_source(parity.EXAMPLE / "input" / "source.txt")
        src, iso = c.sources(), c.isotopes()
    ref_src = refio.parse_sources(parity.EXAMPLE / "input" / "source.txt")
    ref_iso = refio.parse_isotopes(parity.EXAMPLE / "data" / "isotopes.txt")
    assert len(src) == 7 and len(iso) == 4
    for a, b in zip(src, ref_src):
        assert a["natom"] == b["natom"] and a["type"] == b["type"] and a["shape"] == b["shape"]
        assert np.array_equal(a["coeff"], b["coeff"])
    for a, b in zip(iso, ref_iso):
        assert a["halftime"] == b["halftime"] and a["ratio"] == b["ratio"] and np.array_equal(a["coef"], b["coef"])


def test_loader_errors_are_reported_not_fatal(tmp_path):
    with api.Context(device=-1) as c:
        with pytest.raises(api.GpetError) as e:
            c.load_geometry(tmp_path / "missing.geo")
        assert e.value.code == -2
        bad = tmp_path / "bad.in"
        bad.write_text("label\n0\nlabel\nnot-a-number\n")
        with pytest.raises(api.GpetError):
            c.load_config_file(bad)
        with pytest.raises(api.GpetError):
            c.load_tables(tmp_path / "nothing")


def test_parsers_survive_truncated_and_corrupted_files(tmp_path):
    # the reference exits (or reads garbage) on malformed input (CUDA_CALL / FILEEXIST macros, unchecked fscanf); the
    # loaders here must report an error or accept the file -- never crash
    rng = np.random.default_rng(0)
    texts = {"in": (parity.EXAMPLE / "input_PET.in").read_text(), "geo": (parity.EXAMPLE / "input" / "config8.geo").read_text(),
             "src": (parity.EXAMPLE / "input" / "source.txt").read_text(), "iso": (parity.EXAMPLE / "data" / "isotopes.txt").read_text()}
    outcomes = {k: [0, 0] for k in texts}
    for kind, text in texts.items():
        for trial in range(120):
            data = bytearray(text.encode())
            if trial % 2 == 0:
                data = data[: int(rng.integers(0, len(data)))]
            else:
                for _ in range(int(rng.integers(1, 12))):
                    data[int(rng.integers(0, len(data)))] = int(rng.integers(0, 256))
            f = tmp_path / f"{kind}.txt"
            f.write_bytes(bytes(data))
            with api.Context(-1) as c:
                try:
                    if kind == "in":
                        c.load_config_file(f, base_dir=parity.EXAMPLE)
                    elif kind == "geo":
                        c.load_geometry(f)
                        assert 0 <= c.panels().size <= 100000
                    elif kind == "iso":
                        c.load_isotopes(f)
                    else:
                        c.load_isotopes(parity.EXAMPLE / "data" / "isotopes.txt")
                        c.load_source(f)
                    outcomes[kind][0] += 1
                except api.GpetError as e:
                    assert e.code in (-1, -2, -6) and str(e)
                    outcomes[kind][1] += 1
    assert all(err > 0 for _, err in outcomes.values()), outcomes


def test_psf_loader(tmp_path):
    rec = parity.gen_inputs.back_to_back_psf(1000)
    rec.tofile(tmp_path / "psf.dat")
    with api.Context(device=-1) as c:
        c.load_psf(tmp_path / "psf.dat", 0, 1)
        assert c.num_psf() == 2000
        c.load_psf(tmp_path / "psf.dat", 500, 1)
        assert c.num_psf() == 500  # readParticle clamps to the request (initialize.cu:88-92)
    back = refio.read_psf(tmp_path / "psf.dat")
    assert back.shape == (2000, 8) and np.allclose(back[0::2, 4:7], -back[1::2, 4:7])


@pytest.mark.skipif(not (REF / "data" / "input4gPET.lamph").exists(), reason="reference checkout not present")
def test_table_loader_matches_numpy_parser_on_reference_files(tmp_path):
    # the complete CTD set ships with the reference: ASCII loader (header-driven dims) vs numpy parser
    t = refio.read_tables(REF / "data" / "input4gCTD")
    with api.Context(device=-1) as c:
        c.load_tables(REF / "data" / "input4gCTD")
        d = c.table_dims()
        assert (d["nmat"], d["nen"], d["cm_ncp"], d["cm_ne"]) == (7, 2048, 101, 51)
        assert np.array_equal(c.table(0).reshape(7, 2048), t["lamph"])
        assert np.array_equal(c.table(1).reshape(7, 2048), t["compt"])
        assert np.array_equal(c.table(3).reshape(7, 2048), t["rayle"])
        assert np.array_equal(c.table(4).reshape(7, 101, 51), t["cmpsf"]["surf"])
        assert np.array_equal(c.table(5).reshape(7, 101, 51), t["rayff"]["surf"])
        # total = compton + rayleigh + photo, which is why the transport never reads phote
        tot = c.table(1) + c.table(2) + c.table(3)
        assert np.allclose(tot, c.table(0), rtol=3e-4)
        # packed round trip
        c.save_tables_packed(tmp_path / "ctd.gpettab")
    with api.Context(device=-1) as c2:
        c2.load_tables(tmp_path / "ctd.gpettab")
        assert np.array_equal(c2.table(4).reshape(7, 101, 51), t["cmpsf"]["surf"])


@pytest.mark.skipif(not (REF / "data" / "input4gCTD.cmpsf").exists(), reason="reference checkout not present")
def test_surface_generator_reproduces_shipped_ctd_surfaces():
    from tools import gen_tables as g
    ctd = refio.read_matter(REF / "data" / "input4gCTD.matter")
    cm = refio.read_surface(REF / "data" / "input4gCTD.cmpsf", ctd["nmat"])
    rl = refio.read_surface(REF / "data" / "input4gCTD.rayff", ctd["nmat"])
    k = ctd["names"].index("Water")
    sq = cm["sq"][k].astype(np.float64); fq = rl["sq"][k].astype(np.float64)
    s = g.compton_surface(sq[:, 0], sq[:, 2], cm["ncp"], cm["ne"], float(cm["de"]), ncos=4001)
    assert np.abs(s - cm["surf"][k]).max() < 0.012   # grid quantisation of the original generator
    r = g.rayleigh_surface(fq[:, 0], fq[:, 2], rl["ncp"], rl["ne"], float(rl["de"]), ncos=4001)
    assert np.abs(r - rl["surf"][k])[:, :2].max() < 0.004  # low energies: exact; higher: the shipped table is coarse


@pytest.mark.skipif(not (REF / "data" / "input4gCTD.cmpsf").exists(), reason="reference checkout not present")
def test_mixing_rule_recovers_elements_and_reproduces_shipped_compound_blocks():
    """tools/gen_tables.py builds the S(q) / F(q) of the PET materials that have no shipped block (CorticalBone, Pb, the
    tissues) by the additivity rule from elemental functions recovered from the shipped compounds.  The rule
    is checked on what IS shipped: the recovered H, C, O are atoms (S -> Z at large q, F(0)^2 = Z^2), and TissueICRP --
    thirteen elements, N recovered from air, nine minor ones Thomas-Fermi scaled -- comes out of the mix within 1 %."""
    from tools import gen_tables as g
    ctd = refio.read_matter(REF / "data" / "input4gCTD.matter")
    cm = refio.read_surface(REF / "data" / "input4gCTD.cmpsf", ctd["nmat"])
    rl = refio.read_surface(REF / "data" / "input4gCTD.rayff", ctd["nmat"])
    q = cm["sq"][0][:, 0].astype(np.float64)
    S = {n: cm["sq"][i][:, 2].astype(np.float64) for i, n in enumerate(ctd["names"])}
    F2 = {n: rl["sq"][i][:, 2].astype(np.float64) ** 2 for i, n in enumerate(ctd["names"])}
    comp = g.read_compositions(REF / "data" / "input4gCTD.matter")
    assert comp["Water"] == [(1, 2.0), (8, 1.0)] and comp["LSO"] == [(71, 2.0), (14, 1.0), (8, 5.0)]
    el = g.elemental_functions(q, S, F2, comp)
    for z in (1, 6, 7, 8, 71):
        assert abs(el[z][0][-1] - z) < 0.02 * z and abs(np.sqrt(el[z][1][0]) - z) < 0.02 * z, z      # S(inf) = Z, F(0) = Z
        assert np.all(np.diff(el[z][0]) > -5e-3 * z)                                              # S grows with q (to the 6 digits of the files)
    s, f = g.mix(el, comp["TissueICRP"])
    F = np.sqrt(F2["TissueICRP"])
    assert np.max(np.abs(s / S["TissueICRP"] - 1.0)) < 0.01
    big = F2["TissueICRP"] > 1e-3 * F2["TissueICRP"][0]
    assert np.max(np.abs(f[big] / F[big] - 1.0)) < 0.01
    # the compounds the elements were solved from come back exactly
    for name in ("Water", "PE", "PMMA", "DryAir", "LSO", "LYSO"):
        s, f = g.mix(el, comp[name])
        # (up to the clipping of the 6-digit noise in the far tails, where the solved elemental values would dip below zero)
        assert np.allclose(s, S[name], rtol=1e-5, atol=1e-5 * S[name][-1]) and np.allclose(f * f, F2[name], rtol=1e-5, atol=1e-5 * F2[name][0]), name
    # what the PET set gets: shipped blocks verbatim, the rest mixed, bone and lead flagged as Thomas-Fermi scaled
    blocks, pet = g.material_blocks(REF / "data")
    how = {n: h for n, _, _, h in blocks}
    assert [n for n, *_ in blocks] == list(pet["names"]) and how["Water"] == how["LSO"] == "shipped block"
    assert "Thomas-Fermi" in how["CorticalBone"] and "20" in how["CorticalBone"] and "82" in how["Pb"]
    bone = dict((n, (a, b)) for n, a, b, _ in blocks)["CorticalBone"]
    assert abs(bone[0][-1, 2] - 5.288227) < 1e-3          # S -> total Z per molecule of the matter file


@pytest.mark.skipif(not parity.have_tables(), reason="packed tables not built")
def test_packed_pet_tables_are_sane():
    with api.Context(device=-1) as c:
        c.load_tables(parity.PACKED)
        d = c.table_dims()
        assert (d["nmat"], d["nen"], d["cm_ncp"], d["cm_ne"], d["rl_ncp"], d["rl_ne"]) == (10, 4096, 301, 151, 301, 151)
        e = c.table(8)
        lam = c.table(0).reshape(10, 4096)
        i511 = int(round((511000 - e[0]) / (e[-1] - e[0]) * 4095))
        assert abs(lam[1, i511] - 0.09595) < 2e-4     # water at 511 keV
        assert abs(lam[7, i511] - 0.11674) < 2e-4     # LSO
        cm = c.table(4).reshape(10, 301, 151)
        assert np.all(np.diff(cm, axis=1) >= -1e-6) and np.all(cm[:, 0] == -1) and np.all(cm[:, -1] == 1)


# ------------------------------------------------------------------------------------------------ planner
def test_frame_planner_statistics_and_invariance():
    with api.Context(device=-1) as c:
        c.load_isotopes(parity.EXAMPLE / "data" / "isotopes.txt")
        c.load_source(parity.EXAMPLE / "input" / "source.txt")
        c.set_time_window(0, 120)
        c.set_capacity(1 << 19, 1 << 20, 1 << 20)
        nf = c.plan_frames(0)
        pairs = [c.frame_pairs(f) for f in range(nf)]
        total = sum(pairs)
        # 91 875 000 F-18 atoms, 120 s, branching 0.97 -> 1.118e6 pairs
        expect = 91875000 * (1 - 2 ** (-120 / 6586.26)) * 0.97
        assert abs(total - expect) < 6 * np.sqrt(expect)
        assert max(pairs) <= (1 << 19) // 2 and nf >= 5
        # same seed -> same plan; other seed -> other counts
        assert c.plan_frames(0) == nf and [c.frame_pairs(f) for f in range(nf)] == pairs
        c.set_seed(1234)
        c.plan_frames(0)
        assert [c.frame_pairs(f) for f in range(nf)] != pairs


@pytest.mark.parametrize("geo,phantom_half,sources", [
    ("config8", 0.5, [(0, 0, 0.0, 1.5)]),                                   # shipped ring, 1 cm phantom, sources inside 1.5 cm
    ("ring32", 10.0, [(6, 5, -4.0, 4.2), (-7, 3, 6.0, 1.5), (0, -8, 0.0, 3.4)]),   # 32 panels at 30 cm, big phantom, off-centre sources
])
def test_direction_table_never_drops_a_panel_the_exact_test_accepts(tmp_path, geo, phantom_half, sources):
    """gpet_run narrows the panel search with a direction table.  Brute force on the CPU: random
    lines through the reference sphere, the reference's acceptance test (gPET_kernals.cu:966-1007) in float64 for every
    panel, and the accepting panel must be in the table cell of the line's direction."""
    from tools import gen_inputs
    geo_path = parity.EXAMPLE / "input" / "config8.geo"
    if geo == "ring32":
        geo_path = tmp_path / "ring.geo"
        geo_path.write_text(gen_inputs.ring_geo(32, 30.0))
    n = 8
    mat = np.zeros((n, n, n), np.int32); den = np.ones((n, n, n), np.float32)
    src = tmp_path / "src.txt"
    src.write_text(f"{len(sources)}\nheader#\n" + "".join(f"1000 0 2 {x} {y} {z} {r} 0 0\n" for x, y, z, r in sources))
    with api.Context(-1) as c:
        c.load_geometry(geo_path)
        c.set_phantom(mat, den, np.full(3, -phantom_half, np.float32), np.full(3, 2 * phantom_half, np.float32))
        c.load_isotopes(parity.EXAMPLE / "data" / "isotopes.txt")
        c.load_source(src)
        tab, ref = c.direction_table()
        panels = c.panels()
    assert tab is not None and tab.shape == (32, 32, 32)
    assert ref[3] >= np.sqrt(3) * phantom_half and all(np.hypot(np.hypot(x, y), z) + r <= ref[3] for x, y, z, r in sources)
    rng = np.random.default_rng(7)
    m = 400000
    q = rng.normal(size=(m, 3)); q *= (ref[3] * rng.random(m) ** (1 / 3) / np.linalg.norm(q, axis=1))[:, None]; q += ref[:3]   # in the sphere
    v = rng.normal(size=(m, 3)); v /= np.linalg.norm(v, axis=1)[:, None]
    v32 = v.astype(np.float32)
    # the photon may sit anywhere on its line when the search runs (overshoot position): slide it along v
    pos = q + v * rng.uniform(-40, 40, m)[:, None]
    cell = np.clip(np.floor((v32.astype(np.float64) + 1.0) * 16.0).astype(int), 0, 31)
    mask = tab[cell[:, 2], cell[:, 1], cell[:, 0]]
    accepted_any = np.zeros(m, bool)
    for i, p in enumerate(panels):
        ux = np.array([p["UniXx"], p["UniXy"], p["UniXz"]], np.float64); uy = np.array([p["UniYx"], p["UniYy"], p["UniYz"]], np.float64)
        uz = np.array([p["UniZx"], p["UniZy"], p["UniZz"]], np.float64); o = np.array([p["offsetx"], p["offsety"], p["offsetz"]], np.float64)
        lvx = v @ ux
        r = pos - o
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (r @ ux) / lvx
            y2 = r @ uy - t * (v @ uy); z2 = r @ uz - t * (v @ uz)
        ok = (lvx * p["directionx"] >= 0) & (np.abs(y2) < p["lengthy"] / 2) & (np.abs(z2) < p["lengthz"] / 2)
        dropped = ok & ((mask >> np.uint32(i)) & 1 == 0)
        assert not dropped.any(), (geo, i, int(dropped.sum()))
        accepted_any |= ok
    assert accepted_any.mean() > 0.1                       # the test really exercised accepting panels
    uniq, inv = np.unique(mask, return_inverse=True)
    pop = np.array([bin(int(x)).count("1") for x in uniq])[inv]
    # and the table really narrows the search (directions along the ring axis see every panel of the 32-panel ring)
    assert (pop.max() <= 4 and np.median(pop) <= 3) if geo == "config8" else np.median(pop) <= 18


def test_direction_table_reference_sphere_follows_the_inputs(tmp_path):
    # the table is built for (phantom box + source shapes | PSF records); it must widen with them and switch off when a
    # photon's line is not bounded (positron range) or the panels do not fit a 32-bit mask
    n = 8
    mat = np.zeros((n, n, n), np.int32); den = np.ones((n, n, n), np.float32)
    src = tmp_path / "src.txt"
    src.write_text("1\nheader#\n1000 0 1 3 0 0 0.5 2 0\n")          # cylinder r 0.5, h 2 at x = 3
    with api.Context(-1) as c:
        assert c.direction_table() == (None, None)                        # nothing loaded
        c.load_geometry(parity.EXAMPLE / "input" / "config8.geo")
        c.set_phantom(mat, den, np.full(3, -0.5, np.float32), np.full(3, 1.0, np.float32))
        _, ref0 = c.direction_table()
        assert abs(ref0[3] - (np.sqrt(3) / 2 * 1.001 + 1e-3)) < 1e-6 and np.allclose(ref0[:3], 0)
        c.load_isotopes(parity.EXAMPLE / "data" / "isotopes.txt")
        c.load_source(src)
        _, ref1 = c.direction_table()
        assert abs(ref1[3] - ((3 + np.sqrt(0.25 + 1.0)) * 1.001 + 1e-3)) < 1e-5
        c.set_transport(use_positron_range=1)
        assert c.direction_table() == (None, None)
        c.set_transport(use_positron_range=0)
        refio.write_psf(tmp_path / "psf.dat", np.array([0.0, 7.0]), np.array([0.0, 0.0]), np.array([0.0, -2.0]), np.array([1.0, 2.0]),
                        np.array([1.0, 0.0]), np.array([0.0, 1.0]), np.array([0.0, 0.0]), np.array([511e3, 511e3]))
        c.load_psf(tmp_path / "psf.dat", 0, 1)
        _, ref2 = c.direction_table()
        assert abs(ref2[3] - (np.sqrt(49 + 4) * 1.001 + 1e-3)) < 1e-5      # PSF mode: the records, not the sources
    from tools import gen_inputs
    big = tmp_path / "ring40.geo"
    big.write_text(gen_inputs.ring_geo(40, 50.0))
    with api.Context(-1) as c:
        c.load_geometry(big)
        c.set_phantom(mat, den, np.full(3, -0.5, np.float32), np.full(3, 1.0, np.float32))
        assert c.direction_table() == (None, None)                        # 40 panels do not fit the mask


def test_c_example_compiles_as_c99_and_refuses_to_compute_without_a_device(tmp_path):
    # the header is plain C (no C++/torch types in the signatures): examples/c/digitize_replay.c builds with gcc -std=c99
    exe = tmp_path / "digitize_replay"
    libdir = parity.ROOT / "gpet_b200"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{parity.ROOT / 'include'}",
                    str(parity.ROOT / "examples" / "c" / "digitize_replay.c"), f"-L{libdir}", "-lgpet_b200", f"-Wl,-rpath,{libdir}",
                    "-o", str(exe)], check=True)
    parity.random_events(500, np.random.default_rng(3)).tofile(tmp_path / "adder.dat")
    r = subprocess.run([str(exe), str(parity.EXAMPLE / "input" / "config8.geo"), str(tmp_path / "adder.dat"), str(tmp_path / "singles.dat"), "-1"],
                       capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr and not (tmp_path / "singles.dat").exists()


def test_c_run_example_compiles_as_c99_and_refuses_to_compute_without_a_device(tmp_path):
    # examples/c/run_classify.c: the whole-run entry points and the coincidence classes from plain C; on a host-only
    # context the loaders work (the shipped example parses) and gpet_run refuses
    exe = tmp_path / "run_classify"
    libdir = parity.ROOT / "gpet_b200"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{parity.ROOT / 'include'}",
                    str(parity.ROOT / "examples" / "c" / "run_classify.c"), f"-L{libdir}", "-lgpet_b200", f"-Wl,-rpath,{libdir}",
                    "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)], capture_output=True).returncode == 2
    if not parity.have_tables():
        pytest.skip("packed tables not built")
    ex = tmp_path / "ex"
    (ex / "input").mkdir(parents=True); (ex / "data").mkdir()
    (ex / "input_PET.in").write_text((parity.EXAMPLE / "input_PET.in").read_text().replace("200 200 200", "16 16 16"))
    for f in ("config8.geo", "pointsource.txt"):
        (ex / "input" / f).write_text((parity.EXAMPLE / "input" / f).read_text())
    (ex / "data" / "isotopes.txt").write_text((parity.EXAMPLE / "data" / "isotopes.txt").read_text())
    (ex / "data" / "input4gPET.gpettab").write_bytes(parity.PACKED.read_bytes())
    mat, den = parity.gen_inputs.cylinder_phantom(n=16)
    parity.gen_inputs.write_phantom(mat, den, ex / "input" / "cylinder_phantom_mat.dat", ex / "input" / "cylinder_phantom_den.dat")
    r = subprocess.run([str(exe), "input_PET.in", "0.01", "-1"], cwd=ex, capture_output=True, text=True)
    assert r.returncode == 3 and "gpet_run failed" in r.stderr and "no CPU fallback" in r.stderr, r.stderr


def test_cli_rejects_missing_argument():
    exe = parity.ROOT / "bin" / "gpet_b200"
    if not exe.exists():
        pytest.skip("CLI not built")
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 1 and "input_file" in r.stdout  # main.cu:28-33




@pytest.mark.skipif(not (REF / "input_PET.in").exists(), reason="the reference checkout is not on this box (GPU box): the CPU suite runs where it is")
def test_reference_input_files_verbatim(tmp_path):
    """The reference's OWN input_PET.in, input/config8.geo, input/source.txt, input/pointsource.txt and data/isotopes.txt,
    read where they lie (never copied): the product's loaders must parse them, and the relabelled copies under
    examples/small_animal/ must carry exactly the same values."""
    ref_cfg = refio.parse_config(REF / "input_PET.in")
    ex_cfg = refio.parse_config(parity.EXAMPLE / "input_PET.in")
    assert ref_cfg == ex_cfg
    # the product's config parser on the reference's file: every field, via a work directory that links the reference's files
    ex = tmp_path / "ex"
    (ex / "input").mkdir(parents=True); (ex / "data").mkdir()
    os.symlink(REF / "input_PET.in", ex / "input_PET.in")
    for f in ("config8.geo", "pointsource.txt", "source.txt"):
        os.symlink(REF / "input" / f, ex / "input" / f)
    os.symlink(REF / "data" / "isotopes.txt", ex / "data" / "isotopes.txt")
    mat, den = parity.gen_inputs.cylinder_phantom(n=200)        # the phantom blobs are not shipped
    parity.gen_inputs.write_phantom(mat, den, ex / ref_cfg["matfile"], ex / ref_cfg["denfile"])
    with api.Context(device=-1) as c, api.Context(device=-1) as e:
        for ctx_, root in ((c, REF), (e, parity.EXAMPLE)):
            ctx_.load_geometry(root / "input" / "config8.geo")
            ctx_.load_isotopes(root / "data" / "isotopes.txt")
        assert c.panels().tobytes() == e.panels().tobytes() and c.panels().size == 8
        assert [list(x) if hasattr(x, "__len__") else x for x in c.geometry_counts()[0]] == [9, 8, 117, 64]
        for a, b in zip(c.isotopes(), e.isotopes()):
            assert a["halftime"] == b["halftime"] and a["ratio"] == b["ratio"] and np.array_equal(a["coef"], b["coef"])
        for name in ("source.txt", "pointsource.txt"):
            c.load_source(REF / "input" / name); e.load_source(parity.EXAMPLE / "input" / name)
            assert len(c.sources()) == len(e.sources()) > 0
            for a, b in zip(c.sources(), e.sources()):
                assert a["natom"] == b["natom"] and a["type"] == b["type"] and a["shape"] == b["shape"] and np.array_equal(a["coeff"], b["coeff"])
        if parity.have_tables():
            os.symlink(parity.PACKED, ex / "data" / "input4gPET.gpettab")
            c.load_config_file(ex / "input_PET.in", base_dir=ex)           # the whole init chain on the reference's file
            d, t = c.get_digitizer(), c.get_transport()
            assert (d.readout_depth, d.readout_policy, d.dead_level, d.dead_type) == (ref_cfg["rdepth"], ref_cfg["rpolicy"], ref_cfg["dlevel"], ref_cfg["dtype"])
            assert d.threshold_eV == ref_cfg["Eth"] and d.ewin_min == ref_cfg["Ewinmin"] and d.ewin_max == ref_cfg["Ewinmax"]
            assert abs(d.dead_time_us - ref_cfg["dtime"]) < 1e-6 and abs(d.blur_Rref - ref_cfg["Rref"]) < 1e-7
            assert abs(t.noncollinearity_rad - ref_cfg["nonangle"]) < 1e-9 and t.eabs_eV == ref_cfg["eabsph"]
            assert c.sources()[0]["natom"] == 1762974000          # input_PET.in points at pointsource.txt
            assert c.plan_frames(0) >= 1
        assert api.lib().gpet_peek_config_device(str(REF / "input_PET.in").encode()) == ref_cfg.get("device", 0)


def test_division_by_a_shared_reciprocal_is_the_ieee_quotient():
    """k_detector forms the three centroid quotients of an adder / readout merge from ONE correctly rounded reciprocal
    (q = RN(a * rcp); RN(q + (a - b q) * rcp), transport.cu div_rcp) and crystalSearch's four divisions from reciprocals
    that come with the panel.  The identity with IEEE division, emulated here in exact arithmetic, on the operand ranges
    that occur: energies 1e3 .. 2e6 eV, coordinates of a few cm, the module and crystal pitches of the panel files."""
    rng = np.random.default_rng(0)
    n = 4_000_000

    def via_rcp(a, b, rcp):
        q0 = (a * rcp).astype(np.float32)
        rem = (a.astype(np.float64) - b.astype(np.float64) * q0.astype(np.float64)).astype(np.float32)      # fma: one rounding
        return (q0.astype(np.float64) + rem.astype(np.float64) * rcp.astype(np.float64)).astype(np.float32)

    es = rng.uniform(1.0e3, 2.2e6, n).astype(np.float32)
    a = (rng.uniform(-25, 25, n).astype(np.float32) * es * rng.uniform(0.05, 1.0, n).astype(np.float32)).astype(np.float32)
    rcp = (np.float32(1.0) / es).astype(np.float32)
    assert np.array_equal(via_rcp(a, es, rcp), (a / es).astype(np.float32))
    for pitch in (np.float32(1.75) + np.float32(0.02), np.float32(0.21) + np.float32(0.01), np.float32(3.2) + np.float32(0.05)):
        d = np.full(n, pitch, np.float32)
        r = np.full(n, np.float32(np.longdouble(1.0) / np.longdouble(pitch)), np.float32)
        y = rng.uniform(0.0, 30.0, n).astype(np.float32)
        k = np.arange(0, 130, dtype=np.float32) * pitch
        # multiples of the pitch and their neighbours (y = half the panel width + a coordinate: zero or a normal number, never denormal)
        y[:389] = np.concatenate([k, np.nextafter(k[1:], np.float32(1e9)), np.nextafter(k[1:], np.float32(-1e9)), [np.float32(0.0)]])
        assert np.array_equal(via_rcp(y, d, r), (y / d).astype(np.float32))
