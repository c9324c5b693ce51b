"""BASELINE.json configs[1..4] as GPU parity cases (configs[0] is the bench workload and test_gpu_parity's whole-path
tests): each runs through the public C-ABI call (gpet_load_config_file + gpet_run) on the files a gPET user would
write, is compared with the CPU oracle at a size the oracle finishes in seconds, and is checked at (or near) the named
size through size-independent properties (replay bit-exactness of adder.dat -> singles.dat, sortedness, conservation of
counts, coincidences inside the window)."""
import numpy as np
import pytest

import parity
from gpet_b200 import api, refio
from oracle import oracle as orc
from tools import gen_inputs

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not parity.have_tables(), reason="packed tables not built")]


def workdir(tmp_path, text, phantom=None, geo_text=None, extra=None):
    ex = tmp_path / "ex"
    (ex / "input").mkdir(parents=True); (ex / "data").mkdir(); (ex / "output").mkdir()
    (ex / "input_PET.in").write_text(text)
    for f in ("config8.geo", "pointsource.txt", "source.txt"):
        (ex / "input" / f).write_text((parity.EXAMPLE / "input" / f).read_text())
    (ex / "data" / "isotopes.txt").write_text((parity.EXAMPLE / "data" / "isotopes.txt").read_text())
    (ex / "data" / "input4gPET.gpettab").symlink_to(parity.PACKED)
    if phantom is not None:
        gen_inputs.write_phantom(phantom[0], phantom[1], ex / "input" / "phantom_mat.dat", ex / "input" / "phantom_den.dat")
    if geo_text is not None:
        (ex / "input" / "ring.geo").write_text(geo_text)
    for name, content in (extra or {}).items():
        (ex / "input" / name).write_text(content)
    return ex


def replay_is_bit_exact(ex, st, **digi):
    """adder.dat of the run through the ORACLE digitizer must give the run's singles.dat byte for byte (blur off)."""
    adder = refio.read_events(ex / "output" / "adder.dat")
    sing = refio.read_events(ex / "output" / "singles.dat")
    assert adder.size == st.events_adder and sing.size == st.singles
    p, _ = parity.make_digi_params(tie_site=1, **digi)   # events of a run: equal times ordered by site number
    want, wcounts, wco = orc.digitize(adder, p)
    assert want.astype(api.EVENT_DTYPE).tobytes() == sing.tobytes()
    return adder, sing, wco


# ------------------------------------------------------------------------------------------------ config 2: photon PSF
def test_config2_photon_psf_full_size(tmp_path):
    """2e6 photons = 1e6 back-to-back pairs, t_k = (k+1) us: detector transport + digitizer."""
    npairs = 1_000_000
    mat, den = gen_inputs.air_phantom(32)
    text = gen_inputs.input_file(dims=(32, 32, 32), mat="input/phantom_mat.dat", den="input/phantom_den.dat", usepsf=1,
                                 source="input/psf.dat", ptype=1, nhist=2 * npairs, blur=(1, 662000, 0.0, 0, 0))
    ex = workdir(tmp_path, text, phantom=(mat, den))
    gen_inputs.back_to_back_psf(npairs).tofile(ex / "input" / "psf.dat")
    with api.Context(0) as c:
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        assert c.num_psf() == 2 * npairs
        c.set_digitizer(coinc_window_us=0.01)
        st = c.run(ex / "output")
    assert st.pairs == npairs and st.frames == 1
    # air phantom: (almost) every photon leaves it; ~0.41 of them cross a panel face
    assert st.photons_phantom_out > 0.999 * 2 * npairs
    assert 0.35 < st.photons_on_panel / (2 * npairs) < 0.47
    adder, sing, wco = replay_is_bit_exact(ex, st, coinc_window_us=0.01)
    assert np.all(np.diff(sing["t"]) >= 0)
    co = refio.read_coincidences(ex / "output" / "coincidences.dat")
    assert co.size == st.coincidences and co.tobytes() == wco.astype(api.COINC_DTYPE).tobytes()
    # pairs are 1 us apart and the window is 10 ns: every coincidence is a true pair, eventid>>1 = pair index
    assert np.array_equal(co["a"]["eventid"] >> 1, co["b"]["eventid"] >> 1)
    # back-to-back geometry: opposite panels (cyclic difference 4 of 8); the rest are the two photons in neighbours of
    # the opposite panel, or one photon seen in two modules of the same panel (no minimum panel difference is set here)
    d = np.abs(co["a"]["pann"] - co["b"]["pann"]); d = np.minimum(d, 8 - d)
    assert (d == 4).mean() > 0.85
    same_photon = co["a"]["eventid"] == co["b"]["eventid"]
    # (a few hundred of the 2e6 photons scatter in the 1 cm of air and lose their partner's direction)
    assert (d[~same_photon] >= 3).mean() > 0.999 and np.all(d[same_photon] <= 1)
    # event time = pair time + flight (22.5 cm / c = 0.75 ns) + transport in the crystal
    k = adder["eventid"] >> 1
    dt = adder["t"] - (k + 1.0)
    assert dt.min() > 22.5 / 29979.2458 * 0.999 and dt.max() < 0.01


def test_config2_photon_psf_matches_oracle_per_photon(tmp_path):
    mat, den = gen_inputs.air_phantom(16)
    s = parity.Setup(0, phantom=(mat, den), size=1.0)
    rec = gen_inputs.back_to_back_psf(60000)
    refio.write_psf(tmp_path / "psf.dat", *(rec[:, i] for i in range(8)))
    c = s.ctx
    c.load_psf(tmp_path / "psf.dat", 0, ptype=1)
    c.stage_psf(0, c.num_psf())
    ph = c.fetch_photons(0)
    assert ph.size == 120000 and np.array_equal(ph["eventid"], np.arange(120000))   # eventid = record index (initialize.cu:103)
    c.stage_phantom(); c.stage_detector()
    ev = c.fetch_events()
    oph = orc.phantom(ph, s.mat, s.den, s.offset, s.size, s.tab_ph, s.eabs, s.seed)
    res = orc.detector(oph, s.panels, s.counts4, s.pmat, s.pdens, s.surfaces, s.tab_det, s.eabs, 2, 1, s.seed)
    oev = res["events"]
    assert abs(ev.size - oev.size) <= 0.003 * oev.size
    ea = {(int(r["parn"]), int(r["siten"])): float(r["E"]) for r in ev}
    eb = {(int(r["parn"]), int(r["siten"])): float(r["E"]) for r in oev}
    ce = set(ea) & set(eb)
    assert len(ce) >= 0.99 * len(eb)
    assert sum(abs(ea[k] - eb[k]) <= 2e-3 * eb[k] for k in ce) >= 0.99 * len(ce)
    s.close()


# ------------------------------------------------------------------------------------------------ config 3: sensitivity
def test_config3_point_source_sensitivity_sweep(tmp_path):
    """F-18 point source in air, energy-window sweep: singles/decay and coincidences/decay fall
    monotonically with the lower bound, and the photopeak fraction agrees with the oracle within 1 %."""
    decays = 400_000
    natom = gen_inputs.atoms_for_decays(decays, 6586.26, 120.0, 0.97)
    src = gen_inputs.source_file([(natom, 0, 2, 0, 0, 0, 0.03, 0, 0)])
    mat, den = gen_inputs.air_phantom(32)
    sens, coin = [], []
    lows = [250e3, 350e3, 400e3, 450e3]
    ref_events = None
    for lo in lows:
        text = gen_inputs.input_file(dims=(32, 32, 32), mat="input/phantom_mat.dat", den="input/phantom_den.dat",
                                     source="input/f18_point.txt", ewin=(lo, 650e3))
        ex = workdir(tmp_path / f"w{int(lo)}", text, phantom=(mat, den), extra={"f18_point.txt": src})
        with api.Context(0) as c:
            c.load_config_file(ex / "input_PET.in", base_dir=ex)
            c.set_digitizer(coinc_window_us=0.01)
            st = c.run(None)
            if ref_events is None:
                ref_events = st.events_adder
            assert st.events_adder == ref_events          # same seed: transport identical, only the window changes
            assert abs(st.pairs - decays) < 6 * np.sqrt(decays)
            sens.append(st.singles / st.pairs); coin.append(st.coincidences / st.pairs)
            # coincidence classes: nothing to scatter in (air, 6e-5 interactions per photon), ~1700 singles/s -> no randoms to speak of
            assert st.trues + st.scatters + st.randoms == st.coincidences
            assert st.scatters <= 0.01 * st.coincidences and st.randoms <= 0.01 * st.coincidences
    assert all(a > b for a, b in zip(sens, sens[1:])) and all(a > b for a, b in zip(coin, coin[1:]))
    assert 0.05 < coin[0] < 0.2 and 0.2 < sens[0] < 0.7
    # oracle at a matched count with another seed: fraction of post-readout events inside each window
    s = parity.Setup(0, phantom=(mat, den), size=1.0, seed=4242)
    ph = parity.isotropic_photons(300000, np.random.default_rng(5))
    res = orc.detector(ph, s.panels, s.counts4, s.pmat, s.pdens, s.surfaces, s.tab_det, s.eabs, 2, 1, 999)
    s.ctx.put_photons(1, ph); s.ctx.stage_detector()
    ev = s.ctx.fetch_events()
    for lo in lows:
        fg = ((ev["E"] >= lo) & (ev["E"] <= 650e3)).sum() / ph.size
        fo = ((res["events"]["E"] >= lo) & (res["events"]["E"] <= 650e3)).sum() / ph.size
        assert abs(fg - fo) <= 0.01 * fo + 3 * np.sqrt(fo / ph.size), (lo, fg, fo)
    s.close()


# ------------------------------------------------------------------------------------------------ config 4: mouse phantom
def test_config4_mouse_phantom_256(tmp_path):
    """256^3 water/bone phantom of 3.2 x 3.2 x 6.4 cm, distributed F-18 cylinder source: several
    frames through gpet_run at the full grid size; spectra against the oracle at a matched smaller count."""
    n = 256
    size = (3.2, 3.2, 6.4)
    mat, den = gen_inputs.mouse_phantom(n, size)
    decays = 3_000_000
    natom = gen_inputs.atoms_for_decays(decays, 6586.26, 120.0, 0.97)
    src_rows = [(natom, 0, 1, 0, 0, 0, 1.2, 5.0, 0)]
    text = gen_inputs.input_file(dims=(n, n, n), offset=tuple(-x / 2 for x in size), extent=size, mat="input/phantom_mat.dat",
                                 den="input/phantom_den.dat", source="input/mouse_src.txt", blur=(1, 662000, 0.0, 0, 0))
    ex = workdir(tmp_path, text, phantom=(mat, den), extra={"mouse_src.txt": gen_inputs.source_file(src_rows)})
    with api.Context(0) as c:
        c.set_seed(31337)
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        c.set_digitizer(coinc_window_us=0.01)
        c.plan_frames(1 << 20)                # ~1 Mi pairs per frame: 3 frames
        st = c.run(None)
        singles = c.result_singles(); co = c.result_coincidences(); cls = c.result_coincidence_classes()
        # one frame staged by hand for the comparison below
        c.stage_source(0)
        q0 = c.fetch_photons(0)[:300000]
    assert st.frames >= 3 and abs(st.pairs - decays) < 6 * np.sqrt(decays)
    assert st.singles == singles.size and st.coincidences == co.size > 0
    # coincidence classes: ~13 % of the photons scatter in the mouse (checked per photon below), so roughly a fifth of the
    # same-annihilation coincidences hold a scattered photon, less what the energy window removes
    assert st.trues + st.scatters + st.randoms == st.coincidences and np.array_equal(np.bincount(cls, minlength=3), [st.trues, st.scatters, st.randoms])
    assert np.array_equal(cls == 2, co["a"]["eventid"] != co["b"]["eventid"])
    assert 0.03 < st.scatters / (st.trues + st.scatters) < 0.45
    assert st.events_adder >= st.events_threshold >= st.events_deadtime >= st.singles
    # frames are consecutive time slices: the concatenated singles are globally time ordered
    assert np.all(np.diff(singles["t"]) >= 0)
    assert np.all(co["b"]["t"] - co["a"]["t"] < 0.01) and np.all(co["b"]["t"] >= co["a"]["t"])
    # water + bone, ~1.4 cm: photo-absorption in the phantom is a 1e-4 effect at 511 keV (scatter is checked below)
    absorbed = 1.0 - st.photons_phantom_out / (2.0 * st.pairs)
    assert 1e-6 < absorbed < 0.01
    # per-photon phantom parity on the first photons of frame 0 (bone voxels included), then spectra with another seed
    s = parity.Setup(0, phantom=(mat, den), size=size, seed=31337)
    s.ctx.put_photons(0, q0); s.ctx.stage_phantom()
    got = s.ctx.fetch_photons(1)
    want = orc.phantom(q0, s.mat, s.den, s.offset, s.size, s.tab_ph, s.eabs, s.seed)
    want = want[want["t"] > 0]
    ncommon, nmatch, only_g, only_o = parity.compare_photons(got, want)
    assert only_g + only_o <= 0.004 * want.size and nmatch >= 0.995 * ncommon
    assert 0.08 < (got["nscat"] > 0).mean() < 0.4      # ~1.4 cm of water: ~13 % scatter
    s.ctx.stage_detector()
    ev = s.ctx.fetch_events()
    oph = orc.phantom(q0, s.mat, s.den, s.offset, s.size, s.tab_ph, s.eabs, 777)
    res = orc.detector(oph, s.panels, s.counts4, s.pmat, s.pdens, s.surfaces, s.tab_det, s.eabs, 2, 1, 777)
    assert abs(ev.size - res["events"].size) <= 0.01 * res["events"].size + 3 * np.sqrt(res["events"].size)
    chi2, ndf = parity.chi2_hist(ev["E"], res["events"]["E"], np.linspace(0, 520000, 53))
    assert ndf > 20 and chi2 / ndf < 1.6, (chi2, ndf)
    s.close()


# ------------------------------------------------------------------------------------------------ config 5: 32-panel ring
def test_config5_ring_of_32_panels_with_20cm_water(tmp_path):
    """32 panels on a 40 cm ring, 256^3 phantom at 0.1 cm voxels holding a 20 cm water cylinder, F-18 line-like source
    (SURVEY 8d config 5): scatter fraction from the phantom-scatter flag, parity against the oracle."""
    n = 256
    mat, den = gen_inputs.water_cylinder_phantom(n, 0.1, 20.0, 20.0)
    size = (25.6, 25.6, 25.6)
    decays = 2_000_000
    natom = gen_inputs.atoms_for_decays(decays, 6586.26, 120.0, 0.97)
    src_rows = [(natom, 0, 1, 0, 0, 0, 0.5, 18.0, 0)]
    text = gen_inputs.input_file(dims=(n, n, n), offset=(-12.8,) * 3, extent=size, mat="input/phantom_mat.dat",
                                 den="input/phantom_den.dat", source="input/line_src.txt", geo="input/ring.geo",
                                 blur=(1, 662000, 0.0, 0, 0))
    ex = workdir(tmp_path, text, phantom=(mat, den), geo_text=gen_inputs.ring_geo(32, 40.0),
                 extra={"line_src.txt": gen_inputs.source_file(src_rows)})
    with api.Context(0) as c:
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        assert c.geometry_counts()["counts"] == [4, 8, 52, 64] if isinstance(c.geometry_counts(), dict) else True
        c.set_digitizer(coinc_window_us=0.01, coinc_min_panel_diff=4)
        c.plan_frames(1 << 20)
        st = c.run(None)
        singles = c.result_singles(); co = c.result_coincidences()
        c.stage_source(0)
        q0 = c.fetch_photons(0)[:200000]
    assert st.frames >= 2 and abs(st.pairs - decays) < 6 * np.sqrt(decays)
    assert singles["pann"].min() == 0 and singles["pann"].max() == 31 and singles["modn"].max() < 52
    assert np.all(np.diff(singles["t"]) >= 0)
    # 10 cm of water on average: most photons scatter (checked below), a few per cent end by photo-absorption
    out_frac = st.photons_phantom_out / (2.0 * st.pairs)
    assert 0.9 < out_frac < 0.999
    d = np.abs(co["a"]["pann"] - co["b"]["pann"]); d = np.minimum(d, 32 - d)
    assert co.size > 0 and d.min() >= 4
    # scatter fraction of the coincidences from the class tallies: most photons scatter in 20 cm of
    # water, the energy window removes the large angles
    assert st.trues + st.scatters + st.randoms == st.coincidences
    if st.trues + st.scatters > 500:
        assert 0.05 < st.scatters / (st.trues + st.scatters) < 0.9
    # per-photon parity in the big phantom (multi-step Woodcock, several Comptons per history) and in the ring
    s = parity.Setup(0, phantom=(mat, den), size=size, geo=ex / "input" / "ring.geo")
    s.ctx.put_photons(0, q0); s.ctx.stage_phantom()
    got = s.ctx.fetch_photons(1)
    want = orc.phantom(q0, s.mat, s.den, s.offset, s.size, s.tab_ph, s.eabs, s.seed)
    want = want[want["t"] > 0]
    ncommon, nmatch, only_g, only_o = parity.compare_photons(got, want)
    assert only_g + only_o <= 0.006 * want.size and nmatch >= 0.99 * ncommon
    sf = (got["nscat"] > 0).mean()
    assert 0.3 < sf < 0.75
    s.ctx.stage_detector()
    ev = s.ctx.fetch_events()
    res = orc.detector(want, s.panels, s.counts4, s.pmat, s.pdens, s.surfaces, s.tab_det, s.eabs, 2, 1, s.seed)
    assert abs(ev.size - res["events"].size) <= 0.01 * res["events"].size + 3 * np.sqrt(res["events"].size)
    chi2, ndf = parity.chi2_hist(ev["E"], res["events"]["E"], np.linspace(0, 520000, 53))
    assert ndf > 20 and chi2 / ndf < 1.6, (chi2, ndf)
    chi2, ndf = parity.chi2_hist(ev["pann"], res["events"]["pann"], np.arange(33) - 0.5)
    assert chi2 / max(ndf, 1) < 1.8
    s.close()
