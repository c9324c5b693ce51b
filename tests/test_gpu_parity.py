"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.
Bit-exact for the deterministic digitizer stages; per-history agreement + chi-square for the transport (fast-math
intrinsics differ in the last bits from libm, so a small fraction of histories may branch differently)."""
import numpy as np
import pytest

import parity
from gpet_b200 import api, refio
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    # moduleN / crystalN of setSitenum come from the panel geometry (initialize.cu:1074-1086), as in the reference
    c.load_geometry(parity.EXAMPLE / "input" / "config8.geo")
    yield c
    c.close()


def run_both(ctx, ev, **kw):
    p, d = parity.make_digi_params(**kw)
    parity.apply_digi_params(ctx, d)
    got, counts = ctx.digitize(ev)
    co = ctx.fetch_coincidences() if d["coinc_window_us"] > 0 else np.zeros(0, api.COINC_DTYPE)
    want, wcounts, wco = orc.digitize(ev, p)
    return got, counts, co, want.astype(api.EVENT_DTYPE), wcounts, wco.astype(api.COINC_DTYPE)


# ------------------------------------------------------------------------------------------------ digitizer: bit exact
@pytest.mark.parametrize("case,d,ev", list(parity.kat_cases()), ids=lambda v: v["name"] if isinstance(v, dict) and "name" in v else None)
def test_digitizer_known_answers_on_gpu(ctx, case, d, ev):
    parity.apply_digi_params(ctx, d)
    singles, counts = ctx.digitize(ev)
    coinc = ctx.fetch_coincidences() if d.get("coinc_window_us", 0) > 0 else []
    parity.check_kat(case, singles, counts, coinc)


@pytest.mark.parametrize("n", [0, 1, 7, 2047, 2048, 2049, 40000, 300000])
@pytest.mark.parametrize("dead_type", [0, 1])
def test_digitizer_bit_exact_sizes(ctx, n, dead_type):
    rng = np.random.default_rng(100 + n)
    ev = parity.random_events(n, rng, tmax=max(10.0, n * 0.05), nsites=300)
    got, counts, _, want, wcounts, _ = run_both(ctx, ev, dead_type=dead_type)
    assert list(counts) == list(wcounts)
    assert parity.events_equal(got, want)


@pytest.mark.parametrize("dead_level", [0, 1, 2, 3])
@pytest.mark.parametrize("dead_type", [0, 1])
def test_digitizer_bit_exact_levels_ties_and_dead_records(ctx, dead_level, dead_type):
    rng = np.random.default_rng(7 + dead_level)
    ev = parity.random_events(60000, rng, tmax=2.0e4, nsites=936, tie_fraction=0.1, dead_fraction=0.03)
    got, counts, co, want, wcounts, wco = run_both(ctx, ev, dead_level=dead_level, dead_type=dead_type,
                                                   coinc_window_us=0.02, coinc_policy=dead_type, coinc_min_panel_diff=1)
    assert list(counts) == list(wcounts)
    assert parity.events_equal(got, want)
    assert co.size == wco.size and co.tobytes() == wco.tobytes()


def test_digitizer_bit_exact_late_times(ctx):
    # t ~ 1e8 us: fp32 tdead quantised to 8 us; the closed form must still agree bit for bit
    rng = np.random.default_rng(5)
    ev = parity.random_events(50000, rng, tmax=3.0e4, nsites=200)
    ev["t"] += 1.0e8
    for dead_type in (0, 1):
        got, counts, _, want, wcounts, _ = run_both(ctx, ev, dead_type=dead_type)
        assert list(counts) == list(wcounts) and parity.events_equal(got, want)


@pytest.mark.parametrize("dead_type", [0, 1])
def test_digitizer_clustered_times_take_the_lsd_fallback(ctx, dead_type):
    """Times clustered far below the key range overfill one slice of the bucket sort: the device switches to the LSD radix
    passes (counters[5]) and the result is still the oracle's, bit for bit."""
    rng = np.random.default_rng(77)
    ev = parity.random_events(50000, rng, tmax=1.0, nsites=936, tie_fraction=0.05)
    ev["t"] += 1.0e6                       # 50 000 events inside one microsecond ...
    ev["t"][:3] = [2.0, 7.5e11, 3.0]      # ... and three outliers that stretch the range by 18 orders of magnitude
    got, counts, co, want, wcounts, wco = run_both(ctx, ev, dead_type=dead_type, dead_time_us=1e-4, coinc_window_us=1e-5)
    assert list(counts) == list(wcounts)
    assert parity.events_equal(got, want)
    assert co.tobytes() == wco.tobytes()


def test_digitizer_negative_site_and_unsorted_input(ctx):
    rng = np.random.default_rng(11)
    ev = parity.random_events(5000, rng, tmax=500.0, nsites=64)
    ev["siten"] -= 32  # std::sort compares siten as signed int
    got, counts, _, want, wcounts, _ = run_both(ctx, ev)
    assert list(counts) == list(wcounts) and parity.events_equal(got, want)


def test_digitizer_blur_matches_oracle_within_float_tolerance(ctx):
    # blur on: same Philox normals on both sides; device logf/cosf differ from libm in the last bits
    rng = np.random.default_rng(3)
    ev = parity.random_events(40000, rng, tmax=1.0e6, nsites=936)
    got, counts, _, want, wcounts, _ = run_both(ctx, ev, blur_Rref=0.05, blur_space=0.02, time_blur_sigma_us=1e-4)
    assert abs(int(counts[3]) - int(wcounts[3])) <= 4
    a = {(int(p), int(s)): e for p, s, e in zip(got["parn"], got["siten"], got["E"])}
    b = {(int(p), int(s)): e for p, s, e in zip(want["parn"], want["siten"], want["E"])}
    common = set(a) & set(b)
    assert len(common) >= 0.999 * len(b)
    rel = np.array([abs(a[k] - b[k]) / b[k] for k in common])
    assert rel.max() < 1e-5
    # the blur really happened: sigma/E = R/2.35482
    orig = {(int(p), int(s)): e for p, s, e in zip(ev["parn"], ev["siten"], ev["E"])}
    z = np.array([(a[k] - orig[k]) / orig[k] for k in common])
    assert abs(z.std() - 0.05 / 2.35482) < 0.0015


def test_digitizer_full_batch_properties(ctx):
    # reference batch scale: 3 * NPART event slots (gPET.cu:26); size-independent properties
    rng = np.random.default_rng(9)
    n = 1_500_000
    ev = parity.random_events(n, rng, tmax=4.0e6, nsites=59904)
    p, d = parity.make_digi_params(coinc_window_us=0.01)
    parity.apply_digi_params(ctx, d)
    s, counts = ctx.digitize(ev)
    assert counts[0] == n and counts[0] >= counts[1] >= counts[2] >= counts[3] == s.size
    assert np.all(np.diff(s["t"]) >= 0)
    assert np.all((s["E"] >= 30000) & (s["E"] <= 700000))
    # set property: every single is one of the inputs, unchanged
    key_in = set(zip(ev["parn"].tolist(), ev["siten"].tolist()))
    assert set(zip(s["parn"].tolist(), s["siten"].tolist())) <= key_in
    # idempotence
    s2, c2 = ctx.digitize(s)
    assert parity.events_equal(s, s2)
    # thresholder count is an exact, order-free quantity
    assert counts[1] == int(((ev["E"] >= 50000) & (ev["E"] <= 2.0e6)).sum())
    # oracle agrees on the whole thing as well
    want, wcounts, _ = orc.digitize(ev, p)
    assert list(counts) == list(wcounts) and parity.events_equal(s, want.astype(api.EVENT_DTYPE))


# ------------------------------------------------------------------------------------------------ transport
needs_tables = pytest.mark.skipif(not parity.have_tables(), reason="packed tables not built")


@needs_tables
def test_source_sampling_matches_oracle():
    s = parity.Setup(0, phantom="air", n=16, capacity=(1 << 20, 1 << 20, 1 << 20))
    c = s.ctx
    c.load_isotopes(parity.EXAMPLE / "data" / "isotopes.txt")
    c.load_source(parity.EXAMPLE / "input" / "source.txt")
    c.set_time_window(0, 30)
    nf = c.plan_frames(0)
    assert nf >= 1
    f = nf - 1 if nf > 1 else 0
    fr = c.frame(f)
    c.stage_source(f)
    got = c.fetch_photons(0)
    src, iso = c.sources(), c.isotopes()
    tau = np.array([np.float64(iso[x["type"]]["halftime"]) * 1.442695 for x in src])
    frac = -np.expm1(-fr["dt_s"] / tau)
    want = orc.source(np.cumsum(fr["pairs"]), [x["shape"] for x in src], np.concatenate([x["coeff"] for x in src]),
                      tau, frac, fr["t0_s"], fr["first_pair"], 0.0037056, int(fr["pairs"].sum()), s.seed)
    assert got.size == want.size == 2 * int(fr["pairs"].sum())
    assert np.array_equal(got["parn"], want["parn"]) and np.array_equal(got["eventid"], want["eventid"])
    # the GPU places uniform angles with the SFU sine / cosine (4e-7 absolute): positions and the first photon's direction
    # agree to 2e-5 everywhere.  The partner's direction goes through rotate(-cosf(delta)), and for |delta| < ~5e-4 rad
    # fp32 cosf(delta) is within an ulp or two of 1: a last-bit change of delta then moves sin(theta) by ~1e-4 (in the
    # reference as well), so a handful of partners per million may differ by a few 1e-4 -- far below the 3.7e-3 rad sigma
    for f_ in ("x", "y", "z", "vx", "vy", "vz"):
        d = np.abs(got[f_].astype(np.float64) - want[f_])
        assert d[0::2].max() < 2e-5, f_
        assert (d[1::2] > 2e-5).mean() < 1e-4 and d[1::2].max() < 1e-3, f_
    assert np.allclose(got["E"], want["E"], rtol=1e-6)
    assert np.allclose(got["t"], want["t"], rtol=1e-12, atol=1e-6)
    # physics: times inside the frame, pairs back to back within the acollinearity
    assert got["t"].min() >= fr["t0_s"] * 1e6 and got["t"].max() <= (fr["t0_s"] + fr["dt_s"]) * 1e6 * (1 + 1e-9)
    cosang = (got["vx"][0::2] * got["vx"][1::2] + got["vy"][0::2] * got["vy"][1::2] + got["vz"][0::2] * got["vz"][1::2])
    assert np.all(cosang < -0.999)
    assert abs(np.degrees(np.arccos(-cosang.clip(-1, 1))).std() - np.degrees(0.0037056) * 0.66) < 0.08
    s.close()


@needs_tables
def test_phantom_transport_matches_oracle_per_photon():
    # 10 cm water cube so that a sizeable fraction of photons interacts (Compton, Rayleigh, photo-absorption)
    n = 48
    mat = np.ones((n, n, n), np.int32); den = np.ones((n, n, n), np.float32)
    s = parity.Setup(0, phantom=(mat, den), size=10.0)
    rng = np.random.default_rng(21)
    ph = parity.isotropic_photons(200000, rng, pos_sigma=1.0)
    ph["E"][::3] = 140000.0
    ph["E"][1::7] = 30000.0
    s.ctx.put_photons(0, ph)
    s.ctx.stage_phantom()
    got = s.ctx.fetch_photons(1)
    want = orc.phantom(ph, s.mat, s.den, s.offset, s.size, s.tab_ph, s.eabs, s.seed)
    want = want[want["t"] > 0]
    ncommon, nmatch, only_g, only_o = parity.compare_photons(got, want)
    assert abs(got.size - want.size) <= 0.002 * want.size
    assert only_g + only_o <= 0.004 * want.size
    assert nmatch >= 0.995 * ncommon
    # scattered fraction is substantial, otherwise the test proves nothing
    assert (got["nscat"] > 0).mean() > 0.25
    s.close()


@needs_tables
def test_phantom_recording_sphere_matches_oracle():
    # RECORDPSF == -1 branch of photon() (gPET_kernals.cu:288-294, getDistance :148-171): escaped photons end on the sphere
    n = 32
    mat = np.ones((n, n, n), np.int32); den = np.ones((n, n, n), np.float32)
    s = parity.Setup(0, phantom=(mat, den), size=4.0)
    sphere = (0.1, -0.2, 0.05, 12.0)
    s.ctx.set_transport(record_psf=1, record_sphere=sphere)
    rng = np.random.default_rng(23)
    ph = parity.isotropic_photons(100000, rng, pos_sigma=0.5)
    s.ctx.put_photons(0, ph)
    s.ctx.stage_phantom()
    got = s.ctx.fetch_photons(1)
    want = orc.phantom(ph, s.mat, s.den, s.offset, s.size, s.tab_ph, s.eabs, s.seed, record_sphere=sphere)
    want = want[want["t"] > 0]
    ncommon, nmatch, only_g, only_o = parity.compare_photons(got, want)
    assert only_g + only_o <= 0.004 * want.size and nmatch >= 0.995 * ncommon
    # every photon that left the grid alive (not those stopped below the absorption energy inside) sits on the sphere
    r = np.sqrt((got["x"] - sphere[0]) ** 2 + (got["y"] - sphere[1]) ** 2 + (got["z"] - sphere[2]) ** 2)
    assert (np.abs(r - sphere[3]) < 1e-3).mean() > 0.99
    # and the fused front end takes the same branch: same photons reach the panels as through the staged kernels
    s.ctx.put_photons(0, ph)
    s.ctx.stage_phantom(); s.ctx.stage_detector()
    ev_staged = s.ctx.fetch_events()
    s.ctx.put_photons(0, ph)
    s.ctx.stage_front(-1); s.ctx.stage_panel_transport()
    ev_fused = s.ctx.fetch_events()
    assert ev_staged.size == ev_fused.size > 0
    assert np.array_equal(np.sort(ev_staged, order=["parn", "t", "cryn"]), np.sort(ev_fused, order=["parn", "t", "cryn"]))
    s.close()


@needs_tables
def test_detector_transport_matches_oracle_per_photon():
    s = parity.Setup(0, phantom="air", n=16)
    rng = np.random.default_rng(22)
    ph = parity.isotropic_photons(200000, rng)
    ph["E"][::4] = 250000.0
    s.ctx.put_photons(1, ph)
    s.ctx.stage_detector()
    hits = parity.hits_by_photon(s.ctx.fetch_hits())
    ev = s.ctx.fetch_events()
    res = orc.detector(ph, s.panels, s.counts4, s.pmat, s.pdens, s.surfaces, s.tab_det, s.eabs, 2, 1, s.seed)
    ohits = parity.hits_by_photon(res["hits"]); oev = res["events"]
    # a photon that deposits in more than 6 distinct crystals loses the extra deposits in the adder on both sides
    # (reference: Event events[4] without a bound check, gPET_kernals.cu:852-853) -- must stay a ~1e-5 effect
    assert res["adder_overflow"] <= 5
    assert abs(hits.size - ohits.size) <= 0.003 * ohits.size
    assert abs(ev.size - oev.size) <= 0.003 * oev.size
    # per-photon hit sequences
    def seqs(h):
        d = {}
        for r in h:
            d.setdefault(int(r["parn"]), []).append((int(r["pann"]), int(r["modn"]), int(r["cryn"]), int(r["type"]), float(r["E"])))
        return d
    a, b = seqs(hits), seqs(ohits)
    common = set(a) & set(b)
    assert len(common) >= 0.997 * len(b)
    same = 0
    for k in common:
        x, y = a[k], b[k]
        if len(x) == len(y) and all(u[:4] == v[:4] and abs(u[4] - v[4]) <= 2e-3 * max(v[4], 1.0) for u, v in zip(x, y)):
            same += 1
    assert same >= 0.99 * len(common), (same, len(common))
    # events: energy per (parn, siten) agrees
    ea = {(int(r["parn"]), int(r["siten"])): float(r["E"]) for r in ev}
    eb = {(int(r["parn"]), int(r["siten"])): float(r["E"]) for r in oev}
    ce = set(ea) & set(eb)
    assert len(ce) >= 0.99 * len(eb)
    close = sum(abs(ea[k] - eb[k]) <= 2e-3 * eb[k] for k in ce)
    assert close >= 0.99 * len(ce)
    # geometry coverage of config8 (~0.41 of photons enter a panel)
    frac = len(set(int(x) for x in hits["parn"])) / ph.size
    assert 0.2 < frac < 0.5
    s.close()


@needs_tables
def test_detector_quadric_surfaces_exclude_hits_like_the_oracle():
    # crystalSearch (gPET_kernals.cu:1241-1245): a point where a quadric is < 0 is not in a crystal (no hit, gap material).
    # Two surfaces: the shipped inert one (dropped on the host) and a real one, x + 1 < 0: the rear half of the 2 cm panels.
    s = parity.Setup(0, phantom="air", n=16)
    surfaces = np.array([0, 0, 0, 0, 0, 0, 0, 0, 0, 1,   0, 0, 0, 0, 0, 0, 1, 0, 0, 1], np.float32)
    s.ctx.set_transport(nsurface=2, surface=list(surfaces))
    rng = np.random.default_rng(24)
    ph = parity.isotropic_photons(150000, rng)
    s.ctx.put_photons(1, ph)
    s.ctx.stage_detector()
    hits = s.ctx.fetch_hits(); ev = s.ctx.fetch_events()
    res = orc.detector(ph, s.panels, s.counts4, s.pmat, s.pdens, surfaces, s.tab_det, s.eabs, 2, 1, s.seed)
    assert hits.size > 10000 and abs(hits.size - res["hits"].size) <= 0.003 * res["hits"].size
    assert abs(ev.size - res["events"].size) <= 0.003 * res["events"].size
    assert hits["x"].min() >= -1.0 - 1e-5 and res["hits"]["x"].min() >= -1.0 - 1e-5      # nothing recorded behind the surface
    a = {(int(r["parn"]), int(r["siten"])): float(r["E"]) for r in ev}
    b = {(int(r["parn"]), int(r["siten"])): float(r["E"]) for r in res["events"]}
    common = set(a) & set(b)
    assert len(common) >= 0.99 * len(b) and sum(abs(a[k] - b[k]) <= 2e-3 * b[k] for k in common) >= 0.99 * len(common)
    # with the inert surface alone the rear half records hits again
    s.ctx.set_transport(nsurface=1, surface=list(surfaces[:10]))
    s.ctx.put_photons(1, ph)
    s.ctx.stage_detector()
    assert s.ctx.fetch_hits()["x"].min() < -1.5
    s.close()


@needs_tables
def test_pipeline_spectra_agree_statistically_with_independent_seeds():
    # statistical parity: GPU run with one seed vs oracle run with another on the same inputs; chi-square / ndf
    s = parity.Setup(0, phantom="cylinder", n=32, size=2.0, seed=1111)
    rng = np.random.default_rng(23)
    ph = parity.isotropic_photons(400000, rng)
    s.ctx.put_photons(0, ph)
    s.ctx.stage_phantom(); s.ctx.stage_detector()
    hits = s.ctx.fetch_hits(); ev = s.ctx.fetch_events()
    oph = orc.phantom(ph, s.mat, s.den, s.offset, s.size, s.tab_ph, s.eabs, 2222)
    res = orc.detector(oph, s.panels, s.counts4, s.pmat, s.pdens, s.surfaces, s.tab_det, s.eabs, 2, 1, 2222)
    # sensitivity within 1 % (north_star tolerance), spectra by chi-square
    assert abs(ev.size - res["events"].size) <= 0.01 * res["events"].size + 3 * np.sqrt(res["events"].size)
    assert abs(hits.size - res["hits"].size) <= 0.01 * res["hits"].size + 3 * np.sqrt(res["hits"].size)
    bins = np.linspace(0, 520000, 53)
    chi2, ndf = parity.chi2_hist(hits["E"], res["hits"]["E"], bins)
    assert ndf > 20 and chi2 / ndf < 1.6, (chi2, ndf)
    chi2, ndf = parity.chi2_hist(ev["E"], res["events"]["E"], bins)
    assert ndf > 20 and chi2 / ndf < 1.6, (chi2, ndf)
    chi2, ndf = parity.chi2_hist(ev["modn"], res["events"]["modn"], np.arange(118) - 0.5)
    assert chi2 / max(ndf, 1) < 1.6
    s.close()


@needs_tables
@pytest.mark.parametrize("mode", ["source", "queue"])
def test_fused_front_end_equals_staged_kernels(mode):
    """gpet_stage_front (one kernel, what gpet_run uses) must deliver exactly the photons of the staged
    source -> phantom -> panel-entry kernels: same Philox counters, same device functions."""
    n = 40
    mat = np.ones((n, n, n), np.int32); den = np.ones((n, n, n), np.float32)
    s = parity.Setup(0, phantom=(mat, den), size=4.0, capacity=(1 << 20, 1 << 21, 1 << 20))
    c = s.ctx
    if mode == "source":
        c.load_isotopes(parity.EXAMPLE / "data" / "isotopes.txt")
        c.load_source(parity.EXAMPLE / "input" / "source.txt")
        c.set_time_window(0, 30)
        assert c.plan_frames(0) >= 1
        c.stage_source(0)
    else:
        ph = parity.isotropic_photons(300001, np.random.default_rng(3), pos_sigma=0.8)
        ph["t"][::1000] = 0.0        # dead records are skipped (gPET_kernals.cu:272)
        c.put_photons(0, ph)
    c.stage_phantom()
    n1 = c.queue_size(1)
    c.stage_detector()
    staged = c.fetch_photons(2)
    ev_staged = c.fetch_events()
    if mode == "source":
        c.stage_front(0)
    else:
        c.stage_front(-1)
    fused = c.fetch_photons(2)
    assert c.queue_size(1) == n1 and n1 > 100000          # same tally of photons leaving the phantom
    c.stage_panel_transport()
    ev_fused = c.fetch_events()
    assert staged.size == fused.size > 50000
    a = staged[np.argsort(staged["parn"], kind="stable")]; b = fused[np.argsort(fused["parn"], kind="stable")]
    assert a.tobytes() == b.tobytes()
    assert (a["nscat"] >= 0).all() and (a["nscat"] < 8).all()        # queue 2 carries the panel index in this word
    ka = np.lexsort((ev_staged["siten"], ev_staged["parn"])); kb = np.lexsort((ev_fused["siten"], ev_fused["parn"]))
    assert ev_staged[ka].tobytes() == ev_fused[kb].tobytes()
    s.close()


# ------------------------------------------------------------------------------------------------ whole path
def make_example_dir(tmp_path, n=32, source="pointsource.txt", window="0 120"):
    ex = tmp_path / "ex"
    (ex / "input").mkdir(parents=True); (ex / "data").mkdir(); (ex / "output").mkdir()
    text = (parity.EXAMPLE / "input_PET.in").read_text().replace("200 200 200", f"{n} {n} {n}")
    text = text.replace("input/pointsource.txt", f"input/{source}").replace("0 120\n", window + "\n")
    (ex / "input_PET.in").write_text(text)
    for f in ("config8.geo", "pointsource.txt", "source.txt"):
        (ex / "input" / f).write_text((parity.EXAMPLE / "input" / f).read_text())
    (ex / "data" / "isotopes.txt").write_text((parity.EXAMPLE / "data" / "isotopes.txt").read_text())
    (ex / "data" / "input4gPET.gpettab").write_bytes(parity.PACKED.read_bytes())
    mat, den = parity.gen_inputs.cylinder_phantom(n=n)
    parity.gen_inputs.write_phantom(mat, den, ex / "input" / "cylinder_phantom_mat.dat", ex / "input" / "cylinder_phantom_den.dat")
    return ex


@needs_tables
def test_run_shipped_example_writes_reference_layouts(tmp_path):
    ex = make_example_dir(tmp_path)
    with api.Context(0) as c:
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        c.set_digitizer(coinc_window_us=0.01)
        st = c.run(ex / "output")
        singles = c.result_singles()
    # 1 762 974 000 atoms, T1/2 820 500 s, 120 s -> ~178 711 pairs
    assert abs(st.pairs - 178711) < 6 * np.sqrt(178711)
    assert st.overflow_hits == st.overflow_events == 0
    assert st.overflow_adder <= 1e-4 * st.pairs     # deposits beyond 6 distinct crystals per photon (counted)
    assert st.events_adder >= st.events_threshold >= st.events_deadtime >= st.singles > 0
    ids, f = refio.read_hits(ex / "output" / "HitsID.dat", ex / "output" / "Hits.dat")
    assert ids.shape == f.shape == (st.hits, 5)
    assert set(np.unique(ids[:, 4])) <= {1, 2, 4} and ids[:, 1].min() >= 0 and ids[:, 1].max() <= 7
    assert ids[:, 2].max() < 117 and ids[:, 3].max() < 64
    adder = refio.read_events(ex / "output" / "adder.dat")
    sing = refio.read_events(ex / "output" / "singles.dat")
    assert adder.size == st.events_adder and sing.size == st.singles == singles.size
    assert sing.tobytes() == singles.tobytes()
    co = refio.read_coincidences(ex / "output" / "coincidences.dat")
    assert co.size == st.coincidences > 0
    assert np.all(co["b"]["t"] - co["a"]["t"] < 0.01)
    # most coincidences are true pairs (same annihilation) for a point source in a 1 cm phantom
    assert (co["a"]["eventid"] == co["b"]["eventid"]).mean() > 0.9
    # sensitivity: ~0.41 of the photons reach a panel
    assert 0.3 < st.photons_on_panel / (2 * st.pairs) < 0.5
    # replay pin: adder.dat is written BEFORE blur, as the reference does (gPET.cu:383-388); through the digitizer alone, blur
    # on as shipped and the same Philox key, it reproduces singles.dat bit for bit (the blur stream of an event is keyed by
    # its photon and readout site, both in the record)
    # Equal times: a run orders them by the dead-time site number (the order of adder.dat is whatever the detector kernel's
    # atomics made it), a replay keeps the input order -- so the replay gets the list in site order (stable: nothing else moves)
    site = (adder["pann"].astype(np.int64) * 117 + adder["modn"]) * 64 + adder["cryn"]
    with api.Context(0) as c2:
        c2.load_config_file(ex / "input_PET.in", base_dir=ex)
        again, _ = c2.digitize(adder[np.argsort(site, kind="stable")])
    # the dump really is the unblurred list: the photopeak of the adder is as narrow as the acollinearity leaves it
    # (E = 511 keV (1 +- delta / 2), sigma 0.95 keV), that of the singles carries the 5 % energy blur (sigma 11 keV)
    peak = adder["E"][(adder["E"] > 505e3) & (adder["E"] < 517e3)]
    assert peak.size > 1000 and np.std(peak) < 1500.0 < 5000.0 < np.std(sing["E"][(sing["E"] > 450e3) & (sing["E"] < 570e3)])
    assert again.tobytes() == sing.tobytes()


@needs_tables
def test_run_writes_phase_space_dumps_like_outputpsf(tmp_path):
    # OUTPUTPSF == 2 (gPET.cu:296-351): outsource/idsource/timesource after source sampling, out/id/timephantom after the
    # phantom; layouts of readOutput.m:36-54.  The dumps force the staged kernels: the singles must not change.
    ex = make_example_dir(tmp_path, source="source.txt", window="0 10")
    od = ex / "output"
    with api.Context(0) as c:
        c.set_capacity(1 << 17, 1 << 19, 1 << 18)
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        st0 = c.run(None)
        s0 = c.result_singles().copy()
        c.set_psf_output(2)
        st = c.run(od)
        s1 = c.result_singles().copy()
    assert st.frames >= 2 and st.pairs == st0.pairs
    assert s0.tobytes() == s1.tobytes()
    src, sid, stime = refio.read_psf_triplet(od / "outsource.dat", od / "idsource.dat", od / "timesource.dat")
    assert src.shape == (2 * st.pairs, 7) and sid.size == stime.size == 2 * st.pairs
    assert np.array_equal(sid[0::2], sid[1::2]) and np.array_equal(stime[0::2], stime[1::2])    # the two photons of a pair
    assert np.array_equal(src[0::2, :3], src[1::2, :3])
    assert np.allclose(np.linalg.norm(src[:, 3:6], axis=1), 1.0, atol=1e-4)
    assert np.all(np.abs(src[:, 6] - 510999.1) < 30e3) and stime.min() > 0 and stime.max() <= 10e6
    # back to back up to the acollinearity (sigma 0.0037 rad)
    assert np.all(np.einsum("ij,ij->i", src[0::2, 3:6], src[1::2, 3:6]) < -0.999)
    ph, pid, ptime = refio.read_psf_triplet(od / "outphantom.dat", od / "idphantom.dat", od / "timephantom.dat")
    assert ph.shape == (st.photons_phantom_out, 7) and pid.size == ptime.size == st.photons_phantom_out
    assert 0.8 * 2 * st.pairs < pid.size <= 2 * st.pairs and np.all(np.isin(pid, sid))
    assert np.all(ph[:, 6] <= src[:, 6].max()) and ptime.min() > 0
    # PSF-input mode, OUTPUTPSF == 1 (gPET.cu:63-88): the uploaded photons come back as the source triplet
    rng = np.random.default_rng(5)
    n = 20000
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1)[:, None]
    t = np.repeat(np.arange(1, n + 1, dtype=np.float64), 2)
    vv = np.empty((2 * n, 3)); vv[0::2] = v; vv[1::2] = -v
    refio.write_psf(ex / "input" / "psf.dat", np.zeros(2 * n), np.zeros(2 * n), np.zeros(2 * n), t, vv[:, 0], vv[:, 1], vv[:, 2],
                    np.full(2 * n, 511000.0))
    od2 = tmp_path / "out_psf"; od2.mkdir()
    with api.Context(0) as c:
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        c.load_psf(ex / "input" / "psf.dat", 0, 1)
        c.set_psf_output(1)
        st = c.run(od2)
    src, sid, stime = refio.read_psf_triplet(od2 / "outsource.dat", od2 / "idsource.dat", od2 / "timesource.dat")
    assert src.shape == (2 * n, 7) and np.array_equal(sid, np.arange(2 * n)) and np.array_equal(stime, t)
    assert np.array_equal(src[:, 3:6], vv.astype(np.float32)) and not (od2 / "outphantom.dat").exists()
    assert st.singles > 0


@needs_tables
def test_noise_singles_match_oracle_and_enter_the_digitizer(tmp_path):
    # addnoise (gPET_kernals.cu:699-735): same Philox streams on both sides; fp64 log / cos may differ in the last bit
    s = parity.Setup(0, phantom="air", n=16)
    c = s.ctx
    lo, hi = 2.5e5, 7.5e5
    p, d = parity.make_digi_params(seed=s.seed, dead_level=2, ewin_min=350000.0, ewin_max=650000.0, coinc_window_us=0.0)
    parity.apply_digi_params(c, d)
    c.set_digitizer(noise_mean_gap_us=1.5, noise_Emean_eV=4.0e5, noise_sigma_eV=3.0e4, noise_interval_us=500.0)
    c.put_events(np.zeros(0, api.EVENT_DTYPE))
    c.stage_noise(lo, hi)
    got = c.fetch_events()
    want = orc.noise(lo, hi, 1.5, 4.0e5, 3.0e4, 500.0, 8, 117, 64, s.seed).astype(api.EVENT_DTYPE)
    assert abs(want.size - (hi - lo) / 1.5) < 5 * np.sqrt((hi - lo) / 1.5)
    assert got.size == want.size
    got = got[np.argsort(got["eventid"].view(np.uint32), kind="stable")]
    want = want[np.argsort(want["eventid"].view(np.uint32), kind="stable")]
    for k in ("parn", "pann", "modn", "cryn", "siten", "eventid", "x", "y", "z"):
        assert np.array_equal(got[k], want[k]), k
    assert np.allclose(got["t"], want["t"], rtol=1e-13, atol=0) and np.allclose(got["E"], want["E"], rtol=1e-6)
    # noise + digitizer: the GPU's own noise list through the oracle digitizer gives the GPU's singles bit for bit
    c.stage_digitize()
    singles = c.fetch_singles()
    osingles, _, _ = orc.digitize(got, p)
    assert 0.5 * got.size < singles.size < got.size and singles.tobytes() == osingles.astype(api.EVENT_DTYPE).tobytes()
    s.close()
    # gpet_run injects the noise per frame: more singles than without, all noise singles inside the acquisition window
    ex = make_example_dir(tmp_path, source="source.txt", window="0 10")
    with api.Context(0) as c:
        c.set_capacity(1 << 17, 1 << 19, 1 << 18)
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        st0 = c.run(None)
        c.set_digitizer(noise_mean_gap_us=100.0, noise_Emean_eV=4.5e5, noise_sigma_eV=1.0e4, noise_interval_us=1.0e4)
        st1 = c.run(None)
        sg = c.result_singles().copy()
    nz = sg[sg["parn"] == -1]
    assert st1.frames >= 2 and st1.events_adder - st0.events_adder == pytest.approx(1e7 / 100.0, abs=5 * np.sqrt(1e5))
    assert 0.5 * 1e5 < nz.size <= st1.events_adder - st0.events_adder
    assert nz["t"].min() >= 0 and nz["t"].max() < 1e7 and np.all(np.diff(sg["t"][np.argsort(sg["t"], kind="stable")]) >= 0)


@needs_tables
def test_run_repeats_with_the_lsd_fallback_when_decay_times_cluster(tmp_path):
    # gpet_run's first attempt skips the idle launch of the LSD fallback (source-mode times are spread over the frame).
    # A half-life of 0.1 ms in a 10 s frame puts every decay into the first slices of the bucket sort: the overflow flag
    # comes back with the frame's counters and the run is repeated with the fallback enqueued.
    ex = make_example_dir(tmp_path, source="source.txt", window="0 10")
    (ex / "data" / "isotopes.txt").write_text((parity.EXAMPLE / "data" / "isotopes.txt").read_text().replace("6586.26 0.97", "0.0001 0.97"))
    (ex / "input" / "source.txt").write_text("1\natoms, isotope row, shape, centre x y z (cm), three shape parameters#\n"
                                             "150000 0 1 0 0 0 0.3 0.5 0\n")
    with api.Context(0) as c:
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        c.set_digitizer(coinc_window_us=0.01, blur_Rref=0.0)
        c.profile(True)
        st = c.run(None)
        kt = c.kernel_times()
        c.profile(False)
        sg = c.result_singles().copy()
        st2 = c.run_resident()
    assert st.frames == 1 and abs(st.pairs - 150000 * 0.97) < 6 * np.sqrt(150000 * 0.03 * 0.97) + 1
    assert "k_lsd_fallback" in kt and kt["k_bucket_scan"][1] == 2      # two attempts, the second with the fallback
    assert sg.size == st.singles > 20000 and np.all(np.diff(sg["t"]) >= 0) and sg["t"].max() < 5e3
    assert st2.singles == st.singles and st2.coincidences == st.coincidences
    # the same events through the replay entry (which always carries the fallback) give the same singles
    with api.Context(0) as c:
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        c.set_digitizer(coinc_window_us=0.01, blur_Rref=0.0)
        c.plan_frames(0)
        c.stage_front(0); c.stage_panel_transport()
        ev = c.fetch_events()
        again, _ = c.digitize(ev)
    assert again.size == sg.size
    k = lambda e: np.sort(e, order=["t", "parn", "siten"]).tobytes()
    assert k(again) == k(sg)


@needs_tables
@pytest.mark.parametrize("case", ["ring8_small_phantom", "ring32_offset_sources", "psf_input"])
def test_direction_table_of_the_panel_search_changes_nothing(tmp_path, monkeypatch, case):
    # gpet_run narrows the panel search with a direction table built for the run's phantom, sources and geometry
    # (conservative by construction); GPET_NO_DIRMASK=1 runs the plain search.  Same singles, byte for byte.
    import os
    if case == "ring8_small_phantom":
        ex = make_example_dir(tmp_path, source="source.txt", window="0 12")
    elif case == "ring32_offset_sources":
        # 32 panels at 30 cm, sources far off centre (8 cm) and a big phantom: wide candidate sets
        ex = make_example_dir(tmp_path, n=40, source="source.txt", window="0 12")
        (ex / "input" / "config8.geo").write_text(parity.gen_inputs.ring_geo(32, 30.0))
        text = (ex / "input_PET.in").read_text().replace("-0.5 -0.5 -0.5", "-10 -10 -10").replace("1 1 1\n", "20 20 20\n")
        (ex / "input_PET.in").write_text(text)
        mat, den = parity.gen_inputs.water_cylinder_phantom(40, 0.5, 16.0, 16.0)
        parity.gen_inputs.write_phantom(mat, den, ex / "input" / "cylinder_phantom_mat.dat", ex / "input" / "cylinder_phantom_den.dat")
        (ex / "input" / "source.txt").write_text("3\natoms, isotope row, shape, centre x y z (cm), three shape parameters#\n"
                                                 "4000000 0 1 6 5 -4 1.0 4 0\n3000000 0 2 -7 3 6 1.5 0 0\n3000000 0 0 0 -8 0 2 2 6\n")
    else:
        ex = make_example_dir(tmp_path, window="0 12")
        rng = np.random.default_rng(11)
        n = 60000
        v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1)[:, None]
        pos = rng.uniform(-3, 3, size=(n, 3))
        refio.write_psf(ex / "input" / "psf.dat", pos[:, 0], pos[:, 1], pos[:, 2], np.arange(1, n + 1, dtype=np.float64), v[:, 0], v[:, 1], v[:, 2],
                        np.full(n, 511000.0))
    def run(no_table):
        if no_table:
            monkeypatch.setenv("GPET_NO_DIRMASK", "1")
        else:
            monkeypatch.delenv("GPET_NO_DIRMASK", raising=False)
        with api.Context(0) as c:
            c.set_capacity(1 << 18, 1 << 20, 1 << 19)
            c.load_config_file(ex / "input_PET.in", base_dir=ex)
            if case == "psf_input":
                c.load_psf(ex / "input" / "psf.dat", 0, 1)
            c.set_digitizer(coinc_window_us=0.01)
            st = c.run(None)
            return st, c.result_singles().copy()
    st_a, s_a = run(False)
    st_b, s_b = run(True)
    assert st_a.singles == st_b.singles > 5000 and st_a.photons_on_panel == st_b.photons_on_panel and st_a.hits == st_b.hits
    assert s_a.tobytes() == s_b.tobytes()


@needs_tables
def test_run_edge_cases_empty_acquisition_and_overflow(tmp_path):
    ex = make_example_dir(tmp_path, source="source.txt", window="0 4")
    # (1) no atoms: every frame is empty; the run succeeds with empty results
    (ex / "input" / "source.txt").write_text("1\natoms, isotope row, shape, centre x y z (cm), three shape parameters#\n0 0 1 0 0 0 0.3 0.5 0\n")
    with api.Context(0) as c:
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        c.set_digitizer(coinc_window_us=0.01)
        st = c.run(None)
        assert st.pairs == 0 and st.singles == 0 and st.coincidences == 0 and c.result_singles().size == 0
        assert c.run_resident().singles == 0
    # (2) an event buffer far too small for the frame: the run reports GPET_ERR_CAPACITY, it does not drop records silently
    (ex / "input" / "source.txt").write_text((parity.EXAMPLE / "input" / "source.txt").read_text())
    with api.Context(0) as c:
        c.set_capacity(1 << 18, 1 << 12, 1 << 11)      # 2048 events, 4096 hits for ~37 k pairs
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        with pytest.raises(api.GpetError) as e:
            c.run(None)
        assert e.value.code == -5
        # and the context is still usable afterwards with sane capacities? (a new context: capacities are fixed at first use)
    with api.Context(0) as c:
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        st = c.run(None)
        assert st.singles > 10000 and st.overflow_events == 0 and st.overflow_hits == 0
    # (3) a frame larger than the photon queue cannot be planned around: frames are cut to the capacity instead
    with api.Context(0) as c:
        c.set_capacity(1 << 12, 1 << 16, 1 << 15)      # 2048 pairs per frame
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        st2 = c.run(None)
        # (the planner draws the decays per frame: another frame plan is another, statistically equal, acquisition)
        assert st2.frames > 10 and abs(st2.pairs - st.pairs) < 6 * np.sqrt(st.pairs) and abs(st2.singles - st.singles) < 0.03 * st.singles


def test_c_example_replays_adder_dat_like_the_oracle(tmp_path):
    # examples/c/digitize_replay.c: adder.dat -> singles.dat through the C ABI from plain C, against the oracle
    import subprocess
    exe = tmp_path / "digitize_replay"
    libdir = parity.ROOT / "gpet_b200"
    subprocess.run(["gcc", "-std=c99", f"-I{parity.ROOT / 'include'}", str(parity.ROOT / "examples" / "c" / "digitize_replay.c"),
                    f"-L{libdir}", "-lgpet_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    ev = parity.random_events(30000, np.random.default_rng(4), tmax=2.0e5)
    ev.tofile(tmp_path / "adder.dat")
    r = subprocess.run([str(exe), str(parity.EXAMPLE / "input" / "config8.geo"), str(tmp_path / "adder.dat"), str(tmp_path / "singles.dat"), "0"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    p, _ = parity.make_digi_params(ewin_min=350000.0, ewin_max=650000.0)
    want, counts, _ = orc.digitize(ev, p, want_coinc=False)
    got = refio.read_events(tmp_path / "singles.dat")
    assert got.size == want.size > 5000 and got.tobytes() == want.astype(api.EVENT_DTYPE).tobytes()
    assert f"singles {want.size}" in r.stdout


@needs_tables
def test_c_example_runs_an_acquisition_with_compact_delivery(tmp_path):
    # examples/c/run_classify.c on the GPU: gpet_run from plain C99 with index-pair coincidences and 32-byte singles, the
    # expansion checked inside the program (gpet_result_singles vs gpet_expand_singles), class totals adding up
    import re
    import subprocess
    exe = tmp_path / "run_classify"
    libdir = parity.ROOT / "gpet_b200"
    subprocess.run(["gcc", "-std=c99", f"-I{parity.ROOT / 'include'}", str(parity.ROOT / "examples" / "c" / "run_classify.c"),
                    f"-L{libdir}", "-lgpet_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    ex = make_example_dir(tmp_path, n=32)
    r = subprocess.run([str(exe), "input_PET.in", "0.01", "0"], cwd=ex, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    m = re.search(r"coincidences (\d+): trues (\d+), scatters (\d+), randoms (\d+)", r.stdout)
    assert m and int(m.group(1)) == int(m.group(2)) + int(m.group(3)) + int(m.group(4)) > 1000
    assert "first single: t = " in r.stdout and re.search(r"singles (\d+)", r.stdout)


@needs_tables
def test_run_has_no_hidden_host_work(tmp_path):
    # The wall clock of a resident run must stay close to its device time (CUDA events inside the run): host work that
    # creeps into gpet_run (the direction table was once rebuilt at every call: +0.4 ms) doubles the step of bench.py
    # without showing in any kernel time.
    import time
    ex = make_example_dir(tmp_path, n=64, source="source.txt", window="0 120")
    with api.Context(0) as c:
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        c.set_digitizer(coinc_window_us=0.01)
        c.plan_frames(0)
        best = None
        for _ in range(6):
            t0 = time.perf_counter()
            st = c.run_resident()
            wall = (time.perf_counter() - t0) * 1e3
            if best is None or wall < best[0]:
                best = (wall, st.ms_total)
    assert st.pairs > 1.0e6
    assert best[0] < 1.3 * best[1] + 0.1, best


@needs_tables
def test_run_is_reproducible_and_shards_by_frame(tmp_path):
    ex = make_example_dir(tmp_path, source="source.txt", window="0 20")
    def run(rank, world, seed=77):
        with api.Context(0) as c:
            c.set_seed(seed)
            c.set_capacity(1 << 17, 1 << 19, 1 << 18)
            c.load_config_file(ex / "input_PET.in", base_dir=ex)
            c.set_shard(rank, world)
            st = c.run(None)
            return st, c.result_singles()
    st_a, s_a = run(0, 1)
    st_b, s_b = run(0, 1)
    assert st_a.frames >= 3
    assert s_a.tobytes() == s_b.tobytes()           # same seed: bit-identical singles
    st0, s0 = run(0, 2); st1, s1 = run(1, 2)
    assert st0.pairs + st1.pairs == st_a.pairs and st0.singles + st1.singles == st_a.singles
    merged = np.concatenate([s0, s1])
    merged = merged[np.argsort(merged["t"], kind="stable")]
    assert merged.tobytes() == s_a[np.argsort(s_a["t"], kind="stable")].tobytes()   # G-invariant results
    _, s_c = run(0, 1, seed=78)
    assert s_c.tobytes() != s_a.tobytes()


@needs_tables
def test_coincidence_pairs_equal_coincidence_records(tmp_path):
    """GPET_COINC_PAIRS (8-byte index pairs into the singles list, base carried on the device from frame to frame) is the
    same answer as GPET_COINC_RECORDS: same records on demand, same coincidences.dat."""
    ex = make_example_dir(tmp_path, source="source.txt", window="0 20")
    out = {}
    for fmt in (api.Context.COINC_RECORDS, api.Context.COINC_PAIRS):
        od = tmp_path / f"out{fmt}"
        od.mkdir()
        with api.Context(0) as c:
            c.set_seed(5)
            c.set_capacity(1 << 17, 1 << 19, 1 << 18)
            c.load_config_file(ex / "input_PET.in", base_dir=ex)
            c.set_digitizer(coinc_window_us=0.01)
            c.set_coincidence_format(fmt)
            st = c.run(None)
            out[fmt] = (st, c.result_singles(), c.result_coincidences(), c.result_coincidence_pairs())
            st2 = c.run(od)      # file path (frame by frame)
            assert st2.coincidences == st.coincidences
            out[fmt] += (refio.read_coincidences(od / "coincidences.dat"),)
    (st_r, s_r, co_r, p_r, f_r), (st_p, s_p, co_p, p_p, f_p) = out[0], out[1]
    assert st_r.frames >= 3 and st_r.coincidences == st_p.coincidences > 100
    assert s_r.tobytes() == s_p.tobytes()
    assert p_r.shape[0] == 0 and p_p.shape[0] == st_p.coincidences
    assert co_r.tobytes() == co_p.tobytes() == f_r.tobytes() == f_p.tobytes()
    assert np.all(p_p[:, 0] < p_p[:, 1]) and p_p.max() < s_p.size
    assert s_p[p_p[:, 0]].tobytes() == co_p["a"].tobytes() and s_p[p_p[:, 1]].tobytes() == co_p["b"].tobytes()


@needs_tables
@pytest.mark.parametrize("dead_level", [1, 2, 3])
def test_compact_singles_expand_to_the_48_byte_records(tmp_path, dead_level):
    """GPET_SINGLES_COMPACT (32-byte singles on the way to the host) is the same answer as the 48-byte Event: the expanded
    records, the coincidences gathered through the index pairs and the class bytes are byte-identical, at every dead-time
    level (siten is rebuilt from the level and the ids); file runs ignore the format; PSF input and noise refuse it."""
    ex = make_example_dir(tmp_path, source="source.txt", window="0 20")
    out = {}
    for fmt in (api.Context.SINGLES_RECORDS, api.Context.SINGLES_COMPACT):
        with api.Context(0) as c:
            c.set_seed(9)
            c.set_capacity(1 << 17, 1 << 19, 1 << 18)
            c.load_config_file(ex / "input_PET.in", base_dir=ex)
            c.set_digitizer(coinc_window_us=0.01, dead_level=dead_level)
            c.set_coincidence_format(api.Context.COINC_PAIRS)
            c.set_singles_format(fmt)
            st = c.run(None)
            out[fmt] = (st, c.result_singles(), c.result_coincidences(), c.result_coincidence_classes())
            if fmt == api.Context.SINGLES_COMPACT:
                cs = c.result_singles_compact()
                od = tmp_path / "files"
                od.mkdir()
                st_f = c.run(od)                       # a file run keeps the reference layout
                assert st_f.singles == st.singles
                assert refio.read_events(od / "singles.dat").tobytes() == out[fmt][1].tobytes()
                c.set_digitizer(noise_mean_gap_us=5.0, noise_interval_us=100.0)
                with pytest.raises(api.GpetError):
                    c.run(None)
            else:
                with pytest.raises(api.GpetError):
                    c.result_singles_compact()
    (st_r, s_r, co_r, k_r), (st_c, s_c, co_c, k_c) = out[0], out[1]
    assert st_r.frames >= 3 and st_r.singles == st_c.singles == s_r.size > 1000 and st_r.coincidences > 100
    assert s_r.tobytes() == s_c.tobytes()
    assert co_r.tobytes() == co_c.tobytes() and k_r.tobytes() == k_c.tobytes()
    assert cs.size == s_r.size and cs.dtype.itemsize == 32
    assert np.array_equal(cs["t"], s_r["t"]) and np.array_equal(cs["E"], s_r["E"]) and np.array_equal(cs["eventid"], s_r["eventid"])
    assert np.array_equal(cs["ids"] & 0xff, s_r["pann"]) and np.array_equal(cs["ids"] >> 31, s_r["parn"] & 1)


# ------------------------------------------------------------------------------------------------ coincidence classes
def paired_events(npairs, rng, tmax, nnoise=0, pair_shift=0):
    """Two events per annihilation a few ns apart (photon numbers 2k, 2k+1), dense enough in time for randoms and
    multiples, plus noise-like events (parn = -1)."""
    n = 2 * npairs
    ev = parity.random_events(n + nnoise, rng, tmax=tmax, nsites=936)
    ev["parn"][:n] = np.arange(n, dtype=np.int32)
    ev["eventid"][:n] = ev["parn"][:n] >> 1 if pair_shift == 0 else ev["parn"][:n]
    ev["t"][1:n:2] = ev["t"][0:n:2] + rng.uniform(0.0, 0.008, npairs)
    ev["E"][:n] = rng.uniform(200e3, 600e3, n).astype(np.float32)
    ev["parn"][n:] = -1
    ev["eventid"][n:] = (0x80000000 | np.arange(nnoise, dtype=np.int64)).astype(np.uint32).view(np.int32)
    return ev[rng.permutation(ev.size)]


@pytest.mark.parametrize("policy", [0, 1])
@pytest.mark.parametrize("pair_shift", [0, 1])
def test_coincidence_classes_of_replayed_events_match_oracle(ctx, policy, pair_shift):
    rng = np.random.default_rng(40 + 2 * policy + pair_shift)
    ev = paired_events(60_000, rng, tmax=1.0e4, nnoise=500, pair_shift=pair_shift)
    p, d = parity.make_digi_params(coinc_window_us=0.01, coinc_policy=policy, ewin_min=100e3)
    parity.apply_digi_params(ctx, d)
    ctx.set_digitizer(coinc_pair_shift=pair_shift)
    want_s, _, want_co = orc.digitize(ev, p)
    for frac in (0.15, 0.0, 0.4):   # marks of an earlier list must not survive put_events
        scattered = rng.choice(ev["parn"][ev["parn"] >= 0], int(frac * ev.size), replace=False).astype(np.int32)
        ctx.put_events(ev)
        ctx.mark_scattered(scattered)
        ctx.stage_digitize()
        co = ctx.fetch_coincidences()
        cls, totals = ctx.fetch_coincidence_classes()
        assert co.tobytes() == want_co.astype(api.COINC_DTYPE).tobytes() and co.size > 5000
        want_cls, want_totals = orc.classify(want_co, scattered, pair_shift)
        assert np.array_equal(cls, want_cls) and list(totals) == list(want_totals)
        assert int(totals.sum()) == co.size and totals[2] > 20 and totals[0] > 1000
        assert (totals[1] > 100) == (frac > 0)
    # one-byte tags: the serial wraps every 255 event lists / frames and the table is cleared then; marks set before the
    # wrap must be gone after it, marks set after it must work
    ctx.put_events(ev)
    ctx.mark_scattered(ev["parn"][ev["parn"] >= 0])
    for _ in range(254):          # + the put below = 255 serials later: the same serial value again
        ctx.put_events(ev[:1])
    scattered = rng.choice(ev["parn"][ev["parn"] >= 0], ev.size // 5, replace=False).astype(np.int32)
    ctx.put_events(ev)
    ctx.mark_scattered(scattered)
    ctx.stage_digitize()
    cls, totals = ctx.fetch_coincidence_classes()
    want_cls, want_totals = orc.classify(want_co, scattered, pair_shift)
    assert np.array_equal(cls, want_cls) and list(totals) == list(want_totals)
    ctx.set_digitizer(coinc_pair_shift=0)
    # the replay entry itself forgets the marks as well
    ctx.digitize(ev)
    cls, totals = ctx.fetch_coincidence_classes()
    assert totals[1] == 0 and np.all(cls != 1)


@needs_tables
def test_run_classifies_coincidences_like_the_staged_scatter_counts(tmp_path):
    """gpet_run's classes (scatter tags written by the fused front end, or at panel entry by the staged kernels when
    phase-space dumps are on) against the nscat of the same photons fetched after the staged phantom kernel."""
    ex = make_example_dir(tmp_path, source="source.txt", window="0 20")
    def setup(c):
        c.set_seed(11)
        c.set_capacity(1 << 17, 1 << 19, 1 << 18)
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        c.set_digitizer(coinc_window_us=0.2, coinc_policy=1)
    scattered = []
    with api.Context(0) as c:
        setup(c)
        nf = c.plan_frames()
        for f in range(nf):
            c.stage_source(f)
            c.stage_phantom()
            ph = c.fetch_photons(1)
            scattered.append(ph["parn"][ph["nscat"] > 0])
    scattered = np.concatenate(scattered).astype(np.int32)
    od = tmp_path / "out"; od.mkdir()
    od2 = tmp_path / "out2"; od2.mkdir()
    with api.Context(0) as c:
        setup(c)
        st = c.run(None)
        co, cls = c.result_coincidences(), c.result_coincidence_classes()
        c.set_coincidence_format(api.Context.COINC_PAIRS)
        st_p = c.run(None)
        cls_p = c.result_coincidence_classes()
        st_r = c.run_resident()
        st_f = c.run(od)             # frame by frame, fused front end
        c.set_psf_output(2)
        st_s = c.run(od2)            # staged kernels: tags written at panel entry
    assert st.frames >= 3 and st.coincidences == co.size == cls.size > 300
    want_cls, want_totals = orc.classify(co, scattered, 0)
    assert np.array_equal(cls, want_cls)
    assert [st.trues, st.scatters, st.randoms] == [int(v) for v in want_totals]
    assert st.trues > 100 and st.scatters > 10 and st.randoms > 10
    assert np.array_equal(cls_p, cls)
    for other in (st_p, st_r, st_f, st_s):
        assert [other.trues, other.scatters, other.randoms] == [st.trues, st.scatters, st.randoms]
    assert np.fromfile(od / "coincidences_class.dat", np.uint8).tobytes() == cls.tobytes()
    assert np.fromfile(od2 / "coincidences_class.dat", np.uint8).tobytes() == cls.tobytes()
    assert refio.read_coincidences(od / "coincidences.dat").tobytes() == co.tobytes()


# ------------------------------------------------------------------------------------------------ positron range / positron PSF
def water_box(n=40, size=2.0):
    mat = np.ones((n, n, n), np.int32); den = np.ones((n, n, n), np.float32)
    den[: n // 2] = 1.85   # a denser half so that the density-scaled march matters
    return mat, den


@needs_tables
def test_source_with_positron_range_matches_oracle():
    # S4 + S5 (gPET_kernals.cu:347-443): O-15 (row 2, endpoint 1.738 MeV kinetic) in a 2 cm water/bone box
    mat, den = water_box()
    s = parity.Setup(0, phantom=(mat, den), size=2.0, capacity=(1 << 20, 1 << 20, 1 << 20))
    c = s.ctx
    c.load_isotopes(parity.EXAMPLE / "data" / "isotopes.txt")
    src_file = parity.ROOT / "tests" / "golden" / "source_o15_box.txt"
    c.load_source(src_file)
    c.set_transport(noncollinearity_rad=0.0037056, use_positron_range=1, eabs_eV=s.eabs, nsurface=1, surface=list(s.surfaces), record_hits=1)
    c.set_time_window(0, 10)
    nf = c.plan_frames(0)
    fr = c.frame(0)
    c.stage_source(0)
    got = c.fetch_photons(0)
    src, iso = c.sources(), c.isotopes()
    tau = np.array([np.float64(iso[x["type"]]["halftime"]) * 1.442695 for x in src])
    frac = -np.expm1(-fr["dt_s"] / tau)
    iso_coef = np.array([i["coef"] for i in iso], np.float32)
    args = (np.cumsum(fr["pairs"]), [x["shape"] for x in src], np.concatenate([x["coeff"] for x in src]), tau, frac, fr["t0_s"],
            fr["first_pair"], 0.0037056, int(fr["pairs"].sum()), s.seed)
    want = orc.source_with_range(*args, [x["type"] for x in src], iso_coef, s.den, s.offset, s.size)
    plain = orc.source(*args)
    assert got.size == want.size > 20000
    d = np.sqrt((got["x"] - want["x"]) ** 2 + (got["y"] - want["y"]) ** 2 + (got["z"] - want["z"]) ** 2)
    assert (d < 2e-4).mean() > 0.995           # a few walks cross a voxel face differently (powf/logf last bits)
    for f_ in ("vx", "vy", "vz"):   # partners with |delta| < ~5e-4 rad: see test_source_sampling_matches_oracle
        dv = np.abs(got[f_].astype(np.float64) - want[f_])
        assert dv[0::2].max() < 2e-5 and (dv[1::2] > 2e-5).mean() < 1e-4 and dv[1::2].max() < 1e-3
    assert np.allclose(got["t"], want["t"], rtol=1e-12, atol=1e-6)
    # the range really moved the annihilation points: mm scale for O-15 in water, both photons of a pair share the point
    moved = np.sqrt((want["x"] - plain["x"]) ** 2 + (want["y"] - plain["y"]) ** 2 + (want["z"] - plain["z"]) ** 2)
    inside = np.abs(plain["x"]) < 0.5
    assert 0.02 < np.median(moved[inside]) < 0.3
    assert np.array_equal(got["x"][0::2], got["x"][1::2]) and np.array_equal(got["z"][0::2], got["z"][1::2])
    s.close()


@needs_tables
@pytest.mark.parametrize("use_prange", [0, 1])
def test_positron_psf_matches_oracle_and_runs_end_to_end(tmp_path, use_prange):
    # S6 setPositionForPhoton (gPET_kernals.cu:563-604): positron phase space -> photon pairs
    mat, den = water_box()
    s = parity.Setup(0, phantom=(mat, den), size=2.0)
    rng = np.random.default_rng(31)
    n = 30000
    ct = rng.uniform(-1, 1, n); phi = rng.uniform(0, 2 * np.pi, n); st = np.sqrt(1 - ct * ct)
    psf = tmp_path / "positrons.dat"
    refio.write_psf(psf, rng.uniform(-0.6, 0.6, n), rng.uniform(-0.6, 0.6, n), rng.uniform(-0.6, 0.6, n),
                    1.0 + np.arange(n) * 2.0, st * np.cos(phi), st * np.sin(phi), ct, rng.uniform(5e4, 1.5e6, n))
    c = s.ctx
    c.set_transport(noncollinearity_rad=0.0037056, use_positron_range=use_prange, eabs_eV=s.eabs, nsurface=1,
                    surface=list(s.surfaces), record_hits=1)
    c.load_psf(psf, 0, ptype=0)
    assert c.num_psf() == n
    c.stage_psf(0, n)
    got = c.fetch_photons(0)
    pos = np.zeros(n, api.PHOTON_DTYPE)
    rec = refio.read_psf(psf)
    for col, f_ in enumerate(("x", "y", "z", "t", "vx", "vy", "vz", "E")):   # record layout of initialize.cu:79-104
        pos[f_] = rec[:, col]
    want = orc.psf_positron(pos, 0, s.den, s.offset, s.size, 0.0037056, use_prange, s.seed)
    assert got.size == want.size == 2 * n
    assert np.array_equal(got["parn"], want["parn"]) and np.array_equal(got["eventid"], np.repeat(np.arange(n), 2))
    d = np.sqrt((got["x"] - want["x"]) ** 2 + (got["y"] - want["y"]) ** 2 + (got["z"] - want["z"]) ** 2)
    assert (d < 2e-4).mean() > 0.995
    for f_ in ("vx", "vy", "vz"):
        assert np.allclose(got[f_], want[f_], atol=2e-5)
    # both photons of a pair carry the positron's time (the reference drops the second one: consciously fixed)
    assert np.array_equal(got["t"][0::2], got["t"][1::2]) and np.all(got["t"] > 0)
    cosang = got["vx"][0::2] * got["vx"][1::2] + got["vy"][0::2] * got["vy"][1::2] + got["vz"][0::2] * got["vz"][1::2]
    assert np.all(cosang < -0.999)
    # whole path in PSF mode (simulateParticle, gPET.cu:13-199)
    st_ = c.run(None)
    assert st_.pairs == n and st_.singles > 0 and st_.frames == 1
    s.close()


# ------------------------------------------------------------------------------------------------ round 2: exchange, halo, 64-bit ids
def _as_u8(ev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(ev).view(np.uint8).reshape(-1, 48))


@pytest.mark.parametrize("dead_type,dead_level,world", [(0, 3, 4), (1, 2, 3), (0, 1, 2)])
def test_time_slice_exchange_through_the_cuda_digitizer_equals_one_list(ctx, dead_type, dead_level, world):
    """The multi-GPU data path of SURVEY 8e on one GPU: `world` virtual ranks each hold the events of their share of the
    photons; the events are routed by time slice with the dead-time / coincidence-window halo (gpet_b200.multi), every
    slice goes through the CUDA digitizer with its emit window (device-to-device: gpet_put_events_device), and the union
    of the slices must be the CUDA digitization of all events as ONE list -- singles and coincidences, byte for byte."""
    import torch
    from gpet_b200 import multi
    rng = np.random.default_rng(321)
    T = 5.0e4
    ev = parity.random_events(150000, rng, tmax=T, dead_fraction=0.01)
    p, d = parity.make_digi_params(dead_type=dead_type, dead_level=dead_level, dead_time_us=2.2, coinc_window_us=0.5, coinc_policy=1)
    parity.apply_digi_params(ctx, d)
    want_s, want_counts = ctx.digitize(ev)
    want_c = ctx.fetch_coincidences()
    assert want_c.size > 500 and want_counts[1] > want_counts[2] > 0        # dead time kills, windows fill
    edges = multi.slice_edges(0.0, T, world)
    hb, hf = multi.halo_for(d["dead_time_us"], d["coinc_window_us"])
    per_rank = [_as_u8(ev[ev["parn"] % world == r]) for r in range(world)]
    lists = multi.exchange_local(per_rank, edges, hb, hf)
    got_s, got_c, received = [], [], 0
    for r in range(world):
        dev = lists[r].cuda()
        received += dev.shape[0]
        s, c, flag = multi.digitize_slice(ctx, dev, edges, r, hb)
        assert flag == 0
        got_s.append(s); got_c.append(c)
    assert np.concatenate(got_s).tobytes() == want_s.tobytes()
    assert np.concatenate(got_c).tobytes() == want_c.tobytes()
    alive = int((ev["t"] < 1e19).sum())
    assert alive <= received < 1.05 * alive                                 # halo copies are a small overhead
    # the same through the oracle: the emit-window semantics are the oracle's list cut to the slice
    o_s, _, o_c = orc.digitize(np.ascontiguousarray(lists[1].numpy()).view(api.EVENT_DTYPE).reshape(-1), p)
    lo, hi = edges[1], edges[2]
    assert got_s[1].tobytes() == o_s[(o_s["t"] >= lo) & (o_s["t"] < hi)].astype(api.EVENT_DTYPE).tobytes()
    assert got_c[1].tobytes() == o_c[(o_c["a"]["t"] >= lo) & (o_c["a"]["t"] < hi)].astype(api.COINC_DTYPE).tobytes()
    ctx.clear_emit_window()


def test_a_halo_that_is_too_short_is_reported_not_silently_wrong(ctx):
    """Non-paralyzable dead time with a long tau: chains of kills run across the cut.  A halo shorter than one dead time is
    refused outright; with a halo of two dead times the digitizer cannot always know where a chain started and must then
    say so (emit_counts()[2]) -- flagged or exact, never silently wrong; with a long halo it is exact and says nothing."""
    from gpet_b200 import multi
    rng = np.random.default_rng(5)
    T = 2.0e4
    ev = parity.random_events(9000, rng, tmax=T, nsites=8)                  # 8 modules of panel 0: ~1.3 events per dead time and site
    p, d = parity.make_digi_params(dead_type=1, dead_level=2, dead_time_us=40.0, coinc_window_us=0.5)
    parity.apply_digi_params(ctx, d)
    want_s, counts = ctx.digitize(ev)
    assert counts[1] - counts[2] > 1000                                     # dead time really kills
    edges = multi.slice_edges(0.0, T, 2)
    with pytest.raises(api.GpetError):
        ctx.set_emit_window(float(edges[1]), float(edges[2]), float(edges[1]) - 39.0)
    flags = {}
    for chains in (2.0, 64.0):
        hb, hf = multi.halo_for(d["dead_time_us"], d["coinc_window_us"], chains=chains)
        lists = multi.exchange_local([_as_u8(ev)], edges, hb, hf)
        s1, _, flag = multi.digitize_slice(ctx, lists[1].cuda(), edges, 1, hb)
        flags[chains] = flag
        assert flag or s1.tobytes() == want_s[want_s["t"] >= edges[1]].tobytes(), chains
    assert flags[64.0] == 0
    ctx.clear_emit_window()


@needs_tables
def test_what_frame_cuts_drop_is_counted(tmp_path):
    """gpet_run digitizes frame by frame, as the reference digitizes epoch by epoch (gPET.cu:385-424): dead time and
    coincidence windows do not reach across a cut.  The loss is measured here against the digitization of the union of
    all frames' events as one list (the adder.dat of the run through gpet_digitize): a handful of records per cut."""
    ex = make_example_dir(tmp_path, source="source.txt", window="0 30")
    with api.Context(0) as c:
        c.set_capacity(1 << 17, 1 << 19, 1 << 18)
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        c.set_digitizer(coinc_window_us=0.01)
        st = c.run(ex / "output")
        per_frame_singles = c.result_singles().copy()
        per_frame_co = c.result_coincidences().copy()
        adder = refio.read_events(ex / "output" / "adder.dat")
        one_list, counts = c.digitize(adder)                               # same blur streams: keyed by photon and site
        one_co = c.fetch_coincidences()
    assert st.frames >= 5
    cuts = st.frames - 1
    # singles: an event right after a cut may have a killer before it; coincidences: a pair may straddle a cut
    extra_singles = per_frame_singles.size - one_list.size
    lost_co = one_co.size - per_frame_co.size
    assert 0 <= extra_singles <= 2 * cuts and 0 <= lost_co <= 2 * cuts
    key = lambda e: set(zip(e["parn"].tolist(), e["t"].tolist()))           # noqa: E731
    assert key(one_list) <= key(per_frame_singles)
    print(f"frame cuts: {cuts}; singles kept that a cut-free dead time kills: {extra_singles}; coincidences lost at cuts: {lost_co} "
          f"of {one_co.size}")


@needs_tables
def test_history_numbers_beyond_32_bits(tmp_path):
    """gpet_set_first_pair: an acquisition whose pairs are numbered from 2^33 + 12345 on (the 1e10-decay regime; the
    reference stops at 32-bit atom and thread numbers, gPET.h:50).  Every Philox stream is keyed by the 64-bit index: the
    CUDA path agrees with the oracle per photon, the records carry the low 31 bits, and the run differs from the same
    acquisition numbered from 0 (no stream reuse)."""
    base = (1 << 33) + 12345
    s = parity.Setup(0, phantom="cylinder", n=32)
    c = s.ctx
    c.load_isotopes(parity.EXAMPLE / "data" / "isotopes.txt")
    c.load_source(parity.EXAMPLE / "input" / "source.txt")
    c.set_time_window(0, 2)
    out = {}
    for first in (0, base):
        c.set_first_pair(first)
        c.plan_frames(0)
        assert int(c.frame(0)["first_pair"]) == first
        c.stage_front(0)
        c.stage_panel_transport()
        ev = c.fetch_events()
        c.stage_source(0)
        q0 = c.fetch_photons(0)
        c.stage_phantom(); c.stage_detector()
        ev_staged = c.fetch_events()
        assert np.array_equal(np.sort(ev, order=["parn", "t", "cryn"]), np.sort(ev_staged, order=["parn", "t", "cryn"]))
        out[first] = (q0, ev)
    q0, ev = out[base]
    npairs = q0.size // 2
    assert npairs > 5000
    assert np.array_equal(q0["eventid"], ((base + np.arange(q0.size) // 2) & 0x7fffffff).astype(np.int32))
    assert np.array_equal(q0["parn"], ((2 * base + np.arange(q0.size)) & 0x7fffffff).astype(np.int32)) and q0["parn"].min() >= 0
    # the oracle with the same index base: per-photon parity through phantom and detector
    orc.set_id_base(2 * base)
    try:
        oph = orc.phantom(q0, s.mat, s.den, s.offset, s.size, s.tab_ph, s.eabs, s.seed)
        res = orc.detector(oph, s.panels, s.counts4, s.pmat, s.pdens, s.surfaces, s.tab_det, s.eabs, 2, 1, s.seed)
    finally:
        orc.set_id_base(0)
    ea = {(int(r["parn"]), int(r["siten"])): float(r["E"]) for r in ev}
    eb = {(int(r["parn"]), int(r["siten"])): float(r["E"]) for r in res["events"]}
    common = set(ea) & set(eb)
    assert len(common) >= 0.99 * len(eb) and sum(abs(ea[k] - eb[k]) <= 2e-3 * eb[k] for k in common) >= 0.99 * len(common)
    # numbered from 0 the same decays take other streams: same statistics, different histories
    ev0 = out[0][1]
    assert abs(ev0.size - ev.size) < 6 * np.sqrt(ev.size) and not np.array_equal(np.sort(ev0["E"]), np.sort(ev["E"]))
    s.close()


@needs_tables
def test_file_run_streams_and_keeps_files_complete(tmp_path):
    """File runs hand their records to a writer thread; the files must be complete and identical to the in-memory results
    when gpet_run returns, frame after frame."""
    ex = make_example_dir(tmp_path, source="source.txt", window="0 20")
    with api.Context(0) as c:
        c.set_capacity(1 << 17, 1 << 19, 1 << 18)
        c.load_config_file(ex / "input_PET.in", base_dir=ex)
        c.set_digitizer(coinc_window_us=0.01)
        st0 = c.run(None)
        s0 = c.result_singles().copy(); co0 = c.result_coincidences().copy(); cls0 = c.result_coincidence_classes().copy()
        st = c.run(ex / "output")
    assert st.frames >= 3 and st.singles == st0.singles
    assert refio.read_events(ex / "output" / "singles.dat").tobytes() == s0.tobytes()
    assert refio.read_coincidences(ex / "output" / "coincidences.dat").tobytes() == co0.tobytes()
    assert np.fromfile(ex / "output" / "coincidences_class.dat", np.uint8).tobytes() == cls0.tobytes()
    assert refio.read_events(ex / "output" / "adder.dat").size == st.events_adder
