"""Shared helpers of the parity tests: build oracle inputs, run the CUDA path through the C ABI, compare.
The oracle (oracle/) is the checker here, never the thing under test or shipped."""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from gpet_b200 import api, refio  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tools import gen_inputs  # noqa: E402

EXAMPLE = ROOT / "examples" / "small_animal"
DATA = ROOT / "gpet_b200" / "_data"
PACKED = DATA / "input4gPET.gpettab"
GOLDEN = ROOT / "tests" / "golden"
SEED = 0x67504554


def have_tables():
    return PACKED.exists()


# ------------------------------------------------------------------------------------------------ digitizer
def make_digi_params(**kw):
    d = dict(readout_depth=2, readout_policy=1, threshold_eV=50000.0, blur_policy=1, blur_Eref=662000.0, blur_Rref=0.0,
             blur_slope=0.0, blur_space=0.0, dead_level=3, dead_type=0, dead_time_us=2.2, ewin_min=30000.0,
             ewin_max=700000.0, time_blur_sigma_us=0.0, coinc_window_us=0.0, coinc_policy=0, coinc_min_panel_diff=0,
             npanels=8, moduleN=117, crystalN=64, seed=SEED, tie_site=0)
    d.update(kw)
    p = orc.DigiParams()
    for k, v in d.items():
        setattr(p, k, v)
    return p, d


def apply_digi_params(ctx, d):
    keys = [f[0] for f in api.DigitizerParams._fields_]
    ctx.set_digitizer(**{k: v for k, v in d.items() if k in keys})


def random_events(n, rng, tmax=4.0e6, nsites=936, npanels=8, moduleN=117, crystalN=64, tie_fraction=0.0, dead_fraction=0.0):
    """adder.dat-like list: (parn, site) unique, random energies around the windows, times uniform in [0, tmax)."""
    ev = np.zeros(n, api.EVENT_DTYPE)
    ev["parn"] = rng.permutation(n).astype(np.int32)
    site = rng.integers(0, nsites, n)
    ev["pann"] = site // moduleN
    ev["modn"] = site % moduleN
    ev["cryn"] = rng.integers(0, crystalN, n)
    ev["siten"] = site
    ev["eventid"] = ev["parn"] // 2
    ev["t"] = rng.uniform(1.0, tmax, n)
    if tie_fraction > 0 and n > 1:
        k = int(n * tie_fraction)
        src = rng.integers(0, n, k); dst = rng.integers(0, n, k)
        ev["t"][dst] = ev["t"][src]
    ev["E"] = rng.choice([20000.0, 45000.0, 200000.0, 511000.0, 650000.0, 900000.0, 2.5e6], n).astype(np.float32) * \
        rng.uniform(0.9, 1.1, n).astype(np.float32)
    ev["x"] = rng.uniform(-1, 0, n); ev["y"] = rng.uniform(-8, 8, n); ev["z"] = rng.uniform(-11, 11, n)
    if dead_fraction > 0:
        ev["t"][rng.random(n) < dead_fraction] = 1e20
    return ev


def events_equal(a, b):
    return a.size == b.size and a.tobytes() == b.tobytes()


def kat_cases():
    j = json.loads((GOLDEN / "digitizer_kat.json").read_text())
    for c in j["cases"]:
        d = dict(j["defaults"]); d.update(c["params"])
        ev = np.zeros(len(c["events"]), api.EVENT_DTYPE)
        for i, r in enumerate(c["events"]):
            ev[i] = (r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], 0, 0, 0)
        yield c, d, ev


def check_kat(case, singles, counts, coinc):
    assert [int(x) for x in singles["parn"]] == case["expected"], case["name"]
    assert [int(x) for x in counts] == case["counts"], case["name"]
    if "expected_siten" in case:
        assert [int(x) for x in singles["siten"]] == case["expected_siten"], case["name"]
    if "coincidences" in case:
        got = [[int(c["a"]["parn"]), int(c["b"]["parn"])] for c in coinc]
        assert got == case["coincidences"], case["name"]
    assert np.all(np.diff(singles["t"]) >= 0)


# ------------------------------------------------------------------------------------------------ transport setup
class Setup:
    """Everything both sides need for transport parity: the product context and the oracle's views of the same inputs,
    the latter built from the independent numpy parsers (gpet_b200.refio) + the product's packed tables."""

    def __init__(self, device=0, phantom="cylinder", n=64, size=1.0, geo=None, capacity=None, seed=SEED):
        self.ctx = api.Context(device)
        c = self.ctx
        c.set_seed(seed)
        self.seed = seed
        if capacity:
            c.set_capacity(*capacity)
        c.load_tables(PACKED)
        geo = geo or (EXAMPLE / "input" / "config8.geo")
        c.load_geometry(geo)
        if isinstance(phantom, str) and phantom == "cylinder":
            mat, den = gen_inputs.cylinder_phantom(n=n, size=size, radius=size / 2)
        elif isinstance(phantom, str) and phantom == "air":
            mat, den = gen_inputs.air_phantom(n)
        else:
            mat, den = phantom
        self.mat, self.den = mat, den
        sz = np.broadcast_to(np.asarray(size, np.float64), (3,))
        self.offset = (-sz / 2).astype(np.float32)
        self.size = sz.astype(np.float32)
        c.set_phantom(mat, den, self.offset, self.size)
        # oracle side
        self.panels, self.pmat, self.pdens, self.counts4 = refio.parse_geometry(geo)
        d = c.table_dims()
        nm, ne = d["nmat"], d["nen"]
        self.energy = c.table(8)
        lamph = c.table(0).reshape(nm, ne); compt = c.table(1).reshape(nm, ne); rayle = c.table(3).reshape(nm, ne)
        cmpsf = c.table(4).reshape(nm, d["cm_ncp"], d["cm_ne"]); rayff = c.table(5).reshape(nm, d["rl_ncp"], d["rl_ne"])
        maxden_ph = np.zeros(nm, np.float32)
        for m in np.unique(mat):
            maxden_ph[m] = den[mat == m].max()
        maxden_det = np.zeros(nm, np.float32)
        for i in range(2):
            maxden_det[self.pmat[i]] = max(maxden_det[self.pmat[i]], self.pdens[i])
        self.tab_ph = orc.TableSet(self.energy, lamph, compt, rayle, cmpsf, d["cm_dcp"], d["cm_de"], rayff, d["rl_dcp"],
                                   d["rl_de"], orc.build_majorant(lamph, maxden_ph))
        self.tab_det = self.tab_ph.with_majorant(orc.build_majorant(lamph, maxden_det))
        self.eabs = 1000.0
        self.surfaces = np.array([0, 0, 0, 0, 0, 0, 0, 0, 0, 1], np.float32)
        c.set_transport(noncollinearity_rad=0.0037056, eabs_eV=self.eabs, nsurface=1, surface=list(self.surfaces), record_hits=1)

    def close(self):
        self.ctx.close()


def isotropic_photons(n, rng, pos_sigma=0.0, energy=511000.0, t0=1.0):
    ph = np.zeros(n, api.PHOTON_DTYPE)
    ct = rng.uniform(-1, 1, n); phi = rng.uniform(0, 2 * np.pi, n); st = np.sqrt(1 - ct * ct)
    ph["vx"] = st * np.cos(phi); ph["vy"] = st * np.sin(phi); ph["vz"] = ct
    if pos_sigma > 0:
        ph["x"] = rng.normal(0, pos_sigma, n); ph["y"] = rng.normal(0, pos_sigma, n); ph["z"] = rng.normal(0, pos_sigma, n)
    ph["E"] = energy
    ph["t"] = t0 + np.arange(n) * 1.0
    ph["eventid"] = np.arange(n) // 2
    ph["parn"] = np.arange(n)
    return ph


def compare_photons(a, b, rtol=2e-4, atol=2e-4):
    """Per-photon comparison of two phase-space lists keyed by parn.  Returns (n_common, n_matching, only_a, only_b)."""
    ia = np.argsort(a["parn"], kind="stable"); ib = np.argsort(b["parn"], kind="stable")
    a, b = a[ia], b[ib]
    common, xa, xb = np.intersect1d(a["parn"], b["parn"], return_indices=True)
    a2, b2 = a[xa], b[xb]
    ok = np.ones(common.size, bool)
    for f in ("x", "y", "z", "vx", "vy", "vz"):
        ok &= np.isclose(a2[f], b2[f], rtol=rtol, atol=atol)
    ok &= np.isclose(a2["E"], b2["E"], rtol=1e-4)
    ok &= np.isclose(a2["t"], b2["t"], rtol=0, atol=1e-6)
    ok &= a2["nscat"] == b2["nscat"]
    return common.size, int(ok.sum()), a.size - common.size, b.size - common.size


def hits_by_photon(h):
    order = np.lexsort((h["t"], h["parn"]))
    return h[order]


def chi2_hist(a, b, bins):
    """Two-sample chi-square statistic / ndf between samples a and b on common bins."""
    ha, _ = np.histogram(a, bins); hb, _ = np.histogram(b, bins)
    m = (ha + hb) > 10
    na, nb = ha.sum(), hb.sum()
    if m.sum() < 2 or na == 0 or nb == 0:
        return 0.0, 0
    k1, k2 = np.sqrt(nb / na), np.sqrt(na / nb)
    chi2 = (((k1 * ha[m] - k2 * hb[m]) ** 2) / (ha[m] + hb[m])).sum()
    return float(chi2), int(m.sum() - 1)


# ------------------------------------------------------------------------------------------------ smoke
def smoke():
    """One small pass of the hot path on cuda:0 against the oracle (used by __graft_entry__.smoke())."""
    rng = np.random.default_rng(1)
    # digitizer replay, bit exact
    ctx = api.Context(0)
    p, d = make_digi_params()
    apply_digi_params(ctx, d)
    ev = random_events(20000, rng, tmax=2.0e4)
    got, counts = ctx.digitize(ev)
    want, wcounts, _ = orc.digitize(ev, p)
    assert events_equal(got, want.astype(api.EVENT_DTYPE)), "digitizer replay differs from the oracle"
    assert list(counts) == list(wcounts)
    ctx.close()
    if not have_tables():
        print("smoke: packed tables missing, transport part skipped")
        return
    # PSF photons -> phantom -> detector -> digitizer
    s = Setup(0, phantom="cylinder", n=32)
    ph = isotropic_photons(20000, rng)
    s.ctx.put_photons(0, ph)
    s.ctx.stage_phantom()
    s.ctx.stage_detector()
    s.ctx.stage_digitize()
    ev_gpu = s.ctx.fetch_events()
    oph = orc.phantom(ph, s.mat, s.den, s.offset, s.size, s.tab_ph, s.eabs, s.seed)
    res = orc.detector(oph, s.panels, s.counts4, s.pmat, s.pdens, s.surfaces, s.tab_det, s.eabs, 2, 1, s.seed)
    n_o, n_g = res["events"].size, ev_gpu.size
    assert abs(n_o - n_g) <= max(20, 0.01 * n_o), f"event counts differ: oracle {n_o} gpu {n_g}"
    singles = s.ctx.fetch_singles()
    assert singles.size > 0 and np.all(np.diff(singles["t"]) >= 0)
    print(f"smoke: {n_g} events (oracle {n_o}), {singles.size} singles")
    s.close()
