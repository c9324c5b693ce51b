"""Unit tests of the CPU oracle's file-local functions, reached through tests/c/oracle_units.c: against physics where a
closed form exists (Klein-Nishina, rotation geometry) and against plain-Python restatements of the reference lines
(crystalSearch, adder, readout).  CPU only; the oracle is the thing under test here."""
import ctypes as C
import subprocess
from fractions import Fraction

import numpy as np
import pytest

import parity
from gpet_b200 import refio
from oracle import oracle as orc

f32 = np.float32
MC2 = 510.9991e3
MAXT = 1e20


@pytest.fixture(scope="module")
def units(tmp_path_factory):
    so = tmp_path_factory.mktemp("units") / "liborc_units.so"
    subprocess.run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-Wno-unused-function",
                    "-o", str(so), str(parity.ROOT / "tests" / "c" / "oracle_units.c"), "-lm"], check=True)
    return C.CDLL(str(so))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------------------------------------ P6 rotate
def test_rotate_turns_by_the_polar_angle_and_keeps_the_norm(units):
    """rotate (gPET_kernals.cu:172-254): the new direction has unit length and makes the angle theta with the old one;
    for a fixed direction, phi sweeps the cone uniformly (the azimuth of the new direction around the old one equals phi
    up to a constant)."""
    rng = np.random.default_rng(1)
    n = 20000
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:200] *= rng.uniform(0.5, 2.0, (200, 1))            # not normalised on input: renormalised (lines 205-212)
    d[200:210] = [0, 0, 1]; d[210:220] = [0, 0, -1]       # rho = 0: the two special cases (lines 236-252)
    d = d.astype(f32)
    costh = rng.uniform(-1, 1, n).astype(f32)
    costh[:5] = [1, -1, 0, 1, -1]
    phi = rng.uniform(0, 2 * np.pi, n).astype(f32)
    out = d.copy()
    units.u_rotate(_p(out), C.c_int64(n), _p(costh), _p(phi))
    d0 = d.astype(np.float64)
    d0 /= np.linalg.norm(d0, axis=1, keepdims=True)
    o = out.astype(np.float64)
    unit = np.ones(n, bool); unit[:200] = False
    assert np.abs(np.linalg.norm(o[unit], axis=1) - 1).max() < 2e-4   # the reference itself tolerates 1e-4 before it renormalises
    assert np.abs((o[unit] * d0[unit]).sum(1) - costh[unit]).max() < 2e-4
    # Not normalised on input: the reference renormalises (u, v, w) but goes on with the rho2 of the vector as it came in
    # (gPET_kernals.cu:203-212 vs 222), so the result is NOT a unit vector, whatever the header comment says.  Transport
    # never gets there (directions stay within 1e-4 of unit length); the oracle keeps the statement order all the same.
    u, v, w = d[:200, 0].copy(), d[:200, 1].copy(), d[:200, 2].copy()
    rho2 = u * u + v * v
    norm = f32(1) / np.sqrt(rho2 + w * w)
    u, v, w = u * norm, v * norm, w * norm
    sp, cp = np.sin(phi[:200]), np.cos(phi[:200])
    c2 = costh[:200] * costh[:200]
    sthrho = np.where(c2 < 1, np.sqrt(np.maximum(f32(1) - c2, 0) / rho2), f32(0)).astype(f32)
    urho, vrho = u * sthrho, v * sthrho
    lit = np.stack([u * costh[:200] - vrho * sp + w * urho * cp, v * costh[:200] + urho * sp + w * vrho * cp,
                    w * costh[:200] - rho2 * sthrho * cp], 1)
    assert np.abs(out[:200] - lit).max() < 2e-6 and np.abs(np.linalg.norm(o[:200], axis=1) - 1).max() > 0.05
    # azimuth: same direction and polar angle, phi and phi + pi/2 -> the two results are perpendicular around the axis
    k = 5000
    a = np.repeat(d[1000:1001], k, 0).copy(); b = a.copy()
    ct = np.full(k, 0.3, f32); ph = rng.uniform(0, 2 * np.pi, k).astype(f32)
    units.u_rotate(_p(a), C.c_int64(k), _p(ct), _p(ph))
    units.u_rotate(_p(b), C.c_int64(k), _p(ct), _p((ph + f32(np.pi / 2)).astype(f32)))
    axis = d0[1000]
    pa = a - np.outer(a @ axis, axis); pb = b - np.outer(b @ axis, axis)
    assert np.abs((pa * pb).sum(1)).max() < 1e-4 * (1 - 0.09)       # perpendicular transverse parts
    assert np.abs(np.einsum("ij,ij->i", np.cross(pa, pb), np.tile(axis, (k, 1))) - (1 - 0.09)).max() < 1e-3   # right-handed, |.| = sin^2


# ------------------------------------------------------------------------------------------------ X3 Klein-Nishina
@pytest.mark.parametrize("E", [60e3, 140e3, 511e3, 1.0e6])
def test_compton_sampler_follows_klein_nishina(units, E):
    """comsam, free electron (gPET_kernals.cu:90-126): the sampled energy fraction follows
    dsigma/d(eps) ~ (1/eps + eps) (1 - eps sin^2(theta) / (1 + eps^2)) on [1/(1+2k), 1], and cos(theta) is Compton's
    relation for that fraction."""
    n = 400000
    ef = np.zeros(n, f32); ct = np.zeros(n, f32)
    units.u_compton_kn(C.c_float(E), C.c_uint64(77), C.c_int64(n), _p(ef), _p(ct))
    k = E / MC2
    emin = 1.0 / (1.0 + 2.0 * k)
    assert ef.min() >= emin * (1 - 1e-5) and ef.max() <= 1.0
    assert np.abs(ct - (1.0 - (1.0 - ef.astype(np.float64)) / (ef.astype(np.float64) * k))).max() < 2e-4
    edges = np.linspace(emin, 1.0, 41)
    xs = np.linspace(emin, 1.0, 40 * 400 + 1)
    c = 1.0 - (1.0 - xs) / (xs * k)
    pdf = (1.0 / xs + xs) * (1.0 - xs * (1.0 - c * c) / (1.0 + xs * xs))
    cdf = np.concatenate([[0.0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(xs))])
    expect = np.diff(cdf[::400]) / cdf[-1] * n
    got, _ = np.histogram(ef, edges)
    chi2 = ((got - expect) ** 2 / expect).sum()
    assert chi2 / 39 < 1.8, (E, chi2)


# ------------------------------------------------------------------------------------------------ X2 crystalSearch
def crystal_search_literal(p, moduleNy, crystalNy, surfaces, x, y, z):
    """gPET_kernals.cu:1236-1279 in float32, statement by statement.  Returns (m_id, M_id, L_id), -1 = not reached."""
    x, y, z = f32(x), f32(y), f32(z)
    for s in surfaces:
        s = [f32(v) for v in s]
        q = (s[0] * x * x + s[1] * y * y + s[2] * z * z + s[3] * x * y + s[4] * x * z + s[5] * y * z + s[6] * x + s[7] * y + s[8] * z + s[9])
        if q < 0:
            return 1, -1, -1
    y = f32(p["lengthy"] / f32(2)) + y
    z = f32(p["lengthz"] / f32(2)) + z
    py, pz = f32(p["MODy"] + p["Mspacey"]), f32(p["MODz"] + p["Mspacez"])
    My = int(f32(y / py)) if np.floor(f32(y / py)) > 0 else 0
    Mz = int(f32(z / pz)) if np.floor(f32(z / pz)) > 0 else 0
    M = Mz * moduleNy + My
    y = f32(y - f32(f32(My) * py)); z = f32(z - f32(f32(Mz) * pz))
    if y > p["MODy"] or z > p["MODz"]:
        return 1, M, -1
    cy, cz = f32(p["LSOy"] + p["spacey"]), f32(p["LSOz"] + p["spacez"])
    Ly = int(f32(y / cy)) if np.floor(f32(y / cy)) > 0 else 0
    Lz = int(f32(z / cz)) if np.floor(f32(z / cz)) > 0 else 0
    L = Lz * crystalNy + Ly
    y = f32(y - f32(f32(Ly) * cy)); z = f32(z - f32(f32(Lz) * cz))
    if y > p["LSOy"] or z > p["LSOz"]:
        return 1, M, L
    return 0, M, L


def test_crystal_search_equals_a_literal_walk_of_the_reference(units):
    panels, mat, dens, counts = refio.parse_geometry(parity.EXAMPLE / "input" / "config8.geo")
    moduleNy, crystalNy, moduleN, crystalN = counts
    p = panels[0]
    rng = np.random.default_rng(2)
    n = 4000
    xyz = np.stack([rng.uniform(-p["lengthx"], 0, n), rng.uniform(-p["lengthy"] / 2, p["lengthy"] / 2, n),
                    rng.uniform(-p["lengthz"] / 2, p["lengthz"] / 2, n)], 1).astype(f32)
    # points on and next to module / crystal borders
    pitch = f32(p["MODy"] + p["Mspacey"])
    xyz[:200, 1] = (-p["lengthy"] / 2 + pitch * rng.integers(0, moduleNy, 200) + rng.choice([0.0, 1e-6, -1e-6, float(p["MODy"])], 200)).astype(f32)
    excluded = []
    for surfaces in ([], [[0, 0, 0, 0, 0, 0, 0, 0, 0, 1]], [[0, 1, 1, 0, 0, 0, 0, 0, 0, -4.0]]):
        sf = np.asarray(surfaces, f32).ravel() if surfaces else np.zeros(10, f32)
        ids = np.zeros((n, 3), np.int32)
        pc = np.ascontiguousarray(panels[:1])
        units.u_crystal_search(_p(pc), C.c_int(moduleNy), C.c_int(crystalNy), C.c_int(len(surfaces)), _p(sf), C.c_int64(n), _p(xyz), _p(ids))
        want = np.array([crystal_search_literal(p, moduleNy, crystalNy, surfaces, *xyz[i]) for i in range(n)], np.int32)
        # the ids the reference leaves unset on an early return are the caller's previous values: compare what is defined
        assert np.array_equal(ids[:, 0], want[:, 0])
        reached_M = want[:, 1] >= 0
        assert np.array_equal(ids[reached_M, 1], want[reached_M, 1])
        reached_L = want[:, 2] >= 0
        assert np.array_equal(ids[reached_L, 2], want[reached_L, 2])
        assert (want[:, 0] == 0).sum() > 0.3 * n and ids[:, 1].max() < moduleN and ids[:, 2].max() < crystalN
        excluded.append(int((want[:, 0] == 1).sum()))
    # the inert surface of the shipped input changes nothing; y^2 + z^2 - 4 < 0 takes the crystals within 2 cm of the x axis out
    assert excluded[0] == excluded[1] < excluded[2] and excluded[2] - excluded[0] > 50


# ------------------------------------------------------------------------------------------------ D1 adder, D2 readout
def _fma32(a, b, c):
    """float32 fma(a, b, c): exact product and sum, one rounding (the contraction is spelled out)."""
    return f32(float(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c))))


def _centroid(x0, e0, x1, e1):
    return f32(_fma32(x0, e0, f32(x1 * e1)) / f32(e0 + e1))


def adder_readout_literal(hits, depth, policy, moduleN, cap=6):
    """adder (gPET_kernals.cu:737-755) over the hits of one photon, then readout (:756-813)."""
    ev = []
    for h in hits:
        h = h.copy()
        for e in ev:
            if e["siten"] == h["siten"]:
                for ax in "xyz":
                    e[ax] = _centroid(e[ax], e["E"], h[ax], h["E"])
                e["E"] = f32(e["E"] + h["E"])
                break
        else:
            if len(ev) < cap:
                ev.append(h)
    if depth == 3 or not ev:
        return ev
    if policy == 1:
        depth = 2
    for e in ev:
        e["siten"] = 0 if depth == 0 else e["pann"] if depth == 1 else e["pann"] * moduleN + e["modn"]
    out = []
    for i in range(len(ev)):
        e0 = ev[i].copy()
        if e0["t"] > MAXT * 0.1:
            continue
        for j in range(i + 1, len(ev)):
            e = ev[j]
            if e["t"] > MAXT * 0.1:
                continue
            if e["parn"] == e0["parn"] and e["siten"] == e0["siten"]:
                if policy == 1:
                    for ax in "xyz":
                        e0[ax] = _centroid(e0[ax], e0["E"], e[ax], e["E"])
                    e0["E"] = f32(e0["E"] + e["E"])
                    e["t"] = MAXT
                    continue
                e0 = e0 if e0["E"] > e["E"] else e.copy()
                e["t"] = MAXT
        out.append(e0)
    return out


@pytest.mark.parametrize("policy", [0, 1])
@pytest.mark.parametrize("depth", [0, 1, 2, 3])
def test_adder_and_readout_equal_a_literal_walk_of_the_reference(units, depth, policy):
    rng = np.random.default_rng(10 + 2 * depth + policy)
    moduleN, crystalN = 117, 64
    nout_seen = set()
    for trial in range(300):
        n = int(rng.integers(1, 9))
        hits = np.zeros(n, orc.EVENT_DTYPE)
        hits["parn"] = 4242
        hits["pann"] = 3
        hits["modn"] = rng.choice([5, 6, 7], n)                     # few modules / crystals: merges at every level
        hits["cryn"] = rng.choice([1, 2, 3], n)
        hits["siten"] = (hits["pann"] * moduleN + hits["modn"]) * crystalN + hits["cryn"]
        hits["eventid"] = 2121
        hits["t"] = 100.0 + np.sort(rng.uniform(0, 1e-3, n))
        hits["E"] = rng.choice([30e3, 120e3, 120e3, 341e3], n).astype(f32) * rng.choice([1.0, 1.0, 0.7], n).astype(f32)
        for ax in "xyz":
            hits[ax] = rng.uniform(-2, 2, n)
        out = np.zeros(8, orc.EVENT_DTYPE)
        ovf = C.c_int()
        nout = units.u_adder_readout(_p(hits), C.c_int(n), C.c_int(depth), C.c_int(policy), C.c_int(moduleN), _p(out), C.byref(ovf))
        want = adder_readout_literal([dict(zip(hits.dtype.names, h)) for h in hits.tolist()], depth, policy, moduleN)
        assert nout == len(want), (trial, nout, len(want))
        for k, w in enumerate(want):
            got = dict(zip(out.dtype.names, out[k].tolist()))
            for name in out.dtype.names:
                assert got[name] == (float(w[name]) if name in "Exyzt" else w[name]), (trial, k, name, got[name], w[name])
        nout_seen.add(nout)
    assert len(nout_seen) >= 3 or (depth <= 1 and policy == 0)      # one panel: world and panel level leave one event


# ------------------------------------------------------------------------------------------------ S2, S3 source sampling
def _chi2_uniform(u, bins=40):
    got, _ = np.histogram(u, bins, (0.0, 1.0))
    e = u.size / bins
    return ((got - e) ** 2 / e).sum() / (bins - 1)


def test_source_oracle_samples_what_setposition_specifies():
    """orc_source against the physics setPosition / getPositionFromShape state (gPET_kernals.cu:445-561): positions
    uniform in box / cylinder / sphere, decay times from the exponential truncated to the frame, photon 1 isotropic,
    photon 2 turned by a Gaussian acollinearity with E = mc2 -+ delta mc2 / 2, ids and times shared as the reference does."""
    n = 300000
    shapes = [0, 1, 2]
    coeff = np.array([[1.0, -2.0, 0.5, 0.4, 0.6, 0.8], [-3.0, 1.0, 2.0, 0.7, 1.5, 0.0], [4.0, 4.0, -1.0, 0.9, 0.0, 0.0]], np.float32)
    cum = np.array([n // 3, 2 * n // 3, n], np.uint64)
    tau = np.array([6586.26 * 1.442695, 122.24 * 1.442695, 1223.4 * 1.442695])   # F-18, O-15, C-11 mean lives (s)
    dt = 90.0
    frac = 1.0 - np.exp(-dt / tau)
    sigma = 0.0037056
    first = 5_000_000_000                                                        # photon numbers wrap modulo 2^32
    ph = orc.source(cum, shapes, coeff.ravel(), tau, frac, 30.0, first, sigma, n, 99)
    a, b = ph[0::2], ph[1::2]
    assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["t"], b["t"]) and np.array_equal(a["eventid"], b["eventid"])
    k = np.arange(n, dtype=np.uint64)
    assert np.array_equal(a["eventid"].view(np.uint32), ((first + k) & 0xFFFFFFFF).astype(np.uint32))
    assert np.array_equal(a["parn"].view(np.uint32), ((2 * (first + k)) & 0xFFFFFFFF).astype(np.uint32))
    assert np.array_equal(b["parn"].view(np.uint32), ((2 * (first + k) + 1) & 0xFFFFFFFF).astype(np.uint32))
    assert np.all(a["nscat"] == 0)
    lo = 0
    for s, hi in enumerate(cum.astype(int)):
        q = a[lo:hi]
        c = coeff[s]
        dx, dy, dz = q["x"] - c[0], q["y"] - c[1], q["z"] - c[2]
        if shapes[s] == 0:       # box: centre and full lengths
            for d, full in ((dx, c[3]), (dy, c[4]), (dz, c[5])):
                assert np.abs(d).max() <= full / 2 * (1 + 1e-6) and _chi2_uniform(d / full + 0.5) < 1.7
        elif shapes[s] == 1:     # cylinder along z: radius, height
            r2 = (dx * dx + dy * dy) / c[3] ** 2
            assert r2.max() <= 1 + 1e-5 and _chi2_uniform(r2) < 1.7 and _chi2_uniform(dz / c[4] + 0.5) < 1.7
            assert _chi2_uniform(np.arctan2(dy, dx) / (2 * np.pi) + 0.5) < 1.7
        else:                    # sphere: radius (cbrtf of a uniform)
            r3 = (dx * dx + dy * dy + dz * dz) ** 1.5 / c[3] ** 3
            assert r3.max() <= 1 + 1e-5 and _chi2_uniform(r3) < 1.7
            assert _chi2_uniform(dz / np.sqrt(dx * dx + dy * dy + dz * dz) / 2 + 0.5) < 1.7
        # decay time: inverse transform of the exponential truncated to [0, dt)
        p = q["t"] * 1e-6 - 30.0
        assert p.min() >= 0 and p.max() < dt
        assert _chi2_uniform((1.0 - np.exp(-p / tau[s])) / frac[s]) < 1.7
        lo = hi
    # photon 1 isotropic
    assert np.abs(np.sqrt(a["vx"] ** 2 + a["vy"] ** 2 + a["vz"] ** 2) - 1).max() < 1e-5
    assert _chi2_uniform(a["vz"] / 2 + 0.5) < 1.7 and _chi2_uniform(np.arctan2(a["vy"], a["vx"]) / (2 * np.pi) + 0.5) < 1.7
    # photon 2: angle pi - |delta| to photon 1, delta ~ N(0, sigma); energies mc2 +- delta mc2 / 2
    da = (a["E"].astype(np.float64) - MC2) / (MC2 / 2)
    db = (b["E"].astype(np.float64) - MC2) / (MC2 / 2)
    assert np.abs(da + db).max() < 1e-6 * 4 and abs(da.mean()) < 4 * sigma / np.sqrt(n)
    assert abs(da.std() / sigma - 1) < 0.01
    cosang = (a["vx"].astype(np.float64) * b["vx"] + a["vy"].astype(np.float64) * b["vy"] + a["vz"].astype(np.float64) * b["vz"])
    cross = np.linalg.norm(np.cross(np.stack([a["vx"], a["vy"], a["vz"]], 1).astype(np.float64),
                                    np.stack([b["vx"], b["vy"], b["vz"]], 1).astype(np.float64)), axis=1)
    assert cosang.max() < -0.999                                   # back to back
    assert abs(np.sqrt((cross ** 2).mean()) / sigma - 1) < 0.01    # sin|delta| ~ |delta|
    # the same delta turns the photon and splits the energy -- up to the reference's fp32 rotate(-cos(delta)): cos(delta)
    # is 1 - delta^2/2 rounded to a float, which resolves angles only to sqrt(2 * 2^-24) = 3.5e-4 rad near zero
    assert np.abs(cross - np.abs(da)).max() < 4e-4 and (cross == 0).mean() > 0.02


# ------------------------------------------------------------------------------------------------ P1 Woodcock tracking
@pytest.mark.skipif(not parity.have_tables(), reason="packed tables not built")
@pytest.mark.parametrize("E", [511e3, 140e3])
def test_phantom_oracle_attenuates_like_beer_lambert_through_layers(E):
    """photon() (gPET_kernals.cu:256-345) is Woodcock tracking: whatever the majorant, the photons that cross water,
    bone and air without a real interaction must be exp(-sum mu_i L_i) of the beam, mu_i = Sigma_tot(E) rho_i from the
    same tables.  Host-only context: tables and majorants come from the product's loaders, the walk is the oracle's."""
    n = 64
    mat = np.zeros((n, n, n), np.int32); den = np.full((n, n, n), 1.2048e-3, np.float32)      # air, x fastest
    mat[:, :, :22] = 1; den[:, :, :22] = 1.0                                                  # water   x < -1.0
    mat[:, :, 22:37] = 3; den[:, :, 22:37] = 1.85                                             # bone    -1.0 <= x < 0.5
    s = parity.Setup(-1, phantom=(mat, den), size=6.4)
    nph = 400000
    ph = np.zeros(nph, orc.PHOTON_DTYPE)
    ph["x"] = -3.0; ph["y"] = 0.05; ph["z"] = 0.05
    ph["vx"] = 1.0
    ph["E"] = E
    ph["t"] = 1.0 + np.arange(nph)
    ph["parn"] = np.arange(nph); ph["eventid"] = np.arange(nph) // 2
    out = orc.phantom(ph, s.mat, s.den, s.offset, s.size, s.tab_ph, s.eabs, 4321)
    alive = out["t"] > 0
    straight = alive & (out["nscat"] == 0)
    assert np.all(out["E"][straight] == np.float32(E)) and np.all(out["vx"][straight] == 1.0)
    assert np.all(out["x"][straight] > 3.1)                     # they left through the far face (overshoot kept, quirk 2)
    # time of flight of the straight ones: path / c
    path = out["x"][straight].astype(np.float64) + 3.0
    assert np.abs(out["t"][straight] - ph["t"][straight] - path / 29979.2458).max() < 1e-6
    energy = s.energy.astype(np.float64)
    dims = s.ctx.table_dims()
    lam = np.array([np.interp(E, energy, row.astype(np.float64)) for row in s.ctx.table(0).reshape(dims["nmat"], dims["nen"])])
    mu = lam[1] * 1.0 * 2.0 + lam[3] * 1.85 * 1.5 + lam[0] * 1.2048e-3 * 2.7
    want = np.exp(-mu)
    got = straight.mean()
    assert abs(got - want) < 4 * np.sqrt(want * (1 - want) / nph) + 2e-4, (E, got, want)
    # the interacting rest: scattered (alive, nscat > 0) or photo-absorbed (dead); at 511 keV in water and bone nearly all scatter
    absorbed = (~alive).mean()
    assert (absorbed < 0.02) if E > 300e3 else (0.01 < absorbed < 0.2)
    s.close()


# ------------------------------------------------------------------------------------------------ X1 detector transport
@pytest.mark.skipif(not parity.have_tables(), reason="packed tables not built")
def test_detector_oracle_stops_photons_like_beer_lambert_in_one_crystal():
    """photonde (gPET_kernals.cu:839-1233): a beam down the axis of one crystal stays in crystal material over the panel's
    whole depth.  Compton and photoelectric interactions are recorded as hits, Rayleigh ones only turn the photon, so the
    first hits that still lie ON the beam axis are the photons whose first interaction of any kind was a recorded one:
    (1 - f_Rayleigh) (1 - exp(-mu L)) of the beam, at depths distributed as exp(-mu x), Compton : photoelectric as the
    partial cross sections; the deposited energies of a photon never add up to more than it brought."""
    s = parity.Setup(-1, phantom="air", n=8)
    p = s.panels[0]
    moduleNy, crystalNy = int(s.counts4[0]), int(s.counts4[1])
    ly = -p["lengthy"] / 2 + 4 * (p["MODy"] + p["Mspacey"]) + 3 * (p["LSOy"] + p["spacey"]) + p["LSOy"] / 2     # centre of a crystal
    lz = -p["lengthz"] / 2 + 6 * (p["MODz"] + p["Mspacez"]) + 2 * (p["LSOz"] + p["spacez"]) + p["LSOz"] / 2
    ux = np.array([p["UniXx"], p["UniXy"], p["UniXz"]], np.float64); uy = np.array([p["UniYx"], p["UniYy"], p["UniYz"]], np.float64)
    uz = np.array([p["UniZx"], p["UniZy"], p["UniZz"]], np.float64)
    o = np.array([p["offsetx"], p["offsety"], p["offsetz"]], np.float64)
    depth_dir = ux * float(p["directionx"])
    start = o - 0.05 * depth_dir + ly * uy + lz * uz
    nph = 200000
    ph = np.zeros(nph, orc.PHOTON_DTYPE)
    ph["x"], ph["y"], ph["z"] = start
    ph["vx"], ph["vy"], ph["vz"] = depth_dir
    ph["E"] = 511e3
    ph["t"] = 1.0 + np.arange(nph)
    ph["parn"] = np.arange(nph); ph["eventid"] = np.arange(nph) // 2
    res = orc.detector(ph, s.panels, s.counts4, s.pmat, s.pdens, s.surfaces, s.tab_det, s.eabs, 2, 1, 2468)
    assert res["entered"] == nph
    hits = res["hits"]
    assert np.all(hits["pann"] == p["panel"])
    assert np.all(np.diff(hits["parn"]) >= 0)                            # hits are appended in photon order
    first = hits[np.r_[True, hits["parn"][1:] != hits["parn"][:-1]]]
    on_axis = first[(np.abs(first["y"] - ly) < 1e-5) & (np.abs(first["z"] - lz) < 1e-5)]
    assert on_axis.size > 0.9 * first.size
    assert np.all(on_axis["modn"] == 6 * moduleNy + 4) and np.all(on_axis["cryn"] == 2 * crystalNy + 3)
    dims = s.ctx.table_dims()
    xs = {k: np.interp(511e3, s.energy.astype(np.float64), s.ctx.table(k).reshape(dims["nmat"], dims["nen"])[s.pmat[0]].astype(np.float64))
          for k in (0, 1, 3)}                                               # total, Compton, Rayleigh (cm^2/g)
    mu = xs[0] * float(s.pdens[0])
    L = float(p["lengthx"])
    f_c, f_r = xs[1] / xs[0], xs[3] / xs[0]
    want = (1 - f_r) * (1 - np.exp(-mu * L))
    got = on_axis.size / nph
    assert abs(got - want) < 4 * np.sqrt(want * (1 - want) / nph) + 2e-4, (got, want)
    depth = np.abs(on_axis["x"].astype(np.float64))
    assert depth.max() <= L * (1 + 1e-6)
    assert _chi2_uniform((1 - np.exp(-mu * depth)) / (1 - np.exp(-mu * L))) < 1.7
    share_c = f_c / (1 - f_r)
    assert abs(np.isin(on_axis["type"], (1, 2)).mean() - share_c) < 4 * np.sqrt(share_c * (1 - share_c) / on_axis.size) + 1e-3
    assert np.all(np.isin(hits["type"], (1, 2, 4)))                          # Rayleigh leaves no hit
    dep = np.bincount(hits["parn"], weights=hits["E"].astype(np.float64), minlength=nph)
    assert dep.max() <= 511e3 * (1 + 1e-6) and (np.abs(dep - 511e3) < 1.0).mean() > 0.3      # full absorption is common in LSO
    s.close()


# ------------------------------------------------------------------------------------------------ D3 blur
@pytest.mark.parametrize("policy", [0, 1])
def test_blur_oracle_has_the_specified_resolution(policy):
    """blur (gPET_kernals.cu:814-837): E' = E + N(0,1) R E / 2.35482 with R = sqrt(Eref/E) Rref (policy 0) or
    Rref + slope (E - Eref) / 1e6 (policy 1); spatial blur N(0, Sblur) per axis; time blur (extension) N(0, sigma_t)."""
    n = 200000
    rng = np.random.default_rng(3)
    ev = np.zeros(n, orc.EVENT_DTYPE)
    ev["parn"] = np.arange(n); ev["eventid"] = np.arange(n) // 2
    ev["siten"] = np.arange(n)                       # one event per site: dead time has nothing to do
    ev["pann"] = 0; ev["modn"] = np.arange(n) % 117; ev["cryn"] = np.arange(n) % 64
    ev["t"] = 10.0 + 5.0 * np.arange(n)
    ev["E"] = np.where(np.arange(n) % 2 == 0, 511e3, 300e3).astype(np.float32)
    Eref, Rref, slope, sblur, tblur = 511e3, 0.12, 0.2, 0.07, 0.3
    p, d = parity.make_digi_params(dead_level=3, threshold_eV=0.0, ewin_min=0.0, ewin_max=2e6, blur_policy=policy, blur_Eref=Eref,
                                   blur_Rref=Rref, blur_slope=slope, blur_space=sblur, time_blur_sigma_us=tblur)
    s, counts, _ = orc.digitize(ev, p)
    assert s.size == n
    s = s[np.argsort(s["parn"])]
    for E0 in (511e3, 300e3):
        m = ev["E"] == np.float32(E0)
        R = np.sqrt(Eref / E0) * Rref if policy == 0 else Rref + slope * (E0 - Eref) / 1e6
        sig = R * E0 / 2.35482
        z = (s["E"][m].astype(np.float64) - E0) / sig
        assert abs(z.mean()) < 4 / np.sqrt(m.sum()) and abs(z.std() - 1) < 0.01
        from math import erf
        u = 0.5 * (1 + np.vectorize(erf)(z / np.sqrt(2)))
        assert _chi2_uniform(u) < 1.7                                   # Gaussian, not just the right width
    for ax in "xyz":
        dxyz = s[ax].astype(np.float64) - ev[ax]
        assert abs(dxyz.mean()) < 4 * sblur / np.sqrt(n) and abs(dxyz.std() / sblur - 1) < 0.01
    dt = s["t"] - ev["t"]
    assert abs(dt.mean()) < 4 * tblur / np.sqrt(n) and abs(dt.std() / tblur - 1) < 0.01
    assert abs(np.corrcoef(dt, s["E"].astype(np.float64) - ev["E"])[0, 1]) < 0.01     # independent draws (Box-Muller sine / cosine)


# ------------------------------------------------------------------------------------------------ S4 positron energy
@pytest.mark.parametrize("row", [0, 1, 2])
def test_positron_energy_oracle_follows_the_fitted_spectrum(units, row):
    """sampleEkPositron (gPET_kernals.cu:420-443): rejection sampling of the total energy under the six-coefficient
    polynomial of data/isotopes.txt (cut at zero and at the stated pdf maximum), returned as kinetic energy in eV."""
    lines = (parity.EXAMPLE / "data" / "isotopes.txt").read_text().splitlines()
    vals = np.array(lines[2 + row].split(), np.float64)
    coef8 = vals[2:10].astype(f32)                     # Emax (MeV, total), pdf maximum, six coefficients
    n = 300000
    ek = np.zeros(n, f32)
    units.u_sample_ek_positron(_p(coef8), C.c_uint64(5), C.c_int64(n), _p(ek))
    emax = float(coef8[0])
    E = ek.astype(np.float64) * 1e-6 + 0.511
    assert E.min() >= 0.511 and E.max() <= emax * (1 + 1e-6)
    xs = np.linspace(0.511, emax, 40 * 200 + 1)
    pdf = np.clip(sum(float(coef8[2 + i]) * xs ** (5 - i) for i in range(6)), 0.0, float(coef8[1]))
    cdf = np.concatenate([[0.0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(xs))])
    expect = np.diff(cdf[::200]) / cdf[-1] * n
    got, _ = np.histogram(E, np.linspace(0.511, emax, 41))
    m = expect > 20
    assert m.sum() > 30 and (((got - expect) ** 2 / np.maximum(expect, 1e-9))[m]).sum() / (m.sum() - 1) < 1.8
    assert 0.2 * (emax - 0.511) < E.mean() - 0.511 < 0.5 * (emax - 0.511)      # a beta spectrum: mean near a third of the end point


# ------------------------------------------------------------------------------------------------ S5 positron range
def test_positron_range_oracle_walks_a_water_equivalent_gaussian_path():
    """setPositronRange (gPET_kernals.cu:347-418): the positron is displaced along its direction by a water-equivalent
    length r = |N_3(0, sigma)|, sigma = Rex / 2, Rex = 0.1 b1 E^2 / (b2 + E) cm (E kinetic, MeV); the geometric length is
    r / rho in a uniform medium, and the density-weighted path is r across a density step (up to the reference's habit
    of weighting a step with the density of the voxel it ENTERS: one voxel of slack)."""
    from scipy.stats import chi2 as chi2_dist
    n, nv = 100000, 64
    rng = np.random.default_rng(8)
    pos = np.zeros(n, orc.PHOTON_DTYPE)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    pos["vx"], pos["vy"], pos["vz"] = d.T.astype(f32)
    pos["x"] = 0.013; pos["y"] = -0.021; pos["z"] = 0.034           # off the voxel borders
    pos["E"] = 0.6e6
    pos["t"] = 1.0 + np.arange(n)
    ekin = 0.6
    sigma = 0.1 * 5.44040782 * ekin * ekin / (0.369516529 + ekin) / 2
    off, size = (-3.2, -3.2, -3.2), (6.4, 6.4, 6.4)
    start = np.array([0.013, -0.021, 0.034])

    def displaced(dens):
        out = orc.psf_positron(pos, 0, dens, off, size, 0.0037056, True, 77)
        a = out[0::2]
        assert np.array_equal(a["x"], out[1::2]["x"]) and np.array_equal(a["t"], pos["t"])     # both photons from the end point, at the positron's time
        return np.stack([a["x"], a["y"], a["z"]], 1).astype(np.float64) - start

    d1 = displaced(np.full((nv, nv, nv), 1.0, f32))
    r = np.linalg.norm(d1, axis=1)
    dir32 = np.stack([pos["vx"], pos["vy"], pos["vz"]], 1).astype(np.float64)
    assert np.abs(np.cross(d1, dir32)).max() < 2e-5 and (d1 * dir32).sum(1).min() > 0          # along the direction
    assert abs((r * r).mean() / (3 * sigma * sigma) - 1) < 0.015
    assert _chi2_uniform(chi2_dist.cdf((r / sigma) ** 2, 3)) < 1.7                              # |N_3(0, sigma)|
    d2 = displaced(np.full((nv, nv, nv), 2.0, f32))
    assert np.abs(np.linalg.norm(d2, axis=1) - r / 2).max() < 2e-4                              # half as far at twice the density
    # density step at x = 0.1: rho 1 below, rho 4 above
    dens = np.full((nv, nv, nv), 1.0, f32)
    dens[:, :, 33:] = 4.0                                                                       # voxel 33 starts at x = 0.1
    d3 = displaced(dens)
    x_end = start[0] + d3[:, 0]
    frac = np.where(dir32[:, 0] > 0, np.clip((0.1 - start[0]) / np.maximum(x_end - start[0], 1e-12), 0, 1), 1.0)   # share of the path below the step
    length = np.linalg.norm(d3, axis=1)
    weq = length * (frac * 1.0 + (1 - frac) * 4.0)
    crossed = x_end > 0.1
    away = dir32[:, 0] < 0
    assert np.abs(length[away] - r[away]).max() < 2e-4                                          # never near the step: as in water
    # towards the step: the last stretch in water already counts with the density of the voxel it enters, so fewer cross
    # than a true water-equivalent walk would let (the 19 % that pass x = 0.1 in water), and the density-weighted length is
    # off by up to a voxel's worth
    in_water = (start[0] + d1[:, 0] > 0.1).mean()
    assert 0.15 < in_water < 0.25 and 0.02 < crossed.mean() < 0.5 * in_water
    assert np.abs(weq[~away] - r[~away]).max() < 0.1 * np.sqrt(3) * 3.0 + 1e-3
    assert np.all(weq[~away] <= r[~away] + 2e-4)                                                # the habit only ever shortens the walk


# ------------------------------------------------------------------------------------------------ P4, P5 surface reads
def test_surface_lookup_is_the_linear_texture_fetch_of_the_reference(units):
    """comsam (table) / rylsam read cos(theta) with tex3D(E idE + 0.5, U idCP + 0.5, mat + 0.5) from a linear-filter,
    clamp-addressed texture (gPET_kernals.cu:80-85, 140-145; initialize.cu:552-565): texel centres sit at i + 0.5, so this
    is the bilinear interpolation of the table at (E / dE, U / dCP), held at the borders."""
    from scipy.interpolate import RegularGridInterpolator
    rng = np.random.default_rng(4)
    nmat, ncp, ne = 3, 31, 17
    surf = np.sort(rng.uniform(-1, 1, (nmat, ncp, ne)).astype(f32), axis=1)        # an inverse CDF grows with U
    n = 50000
    xe = rng.uniform(-1.5, ne + 0.5, n).astype(f32)
    xcp = rng.uniform(-1.5, ncp + 0.5, n).astype(f32)
    xe[:100] = rng.integers(0, ne, 100); xcp[:100] = rng.integers(0, ncp, 100)      # exactly on the nodes
    for mat in range(nmat):
        out = np.zeros(n, f32)
        units.u_surface_lookup(_p(surf), C.c_int(mat), C.c_int(ncp), C.c_int(ne), C.c_int64(n), _p(xe), _p(xcp), _p(out))
        interp = RegularGridInterpolator((np.arange(ncp), np.arange(ne)), surf[mat].astype(np.float64))
        want = interp(np.stack([np.clip(xcp, 0, ncp - 1), np.clip(xe, 0, ne - 1)], 1))
        assert np.abs(out - want).max() < 3e-7 * 4
        # on a node the table value itself (to an ulp at the last row / column, where the fetch is p0 + 1 * (p1 - p0))
        assert np.abs(out[:100] - surf[mat][xcp[:100].astype(int), xe[:100].astype(int)]).max() < 1.5e-7


# ------------------------------------------------------------------------------------------------ getDistance
@pytest.mark.skipif(not parity.have_tables(), reason="packed tables not built")
def test_recording_sphere_oracle_puts_escaped_photons_on_the_sphere():
    """getDistance + the RECORDPSF == -1 branch of photon() (gPET_kernals.cu:148-171, 288-294): a photon that leaves the
    phantom is moved along its direction onto the recording sphere, its clock changed by the path."""
    s = parity.Setup(-1, phantom="cylinder", n=32)
    rng = np.random.default_rng(6)
    ph = parity.isotropic_photons(50000, rng, pos_sigma=0.1)
    sphere = (0.1, -0.2, 0.05, 5.0)
    plain = orc.phantom(ph, s.mat, s.den, s.offset, s.size, s.tab_ph, s.eabs, s.seed)
    rec = orc.phantom(ph, s.mat, s.den, s.offset, s.size, s.tab_ph, s.eabs, s.seed, record_sphere=sphere)
    alive = plain["t"] > 0
    assert np.array_equal(alive, rec["t"] > 0) and alive.mean() > 0.9
    a, b = plain[alive], rec[alive]
    pa = np.stack([a["x"], a["y"], a["z"]], 1).astype(np.float64); pb = np.stack([b["x"], b["y"], b["z"]], 1).astype(np.float64)
    v = np.stack([a["vx"], a["vy"], a["vz"]], 1).astype(np.float64)
    assert np.array_equal(a["vx"], b["vx"]) and np.array_equal(a["E"], b["E"]) and np.array_equal(a["nscat"], b["nscat"])
    assert np.abs(np.linalg.norm(pb - np.array(sphere[:3]), axis=1) - sphere[3]).max() < 5e-4      # fp32 quadratic formula
    step = ((pb - pa) * v).sum(1)
    # along the direction -- forwards, or BACKWARDS for the photons whose last free flight (mean free path ~10 cm against a
    # 1 cm phantom) carried them beyond the sphere: getDistance then returns the negative root and the clock runs back
    assert np.abs(np.cross(pb - pa, v)).max() < 5e-4 and (step > 0).mean() > 0.2 and (step < 0).mean() > 0.2
    assert np.abs((b["t"] - a["t"]) - step / 29979.2458).max() < 1e-7
    s.close()


# ------------------------------------------------------------------------------------------------ X1 panel entry, P7 majorants
@pytest.mark.skipif(not parity.have_tables(), reason="packed tables not built")
def test_detector_oracle_enters_the_panel_the_ray_hits(tmp_path):
    """photonde prologue (gPET_kernals.cu:963-1009): the panel a photon enters and the local frame its hits are written
    in, against a plain ray / rectangle intersection in numpy -- every first hit that has not been turned before lies on
    the photon's ray when mapped back with the panel's axes, in the panel the ray meets, and the photons the oracle lets
    in are the ones whose ray meets a front face."""
    s = parity.Setup(-1, phantom="air", n=8)
    rng = np.random.default_rng(12)
    nph = 60000
    ph = parity.isotropic_photons(nph, rng, pos_sigma=0.5)
    res = orc.detector(ph, s.panels, s.counts4, s.pmat, s.pdens, s.surfaces, s.tab_det, s.eabs, 2, 1, 99)
    o0 = np.stack([ph["x"], ph["y"], ph["z"]], 1).astype(np.float64); v = np.stack([ph["vx"], ph["vy"], ph["vz"]], 1).astype(np.float64)
    # independent: first front face (plane through the offset, normal = local x) the ray meets inside its rectangle, moving inwards
    best_t = np.full(nph, np.inf); best_p = np.full(nph, -1)
    for k, p in enumerate(s.panels):
        o = np.array([p["offsetx"], p["offsety"], p["offsetz"]], np.float64)
        ux, uy, uz = (np.array([p[a + "x"], p[a + "y"], p[a + "z"]], np.float64) for a in ("UniX", "UniY", "UniZ"))
        vn = v @ ux
        with np.errstate(divide="ignore", invalid="ignore"):
            t = ((o - o0) @ ux) / vn
        hit = o0 + t[:, None] * v - o
        ok = (t > 0) & (vn * float(p["directionx"]) > 0) & (np.abs(hit @ uy) < p["lengthy"] / 2) & (np.abs(hit @ uz) < p["lengthz"] / 2)
        better = ok & (t < best_t)
        best_t[better] = t[better]; best_p[better] = k
    assert abs(res["entered"] - (best_p >= 0).sum()) <= 3                      # rays through the very edge of a face may round either way
    hits = res["hits"]
    first = hits[np.r_[True, hits["parn"][1:] != hits["parn"][:-1]]]
    assert np.all(best_p[first["parn"]] >= 0)                                  # no hit without an entry
    pan = s.panels[first["pann"]]
    g = (np.stack([pan["offsetx"], pan["offsety"], pan["offsetz"]], 1).astype(np.float64)
         + first["x"][:, None].astype(np.float64) * np.stack([pan["UniXx"], pan["UniXy"], pan["UniXz"]], 1)
         + first["y"][:, None].astype(np.float64) * np.stack([pan["UniYx"], pan["UniYy"], pan["UniYz"]], 1)
         + first["z"][:, None].astype(np.float64) * np.stack([pan["UniZx"], pan["UniZy"], pan["UniZz"]], 1))
    off_ray = np.linalg.norm(np.cross(g - o0[first["parn"]], v[first["parn"]]), axis=1)
    straight = off_ray < 2e-3
    assert straight.mean() > 0.9                                               # the rest was Rayleigh-scattered before its first hit
    assert np.array_equal(first["pann"][straight], best_p[first["parn"]][straight])
    depth = ((g - o0[first["parn"]]) * v[first["parn"]]).sum(1) - best_t[first["parn"]]
    assert depth[straight].min() > -1e-3                                       # behind the front face
    # flight time of the straight ones: path / c
    path = ((g - o0[first["parn"]]) * v[first["parn"]]).sum(1)
    assert np.abs(first["t"][straight] - ph["t"][first["parn"]][straight] - path[straight] / 29979.2458).max() < 1e-6
    s.close()


@pytest.mark.skipif(not parity.have_tables(), reason="packed tables not built")
def test_majorants_bound_every_material_at_every_energy():
    """iniwck (initialize.cu:773-829, 919-966): Woodcock tracking is unbiased only if 1 / lambda_min(E) >= Sigma_tot(E) rho for
    every material at its largest density -- for the oracle's majorant and for the product's (host side of libgpet_b200)."""
    s = parity.Setup(-1, phantom="cylinder", n=32)
    dims = s.ctx.table_dims()
    lamph = s.ctx.table(0).reshape(dims["nmat"], dims["nen"]).astype(np.float64)
    maxden = np.zeros(dims["nmat"])
    for m in np.unique(s.mat):
        maxden[m] = s.den[s.mat == m].max()
    need = (lamph * maxden[:, None]).max(0)
    maj_oracle = orc.build_majorant(lamph.astype(f32), maxden.astype(f32)).astype(np.float64)
    assert np.all(maj_oracle >= need * (1 - 1e-6)) and np.all(maj_oracle <= need * (1 + 1e-5))       # tight: the maximum itself
    maj_product = s.ctx.table(6).astype(np.float64)
    assert maj_product.size == dims["nen"] and np.all(maj_product >= need * (1 - 1e-6)) and np.all(maj_product <= need * (1 + 1e-5))
    s.close()


# ------------------------------------------------------------------------------------------------ whole transport vs the REAL reference
@pytest.mark.skipif(not parity.have_tables(), reason="packed tables not built")
@pytest.mark.parametrize("case", [0, 1])
def test_oracle_transport_reproduces_the_reference_binarys_rates(case):
    """Known answers from the reference itself: tests/golden/reference_transport_rates.json holds the per-pair counters the
    reference's own CUDA build printed on the shipped example (tools/ref_pin.py on a B200, tools/make_reference_rates_fixture.py).
    The CPU oracle -- source, phantom, detector, adder / readout, thresholder, dead time, energy window -- on the same
    inputs must give the same hits, events and singles per annihilation pair within BASELINE.json's 1 % (plus 3 sigma of
    both samples)."""
    import json
    ref = json.loads((parity.GOLDEN / "reference_transport_rates.json").read_text())["cases"][case]
    window = float(ref["window_s"].split()[1])
    s = parity.Setup(device=-1, phantom=parity.gen_inputs.cylinder_phantom(n=200), size=1.0)
    src = refio.parse_sources(parity.EXAMPLE / "input" / ref["source"])
    iso = refio.parse_isotopes(parity.EXAMPLE / "data" / "isotopes.txt")
    tau = np.array([np.float64(iso[x["type"]]["halftime"]) * 1.442695 for x in src])
    frac = -np.expm1(-window / tau)
    weight = np.array([x["natom"] * frac[i] * iso[x["type"]]["ratio"] for i, x in enumerate(src)], np.float64)
    n = 300000
    per_src = np.floor(weight / weight.sum() * n).astype(np.uint64)
    per_src[-1] += np.uint64(n - int(per_src.sum()))
    ph = orc.source(np.cumsum(per_src), [x["shape"] for x in src], np.concatenate([x["coeff"] for x in src]), tau, frac, 0.0, 0,
                    0.0037056, n, 2024)
    ph = orc.phantom(ph, s.mat, s.den, s.offset, s.size, s.tab_ph, s.eabs, 2024)
    res = orc.detector(ph, s.panels, s.counts4, s.pmat, s.pdens, s.surfaces, s.tab_det, s.eabs, 2, 1, 2024)
    cfg = refio.parse_config(parity.EXAMPLE / "input_PET.in")
    p, _ = parity.make_digi_params(readout_depth=cfg["rdepth"], readout_policy=cfg["rpolicy"], threshold_eV=float(cfg["Eth"]),
                                   blur_policy=1, blur_Rref=0.0, blur_slope=0.0, blur_space=0.0, dead_level=cfg["dlevel"],
                                   dead_type=cfg["dtype"], dead_time_us=float(cfg["dtime"]), ewin_min=float(cfg["Ewinmin"]),
                                   ewin_max=float(cfg["Ewinmax"]))
    singles, counts, _ = orc.digitize(res["events"], p)
    got = {"hits": res["hits"].size / n, "events_adder": res["events"].size / n, "events_threshold": int(counts[1]) / n,
           "singles": singles.size / n}
    for k, g in got.items():
        want = ref[k + "_per_pair"]
        sig = np.sqrt(want / n + want / ref["reference_pairs"])                 # Poisson on both counts
        assert abs(g - want) < 0.01 * want + 3 * sig, (ref["source"], k, g, want)
    s.close()


# ------------------------------------------------------------------------------------------------ P4, P5, P6 in the phantom walk
@pytest.mark.skipif(not parity.have_tables(), reason="packed tables not built")
def test_phantom_oracle_single_scatters_obey_compton_kinematics():
    """photon() (gPET_kernals.cu:304-334): a photon that interacted once leaves either with its energy untouched (Rayleigh)
    or with E' = E / (1 + E / mc2 (1 - cos theta)) for the angle it was turned by (Compton with the cmpsf angle, the
    energy from the free-electron relation, :66-88), and the shares follow the partial cross sections of water."""
    n = 64
    mat = np.ones((n, n, n), np.int32); den = np.ones((n, n, n), np.float32)                      # water
    s = parity.Setup(-1, phantom=(mat, den), size=6.4)
    nph = 300000
    ph = np.zeros(nph, orc.PHOTON_DTYPE)
    ph["x"] = -3.0; ph["y"] = 0.05; ph["z"] = 0.05
    ph["vx"] = 1.0
    ph["E"] = 511e3
    ph["t"] = 1.0 + np.arange(nph)
    ph["parn"] = np.arange(nph); ph["eventid"] = np.arange(nph) // 2
    out = orc.phantom(ph, s.mat, s.den, s.offset, s.size, s.tab_ph, s.eabs, 555)
    once = out[(out["t"] > 0) & (out["nscat"] == 1)]
    assert once.size > 0.2 * nph
    norm = np.sqrt(once["vx"].astype(np.float64) ** 2 + once["vy"].astype(np.float64) ** 2 + once["vz"].astype(np.float64) ** 2)
    assert np.abs(norm - 1).max() < 1e-5
    cos_t = once["vx"].astype(np.float64) / norm                                                   # the beam flew along +x
    rayleigh = once["E"] == np.float32(511e3)
    kappa = 511e3 / MC2
    expect = 511e3 / (1.0 + kappa * (1.0 - cos_t[~rayleigh]))
    assert np.abs(once["E"][~rayleigh] / expect - 1).max() < 2e-5
    assert cos_t[rayleigh].mean() > 0.98                                                           # coherent scattering is forward peaked at 511 keV
    dims = s.ctx.table_dims()
    xs = {k: np.interp(511e3, s.energy.astype(np.float64), s.ctx.table(k).reshape(dims["nmat"], dims["nen"])[1].astype(np.float64)) for k in (1, 3)}
    share = xs[3] / (xs[1] + xs[3])                                                                # Rayleigh among the scatters
    # the once-scattered sample is biased by what happens afterwards (a Compton photon has less energy and a longer way out),
    # so only the order of magnitude is an invariant here
    assert 0.3 * share < rayleigh.mean() < 3 * share
    # energy spectrum of the Compton ones stays inside the kinematic limits
    assert once["E"][~rayleigh].min() >= 511e3 / (1 + 2 * kappa) * (1 - 1e-5)
    s.close()


# ------------------------------------------------------------------------------------------------ S6 positron PSF
def test_positron_psf_oracle_makes_back_to_back_pairs_at_the_positrons_time():
    """setPositionForPhoton (gPET_kernals.cu:563-604) without positron range: both photons start at the positron's
    position AND time (the reference forgets the second photon's time and thereby loses it -- fixed),
    photon 1 isotropic and independent of the positron's own direction, photon 2 back to back up to the acollinearity,
    ids 2i / 2i+1 counted from the batch offset."""
    n = 200000
    rng = np.random.default_rng(21)
    pos = np.zeros(n, orc.PHOTON_DTYPE)
    pos["x"], pos["y"], pos["z"] = rng.uniform(-1, 1, (3, n)).astype(f32)
    pos["vx"] = 1.0                                   # every positron flies along +x: must not show in the photons
    pos["E"] = 0.3e6
    pos["t"] = rng.uniform(1.0, 1e6, n)
    first = 7_000_000
    sigma = 0.0037056
    out = orc.psf_positron(pos, first, np.ones((4, 4, 4), f32), (-1, -1, -1), (2, 2, 2), sigma, False, 31)
    a, b = out[0::2], out[1::2]
    for q in (a, b):
        assert np.array_equal(q["x"], pos["x"]) and np.array_equal(q["y"], pos["y"]) and np.array_equal(q["z"], pos["z"])
        assert np.array_equal(q["t"], pos["t"])
        assert np.array_equal(q["eventid"], first + np.arange(n))
    assert np.array_equal(a["parn"], 2 * (first + np.arange(n))) and np.array_equal(b["parn"], a["parn"] + 1)
    assert _chi2_uniform(a["vz"] / 2 + 0.5) < 1.7 and _chi2_uniform(a["vx"] / 2 + 0.5) < 1.7
    assert _chi2_uniform(np.arctan2(a["vy"], a["vx"]) / (2 * np.pi) + 0.5) < 1.7
    va = np.stack([a["vx"], a["vy"], a["vz"]], 1).astype(np.float64); vb = np.stack([b["vx"], b["vy"], b["vz"]], 1).astype(np.float64)
    assert (va * vb).sum(1).max() < -0.999
    delta = (a["E"].astype(np.float64) - MC2) / (MC2 / 2)
    assert abs(delta.std() / sigma - 1) < 0.01 and np.abs(delta + (b["E"].astype(np.float64) - MC2) / (MC2 / 2)).max() < 4e-6
    assert abs(np.sqrt((np.linalg.norm(np.cross(va, vb), axis=1) ** 2).mean()) / sigma - 1) < 0.01
