"""Python readers/writers for gPET's file formats.

Plays the role of output/readOutput.m (the reference's only definition of the output layouts, readOutput.m:1-54) and
gives the tests an independent (numpy) restatement of the input parsers: input_PET.in (main.cu:50-184), .geo
(detector.cu:64-285), source.txt / isotopes.txt (initialize.cu:10-31, 116-144), psf.dat (initialize.cu:76-115) and
the input4gPET.* tables (initialize.cu:279-748).  Nothing here is on the product's compute path.
"""
from __future__ import annotations

import re
from pathlib import Path

import numpy as np

from .api import EVENT_DTYPE, COINC_DTYPE, PANEL_DTYPE, PANEL_FIELDS

# ------------------------------------------------------------------------------------------------ outputs
def read_events(path):
    """singles.dat / adder.dat: raw 48-byte Event records (readOutput.m:18-34)."""
    return np.fromfile(path, EVENT_DTYPE)


def write_events(path, ev, append=False):
    with open(path, "ab" if append else "wb") as f:
        np.ascontiguousarray(ev, EVENT_DTYPE).tofile(f)


def read_coincidences(path):
    return np.fromfile(path, COINC_DTYPE)


def read_coincidence_classes(path):
    """coincidences_class.dat: one uint8 per record of coincidences.dat (0 true, 1 scatter, 2 random)."""
    return np.fromfile(path, np.uint8)


def read_hits(hits_id_path, hits_path):
    """HitsID.dat (5 x int32 per hit) + Hits.dat (5 x float32 per hit) (readOutput.m:3-16).
    Columns: particle id, panel, module, crystal, type | E, t, local x, y, z."""
    ids = np.fromfile(hits_id_path, "<i4").reshape(-1, 5)
    f = np.fromfile(hits_path, "<f4").reshape(-1, 5)
    return ids, f


def read_psf_triplet(out_path, id_path, time_path):
    """out*.dat 7 x float32, id*.dat int32, time*.dat float64 (readOutput.m:36-54)."""
    return (np.fromfile(out_path, "<f4").reshape(-1, 7), np.fromfile(id_path, "<i4"), np.fromfile(time_path, "<f8"))


def write_psf(path, x, y, z, t, vx, vy, vz, E):
    """psf.dat input: 8 x float64 per record, x y z t vx vy vz E (initialize.cu:79-104)."""
    rec = np.stack([x, y, z, t, vx, vy, vz, E], axis=1).astype("<f8")
    rec.tofile(path)


# ------------------------------------------------------------------------------------------------ scanf-like cursor
class _Scanner:
    def __init__(self, text):
        self.s = text
        self.p = 0

    def ws(self):
        while self.p < len(self.s) and self.s[self.p].isspace():
            self.p += 1

    def line(self, maxlen=0):
        start = self.p
        while self.p < len(self.s):
            if maxlen and self.p - start >= maxlen - 1:
                break
            c = self.s[self.p]
            self.p += 1
            if c == "\n":
                break
        return self.s[start:self.p]

    _num = re.compile(r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)")
    _int = re.compile(r"[-+]?\d+")

    def f32(self):
        self.ws()
        m = self._num.match(self.s, self.p)
        if not m:
            raise ValueError(f"number expected at offset {self.p}")
        self.p = m.end()
        self.ws()
        return np.float32(m.group(0))

    def i32(self):
        self.ws()
        m = self._int.match(self.s, self.p)
        if not m:
            raise ValueError(f"integer expected at offset {self.p}")
        self.p = m.end()
        self.ws()
        return int(m.group(0))

    def word(self):
        self.ws()
        start = self.p
        while self.p < len(self.s) and not self.s[self.p].isspace():
            self.p += 1
        w = self.s[start:self.p]
        self.ws()
        return w

    def ignore_through(self, n, delim):
        cnt = 0
        while self.p < len(self.s) and cnt < n:
            c = self.s[self.p]
            self.p += 1
            cnt += 1
            if c == delim:
                break

    def next_is_number(self):
        self.ws()
        if self.p >= len(self.s):
            return False
        return bool(re.match(r"[-+.]?\d", self.s[self.p:self.p + 2])) or bool(re.match(r"[-+]\.\d", self.s[self.p:self.p + 3]))

    def skip_labels(self):
        while self.p < len(self.s) and not self.next_is_number():
            self.line()


def parse_config(path):
    s = _Scanner(Path(path).read_text())
    L = 200
    c = {}
    s.line(L); c["device"] = s.i32()
    s.line(L); c["nonangle"] = s.f32()
    s.line(L); c["pdim"] = [s.i32() for _ in range(3)]
    s.line(L); c["poffset"] = [s.f32() for _ in range(3)]
    s.line(L); c["psize"] = [s.f32() for _ in range(3)]
    s.line(L); c["matfile"] = s.word()
    s.line(L); c["denfile"] = s.word()
    s.line(L); c["nhist"] = s.i32()
    s.line(L); c["usepsf"] = s.i32()
    s.line(L); c["sourcefile"] = s.word()
    s.line(L); c["ptype"] = s.i32()
    s.line(L); c["useprange"] = s.i32()
    s.line(L); c["tstart"] = s.f32(); c["tend"] = s.f32()
    s.line(L); c["recordsphere"] = [s.f32() for _ in range(4)]
    s.line(L); c["eabsph"] = s.f32()
    s.line(L); c["geofile"] = s.word()
    s.line(L); c["nsurface"] = s.i32()
    c["surface"] = [s.f32() for _ in range(10 * c["nsurface"])]
    s.line(L); c["rdepth"] = s.i32(); c["rpolicy"] = s.i32()
    s.line(L); c["Eth"] = s.f32()
    s.line(L); c["blurpolicy"] = s.i32(); c["Eref"] = s.f32(); c["Rref"] = s.f32(); c["Eslope"] = s.f32(); c["Sblur"] = s.f32()
    s.line(L); c["dlevel"] = s.i32(); c["dtype"] = s.i32(); c["dtime"] = s.f32()
    s.line(L); c["Ewinmin"] = s.f32(); c["Ewinmax"] = s.f32()
    return c


def _rot(rot, ang, v):
    f = np.float32
    ca, sa = f(np.cos(ang, dtype=np.float32)), f(np.sin(ang, dtype=np.float32))
    one = f(1)
    return np.array([
        (one - ca) * (v[0] * rot[0]) * rot[0] + ca * v[0] + sa * (rot[1] * v[2] - rot[2] * v[1]),
        (one - ca) * (v[1] * rot[1]) * rot[1] + ca * v[1] + sa * (rot[2] * v[0] - rot[0] * v[2]),
        (one - ca) * (v[2] * rot[2]) * rot[2] + ca * v[2] + sa * (rot[0] * v[1] - rot[1] * v[0])], np.float32)


def parse_geometry(path):
    """Returns (panels[PANEL_DTYPE], mat[2], dens[2], counts(moduleNy, crystalNy, moduleN, crystalN))."""
    s = _Scanner(Path(path).read_text())
    L = 256
    s.line(L); count = s.i32()
    s.line(L); rot = np.array([s.f32() for _ in range(3)], np.float32)
    s.line(L); ang_deg = s.f32()
    s.line(L)
    mat = np.zeros(2, np.int32); dens = np.zeros(2, np.float32)
    for i in range(2):
        mat[i] = s.i32(); dens[i] = s.f32()
    s.line(L)
    p = np.zeros(count, PANEL_DTYPE)
    s.line(L); p["panel"][0] = s.i32()
    for names in (("lengthx", "lengthy", "lengthz"), ("MODx", "MODy", "MODz"), ("Mspacex", "Mspacey", "Mspacez"),
                  ("LSOx", "LSOy", "LSOz"), ("spacex", "spacey", "spacez"), ("directionx", "directiony", "directionz"),
                  ("offsetx", "offsety", "offsetz"), ("UniXx", "UniXy", "UniXz"), ("UniYx", "UniYy", "UniYz"),
                  ("UniZx", "UniZy", "UniZz")):
        s.line(L)
        for nme in names:
            p[nme][0] = s.f32()
    PI = np.float32(3.1415926535897932384626433)
    for i in range(1, count):
        p[i] = p[0]
        p["panel"][i] = i
        ang = np.float32(np.float32(np.float32(ang_deg * PI) / np.float32(180.0)) * np.float32(i))
        for pre in ("offset", "UniX", "UniY", "UniZ"):
            v = np.array([p[pre + a][0] for a in "xyz"], np.float32)
            o = _rot(rot, ang, v)
            for k, a in enumerate("xyz"):
                p[pre + a][i] = o[k]
    q = p[0]
    f = np.float32
    Mn = int(np.floor(f(q["lengthy"]) / (f(q["MODy"]) + f(q["Mspacey"]))) + 1)
    Ln = int(np.floor(f(q["MODy"]) / (f(q["LSOy"]) + f(q["spacey"]))) + 1)
    moduleNy, crystalNy = Mn, Ln
    Mn = int(Mn * (np.floor(f(q["lengthz"]) / (f(q["MODz"]) + f(q["Mspacez"]))) + 1))
    Ln = int(Ln * (np.floor(f(q["MODz"]) / (f(q["LSOz"]) + f(q["spacez"]))) + 1))
    return p, mat, dens, np.array([moduleNy, crystalNy, Mn, Ln], np.int32)


def parse_sources(path):
    s = _Scanner(Path(path).read_text())
    n = s.i32()
    s.ignore_through(512, "#")
    out = []
    for _ in range(n):
        natom = s.i32(); ty = s.i32(); sh = s.i32()
        out.append(dict(natom=natom, type=ty, shape=sh, coeff=np.array([s.f32() for _ in range(6)], np.float32)))
    return out


def parse_isotopes(path):
    s = _Scanner(Path(path).read_text())
    n = s.i32()
    s.ignore_through(512, "#")
    out = []
    for _ in range(n):
        hl = s.f32(); ra = s.f32()
        out.append(dict(halftime=hl, ratio=ra, coef=np.array([s.f32() for _ in range(8)], np.float32)))
    return out


def read_psf(path, max_particles=0):
    rec = np.fromfile(path, "<f8")
    rec = rec[: (rec.size // 8) * 8].reshape(-1, 8)
    if max_particles:
        rec = rec[:max_particles]
    return rec


# ------------------------------------------------------------------------------------------------ tables
def _numbers(text):
    return np.array(text.split(), dtype=np.float64)


def read_table_1d(path, nmat):
    """.lamph/.compt/.phote/.rayle -> (energy[nen] float32, values[nmat, nen] float32)."""
    s = _Scanner(Path(path).read_text())
    energy = None
    vals = []
    for _ in range(nmat):
        s.skip_labels()
        nd = s.i32()
        for _k in range(4):
            s.f32()
        s.skip_labels()
        a = np.array([s.f32() for _ in range(2 * nd)], np.float32).reshape(nd, 2)
        if energy is None:
            energy = a[:, 0].copy()
        vals.append(a[:, 1].copy())
    return energy, np.stack(vals)


def read_matter(path):
    s = _Scanner(Path(path).read_text())
    s.skip_labels(); eminph = s.f32(); s.f32(); emax = s.f32()
    s.skip_labels(); s.f32(); s.f32()
    s.skip_labels(); s.f32(); s.f32(); s.f32()
    s.skip_labels(); nmat = s.i32()
    names, dens = [], []
    for _ in range(nmat):
        while True:
            ln = s.line()
            if "MATERIAL:" in ln:
                names.append(ln.split("MATERIAL:")[1].strip())
                break
        s.skip_labels(); dens.append(s.f32())
        s.skip_labels(); nel = s.i32()
        for _k in range(nel):
            s.i32(); s.f32()
        s.skip_labels(); s.f32(); s.f32(); s.f32()
        s.skip_labels(); s.f32()
        s.skip_labels(); s.f32(); s.f32()
    return dict(eminph=eminph, emax=emax, nmat=nmat, names=names, refdens=np.array(dens, np.float32))


def read_surface(path, nmat):
    """.cmpsf/.rayff -> dict(q, f (the S(q)/F(q) blocks), ncp, ne, dcp, de, surf[nmat, ncp, ne])."""
    text = Path(path).read_text()
    s = _Scanner(text)
    blocks, surfs = [], []
    meta = None
    for _ in range(nmat):
        s.skip_labels()
        nd = s.i32(); s.f32(); s.f32(); s.f32()
        s.skip_labels()
        qb = np.array([s.f32() for _ in range(3 * nd)], np.float32).reshape(nd, 3)
        blocks.append(qb)
        s.skip_labels()
        ncp = s.i32(); s.f32(); s.f32(); dcp = s.f32(); ne = s.i32(); s.f32(); s.f32(); de = s.f32()
        meta = (ncp, ne, dcp, de)
        for _k in range(ncp + ne):
            s.f32()
        # fast path for the big block
        m = re.compile(r"\s*((?:[-+]?\d\S*\s+){%d})" % (ncp * ne)).match(text, s.p)
        if m:
            arr = np.array(m.group(1).split(), np.float32)
            s.p = m.end()
        else:
            arr = np.array([s.f32() for _ in range(ncp * ne)], np.float32)
        surfs.append(arr.reshape(ncp, ne))
    return dict(sq=blocks, ncp=meta[0], ne=meta[1], dcp=np.float32(meta[2]), de=np.float32(meta[3]), surf=np.stack(surfs))


def read_tables(prefix):
    """Whole table set as numpy arrays (independent restatement of rmater/rlamph/rcompt/rcmpsf/rphote/rrayle/rrayff)."""
    prefix = str(prefix)
    m = read_matter(prefix + ".matter")
    e, lamph = read_table_1d(prefix + ".lamph", m["nmat"])
    _, compt = read_table_1d(prefix + ".compt", m["nmat"])
    _, phote = read_table_1d(prefix + ".phote", m["nmat"])
    _, rayle = read_table_1d(prefix + ".rayle", m["nmat"])
    cm = read_surface(prefix + ".cmpsf", m["nmat"])
    rl = read_surface(prefix + ".rayff", m["nmat"])
    return dict(matter=m, energy=e, lamph=lamph, compt=compt, phote=phote, rayle=rayle, cmpsf=cm, rayff=rl)
