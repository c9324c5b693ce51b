"""ctypes binding of the CPU oracle (oracle/gpet_oracle.c).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs -- never by the product."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libgpet_oracle.so"

EVENT_DTYPE = np.dtype(
    [("parn", "<i4"), ("pann", "<i4"), ("modn", "<i4"), ("cryn", "<i4"), ("siten", "<i4"), ("eventid", "<i4"),
     ("t", "<f8"), ("E", "<f4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4")], align=True)
COINC_DTYPE = np.dtype([("a", EVENT_DTYPE), ("b", EVENT_DTYPE)], align=True)
HIT_DTYPE = np.dtype(
    [("parn", "<i4"), ("pann", "<i4"), ("modn", "<i4"), ("cryn", "<i4"), ("type", "<i4"),
     ("E", "<f4"), ("t32", "<f4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("t", "<f8")], align=True)
PHOTON_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("E", "<f4"), ("vx", "<f4"), ("vy", "<f4"), ("vz", "<f4"),
     ("nscat", "<i4"), ("t", "<f8"), ("eventid", "<i4"), ("parn", "<i4")], align=True)


class Tables(C.Structure):
    _fields_ = [("nmat", C.c_int32), ("nen", C.c_int32), ("e0", C.c_float), ("e1", C.c_float),
                ("lamph", C.c_void_p), ("compt", C.c_void_p), ("rayle", C.c_void_p), ("maj", C.c_void_p),
                ("cm_ncp", C.c_int32), ("cm_ne", C.c_int32), ("rl_ncp", C.c_int32), ("rl_ne", C.c_int32),
                ("cm_dcp", C.c_float), ("cm_de", C.c_float), ("rl_dcp", C.c_float), ("rl_de", C.c_float),
                ("cmpsf", C.c_void_p), ("rayff", C.c_void_p)]


class DigiParams(C.Structure):
    _fields_ = [("readout_depth", C.c_int32), ("readout_policy", C.c_int32), ("threshold_eV", C.c_float),
                ("blur_policy", C.c_int32), ("blur_Eref", C.c_float), ("blur_Rref", C.c_float),
                ("blur_slope", C.c_float), ("blur_space", C.c_float),
                ("dead_level", C.c_int32), ("dead_type", C.c_int32), ("dead_time_us", C.c_float),
                ("ewin_min", C.c_float), ("ewin_max", C.c_float),
                ("time_blur_sigma_us", C.c_float), ("coinc_window_us", C.c_float),
                ("coinc_policy", C.c_int32), ("coinc_min_panel_diff", C.c_int32),
                ("npanels", C.c_int32), ("moduleN", C.c_int32), ("crystalN", C.c_int32), ("seed", C.c_uint64),
                ("tie_site", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB.exists():
            subprocess.run(["make", "-C", str(HERE)], check=True)
        _lib = C.CDLL(str(LIB))
        _lib.orc_digitize.restype = C.c_int64
        _lib.orc_detector.restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def set_id_base(base):
    """global index of the first photon of the frame the following calls work on (64-bit history numbers); 0 = ids as they are"""
    lib().orc_set_id_base(C.c_uint64(int(base)))


def philox(ctr, key):
    c = np.asarray(ctr, np.uint32); k = np.asarray(key, np.uint32); o = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def build_majorant(lamph, maxdens):
    lamph = np.ascontiguousarray(lamph, np.float32)
    nmat, nen = lamph.shape
    md = np.ascontiguousarray(maxdens, np.float32)
    out = np.zeros(nen, np.float32)
    lib().orc_build_majorant(C.c_int(nmat), C.c_int(nen), _p(lamph), _p(md), _p(out))
    return out


class TableSet:
    """Holds numpy arrays alive and exposes the C struct.  `t` is the dict returned by gpet_b200.refio.read_tables or an
    equivalent built from the product's getters."""

    def __init__(self, energy, lamph, compt, rayle, cmpsf, cm_dcp, cm_de, rayff, rl_dcp, rl_de, maj):
        self.a = [np.ascontiguousarray(x, np.float32) for x in (lamph, compt, rayle, maj, cmpsf, rayff)]
        lamph, compt, rayle, maj, cmpsf, rayff = self.a
        s = Tables()
        s.nmat, s.nen = lamph.shape
        s.e0, s.e1 = float(energy[0]), float(energy[-1])
        s.lamph, s.compt, s.rayle, s.maj = _p(lamph), _p(compt), _p(rayle), _p(maj)
        s.cm_ncp, s.cm_ne = cmpsf.shape[1], cmpsf.shape[2]
        s.rl_ncp, s.rl_ne = rayff.shape[1], rayff.shape[2]
        s.cm_dcp, s.cm_de, s.rl_dcp, s.rl_de = float(cm_dcp), float(cm_de), float(rl_dcp), float(rl_de)
        s.cmpsf, s.rayff = _p(cmpsf), _p(rayff)
        self.c = s

    def with_majorant(self, maj):
        lamph, compt, rayle, _, cmpsf, rayff = self.a
        e = (self.c.e0, self.c.e1)
        return TableSet(e, lamph, compt, rayle, cmpsf, self.c.cm_dcp, self.c.cm_de, rayff, self.c.rl_dcp, self.c.rl_de, maj)


def source(cum_pairs, shape, coeff, tau_s, frac, t0_s, first_pair, nonangle, npairs, seed):
    cum = np.ascontiguousarray(cum_pairs, np.uint64); sh = np.ascontiguousarray(shape, np.int32)
    co = np.ascontiguousarray(coeff, np.float32); tau = np.ascontiguousarray(tau_s, np.float64)
    fr = np.ascontiguousarray(frac, np.float64)
    out = np.zeros(2 * npairs, PHOTON_DTYPE)
    lib().orc_source(C.c_int(len(cum)), _p(cum), _p(sh), _p(co), _p(tau), _p(fr), C.c_double(t0_s), C.c_uint64(first_pair),
                     C.c_float(nonangle), C.c_uint64(npairs), C.c_uint64(seed), _p(out))
    return out


def source_with_range(cum_pairs, shape, coeff, tau_s, frac, t0_s, first_pair, nonangle, npairs, seed, types, iso_coef,
                      dens, offset, size):
    """orc_source with S4 + S5 (positron kinetic energy and range, gPET_kernals.cu:347-443) switched on."""
    cum = np.ascontiguousarray(cum_pairs, np.uint64); sh = np.ascontiguousarray(shape, np.int32)
    co = np.ascontiguousarray(coeff, np.float32); tau = np.ascontiguousarray(tau_s, np.float64)
    fr = np.ascontiguousarray(frac, np.float64); ty = np.ascontiguousarray(types, np.int32)
    ic = np.ascontiguousarray(iso_coef, np.float32).ravel(); dens = np.ascontiguousarray(dens, np.float32)
    nz, ny, nx = dens.shape
    dim = np.array([nx, ny, nz], np.int32); off = np.asarray(offset, np.float32); sz = np.asarray(size, np.float32)
    out = np.zeros(2 * npairs, PHOTON_DTYPE)
    lib().orc_source_ex(C.c_int(len(cum)), _p(cum), _p(sh), _p(co), _p(tau), _p(fr), C.c_double(t0_s), C.c_uint64(first_pair),
                        C.c_float(nonangle), C.c_uint64(npairs), C.c_uint64(seed), C.c_int(1), _p(ty), _p(ic), _p(dens),
                        _p(dim), _p(off), _p(sz), _p(out))
    return out


def psf_positron(positrons, first, dens, offset, size, nonangle, use_prange, seed):
    """setPositionForPhoton (gPET_kernals.cu:563-604): positron records -> photon pairs."""
    pos = np.ascontiguousarray(positrons, PHOTON_DTYPE); dens = np.ascontiguousarray(dens, np.float32)
    nz, ny, nx = dens.shape
    dim = np.array([nx, ny, nz], np.int32); off = np.asarray(offset, np.float32); sz = np.asarray(size, np.float32)
    out = np.zeros(2 * pos.size, PHOTON_DTYPE)
    lib().orc_psf_positron(_p(pos), C.c_int64(pos.size), C.c_uint64(first), _p(dens), _p(dim), _p(off), _p(sz),
                           C.c_float(nonangle), C.c_int(1 if use_prange else 0), C.c_uint64(seed), _p(out))
    return out


def phantom(photons, mat, dens, offset, size, tables: TableSet, eabs, seed, record_sphere=None):
    """photon() (gPET_kernals.cu:256-345); record_sphere = (x, y, z, r) switches on the RECORDPSF == -1 branch."""
    ph = np.ascontiguousarray(photons, PHOTON_DTYPE).copy()
    mat = np.ascontiguousarray(mat, np.int32); dens = np.ascontiguousarray(dens, np.float32)
    nz, ny, nx = mat.shape
    dim = np.array([nx, ny, nz], np.int32); off = np.asarray(offset, np.float32); sz = np.asarray(size, np.float32)
    rec = None if record_sphere is None else np.ascontiguousarray(record_sphere, np.float32)
    lib().orc_phantom_ex(_p(ph), C.c_int64(ph.size), _p(mat), _p(dens), _p(dim), _p(off), _p(sz), C.byref(tables.c),
                         C.c_float(eabs), C.c_uint64(seed), _p(rec) if rec is not None else None)
    return ph


def noise(t_lo_us, t_hi_us, mean_gap_us, Emean, sigma, interval_us, npanels, moduleN, crystalN, seed, cap=1 << 22):
    """addnoise (gPET_kernals.cu:699-735) restated: the noise events with t_lo <= t < t_hi, in slice order."""
    out = np.zeros(cap, EVENT_DTYPE)
    lib().orc_noise.restype = C.c_int64
    n = lib().orc_noise(C.c_double(t_lo_us), C.c_double(t_hi_us), C.c_float(mean_gap_us), C.c_float(Emean), C.c_float(sigma),
                        C.c_float(interval_us), C.c_int32(npanels), C.c_int32(moduleN), C.c_int32(crystalN), C.c_uint64(seed),
                        _p(out), C.c_int64(cap))
    assert n <= cap, "noise event buffer too small"
    return out[:n]


def detector(photons, panels, counts4, pmat, pdens, surfaces, tables: TableSet, eabs, rdepth, rpolicy, seed,
             hit_cap=None, ev_cap=None):
    ph = np.ascontiguousarray(photons, PHOTON_DTYPE)
    n = ph.size
    hit_cap = hit_cap or max(16 * n, 1024); ev_cap = ev_cap or max(8 * n, 1024)
    hits = np.zeros(hit_cap, HIT_DTYPE); ev = np.zeros(ev_cap, EVENT_DTYPE)
    nh = C.c_int64(); ne = C.c_int64(); ovf = C.c_int64()
    panels = np.ascontiguousarray(panels)
    c4 = np.ascontiguousarray(counts4, np.int32); pm = np.ascontiguousarray(pmat, np.int32)
    pd = np.ascontiguousarray(pdens, np.float32); sf = np.ascontiguousarray(surfaces, np.float32).ravel()
    nsurf = sf.size // 10
    if sf.size == 0:
        sf = np.zeros(10, np.float32)
    entered = lib().orc_detector(_p(ph), C.c_int64(n), _p(panels), C.c_int(panels.size), _p(c4), _p(pm), _p(pd),
                                 C.c_int(nsurf), _p(sf), C.byref(tables.c), C.c_float(eabs), C.c_int(rdepth),
                                 C.c_int(rpolicy), C.c_uint64(seed), _p(hits), C.c_int64(hit_cap), C.byref(nh),
                                 _p(ev), C.c_int64(ev_cap), C.byref(ne), C.byref(ovf))
    assert nh.value <= hit_cap and ne.value <= ev_cap
    return dict(entered=entered, hits=hits[:nh.value], events=ev[:ne.value], adder_overflow=ovf.value)


def digitize(events, params: DigiParams, want_coinc=True):
    ev = np.ascontiguousarray(events, EVENT_DTYPE)
    n = ev.size
    work = np.zeros(max(n, 1), EVENT_DTYPE); out = np.zeros(max(n, 1), EVENT_DTYPE)
    counts = np.zeros(4, np.uint64)
    ccap = max(4 * n, 16)
    co = np.zeros(ccap, COINC_DTYPE); nc = C.c_int64()
    ns = lib().orc_digitize(_p(ev), C.c_int64(n), C.byref(params), _p(work), _p(out), _p(counts), _p(co), C.c_int64(ccap),
                            C.byref(nc))
    assert nc.value <= ccap
    return out[:ns], counts, co[:nc.value]


def classify(coincidences, scattered_parn, pair_shift=0):
    """(classes uint8[n], totals uint64[3]): 0 true, 1 scatter, 2 random (orc_classify)."""
    co = np.ascontiguousarray(coincidences, COINC_DTYPE)
    sc = np.ascontiguousarray(scattered_parn, np.int32)
    cls = np.zeros(max(co.size, 1), np.uint8)
    totals = np.zeros(3, np.uint64)
    lib().orc_classify(_p(co), C.c_int64(co.size), _p(sc), C.c_int64(sc.size), C.c_int32(pair_shift), _p(cls), _p(totals))
    return cls[:co.size], totals
