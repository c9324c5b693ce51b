"""ctypes mirror of include/gpet_b200.h -- the host-side face of libgpet_b200.so.

The reference has no Python layer (it is one CUDA program); this module exists so that tests, bench.py and
scripts can drive the C ABI.  Every method maps 1:1 onto an `extern "C"` entry point; nothing is computed here and
there is no CPU fallback: if the CUDA library is missing, importing `lib()` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libgpet_b200.so"

# numpy dtypes of the records (byte-identical to the C structs)
EVENT_DTYPE = np.dtype(
    [("parn", "<i4"), ("pann", "<i4"), ("modn", "<i4"), ("cryn", "<i4"), ("siten", "<i4"), ("eventid", "<i4"),
     ("t", "<f8"), ("E", "<f4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4")], align=True)
assert EVENT_DTYPE.itemsize == 48
COINC_DTYPE = np.dtype([("a", EVENT_DTYPE), ("b", EVENT_DTYPE)], align=True)
# gpet_single_compact (include/gpet_b200.h): ids = pann | modn << 8 | cryn << 20 | (parn & 1) << 31
COMPACT_DTYPE = np.dtype([("t", "<f8"), ("E", "<f4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("eventid", "<i4"), ("ids", "<u4")], align=True)
assert COMPACT_DTYPE.itemsize == 32
assert COINC_DTYPE.itemsize == 96
HIT_DTYPE = np.dtype(
    [("parn", "<i4"), ("pann", "<i4"), ("modn", "<i4"), ("cryn", "<i4"), ("type", "<i4"),
     ("E", "<f4"), ("t32", "<f4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("t", "<f8")], align=True)
assert HIT_DTYPE.itemsize == 48
PHOTON_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("E", "<f4"), ("vx", "<f4"), ("vy", "<f4"), ("vz", "<f4"),
     ("nscat", "<i4"), ("t", "<f8"), ("eventid", "<i4"), ("parn", "<i4")], align=True)
assert PHOTON_DTYPE.itemsize == 48
PANEL_FIELDS = ["panel", "lengthx", "lengthy", "lengthz", "MODx", "MODy", "MODz", "Mspacex", "Mspacey", "Mspacez",
                "LSOx", "LSOy", "LSOz", "spacex", "spacey", "spacez", "offsetx", "offsety", "offsetz",
                "directionx", "directiony", "directionz", "UniXx", "UniXy", "UniXz", "UniYx", "UniYy", "UniYz",
                "UniZx", "UniZy", "UniZz"]
PANEL_DTYPE = np.dtype([("panel", "<i4")] + [(f, "<f4") for f in PANEL_FIELDS[1:]], align=True)
assert PANEL_DTYPE.itemsize == 124

GPET_MAX_SURFACES = 5


class DigitizerParams(C.Structure):
    _fields_ = [("readout_depth", C.c_int32), ("readout_policy", C.c_int32), ("threshold_eV", C.c_float),
                ("blur_policy", C.c_int32), ("blur_Eref", C.c_float), ("blur_Rref", C.c_float),
                ("blur_slope", C.c_float), ("blur_space", C.c_float),
                ("dead_level", C.c_int32), ("dead_type", C.c_int32), ("dead_time_us", C.c_float),
                ("ewin_min", C.c_float), ("ewin_max", C.c_float),
                ("time_blur_sigma_us", C.c_float), ("coinc_window_us", C.c_float),
                ("coinc_policy", C.c_int32), ("coinc_min_panel_diff", C.c_int32),
                ("noise_mean_gap_us", C.c_float), ("noise_Emean_eV", C.c_float), ("noise_sigma_eV", C.c_float),
                ("noise_interval_us", C.c_float), ("coinc_pair_shift", C.c_int32)]


class TransportParams(C.Structure):
    _fields_ = [("noncollinearity_rad", C.c_float), ("use_positron_range", C.c_int32), ("eabs_eV", C.c_float),
                ("nsurface", C.c_int32), ("surface", C.c_float * (10 * GPET_MAX_SURFACES)),
                ("record_hits", C.c_int32), ("record_psf", C.c_int32), ("record_sphere", C.c_float * 4)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("pairs", "photons_phantom_out", "photons_on_panel", "hits", "events_adder", "events_threshold",
                 "events_deadtime", "singles", "coincidences", "overflow_hits", "overflow_events", "overflow_adder",
                 "frames", "kernel_launches")] + \
               [(n, C.c_double) for n in ("ms_source", "ms_phantom", "ms_detector", "ms_digitizer", "ms_total")] + \
               [(n, C.c_uint64) for n in ("trues", "scatters", "randoms")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# name -> (restype, argtypes); must list every symbol include/gpet_b200.h declares (checked by tests)
_P = C.c_void_p
_SIGS = {
    "gpet_abi_version": (C.c_int, []),
    "gpet_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "gpet_destroy": (None, [_P]),
    "gpet_last_error": (C.c_char_p, [_P]),
    "gpet_set_stream": (C.c_int, [_P, _P]),
    "gpet_set_seed": (C.c_int, [_P, C.c_uint64]),
    "gpet_set_capacity": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_uint64]),
    "gpet_load_config_file": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_char_p]),
    "gpet_load_tables": (C.c_int, [_P, C.c_char_p]),
    "gpet_save_tables_packed": (C.c_int, [_P, C.c_char_p]),
    "gpet_load_phantom_files": (C.c_int, [_P, C.c_char_p, C.c_char_p, _P, _P, _P]),
    "gpet_set_phantom": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "gpet_load_geometry": (C.c_int, [_P, C.c_char_p]),
    "gpet_load_isotopes": (C.c_int, [_P, C.c_char_p]),
    "gpet_load_source": (C.c_int, [_P, C.c_char_p]),
    "gpet_load_psf": (C.c_int, [_P, C.c_char_p, C.c_int64, C.c_int]),
    "gpet_set_digitizer": (C.c_int, [_P, C.POINTER(DigitizerParams)]),
    "gpet_get_digitizer": (C.c_int, [_P, C.POINTER(DigitizerParams)]),
    "gpet_set_transport": (C.c_int, [_P, C.POINTER(TransportParams)]),
    "gpet_get_transport": (C.c_int, [_P, C.POINTER(TransportParams)]),
    "gpet_set_time_window": (C.c_int, [_P, C.c_float, C.c_float]),
    "gpet_set_source_atoms": (C.c_int, [_P, C.c_int, C.c_uint64]),
    "gpet_get_num_panels": (C.c_int, [_P]),
    "gpet_get_panels": (C.c_int, [_P, _P, C.c_int]),
    "gpet_get_geometry_counts": (C.c_int, [_P, _P, _P, _P]),
    "gpet_get_table_dims": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gpet_get_table": (C.c_int64, [_P, C.c_int, _P, C.c_int64]),
    "gpet_get_num_sources": (C.c_int, [_P]),
    "gpet_get_source": (C.c_int, [_P, C.c_int, _P, _P, _P, _P]),
    "gpet_get_num_isotopes": (C.c_int, [_P]),
    "gpet_get_isotope": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "gpet_get_num_psf": (C.c_int64, [_P]),
    "gpet_plan_frames": (C.c_int64, [_P, C.c_uint64]),
    "gpet_frame_pairs": (C.c_int64, [_P, C.c_int64]),
    "gpet_get_frame": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P]),
    "gpet_stage_source": (C.c_int, [_P, C.c_int64]),
    "gpet_stage_psf": (C.c_int, [_P, C.c_int64, C.c_int64]),
    "gpet_set_psf_output": (C.c_int, [_P, C.c_int]),
    "gpet_stage_phantom": (C.c_int, [_P]),
    "gpet_stage_detector": (C.c_int, [_P]),
    "gpet_stage_front": (C.c_int, [_P, C.c_int64]),
    "gpet_stage_panel_transport": (C.c_int, [_P]),
    "gpet_stage_digitize": (C.c_int, [_P]),
    "gpet_stage_noise": (C.c_int, [_P, C.c_double, C.c_double]),
    "gpet_queue_size": (C.c_int64, [_P, C.c_int]),
    "gpet_put_photons": (C.c_int, [_P, C.c_int, _P, C.c_int64]),
    "gpet_fetch_photons": (C.c_int64, [_P, C.c_int, _P, C.c_int64]),
    "gpet_put_events": (C.c_int, [_P, _P, C.c_int64]),
    "gpet_fetch_events": (C.c_int64, [_P, _P, C.c_int64]),
    "gpet_fetch_hits": (C.c_int64, [_P, _P, C.c_int64]),
    "gpet_fetch_singles": (C.c_int64, [_P, _P, C.c_int64]),
    "gpet_fetch_coincidences": (C.c_int64, [_P, _P, C.c_int64]),
    "gpet_fetch_coincidence_classes": (C.c_int64, [_P, _P, C.c_int64, _P]),
    "gpet_mark_scattered": (C.c_int, [_P, _P, C.c_int64]),
    "gpet_last_counts": (C.c_int, [_P, _P]),
    "gpet_digitize": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int64, C.POINTER(C.c_int64), _P]),
    "gpet_run": (C.c_int, [_P, C.c_char_p, C.POINTER(Stats)]),
    "gpet_run_resident": (C.c_int, [_P, C.POINTER(Stats)]),
    "gpet_result_singles": (C.c_int64, [_P, C.POINTER(_P)]),
    "gpet_result_coincidences": (C.c_int64, [_P, C.POINTER(_P)]),
    "gpet_set_coincidence_format": (C.c_int, [_P, C.c_int]),
    "gpet_set_singles_format": (C.c_int, [_P, C.c_int]),
    "gpet_result_singles_compact": (C.c_int64, [_P, C.POINTER(_P)]),
    "gpet_expand_singles": (C.c_int, [_P, _P, C.c_int64, _P]),
    "gpet_result_coincidence_pairs": (C.c_int64, [_P, C.POINTER(_P)]),
    "gpet_result_coincidence_classes": (C.c_int64, [_P, C.POINTER(_P)]),
    "gpet_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "gpet_get_spectrum": (C.c_int, [_P, _P, C.c_int]),
    "gpet_set_spectrum": (C.c_int, [_P, C.c_int, C.c_float, C.c_float]),
    "gpet_set_shard": (C.c_int, [_P, C.c_int, C.c_int]),
    "gpet_set_first_pair": (C.c_int, [_P, C.c_uint64]),
    "gpet_set_emit_window": (C.c_int, [_P, C.c_double, C.c_double, C.c_double]),
    "gpet_clear_emit_window": (C.c_int, [_P]),
    "gpet_get_emit_counts": (C.c_int, [_P, _P]),
    "gpet_copy_events_to_device": (C.c_int64, [_P, _P, C.c_int64]),
    "gpet_put_events_device": (C.c_int, [_P, _P, C.c_int64]),
    "gpet_peek_config_device": (C.c_int, [C.c_char_p]),
    "gpet_get_direction_table": (C.c_int64, [_P, _P, C.c_int64, _P]),
    "gpet_profile_enable": (C.c_int, [_P, C.c_int]),
    "gpet_profile_count": (C.c_int, [_P]),
    "gpet_profile_get": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
}

_lib = None


ABI_VERSION = 5   # GPET_ABI_VERSION of include/gpet_b200.h


def lib():
    """Load libgpet_b200.so (built in-tree by `make` / __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()); "
                               "gpet_b200 has no CPU fallback")
        l = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.gpet_abi_version() != ABI_VERSION:   # the ctypes structures below mirror the header of exactly this version
            raise RuntimeError(f"{LIB_PATH} has ABI version {l.gpet_abi_version()}, this module expects {ABI_VERSION}: rebuild (`make`)")
        _lib = l
    return _lib


class GpetError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gpet error {code}: {msg}")
        self.code = code


def _b(s):
    return None if s is None else os.fspath(s).encode()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One gpet_ctx.  device=-1 gives a host-only context (parsers/getters only)."""

    def __init__(self, device=0):
        self._l = lib()
        h = C.c_void_p()
        rc = self._l.gpet_create(int(device), C.byref(h))
        if rc != 0:
            raise GpetError(rc, "gpet_create failed (no usable CUDA device?)")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._l.gpet_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc < 0:
            raise GpetError(rc, (self._l.gpet_last_error(self._h) or b"").decode())
        return rc

    # ---- configuration
    def set_stream(self, cuda_stream_ptr):
        self._ck(self._l.gpet_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_seed(self, seed):
        self._ck(self._l.gpet_set_seed(self._h, seed))

    def set_capacity(self, photons, hits, events):
        self._ck(self._l.gpet_set_capacity(self._h, photons, hits, events))

    def load_config_file(self, path, base_dir=None, data_dir=None):
        self._ck(self._l.gpet_load_config_file(self._h, _b(path), _b(base_dir), _b(data_dir)))

    def load_tables(self, prefix):
        self._ck(self._l.gpet_load_tables(self._h, _b(prefix)))

    def save_tables_packed(self, path):
        self._ck(self._l.gpet_save_tables_packed(self._h, _b(path)))

    def load_phantom_files(self, mat, den, dim, offset, size):
        d = np.asarray(dim, np.int32); o = np.asarray(offset, np.float32); s = np.asarray(size, np.float32)
        self._ck(self._l.gpet_load_phantom_files(self._h, _b(mat), _b(den), _ptr(d), _ptr(o), _ptr(s)))

    def set_phantom(self, mat, dens, offset, size):
        """mat/dens: arrays of shape (nz, ny, nx) (x fastest, as in the .dat files)."""
        mat = np.ascontiguousarray(mat, np.int32); dens = np.ascontiguousarray(dens, np.float32)
        nz, ny, nx = mat.shape
        d = np.asarray([nx, ny, nz], np.int32); o = np.asarray(offset, np.float32); s = np.asarray(size, np.float32)
        self._ck(self._l.gpet_set_phantom(self._h, _ptr(mat), _ptr(dens), _ptr(d), _ptr(o), _ptr(s)))

    def load_geometry(self, path):
        self._ck(self._l.gpet_load_geometry(self._h, _b(path)))

    def load_isotopes(self, path):
        self._ck(self._l.gpet_load_isotopes(self._h, _b(path)))

    def load_source(self, path):
        self._ck(self._l.gpet_load_source(self._h, _b(path)))

    def load_psf(self, path, max_particles=0, ptype=1):
        self._ck(self._l.gpet_load_psf(self._h, _b(path), max_particles, ptype))

    def get_digitizer(self):
        p = DigitizerParams()
        self._ck(self._l.gpet_get_digitizer(self._h, C.byref(p)))
        return p

    def set_digitizer(self, p=None, **kw):
        if p is None:
            p = self.get_digitizer()
        for k, v in kw.items():
            setattr(p, k, v)
        self._ck(self._l.gpet_set_digitizer(self._h, C.byref(p)))

    def get_transport(self):
        p = TransportParams()
        self._ck(self._l.gpet_get_transport(self._h, C.byref(p)))
        return p

    def set_transport(self, p=None, **kw):
        if p is None:
            p = self.get_transport()
        for k, v in kw.items():
            if k in ("surface", "record_sphere"):
                for i, x in enumerate(v):
                    getattr(p, k)[i] = x
            else:
                setattr(p, k, v)
        self._ck(self._l.gpet_set_transport(self._h, C.byref(p)))

    def direction_table(self):
        """(table[32, 32, 32] indexed [iz, iy, ix], reference sphere (x, y, z, r)) or (None, None) when it does not apply."""
        ref = np.zeros(4, np.float64)
        n = self._ck(self._l.gpet_get_direction_table(self._h, None, 0, _ptr(ref)))
        if n == 0:
            return None, None
        tab = np.zeros(n, np.uint32)
        self._ck(self._l.gpet_get_direction_table(self._h, _ptr(tab), n, _ptr(ref)))
        nb = round(n ** (1 / 3))
        return tab.reshape(nb, nb, nb), ref

    def set_psf_output(self, mode):
        """OUTPUTPSF of the reference (constants.h:5): 0 none, 1 source photons in PSF mode, 2 source + phantom dumps."""
        self._ck(self._l.gpet_set_psf_output(self._h, mode))

    def set_time_window(self, t0, t1):
        self._ck(self._l.gpet_set_time_window(self._h, t0, t1))

    def set_source_atoms(self, i, natom):
        self._ck(self._l.gpet_set_source_atoms(self._h, i, natom))

    def profile(self, on=True):
        """Switch per-kernel CUDA-event timing on (clears earlier numbers) or off."""
        self._ck(self._l.gpet_profile_enable(self._h, 1 if on else 0))

    def kernel_times(self):
        """{kernel name: (total ms, launches)} accumulated since profile(True); synchronises the stream."""
        n = self._l.gpet_profile_count(self._h)
        if n < 0:
            self._ck(n)
        out = {}
        buf = C.create_string_buffer(128)
        for i in range(n):
            ms, cnt = C.c_double(), C.c_uint64()
            self._ck(self._l.gpet_profile_get(self._h, i, buf, 128, C.byref(ms), C.byref(cnt)))
            out[buf.value.decode()] = (ms.value, int(cnt.value))
        return out

    def set_shard(self, rank, world):
        self._ck(self._l.gpet_set_shard(self._h, rank, world))

    def set_emit_window(self, lo_us, hi_us, halo_start_us=float("-inf")):
        """digitize a time slice [lo, hi) given with its halo (gpet_set_emit_window); None clears"""
        self._ck(self._l.gpet_set_emit_window(self._h, float(lo_us), float(hi_us), float(halo_start_us)))

    def clear_emit_window(self):
        self._ck(self._l.gpet_clear_emit_window(self._h))

    def emit_counts(self):
        """(singles before the window, singles inside it, halo-too-short flag, coincidences emitted) of the last digitizer pass"""
        out = (C.c_uint64 * 4)()
        self._ck(self._l.gpet_get_emit_counts(self._h, out))
        return int(out[0]), int(out[1]), int(out[2]), int(out[3])

    def copy_events_to_device(self, dst_ptr, cap):
        return self._ck(self._l.gpet_copy_events_to_device(self._h, C.c_void_p(dst_ptr), int(cap)))

    def put_events_device(self, src_ptr, n):
        self._ck(self._l.gpet_put_events_device(self._h, C.c_void_p(src_ptr), int(n)))

    def set_first_pair(self, first_pair):
        """global 64-bit index of the acquisition's first annihilation pair (gpet_set_first_pair)"""
        self._ck(self._l.gpet_set_first_pair(self._h, int(first_pair)))

    def set_spectrum(self, nbins, emin, emax):
        self._ck(self._l.gpet_set_spectrum(self._h, nbins, emin, emax))

    # ---- getters
    def panels(self):
        n = self._ck(self._l.gpet_get_num_panels(self._h))
        out = np.zeros(n, PANEL_DTYPE)
        self._ck(self._l.gpet_get_panels(self._h, _ptr(out), n))
        return out

    def geometry_counts(self):
        c = np.zeros(4, np.int32); m = np.zeros(2, np.int32); d = np.zeros(2, np.float32)
        self._ck(self._l.gpet_get_geometry_counts(self._h, _ptr(c), _ptr(m), _ptr(d)))
        return c, m, d

    def table_dims(self):
        nmat = C.c_int32(); nen = C.c_int32(); e0 = C.c_float(); e1 = C.c_float()
        cmd = np.zeros(2, np.int32); cms = np.zeros(2, np.float32); rld = np.zeros(2, np.int32); rls = np.zeros(2, np.float32)
        self._ck(self._l.gpet_get_table_dims(self._h, C.byref(nmat), C.byref(nen), C.byref(e0), C.byref(e1),
                                              _ptr(cmd), _ptr(cms), _ptr(rld), _ptr(rls)))
        return dict(nmat=nmat.value, nen=nen.value, e0=e0.value, e1=e1.value, cm_ncp=int(cmd[0]), cm_ne=int(cmd[1]),
                    cm_dcp=float(cms[0]), cm_de=float(cms[1]), rl_ncp=int(rld[0]), rl_ne=int(rld[1]),
                    rl_dcp=float(rls[0]), rl_de=float(rls[1]))

    def table(self, which):
        d = self.table_dims()
        sizes = {0: d["nmat"] * d["nen"], 1: d["nmat"] * d["nen"], 2: d["nmat"] * d["nen"], 3: d["nmat"] * d["nen"],
                 4: d["nmat"] * d["cm_ncp"] * d["cm_ne"], 5: d["nmat"] * d["rl_ncp"] * d["rl_ne"],
                 6: d["nen"], 7: d["nen"], 8: d["nen"]}
        out = np.zeros(sizes[which], np.float32)
        n = self._ck(self._l.gpet_get_table(self._h, which, _ptr(out), out.size))
        return out[:n]

    def sources(self):
        n = self._ck(self._l.gpet_get_num_sources(self._h))
        res = []
        for i in range(n):
            na = C.c_uint64(); ty = C.c_int32(); sh = C.c_int32(); co = np.zeros(6, np.float32)
            self._ck(self._l.gpet_get_source(self._h, i, C.byref(na), C.byref(ty), C.byref(sh), _ptr(co)))
            res.append(dict(natom=na.value, type=ty.value, shape=sh.value, coeff=co))
        return res

    def isotopes(self):
        n = self._ck(self._l.gpet_get_num_isotopes(self._h))
        res = []
        for i in range(n):
            hl = C.c_float(); ra = C.c_float(); co = np.zeros(8, np.float32)
            self._ck(self._l.gpet_get_isotope(self._h, i, C.byref(hl), C.byref(ra), _ptr(co)))
            res.append(dict(halftime=hl.value, ratio=ra.value, coef=co))
        return res

    def num_psf(self):
        return self._ck(self._l.gpet_get_num_psf(self._h))

    # ---- stages
    def plan_frames(self, max_pairs=0):
        return self._ck(self._l.gpet_plan_frames(self._h, max_pairs))

    def frame_pairs(self, f):
        return self._ck(self._l.gpet_frame_pairs(self._h, f))

    def frame(self, f):
        ns = self._ck(self._l.gpet_get_num_sources(self._h))
        t0 = C.c_double(); dt = C.c_double(); fp = C.c_uint64(); pairs = np.zeros(ns, np.uint64)
        self._ck(self._l.gpet_get_frame(self._h, f, C.byref(t0), C.byref(dt), C.byref(fp), _ptr(pairs)))
        return dict(t0_s=t0.value, dt_s=dt.value, first_pair=fp.value, pairs=pairs)

    def stage_source(self, f):
        self._ck(self._l.gpet_stage_source(self._h, f))

    def stage_psf(self, first, n):
        self._ck(self._l.gpet_stage_psf(self._h, first, n))

    def stage_phantom(self):
        self._ck(self._l.gpet_stage_phantom(self._h))

    def stage_detector(self):
        self._ck(self._l.gpet_stage_detector(self._h))

    def stage_noise(self, t_lo_us, t_hi_us):
        self._ck(self._l.gpet_stage_noise(self._h, t_lo_us, t_hi_us))

    def stage_front(self, f=-1):
        """Fused source (frame f >= 0) or queue 0 (f = -1) -> phantom -> panel entry -> queue 2."""
        self._ck(self._l.gpet_stage_front(self._h, f))

    def stage_panel_transport(self):
        self._ck(self._l.gpet_stage_panel_transport(self._h))

    def stage_digitize(self):
        self._ck(self._l.gpet_stage_digitize(self._h))

    def queue_size(self, which):
        return self._ck(self._l.gpet_queue_size(self._h, which))

    def put_photons(self, which, photons):
        photons = np.ascontiguousarray(photons, PHOTON_DTYPE)
        self._ck(self._l.gpet_put_photons(self._h, which, _ptr(photons), photons.size))

    def fetch_photons(self, which):
        n = self.queue_size(which)
        out = np.zeros(n, PHOTON_DTYPE)
        n = self._ck(self._l.gpet_fetch_photons(self._h, which, _ptr(out), out.size))
        return out[:n]

    def put_events(self, events):
        events = np.ascontiguousarray(events, EVENT_DTYPE)
        self._ck(self._l.gpet_put_events(self._h, _ptr(events), events.size))

    def _fetch(self, fn, dtype, cap):
        out = np.zeros(cap, dtype)
        n = self._ck(fn(self._h, _ptr(out), out.size))
        return out[:n]

    def fetch_events(self, cap=1 << 22):
        return self._fetch(self._l.gpet_fetch_events, EVENT_DTYPE, cap)

    def fetch_hits(self, cap=1 << 22):
        return self._fetch(self._l.gpet_fetch_hits, HIT_DTYPE, cap)

    def fetch_singles(self, cap=1 << 22):
        return self._fetch(self._l.gpet_fetch_singles, EVENT_DTYPE, cap)

    def fetch_coincidences(self, cap=1 << 21):
        return self._fetch(self._l.gpet_fetch_coincidences, COINC_DTYPE, cap)

    TRUE, SCATTER, RANDOM = 0, 1, 2   # coincidence classes

    def fetch_coincidence_classes(self, cap=1 << 21):
        """(classes uint8[n], totals uint64[3] = trues, scatters, randoms) of the last frame's coincidences."""
        out = np.zeros(cap, np.uint8)
        totals = np.zeros(3, np.uint64)
        n = self._ck(self._l.gpet_fetch_coincidence_classes(self._h, _ptr(out), out.size, _ptr(totals)))
        return out[:n], totals

    def mark_scattered(self, parn):
        """Replay only: photons (by gpet_event.parn) that scattered in the phantom; call after put_events."""
        parn = np.ascontiguousarray(parn, np.int32)
        self._ck(self._l.gpet_mark_scattered(self._h, _ptr(parn), parn.size))

    def last_counts(self):
        c = np.zeros(4, np.uint64)
        self._ck(self._l.gpet_last_counts(self._h, _ptr(c)))
        return c

    # ---- whole path
    def digitize(self, events, out=None):
        """adder.dat-format list (host) -> (singles, counts[4]).  The bit-exact replay entry."""
        events = np.ascontiguousarray(events, EVENT_DTYPE)
        if out is None:
            out = np.zeros(max(events.size, 1), EVENT_DTYPE)
        n_out = C.c_int64()
        counts = np.zeros(4, np.uint64)
        self._ck(self._l.gpet_digitize(self._h, _ptr(events), events.size, _ptr(out), out.size, C.byref(n_out), _ptr(counts)))
        return out[:n_out.value], counts

    def run(self, output_dir=None):
        st = Stats()
        self._ck(self._l.gpet_run(self._h, _b(output_dir), C.byref(st)))
        return st

    def run_resident(self):
        st = Stats()
        self._ck(self._l.gpet_run_resident(self._h, C.byref(st)))
        return st

    def result_singles(self):
        p = C.c_void_p()
        n = self._ck(self._l.gpet_result_singles(self._h, C.byref(p)))
        if n == 0:
            return np.zeros(0, EVENT_DTYPE)
        buf = (C.c_char * (n * EVENT_DTYPE.itemsize)).from_address(p.value)
        return np.frombuffer(buf, EVENT_DTYPE, n).copy()

    def result_coincidences(self):
        p = C.c_void_p()
        n = self._ck(self._l.gpet_result_coincidences(self._h, C.byref(p)))
        if n == 0:
            return np.zeros(0, COINC_DTYPE)
        buf = (C.c_char * (n * COINC_DTYPE.itemsize)).from_address(p.value)
        return np.frombuffer(buf, COINC_DTYPE, n).copy()

    COINC_RECORDS, COINC_PAIRS = 0, 1
    SINGLES_RECORDS, SINGLES_COMPACT = 0, 1

    def set_singles_format(self, fmt):
        """How run(None) brings singles to the host: 48-byte records, or 32-byte gpet_single_compact records that
        result_singles() expands on demand (byte-identical)."""
        self._ck(self._l.gpet_set_singles_format(self._h, int(fmt)))

    def expand_singles(self, compact):
        """48-byte records from 32-byte compact singles (host only; needs the geometry and the digitizer parameters)."""
        compact = np.ascontiguousarray(compact, COMPACT_DTYPE)
        out = np.zeros(compact.size, EVENT_DTYPE)
        self._ck(self._l.gpet_expand_singles(self._h, _ptr(compact), compact.size, _ptr(out)))
        return out

    def result_singles_compact(self):
        p = C.c_void_p()
        n = self._ck(self._l.gpet_result_singles_compact(self._h, C.byref(p)))
        if n == 0:
            return np.zeros(0, COMPACT_DTYPE)
        buf = (C.c_char * (n * COMPACT_DTYPE.itemsize)).from_address(p.value)
        return np.frombuffer(buf, COMPACT_DTYPE, n).copy()

    def set_coincidence_format(self, fmt):
        self._ck(self._l.gpet_set_coincidence_format(self._h, int(fmt)))

    def result_coincidence_pairs(self):
        """(n, 2) uint32 indices into result_singles() (COINC_PAIRS mode)."""
        p = C.c_void_p()
        n = self._ck(self._l.gpet_result_coincidence_pairs(self._h, C.byref(p)))
        if n == 0:
            return np.zeros((0, 2), np.uint32)
        buf = (C.c_char * (n * 8)).from_address(p.value)
        return np.frombuffer(buf, np.uint32, 2 * n).reshape(n, 2).copy()

    def result_coincidence_classes(self):
        """uint8 class per coincidence of the last run (0 true, 1 scatter, 2 random)."""
        p = C.c_void_p()
        n = self._ck(self._l.gpet_result_coincidence_classes(self._h, C.byref(p)))
        if n == 0:
            return np.zeros(0, np.uint8)
        buf = (C.c_char * n).from_address(p.value)
        return np.frombuffer(buf, np.uint8, n).copy()

    def stats(self):
        st = Stats()
        self._ck(self._l.gpet_get_stats(self._h, C.byref(st)))
        return st

    def spectrum(self, nbins):
        out = np.zeros(nbins, np.uint64)
        self._ck(self._l.gpet_get_spectrum(self._h, _ptr(out), nbins))
        return out
