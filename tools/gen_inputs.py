#!/usr/bin/env python
"""Synthetic inputs for the BASELINE.json configurations (the reference checkout lacks the phantom and PSF blobs,
SURVEY F6): phantom volumes in the reference's mat.dat/den.dat layout (int32 / float32, x fastest,
initialize.cu:55-66) and psf.dat records (8 x fp64, initialize.cu:79-104)."""
from __future__ import annotations

import argparse
from pathlib import Path

import numpy as np

AIR_RHO = np.float32(1.2048e-3)


def cylinder_phantom(n=200, size=1.0, radius=0.5, water_mat=1, air_mat=0):
    """Config 1: water cylinder (axis z, full height) of `radius` cm in air, n^3 voxels over `size` cm."""
    c = (np.arange(n, dtype=np.float64) + 0.5) * (size / n) - size / 2
    inside = (c[None, :] ** 2 + c[:, None] ** 2) <= radius * radius  # [y, x]
    mat2 = np.where(inside, water_mat, air_mat).astype(np.int32)
    den2 = np.where(inside, np.float32(1.0), AIR_RHO).astype(np.float32)
    mat = np.broadcast_to(mat2[None], (n, n, n)).copy()
    den = np.broadcast_to(den2[None], (n, n, n)).copy()
    return mat, den  # [z, y, x]


def air_phantom(n=200):
    return np.zeros((n, n, n), np.int32), np.full((n, n, n), AIR_RHO, np.float32)


def uniform_phantom(n, mat_id, rho):
    return np.full((n, n, n), mat_id, np.int32), np.full((n, n, n), rho, np.float32)


def write_phantom(mat, den, mat_path, den_path):
    np.ascontiguousarray(mat, "<i4").tofile(mat_path)
    np.ascontiguousarray(den, "<f4").tofile(den_path)


def mouse_phantom(n=256, size=(3.2, 3.2, 6.4), water_mat=1, bone_mat=3, air_mat=0):
    """Config 4: water ellipsoid with semi-axes (1.4, 1.4, 3.0) cm, CorticalBone (rho 1.85) cylinder of r = 0.15 cm along z
    at (0, -0.8), air elsewhere; n^3 voxels over `size` cm centred on the origin.  Returns [z, y, x] volumes."""
    cx = (np.arange(n, dtype=np.float64) + 0.5) * (size[0] / n) - size[0] / 2
    cy = (np.arange(n, dtype=np.float64) + 0.5) * (size[1] / n) - size[1] / 2
    cz = (np.arange(n, dtype=np.float64) + 0.5) * (size[2] / n) - size[2] / 2
    X, Y, Z = cx[None, None, :], cy[None, :, None], cz[:, None, None]
    water = (X / 1.4) ** 2 + (Y / 1.4) ** 2 + (Z / 3.0) ** 2 <= 1.0
    bone = ((X - 0.0) ** 2 + (Y + 0.8) ** 2 <= 0.15 ** 2) & water
    mat = np.full((n, n, n), air_mat, np.int32)
    den = np.full((n, n, n), AIR_RHO, np.float32)
    mat[water] = water_mat; den[water] = 1.0
    mat[bone] = bone_mat; den[bone] = 1.85
    return mat, den


def water_cylinder_phantom(n=256, voxel=0.1, diameter=20.0, height=20.0, water_mat=1, air_mat=0):
    """Config 5: `diameter` x `height` cm water cylinder (axis z) centred in an n^3 grid of `voxel` cm voxels."""
    c = (np.arange(n, dtype=np.float64) + 0.5) * voxel - n * voxel / 2
    disk = (c[None, :] ** 2 + c[:, None] ** 2) <= (diameter / 2) ** 2   # [y, x]
    slab = np.abs(c) <= height / 2                                      # [z]
    inside = slab[:, None, None] & disk[None]
    mat = np.where(inside, water_mat, air_mat).astype(np.int32)
    den = np.where(inside, np.float32(1.0), AIR_RHO).astype(np.float32)
    return mat, den


def ring_geo(npanels=32, radius=40.0, modules_y=4, modules_z=13, crystal_mat=7, crystal_rho=7.4):
    """Config 5: a ring of `npanels` panels at `radius` cm about z in the reference's .geo layout (detector.cu:76-207):
    the same 1.75 cm modules (0.02 gap) of 8x8 crystals of 0.21 cm (0.01 gap) as config8.geo; only panel 0 is described,
    the others are rotated copies."""
    ly = modules_y * 1.75 + (modules_y - 1) * 0.02
    lz = modules_z * 1.75 + (modules_z - 1) * 0.02
    return f"""number of panels
{npanels}
rotation axis (global frame)
0 0 1
rotation step between panels (degrees, counter-clockwise positive)
{360.0 / npanels:g}
material id and density (g/cm3): crystal, then gap
{crystal_mat}          {crystal_rho:g}
0        0.001025
-------------------------------------------------
index of the first panel
0
panel size x y z (cm)
2 {ly:.2f} {lz:.2f}
module size x y z (cm)
2 1.75 1.75
module gap x y z (cm)
2 0.02 0.02
crystal size x y z (cm)
2 0.21 0.21
crystal gap x y z (cm)
2 0.01 0.01
growth direction along local x y z
-1  1 1
centre of the panel face looking at the phantom (cm)
0  {-radius:g}   0
local x axis in the global frame
0  1   0
local y axis in the global frame
1  0   0
local z axis in the global frame
0   0  -1
-------------------------------------------------
"""


def source_file(rows):
    """source.txt text (initialize.cu:120-140): rows of (natom, isotope row, shape, c0..c5)."""
    out = [str(len(rows)), "atoms, isotope row, shape (0 box, 1 cylinder, 2 sphere), centre x y z (cm), three shape parameters#"]
    for r in rows:
        out.append(" ".join(str(v) for v in r))
    return "\n".join(out) + "\n"


def atoms_for_decays(decays, halflife_s, window_s, ratio=1.0):
    """natom such that the expected number of emitted pairs in [0, window_s] is `decays`."""
    frac = -np.expm1(-window_s * np.log(2.0) / halflife_s)
    return int(round(decays / (frac * ratio)))


def input_file(dims=(200, 200, 200), offset=(-0.5, -0.5, -0.5), extent=(1, 1, 1), mat="input/cylinder_phantom_mat.dat",
               den="input/cylinder_phantom_den.dat", nhist=1000000, usepsf=0, source="input/pointsource.txt", ptype=0, prange=0,
               window=(0, 120), eabs=1e3, geo="input/config8.geo", readout=(2, 1), threshold=50000, blur=(1, 662000, 0.05, 0, 0),
               deadtime=(3, 0, 2.2), ewin=(30000, 700000), acollinearity=0.0037056):
    """input_PET.in text in the reference's positional label/value layout (main.cu:52-182)."""
    j = lambda v: " ".join(f"{x:g}" if isinstance(x, float) else str(x) for x in v)   # noqa: E731
    return f"""GPU index:
0
annihilation photon acollinearity, Gaussian sigma (rad):
{acollinearity:g}
phantom voxel counts nx ny nz:
{j(dims)}
phantom corner offset x y z (cm):
{j(offset)}
phantom extent x y z (cm):
{j(extent)}
phantom material volume (int32, x fastest):
{mat}
phantom density volume (float32, x fastest):
{den}
number of phase-space histories (PSF mode only):
{nhist}
read a phase-space file as source (0 no, 1 yes):
{usepsf}
source description file:
{source}
phase-space particle type (0 positron, 1 photon):
{ptype}
positron range (0 off, 1 on):
{prange}
acquisition start and end (s):
{j(window)}
photon-PSF recording sphere centre x y z and radius (cm):
0 0 0 5
photon absorption energy (eV):
{eabs:g}
detector geometry file:
{geo}
quadric exclusion surfaces: count, then ten coefficients each (x2 y2 z2 xy xz yz x y z 1):
1
0 0 0 0 0 0 0 0 0 1
readout depth and policy:
{j(readout)}
energy threshold before dead time (eV):
{threshold:g}
energy blur policy, reference energy (eV), reference resolution, slope (1/MeV), spatial blur sigma (cm):
{j(blur)}
dead-time level, type (0 paralyzable, 1 non-paralyzable), duration (us):
{j(deadtime)}
energy window lower and upper bound (eV):
{j(ewin)}
"""


def back_to_back_psf(npairs, seed=20201001, energy=511000.0, dt_us=1.0):
    """Config 2: back-to-back 511 keV pairs from the origin, isotropic, pair k at t = (k+1)*dt_us."""
    rng = np.random.default_rng(seed)
    ct = rng.uniform(-1.0, 1.0, npairs)
    phi = rng.uniform(0.0, 2 * np.pi, npairs)
    st = np.sqrt(1 - ct * ct)
    v = np.stack([st * np.cos(phi), st * np.sin(phi), ct], axis=1)
    rec = np.zeros((2 * npairs, 8), "<f8")
    t = (np.arange(npairs) + 1) * dt_us
    rec[0::2, 3] = t
    rec[1::2, 3] = t
    rec[0::2, 4:7] = v
    rec[1::2, 4:7] = -v
    rec[:, 7] = energy
    return rec


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True, help="directory that receives input/cylinder_phantom_{mat,den}.dat and input/psf.dat")
    ap.add_argument("--psf-pairs", type=int, default=1000000)
    a = ap.parse_args()
    out = Path(a.out) / "input"
    out.mkdir(parents=True, exist_ok=True)
    m, d = cylinder_phantom()
    write_phantom(m, d, out / "cylinder_phantom_mat.dat", out / "cylinder_phantom_den.dat")
    back_to_back_psf(a.psf_pairs).tofile(out / "psf.dat")
    print("wrote", out)


# ------------------------------------------------------------------------------------------------ named configurations
# The BASELINE.json configurations as files a gPET user would write, shared by tests/test_reference_stats.py,
# tools/ref_stats.py (which runs the REFERENCE BINARY on them) and the scale runs.  `decays` = expected annihilation
# pairs in the 0-120 s window (F-18 row 0 of isotopes.txt: T1/2 6586.26 s, branching 0.97).
STATS_CONFIGS = ("config1_water10", "config2_psf", "config4_mouse", "config5_ring")


def stats_config(name, decays, blur=(1, 662000, 0.05, 0, 0), deadtime=(3, 0, 2.2)):
    """Returns dict(text, phantom=(mat, den) | None, geo_text | None, extra={file: text}, psf=records | None, npanels, moduleN)."""
    natom = atoms_for_decays(decays, 6586.26, 120.0, 0.97)
    if name == "config1_water10":
        # shipped 8-panel geometry, but a phantom that scatters: 10 cm water cylinder (200^3 over 10 cm), F-18 rod r 0.5 x 8 cm
        n = 200
        mat, den = cylinder_phantom(n=n, size=10.0, radius=5.0)
        text = input_file(dims=(n, n, n), offset=(-5.0,) * 3, extent=(10.0,) * 3, mat="input/phantom_mat.dat", den="input/phantom_den.dat",
                          source="input/src.txt", blur=blur, deadtime=deadtime)
        return dict(text=text, phantom=(mat, den), geo_text=None, extra={"src.txt": source_file([(natom, 0, 1, 0, 0, 0, 0.5, 8.0, 0)])},
                    psf=None, npanels=8, moduleN=117)
    if name == "config2_psf":
        mat, den = air_phantom(32)
        text = input_file(dims=(32, 32, 32), mat="input/phantom_mat.dat", den="input/phantom_den.dat", usepsf=1, source="input/psf.dat",
                          ptype=1, nhist=2 * decays, blur=blur, deadtime=deadtime)
        return dict(text=text, phantom=(mat, den), geo_text=None, extra={}, psf=back_to_back_psf(decays), npanels=8, moduleN=117)
    if name == "config4_mouse":
        n = 256
        size = (3.2, 3.2, 6.4)
        text = input_file(dims=(n, n, n), offset=tuple(-x / 2 for x in size), extent=size, mat="input/phantom_mat.dat",
                          den="input/phantom_den.dat", source="input/src.txt", blur=blur, deadtime=deadtime)
        return dict(text=text, phantom=mouse_phantom(n, size), geo_text=None,
                    extra={"src.txt": source_file([(natom, 0, 1, 0, 0, 0, 1.2, 5.0, 0)])}, psf=None, npanels=8, moduleN=117)
    if name == "config5_ring":
        n = 256
        text = input_file(dims=(n, n, n), offset=(-12.8,) * 3, extent=(25.6,) * 3, mat="input/phantom_mat.dat", den="input/phantom_den.dat",
                          source="input/src.txt", geo="input/ring.geo", blur=blur, deadtime=deadtime)
        return dict(text=text, phantom=water_cylinder_phantom(n, 0.1, 20.0, 20.0), geo_text=ring_geo(32, 40.0),
                    extra={"src.txt": source_file([(natom, 0, 1, 0, 0, 0, 0.5, 18.0, 0)])}, psf=None, npanels=32, moduleN=52)
    raise KeyError(name)


def write_workdir(ex, cfg, example_dir, packed_tables=None):
    """Lay `cfg` (stats_config) out under `ex` the way the reference expects: input_PET.in, input/, data/, output/."""
    ex = Path(ex)
    (ex / "input").mkdir(parents=True, exist_ok=True)
    (ex / "data").mkdir(exist_ok=True)
    (ex / "output").mkdir(exist_ok=True)
    example_dir = Path(example_dir)
    (ex / "input_PET.in").write_text(cfg["text"])
    (ex / "input" / "config8.geo").write_text((example_dir / "input" / "config8.geo").read_text())
    (ex / "data" / "isotopes.txt").write_text((example_dir / "data" / "isotopes.txt").read_text())
    if packed_tables is not None and not (ex / "data" / "input4gPET.gpettab").exists():
        (ex / "data" / "input4gPET.gpettab").symlink_to(packed_tables)
    if cfg["phantom"] is not None:
        write_phantom(cfg["phantom"][0], cfg["phantom"][1], ex / "input" / "phantom_mat.dat", ex / "input" / "phantom_den.dat")
    if cfg["geo_text"] is not None:
        (ex / "input" / "ring.geo").write_text(cfg["geo_text"])
    for name, content in cfg["extra"].items():
        (ex / "input" / name).write_text(content)
    if cfg["psf"] is not None:
        np.ascontiguousarray(cfg["psf"], "<f8").tofile(ex / "input" / "psf.dat")
    return ex
