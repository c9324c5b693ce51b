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
