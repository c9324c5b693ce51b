#!/usr/bin/env python
"""Regenerate the two cross-section files the reference checkout lacks:
`input4gPET.cmpsf` (Compton inverse-CDF surfaces with binding, from the incoherent scattering function S(q)) and
`input4gPET.rayff` (Rayleigh inverse-CDF surfaces from the form factor F(q)), in the exact text layout that
rcmpsf / rrayff parse (initialize.cu:483-566, 664-748).

Recipe (validated against the complete input4gCTD set, see tests/test_tables.py):
  Compton : pdf(cos t) ~ KN(k, cos t) * S(q),  k = E/mc^2, eps = 1/(1 + k(1-cos t)),
            KN = eps^2 (eps + 1/eps - sin^2 t),  q = sqrt(k^2 + k'^2 - 2 k k' cos t), k' = k eps
  Rayleigh: pdf(cos t) ~ (1 + cos^2 t) * F(q)^2,  q = k sqrt(2 (1 - cos t))
  S and F are interpolated in log q on the 64-point grid of the files; the cumulative from cos t = -1 is inverted on
  the CP grid.  Column E = 0 duplicates column 1 (as in the shipped CTD surfaces).

S(q)/F(q) blocks exist in this container only for the CTD materials.  PET materials without a block borrow the
closest one (documented in MATERIAL_SOURCE below); configs 1-3 of BASELINE.json use only DryAir, Water and LSO,
which are exact.
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gpet_b200 import refio  # noqa: E402

MC2 = 510.9991e3

# PET material -> CTD material providing S(q)/F(q)   (exact when the names coincide)
MATERIAL_SOURCE = {
    "DryAir": "DryAir", "Water": "Water", "PMMA": "PMMA", "LSO": "LSO", "LYSO": "LYSO",
    "TissueICRU": "TissueICRP",      # near-identical soft tissue composition
    "MuscleStriated": "TissueICRP",  # approximation (soft tissue)
    "Brain": "TissueICRP",           # approximation (soft tissue)
    "CorticalBone": "Water",         # approximation: no S/F source here; only the angular shape is affected
    "Pb": "LSO",                     # approximation: heaviest available block
}


def _interp_logq(qgrid, vals, q):
    lq = np.log(np.maximum(q, qgrid[0]))
    return np.interp(lq, np.log(qgrid), vals)


def compton_surface(qgrid, sq, ncp, ne, de, ncos=8001):
    cp = np.linspace(0.0, 1.0, ncp)
    cos_t = np.linspace(-1.0, 1.0, ncos)
    surf = np.zeros((ncp, ne), np.float64)
    for ie in range(1, ne):
        k = ie * de / MC2
        eps = 1.0 / (1.0 + k * (1.0 - cos_t))
        kn = eps * eps * (eps + 1.0 / eps - (1.0 - cos_t * cos_t))
        kp = k * eps
        q = np.sqrt(np.maximum(k * k + kp * kp - 2.0 * k * kp * cos_t, 0.0))
        pdf = kn * _interp_logq(qgrid, sq, q)
        cdf = np.concatenate([[0.0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(cos_t))])
        cdf /= cdf[-1]
        surf[:, ie] = np.interp(cp, cdf, cos_t)
    surf[:, 0] = surf[:, 1]
    surf[0, :] = -1.0
    surf[-1, :] = 1.0
    return surf


def rayleigh_surface(qgrid, fq, ncp, ne, de, ncos=8001):
    cp = np.linspace(0.0, 1.0, ncp)
    # forward peaked at high energy: refine the grid towards cos t = 1
    u = np.linspace(0.0, 1.0, ncos)
    cos_t = 1.0 - 2.0 * u ** 3
    cos_t = cos_t[::-1]
    surf = np.zeros((ncp, ne), np.float64)
    for ie in range(1, ne):
        k = ie * de / MC2
        q = k * np.sqrt(2.0 * (1.0 - cos_t))
        f = _interp_logq(qgrid, fq, q)
        pdf = (1.0 + cos_t * cos_t) * f * f
        cdf = np.concatenate([[0.0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(cos_t))])
        cdf /= cdf[-1]
        surf[:, ie] = np.interp(cp, cdf, cos_t)
    # E -> 0: F(q) -> Z, pure Thomson (1 + cos^2 t) law (this is what column 0 of the shipped CTD file holds)
    pdf = 1.0 + cos_t * cos_t
    cdf = np.concatenate([[0.0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(cos_t))])
    surf[:, 0] = np.interp(cp, cdf / cdf[-1], cos_t)
    surf[0, :] = -1.0
    surf[-1, :] = 1.0
    return surf


def write_surface_file(path, set_name, label, blocks, surfaces, ncp, ne, emax):
    dcp = 1.0 / (ncp - 1)
    de = emax / (ne - 1)
    with open(path, "w") as f:
        f.write(" This file is part of set:\n %s\n" % set_name)
        for qb, surf in zip(blocks, surfaces):
            nd = qb.shape[0]
            f.write(" ndata, logqmin, logqmax, dlogq:\n")
            f.write("  %d  %.6e  %.6e  %.6e\n" % (nd, qb[0, 1], qb[-1, 1], (qb[-1, 1] - qb[0, 1]) / (nd - 1)))
            f.write("  q(m_e*c) -- logq -- %s\n" % label)
            for r in qb:
                f.write(" %.6e  %.6e  %.6e\n" % (r[0], r[1], r[2]))
            f.write("\n nCP, CPmin, CPmax, dCP, nE, Emin, Emax, dE:\n")
            f.write("  %d %.6e %.6e %.6e %d %.6e %.6e %.6e\n" % (ncp, 0.0, 1.0, dcp, ne, 0.0, emax, de))
            f.write(" " + " ".join("%.6e" % (i * dcp) for i in range(ncp)) + " \n")
            f.write(" " + " ".join("%.6e" % (i * de) for i in range(ne)) + " \n")
            for row in surf:
                f.write(" " + " ".join("%.6e" % v for v in row) + " \n")
            f.write("\n")


def generate(ref_data_dir, out_dir, set_name="input4gPET", ncp=301, ne=151):
    ref = Path(ref_data_dir)
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    pet = refio.read_matter(ref / "input4gPET.matter")
    ctd = refio.read_matter(ref / "input4gCTD.matter")
    ctd_cm = refio.read_surface(ref / "input4gCTD.cmpsf", ctd["nmat"])
    ctd_rl = refio.read_surface(ref / "input4gCTD.rayff", ctd["nmat"])
    emax = float(pet["emax"])
    de = emax / (ne - 1)
    cm_blocks, rl_blocks, cm_surf, rl_surf = [], [], [], []
    cache = {}
    for name in pet["names"]:
        src = MATERIAL_SOURCE.get(name)
        if src is None or src not in ctd["names"]:
            raise SystemExit(f"no S(q)/F(q) source for material {name}")
        k = ctd["names"].index(src)
        if src not in cache:
            sq, fq = ctd_cm["sq"][k].astype(np.float64), ctd_rl["sq"][k].astype(np.float64)
            cache[src] = (compton_surface(sq[:, 0], sq[:, 2], ncp, ne, de), rayleigh_surface(fq[:, 0], fq[:, 2], ncp, ne, de))
        cm_blocks.append(ctd_cm["sq"][k]); rl_blocks.append(ctd_rl["sq"][k])
        cm_surf.append(cache[src][0]); rl_surf.append(cache[src][1])
    write_surface_file(out / f"{set_name}.cmpsf", set_name, "sf", cm_blocks, cm_surf, ncp, ne, emax)
    write_surface_file(out / f"{set_name}.rayff", set_name, "ff", rl_blocks, rl_surf, ncp, ne, emax)
    return out / f"{set_name}.cmpsf", out / f"{set_name}.rayff"


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--ref-data", default="/root/reference/data")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    print(*generate(a.ref_data, a.out))
