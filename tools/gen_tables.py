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

S(q)/F(q) blocks exist in this container only for the CTD materials (DryAir, Water, PMMA, PE, LSO, LYSO, TissueICRP).
PET materials with the same name take their block as it is.  The others (TissueICRU, CorticalBone, MuscleStriated, Brain,
Pb) are MIXED from elemental functions by the additivity rule the CTD blocks themselves obey:

    S_material(q) = sum_i n_i S_i(q),      F_material(q)^2 = sum_i n_i F_i(q)^2       (n_i: atoms of element i per molecule)

The elemental functions are not shipped either; they are recovered from the shipped compound blocks, which the additivity
rule makes a linear system per q:  H, C, O exactly from Water (H2O), PE (C2H4) and PMMA (C5H8O2) -- the solution has
S_H -> 1, S_C -> 6, S_O -> 8 at large q and F_H(0)^2 = 1, F_C(0)^2 = 36, F_O(0)^2 = 64, i.e. the blocks really are
independent-atom sums --, then N from DryAir, Lu from LSO and Y from LYSO.  Elements that occur in no shipped compound
(Na, Mg, Si, P, S, Cl, Ar, K, Ca, Fe, Zn, Pb) are scaled from the nearest recovered element with the Thomas-Fermi rule
S_Z(q) = Z s(q Z^-2/3), F_Z(q) = Z f(q Z^-1/3): an approximation that ignores shell structure, flagged in MIXED below and in
DESIGN.md; it enters configs 4 (bone: Ca, P) only.  tests/test_oracle_and_host.py checks the mixing against the shipped
TissueICRP and LYSO blocks.
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gpet_b200 import refio  # noqa: E402

MC2 = 510.9991e3

# PET materials whose S(q)/F(q) block is shipped under the same name in the CTD set; everything else is mixed
SHIPPED_BLOCK = ("DryAir", "Water", "PMMA", "LSO", "LYSO")
# element -> the recovered element it is Thomas-Fermi scaled from
TF_PARENT = {11: 8, 12: 8, 14: 8, 15: 8, 16: 8, 17: 8, 18: 8, 19: 8, 20: 8, 26: 8, 30: 8, 39: 71, 82: 71}


def read_compositions(matter_path):
    """{material: [(Z, atoms per molecule), ...]} from a *.matter file (`No elements in molecule:` blocks)."""
    out, name, lines = {}, None, Path(matter_path).read_text().splitlines()
    i = 0
    while i < len(lines):
        ln = lines[i].strip()
        if ln.startswith("MATERIAL:"):
            name = ln.split(":", 1)[1].strip()
        elif ln.startswith("No elements in molecule") and name:
            n = int(lines[i + 1].split()[0])
            out[name] = [(int(lines[i + 2 + k].split()[0]), float(lines[i + 2 + k].split()[1])) for k in range(n)]
            i += 1 + n
        i += 1
    return out


def tf_scaled(q, s_ref, f2_ref, z_ref, z):
    """Thomas-Fermi scaling of an element's S(q) and F(q)^2 from element z_ref to element z (log-q interpolation, clamped)."""
    lq = np.log(q)
    s = (z / z_ref) * np.interp(lq + (2.0 / 3.0) * np.log(z_ref / z), lq, s_ref)
    f2 = (z / z_ref) ** 2 * np.interp(lq + (1.0 / 3.0) * np.log(z_ref / z), lq, f2_ref)
    return s, f2


def elemental_functions(q, S, F2, comp):
    """{Z: (S_Z(q), F_Z(q)^2)} recovered from the shipped compound blocks (S, F2: {compound: array}; comp: compositions)."""
    el = {}
    for T, k in ((S, 0), (F2, 1)):
        H = (2.5 * T["PE"] + 2.0 * T["Water"] - T["PMMA"]) / 6.0      # Water = 2H + O, PE = 2C + 4H, PMMA = 5C + 8H + 2O
        C = (T["PE"] - 4.0 * H) / 2.0
        O = T["Water"] - 2.0 * H
        for z, v in ((1, H), (6, C), (8, O)):
            el.setdefault(z, [None, None])[k] = np.maximum(v, 0.0)

    def scaled(z):
        if z not in el:
            p = TF_PARENT[z]
            if p not in el:
                raise KeyError(z)
            el[z] = list(tf_scaled(q, el[p][0], el[p][1], p, z))
        return el[z]

    def solve_for(z, compound):
        rest_s, rest_f = np.zeros_like(q), np.zeros_like(q)
        n_z = 0.0
        for zz, n in comp[compound]:
            if zz == z:
                n_z = n
            else:
                e = scaled(zz)
                rest_s += n * e[0]; rest_f += n * e[1]
        el[z] = [np.maximum((S[compound] - rest_s) / n_z, 0.0), np.maximum((F2[compound] - rest_f) / n_z, 0.0)]

    solve_for(7, "DryAir")     # 78 % of the atoms of air; C and O recovered above, Ar scaled
    solve_for(71, "LSO")       # Lu2 Si O5 with Si scaled from O
    solve_for(39, "LYSO")      # Lu1.8 Y0.2 Si O5
    for z in TF_PARENT:
        scaled(z)
    return {z: (v[0], v[1]) for z, v in el.items()}


def mix(elements, composition):
    """(S(q), F(q)) of a material from its composition by the additivity rule"""
    s = sum(n * elements[z][0] for z, n in composition)
    f2 = sum(n * elements[z][1] for z, n in composition)
    return s, np.sqrt(np.maximum(f2, 0.0))


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


def material_blocks(ref_data_dir):
    """For every PET material, in file order: (name, q-block of the .cmpsf file [q, ln q, S], q-block of the .rayff file
    [q, ln q, F], how it was obtained)."""
    ref = Path(ref_data_dir)
    pet = refio.read_matter(ref / "input4gPET.matter")
    ctd = refio.read_matter(ref / "input4gCTD.matter")
    ctd_cm = refio.read_surface(ref / "input4gCTD.cmpsf", ctd["nmat"])
    ctd_rl = refio.read_surface(ref / "input4gCTD.rayff", ctd["nmat"])
    q = ctd_cm["sq"][0][:, 0].astype(np.float64)
    S = {n: ctd_cm["sq"][i][:, 2].astype(np.float64) for i, n in enumerate(ctd["names"])}
    F2 = {n: ctd_rl["sq"][i][:, 2].astype(np.float64) ** 2 for i, n in enumerate(ctd["names"])}
    elements = elemental_functions(q, S, F2, read_compositions(ref / "input4gCTD.matter"))
    pet_comp = read_compositions(ref / "input4gPET.matter")
    out = []
    for name in pet["names"]:
        if name in SHIPPED_BLOCK and name in ctd["names"]:
            k = ctd["names"].index(name)
            out.append((name, ctd_cm["sq"][k].astype(np.float64), ctd_rl["sq"][k].astype(np.float64), "shipped block"))
            continue
        if name not in pet_comp:
            raise SystemExit(f"no composition for material {name}")
        s, f = mix(elements, pet_comp[name])
        scaled = sorted(z for z, _ in pet_comp[name] if z in TF_PARENT)
        how = "mixed from recovered elements" + (f" (Thomas-Fermi scaled: Z = {scaled})" if scaled else "")
        lq = np.log(q)
        out.append((name, np.stack([q, lq, s], 1), np.stack([q, lq, f], 1), how))
    return out, pet


def generate(ref_data_dir, out_dir, set_name="input4gPET", ncp=301, ne=151):
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    blocks, pet = material_blocks(ref_data_dir)
    emax = float(pet["emax"])
    de = emax / (ne - 1)
    cm_blocks, rl_blocks, cm_surf, rl_surf = [], [], [], []
    for name, sq, fq, how in blocks:
        cm_blocks.append(sq); rl_blocks.append(fq)
        cm_surf.append(compton_surface(sq[:, 0], sq[:, 2], ncp, ne, de))
        rl_surf.append(rayleigh_surface(fq[:, 0], fq[:, 2], ncp, ne, de))
    write_surface_file(out / f"{set_name}.cmpsf", set_name, "sf", cm_blocks, cm_surf, ncp, ne, emax)
    write_surface_file(out / f"{set_name}.rayff", set_name, "ff", rl_blocks, rl_surf, ncp, ne, emax)
    return out / f"{set_name}.cmpsf", out / f"{set_name}.rayff"


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--ref-data", default="/root/reference/data")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    print(*generate(a.ref_data, a.out))
