/*
 * gpet_oracle.c -- CPU restatement of gPET's Monte-Carlo hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or call
 * this file.  The product (libgpet_b200.so) never links or calls it and has no CPU fallback.
 *
 * What it restates (file:line under /root/reference):
 *   source sampling   setPosition / getPositionFromShape      gPET_kernals.cu:445-561
 *   phantom transport photon, getAbsVox, table lookups, comsam (table), rylsam, rotate   gPET_kernals.cu:19-345
 *   detector          photonde, crystalSearch, comsam (Klein-Nishina), adder, readout    gPET_kernals.cu:737-813, 839-1279
 *   digitizer         blur, energywindow, setSitenum, deadtime + host sorts / orderevents  gPET_kernals.cu:607-698, 814-837;
 *                     gPET.cu:380-424; detector.cu:309-385
 * Control flow follows the reference (one photon at a time, while(1) loops, three sorts in the digitizer), NOT the
 * restructured CUDA product.  Two deliberate, documented differences from the reference, shared with the product
 * because the reference's own behaviour is not reproducible:
 *   - random numbers: Philox4x32-10 (Salmon et al. SC'11, Random123) streams keyed by photon id instead of
 *     time-seeded XORWOW state; one 4-word block per Woodcock flight / rejection round (DESIGN.md "RNG contract");
 *   - dead time uses the race-free "snapshot-start" semantics of SURVEY 8(a) D7.
 *
 * PARITY PINNING: the reference ships no tests, golden vectors or fixtures.  Philox is pinned on the
 * published Random123 known-answer vectors (tests/test_oracle.py); the digitizer is pinned on hand-derived
 * known-answer cases written from the reference source (tests/golden/digitizer_kat.json) and -- when the patched
 * reference binary has been run on a GPU box -- on its adder.dat -> singles.dat pairs (tests/golden/ref_*).
 * Transport is stochastic and time-seeded in the reference: parity there is statistical ("parity unpinned" for
 * the bit level).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXT 1e20
#define ORC_MC2 510.9991e3f
#define ORC_IMC2 1.95695060911e-6f
#define ORC_TWOPI 6.2831853071795864769252867f
#define ORC_SPE 29979.2458 /* cm/us, gPET_kernals.cu:264 */

/* ------------------------------------------------------------------------------------------------ Philox4x32-10 */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t x0 = ctr[0], x1 = ctr[1], x2 = ctr[2], x3 = ctr[3], a = key[0], b = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * x0, p1 = (uint64_t)0xCD9E8D57u * x2;
        uint32_t y0 = (uint32_t)(p1 >> 32) ^ x1 ^ a, y1 = (uint32_t)p1;
        uint32_t y2 = (uint32_t)(p0 >> 32) ^ x3 ^ b, y3 = (uint32_t)p0;
        x0 = y0; x1 = y1; x2 = y2; x3 = y3;
        a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
}

typedef struct { uint32_t key[2], ctr[4]; } orc_rng;

static void rng_init(orc_rng* g, uint64_t seed, uint64_t index, uint32_t stream) {
    g->key[0] = (uint32_t)seed; g->key[1] = (uint32_t)(seed >> 32);
    g->ctr[0] = (uint32_t)index; g->ctr[1] = (uint32_t)(index >> 32); g->ctr[2] = stream; g->ctr[3] = 0;
}
static void rng_next(orc_rng* g, uint32_t r[4]) { orc_philox4x32_10(g->ctr, g->key, r); g->ctr[3]++; }
/* (0,1] like curand_uniform */
static float u01(uint32_t x) { return fmaf((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f); }
static double u01d(uint32_t a, uint32_t b) {
    uint64_t k = ((uint64_t)a << 21) | (uint64_t)(b >> 11);
    return (double)k * 1.1102230246251565e-16 + 5.551115123125783e-17;
}

/* 64-bit history numbers (gpet_b200.h gpet_set_first_pair; the reference stops at 32 bits, gPET.h:50): records keep the low
 * 31 bits, Philox streams are keyed by the full index, rebuilt from the record and the index of the frame's first photon.
 * id_base == 0: the 32-bit id as it is.  Set by orc_set_id_base() before a call (test infrastructure: one global). */
static uint64_t g_id_base = 0;
void orc_set_id_base(uint64_t base) { g_id_base = base; }
static uint64_t photon_index(int32_t parn) {
    if (g_id_base == 0) return (uint64_t)(uint32_t)parn;
    return g_id_base + (uint64_t)(((uint32_t)parn - (uint32_t)g_id_base) & 0x7fffffffu);
}

enum { ST_SOURCE = 1, ST_PHANTOM = 2, ST_DETECTOR = 3, ST_BLUR = 4, ST_PLAN = 5, ST_PSF_POSITRON = 6, ST_NOISE = 7 };

/* ------------------------------------------------------------------------------------------------ records */
typedef struct { int32_t parn, pann, modn, cryn, siten, eventid; double t; float E, x, y, z; } orc_event; /* gPET.h:87-92 */
typedef struct { orc_event a, b; } orc_coinc;
typedef struct { int32_t parn, pann, modn, cryn, type; float E, t32, x, y, z; double t; } orc_hit;
typedef struct { float x, y, z, E, vx, vy, vz; int32_t nscat; double t; int32_t eventid, parn; } orc_photon;

typedef struct {  /* same field order as gpet_panel / object_t (gPET.h:54-78) */
    int32_t panel;
    float lengthx, lengthy, lengthz, MODx, MODy, MODz, Mspacex, Mspacey, Mspacez, LSOx, LSOy, LSOz;
    float spacex, spacey, spacez, offsetx, offsety, offsetz, directionx, directiony, directionz;
    float UniXx, UniXy, UniXz, UniYx, UniYy, UniYz, UniZx, UniZy, UniZz;
} orc_panel;

typedef struct {
    int32_t nmat, nen;
    float e0, e1;                 /* first / last energy of the 1-D grid */
    const float *lamph, *compt, *rayle;  /* [mat][ie] */
    const float *maj;             /* Sigma_max on the grid (1/cm) */
    int32_t cm_ncp, cm_ne, rl_ncp, rl_ne;
    float cm_dcp, cm_de, rl_dcp, rl_de;
    const float *cmpsf, *rayff;   /* [mat][icp][ie] */
} orc_tables;

typedef struct {
    int32_t readout_depth, readout_policy;
    float threshold_eV;
    int32_t blur_policy; float blur_Eref, blur_Rref, blur_slope, blur_space;
    int32_t dead_level, dead_type; float dead_time_us;
    float ewin_min, ewin_max;
    float time_blur_sigma_us, coinc_window_us; int32_t coinc_policy, coinc_min_panel_diff;
    int32_t npanels, moduleN, crystalN;
    uint64_t seed;
    /* equal times: 0 input order; 1 site number first, then input order -- the rule gpet_run applies to the events of its
     * own detector kernel, whose order in the buffer is not defined (DigitizerDev::tie_site) */
    int32_t tie_site;
} orc_digi_params;

/* ------------------------------------------------------------------------------------------------ table lookups */
/* itphip_G / icptip / irylip / lamwck (gPET_kernals.cu:31-64): linear interpolation at index idl*(e - e0) */
static void energy_index(const orc_tables* tb, float E, int* i, float* f) {
    float ide = (float)(tb->nen - 1) / (tb->e1 - tb->e0);  /* initialize.cu:421 */
    float x = ide * (E - tb->e0);
    if (x < 0.f) x = 0.f;
    if (x > (float)(tb->nen - 1)) x = (float)(tb->nen - 1);
    int k = (int)x;
    if (k > tb->nen - 2) k = tb->nen - 2;
    *i = k; *f = x - (float)k;
}
static float lerp1(const float* t, int i, float f) { return fmaf(f, t[i + 1] - t[i], t[i]); }

/* comsam(table) / rylsam surface read (gPET_kernals.cu:80-85, 140-145): bilinear, border clamp, result in [-1,1] */
static float surface_lookup(const float* surf, int mat, int ncp, int ne, float xe, float xcp) {
    if (xe < 0.f) xe = 0.f;
    if (xe > (float)(ne - 1)) xe = (float)(ne - 1);
    if (xcp < 0.f) xcp = 0.f;
    if (xcp > (float)(ncp - 1)) xcp = (float)(ncp - 1);
    int ie = (int)xe, ic = (int)xcp;
    if (ie > ne - 2) ie = ne - 2;
    if (ic > ncp - 2) ic = ncp - 2;
    float fe = xe - (float)ie, fc = xcp - (float)ic;
    const float* p = surf + ((size_t)mat * ncp + ic) * ne + ie;
    float a = fmaf(fe, p[1] - p[0], p[0]), b = fmaf(fe, p[ne + 1] - p[ne], p[ne]);
    float c = fmaf(fc, b - a, a);
    if (c > 1.f) c = 1.f;
    if (c < -1.f) c = -1.f;
    return c;
}

/* majorant on the table grid: max_m lamph[m][i] * maxdens[m]  (role of iniwck, initialize.cu:773-829, 919-966) */
void orc_build_majorant(int nmat, int nen, const float* lamph, const float* maxdens, float* out) {
    for (int i = 0; i < nen; i++) {
        float y = 0.f;
        for (int m = 0; m < nmat; m++) {
            float v = lamph[(size_t)m * nen + i] * maxdens[m];
            if (v > y) y = v;
        }
        out[i] = y;
    }
}

/* rotate (gPET_kernals.cu:172-254) */
static void rotate_dir(float* u, float* v, float* w, float costh, float phi) {
    float rho2 = (*u) * (*u) + (*v) * (*v);
    float norm = rho2 + (*w) * (*w);
    if (fabsf(norm - 1.0f) > 1.0e-4f) {
        norm = 1.0f / sqrtf(norm);
        *u *= norm; *v *= norm; *w *= norm;
    }
    float sinphi = sinf(phi), cosphi = cosf(phi);
    float c2 = costh * costh;
    if (rho2 > 1.0e-20f) {
        float sthrho = c2 < 1.0f ? sqrtf((1.0f - c2) / rho2) : 0.0f;
        float urho = (*u) * sthrho, vrho = (*v) * sthrho;
        float un = (*u) * costh - vrho * sinphi + (*w) * urho * cosphi;
        float vn = (*v) * costh + urho * sinphi + (*w) * vrho * cosphi;
        float wn = (*w) * costh - rho2 * sthrho * cosphi;
        *u = un; *v = vn; *w = wn;
    } else {
        float sinth = c2 < 1.0f ? sqrtf(1.0f - c2) : 0.0f;
        *v = sinth * sinphi;
        if (*w > 0.0f) { *u = sinth * cosphi; *w = costh; }
        else { *u = -sinth * cosphi; *w = -costh; }
    }
}

/* ------------------------------------------------------------------------------------------------ positron (S4, S5) */
typedef struct {  /* voxel grid as the range walk sees it */
    const float* dens; int nx, ny, nz; float ox, oy, oz, dx, dy, dz, idx, idy, idz;
} orc_grid;

/* sampleEkPositron (gPET_kernals.cu:420-443) */
static float sample_ek_positron(const float* coef, orc_rng* g) {
    float E, u, sumE;
    do {
        uint32_t r[4];
        rng_next(g, r);
        E = (float)((double)u01(r[0]) * ((double)coef[0] - 0.511) + 0.511);
        u = coef[1] * u01(r[1]);
        sumE = 0.f;
        for (int i = 0; i < 6; i++) sumE += coef[2 + i] * powf(E, (float)(5 - i));
    } while (u > sumE);
    return (float)(((double)E - 0.511) * 1e6);
}

static float grid_density_clamped(const orc_grid* gr, int ix, int iy, int iz) {  /* clamp-addressed point texture */
    if (ix < 0) ix = 0;
    if (ix > gr->nx - 1) ix = gr->nx - 1;
    if (iy < 0) iy = 0;
    if (iy > gr->ny - 1) iy = gr->ny - 1;
    if (iz < 0) iz = 0;
    if (iz > gr->nz - 1) iz = gr->nz - 1;
    return gr->dens[((size_t)iz * gr->ny + iy) * gr->nx + ix];
}

/* setPositronRange (gPET_kernals.cu:347-418), statement by statement (quirks kept: density of the voxel being
 * entered; `step` keeps its 1000 sentinel for a positron outside the phantom) */
static void positron_range(const orc_grid* gr, float* px, float* py, float* pz, float vx, float vy, float vz, float ekin_eV,
                           int usedirection, orc_rng* g) {
    const float ekin = (float)((double)ekin_eV / 1e6);
    float b1 = 5.44040782f, b2 = 0.369516529f;
    const float Rex = (float)(0.1 * (double)b1 * (double)ekin * (double)ekin / (double)(b2 + ekin));
    const float sigma = Rex / (2 * 1.0f);
    uint32_t q[4];
    rng_next(g, q);
    const float ra = sqrtf(-2.0f * logf(u01(q[0]))), rb = sqrtf(-2.0f * logf(u01(q[2])));
    float dx = sigma * (ra * cosf(ORC_TWOPI * u01(q[1])));
    float dy = sigma * (ra * sinf(ORC_TWOPI * u01(q[1])));
    float dz = sigma * (rb * cosf(ORC_TWOPI * u01(q[3])));
    const float r = sqrtf(dx * dx + dy * dy + dz * dz);
    if (usedirection) {
        const float tmp = sqrtf(vx * vx + vy * vy + vz * vz);
        dx = r * vx / tmp; dy = r * vy / tmp; dz = r * vz / tmp;
    }
    float s = 0.f, step;
    int ix = (int)((*px - gr->ox) * gr->idx), iy = (int)((*py - gr->oy) * gr->idy), iz = (int)((*pz - gr->oz) * gr->idz);
    int w = (ix <= 0 || ix >= gr->nx || iy <= 0 || iy >= gr->ny || iz <= 0 || iz >= gr->nz) ? -1 : 1;
    int guard = 0;
    while (s < r && guard++ < 100000) {
        step = 1000.f;
        if (w > 0) {
            b1 = (gr->ox + (ix + (dx > 0.f)) * gr->dx - *px) / dx;
            if (step > b1) { step = b1; w = 1; }
            b1 = (gr->oy + (iy + (dy > 0.f)) * gr->dy - *py) / dy;
            if (step > b1) { step = b1; w = 2; }
            b1 = (gr->oz + (iz + (dz > 0.f)) * gr->dz - *pz) / dz;
            if (step > b1) { step = b1; w = 3; }
            if (w == 1) ix += (dx > 0.f) ? 1 : -1;
            else if (w == 2) iy += (dy > 0.f) ? 1 : -1;
            else iz += (dz > 0.f) ? 1 : -1;
            b2 = grid_density_clamped(gr, ix, iy, iz);
            step = step * r;
            s += step * b2;
            if (s > r) step += (r - s) / b2;
        } else {
            step += (float)((double)(r - s) / 0.0012905);
            s = r + 100.f;
        }
        *px += step * dx / r; *py += step * dy / r; *pz += step * dz / r;
        if (*px < gr->ox || *px > (gr->ox + gr->nx * gr->dx)) w = -1;
        if (*py < gr->oy || *py > (gr->oy + gr->ny * gr->dy)) w = -1;
        if (*pz < gr->oz || *pz > (gr->oz + gr->nz * gr->dz)) w = -1;
    }
}

static void make_grid(orc_grid* gr, const float* dens, const int32_t dim[3], const float offset[3], const float size[3]) {
    gr->dens = dens; gr->nx = dim[0]; gr->ny = dim[1]; gr->nz = dim[2];
    gr->ox = offset[0]; gr->oy = offset[1]; gr->oz = offset[2];
    gr->dx = size[0] / dim[0]; gr->dy = size[1] / dim[1]; gr->dz = size[2] / dim[2];
    gr->idx = 1.0f / gr->dx; gr->idy = 1.0f / gr->dy; gr->idz = 1.0f / gr->dz;
}

/* setPositionForPhoton (gPET_kernals.cu:563-604): positron i -> photons 2i, 2i+1; the second photon's time is set
 * (the reference leaves it 0 and thereby drops the photon: consciously fixed) */
void orc_psf_positron(const orc_photon* pos, int64_t n, uint64_t first, const float* dens, const int32_t dim[3],
                      const float offset[3], const float size[3], float nonangle, int use_prange, uint64_t seed,
                      orc_photon* out) {
    orc_grid gr;
    if (use_prange) make_grid(&gr, dens, dim, offset, size);
    for (int64_t i = 0; i < n; i++) {
        const orc_photon* e = pos + i;
        uint64_t gi = first + (uint64_t)i;
        orc_rng g;
        rng_init(&g, seed, gi, (uint32_t)ST_PSF_POSITRON << 24);
        uint32_t r0[4], r1[4];
        rng_next(&g, r0);
        float x = e->x, y = e->y, z = e->z;
        float ct = -1.f + 2.f * u01(r0[0]);
        float phi = ORC_TWOPI * u01(r0[1]);
        float st = sqrtf(1.f - ct * ct);
        float vx = st * cosf(phi), vy = st * sinf(phi), vz = ct;
        float phi2 = ORC_TWOPI * u01(r0[2]);
        rng_next(&g, r1);
        float gn = sqrtf(-2.f * logf(u01(r1[0]))) * cosf(ORC_TWOPI * u01(r1[1]));
        float delta = gn * nonangle;
        if (use_prange) positron_range(&gr, &x, &y, &z, e->vx, e->vy, e->vz, e->E, 1, &g);
        for (int which = 0; which < 2; which++) {
            orc_photon* p = out + 2 * i + which;
            float ax = vx, ay = vy, az = vz, E;
            if (which == 0) E = ORC_MC2 + delta * ORC_MC2 * 0.5f;
            else { rotate_dir(&ax, &ay, &az, -cosf(delta), phi2); E = ORC_MC2 - delta * ORC_MC2 * 0.5f; }
            p->x = x; p->y = y; p->z = z; p->E = E; p->vx = ax; p->vy = ay; p->vz = az; p->nscat = 0;
            p->t = e->t; p->eventid = (int32_t)((uint32_t)gi & 0x7fffffffu); p->parn = (int32_t)((uint32_t)(2 * gi + which) & 0x7fffffffu);
        }
    }
}

/* ------------------------------------------------------------------------------------------------ source (S2, S3) */
/* One frame: pairs k = 0..npairs-1, source index from the inclusive prefix cum_pairs[]; decay time is the
 * truncated exponential inside [t0, t0+dt) that the per-atom test of gPET_kernals.cu:519-521 induces. */
void orc_source_ex(int nsource, const uint64_t* cum_pairs, const int32_t* shape, const float* coeff, const double* tau_s,
                   const double* frac, double t0_s, uint64_t first_pair, float nonangle, uint64_t npairs, uint64_t seed,
                   int use_prange, const int32_t* type, const float* iso_coef, const float* dens, const int32_t dim[3],
                   const float offset[3], const float size[3], orc_photon* out);

void orc_source(int nsource, const uint64_t* cum_pairs, const int32_t* shape, const float* coeff, const double* tau_s,
                const double* frac, double t0_s, uint64_t first_pair, float nonangle, uint64_t npairs, uint64_t seed,
                orc_photon* out) {
    orc_source_ex(nsource, cum_pairs, shape, coeff, tau_s, frac, t0_s, first_pair, nonangle, npairs, seed, 0, NULL, NULL, NULL,
                  NULL, NULL, NULL, out);
}

/* use_prange = 1 adds S4 + S5 (gPET_kernals.cu:529-533): `type` = isotope row per source, `iso_coef` = 8 floats per row */
void orc_source_ex(int nsource, const uint64_t* cum_pairs, const int32_t* shape, const float* coeff, const double* tau_s,
                   const double* frac, double t0_s, uint64_t first_pair, float nonangle, uint64_t npairs, uint64_t seed,
                   int use_prange, const int32_t* type, const float* iso_coef, const float* dens, const int32_t dim[3],
                   const float offset[3], const float size[3], orc_photon* out) {
    orc_grid gr;
    if (use_prange) make_grid(&gr, dens, dim, offset, size);
    for (uint64_t k = 0; k < npairs; k++) {
        int s = 0;
        while (s < nsource - 1 && k >= cum_pairs[s]) s++;
        uint64_t gk = first_pair + k;
        orc_rng g;
        rng_init(&g, seed, gk, (uint32_t)ST_SOURCE << 24);
        uint32_t r0[4], r1[4], r2[4];
        rng_next(&g, r0); rng_next(&g, r1); rng_next(&g, r2);
        double ptime = -tau_s[s] * log1p(-u01d(r0[0], r0[1]) * frac[s]);
        double t_us = (t0_s + ptime) * 1e6;  /* gPET_kernals.cu:544 */
        const float* c = coeff + 6 * s;
        float u0 = u01(r1[0]), u1 = u01(r1[1]), u2 = u01(r1[2]);
        float x, y, z;
        int sh = shape[s];
        if (sh < 0 || sh > 2) sh = 0;
        if (sh == 0) {
            x = c[0] + c[3] * (-1.f + 2.f * u0) * 0.5f;
            y = c[1] + c[4] * (-1.f + 2.f * u1) * 0.5f;
            z = c[2] + c[5] * (-1.f + 2.f * u2) * 0.5f;
        } else if (sh == 1) {
            float phi = ORC_TWOPI * u0, rr = c[3] * sqrtf(u1);
            x = c[0] + rr * cosf(phi);
            y = c[1] + rr * sinf(phi);
            z = c[2] + c[4] * (-1.f + 2.f * u2) * 0.5f;
        } else {
            float phi = ORC_TWOPI * u0, ct = -1.f + 2.f * u1, rr = c[3] * cbrtf(u2);
            float st = sqrtf(1.f - ct * ct);
            x = c[0] + rr * st * cosf(phi);
            y = c[1] + rr * st * sinf(phi);
            z = c[2] + rr * ct;
        }
        float ct = -1.f + 2.f * u01(r0[2]);
        float phi = ORC_TWOPI * u01(r0[3]);
        float st = sqrtf(1.f - ct * ct);
        float vx = st * cosf(phi), vy = st * sinf(phi), vz = ct;
        float phi2 = ORC_TWOPI * u01(r2[0]);
        float gn = sqrtf(-2.f * logf(u01(r2[1]))) * cosf(ORC_TWOPI * u01(r2[2]));
        float delta = gn * nonangle;
        if (use_prange) {
            float ek = sample_ek_positron(iso_coef + 8 * type[s], &g);
            positron_range(&gr, &x, &y, &z, 0.f, 0.f, 0.f, ek, 0, &g);
        }
        for (int which = 0; which < 2; which++) {
            orc_photon* p = out + 2 * k + which;
            float ax = vx, ay = vy, az = vz, E;
            if (which == 0) E = ORC_MC2 + delta * ORC_MC2 * 0.5f;
            else { rotate_dir(&ax, &ay, &az, -cosf(delta), phi2); E = ORC_MC2 - delta * ORC_MC2 * 0.5f; }
            p->x = x; p->y = y; p->z = z; p->E = E; p->vx = ax; p->vy = ay; p->vz = az; p->nscat = 0;
            p->t = t_us; p->eventid = (int32_t)((uint32_t)gk & 0x7fffffffu); p->parn = (int32_t)((uint32_t)(2 * gk + which) & 0x7fffffffu);
        }
    }
}

/* ------------------------------------------------------------------------------------------------ phantom (P1) */
/* getDistance (gPET_kernals.cu:148-171): distance along the direction to the PSF-recording sphere rec = (x, y, z, r) */
static float record_distance(const float rec[4], const orc_photon* p) {
    float cx = p->x - rec[0], cy = p->y - rec[1], cz = p->z - rec[2];
    float a = p->vx * p->vx + p->vy * p->vy + p->vz * p->vz;
    float b = 2.0f * (p->vx * cx + p->vy * cy + p->vz * cz);
    float c = (cx * cx + cy * cy + cz * cz) - rec[3] * rec[3];
    float disc = b * b - 4 * a * c;
    if (disc < 0) return 0.f;
    if (c < 0) return (-b + sqrtf(disc)) / (2 * a);
    if (b < 0) return (-b - sqrtf(disc)) / (2 * a);
    return (-b + sqrtf(disc)) / (2 * a);
}

/* photon() (gPET_kernals.cu:256-345), in place: t = -0.5 marks a photo-absorbed photon.  rec != NULL switches on the
 * RECORDPSF == -1 branch (:288-294): a photon that leaves the phantom is moved onto the recording sphere. */
void orc_phantom_ex(orc_photon* ph, int64_t n, const int32_t* mat, const float* dens, const int32_t dim[3],
                    const float offset[3], const float size[3], const orc_tables* tb, float eabs, uint64_t seed, const float* rec);
void orc_phantom(orc_photon* ph, int64_t n, const int32_t* mat, const float* dens, const int32_t dim[3],
                 const float offset[3], const float size[3], const orc_tables* tb, float eabs, uint64_t seed) {
    orc_phantom_ex(ph, n, mat, dens, dim, offset, size, tb, eabs, seed, NULL);
}
void orc_phantom_ex(orc_photon* ph, int64_t n, const int32_t* mat, const float* dens, const int32_t dim[3],
                    const float offset[3], const float size[3], const orc_tables* tb, float eabs, uint64_t seed, const float* rec) {
    float idx = 1.0f / (size[0] / dim[0]), idy = 1.0f / (size[1] / dim[1]), idz = 1.0f / (size[2] / dim[2]);
    for (int64_t id = 0; id < n; id++) {
        orc_photon* p = ph + id;
        if (p->E < 0.f || p->t <= 0.0) continue;  /* :272 */
        orc_rng g;
        rng_init(&g, seed, photon_index(p->parn), (uint32_t)ST_PHANTOM << 24);
        for (;;) {
            uint32_t r[4];
            rng_next(&g, r);
            int ie; float fe;
            energy_index(tb, p->E, &ie, &fe);
            float lammin = 1.0f / lerp1(tb->maj, ie, fe);
            float s = -lammin * logf(u01(r[0]));
            p->x = fmaf(s, p->vx, p->x); p->y = fmaf(s, p->vy, p->y); p->z = fmaf(s, p->vz, p->z);
            p->t += (double)s / ORC_SPE;
            /* getAbsVox (:19-29): truncation, voxel layer 0 counts as outside */
            int ix = (int)((p->x - offset[0]) * idx), iy = (int)((p->y - offset[1]) * idy), iz = (int)((p->z - offset[2]) * idz);
            if (ix <= 0 || ix >= dim[0] || iy <= 0 || iy >= dim[1] || iz <= 0 || iz >= dim[2]) {
                if (rec) {
                    float rr = record_distance(rec, p);
                    p->x = fmaf(rr, p->vx, p->x); p->y = fmaf(rr, p->vy, p->y); p->z = fmaf(rr, p->vz, p->z);
                    p->t += (double)rr / ORC_SPE;
                }
                break;
            }
            size_t v = ((size_t)iz * dim[1] + iy) * dim[0] + ix;
            float rho = dens[v];
            int m = mat[v];
            float lamden = lammin * rho;
            float tot = lerp1(tb->lamph + (size_t)m * tb->nen, ie, fe);
            float prob = 1.0f - lamden * tot;
            float u = u01(r[1]);
            if (u < prob) continue;
            prob += lamden * lerp1(tb->compt + (size_t)m * tb->nen, ie, fe);
            if (u < prob) {
                float costh = surface_lookup(tb->cmpsf, m, tb->cm_ncp, tb->cm_ne, p->E * (1.0f / tb->cm_de), u01(r[2]) * (1.0f / tb->cm_dcp));
                float efrac = 1.0f / (1.0f + p->E * ORC_IMC2 * (1.0f - costh));
                float phi = ORC_TWOPI * u01(r[3]);
                p->E *= efrac;
                p->nscat++;
                if (p->E < eabs) break;  /* :319-320: still alive, handed to the detector */
                rotate_dir(&p->vx, &p->vy, &p->vz, costh, phi);
                continue;
            }
            prob += lamden * lerp1(tb->rayle + (size_t)m * tb->nen, ie, fe);
            if (u < prob) {
                float costh = surface_lookup(tb->rayff, m, tb->rl_ncp, tb->rl_ne, p->E * (1.0f / tb->rl_de), u01(r[2]) * (1.0f / tb->rl_dcp));
                float phi = ORC_TWOPI * u01(r[3]);
                p->nscat++;
                rotate_dir(&p->vx, &p->vy, &p->vz, costh, phi);
                continue;
            }
            p->t = -0.5;  /* photoelectric (:336) */
            break;
        }
    }
}

/* ------------------------------------------------------------------------------------------------ detector (X1, X2, D1, D2) */
/* crystalSearch (:1236-1279) */
static void crystal_search(const orc_panel* pd, int moduleNy, int crystalNy, int nsurface, const float* surface,
                           float px, float py, float pz, int* m_id, int* M_id, int* L_id) {
    *m_id = 1; *M_id = -1; *L_id = -1;
    for (int k = 0; k < nsurface; k++) {
        const float* c = surface + 10 * k;
        float q = c[0] * px * px + c[1] * py * py + c[2] * pz * pz + c[3] * px * py + c[4] * px * pz + c[5] * py * pz +
                  c[6] * px + c[7] * py + c[8] * pz + c[9];
        if (q < 0.f) return;
    }
    float y = pd->lengthy / 2 + py, z = pd->lengthz / 2 + pz;
    float my = y / (pd->MODy + pd->Mspacey), mz = z / (pd->MODz + pd->Mspacez);
    int My = floorf(my) > 0.f ? (int)my : 0, Mz = floorf(mz) > 0.f ? (int)mz : 0;
    *M_id = Mz * moduleNy + My;
    y = y - My * (pd->MODy + pd->Mspacey);
    z = z - Mz * (pd->MODz + pd->Mspacez);
    if (y > pd->MODy || z > pd->MODz) return;
    float cy = y / (pd->LSOy + pd->spacey), cz = z / (pd->LSOz + pd->spacez);
    int Ly = floorf(cy) > 0.f ? (int)cy : 0, Lz = floorf(cz) > 0.f ? (int)cz : 0;
    *L_id = Lz * crystalNy + Ly;
    y = y - Ly * (pd->LSOy + pd->spacey);
    z = z - Lz * (pd->LSOz + pd->spacez);
    if (y > pd->LSOy || z > pd->LSOz) return;
    *m_id = 0;
}

/* comsam, free-electron Klein-Nishina (:90-126) */
static void compton_kn(float E, orc_rng* g, float* efrac, float* costh) {
    float e0 = E * ORC_IMC2, twoe = 2.0f * e0;
    float kmin2 = 1.0f / ((1.0f + twoe) * (1.0f + twoe));
    float loge = logf(1.0f + twoe);
    for (;;) {
        uint32_t r[4];
        rng_next(g, r);
        if (u01(r[0]) * (loge + twoe * (1.0f + e0) * kmin2) < loge) *efrac = expf(-u01(r[1]) * loge);
        else *efrac = sqrtf(kmin2 + u01(r[1]) * (1.0f - kmin2));
        float mess = e0 * e0 * (*efrac) * (1.0f + (*efrac) * (*efrac));
        if (u01(r[2]) * mess <= mess - (1.0f - *efrac) * ((1.0f + twoe) * (*efrac) - 1.0f)) break;
    }
    *costh = 1.0f - (1.0f - *efrac) / ((*efrac) * e0);
}

#define ORC_MAXEV 6  /* the reference's Event events[4] has no bound check; 6 slots + overflow count */

/* adder (:737-755); the centroid contraction is spelled fma(x_i, E_i, x*E)/(E_i+E) */
static int adder(orc_event* ev, int* cnt, const orc_event* e) {
    for (int i = 0; i < *cnt; i++) {
        if (e->siten == ev[i].siten) {
            float es = ev[i].E + e->E;
            ev[i].x = fmaf(ev[i].x, ev[i].E, e->x * e->E) / es;
            ev[i].y = fmaf(ev[i].y, ev[i].E, e->y * e->E) / es;
            ev[i].z = fmaf(ev[i].z, ev[i].E, e->z * e->E) / es;
            ev[i].E = es;
            return 1;
        }
    }
    if (*cnt >= ORC_MAXEV) return 0;
    ev[(*cnt)++] = *e;
    return 1;
}

/* readout (:756-813); depth 3 returns the adder list unchanged (the reference leaves counts[1]=4 there: fixed) */
static int readout(orc_event* ev, int cnt, int depth, int policy, int moduleN) {
    if (depth == 3) return cnt;
    if (policy == 1) depth = 2;
    for (int i = 0; i < cnt; i++) {
        if (depth == 0) ev[i].siten = 0;
        else if (depth == 1) ev[i].siten = ev[i].pann;
        else ev[i].siten = ev[i].pann * moduleN + ev[i].modn;
    }
    int ind = 0;
    for (int i = 0; i < cnt; i++) {
        orc_event e0 = ev[i];
        if (e0.t > ORC_MAXT * 0.1) continue;
        for (int j = i + 1; j < cnt; j++) {
            orc_event e = ev[j];
            if (e.t > ORC_MAXT * 0.1) continue;  /* already merged (the reference re-merges dead slots only into dead ones) */
            if (e.parn == e0.parn && e.siten == e0.siten) {
                if (policy == 1) {
                    float es = e0.E + e.E;
                    e0.x = fmaf(e0.x, e0.E, e.x * e.E) / es;
                    e0.y = fmaf(e0.y, e0.E, e.y * e.E) / es;
                    e0.z = fmaf(e0.z, e0.E, e.z * e.E) / es;
                    e0.E = es;
                    ev[j].t = ORC_MAXT;
                    continue;
                }
                e0 = (e0.E > e.E) ? e0 : e;
                ev[j].t = ORC_MAXT;
            }
        }
        ev[ind++] = e0;
    }
    return ind;
}

/* photonde (:839-1233).  Returns the number of photons that entered a panel; *nhits / *nevents receive the counts.
 * Hits and events are appended in photon order. */
int64_t orc_detector(const orc_photon* ph, int64_t n, const orc_panel* panels, int npanels, const int32_t counts4[4],
                     const int32_t pmat[2], const float pdens[2], int nsurface, const float* surface,
                     const orc_tables* tb, float eabs, int rdepth, int rpolicy, uint64_t seed, orc_hit* hits,
                     int64_t hit_cap, int64_t* nhits, orc_event* events, int64_t ev_cap, int64_t* nevents,
                     int64_t* adder_overflow) {
    const int moduleNy = counts4[0], crystalNy = counts4[1], moduleN = counts4[2], crystalN = counts4[3];
    int64_t nh = 0, ne = 0, entered = 0, ovf = 0;
    for (int64_t id = 0; id < n; id++) {
        const orc_photon* p = ph + id;
        if (!(p->t > 0.0)) continue;  /* :951 */
        int pa = -1;
        float x = 0, y = 0, z = 0, vx = 0, vy = 0, vz = 0, E = p->E;
        double tof = p->t;
        for (int i = 0; i < npanels; i++) {  /* :963-1009 */
            const orc_panel* pd = panels + i;
            float rx = p->x - pd->offsetx, ry = p->y - pd->offsety, rz = p->z - pd->offsetz;
            float lx = rx * pd->UniXx + ry * pd->UniXy + rz * pd->UniXz;
            float ly = rx * pd->UniYx + ry * pd->UniYy + rz * pd->UniYz;
            float lz = rx * pd->UniZx + ry * pd->UniZy + rz * pd->UniZz;
            float lvx = p->vx * pd->UniXx + p->vy * pd->UniXy + p->vz * pd->UniXz;
            float lvy = p->vx * pd->UniYx + p->vy * pd->UniYy + p->vz * pd->UniYz;
            float lvz = p->vx * pd->UniZx + p->vy * pd->UniZy + p->vz * pd->UniZz;
            if (lvx * pd->directionx >= 0.f) {
                float q = lx / lvx;
                float y2 = ly - q * lvy, z2 = lz - q * lvz;
                if (fabsf(y2) < pd->lengthy / 2 && fabsf(z2) < pd->lengthz / 2) {
                    x = 0.f; y = y2; z = z2; vx = lvx; vy = lvy; vz = lvz;
                    tof += -(double)lx / (ORC_SPE * (double)lvx);  /* :987 */
                    pa = i;
                    break;
                }
            }
        }
        if (pa < 0) continue;
        entered++;
        const orc_panel* pd = panels + pa;
        orc_rng g;
        rng_init(&g, seed, photon_index(p->parn), (uint32_t)ST_DETECTOR << 24);
        orc_event evs[ORC_MAXEV];
        int cnt = 0;
        for (;;) {
            uint32_t r[4];
            rng_next(&g, r);
            int ie; float fe;
            energy_index(tb, E, &ie, &fe);
            float lammin = 1.0f / lerp1(tb->maj, ie, fe);
            float s = -lammin * logf(u01(r[0]));
            x = fmaf(s, vx, x); y = fmaf(s, vy, y); z = fmaf(s, vz, z);
            tof += (double)s / ORC_SPE;
            if (fabsf(y) > pd->lengthy * 0.5f || fabsf(z) > pd->lengthz * 0.5f || x * pd->directionx < 0.f ||
                x * pd->directionx > pd->lengthx)
                break;  /* :1027 */
            int m_id, M_id, L_id;
            crystal_search(pd, moduleNy, crystalNy, nsurface, surface, x, y, z, &m_id, &M_id, &L_id);
            float rho = pdens[m_id];
            int m = pmat[m_id];
            float lamden = lammin * rho;
            float prob = 1.0f - lamden * lerp1(tb->lamph + (size_t)m * tb->nen, ie, fe);
            if (prob < 0.f) prob = 0.f;  /* :1044 */
            float u = u01(r[1]);
            if (u < prob) continue;
            orc_event e;
            e.parn = p->parn; e.pann = pd->panel; e.modn = M_id; e.cryn = L_id;
            e.siten = pd->panel * moduleN * crystalN + M_id * crystalN + L_id;  /* :1073 */
            e.eventid = p->eventid; e.t = tof; e.x = x; e.y = y; e.z = z; e.E = 0.f;
            int nnew = 0, type0 = 0;
            float E0 = 0.f, E1 = 0.f;
            int stop = 0;
            prob += lamden * lerp1(tb->compt + (size_t)m * tb->nen, ie, fe);
            if (u < prob) {
                float efrac, costh;
                compton_kn(E, &g, &efrac, &costh);
                float de = E * (1.0f - efrac);
                float phi = ORC_TWOPI * u01(r[2]);
                if (m_id == 0) { type0 = 1; E0 = de; nnew = 1; }
                E -= de;
                if (E < eabs) {
                    if (m_id == 0) { E1 = E; nnew = 2; }
                    stop = 1;
                } else {
                    rotate_dir(&vx, &vy, &vz, costh, phi);
                }
            } else {
                prob += lamden * lerp1(tb->rayle + (size_t)m * tb->nen, ie, fe);
                if (u < prob) {
                    float costh = surface_lookup(tb->rayff, m, tb->rl_ncp, tb->rl_ne, E * (1.0f / tb->rl_de), u01(r[2]) * (1.0f / tb->rl_dcp));
                    float phi = ORC_TWOPI * u01(r[3]);
                    rotate_dir(&vx, &vy, &vz, costh, phi);
                } else {
                    if (m_id == 0) { type0 = 4; E0 = E; nnew = 1; }
                    stop = 1;
                }
            }
            for (int k = 0; k < nnew; k++) {
                if (nh < hit_cap) {
                    orc_hit* h = hits + nh;
                    h->parn = p->parn; h->pann = pd->panel; h->modn = M_id; h->cryn = L_id; h->type = k ? 2 : type0;
                    h->E = k ? E1 : E0; h->t32 = (float)tof; h->x = x; h->y = y; h->z = z; h->t = tof;
                }
                nh++;
                e.E = k ? E1 : E0;
                if (!adder(evs, &cnt, &e)) ovf++;
            }
            if (stop) break;
        }
        if (cnt) {
            int nout = readout(evs, cnt, rdepth, rpolicy, moduleN);
            for (int k = 0; k < nout; k++) {
                if (ne < ev_cap) events[ne] = evs[k];
                ne++;
            }
        }
    }
    *nhits = nh; *nevents = ne;
    if (adder_overflow) *adder_overflow = ovf;
    return entered;
}

/* ------------------------------------------------------------------------------------------------ noise singles */
/* addnoise (gPET_kernals.cu:699-735): thread `id` walks a Poisson process of mean gap `lambda` through its time slice
 * [id, id+1) * interval; every arrival is an event with E = Emean + sigma * N(0,1), position numbers uniform in (0,1],
 * uniformly drawn panel / module / crystal, parn = -1, crystal-level siten.  The reference never launches it.  Fixed
 * here on purpose (and documented in DESIGN.md): fp64 time accumulation (the reference's fp32 `t` stalls once its ulp
 * exceeds the gaps), int(N * u) clamped to N - 1, eventid = 0x80000000 | (slice << 10 | ordinal), arrivals kept only
 * inside [t_lo, t_hi).  Returns the number of events (written while < cap), in slice order. */
int64_t orc_noise(double t_lo, double t_hi, float lambda_us, float Emean, float sigma, float interval_us, int32_t npanels,
                  int32_t moduleN, int32_t crystalN, uint64_t seed, orc_event* out, int64_t cap) {
    if (!(lambda_us > 0.f) || !(interval_us > 0.f) || !(t_hi > t_lo)) return 0;
    double iv = (double)interval_us, lam = (double)lambda_us;
    int64_t id0 = (int64_t)floor(t_lo / iv), id1 = (int64_t)ceil(t_hi / iv), n = 0;
    if (id1 <= id0) id1 = id0 + 1;
    for (int64_t id = id0; id < id1; id++) {
        double t = (double)id * iv, tend = (double)(id + 1) * iv;
        orc_rng g;
        rng_init(&g, seed, (uint64_t)id, (uint32_t)ST_NOISE << 24);
        for (uint32_t ord = 0; ord < (1u << 20); ord++) {
            uint32_t r[4], q[4], w[4];
            rng_next(&g, r);
            t = t + (-log((double)u01(r[0]))) * lam;
            if (!(t < tend)) break;
            rng_next(&g, q);
            rng_next(&g, w);
            if (t < t_lo || !(t < t_hi)) continue;
            orc_event e;
            double gs = sqrt(-2.0 * log((double)u01(r[1]))) * cos(6.283185307179586 * (double)u01(r[2]));
            e.E = (float)((double)Emean + (double)sigma * gs);
            e.x = u01(r[3]); e.y = u01(q[0]); e.z = u01(q[1]);
            e.parn = -1;
            e.pann = (int32_t)((float)npanels * u01(q[2])); if (e.pann > npanels - 1) e.pann = npanels - 1;
            e.modn = (int32_t)((float)moduleN * u01(q[3])); if (e.modn > moduleN - 1) e.modn = moduleN - 1;
            e.cryn = (int32_t)((float)crystalN * u01(w[0])); if (e.cryn > crystalN - 1) e.cryn = crystalN - 1;
            e.siten = e.pann * moduleN * crystalN + e.modn * crystalN + e.cryn;
            e.eventid = (int32_t)(0x80000000u | ((((uint32_t)id << 10) | (ord & 1023u)) & 0x7fffffffu));
            e.t = t;
            if (n < cap) out[n] = e;
            n++;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------------ digitizer (D3-D7) */
static void merge_sort_idx(int64_t* idx, int64_t* tmp, int64_t n, int (*less)(int64_t, int64_t, const void*), const void* ctx) {
    /* bottom-up stable merge sort of an index array */
    for (int64_t w = 1; w < n; w *= 2) {
        for (int64_t lo = 0; lo < n; lo += 2 * w) {
            int64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int64_t a = lo, b = mid, k = lo;
            while (a < mid && b < hi) tmp[k++] = less(idx[b], idx[a], ctx) ? idx[b++] : idx[a++];
            while (a < mid) tmp[k++] = idx[a++];
            while (b < hi) tmp[k++] = idx[b++];
        }
        memcpy(idx, tmp, (size_t)n * sizeof(int64_t));
    }
}
typedef struct { const orc_event* e; const int64_t* orig; int tie_site; } sort_ctx;
/* time order; ties are broken by the position in the input list (std::sort in the reference leaves ties undefined) */
static int less_t(int64_t a, int64_t b, const void* ctx) {
    const sort_ctx* c = (const sort_ctx*)ctx;
    if (c->e[a].t != c->e[b].t) return c->e[a].t < c->e[b].t;
    if (c->tie_site && c->e[a].siten != c->e[b].siten) return c->e[a].siten < c->e[b].siten;
    return c->orig[a] < c->orig[b];
}
static int less_site(int64_t a, int64_t b, const void* ctx) { const sort_ctx* c = (const sort_ctx*)ctx; return c->e[a].siten < c->e[b].siten; }

/* quicksort_h(by t) (detector.cu:354-367) made deterministic; works on the first n records in place.  `orig` carries
 * each record's position in the input list and is permuted along. */
static void sort_events(orc_event* ev, int64_t* orig, int64_t n, int by_site, int tie_site) {
    if (n < 2) return;
    int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * (size_t)n * 2);
    orc_event* cp = (orc_event*)malloc(sizeof(orc_event) * (size_t)n);
    int64_t* co = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    sort_ctx c = {ev, orig, tie_site};
    for (int64_t i = 0; i < n; i++) idx[i] = i;
    merge_sort_idx(idx, idx + n, n, by_site ? less_site : less_t, &c);
    for (int64_t i = 0; i < n; i++) { cp[i] = ev[idx[i]]; co[i] = orig[idx[i]]; }
    memcpy(ev, cp, sizeof(orc_event) * (size_t)n);
    memcpy(orig, co, sizeof(int64_t) * (size_t)n);
    free(cp); free(co); free(idx);
}

static int pair_ok(const orc_event* a, const orc_event* b, const orc_digi_params* p) {
    if (p->coinc_min_panel_diff <= 0) return 1;
    int d = abs(a->pann - b->pann);
    if (p->npanels > 0 && p->npanels - d < d) d = p->npanels - d;
    return d >= p->coinc_min_panel_diff;
}

/* Whole chain of gPET.cu:385-424 on a host list.  `work` is scratch of n records.  Returns the number of singles. */
int64_t orc_digitize(const orc_event* in, int64_t n, const orc_digi_params* p, orc_event* work, orc_event* singles,
                     uint64_t counts[4], orc_coinc* coinc, int64_t coinc_cap, int64_t* ncoinc) {
    memcpy(work, in, sizeof(orc_event) * (size_t)n);
    int64_t* orig = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) orig[i] = i;
    int64_t cnt = n;
    counts[0] = (uint64_t)n;
    /* blur (gPET_kernals.cu:814-837) */
    for (int64_t i = 0; i < cnt; i++) {
        orc_event* e = work + i;
        float R = 0.f;
        if (p->blur_policy == 0) R = sqrtf(p->blur_Eref / e->E) * p->blur_Rref;
        if (p->blur_policy == 1) R = (float)((double)p->blur_Rref + (double)(p->blur_slope * (e->E - p->blur_Eref)) / 1e6);
        if (!(R > 0.f)) R = 0.f;
        if (R > 0.f || p->blur_space > 0.f || p->time_blur_sigma_us > 0.f) {
            orc_rng g;
            /* stream = the photon; noise events (parn == -1, orc_noise) are told apart by their event id */
            uint64_t who = e->parn == -1 ? ((1ull << 63) | (uint32_t)e->eventid) : photon_index(e->parn);
            rng_init(&g, p->seed, who, ((uint32_t)ST_BLUR << 24) | ((uint32_t)e->siten & 0xFFFFFFu));
            uint32_t r[4];
            rng_next(&g, r);
            float rad = sqrtf(-2.0f * logf(u01(r[0])));
            float g0 = rad * cosf(ORC_TWOPI * u01(r[1]));
            float nre = (g0 * R) * e->E;
            e->E = (float)((double)e->E + (double)nre / 2.35482);
            if (p->blur_space > 0.f) {
                uint32_t q[4];
                rng_next(&g, q);
                float ra = sqrtf(-2.0f * logf(u01(q[0]))), rb = sqrtf(-2.0f * logf(u01(q[2])));
                float a0 = ORC_TWOPI * u01(q[1]), a1 = ORC_TWOPI * u01(q[3]);
                e->x = e->x + p->blur_space * (ra * cosf(a0));
                e->y = e->y + p->blur_space * (ra * sinf(a0));
                e->z = e->z + p->blur_space * (rb * cosf(a1));
            }
            if (p->time_blur_sigma_us > 0.f) {
                float g1 = rad * sinf(ORC_TWOPI * u01(r[1]));
                double tb = e->t + (double)p->time_blur_sigma_us * (double)g1;
                if (tb > 0.0) e->t = tb;
            }
        }
    }
    /* energywindow(Eth, 2000000) + sort by t (gPET.cu:393-397); records that arrive dead stay dead */
    {
        int64_t alive = 0;
        for (int64_t i = 0; i < cnt; i++) {
            if (work[i].E < p->threshold_eV || work[i].E > 2000000.0f || !(work[i].t < ORC_MAXT * 0.1)) work[i].t = ORC_MAXT;
            else alive++;
        }
        sort_events(work, orig, cnt, 0, 0);   /* ties inside one site stay in input order: the site numbers of the tie rule are set below */
        cnt = alive;
    }
    counts[1] = (uint64_t)cnt;
    /* setSitenum (gPET_kernals.cu:607-640) when dlevel != 3, then orderevents (detector.cu:369-385) */
    if (p->dead_level != 3) {
        for (int64_t i = 0; i < cnt; i++) {
            orc_event* e = work + i;
            if (p->dead_level == 0) e->siten = 0;
            else if (p->dead_level == 1) e->siten = e->pann;
            else if (p->dead_level == 2) e->siten = e->pann * p->moduleN + e->modn;
        }
    }
    sort_events(work, orig, cnt, 1, 0);  /* stable by site on a time-sorted list == sort by site, then by t inside each site */
    /* deadtime (gPET_kernals.cu:657-698), snapshot-start semantics: tdead is float, tdead+tau an fp32 sum */
    {
        const float tau = p->dead_time_us;
        char* kill = (char*)calloc((size_t)(cnt > 0 ? cnt : 1), 1);
        float anchor = 0.f;
        for (int64_t i = 0; i < cnt; i++) {
            int same = i > 0 && work[i].siten == work[i - 1].siten;
            if (p->dead_type == 0) {
                if (same && work[i].t < (double)((float)work[i - 1].t + tau)) kill[i] = 1;
            } else {
                if (!same) { anchor = (float)work[i].t; continue; }
                if (work[i].t < (double)(anchor + tau)) kill[i] = 1;
                else anchor = (float)work[i].t;
            }
        }
        int64_t alive = 0;
        for (int64_t i = 0; i < cnt; i++) {
            if (kill[i]) work[i].t = ORC_MAXT; else alive++;
        }
        free(kill);
        sort_events(work, orig, cnt, 0, p->tie_site);
        cnt = alive;
    }
    counts[2] = (uint64_t)cnt;
    /* energywindow(Ewinmin, Ewinmax) + sort (gPET.cu:418-422) */
    {
        int64_t alive = 0;
        for (int64_t i = 0; i < cnt; i++) {
            if (work[i].E < p->ewin_min || work[i].E > p->ewin_max) work[i].t = ORC_MAXT; else alive++;
        }
        sort_events(work, orig, cnt, 0, p->tie_site);
        cnt = alive;
    }
    counts[3] = (uint64_t)cnt;
    free(orig);
    memcpy(singles, work, sizeof(orc_event) * (size_t)cnt);
    /* coincidence sorter (extension; no reference counterpart): windows open sequentially */
    int64_t nc = 0;
    if (p->coinc_window_us > 0.f && ncoinc) {
        const double W = (double)p->coinc_window_us;
        int64_t a = 0;
        while (a < cnt) {
            const double tend = singles[a].t + W;
            int64_t m = 0, valid = 0;
            while (a + 1 + m < cnt && singles[a + 1 + m].t < tend) {
                if (pair_ok(&singles[a], &singles[a + 1 + m], p)) valid++;
                m++;
            }
            int emit = p->coinc_policy == 0 ? (m == 1 && valid == 1) : (valid > 0);
            if (emit) {
                for (int64_t b = a + 1; b <= a + m; b++) {
                    if (!pair_ok(&singles[a], &singles[b], p)) continue;
                    if (coinc && nc < coinc_cap) { coinc[nc].a = singles[a]; coinc[nc].b = singles[b]; }
                    nc++;
                }
            }
            a += m + 1;
        }
    }
    if (ncoinc) *ncoinc = nc;
    return cnt;
}


/* Coincidence classes (extension; the reference has neither a coincidence sorter, SURVEY F2, nor a scatter flag, F11 --
 * parity for this function is against this specification only).  One byte per coincidence record:
 *   2 random  : the singles stem from different annihilations (eventid >> pair_shift differ) or one is a noise single
 *               (parn == -1, addnoise gPET_kernals.cu:699-735 has no particle)
 *   1 scatter : same annihilation and at least one of the two photons is in `scattered` (photon numbers with a Compton
 *               or Rayleigh interaction in the phantom, the nscat > 0 photons of orc_phantom; gPET_kernals.cu:304-334)
 *   0 true    : the rest
 * totals[0..2] = trues, scatters, randoms. */
static int cmp_i32(const void* a, const void* b) {
    const int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return x < y ? -1 : x > y;
}
void orc_classify(const orc_coinc* co, int64_t n, const int32_t* scattered, int64_t nscat, int32_t pair_shift, uint8_t* cls,
                  uint64_t totals[3]) {
    int32_t* sorted = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nscat > 0 ? nscat : 1));
    if (nscat > 0) memcpy(sorted, scattered, sizeof(int32_t) * (size_t)nscat);
    qsort(sorted, (size_t)(nscat > 0 ? nscat : 0), sizeof(int32_t), cmp_i32);
    totals[0] = totals[1] = totals[2] = 0;
    for (int64_t i = 0; i < n; i++) {
        const orc_event *a = &co[i].a, *b = &co[i].b;
        uint8_t c;
        if (a->parn == -1 || b->parn == -1 || (a->eventid >> pair_shift) != (b->eventid >> pair_shift)) c = 2;
        else if (nscat > 0 && (bsearch(&a->parn, sorted, (size_t)nscat, sizeof(int32_t), cmp_i32) ||
                               bsearch(&b->parn, sorted, (size_t)nscat, sizeof(int32_t), cmp_i32))) c = 1;
        else c = 0;
        if (cls) cls[i] = c;
        totals[c]++;
    }
    free(sorted);
}
