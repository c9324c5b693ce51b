// Source sampling, phantom transport and detector transport for sm_100a.
//
// Reference kernels: setPosition (gPET_kernals.cu:483-561), photon (:256-345), photonde (:839-1233) with
// crystalSearch (:1236-1279), adder/readout (:737-813).  What changes here (B200-first):
//   * counter-based Philox streams keyed by the global photon id: no RNG state array, results independent of the
//     launch shape and of the number of GPUs;
//   * persistent warps with lane refill: a lane whose photon is finished pulls the next one, so the Woodcock loop
//     body always runs with (nearly) full warps instead of waiting for the slowest history of a block;
//   * stage boundaries are compact SoA queues filled with warp-aggregated appends (one atomic per warp);
//   * the per-photon adder/readout runs in registers (no local-memory Event[4]); hits leave the SM as contiguous
//     rows already in the HitsID.dat / Hits.dat layout.
#include <algorithm>
#include <cstdlib>

#include "kernels.hpp"
#include "ktimer.hpp"
#include "philox.cuh"

#include "../../include/gpet_b200.h"

namespace gpet {

namespace {

constexpr int kThreads = 256;
constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------- shared device helpers
struct Xs3 { float tot, compt, rayl; };

__device__ __forceinline__ void energy_index(const TablesDev& tb, float E, int& i, float& f) {
    float x = tb.ide * (E - tb.e0);
    x = fminf(fmaxf(x, 0.f), (float)(tb.nen - 1));
    i = min((int)x, tb.nen - 2);
    f = x - (float)i;
}

__device__ __forceinline__ float lerp_table(const float* __restrict__ t, int i, float f) {
    float a = __ldg(t + i), b = __ldg(t + i + 1);
    return fmaf(f, b - a, a);
}

__device__ __forceinline__ Xs3 lerp_xs(const TablesDev& tb, int mat, int i, float f) {
    const float4* p = tb.xs + (size_t)mat * tb.nen + i;
    float4 a = __ldg(p), b = __ldg(p + 1);
    Xs3 r;
    r.tot = fmaf(f, b.x - a.x, a.x);
    r.compt = fmaf(f, b.y - a.y, a.y);
    r.rayl = fmaf(f, b.z - a.z, a.z);
    return r;
}

// bilinear read of an inverse-CDF surface [mat][icp][ie] at (ie = E*ide, icp = u*idcp); the texture unit of the
// reference clamps coordinates at the borders (tex3D(s_tex, ...), gPET_kernals.cu:80-85, 140-145)
__device__ __forceinline__ float surface_lookup(const float* __restrict__ surf, int mat, int ncp, int ne, float xe, float xcp) {
    xe = fminf(fmaxf(xe, 0.f), (float)(ne - 1));
    xcp = fminf(fmaxf(xcp, 0.f), (float)(ncp - 1));
    int ie = min((int)xe, ne - 2), ic = min((int)xcp, ncp - 2);
    float fe = xe - (float)ie, fc = xcp - (float)ic;
    const float* p = surf + ((size_t)mat * ncp + ic) * ne + ie;
    float v00 = __ldg(p), v01 = __ldg(p + 1), v10 = __ldg(p + ne), v11 = __ldg(p + ne + 1);
    float a = fmaf(fe, v01 - v00, v00), b = fmaf(fe, v11 - v10, v10);
    float c = fmaf(fc, b - a, a);
    return fminf(fmaxf(c, -1.f), 1.f);
}

// PENELOPE-style direction rotation (gPET_kernals.cu:172-254), same fast intrinsics as the reference
__device__ __forceinline__ void rotate_dir(float& u, float& v, float& w, float costh, float phi) {
    float rho2 = u * u + v * v;
    float norm = rho2 + w * w;
    if (fabsf(norm - 1.0f) > 1.0e-4f) {
        norm = 1.0f / __fsqrt_rn(norm);
        u *= norm; v *= norm; w *= norm;
    }
    float sinphi, cosphi;
    __sincosf(phi, &sinphi, &cosphi);
    float c2 = costh * costh;
    if (rho2 > 1.0e-20f) {
        float sthrho = c2 < 1.0f ? __fsqrt_rn((1.0f - c2) / rho2) : 0.0f;
        float urho = u * sthrho, vrho = v * sthrho;
        float un = u * costh - vrho * sinphi + w * urho * cosphi;
        float vn = v * costh + urho * sinphi + w * vrho * cosphi;
        float wn = w * costh - rho2 * sthrho * cosphi;
        u = un; v = vn; w = wn;
    } else {
        float sinth = c2 < 1.0f ? __fsqrt_rn(1.0f - c2) : 0.0f;
        v = sinth * sinphi;
        if (w > 0.0f) { u = sinth * cosphi; w = costh; }
        else { u = -sinth * cosphi; w = -costh; }
    }
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// warp-aggregated reservation of `mine` slots per lane in a global counter; returns this lane's first slot
__device__ __forceinline__ unsigned warp_reserve(unsigned* counter, unsigned mine) {
    unsigned incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned y = __shfl_up_sync(kFull, incl, o);
        if (lane_id() >= (unsigned)o) incl += y;
    }
    unsigned total = __shfl_sync(kFull, incl, 31);
    unsigned base = 0;
    if (lane_id() == 31 && total) base = atomicAdd(counter, total);
    base = __shfl_sync(kFull, base, 31);
    return base + incl - mine;
}

// ------------------------------------------------------------------------------------------- S2/S3: source sampling
// Angles that only place a point or a direction uniformly use the SFU sine / cosine (absolute error 4e-7 on [0, 2 pi],
// i.e. 2e-7 cm on a 0.5 cm source radius): two instructions instead of the ~150 of a full-precision sincosf, in a kernel
// that is bound by instruction issue.  The reference calls cosf / sinf here (gPET_kernals.cu:465-477, 536-540); the
// oracle keeps libm, and the parity test of this stage holds both to 2e-5.  NOT used where a result is close to +-1 and
// its distance from 1 matters (cosf(delta) of the acollinearity, rotate_dir keeps the reference's own intrinsics).
__device__ __forceinline__ void fast_sincos(float a, float& sn, float& cs) { sn = __sinf(a); cs = __cosf(a); }

__device__ __forceinline__ void sample_shape(int shape, const float* __restrict__ c, uint4 r, float& x, float& y, float& z) {
    float u0 = u01(r.x), u1 = u01(r.y), u2 = u01(r.z);
    if (shape < 0 || shape > 2) shape = 0;
    if (shape == 0) {  // box: centre + full lengths
        x = c[0] + c[3] * (-1.f + 2.f * u0) * 0.5f;
        y = c[1] + c[4] * (-1.f + 2.f * u1) * 0.5f;
        z = c[2] + c[5] * (-1.f + 2.f * u2) * 0.5f;
    } else if (shape == 1) {  // cylinder along z: radius c3, height c4
        float sn, cs;
        fast_sincos(kTwoPi * u0, sn, cs);
        float rr = c[3] * sqrtf(u1);
        x = c[0] + rr * cs;
        y = c[1] + rr * sn;
        z = c[2] + c[4] * (-1.f + 2.f * u2) * 0.5f;
    } else {  // sphere radius c3
        float sn, cs;
        fast_sincos(kTwoPi * u0, sn, cs);
        float ct = -1.f + 2.f * u1;
        float rr = c[3] * cbrtf(u2);
        float st = sqrtf(1.f - ct * ct);
        x = c[0] + rr * st * cs;
        y = c[1] + rr * st * sn;
        z = c[2] + rr * ct;
    }
}

// S4 sampleEkPositron (gPET_kernals.cu:420-443): rejection sampling of the beta+ spectrum from the fitted polynomial
// (coef[0] = endpoint incl. 0.511 MeV, coef[1] = pdf maximum, coef[2..7] = coefficients of E^5..E^0); kinetic energy in eV.
// One Philox block per round.  The 0.511 literals are doubles in the reference, hence the fp64 detours.
__device__ __forceinline__ float sample_ek_positron(const float* __restrict__ coef, Philox& rng) {
    float E, u, sumE;
    do {
        uint4 r = rng.next();
        E = (float)((double)u01(r.x) * ((double)coef[0] - 0.511) + 0.511);
        u = coef[1] * u01(r.y);
        sumE = 0.f;
#pragma unroll
        for (int i = 0; i < 6; i++) sumE += coef[2 + i] * powf(E, (float)(5 - i));
    } while (u > sumE);
    return (float)(((double)E - 0.511) * 1e6);
}

__device__ __forceinline__ float voxel_density_clamped(const PhantomDev& ph, int ix, int iy, int iz) {
    // the reference reads a point-filtered, clamp-addressed 3-D texture (tex3D(dens_tex, ...), initialize.cu:880-882)
    ix = min(max(ix, 0), ph.nx - 1); iy = min(max(iy, 0), ph.ny - 1); iz = min(max(iz, 0), ph.nz - 1);
    return __uint_as_float(__ldg(ph.vox + ((size_t)iz * ph.ny + iy) * ph.nx + ix) & ~15u);
}

// S5 setPositronRange (gPET_kernals.cu:347-418): Gaussian displacement with sigma = Rex/2, Rex = 0.1*b1*E^2/(b2+E) (E in MeV,
// water), then a density-scaled ray march through the voxels.  Restated statement by statement, including its quirks:
// the density used for a step is that of the voxel being ENTERED, and a positron outside the phantom is moved by
// 1000 cm + (r - s)/0.0012905 (`step` keeps its 1000 sentinel in the else branch).
__device__ __forceinline__ void positron_range(const PhantomDev& ph, float& px, float& py, float& pz, float vx, float vy, float vz,
                                               float ekin_eV, bool usedirection, Philox& rng) {
    const float ekin = (float)((double)ekin_eV / 1e6);
    float b1 = 5.44040782f, b2 = 0.369516529f;
    const float Rex = (float)(0.1 * (double)b1 * (double)ekin * (double)ekin / (double)(b2 + ekin));
    const float sigma = Rex / (2 * 1.0f);
    uint4 q = rng.next();
    const float ra = sqrtf(-2.0f * logf(u01(q.x))), rb = sqrtf(-2.0f * logf(u01(q.z)));
    float dx = sigma * (ra * cosf(kTwoPi * u01(q.y)));
    float dy = sigma * (ra * sinf(kTwoPi * u01(q.y)));
    float dz = sigma * (rb * cosf(kTwoPi * u01(q.w)));
    const float r = sqrtf(dx * dx + dy * dy + dz * dz);
    if (usedirection) {
        const float tmp = sqrtf(vx * vx + vy * vy + vz * vz);
        dx = r * vx / tmp; dy = r * vy / tmp; dz = r * vz / tmp;
    }
    float s = 0.f, step;
    int ix = (int)((px - ph.ox) * ph.idx), iy = (int)((py - ph.oy) * ph.idy), iz = (int)((pz - ph.oz) * ph.idz);
    int w = (ix <= 0 || ix >= ph.nx || iy <= 0 || iy >= ph.ny || iz <= 0 || iz >= ph.nz) ? -1 : 1;
    int guard = 0;
    while (s < r && guard++ < 100000) {
        step = 1000.f;
        if (w > 0) {
            b1 = (ph.ox + (ix + (dx > 0.f)) * ph.dx - px) / dx;
            if (step > b1) { step = b1; w = 1; }
            b1 = (ph.oy + (iy + (dy > 0.f)) * ph.dy - py) / dy;
            if (step > b1) { step = b1; w = 2; }
            b1 = (ph.oz + (iz + (dz > 0.f)) * ph.dz - pz) / dz;
            if (step > b1) { step = b1; w = 3; }
            if (w == 1) ix += (dx > 0.f) ? 1 : -1;
            else if (w == 2) iy += (dy > 0.f) ? 1 : -1;
            else iz += (dz > 0.f) ? 1 : -1;
            b2 = voxel_density_clamped(ph, ix, iy, iz);
            step = step * r;
            s += step * b2;
            if (s > r) step += (r - s) / b2;
        } else {
            step += (float)((double)(r - s) / 0.0012905);
            s = r + 100.f;
        }
        px += step * dx / r; py += step * dy / r; pz += step * dz / r;
        if (px < ph.ox || px > (ph.ox + ph.nx * ph.dx)) w = -1;
        if (py < ph.oy || py > (ph.oy + ph.ny * ph.dy)) w = -1;
        if (pz < ph.oz || pz > (ph.oz + ph.nz * ph.dz)) w = -1;
    }
}

// A photon in registers.
struct Photon {
    float x, y, z, E, vx, vy, vz;
    double t;
    int eid, parn, nscat;
};

__device__ __forceinline__ void store_photon(const PhotonQueue& q, unsigned slot, const Photon& p) {
    q.pos_e[slot] = make_float4(p.x, p.y, p.z, p.E);
    q.dir_n[slot] = make_float4(p.vx, p.vy, p.vz, __int_as_float(p.nscat));
    q.t[slot] = p.t;
    q.ids[slot] = make_int2(p.eid, p.parn);
}

// S2 setPosition (gPET_kernals.cu:483-561) for pair k of the frame: both annihilation photons.  Photon a keeps the
// sampled direction, photon b is the acollinear partner; they share the point and the time.
__device__ __forceinline__ void source_pair(const SourceDev* __restrict__ fr, const PhantomDev& ph, uint64_t seed,
                                            unsigned long long k, Photon& a, Photon& b) {
    // source of pair k = number of inclusive prefix sums at or below k: branch-free binary search over the table padded to 64
    // entries (the prefix sums of the unused entries equal the frame's total, fill_source_dev), 6 steps whatever nsource is
    int s = 0;
#pragma unroll
    for (int step = 32; step > 0; step >>= 1) s += (k >= __ldg(&fr->cum_pairs[s + step - 1])) ? step : 0;
    s = min(s, fr->nsource - 1);
    const unsigned long long gk = fr->first_pair + k;
    Philox rng(seed, gk, (uint32_t)kStageSource << 24);
    uint4 r0 = rng.next();
    // truncated-exponential decay time inside the frame (statistically identical to the reference's per-atom
    // test ptime = -T_half*1.442695*log(U) < slice, gPET_kernals.cu:519-521): ptime = -tau log(1 - u frac).  A frame is short
    // against the mean life (frac = 1 - exp(-dt / tau) ~ 1e-2 for F-18 and the 120 s window), where the series of
    // -log1p(-x) = x + x^2/2 + ... converges to double precision in ten terms: ten DFMAs instead of the ~150 instructions
    // of log1p (3 % of this kernel); longer frames / short-lived isotopes keep log1p
    double ud = u01d(r0.x, r0.y);
    const double xq = ud * fr->frac[s];
    double ptime;
    if (xq < 0.03) {
        double acc = 1.0 / 11.0;
#pragma unroll
        for (int n = 10; n >= 1; n--) acc = fma(acc, xq, 1.0 / (double)n);
        ptime = fr->tau_s[s] * (acc * xq);
    } else {
        ptime = -fr->tau_s[s] * log1p(-xq);
    }
    double t_us = (fr->t0_s + ptime) * 1e6;
    uint4 r1 = rng.next();
    float x, y, z;
    sample_shape(fr->shape[s], fr->coeff + 6 * s, r1, x, y, z);
    // isotropic direction (gPET_kernals.cu:536-540)
    float ct = -1.f + 2.f * u01(r0.z);
    float phi = kTwoPi * u01(r0.w);
    float st = sqrtf(1.f - ct * ct);
    float sphi, cphi;
    fast_sincos(phi, sphi, cphi);
    float vx = st * cphi, vy = st * sphi, vz = ct;
    // acollinearity: delta = N(0,1) * sigma (gPET_kernals.cu:549-555)
    uint4 r2 = rng.next();
    float phi2 = kTwoPi * u01(r2.x);
    float g = sqrtf(-2.f * __logf(u01(r2.y))) * __cosf(kTwoPi * u01(r2.z));
    float delta = g * fr->nonangle;
    if (fr->use_prange) {
        // S4 + S5: positron kinetic energy, then its range (gPET_kernals.cu:529-533); direction is sampled (usedirection 0)
        const float ek = sample_ek_positron(fr->iso_coef + 8 * fr->type[s], rng);
        positron_range(ph, x, y, z, 0.f, 0.f, 0.f, ek, false, rng);
    }
    a.x = b.x = x; a.y = b.y = y; a.z = b.z = z;
    a.t = b.t = t_us;
    // records keep the low 31 bits of the 64-bit history numbers (device_types.cuh photon_index)
    a.eid = b.eid = (int)((unsigned)gk & kIdMask);
    a.nscat = b.nscat = 0;
    a.parn = (int)((unsigned)(2ull * gk) & kIdMask); b.parn = (int)((unsigned)(2ull * gk + 1ull) & kIdMask);
    a.vx = vx; a.vy = vy; a.vz = vz;
    a.E = kMC2 + delta * kMC2 * 0.5f;
    rotate_dir(vx, vy, vz, -cosf(delta), phi2);
    b.vx = vx; b.vy = vy; b.vz = vz;
    b.E = kMC2 - delta * kMC2 * 0.5f;
}

// one thread per annihilation pair: photons 2k and 2k+1 of the queue
__global__ void __launch_bounds__(kThreads) k_source(const SourceDev* __restrict__ fr, unsigned long long npairs,
                                                     PhantomDev ph, PhotonQueue q0, uint64_t seed) {
    const unsigned long long nph = 2ull * npairs;
    if (blockIdx.x == 0 && threadIdx.x == 0) *q0.count = (unsigned)min(nph, (unsigned long long)q0.capacity);
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < npairs && 2ull * k + 1ull < q0.capacity;
         k += (unsigned long long)gridDim.x * blockDim.x) {
        Photon a, b;
        source_pair(fr, ph, seed, k, a, b);
        store_photon(q0, (unsigned)(2ull * k), a);
        store_photon(q0, (unsigned)(2ull * k + 1ull), b);
    }
}

// ------------------------------------------------------------------------------------------- P1: phantom transport
// getDistance (gPET_kernals.cu:148-171): flight length along the direction to the PSF-recording sphere; 0 when the line
// misses it.  Statement by statement, IEEE divide and square root (this file is compiled without FMA contraction).
__device__ __forceinline__ float record_distance(const PhantomDev& ph, const Photon& p) {
    const float cx = p.x - ph.rec[0], cy = p.y - ph.rec[1], cz = p.z - ph.rec[2];
    const float a = p.vx * p.vx + p.vy * p.vy + p.vz * p.vz;
    const float b = 2.0f * (p.vx * cx + p.vy * cy + p.vz * cz);
    const float c = (cx * cx + cy * cy + cz * cz) - ph.rec[3] * ph.rec[3];
    const float disc = b * b - 4 * a * c;
    if (disc < 0) return 0.f;
    if (c < 0) return (-b + sqrtf(disc)) / (2 * a);
    if (b < 0) return (-b - sqrtf(disc)) / (2 * a);
    return (-b + sqrtf(disc)) / (2 * a);
}

// Scatter tags (DetectorDev::scat_tag): read back by the coincidence sorter through the single's parn.  Plain stores: the
// slot belongs to this photon alone within its frame.  The fused front end tags a photon when it scatters (its nscat
// then never has to stay in a register up to the panel search); the staged path carries nscat through queue 1 and tags
// at panel entry.
__device__ __forceinline__ void mark_scattered(const PhantomDev& ph, int parn) {
    if (ph.scat_tag != nullptr) ph.scat_tag[(unsigned)parn & ph.scat_mask] = (unsigned char)ph.scat_serial;
}
__device__ __forceinline__ void mark_scattered(const DetectorDev& det, int parn, int nscat) {
    if (nscat > 0 && det.scat_tag != nullptr) det.scat_tag[(unsigned)parn & det.scat_mask] = (unsigned char)det.scat_serial;
}

// One Woodcock flight (gPET_kernals.cu:277-335).  Returns 0: still inside, 1: the photon leaves the stage alive (escaped,
// keeping the overshoot position -- SURVEY quirk 2 -- or below the absorption energy after a Compton, which the reference
// still hands to the detector stage -- quirk 3), 2: photo-absorbed (tof = -0.5 in the reference).
// Where the flight reads its 1-D tables (majorant, cross sections of the voxel's material) from.  TabGlobal: the read-only
// path (L1 / L2), as the reference reads its linear-filter textures (initialize.cu:425-428).  TabShared: per-block copies in
// shared memory of the majorant and of the float4 rows of up to kSmemMats materials present in the phantom, for the energy
// nodes photons can have (below ~600 keV); anything else falls through to the read-only path.  Same values either way.
constexpr int kSmemMats = 3;
struct TabGlobal {
    const TablesDev& tb;
    __device__ __forceinline__ float maj(int i) const { return __ldg(tb.maj_phantom + i); }
    __device__ __forceinline__ float4 xs(int mat, int i) const { return __ldg(tb.xs + (size_t)mat * tb.nen + i); }
};
struct TabShared {
    const TablesDev& tb;
    const float* s_maj;
    const float4* s_xs;
    int nstage;
    unsigned long long slot_map;   // 4 bits per material id: its slot in s_xs, 15 = not staged
    __device__ __forceinline__ float maj(int i) const { return i < nstage ? s_maj[i] : __ldg(tb.maj_phantom + i); }
    __device__ __forceinline__ float4 xs(int mat, int i) const {
        const unsigned sl = (unsigned)(slot_map >> (4 * mat)) & 15u;
        return (sl < (unsigned)kSmemMats && i < nstage) ? s_xs[sl * nstage + i] : __ldg(tb.xs + (size_t)mat * tb.nen + i);
    }
};

template <class Tab>
__device__ __forceinline__ int phantom_flight(Photon& p, Philox& rng, const PhantomDev& ph, const TablesDev& tb, const Tab& tab, float eabs) {
    uint4 r = rng.next();
    int ie; float fe;
    energy_index(tb, p.E, ie, fe);
    float lammin;
    {
        const float a = tab.maj(ie), b = tab.maj(ie + 1);
        lammin = __fdividef(1.0f, fmaf(fe, b - a, a));
    }
    float s = -lammin * __logf(u01(r.x));
    p.x = fmaf(s, p.vx, p.x); p.y = fmaf(s, p.vy, p.y); p.z = fmaf(s, p.vz, p.z);
    p.t += (double)s * kInvSpeedOfLight;
    int ix = (int)((p.x - ph.ox) * ph.idx), iy = (int)((p.y - ph.oy) * ph.idy), iz = (int)((p.z - ph.oz) * ph.idz);
    if (ix <= 0 || ix >= ph.nx || iy <= 0 || iy >= ph.ny || iz <= 0 || iz >= ph.nz) {
        if (ph.rec_on) {   // RECORDPSF == -1 branch (gPET_kernals.cu:288-294)
            const float rr = record_distance(ph, p);
            p.x = fmaf(rr, p.vx, p.x); p.y = fmaf(rr, p.vy, p.y); p.z = fmaf(rr, p.vz, p.z);
            p.t += (double)rr * kInvSpeedOfLight;
        }
        return 1;
    }
    uint32_t vw = __ldg(ph.vox + ((size_t)iz * ph.ny + iy) * ph.nx + ix);
    int mat = (int)(vw & 15u);
    float rho = __uint_as_float(vw & ~15u);
    Xs3 xs;
    {
        const float4 a = tab.xs(mat, ie), b = tab.xs(mat, ie + 1);
        xs.tot = fmaf(fe, b.x - a.x, a.x);
        xs.compt = fmaf(fe, b.y - a.y, a.y);
        xs.rayl = fmaf(fe, b.z - a.z, a.z);
    }
    float lamden = lammin * rho;
    float prob = 1.0f - lamden * xs.tot;
    float u = u01(r.y);
    if (u < prob) return 0;
    prob += lamden * xs.compt;
    if (u < prob) {
        // Compton with binding effects: cos(theta) from the cmpsf surface (gPET_kernals.cu:66-88)
        float costh = surface_lookup(tb.cmpsf, mat, tb.cm_ncp, tb.cm_ne, p.E * tb.cm_ide, u01(r.z) * tb.cm_idcp);
        float efrac = 1.0f / (1.0f + p.E * kIMC2 * (1.0f - costh));
        float phi = kTwoPi * u01(r.w);
        p.E *= efrac;
        p.nscat++;
        mark_scattered(ph, p.parn);
        if (p.E < eabs) return 1;
        rotate_dir(p.vx, p.vy, p.vz, costh, phi);
        return 0;
    }
    prob += lamden * xs.rayl;
    if (u < prob) {
        float costh = surface_lookup(tb.rayff, mat, tb.rl_ncp, tb.rl_ne, p.E * tb.rl_ide, u01(r.z) * tb.rl_idcp);
        float phi = kTwoPi * u01(r.w);
        p.nscat++;
        mark_scattered(ph, p.parn);
        rotate_dir(p.vx, p.vy, p.vz, costh, phi);
        return 0;
    }
    return 2;
}

__global__ void __launch_bounds__(kThreads) k_phantom(PhotonQueue q0, PhotonQueue q1, PhantomDev ph, TablesDev tb,
                                                      float eabs, uint64_t seed, unsigned long long id_base) {
    const unsigned n = min(*q0.count, q0.capacity);
    const unsigned stride = gridDim.x * blockDim.x;
    unsigned next = blockIdx.x * blockDim.x + threadIdx.x;
    bool active = false;
    Photon p{};
    Philox rng(seed, 0, 0);
    while (true) {
        // ---- refill: idle lanes pull their next photon; deferred until a quarter of the warp is idle
        bool want = !active && next < n;
        unsigned wmask = __ballot_sync(kFull, want);
        unsigned amask = __ballot_sync(kFull, active);
        if (wmask && (__popc(wmask) >= 8 || amask == 0)) {
            while (!active && next < n) {
                float4 pe = q0.pos_e[next];
                float4 dn = q0.dir_n[next];
                double tt = q0.t[next];
                int2 id = q0.ids[next];
                next += stride;
                if (pe.w < 0.f || tt <= 0.0) continue;  // gPET_kernals.cu:272
                p.x = pe.x; p.y = pe.y; p.z = pe.z; p.E = pe.w;
                p.vx = dn.x; p.vy = dn.y; p.vz = dn.z; p.nscat = __float_as_int(dn.w);
                p.t = tt; p.eid = id.x; p.parn = id.y;
                rng = Philox(seed, photon_index(p.parn, id_base), (uint32_t)kStagePhantom << 24);
                active = true;
            }
            amask = __ballot_sync(kFull, active);
        }
        if (amask == 0) {
            if (__ballot_sync(kFull, next < n) == 0) break;
            continue;
        }
        bool done_alive = false;
        if (active) {
            const int r = phantom_flight(p, rng, ph, tb, TabGlobal{tb}, eabs);
            done_alive = r == 1;
            if (r) active = false;
        }
        // ---- warp-aggregated append of the photons that left the phantom alive
        unsigned slot = warp_reserve(q1.count, done_alive ? 1u : 0u);
        if (done_alive && slot < q1.capacity) store_photon(q1, slot, p);
    }
}

// ------------------------------------------------------------------------------------------- X1/X2/D1/D2: detector
constexpr int kSlots = 6;  // distinct crystals per photon kept by the adder (reference: Event events[4], no bound check)
// k_detector's block: 4 x 256 threads per SM.  The kernel retires ~2.2 instructions per cycle and SM whatever its mix (it waits
// on dependent ALU chains -- Philox rounds -- and on L1 / L2 round trips), so resident warps are what it needs: 64 registers
// (no spills) and 55 KB of shared memory per block (adder slots 43 KB + 16 staging rows per warp) hold 32 warps per SM with
// up to 8 panels; more panels fall back to 3 blocks.  Round 1 / early round 2: 3 x 256 at 76 registers and 32 staging rows.
constexpr int kDetThreads = 256;
constexpr int kDetBlocksPerSm = 4;
constexpr unsigned kStage = 16;   // hit rows and event records a warp stages before it flushes them

__device__ __forceinline__ void crystal_search(const PanelDev& pd, const DetectorDev& det, float px, float py, float pz,
                                               int& m_id, int& M_id, int& L_id) {
    m_id = 1; M_id = -1; L_id = -1;
    for (int k = 0; k < det.nsurface; k++) {
        const float* c = det.surface + 10 * k;
        float q = c[0] * px * px + c[1] * py * py + c[2] * pz * pz + c[3] * px * py + c[4] * px * pz + c[5] * py * pz +
                  c[6] * px + c[7] * py + c[8] * pz + c[9];
        if (q < 0.f) return;
    }
    float y = pd.ly / 2 + py, z = pd.lz / 2 + pz;
    float my = __fdiv_rn(y, pd.mody + pd.mspy), mz = __fdiv_rn(z, pd.modz + pd.mspz);
    int My = floorf(my) > 0.f ? (int)my : 0, Mz = floorf(mz) > 0.f ? (int)mz : 0;
    M_id = Mz * det.moduleNy + My;
    y = y - My * (pd.mody + pd.mspy);
    z = z - Mz * (pd.modz + pd.mspz);
    if (y > pd.mody || z > pd.modz) return;
    float cy = __fdiv_rn(y, pd.lsoy + pd.spy), cz = __fdiv_rn(z, pd.lsoz + pd.spz);
    int Ly = floorf(cy) > 0.f ? (int)cy : 0, Lz = floorf(cz) > 0.f ? (int)cz : 0;
    L_id = Lz * det.crystalNy + Ly;
    y = y - Ly * (pd.lsoy + pd.spy);
    z = z - Lz * (pd.lsoz + pd.spz);
    if (y > pd.lsoy || z > pd.lsoz) return;
    m_id = 0;
}

// Klein-Nishina sampling for free electrons at rest (gPET_kernals.cu:90-126); one Philox block per rejection round
__device__ __forceinline__ void compton_kn(float E, Philox& rng, float& efrac, float& costh) {
    float e0 = E * kIMC2;
    float twoe = 2.0f * e0;
    float kmin2 = 1.0f / ((1.0f + twoe) * (1.0f + twoe));
    float loge = __logf(1.0f + twoe);
    for (;;) {
        uint4 r = rng.next();
        if (u01(r.x) * (loge + twoe * (1.0f + e0) * kmin2) < loge) efrac = expf(-u01(r.y) * loge);
        else efrac = sqrtf(kmin2 + u01(r.y) * (1.0f - kmin2));
        float mess = e0 * e0 * efrac * (1.0f + efrac * efrac);
        if (u01(r.z) * mess <= mess - (1.0f - efrac) * ((1.0f + twoe) * efrac - 1.0f)) break;
    }
    costh = 1.0f - (1.0f - efrac) / (efrac * e0);
}

// Panel entry (gPET_kernals.cu:963-1009): the first panel (in index order) whose front face the photon's straight line
// crosses while moving along the panel's growth direction.  On success the photon is returned in that panel's local
// frame with the time of flight to the face added; ov.w carries the panel index.
// The exact test of one panel, statement by statement (gPET_kernals.cu:966-1007).
__device__ __forceinline__ bool panel_entry_one(const PanelDev& pd, int i, const Photon& p, float4& pe, float4& ov, double& t) {
    const float lvx = fmaf(p.vz, pd.uxz, fmaf(p.vy, pd.uxy, p.vx * pd.uxx));
    if (!(lvx * pd.dirx >= 0.f)) return false;
    const float rx = p.x - pd.ox, ry = p.y - pd.oy, rz = p.z - pd.oz;
    const float lx = fmaf(rz, pd.uxz, fmaf(ry, pd.uxy, rx * pd.uxx));
    const float q = __fdiv_rn(lx, lvx);
    const float ly = fmaf(rz, pd.uyz, fmaf(ry, pd.uyy, rx * pd.uyx));
    const float lvy = fmaf(p.vz, pd.uyz, fmaf(p.vy, pd.uyy, p.vx * pd.uyx));
    const float y2 = fmaf(-q, lvy, ly);
    if (!(fabsf(y2) < pd.ly / 2)) return false;
    const float lz = fmaf(rz, pd.uzz, fmaf(ry, pd.uzy, rx * pd.uzx));
    const float lvz = fmaf(p.vz, pd.uzz, fmaf(p.vy, pd.uzy, p.vx * pd.uzx));
    const float z2 = fmaf(-q, lvz, lz);
    if (!(fabsf(z2) < pd.lz / 2)) return false;
    pe = make_float4(0.f, y2, z2, p.E);
    ov = make_float4(lvx, lvy, lvz, __int_as_float(i));
    // flight time to the face: the reference divides in fp64, -lx / (c lvx) (gPET_kernals.cu:1003); the fp32 quotient q is
    // already at hand and good to 6e-8 of a sub-nanosecond flight, so the time path keeps its double accumulation without the
    // fp64 divide (2.75 % of k_front's instructions at 11 lanes)
    t = p.t - (double)q * kInvSpeedOfLight;
    return true;
}

// Panels in shared memory: 33 words apart.  Lanes of a warp read DIFFERENT panels (k_detector: every photon its own
// panel; the candidate walks of the panel search), and with the natural 32-word stride of the 128-byte record the same
// field of any two panels sits in the same bank -- an 8-way conflict per read with 8 panels in play.
struct PanelSm {
    PanelDev d;
    float skew;
};
static_assert(sizeof(PanelSm) == 132, "PanelSm must be 33 words");

__device__ __forceinline__ bool panel_entry(const PanelSm* __restrict__ s_panels, const DetectorDev& det, const Photon& p,
                                            float4& pe, float4& ov, double& t) {
    const int npanels = det.npanels;
    if (det.prefilter) {
        // Phase 1, the same instructions for every lane: panels that can possibly accept.  An accepted crossing point lies
        // on the photon's line and on the face, i.e. inside the sphere of radius^2 r2 around the face centre, so a line
        // that misses the sphere (with a margin far above the rounding of this test) cannot be accepted.  3 of 4 panels
        // fall to the direction test or to this one before the divide of the exact test.
        const float vv = fmaf(p.vz, p.vz, fmaf(p.vy, p.vy, p.vx * p.vx));
        unsigned cand = 0;
        if (det.dirmask) {
            // Phase 0 (gpet_run): the panels this direction can reach at all, from the direction table; phase 1 then
            // runs on those few instead of on every panel (32-panel rings: 6-12 instead of 32)
            const float h = 0.5f * (float)kDirBins;
            const int ix = min(max((int)floorf(fmaf(p.vx, h, h)), 0), kDirBins - 1);
            const int iy = min(max((int)floorf(fmaf(p.vy, h, h)), 0), kDirBins - 1);
            const int iz = min(max((int)floorf(fmaf(p.vz, h, h)), 0), kDirBins - 1);
            unsigned rem = __ldg(det.dirmask + (iz * kDirBins + iy) * kDirBins + ix);
            while (rem) {
                const int i = __ffs(rem) - 1;
                rem &= rem - 1;
                const PanelDev& pd = s_panels[i].d;
                const float lvx = fmaf(p.vz, pd.uxz, fmaf(p.vy, pd.uxy, p.vx * pd.uxx));
                const float rx = p.x - pd.ox, ry = p.y - pd.oy, rz = p.z - pd.oz;
                const float dv = fmaf(rz, p.vz, fmaf(ry, p.vy, rx * p.vx));
                const float dd = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
                const float miss2 = fmaf(dd, vv, -dv * dv);
                const bool far = miss2 > fmaf(1.0e-3f, dd, 1.01f * pd.r2) * vv;
                if (lvx * pd.dirx >= 0.f && !far) cand |= 1u << i;
            }
        } else {
            for (int i = 0; i < npanels; i++) {
                const PanelDev& pd = s_panels[i].d;
                const float lvx = fmaf(p.vz, pd.uxz, fmaf(p.vy, pd.uxy, p.vx * pd.uxx));
                const float rx = p.x - pd.ox, ry = p.y - pd.oy, rz = p.z - pd.oz;
                const float dv = fmaf(rz, p.vz, fmaf(ry, p.vy, rx * p.vx));
                const float dd = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
                const float miss2 = fmaf(dd, vv, -dv * dv);                      // |d x v|^2 = distance^2 * |v|^2
                const bool far = miss2 > fmaf(1.0e-3f, dd, 1.01f * pd.r2) * vv;  // NaN compares false: the exact test decides
                if (lvx * pd.dirx >= 0.f && !far) cand |= 1u << i;
            }
        }
        // Phase 2: the exact test on the candidates, in panel order (first accepting panel wins, as in the reference)
        while (cand) {
            const int i = __ffs(cand) - 1;
            cand &= cand - 1;
            if (panel_entry_one(s_panels[i].d, i, p, pe, ov, t)) return true;
        }
        return false;
    }
    for (int i = 0; i < npanels; i++)
        if (panel_entry_one(s_panels[i].d, i, p, pe, ov, t)) return true;
    return false;
}

__device__ __forceinline__ void stage_panels(PanelSm* s_panels, const DetectorDev& det) {
    for (int i = threadIdx.x; i < det.npanels * 32; i += blockDim.x)
        reinterpret_cast<uint32_t*>(s_panels)[(i >> 5) * 33 + (i & 31)] = reinterpret_cast<const uint32_t*>(det.panels)[i];
    __syncthreads();
}
// k_detector keeps the 128-byte records: it reads a dozen consecutive fields of ONE panel per flight, which the aligned
// layout serves with 16-byte loads (measured: skewed 251 us, aligned 235 us per frame of the 32-panel ring)
__device__ __forceinline__ void stage_panels(PanelDev* s_panels, const DetectorDev& det) {
    for (int i = threadIdx.x; i < det.npanels * 32; i += blockDim.x)
        reinterpret_cast<uint32_t*>(s_panels)[i] = reinterpret_cast<const uint32_t*>(det.panels)[i];
    __syncthreads();
}

// staged form: one thread per photon that left the phantom, appended to the compact queue the transport kernel works on
__global__ void __launch_bounds__(kThreads) k_panel_entry(PhotonQueue q1, DetectorDev det, PhotonQueue q2,
                                                          unsigned* __restrict__ counters) {
    extern __shared__ PanelSm s_panels[];
    stage_panels(s_panels, det);
    const unsigned n = min(*q1.count, q1.capacity);
    const unsigned nround = (n + 31u) & ~31u;  // whole warps stay in the loop for the collective append
    unsigned n_on_panel = 0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += gridDim.x * blockDim.x) {
        bool ok = false;
        float4 pe = make_float4(0, 0, 0, 0), ov = make_float4(0, 0, 0, 0);
        double t = 0.0;
        int2 id = make_int2(0, 0);
        int nscat = 0;
        if (i < n) {
            Photon p;
            const float4 a = q1.pos_e[i], dn = q1.dir_n[i];
            p.x = a.x; p.y = a.y; p.z = a.z; p.E = a.w; p.vx = dn.x; p.vy = dn.y; p.vz = dn.z;
            nscat = __float_as_int(dn.w);
            p.t = q1.t[i];
            id = q1.ids[i];
            if (p.t > 0.0) ok = panel_entry(s_panels, det, p, pe, ov, t);
        }
        unsigned slot = warp_reserve(q2.count, ok ? 1u : 0u);
        if (ok) {
            n_on_panel++;
            mark_scattered(det, id.y, nscat);
            if (slot < q2.capacity) {
                q2.pos_e[slot] = pe;
                q2.dir_n[slot] = ov;
                q2.t[slot] = t;
                q2.ids[slot] = id;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_on_panel += __shfl_xor_sync(kFull, n_on_panel, o);
    if (lane_id() == 0 && n_on_panel) atomicAdd(&counters[8], n_on_panel);
}

// Fused front end: source sampling (or the photon queue q0 in PSF mode) -> phantom transport -> panel entry, photon state
// in registers from birth to the panel face; only the photons that enter a panel (about 0.4 of them in the shipped
// geometry) ever reach HBM.  Same device functions and the same Philox counters as the staged kernels above, hence
// identical photons.  Persistent warps: a lane is in one of three states, and the two expensive state changes (pair
// generation, panel search) are run when enough lanes wait for them so that they execute nearly convergent.
// gen_min / entry_min: waiting lanes that trigger pair generation / the panel search

// Entered photons are staged per warp in shared memory and appended 32 at a time: one atomic on the queue count per
// flush (warp-aggregated atomics on one line serialise at 0.67 ns each, tools/microbench/latency.cu -- with an append
// per panel search and the ticket on the same line that was 2/3 of this kernel's time) and full-line stores.
template <int NW>
struct FrontStageT {
    float4 pe[NW][32];
    float4 ov[NW][32];
    double t[NW][32];
    int2 id[NW][32];
};

template <class FS>
__device__ __forceinline__ void front_flush(FS& st, unsigned warp, unsigned lane, unsigned n, const PhotonQueue& q2) {
    __syncwarp();
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(q2.count, n);
    base = __shfl_sync(kFull, base, 0);
    if (lane < n && base + lane < q2.capacity) {
        q2.pos_e[base + lane] = st.pe[warp][lane];
        q2.dir_n[base + lane] = st.ov[warp][lane];
        q2.t[base + lane] = st.t[warp][lane];
        q2.ids[base + lane] = st.id[warp][lane];
    }
    __syncwarp();
}

// kSmemTab: one block of 1024 threads per SM with the 1-D tables in shared memory (TabShared) instead of four blocks of 256
// reading them through L1 (TabGlobal): same warps per SM, same registers; an A/B of where the tables live (GPET_SMEM_TABLES)
constexpr int kFrontThreadsSmem = 1024;
// (five blocks of 256 at 48 registers spill 128 bytes and are slower: 87 -> 93 us, profiles/r02q_kprof_source_front{4,5}.txt)
template <bool kFromQueue, bool kSmemTab>
__global__ void __launch_bounds__(kSmemTab ? kFrontThreadsSmem : kThreads, kSmemTab ? 1 : 4) k_front(const SourceDev* __restrict__ fr, unsigned long long npairs, PhotonQueue q0,
                                                    PhantomDev ph, TablesDev tb, DetectorDev det, float eabs, uint64_t seed,
                                                    PhotonQueue q2, unsigned* __restrict__ q1_count,
                                                    unsigned* __restrict__ counters, unsigned* __restrict__ ticket, int gen_min,
                                                    int entry_min, unsigned long long id_base_q) {
    extern __shared__ __align__(16) unsigned char s_front[];
    using FrontStage = FrontStageT<(kSmemTab ? kFrontThreadsSmem : kThreads) / 32>;
    FrontStage& stage = *reinterpret_cast<FrontStage*>(s_front);
    PanelSm* s_panels = reinterpret_cast<PanelSm*>(s_front + sizeof(FrontStage));
    // [stage | panels | (kSmemTab) 16-byte aligned: xs rows of the staged materials, then the majorant]
    const size_t tab_off = (sizeof(FrontStage) + (size_t)det.npanels * sizeof(PanelSm) + 15) & ~(size_t)15;
    float4* s_xs = reinterpret_cast<float4*>(s_front + tab_off);
    float* s_maj = reinterpret_cast<float*>(s_xs + (kSmemTab ? kSmemMats * ph.tab_nstage : 0));
    if (kSmemTab) {
        for (int i = threadIdx.x; i < ph.tab_nstage; i += blockDim.x) s_maj[i] = __ldg(tb.maj_phantom + i);
        for (int m = 0; m < 16; m++) {
            const unsigned sl = (unsigned)(ph.tab_slot_map >> (4 * m)) & 15u;
            if (sl >= (unsigned)kSmemMats) continue;
            for (int i = threadIdx.x; i < ph.tab_nstage; i += blockDim.x) s_xs[sl * ph.tab_nstage + i] = __ldg(tb.xs + (size_t)m * tb.nen + i);
        }
    }
    stage_panels(s_panels, det);   // ends with the block barrier that also publishes the tables
    enum { NEED = 0, FLY = 1, ESC = 2 };
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned nunits = kFromQueue ? min(*q0.count, q0.capacity) : (unsigned)min(npairs, (unsigned long long)(q2.capacity / 2));
    // index of the frame's first photon; re-read where a photon starts instead of living in two registers
    auto id_base_of_frame = [&]() -> unsigned long long { return kFromQueue ? id_base_q : 2ull * __ldg(&fr->first_pair); };
    if (!kFromQueue && blockIdx.x == 0 && threadIdx.x == 0) *q0.count = 2u * nunits;
    int state = NEED;
    bool has_b = false, exhausted = false;
    Photon p{}, b{};
    unsigned n_out = 0, n_on = 0, staged = 0;
    Philox rng(seed, 0, 0);
    while (true) {
        if (!kFromQueue && state == NEED && has_b) {   // second photon of the pair
            p = b;
            has_b = false;
            rng = Philox(seed, photon_index(p.parn, id_base_of_frame()), (uint32_t)kStagePhantom << 24);
            state = FLY;
        }
        unsigned need = __ballot_sync(kFull, state == NEED);
        unsigned fly = __ballot_sync(kFull, state == FLY);
        if (!exhausted && need && (__popc(need) >= gen_min || fly == 0)) {
            const unsigned cnt = __popc(need);
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(ticket, cnt);
            base = __shfl_sync(kFull, base, 0);
            if (state == NEED) {
                const unsigned idx = base + __popc(need & lt_mask);
                if (idx < nunits) {
                    bool live = true;
                    if (kFromQueue) {
                        const float4 pe = __ldcs(q0.pos_e + idx), dn = __ldcs(q0.dir_n + idx);
                        p.t = __ldcs(q0.t + idx);
                        const int2 id = __ldcs(q0.ids + idx);
                        p.x = pe.x; p.y = pe.y; p.z = pe.z; p.E = pe.w;
                        p.vx = dn.x; p.vy = dn.y; p.vz = dn.z; p.nscat = __float_as_int(dn.w);
                        p.eid = id.x; p.parn = id.y;
                        live = !(pe.w < 0.f || p.t <= 0.0);  // gPET_kernals.cu:272
                    } else {
                        source_pair(fr, ph, seed, idx, p, b);
                        has_b = true;
                    }
                    if (live) {
                        rng = Philox(seed, photon_index(p.parn, id_base_of_frame()), (uint32_t)kStagePhantom << 24);
                        state = FLY;
                    }
                }
            }
            if (base + cnt >= nunits) exhausted = true;
        }
        if (state == FLY) {
            int r;
            if (kSmemTab) r = phantom_flight(p, rng, ph, tb, TabShared{tb, s_maj, s_xs, ph.tab_nstage, ph.tab_slot_map}, eabs);
            else r = phantom_flight(p, rng, ph, tb, TabGlobal{tb}, eabs);
            if (r == 1) state = ESC;
            else if (r == 2) state = NEED;
        }
        const unsigned esc = __ballot_sync(kFull, state == ESC);
        fly = __ballot_sync(kFull, state == FLY);
        if (esc && (__popc(esc) >= entry_min || fly == 0)) {
            bool ok = false;
            float4 pe = make_float4(0, 0, 0, 0), ov = make_float4(0, 0, 0, 0);
            double t = 0.0;
            if (state == ESC) {
                n_out++;
                ok = panel_entry(s_panels, det, p, pe, ov, t);
                state = NEED;
            }
            const unsigned okmask = __ballot_sync(kFull, ok);
            const unsigned total = __popc(okmask);
            if (staged + total > 32u) {
                front_flush(stage, warp, lane, staged, q2);
                staged = 0;
            }
            if (ok) {
                n_on++;
                const unsigned pos = staged + __popc(okmask & lt_mask);
                stage.pe[warp][pos] = pe;
                stage.ov[warp][pos] = ov;
                stage.t[warp][pos] = t;
                stage.id[warp][pos] = make_int2(p.eid, p.parn);
            }
            staged += total;
        }
        if (exhausted && __ballot_sync(kFull, state != NEED || has_b) == 0) break;
    }
    if (staged) front_flush(stage, warp, lane, staged, q2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_out += __shfl_xor_sync(kFull, n_out, o);
        n_on += __shfl_xor_sync(kFull, n_on, o);
    }
    // one pair of tallies per block (per warp they were 9.5 k adds on one line at the tail of the kernel)
    __shared__ unsigned s_tally[2];
    if (threadIdx.x < 2) s_tally[threadIdx.x] = 0;
    __syncthreads();
    if (lane == 0) {
        if (n_out) atomicAdd(&s_tally[0], n_out);
        if (n_on) atomicAdd(&s_tally[1], n_on);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_tally[0]) atomicAdd(q1_count, s_tally[0]);
        if (s_tally[1]) atomicAdd(&counters[8], s_tally[1]);
    }
}

// Adder slots ("columns") in shared memory.  A slot is one crystal of the photon's panel: key = (module << 16) |
// crystal-in-module (a photon never leaves its panel).  k_detector_v1 gives every thread its own column [slot][thread];
// k_detector gives every WARP a pool of 32 + kSpare columns (below).
struct SlotsSmem {
    int key[kSlots][kDetThreads];
    float E[kSlots][kDetThreads], x[kSlots][kDetThreads], y[kSlots][kDetThreads], z[kSlots][kDetThreads];
    double t[kSlots][kDetThreads];
};

__device__ __forceinline__ int event_siten(int depth, int panel_id, int modn, int cryn, const DetectorDev& det) {
    return depth == 0 ? 0 : depth == 1 ? panel_id : depth == 2 ? panel_id * det.moduleN + modn
                                                             : (panel_id * det.moduleN + modn) * det.crystalN + cryn;
}

constexpr unsigned kDetChunk = 32;   // photons a warp claims from the queue with one ticket atomic

// ---- k_detector -----------------------------------------------------------------------------------------------------
// Photon transport inside a panel (photonde, gPET_kernals.cu:1018-1192, with comsam :90-126, adder / readout :737-813) over
// the compact panel-entry queue.  Persistent warps with lane refill, one lane = one photon from entry to readout.  Against
// round 1's kernel ("v1" below; ncu: 12.3 of 32 lanes per instruction) the divergent parts run differently:
//
//  * Klein-Nishina rejection sampling by the WHOLE warp.  In v1 the 2-6 lanes whose flight ended in a Compton scatter ran
//    the rejection loop nested in the flight loop -- a Philox block and the acceptance test per round, as many rounds as
//    the unluckiest of them needed -- while the other lanes waited (19 % of the kernel's warp instructions at 4 lanes).
//    Here the warp evaluates floor(32 / nC) consecutive rounds of each of the nC scattering photons AT ONCE, lane
//    L = (photon L mod nC, round L div nC), every lane one Philox block and one acceptance test; each photon then takes
//    its first accepted round.  Round k of a photon uses the Philox block it would have used sequentially (block counter
//    + k) and the counter advances past the accepted round only, so the draws, and therefore the results, are those of the
//    sequential loop bit for bit; the blocks of later rounds are simply never used.  One pass settles all photons in
//    > 99 % of the cases (acceptance is ~2/3 per round); the pass repeats for the rest.
//  * The three IEEE divides of an adder / readout centroid share one correctly rounded reciprocal (div3_rn below).
//  * crystal_search divides by four constants of the panel: their reciprocals come with the panel record (div_rcp).
//  * Hits and events are staged per warp in shared memory and leave 32 at a time: one atomic per flush instead of one per
//    loop iteration (the reservation's round trip was 7 % of v1's stall samples), hit rows as full-line vector stores into
//    the SoA hit buffer, events as 1.5 KB of consecutive 16-byte stores.
//
// A measured dead end, for the record (profiles/r02a_*): making the rejection rounds loop iterations of their own (a lane
// state "KN" next to "FLY", both sharing the iteration's Philox block) and pooling readout over spare adder columns took
// the kernel from 130 to 170 us: the scattering lanes no longer fly, the flight body -- the most expensive piece -- ran
// with 15 instead of 24 lanes, and every per-iteration cost was paid 1.5 times as often.

// a / b, b > 0 normal, with rcp = the correctly rounded 1 / b: one multiply and two FMAs give the correctly rounded quotient
// (Markstein's correction: q = RN(a * rcp), r = a - b q exactly, RN(q + r * rcp)).  Checked against IEEE division on
// 2e7 random operands of the magnitudes that occur here (tests/test_oracle_units.py) -- and the parity tests compare the
// events of this kernel with the oracle's plain divisions.
__device__ __forceinline__ float div_rcp(float a, float b, float rcp) {
    const float q = __fmul_rn(a, rcp);
    return __fmaf_rn(__fmaf_rn(-b, q, a), rcp, q);
}

// energy-weighted centroid of two deposits, as the reference forms it: fma(x_i, E_i, x E) / (E_i + E)
struct Centroid { float x, y, z, E; };
__device__ __forceinline__ Centroid merge_centroid(float xi, float yi, float zi, float Ei, float x, float y, float z, float E) {
    Centroid c;
    c.E = __fadd_rn(Ei, E);
    const float rcp = __frcp_rn(c.E);
    // energies are 1e3 .. 2e6 eV and coordinates a few cm: far inside the range where the correction is exact; anything
    // else (a zero or non-finite sum) takes the plain divide
    if (c.E > 1.0f && c.E < 1.0e30f) {
        c.x = div_rcp(__fmaf_rn(xi, Ei, __fmul_rn(x, E)), c.E, rcp);
        c.y = div_rcp(__fmaf_rn(yi, Ei, __fmul_rn(y, E)), c.E, rcp);
        c.z = div_rcp(__fmaf_rn(zi, Ei, __fmul_rn(z, E)), c.E, rcp);
    } else {
        c.x = __fdiv_rn(__fmaf_rn(xi, Ei, __fmul_rn(x, E)), c.E);
        c.y = __fdiv_rn(__fmaf_rn(yi, Ei, __fmul_rn(y, E)), c.E);
        c.z = __fdiv_rn(__fmaf_rn(zi, Ei, __fmul_rn(z, E)), c.E);
    }
    return c;
}

__device__ __forceinline__ bool adder_fast(SlotsSmem& sl, int col, int& n, int key, float E, float x, float y, float z, double t) {
    for (int k = 0; k < n; k++) {
        if (sl.key[k][col] == key) {
            const Centroid c = merge_centroid(sl.x[k][col], sl.y[k][col], sl.z[k][col], sl.E[k][col], x, y, z, E);
            sl.x[k][col] = c.x; sl.y[k][col] = c.y; sl.z[k][col] = c.z; sl.E[k][col] = c.E;
            return true;
        }
    }
    if (n >= kSlots) return false;
    sl.key[n][col] = key; sl.E[n][col] = E; sl.x[n][col] = x; sl.y[n][col] = y; sl.z[n][col] = z; sl.t[n][col] = t;
    n++;
    return true;
}

__device__ __forceinline__ unsigned readout_merge_fast(SlotsSmem& sl, int col, int nslot, int depth, int rpolicy) {
    unsigned deadmask = 0;
#pragma unroll 1
    for (int i = 0; i < nslot - 1; i++) {
        if (deadmask >> i & 1u) continue;
        const int ki = sl.key[i][col];
#pragma unroll 1
        for (int j = i + 1; j < nslot; j++) {
            if (deadmask >> j & 1u) continue;
            const int kj = sl.key[j][col];
            const bool same = depth <= 1 ? true : depth == 2 ? (ki >> 16) == (kj >> 16) : ki == kj;
            if (!same) continue;
            const float Ei = sl.E[i][col], Ej = sl.E[j][col];
            if (rpolicy == 1) {
                const Centroid c = merge_centroid(sl.x[i][col], sl.y[i][col], sl.z[i][col], Ei, sl.x[j][col], sl.y[j][col], sl.z[j][col], Ej);
                sl.x[i][col] = c.x; sl.y[i][col] = c.y; sl.z[i][col] = c.z; sl.E[i][col] = c.E;
            } else if (!(Ei > Ej)) {
                sl.key[i][col] = kj; sl.E[i][col] = Ej; sl.x[i][col] = sl.x[j][col];
                sl.y[i][col] = sl.y[j][col]; sl.z[i][col] = sl.z[j][col]; sl.t[i][col] = sl.t[j][col];
            }
            deadmask |= 1u << j;
        }
    }
    return deadmask;
}

// crystalSearch (gPET_kernals.cu:1236-1279) with the panel's four divisors replaced by their reciprocals (PanelDev::rcp*)
__device__ __forceinline__ void crystal_search_rcp(const PanelDev& pd, const DetectorDev& det, float px, float py, float pz,
                                                   int& m_id, int& M_id, int& L_id) {
    m_id = 1; M_id = -1; L_id = -1;
    for (int k = 0; k < det.nsurface; k++) {
        const float* c = det.surface + 10 * k;
        float q = c[0] * px * px + c[1] * py * py + c[2] * pz * pz + c[3] * px * py + c[4] * px * pz + c[5] * py * pz +
                  c[6] * px + c[7] * py + c[8] * pz + c[9];
        if (q < 0.f) return;
    }
    float y = pd.ly / 2 + py, z = pd.lz / 2 + pz;
    const float dmy = pd.mody + pd.mspy, dmz = pd.modz + pd.mspz;
    float my = div_rcp(y, dmy, pd.rcp_my), mz = div_rcp(z, dmz, pd.rcp_mz);
    int My = floorf(my) > 0.f ? (int)my : 0, Mz = floorf(mz) > 0.f ? (int)mz : 0;
    M_id = Mz * det.moduleNy + My;
    y = y - My * dmy;
    z = z - Mz * dmz;
    if (y > pd.mody || z > pd.modz) return;
    const float dcy = pd.lsoy + pd.spy, dcz = pd.lsoz + pd.spz;
    float cy = div_rcp(y, dcy, pd.rcp_cy), cz = div_rcp(z, dcz, pd.rcp_cz);
    int Ly = floorf(cy) > 0.f ? (int)cy : 0, Lz = floorf(cz) > 0.f ? (int)cz : 0;
    L_id = Lz * det.crystalNy + Ly;
    y = y - Ly * dcy;
    z = z - Lz * dcz;
    if (y > pd.lsoy || z > pd.lsoz) return;
    m_id = 0;
}

// per-warp staging of the rows that leave the SM: 32 hits (SoA pieces) and 32 events (48-byte records)
struct WarpStage {
    int4 hid[kStage];        // hits: parn, pann, modn, cryn
    float4 hf[kStage];       // E, x, y, z
    double ht[kStage];
    int4 ev[kStage * 3];     // events: record k = pieces 3k .. 3k+2
    int htype[kStage];
    unsigned char kn_src[32];   // lane of the j-th scattering photon (cooperative Klein-Nishina pass)
};

constexpr int kKnMaxRounds = 8;   // rounds of one photon evaluated side by side at most (acceptance ~2/3: 8 rounds fail 2e-4 of the time)

// lane -> (photon j = lane mod nC, round k = lane div nC) and the lanes that hold the rounds of photon 0 (bits 0, nC, 2 nC, ...
// for the R = min(32 / nC, kKnMaxRounds) rounds of a pass), by nC = 1..32: tables instead of two integer divisions and a loop
struct KnTables {
    unsigned char jk[33][32];   // j | k << 5
    unsigned rounds[33];        // stride mask
};
__device__ KnTables c_kn;   // read through the L1 (lanes read 32 consecutive bytes: one sector; a __constant__ table would serialise them)

__global__ void __launch_bounds__(kDetThreads, kDetBlocksPerSm) k_detector(PhotonQueue q2, DetectorDev det, TablesDev tb, float eabs,
                                                       int rdepth, int rpolicy, int record_hits, HitBuffer hits, EventBuf ev,
                                                       unsigned* __restrict__ counters, unsigned* __restrict__ ticket, uint64_t seed,
                                                       int refill_min, unsigned long long id_base) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    SlotsSmem& sl = *reinterpret_cast<SlotsSmem*>(s_raw);
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    WarpStage& ws = reinterpret_cast<WarpStage*>(s_raw + sizeof(SlotsSmem))[warp];
    PanelDev* s_panels = reinterpret_cast<PanelDev*>(s_raw + sizeof(SlotsSmem) + sizeof(WarpStage) * (kDetThreads / 32));
    stage_panels(s_panels, det);
    // programmatic dependent launch: everything above overlapped the tail of the front-end kernel; its queue is read below
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const int tid = threadIdx.x;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned n = min(*q2.count, q2.capacity);
    const int depth = (rdepth != 3 && rpolicy == 1) ? 2 : rdepth;
    bool active = false, exhausted = false;
    unsigned chunk_pos = 0, chunk_end = 0;   // the warp's claimed share of the queue (warp-uniform)
    unsigned seen = 0;                       // ticket value after this warp's last claim
    unsigned hstaged = 0, estaged = 0;       // warp-uniform: staged hits, staged events
    const unsigned nwarps = gridDim.x * (kDetThreads / 32);
    float x = 0, y = 0, z = 0, E = 0, vx = 0, vy = 0, vz = 0;
    double t = 0;
    int eid = 0, parn = 0, pa = 0, nslot = 0;
    unsigned n_drop_adder = 0;
    Philox rng(seed, 0, 0);

    auto flush_hits = [&]() {
        __syncwarp();
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(hits.count, hstaged);
        base = __shfl_sync(kFull, base, 0);
        if (lane < hstaged && base + lane < hits.capacity) {
            hits.id4[base + lane] = ws.hid[lane];
            hits.f4[base + lane] = ws.hf[lane];
            hits.t[base + lane] = ws.ht[lane];
            hits.type[base + lane] = ws.htype[lane];
        }
        hstaged = 0;
        __syncwarp();
    };
    auto flush_events = [&]() {
        __syncwarp();
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(ev.count, estaged);
        base = __shfl_sync(kFull, base, 0);
        const unsigned room = base < ev.capacity ? min(estaged, ev.capacity - base) : 0u;
        int4* dst = reinterpret_cast<int4*>(ev.rec + base);
#pragma unroll
        for (unsigned p = 0; p * 32u < 3u * kStage; p++) {
            const unsigned piece = p * 32u + lane;
            if (piece < 3u * room) dst[piece] = ws.ev[piece];
        }
        estaged = 0;
        __syncwarp();
    };

    while (true) {
        // ---- refill idle lanes from the warp's chunk of the queue; a new chunk costs one ticket atomic per kDetChunk photons
        unsigned amask = __ballot_sync(kFull, active);
        if (!exhausted && (__popc(~amask) >= refill_min || amask == 0)) {
            if (chunk_pos == chunk_end) {
                // guided self-scheduling: full chunks while the queue is long, smaller ones as it drains (remaining /
                // 2 x warps, from the ticket value this warp saw last), so that the warps run dry together
                const unsigned left = n > seen ? n - seen : 0u;
                const unsigned want = max(2u, min(kDetChunk, left / (2u * nwarps)));
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(ticket, want);
                base = __shfl_sync(kFull, base, 0);
                seen = base + want;
                chunk_pos = min(base, n);
                chunk_end = min(base + want, n);
                if (base >= n) exhausted = true;
            }
            const unsigned need = ~amask;
            const unsigned avail = chunk_end - chunk_pos;
            const unsigned rank = __popc(need & lt_mask);
            if (!active && rank < avail) {
                const unsigned idx = chunk_pos + rank;
                const float4 pe = __ldcs(q2.pos_e + idx);
                const float4 dn = __ldcs(q2.dir_n + idx);
                t = __ldcs(q2.t + idx);
                const int2 id = __ldcs(q2.ids + idx);
                x = pe.x; y = pe.y; z = pe.z; E = pe.w;
                vx = dn.x; vy = dn.y; vz = dn.z; pa = __float_as_int(dn.w);
                eid = id.x; parn = id.y;
                rng = Philox(seed, photon_index(parn, id_base), (uint32_t)kStageDetector << 24);
                nslot = 0;
                active = true;
            }
            chunk_pos += min((unsigned)__popc(need), avail);
            amask = __ballot_sync(kFull, active);
        }
        if (amask == 0) {
            if (exhausted) break;
            continue;
        }
        // ---- one Woodcock flight per active lane (gPET_kernals.cu:1018-1050) and the choice of the interaction.  Up to two
        // hits per flight (Compton deposit + absorption of the remainder), both at the same point
        int nh = 0, h_key = 0, h_type0 = 0, mid = 1;
        float h_E0 = 0.f, h_E1 = 0.f;
        bool finished = false, turn = false, compton = false;
        float costh = 1.f, phi = 0.f;
        if (active) {
            const PanelDev& pd = s_panels[pa];
            uint4 r = rng.next();
            int ie; float fe;
            energy_index(tb, E, ie, fe);
            float lammin = __fdividef(1.0f, lerp_table(tb.maj_detector, ie, fe));
            float s = -lammin * __logf(u01(r.x));
            x = fmaf(s, vx, x); y = fmaf(s, vy, y); z = fmaf(s, vz, z);
            t += (double)s * kInvSpeedOfLight;
            if (fabsf(y) > pd.ly * 0.5f || fabsf(z) > pd.lz * 0.5f || x * pd.dirx < 0.f || x * pd.dirx > pd.lx) {
                finished = true;  // left the panel
            } else {
                int M_id, L_id;
                crystal_search_rcp(pd, det, x, y, z, mid, M_id, L_id);
                h_key = (M_id << 16) | (L_id & 0xffff);
                float rho = det.dens[mid];
                int mat = det.mat[mid];
                Xs3 xs = lerp_xs(tb, mat, ie, fe);
                float lamden = lammin * rho;
                float prob = fmaxf(1.0f - lamden * xs.tot, 0.f);
                float u = u01(r.y);
                if (u >= prob) {
                    prob += lamden * xs.compt;
                    if (u < prob) {
                        compton = true;
                        phi = kTwoPi * u01(r.z);
                    } else {
                        prob += lamden * xs.rayl;
                        if (u < prob) {
                            costh = surface_lookup(tb.rayff, mat, tb.rl_ncp, tb.rl_ne, E * tb.rl_ide, u01(r.z) * tb.rl_idcp);
                            phi = kTwoPi * u01(r.w);
                            turn = true;
                        } else {
                            if (mid == 0) { h_type0 = 4; h_E0 = E; nh = 1; }
                            finished = true;
                        }
                    }
                }
            }
        }
        // ---- Klein-Nishina sampling of the scattering photons (comsam, gPET_kernals.cu:90-126) by the whole warp
        unsigned cmask = __ballot_sync(kFull, compton);
        float efrac = 1.f;
        bool pending = compton;
        while (cmask) {
            const unsigned nC = __popc(cmask);
            const unsigned R = min(32u / nC, (unsigned)kKnMaxRounds);
            const unsigned mine = __popc(cmask & lt_mask);          // this lane's number among the scattering photons
            if (pending) ws.kn_src[mine] = (unsigned char)lane;
            __syncwarp();
            const unsigned jk = __ldg(&c_kn.jk[nC][lane]);
            const unsigned j = jk & 31u, k = jk >> 5;               // this lane evaluates round k of photon j
            const int src = ws.kn_src[j];
            const float Ej = __shfl_sync(kFull, E, src);
            const unsigned c0 = __shfl_sync(kFull, rng.c0, src), c1 = __shfl_sync(kFull, rng.c1, src), c3 = __shfl_sync(kFull, rng.c3, src);
            bool acc = false;
            float ef = 0.f;
            if (k < R) {
                Philox g(seed, 0, (uint32_t)kStageDetector << 24);
                g.c0 = c0; g.c1 = c1; g.c3 = c3 + k;
                const uint4 q = g.next();
                const float e0 = Ej * kIMC2;
                const float twoe = 2.0f * e0;
                const float kmin2 = 1.0f / ((1.0f + twoe) * (1.0f + twoe));
                const float loge = __logf(1.0f + twoe);
                if (u01(q.x) * (loge + twoe * (1.0f + e0) * kmin2) < loge) ef = expf(-u01(q.y) * loge);
                else ef = sqrtf(kmin2 + u01(q.y) * (1.0f - kmin2));
                const float mess = e0 * e0 * ef * (1.0f + ef * ef);
                acc = u01(q.z) * mess <= mess - (1.0f - ef) * ((1.0f + twoe) * ef - 1.0f);
            }
            const unsigned accmask = __ballot_sync(kFull, acc);
            // the lanes that hold this photon's rounds: mine, mine + nC, ...; the first accepted one is the sampler's answer
            const unsigned pat = __ldg(&c_kn.rounds[nC]) << mine;
            const unsigned hit = pending ? (accmask & pat) : 0u;
            const int win = hit ? __ffs(hit) - 1 : 0;
            const float efw = __shfl_sync(kFull, ef, win);
            if (pending) {
                if (hit) {
                    efrac = efw;
                    rng.c3 += __popc(pat & ((1u << win) - 1u)) + 1u;   // rounds before the accepted one were rejections; its block is the last one consumed
                    pending = false;
                } else {
                    rng.c3 += R;                                  // R rejections
                }
            }
            __syncwarp();
            cmask = __ballot_sync(kFull, pending);
        }
        if (compton) {
            const float e0 = E * kIMC2;
            costh = 1.0f - (1.0f - efrac) / (efrac * e0);
            const float de = E * (1.0f - efrac);
            if (mid == 0) { h_type0 = 1; h_E0 = de; nh = 1; }
            E -= de;
            if (E < eabs) {
                if (mid == 0) { h_E1 = E; nh = 2; }  // type 2: remainder absorbed on the spot
                finished = true;
            } else {
                turn = true;
            }
        }
        if (turn) rotate_dir(vx, vy, vz, costh, phi);
        // ---- adder on the fly
        if (nh >= 1) {
            if (!adder_fast(sl, tid, nslot, h_key, h_E0, x, y, z, t)) n_drop_adder++;
            if (nh == 2 && !adder_fast(sl, tid, nslot, h_key, h_E1, x, y, z, t)) n_drop_adder++;
        }
        // ---- hits into the warp's staging rows
        if (record_hits) {
#pragma unroll
            for (int pass = 0; pass < 2; pass++) {
                const bool has = nh > pass;
                const unsigned m = __ballot_sync(kFull, has);
                if (m == 0u) break;
                const unsigned cnt = __popc(m);
                if (hstaged + cnt > kStage) flush_hits();
                unsigned rank = __popc(m & lt_mask);
                if (cnt > kStage) {   // rare: more deposits in one pass than the staging rows hold; the first kStage leave at once
                    if (has && rank < kStage) {
                        ws.hid[rank] = make_int4(parn, s_panels[pa].id, h_key >> 16, h_key & 0xffff);
                        ws.hf[rank] = make_float4(pass ? h_E1 : h_E0, x, y, z);
                        ws.ht[rank] = t;
                        ws.htype[rank] = pass ? 2 : h_type0;
                    }
                    hstaged = kStage;
                    flush_hits();
                    rank -= kStage;   // wraps for the lanes just served: they fail the test below
                }
                if (has && rank < kStage) {
                    const unsigned p = hstaged + rank;
                    ws.hid[p] = make_int4(parn, s_panels[pa].id, h_key >> 16, h_key & 0xffff);
                    ws.hf[p] = make_float4(pass ? h_E1 : h_E0, x, y, z);
                    ws.ht[p] = t;
                    ws.htype[p] = pass ? 2 : h_type0;
                }
                hstaged += cnt > kStage ? cnt - kStage : cnt;
            }
        }
        // ---- photon finished: readout (gPET_kernals.cu:756-813), events into the warp's staging records
        if (finished) active = false;
        const bool mine_ev = finished && nslot > 0;
        unsigned deadmask = 0;
        if (mine_ev && nslot > 1 && rdepth != 3) deadmask = readout_merge_fast(sl, tid, nslot, depth, rpolicy);
        unsigned ne = mine_ev ? (unsigned)(nslot - __popc(deadmask)) : 0u;
        if (__ballot_sync(kFull, ne != 0u)) {
            // at most kSlots events per lane: rounds of at most one event per lane; a lane whose record does not fit the
            // staging rows this round comes again after the flush
            int knext = 0;
#pragma unroll 1
            while (true) {
                const bool want = ne != 0u;
                const unsigned m = __ballot_sync(kFull, want);
                if (m == 0u) break;
                if (estaged + (unsigned)__popc(m) > kStage) flush_events();
                const unsigned room = kStage - estaged, rank = __popc(m & lt_mask);   // more than kStage at once: the rest next round
                const bool has = want && rank < room;
                if (has) {
                    while (deadmask >> knext & 1u) knext++;
                    const int key = sl.key[knext][tid];
                    const int panel_id = s_panels[pa].id;
                    EventRec r;
                    r.parn = parn; r.pann = panel_id; r.modn = key >> 16; r.cryn = key & 0xffff;
                    r.siten = event_siten(depth, panel_id, r.modn, r.cryn, det);
                    r.eventid = eid;
                    r.t = sl.t[knext][tid]; r.E = sl.E[knext][tid];
                    r.x = sl.x[knext][tid]; r.y = sl.y[knext][tid]; r.z = sl.z[knext][tid];
                    store_event_rec(reinterpret_cast<EventRec*>(ws.ev) + (estaged + rank), r);
                    knext++;
                    ne--;
                }
                estaged += min((unsigned)__popc(m), room);
            }
        }
        if (mine_ev) nslot = 0;
    }
    if (hstaged) flush_hits();
    if (estaged) flush_events();
    // per-warp tallies
    unsigned b = n_drop_adder;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(kFull, b, o);
    if (lane == 0 && b) atomicAdd(&counters[9], b);
}

// hits in the reference's file layout (gPET.cu:367-376; readOutput.m:3-16), for the dumps and gpet_fetch_hits only
__global__ void k_hits_to_rows(HitBuffer h, unsigned n, int* __restrict__ id5, float* __restrict__ f5) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 a = h.id4[i];
        const float4 f = h.f4[i];
        int* hi = id5 + 5ull * i;
        float* hf = f5 + 5ull * i;
        hi[0] = a.x; hi[1] = a.y; hi[2] = a.z; hi[3] = a.w; hi[4] = h.type[i];
        hf[0] = f.x; hf[1] = (float)h.t[i]; hf[2] = f.y; hf[3] = f.z; hf[4] = f.w;
    }
}

__global__ void k_hits_to_aos(HitBuffer h, unsigned n, gpet_hit* __restrict__ out) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 a = h.id4[i];
        const float4 f = h.f4[i];
        gpet_hit o;
        o.parn = a.x; o.pann = a.y; o.modn = a.z; o.cryn = a.w; o.type = h.type[i];
        o.E = f.x; o.t = h.t[i]; o.t32 = (float)o.t; o.x = f.y; o.y = f.z; o.z = f.w;
        out[i] = o;
    }
}

// ------------------------------------------------------------------------------------------- host AoS <-> queue
__global__ void k_aos_to_queue(const gpet_photon* __restrict__ aos, PhotonQueue q, unsigned n) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *q.count = n;
    for (; i < n; i += gridDim.x * blockDim.x) {
        gpet_photon p = aos[i];
        q.pos_e[i] = make_float4(p.x, p.y, p.z, p.E);
        q.dir_n[i] = make_float4(p.vx, p.vy, p.vz, __int_as_float(p.nscat));
        q.t[i] = p.t;
        q.ids[i] = make_int2(p.eventid, p.parn);
    }
}

__global__ void k_queue_to_aos(PhotonQueue q, gpet_photon* __restrict__ aos) {
    const unsigned n = min(*q.count, q.capacity);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 pe = q.pos_e[i], dn = q.dir_n[i];
        int2 id = q.ids[i];
        gpet_photon p;
        p.x = pe.x; p.y = pe.y; p.z = pe.z; p.E = pe.w;
        p.vx = dn.x; p.vy = dn.y; p.vz = dn.z; p.nscat = __float_as_int(dn.w);
        p.t = q.t[i];
        p.eventid = id.x; p.parn = id.y;
        aos[i] = p;
    }
}

// S6 setPositionForPhoton (gPET_kernals.cu:563-604): positron phase space -> annihilation photon pair.  Positron i of
// the batch gives photons 2i and 2i+1 (the reference puts them at i and i+total); both photons carry the positron's
// time -- the reference never writes d_time of the second photon, which silently drops it: fixed here.
__global__ void __launch_bounds__(kThreads) k_psf_positron(const gpet_photon* __restrict__ pos, PhotonQueue q0, unsigned n,
                                                           unsigned long long first, PhantomDev ph, float nonangle, int use_prange,
                                                           uint64_t seed) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *q0.count = min(2u * n, q0.capacity);
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < 2u * n && p < q0.capacity; p += gridDim.x * blockDim.x) {
        const unsigned i = p >> 1;
        const int which = (int)(p & 1u);
        const gpet_photon e = pos[i];
        const unsigned long long gi = first + i;
        Philox rng(seed, gi, (uint32_t)kStagePsfPositron << 24);
        uint4 r0 = rng.next();
        float x = e.x, y = e.y, z = e.z;
        float ct = -1.f + 2.f * u01(r0.x);
        float phi = kTwoPi * u01(r0.y);
        float st = sqrtf(1.f - ct * ct);
        float vx = st * cosf(phi), vy = st * sinf(phi), vz = ct;
        float phi2 = kTwoPi * u01(r0.z);
        uint4 r1 = rng.next();
        float g = sqrtf(-2.f * logf(u01(r1.x))) * cosf(kTwoPi * u01(r1.y));
        float delta = g * nonangle;
        if (use_prange) positron_range(ph, x, y, z, e.vx, e.vy, e.vz, e.E, true, rng);
        float E;
        if (which == 0) {
            E = kMC2 + delta * kMC2 * 0.5f;
        } else {
            rotate_dir(vx, vy, vz, -cosf(delta), phi2);
            E = kMC2 - delta * kMC2 * 0.5f;
        }
        q0.pos_e[p] = make_float4(x, y, z, E);
        q0.dir_n[p] = make_float4(vx, vy, vz, __int_as_float(0));
        q0.t[p] = e.t;
        q0.ids[p] = make_int2((int)((unsigned)gi & kIdMask), (int)((unsigned)(2ull * gi + which) & kIdMask));
    }
}

// launch-shape knobs, overridable from the environment for tuning runs (tools/kprof.py); defaults are the tuned values
int tune(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

template <typename K>
int persistent_grid(K kernel, int num_sms, size_t smem, int threads = kThreads) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (per_sm < 1) per_sm = 1;
    return per_sm * num_sms;
}

// Launch shapes depend on the device (SM count, shared memory) and on the panel count: cached per device, so that
// contexts on different GPUs of one process never share a stale grid (a persistent kernel sized for another device can
// lose its one-wave property; a cooperative one can fail to launch).
constexpr int kMaxDevices = 64;
struct ShapeCache {
    int grid[kMaxDevices][4];
    size_t smem[kMaxDevices][4];
};
int current_device() {
    int d = 0;
    cudaGetDevice(&d);
    return d >= 0 && d < kMaxDevices ? d : 0;
}

}  // namespace

// ================================================================================================ launchers
int launch_source(const SourceDev* frame_dev, unsigned long long npairs, PhantomDev ph, PhotonQueue q0, uint64_t seed,
                  int num_sms, cudaStream_t s) {
    unsigned long long blocks = (npairs + kThreads - 1) / kThreads;
    unsigned long long maxb = (unsigned long long)num_sms * 8;
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    GPET_LAUNCH("k_source", s, k_source<<<(unsigned)blocks, kThreads, 0, s>>>(frame_dev, npairs, ph, q0, seed));
    return 1;
}

int launch_psf_positron(const void* positrons_aos, PhotonQueue q0, unsigned int n_positrons, unsigned long long first,
                        PhantomDev ph, float nonangle, int use_prange, uint64_t seed, int num_sms, cudaStream_t s) {
    unsigned blocks = n_positrons ? (2 * n_positrons + kThreads - 1) / kThreads : 1;
    if (blocks > (unsigned)num_sms * 8) blocks = (unsigned)num_sms * 8;
    GPET_LAUNCH("k_psf_positron", s, k_psf_positron<<<blocks, kThreads, 0, s>>>(static_cast<const gpet_photon*>(positrons_aos), q0,
                                                                              n_positrons, first, ph, nonangle, use_prange, seed));
    return 1;
}

int launch_phantom(PhotonQueue q0, PhotonQueue q1, PhantomDev ph, TablesDev tb, float eabs, uint64_t seed, unsigned long long id_base,
                   int num_sms, cudaStream_t s) {
    static ShapeCache sc{};
    const int dev = current_device();
    if (!sc.grid[dev][0]) sc.grid[dev][0] = persistent_grid(k_phantom, num_sms, 0);
    cudaMemsetAsync(q1.count, 0, sizeof(unsigned), s);
    GPET_LAUNCH("k_phantom", s, k_phantom<<<sc.grid[dev][0], kThreads, 0, s>>>(q0, q1, ph, tb, eabs, seed, id_base));
    return 1;
}

int launch_panel_entry(PhotonQueue q1, PhotonQueue q2, DetectorDev det, unsigned int* counters, int num_sms, cudaStream_t s) {
    const size_t smem_panels = (size_t)det.npanels * sizeof(PanelSm);
    static ShapeCache sc{};
    const int dev = current_device();
    if (!sc.grid[dev][0] || sc.smem[dev][0] != smem_panels) {
        if (smem_panels > 48 * 1024)
            cudaFuncSetAttribute(k_panel_entry, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_panels);
        sc.grid[dev][0] = persistent_grid(k_panel_entry, num_sms, smem_panels);
        sc.smem[dev][0] = smem_panels;
    }
    cudaMemsetAsync(q2.count, 0, sizeof(unsigned), s);
    cudaMemsetAsync(counters + 8, 0, sizeof(unsigned), s);   // photons on a panel
    GPET_LAUNCH("k_panel_entry", s, k_panel_entry<<<sc.grid[dev][0], kThreads, smem_panels, s>>>(q1, det, q2, counters));
    return 1;
}

int launch_front(const SourceDev* frame_dev, unsigned long long npairs, PhotonQueue q0, PhotonQueue q1, PhotonQueue q2,
                 PhantomDev ph, TablesDev tb, DetectorDev det, float eabs, unsigned int* counters, unsigned int* hot, uint64_t seed,
                 unsigned long long id_base, int num_sms, cudaStream_t s, bool reset) {
    // tables in shared memory (A/B, GPET_SMEM_TABLES=1): only when the phantom's materials fit the slots
    static const int want_smem = tune("GPET_SMEM_TABLES", 0);
    const bool smem_tab = want_smem && ph.tab_nstage > 0;
    using FS256 = FrontStageT<kThreads / 32>;
    using FS1024 = FrontStageT<kFrontThreadsSmem / 32>;
    size_t smem = (smem_tab ? sizeof(FS1024) : sizeof(FS256)) + (size_t)det.npanels * sizeof(PanelSm);
    if (smem_tab) smem = ((smem + 15) & ~(size_t)15) + (size_t)ph.tab_nstage * (kSmemMats * sizeof(float4) + sizeof(float));
    else ph.tab_nstage = 0;
    static ShapeCache sc{};
    const int dev = current_device();
    const int v = smem_tab ? 2 : 0;
    if (!sc.grid[dev][v] || sc.smem[dev][v] != smem) {
        if (smem_tab) {
            cudaFuncSetAttribute(k_front<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k_front<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            sc.grid[dev][2] = persistent_grid(k_front<false, true>, num_sms, smem, kFrontThreadsSmem);
            sc.grid[dev][3] = persistent_grid(k_front<true, true>, num_sms, smem, kFrontThreadsSmem);
        } else {
            if (smem > 48 * 1024) {
                cudaFuncSetAttribute(k_front<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cudaFuncSetAttribute(k_front<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            }
            sc.grid[dev][0] = persistent_grid(k_front<false, false>, num_sms, smem);
            sc.grid[dev][1] = persistent_grid(k_front<true, false>, num_sms, smem);
        }
        sc.smem[dev][v] = smem;
    }
    static const int gen_min = tune("GPET_GEN_MIN", 12), entry_min = tune("GPET_ENTRY_MIN", 12);
    unsigned* ticket = hot + kHotTicketFront;
    if (reset) {
        cudaMemsetAsync(q1.count, 0, sizeof(unsigned), s);       // photons that left the phantom (tally only: q1 is not filled)
        cudaMemsetAsync(q2.count, 0, sizeof(unsigned), s);
        cudaMemsetAsync(counters + 8, 0, sizeof(unsigned), s);   // photons on a panel
        cudaMemsetAsync(ticket, 0, sizeof(unsigned), s);
    }
    const int threads = smem_tab ? kFrontThreadsSmem : kThreads;
    if (frame_dev) {
        auto kernel = smem_tab ? k_front<false, true> : k_front<false, false>;
        GPET_LAUNCH("k_front", s, kernel<<<sc.grid[dev][v], threads, smem, s>>>(frame_dev, npairs, q0, ph, tb, det, eabs, seed, q2,
                                                                             q1.count, counters, ticket, gen_min, entry_min, 0ull));
    } else {
        auto kernel = smem_tab ? k_front<true, true> : k_front<true, false>;
        GPET_LAUNCH("k_front<queue>", s, kernel<<<sc.grid[dev][v + 1], threads, smem, s>>>(nullptr, 0ull, q0, ph, tb, det, eabs, seed, q2,
                                                                                         q1.count, counters, ticket, gen_min, entry_min, id_base));
    }
    return 1;
}

template <typename K>
int launch_detector_variant(K kernel, int slot, size_t pool_bytes, PhotonQueue q2, DetectorDev det, TablesDev tb, float eabs, int readout_depth,
                            int readout_policy, int record_hits, HitBuffer hits, EventBuf ev, unsigned int* counters, unsigned* ticket,
                            uint64_t seed, unsigned long long id_base, int num_sms, cudaStream_t s) {
    const size_t smem = pool_bytes + (size_t)det.npanels * sizeof(PanelDev);
    static ShapeCache sc{};
    const int dev = current_device();
    if (!sc.grid[dev][slot] || sc.smem[dev][slot] != smem) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        sc.grid[dev][slot] = persistent_grid(kernel, num_sms, smem, kDetThreads);
        sc.smem[dev][slot] = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)sc.grid[dev][slot]); cfg.blockDim = dim3((unsigned)kDetThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    static const int refill_min = tune("GPET_REFILL_MIN", 4);
    GPET_LAUNCH("k_detector", s, cudaLaunchKernelEx(&cfg, kernel, q2, det, tb, eabs, readout_depth, readout_policy, record_hits, hits, ev,
                                                    counters, ticket, seed, refill_min, id_base));
    return 1;
}

int launch_detector(PhotonQueue q2, DetectorDev det, TablesDev tb, float eabs, int readout_depth, int readout_policy,
                    int record_hits, HitBuffer hits, EventBuf ev, unsigned int* counters, unsigned int* hot, uint64_t seed,
                    unsigned long long id_base, int num_sms, cudaStream_t s, bool reset) {
    unsigned* ticket = hot + kHotTicketDet;
    if (reset) {
        cudaMemsetAsync(hits.count, 0, sizeof(unsigned), s);
        cudaMemsetAsync(ev.count, 0, sizeof(unsigned), s);
        cudaMemsetAsync(counters + 9, 0, sizeof(unsigned), s);     // adder drops
        cudaMemsetAsync(ticket, 0, sizeof(unsigned), s);
    }
    static bool tables_ready[kMaxDevices] = {};
    const int dev = current_device();
    if (!tables_ready[dev]) {
        KnTables h{};
        for (unsigned nC = 1; nC <= 32; nC++) {
            const unsigned R = std::min(32u / nC, (unsigned)kKnMaxRounds);
            for (unsigned l = 0; l < 32; l++) h.jk[nC][l] = (unsigned char)((l % nC) | (std::min(l / nC, 7u) << 5));
            for (unsigned l = 0; l < 32; l++) if (l / nC >= R) h.jk[nC][l] = (unsigned char)((l % nC) | (7u << 5));
            for (unsigned kk = 0; kk < R; kk++) h.rounds[nC] |= 1u << (kk * nC);
        }
        cudaMemcpyToSymbolAsync(c_kn, &h, sizeof(h), 0, cudaMemcpyHostToDevice, s);
        cudaStreamSynchronize(s);   // `h` is a local
        tables_ready[dev] = true;
    }
    constexpr size_t kWarps = kDetThreads / 32;
    return launch_detector_variant(k_detector, 1, sizeof(SlotsSmem) + kWarps * sizeof(WarpStage), q2, det, tb, eabs, readout_depth, readout_policy,
                                   record_hits, hits, ev, counters, ticket, seed, id_base, num_sms, s);
}

int launch_hits_to_rows(HitBuffer hits, unsigned int n, int* id5, float* f5, cudaStream_t s) {
    if (n == 0) return 0;
    const unsigned blocks = std::min((n + kThreads - 1) / kThreads, 4096u);
    GPET_LAUNCH("k_hits_to_rows", s, k_hits_to_rows<<<blocks, kThreads, 0, s>>>(hits, n, id5, f5));
    return 1;
}

int launch_hits_to_aos(HitBuffer hits, unsigned int n, void* aos, cudaStream_t s) {
    if (n == 0) return 0;
    const unsigned blocks = std::min((n + kThreads - 1) / kThreads, 4096u);
    GPET_LAUNCH("k_hits_to_aos", s, k_hits_to_aos<<<blocks, kThreads, 0, s>>>(hits, n, static_cast<gpet_hit*>(aos)));
    return 1;
}

int launch_photons_aos_to_queue(const void* aos, PhotonQueue q, unsigned int n, cudaStream_t s) {
    unsigned blocks = n ? (n + kThreads - 1) / kThreads : 1;
    if (blocks > 4096) blocks = 4096;
    GPET_LAUNCH("k_aos_to_queue", s, k_aos_to_queue<<<blocks, kThreads, 0, s>>>(static_cast<const gpet_photon*>(aos), q, n));
    return 1;
}

int launch_queue_to_photons_aos(PhotonQueue q, void* aos, cudaStream_t s) {
    GPET_LAUNCH("k_queue_to_aos", s, k_queue_to_aos<<<1024, kThreads, 0, s>>>(q, static_cast<gpet_photon*>(aos)));
    return 1;
}

}  // namespace gpet
