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
__device__ __forceinline__ void sample_shape(int shape, const float* __restrict__ c, uint4 r, float& x, float& y, float& z) {
    float u0 = u01(r.x), u1 = u01(r.y), u2 = u01(r.z);
    if (shape < 0 || shape > 2) shape = 0;
    if (shape == 0) {  // box: centre + full lengths
        x = c[0] + c[3] * (-1.f + 2.f * u0) * 0.5f;
        y = c[1] + c[4] * (-1.f + 2.f * u1) * 0.5f;
        z = c[2] + c[5] * (-1.f + 2.f * u2) * 0.5f;
    } else if (shape == 1) {  // cylinder along z: radius c3, height c4
        float phi = kTwoPi * u0;
        float rr = c[3] * sqrtf(u1);
        x = c[0] + rr * cosf(phi);
        y = c[1] + rr * sinf(phi);
        z = c[2] + c[4] * (-1.f + 2.f * u2) * 0.5f;
    } else {  // sphere radius c3
        float phi = kTwoPi * u0;
        float ct = -1.f + 2.f * u1;
        float rr = c[3] * cbrtf(u2);
        float st = sqrtf(1.f - ct * ct);
        x = c[0] + rr * st * cosf(phi);
        y = c[1] + rr * st * sinf(phi);
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

// one thread per photon (two threads per annihilation pair; both recompute the shared pair quantities)
__global__ void __launch_bounds__(kThreads) k_source(const SourceDev* __restrict__ fr, unsigned long long npairs,
                                                     PhantomDev ph, PhotonQueue q0, uint64_t seed) {
    const unsigned long long nph = 2ull * npairs;
    if (blockIdx.x == 0 && threadIdx.x == 0) *q0.count = (unsigned)min(nph, (unsigned long long)q0.capacity);
    for (unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; p < nph && p < q0.capacity;
         p += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long k = p >> 1;
        const int which = (int)(p & 1ull);
        int s = 0;
        while (s < fr->nsource - 1 && k >= fr->cum_pairs[s]) s++;
        const unsigned long long gk = fr->first_pair + k;
        Philox rng(seed, gk, (uint32_t)kStageSource << 24);
        uint4 r0 = rng.next();
        // truncated-exponential decay time inside the frame (statistically identical to the reference's per-atom
        // test ptime = -T_half*1.442695*log(U) < slice, gPET_kernals.cu:519-521)
        double ud = u01d(r0.x, r0.y);
        double ptime = -fr->tau_s[s] * log1p(-ud * fr->frac[s]);
        double t_us = (fr->t0_s + ptime) * 1e6;
        uint4 r1 = rng.next();
        float x, y, z;
        sample_shape(fr->shape[s], fr->coeff + 6 * s, r1, x, y, z);
        // isotropic direction (gPET_kernals.cu:536-540)
        float ct = -1.f + 2.f * u01(r0.z);
        float phi = kTwoPi * u01(r0.w);
        float st = sqrtf(1.f - ct * ct);
        float vx = st * cosf(phi), vy = st * sinf(phi), vz = ct;
        // acollinearity: delta = N(0,1) * sigma (gPET_kernals.cu:549-555)
        uint4 r2 = rng.next();
        float phi2 = kTwoPi * u01(r2.x);
        float g = sqrtf(-2.f * logf(u01(r2.y))) * cosf(kTwoPi * u01(r2.z));
        float delta = g * fr->nonangle;
        if (fr->use_prange) {
            // S4 + S5: positron kinetic energy, then its range (gPET_kernals.cu:529-533); direction is sampled (usedirection 0)
            const float ek = sample_ek_positron(fr->iso_coef + 8 * fr->type[s], rng);
            positron_range(ph, x, y, z, 0.f, 0.f, 0.f, ek, false, rng);
        }
        float E;
        if (which == 0) {
            E = kMC2 + delta * kMC2 * 0.5f;
        } else {
            rotate_dir(vx, vy, vz, -cosf(delta), phi2);
            E = kMC2 - delta * kMC2 * 0.5f;
        }
        q0.pos_e[p] = make_float4(x, y, z, E);
        q0.dir_n[p] = make_float4(vx, vy, vz, __int_as_float(0));
        q0.t[p] = t_us;
        q0.ids[p] = make_int2((int)(unsigned)gk, (int)(unsigned)(2ull * gk + which));
    }
}

// ------------------------------------------------------------------------------------------- P1: phantom transport
__global__ void __launch_bounds__(kThreads) k_phantom(PhotonQueue q0, PhotonQueue q1, PhantomDev ph, TablesDev tb,
                                                      float eabs, uint64_t seed) {
    const unsigned n = min(*q0.count, q0.capacity);
    const unsigned stride = gridDim.x * blockDim.x;
    unsigned next = blockIdx.x * blockDim.x + threadIdx.x;
    bool active = false;
    float x = 0, y = 0, z = 0, E = 0, vx = 0, vy = 0, vz = 0;
    double t = 0;
    int eid = 0, parn = 0, nscat = 0;
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
                x = pe.x; y = pe.y; z = pe.z; E = pe.w;
                vx = dn.x; vy = dn.y; vz = dn.z; nscat = __float_as_int(dn.w);
                t = tt; eid = id.x; parn = id.y;
                rng = Philox(seed, (uint64_t)(uint32_t)parn, (uint32_t)kStagePhantom << 24);
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
            // ---- one Woodcock flight (gPET_kernals.cu:277-296)
            uint4 r = rng.next();
            int ie; float fe;
            energy_index(tb, E, ie, fe);
            float lammin = __fdividef(1.0f, lerp_table(tb.maj_phantom, ie, fe));
            float s = -lammin * __logf(u01(r.x));
            x = fmaf(s, vx, x); y = fmaf(s, vy, y); z = fmaf(s, vz, z);
            t += (double)s * kInvSpeedOfLight;
            int ix = (int)((x - ph.ox) * ph.idx), iy = (int)((y - ph.oy) * ph.idy), iz = (int)((z - ph.oz) * ph.idz);
            if (ix <= 0 || ix >= ph.nx || iy <= 0 || iy >= ph.ny || iz <= 0 || iz >= ph.nz) {
                done_alive = true;  // escaped: keeps the overshoot position
            } else {
                uint32_t vw = __ldg(ph.vox + ((size_t)iz * ph.ny + iy) * ph.nx + ix);
                int mat = (int)(vw & 15u);
                float rho = __uint_as_float(vw & ~15u);
                Xs3 xs = lerp_xs(tb, mat, ie, fe);
                float lamden = lammin * rho;
                float prob = 1.0f - lamden * xs.tot;
                float u = u01(r.y);
                if (u >= prob) {
                    prob += lamden * xs.compt;
                    if (u < prob) {
                        // Compton with binding effects: cos(theta) from the cmpsf surface (gPET_kernals.cu:66-88)
                        float costh = surface_lookup(tb.cmpsf, mat, tb.cm_ncp, tb.cm_ne, E * tb.cm_ide, u01(r.z) * tb.cm_idcp);
                        float efrac = 1.0f / (1.0f + E * kIMC2 * (1.0f - costh));
                        float phi = kTwoPi * u01(r.w);
                        E *= efrac;
                        nscat++;
                        if (E < eabs) done_alive = true;  // still handed to the detector stage
                        else rotate_dir(vx, vy, vz, costh, phi);
                    } else {
                        prob += lamden * xs.rayl;
                        if (u < prob) {
                            float costh = surface_lookup(tb.rayff, mat, tb.rl_ncp, tb.rl_ne, E * tb.rl_ide, u01(r.z) * tb.rl_idcp);
                            float phi = kTwoPi * u01(r.w);
                            nscat++;
                            rotate_dir(vx, vy, vz, costh, phi);
                        } else {
                            active = false;  // photoelectric absorption: history ends (tof = -0.5 in the reference)
                        }
                    }
                }
            }
        }
        // ---- warp-aggregated append of the photons that left the phantom alive
        unsigned slot = warp_reserve(q1.count, done_alive ? 1u : 0u);
        if (done_alive) {
            if (slot < q1.capacity) {
                q1.pos_e[slot] = make_float4(x, y, z, E);
                q1.dir_n[slot] = make_float4(vx, vy, vz, __int_as_float(nscat));
                q1.t[slot] = t;
                q1.ids[slot] = make_int2(eid, parn);
            }
            active = false;
        }
    }
}

// ------------------------------------------------------------------------------------------- X1/X2/D1/D2: detector
constexpr int kSlots = 6;  // distinct crystals per photon kept by the adder (reference: Event events[4], no bound check)

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

// Panel entry (gPET_kernals.cu:963-1009): one thread per photon that left the phantom, convergent loop over the
// panels; photons whose straight line crosses a panel's front face are appended -- already in that panel's local
// frame, with the time of flight to the face added -- to the compact queue the transport kernel works on.
// dir_n.w of the output carries the panel index.
__global__ void __launch_bounds__(kThreads) k_panel_entry(PhotonQueue q1, DetectorDev det, PhotonQueue q2,
                                                          unsigned* __restrict__ counters) {
    extern __shared__ PanelDev s_panels[];
    for (int i = threadIdx.x; i < det.npanels * (int)(sizeof(PanelDev) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(s_panels)[i] = reinterpret_cast<const uint32_t*>(det.panels)[i];
    __syncthreads();
    const unsigned n = min(*q1.count, q1.capacity);
    const unsigned nround = (n + 31u) & ~31u;  // whole warps stay in the loop for the collective append
    unsigned n_on_panel = 0;
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < nround; p += gridDim.x * blockDim.x) {
        bool ok = false;
        float4 pe = make_float4(0, 0, 0, 0), ov = make_float4(0, 0, 0, 0);
        double t = 0.0;
        int2 id = make_int2(0, 0);
        if (p < n) {
            pe = q1.pos_e[p];
            float4 dn = q1.dir_n[p];
            const double tt = q1.t[p];
            id = q1.ids[p];
            if (tt > 0.0) {
                for (int i = 0; i < det.npanels; i++) {
                    const PanelDev& pd = s_panels[i];
                    float rx = pe.x - pd.ox, ry = pe.y - pd.oy, rz = pe.z - pd.oz;
                    float lx = rx * pd.uxx + ry * pd.uxy + rz * pd.uxz;
                    float ly = rx * pd.uyx + ry * pd.uyy + rz * pd.uyz;
                    float lz = rx * pd.uzx + ry * pd.uzy + rz * pd.uzz;
                    float lvx = dn.x * pd.uxx + dn.y * pd.uxy + dn.z * pd.uxz;
                    float lvy = dn.x * pd.uyx + dn.y * pd.uyy + dn.z * pd.uyz;
                    float lvz = dn.x * pd.uzx + dn.y * pd.uzy + dn.z * pd.uzz;
                    if (lvx * pd.dirx >= 0.f) {
                        float q = __fdiv_rn(lx, lvx);
                        float y2 = ly - q * lvy, z2 = lz - q * lvz;
                        if (fabsf(y2) < pd.ly / 2 && fabsf(z2) < pd.lz / 2) {
                            pe = make_float4(0.f, y2, z2, pe.w);
                            ov = make_float4(lvx, lvy, lvz, __int_as_float(i));
                            t = tt + (-(double)lx / (kSpeedOfLight * (double)lvx));
                            ok = true;
                            break;
                        }
                    }
                }
            }
        }
        unsigned slot = warp_reserve(q2.count, ok ? 1u : 0u);
        if (ok) {
            n_on_panel++;
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

// Per-thread adder slots in shared memory, [slot][thread] so that a warp's accesses are conflict free.
struct SlotsSmem {
    int site[kSlots][kThreads];
    float E[kSlots][kThreads], x[kSlots][kThreads], y[kSlots][kThreads], z[kSlots][kThreads];
    double t[kSlots][kThreads];
};

// D1 adder (gPET_kernals.cu:737-755): merge hits of the same crystal; energy-weighted centroid with the
// contraction spelled out: (x_i*E_i + x*E)/(E_i+E) = fma(x_i, E_i, x*E) / (E_i + E)
__device__ __forceinline__ bool adder(SlotsSmem& sl, int& n, int site, float E, float x, float y, float z, double t) {
    const int tid = threadIdx.x;
    for (int k = 0; k < n; k++) {
        if (sl.site[k][tid] == site) {
            const float ek = sl.E[k][tid];
            const float es = __fadd_rn(ek, E);
            sl.x[k][tid] = __fdiv_rn(__fmaf_rn(sl.x[k][tid], ek, __fmul_rn(x, E)), es);
            sl.y[k][tid] = __fdiv_rn(__fmaf_rn(sl.y[k][tid], ek, __fmul_rn(y, E)), es);
            sl.z[k][tid] = __fdiv_rn(__fmaf_rn(sl.z[k][tid], ek, __fmul_rn(z, E)), es);
            sl.E[k][tid] = es;
            return true;
        }
    }
    if (n >= kSlots) return false;
    sl.site[n][tid] = site; sl.E[n][tid] = E; sl.x[n][tid] = x; sl.y[n][tid] = y; sl.z[n][tid] = z; sl.t[n][tid] = t;
    n++;
    return true;
}

// Photon transport inside a panel (gPET_kernals.cu:1018-1192) over the compact panel-entry queue, persistent warps with
// lane refill; adder on the fly, readout (gPET_kernals.cu:756-813) when the photon is finished.
__global__ void __launch_bounds__(kThreads, 3) k_detector(PhotonQueue q2, DetectorDev det, TablesDev tb, float eabs,
                                                          int rdepth, int rpolicy, int record_hits, HitBuffer hits, EventSoA ev,
                                                          unsigned* __restrict__ counters, uint64_t seed) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    SlotsSmem& sl = *reinterpret_cast<SlotsSmem*>(s_raw);
    PanelDev* s_panels = reinterpret_cast<PanelDev*>(s_raw + sizeof(SlotsSmem));
    for (int i = threadIdx.x; i < det.npanels * (int)(sizeof(PanelDev) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(s_panels)[i] = reinterpret_cast<const uint32_t*>(det.panels)[i];
    __syncthreads();

    const int tid = threadIdx.x;
    const unsigned n = min(*q2.count, q2.capacity);
    const unsigned stride = gridDim.x * blockDim.x;
    unsigned next = blockIdx.x * blockDim.x + threadIdx.x;
    const int crysPerPanel = det.moduleN * det.crystalN;
    const int depth = (rdepth != 3 && rpolicy == 1) ? 2 : rdepth;
    bool active = false;
    float x = 0, y = 0, z = 0, E = 0, vx = 0, vy = 0, vz = 0;
    double t = 0;
    int eid = 0, parn = 0, pa = 0, nslot = 0;
    unsigned n_drop_adder = 0;
    Philox rng(seed, 0, 0);
    while (true) {
        // ---- refill from the compact queue: cheap, so idle lanes are topped up as soon as a quarter of the warp idles
        unsigned amask = __ballot_sync(kFull, active);
        unsigned wmask = __ballot_sync(kFull, !active && next < n);
        if (wmask && (__popc(wmask) >= 8 || amask == 0)) {
            if (!active && next < n) {
                float4 pe = q2.pos_e[next];
                float4 dn = q2.dir_n[next];
                t = q2.t[next];
                int2 id = q2.ids[next];
                next += stride;
                x = pe.x; y = pe.y; z = pe.z; E = pe.w;
                vx = dn.x; vy = dn.y; vz = dn.z; pa = __float_as_int(dn.w);
                eid = id.x; parn = id.y;
                rng = Philox(seed, (uint64_t)(uint32_t)parn, (uint32_t)kStageDetector << 24);
                nslot = 0;
                active = true;
            }
            amask = __ballot_sync(kFull, active);
        }
        if (amask == 0) break;  // nothing active and nothing left to pull for any lane of this warp
        // up to two hits per flight (Compton deposit + absorption of the remainder), both at the same point
        int nh = 0, h_mod = -1, h_cry = -1, h_type0 = 0;
        float h_E0 = 0.f, h_E1 = 0.f;
        bool finished = false;
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
                int m_id, M_id, L_id;
                crystal_search(pd, det, x, y, z, m_id, M_id, L_id);
                float rho = det.dens[m_id];
                int mat = det.mat[m_id];
                Xs3 xs = lerp_xs(tb, mat, ie, fe);
                float lamden = lammin * rho;
                float prob = fmaxf(1.0f - lamden * xs.tot, 0.f);
                float u = u01(r.y);
                if (u >= prob) {
                    prob += lamden * xs.compt;
                    if (u < prob) {
                        float efrac, costh;
                        compton_kn(E, rng, efrac, costh);
                        float de = E * (1.0f - efrac);
                        float phi = kTwoPi * u01(r.z);
                        if (m_id == 0) { h_mod = M_id; h_cry = L_id; h_type0 = 1; h_E0 = de; nh = 1; }
                        E -= de;
                        if (E < eabs) {
                            if (m_id == 0) { h_E1 = E; nh = 2; }  // type 2: remainder absorbed on the spot
                            finished = true;
                        } else {
                            rotate_dir(vx, vy, vz, costh, phi);
                        }
                    } else {
                        prob += lamden * xs.rayl;
                        if (u < prob) {
                            float costh = surface_lookup(tb.rayff, mat, tb.rl_ncp, tb.rl_ne, E * tb.rl_ide, u01(r.z) * tb.rl_idcp);
                            float phi = kTwoPi * u01(r.w);
                            rotate_dir(vx, vy, vz, costh, phi);
                        } else {
                            if (m_id == 0) { h_mod = M_id; h_cry = L_id; h_type0 = 4; h_E0 = E; nh = 1; }
                            finished = true;
                        }
                    }
                }
            }
            // adder on the fly
            if (nh >= 1) {
                int site = pa * crysPerPanel + h_mod * det.crystalN + h_cry;
                if (!adder(sl, nslot, site, h_E0, x, y, z, t)) n_drop_adder++;
                if (nh == 2 && !adder(sl, nslot, site, h_E1, x, y, z, t)) n_drop_adder++;
            }
        }
        // ---- hits: rows in file layout, warp-aggregated
        if (record_hits) {
            unsigned hmask = __ballot_sync(kFull, nh > 0);
            if (hmask) {
                unsigned slot = warp_reserve(hits.count, (unsigned)nh);
                for (int k = 0; k < nh; k++) {
                    if (slot + k < hits.capacity) {
                        int* hi = hits.id + 5ull * (slot + k);
                        float* hf = hits.f + 5ull * (slot + k);
                        hi[0] = parn; hi[1] = s_panels[pa].id; hi[2] = h_mod; hi[3] = h_cry; hi[4] = k ? 2 : h_type0;
                        hf[0] = k ? h_E1 : h_E0; hf[1] = (float)t; hf[2] = x; hf[3] = y; hf[4] = z;
                        hits.t[slot + k] = t;
                    }
                }
            }
        }
        // ---- photon finished: readout (gPET_kernals.cu:756-813) and event append
        const bool mine = finished && nslot > 0;
        unsigned fmask = __ballot_sync(kFull, mine);
        if (finished) active = false;
        if (fmask) {
            int cnt = 0;
            const int panel_id = mine ? s_panels[pa].id : 0;
            unsigned deadmask = 0;  // bit k: slot k merged away
            if (mine) {
                // rewrite the slot keys at readout level in place (site -> key), keeping the crystal site in a register copy
                if (rdepth != 3) {
                    for (int i = 0; i < nslot; i++) {
                        if (deadmask >> i & 1u) continue;
                        const int csi = sl.site[i][tid] - pa * crysPerPanel;
                        const int keyi = depth == 0 ? 0 : depth == 1 ? panel_id : depth == 2 ? panel_id * det.moduleN + csi / det.crystalN
                                                                                              : panel_id * crysPerPanel + csi;
                        for (int j = i + 1; j < nslot; j++) {
                            if (deadmask >> j & 1u) continue;
                            const int csj = sl.site[j][tid] - pa * crysPerPanel;
                            const int keyj = depth == 0 ? 0 : depth == 1 ? panel_id : depth == 2 ? panel_id * det.moduleN + csj / det.crystalN
                                                                                                  : panel_id * crysPerPanel + csj;
                            if (keyj != keyi) continue;
                            const float Ei = sl.E[i][tid], Ej = sl.E[j][tid];
                            if (rpolicy == 1) {
                                const float es = __fadd_rn(Ei, Ej);
                                sl.x[i][tid] = __fdiv_rn(__fmaf_rn(sl.x[i][tid], Ei, __fmul_rn(sl.x[j][tid], Ej)), es);
                                sl.y[i][tid] = __fdiv_rn(__fmaf_rn(sl.y[i][tid], Ei, __fmul_rn(sl.y[j][tid], Ej)), es);
                                sl.z[i][tid] = __fdiv_rn(__fmaf_rn(sl.z[i][tid], Ei, __fmul_rn(sl.z[j][tid], Ej)), es);
                                sl.E[i][tid] = es;
                            } else if (!(Ei > Ej)) {
                                // winner-take-all: the larger energy wins the whole record (ties -> the later one)
                                sl.site[i][tid] = sl.site[j][tid]; sl.E[i][tid] = Ej; sl.x[i][tid] = sl.x[j][tid];
                                sl.y[i][tid] = sl.y[j][tid]; sl.z[i][tid] = sl.z[j][tid]; sl.t[i][tid] = sl.t[j][tid];
                            }
                            deadmask |= 1u << j;
                        }
                    }
                }
                cnt = nslot - __popc(deadmask);
            }
            unsigned slot = warp_reserve(ev.count, (unsigned)cnt);
            if (mine) {
                for (int k = 0; k < nslot; k++) {
                    if (deadmask >> k & 1u) continue;
                    if (slot < ev.capacity) {
                        const int cs = sl.site[k][tid] - pa * crysPerPanel;
                        const int mod = cs / det.crystalN;
                        ev.parn[slot] = parn; ev.pann[slot] = panel_id; ev.modn[slot] = mod;
                        ev.cryn[slot] = cs - mod * det.crystalN;
                        ev.siten[slot] = depth == 0 ? 0 : depth == 1 ? panel_id : depth == 2 ? panel_id * det.moduleN + mod
                                                                                            : panel_id * crysPerPanel + cs;
                        ev.eventid[slot] = eid;
                        ev.t[slot] = sl.t[k][tid]; ev.E[slot] = sl.E[k][tid];
                        ev.x[slot] = sl.x[k][tid]; ev.y[slot] = sl.y[k][tid]; ev.z[slot] = sl.z[k][tid];
                    }
                    slot++;
                }
                nslot = 0;
            }
        }
    }
    // per-warp tallies
    unsigned b = n_drop_adder;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(kFull, b, o);
    if (lane_id() == 0 && b) atomicAdd(&counters[9], b);
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
        q0.ids[p] = make_int2((int)(unsigned)gi, (int)(unsigned)(2ull * gi + which));
    }
}

template <typename K>
int persistent_grid(K kernel, int num_sms, size_t smem) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem);
    if (per_sm < 1) per_sm = 1;
    return per_sm * num_sms;
}

}  // namespace

// ================================================================================================ launchers
int launch_source(const SourceDev* frame_dev, unsigned long long npairs, PhantomDev ph, PhotonQueue q0, uint64_t seed,
                  int num_sms, cudaStream_t s) {
    unsigned long long nph = 2ull * npairs;
    unsigned long long blocks = (nph + kThreads - 1) / kThreads;
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

int launch_phantom(PhotonQueue q0, PhotonQueue q1, PhantomDev ph, TablesDev tb, float eabs, uint64_t seed, int num_sms,
                   cudaStream_t s) {
    static int grid = 0;
    if (!grid) grid = persistent_grid(k_phantom, num_sms, 0);
    cudaMemsetAsync(q1.count, 0, sizeof(unsigned), s);
    GPET_LAUNCH("k_phantom", s, k_phantom<<<grid, kThreads, 0, s>>>(q0, q1, ph, tb, eabs, seed));
    return 1;
}

int launch_detector(PhotonQueue q1, PhotonQueue q2, DetectorDev det, TablesDev tb, float eabs, int readout_depth,
                    int readout_policy, int record_hits, HitBuffer hits, EventSoA ev, unsigned int* counters, uint64_t seed,
                    int num_sms, cudaStream_t s) {
    const size_t smem_panels = (size_t)det.npanels * sizeof(PanelDev);
    const size_t smem = sizeof(SlotsSmem) + smem_panels;
    static int grid = 0, grid_entry = 0;
    static size_t grid_smem = 0;
    if (!grid || grid_smem != smem) {
        cudaFuncSetAttribute(k_detector, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (smem_panels > 48 * 1024)
            cudaFuncSetAttribute(k_panel_entry, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_panels);
        grid = persistent_grid(k_detector, num_sms, smem);
        grid_entry = persistent_grid(k_panel_entry, num_sms, smem_panels);
        grid_smem = smem;
    }
    // q2.count, hits.count, ev.count and the two tallies are adjacent words of the counter block? no: reset one by one
    cudaMemsetAsync(q2.count, 0, sizeof(unsigned), s);
    cudaMemsetAsync(hits.count, 0, 2 * sizeof(unsigned), s);   // hits.count, ev.count (adjacent words of the counter block)
    cudaMemsetAsync(counters + 8, 0, 2 * sizeof(unsigned), s);
    GPET_LAUNCH("k_panel_entry", s, k_panel_entry<<<grid_entry, kThreads, smem_panels, s>>>(q1, det, q2, counters));
    GPET_LAUNCH("k_detector", s, k_detector<<<grid, kThreads, smem, s>>>(q2, det, tb, eabs, readout_depth, readout_policy, record_hits, hits, ev, counters,
                                           seed));
    return 2;
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
