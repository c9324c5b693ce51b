// Device-side data layout shared by the kernels and the ABI layer.  All buffers are SoA in HBM.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gpet {

constexpr double kSpeedOfLight = 29979.2458;          // cm/us (gPET_kernals.cu:264)
constexpr double kInvSpeedOfLight = 1.0 / 29979.2458;
constexpr float kMC2 = 510.9991e3f;                   // constants.h:33
constexpr float kIMC2 = 1.95695060911e-6f;            // constants.h:34
constexpr float kTwoPi = 6.2831853071795864769252867f;
constexpr double kMaxT = 1e20;                        // constants.h:18 MAXT

// History numbers.  The records of the reference carry 32-bit ids (gPET.h:50, 87-92); here the global photon index is
// 64 bits wide (1e10 decays = 2e10 photons), the record keeps its low 31 bits (never negative: parn == -1 stays the mark of
// a noise single, eventid's top bit the mark of a noise event) and every Philox stream is keyed by the full index,
// rebuilt from the record and the index of the first photon of the frame (a frame holds far fewer than 2^31 photons):
//   index = base + ((parn - base) mod 2^31).
// base == 0 (stage-level entry points, replayed lists, the first 2^31 photons of a run): the 32-bit id as it is.
constexpr unsigned kIdMask = 0x7fffffffu;
__host__ __device__ __forceinline__ unsigned long long photon_index(int parn, unsigned long long base) {
    if (base == 0ull) return (unsigned long long)(unsigned)parn;
    return base + (unsigned long long)(((unsigned)parn - (unsigned)base) & kIdMask);
}

// Philox stream ids (counter word 2, high byte)
enum Stage : uint32_t {
    kStageSource = 1, kStagePhantom = 2, kStageDetector = 3, kStageBlur = 4, kStagePlan = 5, kStagePsfPositron = 6, kStageNoise = 7
};

// Photon phase-space queue: 48 B per photon, 16-byte vector accesses.
struct PhotonQueue {
    float4* pos_e;        // x, y, z, E
    float4* dir_n;        // vx, vy, vz, bit-cast nscat
    double* t;            // us; <= 0 dead
    int2* ids;            // eventid, parn
    unsigned int* count;  // device counter
    unsigned int capacity;
};

// Hits as structure-of-arrays of 16-byte vectors: a warp's staged hits leave the SM as full-line vector stores (one int4,
// one float4, one double and one int per hit and lane, consecutive hits in consecutive lanes).  The file layout of the
// reference (HitsID.dat: 5 x int32 rows, Hits.dat: 5 x float32 rows, gPET.cu:367-376) is produced from these on demand
// by k_hits_to_rows / k_hits_to_aos (file dumps and gpet_fetch_hits), never on the transport path.
struct HitBuffer {
    int4* id4;            // parn, pann, modn, cryn
    float4* f4;           // E, x, y, z (panel-local)
    double* t;            // fp64 time (the file keeps its float32 image)
    int* type;            // 1 Compton deposit, 2 absorbed remainder, 4 photoelectric
    unsigned int* count;
    unsigned int capacity;
};

// Post-readout events / singles: 48-byte records in the reference's file layout (Event, gPET.h:87-92), 16-byte
// aligned, so a record moves as three 16-byte vectors and adder.dat / singles.dat are plain copies of the buffers.
struct EventRec {
    int parn, pann, modn, cryn, siten, eventid;
    double t;
    float E, x, y, z;
};
static_assert(sizeof(EventRec) == 48, "EventRec must match gpet_event");

struct EventBuf {
    EventRec* rec;
    unsigned int* count;
    unsigned int capacity;
};

__device__ __forceinline__ void store_event_rec(EventRec* dst, const EventRec& r) {
    const long long tb = __double_as_longlong(r.t);
    int4* p = reinterpret_cast<int4*>(dst);
    p[0] = make_int4(r.parn, r.pann, r.modn, r.cryn);
    p[1] = make_int4(r.siten, r.eventid, (int)(unsigned)(tb & 0xffffffffll), (int)(unsigned)((unsigned long long)tb >> 32));
    p[2] = make_int4(__float_as_int(r.E), __float_as_int(r.x), __float_as_int(r.y), __float_as_int(r.z));
}

__device__ __forceinline__ EventRec load_event_rec(const EventRec* src) {
    const int4* p = reinterpret_cast<const int4*>(src);
    const int4 a = p[0], b = p[1], c = p[2];
    EventRec r;
    r.parn = a.x; r.pann = a.y; r.modn = a.z; r.cryn = a.w;
    r.siten = b.x; r.eventid = b.y;
    r.t = __longlong_as_double((long long)(((unsigned long long)(unsigned)b.w << 32) | (unsigned)b.z));
    r.E = __int_as_float(c.x); r.x = __int_as_float(c.y); r.y = __int_as_float(c.z); r.z = __int_as_float(c.w);
    return r;
}

struct PanelDev {  // one panel, 128 B
    float ox, oy, oz;          // offset (front-face centre)
    float uxx, uxy, uxz;       // local x axis in global frame
    float uyx, uyy, uyz;
    float uzx, uzy, uzz;
    float lx, ly, lz;          // panel dimensions
    float dirx;                // expansion direction along local x (+-1)
    float mody, modz;          // module size
    float mspy, mspz;          // module gap
    float lsoy, lsoz;          // crystal size
    float spy, spz;            // crystal gap
    int id;
    float r2;                  // (ly/2)^2 + (lz/2)^2: squared radius of the sphere around the face centre that holds the face
    // correctly rounded reciprocals of the four divisors of crystalSearch: 1/(mody+mspy), 1/(modz+mspz), 1/(lsoy+spy), 1/(lsoz+spz)
    float rcp_my, rcp_mz, rcp_cy, rcp_cz;
    float pad[2];
};
static_assert(sizeof(PanelDev) == 128, "PanelDev must be 128 B");

struct DetectorDev {
    const PanelDev* panels;
    int npanels;
    int moduleNy, crystalNy, moduleN, crystalN;
    int mat[2];
    float dens[2];
    int nsurface;
    float surface[50];
    int prefilter;             // 1: every panel's local axes are orthonormal, so the bounding-sphere rejection of panel_entry is valid
    // Direction table (gpet_run only, nullptr otherwise): kDirBins^3 cells over the direction cube [-1,1]^3, one bit per
    // panel that a photon flying in a direction of that cell can possibly enter, GIVEN that its line passes the
    // reference sphere the table was built for (phantom box + source shapes / PSF points).  Conservative.
    const unsigned* dirmask;
    // Scatter tags (coincidence classification): a photon that enters a panel after at least one
    // Compton or Rayleigh interaction in the phantom stores the serial number of its frame at scat_tag[parn & scat_mask];
    // the coincidence sorter finds it there by the single's parn.  A frame's photon numbers are contiguous and the table
    // holds at least a frame's photons, so slots are unique within a frame.  One BYTE per photon: the serial runs 1..255
    // and the table is cleared when it wraps, i.e. every 255 frames (with 4-byte tags the table was 16 MB and the sorter's
    // 0.7 M random lookups per frame came from DRAM, 23 MB against 8.5 MB without them; 4 MB stays in L2).  nullptr: off.
    unsigned char* scat_tag;
    unsigned scat_mask, scat_serial;
};
constexpr int kDirBins = 32;

struct PhantomDev {
    const uint32_t* vox;   // packed: fp32 density with the material id in the 4 low mantissa bits
    int nx, ny, nz;
    float ox, oy, oz;      // offset
    float idx, idy, idz;   // 1/voxel size
    float dx, dy, dz;      // voxel size
    int rec_on;            // 1: photons leaving the phantom are moved onto the PSF-recording sphere (RECORDPSF == -1, gPET_kernals.cu:288-294)
    float rec[4];          // sphere centre x, y, z and radius (input_PET.in field 14)
    // scatter tags, set for the fused front end only (see DetectorDev::scat_tag; the staged path tags at panel entry)
    unsigned char* scat_tag;
    unsigned scat_mask, scat_serial;
    // shared-memory staging of the 1-D tables in the fused front end (transport.cu TabShared): energy nodes staged (0: off) and
    // 4 bits per material id with its slot (15: not staged); set by the host from the materials present in the phantom
    int tab_nstage;
    unsigned long long tab_slot_map;
};

struct TablesDev {
    // float4 per (material, energy node): Sigma_tot, Sigma_compton, Sigma_rayleigh, Sigma_photo (cm^2/g)
    const float4* xs;
    const float* maj_phantom;   // Sigma_max(E) 1/cm on the same grid
    const float* maj_detector;
    const float* cmpsf;         // [mat][icp][ie]
    const float* rayff;
    int nmat, nen;
    float e0, ide;              // index = ide * (E - e0)
    int cm_ncp, cm_ne, rl_ncp, rl_ne;
    float cm_idcp, cm_ide, rl_idcp, rl_ide;
};

struct SourceDev {   // per-frame source description (<= GPET_MAX_SOURCES entries)
    int nsource;
    unsigned long long cum_pairs[64];  // inclusive prefix of pairs per source in this frame
    int type[64], shape[64];
    float coeff[64 * 6];
    double tau_s[64];       // mean life (s) = T_half * 1.442695 as in gPET_kernals.cu:519
    double frac[64];        // 1 - exp(-dt/tau): truncated-exponential normaliser for this frame
    float iso_coef[16 * 8];
    double t0_s;            // frame start (s, absolute acquisition time)
    unsigned long long first_pair;  // global index of pair 0 of this frame
    float nonangle;
    int use_prange;
};

struct DigitizerDev {
    int readout_depth, readout_policy;
    float Eth;
    int blur_policy; float Eref, Rref, slope, sblur;
    int dlevel, dtype; float dtime;
    float Ewinmin, Ewinmax;
    float tblur; float cwin; int cpolicy; int cmindiff;
    float noise_gap, noise_Emean, noise_sigma, noise_interval;   // addnoise (gPET_kernals.cu:699-735); gap <= 0: off
    int npanels;
    int moduleN, crystalN;
    unsigned long long id_base;   // global index of the frame's first photon (photon_index); 0 for replayed lists
    // Emit window (gpet_set_emit_window; multi-GPU exchange by time slice): the list holds this rank's slice [emit_lo, emit_hi)
    // plus a halo of the neighbouring slices.  Everything is digitized as one list; singles are all written (the sorter
    // needs the halo), counters[10] / [11] count those before / inside the window, and only coincidences whose opening
    // single lies inside the window are emitted.  trust_lo = start of the halo + max(dead time, coincidence window): a
    // decision for an event of the window that depends on anything earlier raises counters[15] (halo too short).
    int emit_on;
    double emit_lo, emit_hi, trust_lo;
    // coincidence classes (k_coinc): same annihilation iff eventid >> pair_shift agree; scatter tags as in DetectorDev
    int pair_shift;
    // equal times: 0 keep the input order (replayed lists: SURVEY quirk 11 as specified), 1 site number first, then input
    // order (events produced by this run's detector kernel, whose order in the buffer is not defined)
    int tie_site;
    const unsigned char* scat_tag;
    unsigned scat_mask, scat_serial;
};

}  // namespace gpet
