// The library context: host-side inputs (parsed files), device buffers, frame plan, results.
// Replaces the reference's host globals (gPETInternal.h:6-63) and static __device__ arrays (gPET_kernals.h:5-73).
#pragma once
#include <string>
#include <vector>

#include "device_types.cuh"
#include "host_io.hpp"
#include "kernels.hpp"
#include "ktimer.hpp"

struct FramePlan {
    double t0_s = 0.0, dt_s = 0.0;             // absolute start (s) and length of the slice
    std::vector<uint64_t> pairs;               // per source
    uint64_t npairs = 0, first_pair = 0;
};

// Growable pinned host buffer: results of gpet_run land here by asynchronous D2H copies.
struct PinnedArena {
    char* p = nullptr;
    size_t cap = 0, size = 0;
};

struct gpet_ctx {
    int device = -1;
    bool has_device = false;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    mutable std::string err;
    uint64_t seed = 0x67504554ull;  // "gPET"

    // ---- host inputs
    gpet::Config cfg;
    gpet::Geometry geo;
    gpet::Isotopes iso;
    gpet::Sources src;
    gpet::Phantom ph;
    gpet::Psf psf;
    gpet::Tables tab;
    bool have_geo = false, have_iso = false, have_src = false, have_ph = false, have_psf = false;
    int usepsf = 0;
    gpet_digitizer_params dig{};
    gpet_transport_params tr{};
    float tstart = 0.f, tend = 1.f;
    std::vector<float> maj_ph, maj_det;
    std::vector<char> ph_present;           // material ids present in the uploaded phantom (shared-memory table staging)
    int rank = 0, world = 1;
    bool emit_on = false;                   // gpet_set_emit_window: digitize a time slice with its halo
    double emit_lo = 0.0, emit_hi = 0.0, emit_trust = 0.0;
    uint64_t first_pair = 0;                // global index of the acquisition's first annihilation pair (gpet_set_first_pair)
    uint64_t id_base = 0;                   // global index of the first photon of the frame now in the queues (photon_index)

    // ---- capacities
    uint64_t cap_photons = 1ull << 22, cap_hits = 1ull << 23, cap_events = 1ull << 22;
    uint64_t max_pairs_per_frame = 0;

    // ---- device state
    bool dev_buffers = false, dev_tables = false, dev_phantom = false, dev_geo = false;
    gpet::PhotonQueue q[3]{};   // after source, after phantom, entered a panel
    gpet::HitBuffer hits{};
    gpet::EventBuf ev{};        // post adder/readout records of the frame ("adder.dat")
    gpet::DigitizerWorkspace ws{};
    void* singles_aos = nullptr;     // = singles_slot[out_slot]: 48-byte records of the frame being digitized
    void* coinc_aos = nullptr;       // = coinc_slot[out_slot]
    unsigned coinc_cap = 0;
    // gpet_run pipelines frames: frame k computes into slot k&1 while the host copies frame k-1 out of the other slot
    void* singles_slot[2] = {nullptr, nullptr};
    void* coinc_slot[2] = {nullptr, nullptr};
    void* pairs_slot[2] = {nullptr, nullptr};            // uint2 index pairs (GPET_COINC_PAIRS)
    void* cls_slot[2] = {nullptr, nullptr};              // one class byte per coincidence (0 true, 1 scatter, 2 random)
    void* cls_aos = nullptr;                             // = cls_slot[out_slot]
    // scatter tags (DetectorDev::scat_tag): table of a power of two >= cap_photons bytes; the serial (1..255) is bumped whenever
    // new photons reach the panel faces or new events are put, so stale tags never match
    unsigned char* d_scat_tag = nullptr;
    unsigned scat_mask = 0, scat_serial = 0;
    unsigned* d_pair_base = nullptr;                     // [2]: singles of the run's earlier frames, alternating by frame
    int psf_output = 0;                                  // OUTPUTPSF of the reference (gpet_set_psf_output)
    int coinc_format = 0;                                // GPET_COINC_RECORDS / GPET_COINC_PAIRS (gpet_run only)
    int singles_format = 0;                              // GPET_SINGLES_RECORDS / GPET_SINGLES_COMPACT (gpet_run(NULL) only)
    bool run_compact = false;                            // the run in flight / the last run delivered 32-byte singles (res_singles holds them)
    void* compact_slot[2] = {nullptr, nullptr};          // 32-byte singles of the frame in each slot (allocated on first use)
    std::vector<char> singles_expanded;                  // 48-byte records built on demand from compact singles
    bool in_run = false;
    bool skip_fallback = false;                          // gpet_run's first attempt: time sort without the LSD fallback kernel
    int64_t run_frame = 0;                               // owned frames launched so far in this run
    bool have_range = false;                             // the time slice of the frame being digitized is known
    gpet::TimeRange range{};
    unsigned* h_slot_counters[2] = {nullptr, nullptr};   // pinned, 32 words each
    cudaEvent_t ev_counters[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    cudaEvent_t ev_emit[2] = {nullptr, nullptr};         // singles of the slot's frame are final (recorded before the coincidence sorter)
    bool early_copy = false;                             // this run starts the singles D2H at ev_emit instead of at the frame's end
    cudaEvent_t ev_run[2] = {nullptr, nullptr};          // start / end of the last run (gpet_stats.ms_total)
    cudaStream_t copy_stream = nullptr;
    int out_slot = 0;
    void* stage_aos = nullptr;       // cap * 48 B staging for AoS <-> SoA conversion
    size_t stage_bytes = 0;
    gpet::PanelDev* d_panels = nullptr;
    unsigned* d_dirmask = nullptr;          // direction table of the panel search (gpet_run only), kDirBins^3 words
    bool dirmask_on = false;                // valid for the run in flight
    double dirmask_key[5] = {0, 0, 0, -1, -1};   // reference point, reference radius, geometry version it was built for
    int geo_version = 0;
    int psf_version = 0;                    // bumped by gpet_load_psf
    mutable double psf_reach_key[4] = {0, 0, 0, -1};   // (o, psf_version) the cached reach below belongs to
    mutable double psf_reach = 0.0;         // largest distance of a PSF record from o
    uint32_t* d_vox = nullptr;
    size_t vox_bytes = 0;                   // size of the voxel grid (access-policy window)
    float4* d_xs = nullptr;
    float *d_maj_ph = nullptr, *d_maj_det = nullptr, *d_cmpsf = nullptr, *d_rayff = nullptr;
    gpet::SourceDev* d_frames = nullptr;
    size_t d_frames_n = 0;
    unsigned long long* d_totals = nullptr;  // accumulated counters of all frames (resident runs)
    unsigned* h_counters = nullptr;          // pinned, 32 words
    unsigned long long* h_totals = nullptr;  // pinned, 32 words
    std::vector<void*> allocs;

    // ---- frames / results
    std::vector<FramePlan> frames;
    bool planned = false;
    PinnedArena res_singles, res_coinc;   // gpet_event / gpet_coincidence records of the last gpet_run
    PinnedArena res_pairs;                // uint32 index pairs of the last gpet_run (GPET_COINC_PAIRS)
    PinnedArena res_cls;                  // class bytes of the last gpet_run's coincidences
    PinnedArena res_adder;                // file runs: post-readout events on their way to adder.dat
    void* file_writer = nullptr;          // the running file run's writer thread (abi.cu FileWriter), else nullptr
    std::vector<char> coinc_expanded;     // records built on demand from res_pairs + res_singles
    bool results_streamed = false;        // the last file run outgrew the arenas: its results are in the files only
    gpet_stats stats{};
    uint64_t last_counts[4] = {0, 0, 0, 0};
    gpet::KernelTimer ktimer;   // per-kernel CUDA-event times (gpet_profile_enable)
    bool profiling = false;
    int64_t psf_first = 0;  // global index of photon 0 of the current PSF batch
};
