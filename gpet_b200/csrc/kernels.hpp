// Host-callable launchers of the sm_100a kernels.  Every launcher is asynchronous on `stream` and returns the
// number of kernel launches it issued (for gpet_stats.kernel_launches).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "device_types.cuh"

namespace gpet {

namespace rsort { struct SortState; }

// Time range (us) the time sort slices.  `dev` != nullptr: the range is read from device memory instead, as the u64
// key images of its ends (dev[0], dev[1]).  The range only shapes the load of the sort, never its result.
struct TimeRange {
    double lo, hi;
    const unsigned long long* dev;
};
TimeRange time_range_us(double t_lo_us, double t_hi_us);

struct DigitizerWorkspace {
    unsigned long long* tkeys[2];  // [0] time keys by event index, then the sorted keys; [1] scatter target / LSD ping-pong
    unsigned int* tvals[2];        // LSD fallback payload ping-pong (event index | window flag)
    int* site_of;                  // dead-time site number by event index
    unsigned int* aux;             // by event index: arrival rank in its time slice | window flag (bit 31)
    uint4* bent;                   // scatter target, one 16-byte entry per event: time key (lo, hi), event index | window flag, site
    unsigned int* lookback[2];     // sort_lookback_words(capacity) status words each (radix passes alternate between them)
    rsort::SortState* st_time;     // LSD fallback bookkeeping (histograms, tile counters)
    unsigned int* grid_bar;        // grid barrier of the cooperative fallback kernel (2 words, zeroed once)
    unsigned int* scan_status[3];  // status words of the chained scans: singles compaction, coincidence compaction (max_tiles
                                   // each), slice counters of the bucket sort (bucket_words() / 2048)
    unsigned int* bcount;          // bucket sort of the time keys: events per slice, slice starts (bucket_words() entries each)
    unsigned int* bstart;
    unsigned long long* minmax;    // [0] ~(smallest), [1] largest time key (k_range, replay entry); zero = empty
    unsigned int* frame_state;     // counters | scan_status[0..2] | minmax | bcount as ONE block: a frame starts by zeroing it
    size_t frame_state_bytes;
    unsigned int max_tiles;
    unsigned int capacity;
    unsigned int* order_t;         // event index | window flag, in time order (first counters[1] entries)
    int* site_t;                   // site number in time order
    unsigned char* kill;           // per time-order position: 1 = removed by dead time (non-paralyzable chain)
    double* stime;                 // per single: time, panel (what the coincidence sorter reads)
    int* span;
    int* spar;                     // per single: photon number and annihilation number (coincidence classes)
    int* seid;
    // [0] n_in [1] after thresholder [2] after deadtime [3] singles [4] coincidences [6],[7] tile tickets of the two
    // compactions [5] time sort fell back to LSD radix [8] photons on a panel [9] adder drops [12..14] true / scatter /
    // random coincidences; [16..19] queue 0, queue 1, hits, events counts, [21] queue 2
    unsigned int* counters;
    // Hot counters, one per 4 KB of their own: warp-aggregated atomics on ONE 128-byte line serialise at 0.67 ns each
    // whatever word they hit (tools/microbench/latency.cu), which bounded k_front and k_detector while tickets and queue
    // counts shared the counter block.  Word offsets below; part of frame_state (zeroed by the frame's memset).
    unsigned int* hot;
    // energy spectrum of the singles; bin b lives at spectrum[b * spectrum_stride] (one bin per 128-byte line for small
    // histograms: every block adds its private histogram at the end, and atomics on one line serialise)
    unsigned long long* spectrum; int spectrum_bins; int spectrum_stride; float spec_emin, spec_emax;
};

// Where one frame's results go (device memory).
struct DigitizerOut {
    void* singles;                 // 48-byte records, time sorted
    unsigned int singles_cap;
    void* singles_compact;         // optional: the same list as 32-byte gpet_single_compact records (k_pack_singles)
    void* coinc;                   // 96-byte coincidence records, or nullptr
    void* pairs;                   // uint2 index pairs into the run's singles list, or nullptr
    unsigned int coinc_cap;
    void* cls;                     // one class byte per coincidence (0 true, 1 scatter, 2 random), or nullptr
    const unsigned int* pair_base_in;   // singles of the run's earlier frames (device word), nullptr = 0
    unsigned int* pair_base_out;        // receives *pair_base_in + this frame's singles, or nullptr
    // optional: right after k_emit_singles the singles count is copied to this pinned host word and the event is
    // recorded, so that the host can start the D2H copy of the singles while the coincidence sorter still runs
    unsigned int* h_singles_count;
    cudaEvent_t ev_after_emit;
};

enum HotWord : unsigned {
    kHotStride = 1024,                 // words between hot counters
    kHotTicketFront = 0 * kHotStride,  // k_front's pair / photon ticket
    kHotQ2 = 1 * kHotStride,           // photons on a panel (queue 2 count)
    kHotTicketDet = 2 * kHotStride,    // k_detector's photon ticket
    kHotHits = 3 * kHotStride,         // hits count
    kHotEvents = 4 * kHotStride,       // post-readout events count (a line of its own: both are bumped by every warp's flushes)
    kHotLines = 5,
    kHotWords = 5 * kHotStride
};

size_t sort_state_bytes();
size_t sort_lookback_words(size_t capacity);   // status words one radix pass needs for `capacity` keys
unsigned scan_tiles(size_t capacity);          // tiles of the compaction scans
unsigned scan_status_stride();                 // words between the status words of two tiles (one per 128-byte line)
unsigned bucket_words();                       // slice counters of the bucket sort

// ---- digitizer (digitizer.cu) --------------------------------------------------------------------------
// range == nullptr: the key range is measured on the device first (replay entry)
// reset: clear the stage's share of the frame state first (false inside gpet_run, where one memset per frame clears all)
// with_fallback == false: the LSD fallback kernel is not enqueued (6.6 us of idle cooperative launch per frame); if a
// slice of the bucket sort overflows, counters[5] is raised, no singles are produced and the caller must run again
int launch_digitize(EventBuf ev, const DigitizerOut& out, const DigitizerDev& p, DigitizerWorkspace& ws, const TimeRange* range,
                    uint64_t seed, int num_sms, cudaStream_t s, bool reset, bool with_fallback);

// the frame's counter block (32 words) and hot counters (kHotLines x 2 words) to pinned host memory by zero-copy stores
int launch_publish_counters(const unsigned* counters, const unsigned* hot, unsigned* h_dst, cudaStream_t s);
// addnoise: events of the noise process with t_lo <= t < t_hi appended to ev (digitizer.cu)
int launch_noise(EventBuf ev, const DigitizerDev& p, double t_lo_us, double t_hi_us, uint64_t seed, int num_sms, cudaStream_t s);

// scatter tags of a host-supplied list of photon numbers (gpet_mark_scattered)
int launch_mark_scattered(const int* parn, unsigned n, unsigned char* tag, unsigned mask, unsigned serial, cudaStream_t s);

// ---- transport (transport.cu) ----------------------------------------------------------------------------
int launch_source(const SourceDev* frame_dev, unsigned long long npairs, PhantomDev ph, PhotonQueue q0,
                  uint64_t seed, int num_sms, cudaStream_t s);
int launch_psf_positron(const void* positrons_aos, PhotonQueue q0, unsigned int n_positrons, unsigned long long first,
                        PhantomDev ph, float nonangle, int use_prange, uint64_t seed, int num_sms, cudaStream_t s);
// id_base: global index of the first photon of the frame in the queue (device_types.cuh photon_index); 0 = ids as they are
int launch_phantom(PhotonQueue q0, PhotonQueue q1, PhantomDev ph, TablesDev tb, float eabs, uint64_t seed, unsigned long long id_base,
                   int num_sms, cudaStream_t s);
int launch_panel_entry(PhotonQueue q1, PhotonQueue q2, DetectorDev det, unsigned int* counters, int num_sms, cudaStream_t s);
// fused source (frame_dev != nullptr) or queue q0 (frame_dev == nullptr) -> phantom -> panel entry -> q2; q1 only counts
int launch_front(const SourceDev* frame_dev, unsigned long long npairs, PhotonQueue q0, PhotonQueue q1, PhotonQueue q2,
                 PhantomDev ph, TablesDev tb, DetectorDev det, float eabs, unsigned int* counters, unsigned int* hot, uint64_t seed,
                 unsigned long long id_base, int num_sms, cudaStream_t s, bool reset);
int launch_detector(PhotonQueue q2, DetectorDev det, TablesDev tb, float eabs, int readout_depth, int readout_policy,
                    int record_hits, HitBuffer hits, EventBuf ev, unsigned int* counters, unsigned int* hot, uint64_t seed,
                    unsigned long long id_base, int num_sms, cudaStream_t s, bool reset);
// the SoA hit buffer in the reference's file layout (HitsID.dat / Hits.dat rows) or as gpet_hit records
int launch_hits_to_rows(HitBuffer hits, unsigned int n, int* id5, float* f5, cudaStream_t s);
int launch_hits_to_aos(HitBuffer hits, unsigned int n, void* aos, cudaStream_t s);
int launch_photons_aos_to_queue(const void* aos, PhotonQueue q, unsigned int n, cudaStream_t s);
int launch_queue_to_photons_aos(PhotonQueue q, void* aos, cudaStream_t s);

}  // namespace gpet
