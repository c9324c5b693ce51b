// Host-callable launchers of the sm_100a kernels.  Every launcher is asynchronous on `stream` and returns the
// number of kernel launches it issued (for gpet_stats.kernel_launches).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "device_types.cuh"

namespace gpet {

namespace rsort { struct SortState; }

struct DigitizerWorkspace {
    unsigned long long* tkeys[2];  // time sort ping-pong: order-preserving u64 image of the fp64 time
    unsigned int* tvals[2];        //   payload: event index
    unsigned int* skeys[2];        // site sort ping-pong: site number with the sign bit flipped
    unsigned int* svals[2];        //   payload: position in the time order
    unsigned int* lookback[2];     // sort_lookback_words(capacity) status words each (radix passes alternate between them)
    rsort::SortState* st_time;     // device-resident sort bookkeeping (histograms, tile counters, current buffer)
    rsort::SortState* st_site;
    unsigned int* scan_status[3];  // status words of the chained scans: singles compaction, coincidence compaction (max_tiles
                                   // each), slice counters of the bucket sort (bucket_words() / 2048)
    unsigned int* bcount;          // bucket sort of the time keys: events per slice, slice starts, scatter cursors
    unsigned int* bstart;          //   (bucket_words() entries each)
    unsigned int* bcur;
    unsigned long long* minmax;    // [0] smallest, [1] largest time key of the alive records
    unsigned int max_tiles;
    unsigned int capacity;
    unsigned int* order_t;         // event index in time order (first counters[1] entries alive)
    unsigned char* kill;           // per time-order position: 1 = removed by dead time
    unsigned int* coinc_cnt;       // per single: coincidences it opens
    // [0] n_in [1] after thresholder [2] after deadtime [3] singles [4] coincidences [6],[7] tile tickets of the two
    // compactions [8] photons on a panel [9] adder drops [10],[11] photon tickets of k_detector / k_front [12] time sort
    // fell back to LSD radix; [16..19] queue 0, queue 1, hits, events counts, [21] queue 2
    unsigned int* counters;
    unsigned long long* spectrum; int spectrum_bins; float spec_emin, spec_emax;
};

size_t sort_state_bytes();
size_t sort_lookback_words(size_t capacity);   // status words one radix pass needs for `capacity` keys
unsigned scan_tiles(size_t capacity);          // status words of the compaction scans
unsigned bucket_words();                       // slice counters of the bucket sort

// ---- digitizer (digitizer.cu) --------------------------------------------------------------------------
int launch_digitize(EventBuf ev, void* singles_aos, unsigned int singles_cap, void* coinc_aos, unsigned int coinc_cap,
                    const DigitizerDev& p, DigitizerWorkspace& ws, uint64_t seed, int num_sms, cudaStream_t s);

// ---- transport (transport.cu) ----------------------------------------------------------------------------
int launch_source(const SourceDev* frame_dev, unsigned long long npairs, PhantomDev ph, PhotonQueue q0,
                  uint64_t seed, int num_sms, cudaStream_t s);
int launch_psf_positron(const void* positrons_aos, PhotonQueue q0, unsigned int n_positrons, unsigned long long first,
                        PhantomDev ph, float nonangle, int use_prange, uint64_t seed, int num_sms, cudaStream_t s);
int launch_phantom(PhotonQueue q0, PhotonQueue q1, PhantomDev ph, TablesDev tb, float eabs, uint64_t seed,
                   int num_sms, cudaStream_t s);
int launch_panel_entry(PhotonQueue q1, PhotonQueue q2, DetectorDev det, unsigned int* counters, int num_sms, cudaStream_t s);
// fused source (frame_dev != nullptr) or queue q0 (frame_dev == nullptr) -> phantom -> panel entry -> q2; q1 only counts
int launch_front(const SourceDev* frame_dev, unsigned long long npairs, PhotonQueue q0, PhotonQueue q1, PhotonQueue q2,
                 PhantomDev ph, TablesDev tb, DetectorDev det, float eabs, unsigned int* counters, uint64_t seed, int num_sms,
                 cudaStream_t s);
int launch_detector(PhotonQueue q2, DetectorDev det, TablesDev tb, float eabs, int readout_depth, int readout_policy,
                    int record_hits, HitBuffer hits, EventBuf ev, unsigned int* counters, uint64_t seed,
                    int num_sms, cudaStream_t s);
int launch_photons_aos_to_queue(const void* aos, PhotonQueue q, unsigned int n, cudaStream_t s);
int launch_queue_to_photons_aos(PhotonQueue q, void* aos, cudaStream_t s);

}  // namespace gpet
