// Host-callable launchers of the sm_100a kernels.  Every launcher is asynchronous on `stream` and returns the
// number of kernel launches it issued (for gpet_stats.kernel_launches).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "device_types.cuh"

namespace gpet {

struct SortWorkspace {
    unsigned long long* keys[2];  // ping-pong
    unsigned int* vals[2];
    unsigned int* tile_hist;      // 256 * max_tiles
    unsigned int max_tiles;
    unsigned int capacity;
};

struct DigitizerWorkspace {
    SortWorkspace sort;
    unsigned int* order_t;     // event index in time order (first count1 entries alive)
    unsigned int* order_s;     // position-in-time-order, sorted by (site, t)
    unsigned int* site_sorted; // site of order_s[p]
    unsigned char* kill;       // per event: 1 = removed by dead time
    unsigned int* flags;       // per time-order position: survives everything
    unsigned int* scan_tmp;    // block sums for the compaction scan
    unsigned int* counters;    // [0] n_in [1] after thresholder [2] after deadtime [3] singles [4] coincidences [5..] scratch
    unsigned long long* spectrum; int spectrum_bins; float spec_emin, spec_emax;
};

// ---- digitizer (digitizer.cu) --------------------------------------------------------------------------
int launch_events_aos_to_soa(const void* aos, EventSoA ev, unsigned int n, cudaStream_t s);
int launch_events_soa_to_aos(EventSoA ev, void* aos, cudaStream_t s);
int launch_digitize(EventSoA ev, EventSoA singles, void* singles_aos, void* coinc_aos, unsigned int coinc_cap,
                    const DigitizerDev& p, DigitizerWorkspace& ws, uint64_t seed, int num_sms, cudaStream_t s);
int launch_radix_sort_pairs(SortWorkspace& ws, const unsigned int* n_dev, int begin_bit, int end_bit,
                            int* result_buffer, int num_sms, cudaStream_t s);

// ---- transport (transport.cu) ----------------------------------------------------------------------------
int launch_source(const SourceDev* frame_dev, unsigned long long npairs, PhantomDev ph, PhotonQueue q0,
                  uint64_t seed, int num_sms, cudaStream_t s);
int launch_psf_positron(PhotonQueue q0, unsigned int n_positrons, PhantomDev ph, float nonangle, int use_prange,
                        uint64_t seed, int num_sms, cudaStream_t s);
int launch_phantom(PhotonQueue q0, PhotonQueue q1, PhantomDev ph, TablesDev tb, float eabs, uint64_t seed,
                   int num_sms, cudaStream_t s);
int launch_detector(PhotonQueue q1, DetectorDev det, TablesDev tb, float eabs, int readout_depth, int readout_policy,
                    int record_hits, HitBuffer hits, EventSoA ev, unsigned int* counters, uint64_t seed,
                    int num_sms, cudaStream_t s);
int launch_photons_aos_to_queue(const void* aos, PhotonQueue q, unsigned int n, cudaStream_t s);
int launch_queue_to_photons_aos(PhotonQueue q, void* aos, cudaStream_t s);

}  // namespace gpet
