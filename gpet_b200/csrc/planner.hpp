// Frame planner: the epoch loop of sampleParticle (gPET.cu:204-208, 260-282) and findT (gPET.cu:439-452).
// The reference finds decays by sweeping every atom each epoch (setPosition, gPET_kernals.cu:497-558: O(N_atoms)
// Bernoulli trials).  Here the number of decays of each source in a time slice is drawn directly from the
// binomial law those trials define, then thinned by the positron branching ratio; the kernel only generates the
// pairs that exist (O(decays)).  Frames are independent of the number of GPUs (all ranks plan all frames).
#pragma once
#include <string>
#include <vector>

#include "ctx.hpp"

namespace gpet {

// first_pair0: global index of the first pair of the acquisition (64-bit history numbers; 0 unless the caller continues or
// shards a longer history sequence, gpet_set_first_pair)
std::string plan_frames(const Sources& src, const Isotopes& iso, float tstart_s, float tend_s, uint64_t max_pairs,
                        uint64_t seed, uint64_t first_pair0, std::vector<FramePlan>& out);
void fill_source_dev(const Sources& src, const Isotopes& iso, const FramePlan& fp, float nonangle, int use_prange,
                     SourceDev& d);

// host Philox4x32-10 (same function as philox.cuh) and the binomial sampler, exposed for tests
void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
uint64_t sample_binomial(uint64_t n, double p, uint64_t seed, uint64_t stream, uint64_t index);

}  // namespace gpet
