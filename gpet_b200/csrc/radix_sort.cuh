// Stable LSD radix sort of (key, value) pairs with 8-bit digits ("onesweep": one sweep per digit, chained scan with
// decoupled look-back), hand-written for sm_100a, as DEVICE functions: the digitizer's two fallback kernels
// (digitizer.cu: k_lsd_time_sort, k_dead_sorted) are persistent cooperative kernels that run all passes of a sort
// with a grid barrier in between, so a sort that is not needed costs one empty launch.
//
// Replaces the reference's D2H -> std::sort -> H2D round trips (quicksort_h / orderevents, detector.cu:354-385).
// Everything the sort needs to know at run time lives on the device: the element count, which of the two ping-pong
// buffers currently holds the data (derived from the per-digit histograms).
//
//   producer step      writes keys/vals into buffer 0, accumulates the histograms of ALL digits (hist_add / hist_flush)
//                      and clears look-back array 0 (clear_lookback)
//   onesweep_pass x P  pass p: skip if one digit value holds every key (high bytes of fp64 times inside a short frame,
//                      high bytes of site numbers); else tiles are claimed in order from a device counter, ranked
//                      inside the tile (warp ballots => stable), chained to their predecessors through a per-digit
//                      status word (aggregate / inclusive-prefix flags) and scattered.  Every pass clears the look-back
//                      array of the next pass.  Which buffer holds the data after p passes follows from the
//                      histograms alone (current_buffer), so consumers need no flag either.
//
// Ranking inside a tile: each warp walks its 32-wide rows in memory order, lanes with equal digits find each other
// with one ballot per digit bit, the lowest lane bumps the warp's digit counter, so ranks are stable without
// per-thread counters.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "ktimer.hpp"

namespace gpet {
namespace rsort {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kItems = 4;
constexpr int kTile = kThreads * kItems;  // 1024 keys per tile: small tiles keep every SM busy at frame sizes of ~1e5 keys
constexpr int kBins = 256;
constexpr int kMaxPasses = 8;

constexpr unsigned kFlagAggregate = 1u << 30;
constexpr unsigned kFlagPrefix = 2u << 30;
constexpr unsigned kValueMask = (1u << 30) - 1u;

// Device-resident bookkeeping of one sort (zeroed by the chain's first kernel).
struct SortState {
    unsigned tile_counter[kMaxPasses];  // next tile to claim, per pass
    unsigned pad[24];
    unsigned hist[kMaxPasses * kBins];  // global digit histograms, all passes
};

template <typename KeyT>
__device__ __forceinline__ unsigned digit_of(KeyT k, int shift) { return (unsigned)(k >> shift) & 0xFFu; }

__device__ __forceinline__ unsigned ld_volatile(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }
__device__ __forceinline__ void st_volatile(unsigned* p, unsigned v) { *reinterpret_cast<volatile unsigned*>(p) = v; }

// Decoupled look-back: sum of the published values of tiles [0, tile).  Predecessors are read in windows of 16
// INDEPENDENT loads (one L2 round trip per window instead of one per tile), walking back until a tile that already
// holds its inclusive prefix; a word that is not published yet is simply read again.
constexpr int kLookbackWindow = 32;
__device__ __forceinline__ unsigned lookback_sum(const unsigned* status, unsigned stride, unsigned tile) {
    unsigned excl = 0;
    int t = (int)tile - 1;
    while (t >= 0) {
        unsigned v[kLookbackWindow];
#pragma unroll
        for (int k = 0; k < kLookbackWindow; k++) v[k] = (t - k >= 0) ? ld_volatile(status + (size_t)(t - k) * stride) : kFlagPrefix;
        int k = 0;
#pragma unroll
        for (; k < kLookbackWindow; k++) {
            if ((v[k] >> 30) == 0u) break;  // not published yet: resume from this tile
            excl += v[k] & kValueMask;
            if (v[k] & kFlagPrefix) return excl;
        }
        t -= k;
    }
    return excl;
}

// ---- helpers for the producer kernel --------------------------------------------------------------------
// Block-level accumulation of the digit histograms of all passes: `sh` is [npass * 256] shared counters (zeroed by
// the caller), flushed once per block.
template <typename KeyT, int NPASS>
__device__ __forceinline__ void hist_add(unsigned* sh, KeyT key) {
#pragma unroll
    for (int p = 0; p < NPASS; p++) atomicAdd(&sh[p * kBins + digit_of(key, 8 * p)], 1u);
}

template <int NPASS>
__device__ __forceinline__ void hist_flush(const unsigned* sh, SortState* st) {
    for (int i = threadIdx.x; i < NPASS * kBins; i += blockDim.x) {
        unsigned c = sh[i];
        if (c) atomicAdd(&st->hist[i], c);
    }
}

// grid-stride clear of the look-back words the first pass will use
__device__ __forceinline__ void clear_lookback(unsigned* lookback, unsigned n) {
    const unsigned words = ((n + kTile - 1) / kTile) * kBins;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) lookback[i] = 0u;
}

// Which ping-pong buffer holds the data before pass `npasses_before` (= after that many passes): a pass whose keys all
// share one digit value is skipped, every other pass flips the buffer.  Computed by every block from the global
// histograms (no device-side flag, no completion counter).  Must be called by all kThreads threads of the block.
__device__ __forceinline__ unsigned current_buffer(const SortState* st, int npasses_before, unsigned n) {
    __shared__ unsigned s_mask[kWarps];
    unsigned trivial_mask = 0;
    for (int q = 0; q < npasses_before; q++)
        if (st->hist[q * kBins + threadIdx.x] == n) trivial_mask |= 1u << q;
    trivial_mask = __reduce_or_sync(0xffffffffu, trivial_mask);
    if ((threadIdx.x & 31) == 0) s_mask[threadIdx.x >> 5] = trivial_mask;
    __syncthreads();
    unsigned m = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) m |= s_mask[w];
    __syncthreads();
    return (unsigned)(npasses_before - __popc(m)) & 1u;
}

// ---- grid barrier of a cooperative launch (all blocks resident) ---------------------------------------------------
// bar[0]: arrivals of the current generation, bar[1]: generation.  Zeroed once at allocation; self-resetting.
__device__ __forceinline__ void grid_barrier(unsigned* bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned gen = ld_volatile(bar + 1);
        __threadfence();
        if (atomicAdd(bar, 1u) == gridDim.x - 1) {
            st_volatile(bar, 0u);
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (ld_volatile(bar + 1) == gen) {}
        }
        __threadfence();
    }
    __syncthreads();
}

// ---- one pass -----------------------------------------------------------------------------------------------
// Shared memory of a pass; one instance per kernel, handed to every call.
struct PassSmem {
    unsigned whist[kWarps][kBins];  // per-warp digit counters -> exclusive warp offsets
    unsigned gbase[kBins];
    unsigned wsum[kWarps];
    unsigned s_tile;
};

template <typename KeyT>
__device__ __forceinline__ void onesweep_pass(PassSmem& sm, KeyT* keys0, KeyT* keys1, unsigned* vals0, unsigned* vals1,
                                              unsigned n, SortState* st, unsigned* lookback0, unsigned* lookback1, int pass) {
    auto& whist = sm.whist;
    auto& gbase = sm.gbase;
    auto& wsum = sm.wsum;
    unsigned& s_tile = sm.s_tile;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __syncthreads();  // the previous pass may still be reading the shared arrays
    // the tile ticket is requested first: its round trip overlaps the loads below (a skipped pass wastes it, harmlessly)
    if (tid == 0) s_tile = atomicAdd(&st->tile_counter[pass], 1u);
    const unsigned ntiles = (n + kTile - 1) / kTile;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int shift = 8 * pass;
    unsigned* lb = (pass & 1) ? lookback1 : lookback0;
    unsigned* lb_next = (pass & 1) ? lookback0 : lookback1;

    // exclusive scan over the digits of this pass's global histogram; a pass whose keys all share one digit is the
    // identity permutation and is skipped
    const unsigned cnt_d = st->hist[pass * kBins + tid];
    const unsigned cur = current_buffer(st, pass, n);
    const int trivial = __syncthreads_or(cnt_d == n);
    if (!trivial) {
        unsigned x = cnt_d;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= (unsigned)o) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        unsigned wp = 0;
        for (unsigned w = 0; w < warp; w++) wp += wsum[w];
        const unsigned digit_base = wp + x - cnt_d;
        const KeyT* __restrict__ keys_in = cur ? keys1 : keys0;
        const unsigned* __restrict__ vals_in = cur ? vals1 : vals0;
        KeyT* __restrict__ keys_out = cur ? keys0 : keys1;
        unsigned* __restrict__ vals_out = cur ? vals0 : vals1;
        // with no more tiles than blocks every tile is claimed by some block's first ticket: no second round
        const bool single_round = ntiles <= gridDim.x;

        while (true) {
#pragma unroll
            for (int w = 0; w < kWarps; w++) whist[w][tid] = 0;
            __syncthreads();
            const unsigned tile = s_tile;
            if (tile >= ntiles) break;
            const unsigned wbase = tile * kTile + warp * (32 * kItems);
            KeyT key[kItems];
            unsigned val[kItems], rank[kItems], dig[kItems];
#pragma unroll
            for (int k = 0; k < kItems; k++) {
                unsigned i = wbase + k * 32 + lane;
                bool valid = i < n;
                key[k] = valid ? keys_in[i] : (KeyT) ~(KeyT)0;
                val[k] = valid ? vals_in[i] : 0u;
                dig[k] = valid ? digit_of(key[k], shift) : 0xFFu;
            }
#pragma unroll
            for (int k = 0; k < kItems; k++) {
                // lanes holding the same digit: eight ballots, one per digit bit (constant cost; __match_any_sync walks
                // the distinct values one by one and is ~8x slower on random digits)
                unsigned peers = 0xffffffffu;
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    const bool bit = (dig[k] >> b) & 1u;
                    const unsigned m = __ballot_sync(0xffffffffu, bit);
                    peers &= bit ? m : ~m;
                }
                unsigned leader = __ffs(peers) - 1;
                unsigned old = 0;
                if (lane == leader) {
                    old = whist[warp][dig[k]];
                    whist[warp][dig[k]] = old + __popc(peers);
                }
                old = __shfl_sync(0xffffffffu, old, leader);
                rank[k] = old + __popc(peers & lt_mask);
                __syncwarp();
            }
            __syncthreads();
            // digit = tid: exclusive prefix over warps, tile total, chained scan over the preceding tiles
            unsigned run = 0;
#pragma unroll
            for (int w = 0; w < kWarps; w++) {
                unsigned c = whist[w][tid];
                whist[w][tid] = run;
                run += c;
            }
            // the padding of a ragged last tile was counted under digit 0xFF: take it out again
            if (tid == 0xFFu) {
                unsigned tile_end = (tile + 1) * kTile;
                if (tile_end > n) run -= tile_end - n;
            }
            unsigned excl = 0;
            if (tile == 0) {
                st_volatile(&lb[tid], kFlagPrefix | run);
            } else {
                st_volatile(&lb[tile * kBins + tid], kFlagAggregate | run);
                excl = lookback_sum(lb + tid, kBins, tile);
                st_volatile(&lb[tile * kBins + tid], kFlagPrefix | (excl + run));
            }
            gbase[tid] = digit_base + excl;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kItems; k++) {
                unsigned i = wbase + k * 32 + lane;
                if (i < n) {
                    unsigned pos = gbase[dig[k]] + whist[warp][dig[k]] + rank[k];
                    keys_out[pos] = key[k];
                    vals_out[pos] = val[k];
                }
            }
            if (single_round) break;
            __syncthreads();
            if (tid == 0) s_tile = atomicAdd(&st->tile_counter[pass], 1u);
        }
    }
    // clear the look-back array of the next pass (last used two passes ago)
    for (unsigned i = blockIdx.x * kThreads + tid; i < ntiles * kBins; i += gridDim.x * kThreads) lb_next[i] = 0u;
}

// one pass as a kernel of its own; `only_if` (may be null) points at a device word: the pass returns at once when it is 0
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) k_onesweep(KeyT* keys0, KeyT* keys1, unsigned* vals0, unsigned* vals1,
                                                       const unsigned* __restrict__ n_ptr, SortState* st,
                                                       unsigned* lookback0, unsigned* lookback1, int pass,
                                                       const unsigned* __restrict__ only_if) {
    __shared__ PassSmem sm;
    if (only_if && *only_if == 0u) return;
    onesweep_pass<KeyT>(sm, keys0, keys1, vals0, vals1, *n_ptr, st, lookback0, lookback1, pass);
}

}  // namespace rsort

// Issues the passes of a sort whose producer has already filled buffer 0, the histograms and look-back array 0.
// Returns the number of launches.  Which buffer holds the result follows from the histograms (rsort::current_buffer).
template <typename KeyT>
inline int radix_sort_passes(KeyT* keys[2], unsigned* vals[2], const unsigned* n_dev, rsort::SortState* st,
                             unsigned* lookback[2], int npasses, int grid, cudaStream_t s, const unsigned* only_if = nullptr) {
    for (int p = 0; p < npasses; p++)
        GPET_LAUNCH(sizeof(KeyT) == 8 ? "k_onesweep<u64>" : "k_onesweep<u32>", s,
                    rsort::k_onesweep<KeyT><<<grid, rsort::kThreads, 0, s>>>(keys[0], keys[1], vals[0], vals[1], n_dev, st,
                                                                             lookback[0], lookback[1], p, only_if));
    return npasses;
}

}  // namespace gpet
