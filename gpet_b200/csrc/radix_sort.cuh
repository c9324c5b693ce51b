// Stable LSD radix sort of (key, value) pairs with 8-bit digits, one kernel per digit ("onesweep": chained scan with
// decoupled look-back), hand-written for sm_100a.
//
// Replaces the reference's D2H -> std::sort -> H2D round trips (quicksort_h / orderevents, detector.cu:354-385).
// Everything the sort needs to know at run time lives on the device: the element count, which of the two ping-pong
// buffers currently holds the data (SortState::cur), the per-digit histograms.  The launch sequence is therefore
// static and can sit inside a CUDA graph:
//
//   producer kernel    writes keys/vals into buffer 0, accumulates the histograms of ALL digits (hist_accumulate) and
//                      clears look-back array 0 (clear_lookback)
//   k_onesweep x P     pass p: skip if one digit value holds every key (high bytes of fp64 times inside a short frame,
//                      high bytes of site numbers); else tiles are claimed in order from a device counter, ranked
//                      inside the tile (warp match-any => stable), chained to their predecessors through a per-digit
//                      status word (aggregate / inclusive-prefix flags) and scattered.  Every pass clears the look-back
//                      array of the next pass; the last block to finish flips SortState::cur.
//
// Ranking inside a tile: each warp walks its 32-wide rows in memory order, lanes with equal digits find each other
// with __match_any_sync, the lowest lane bumps the warp's digit counter, so ranks are stable without per-thread
// counters.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gpet {
namespace rsort {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;  // 2048 keys per tile
constexpr int kBins = 256;
constexpr int kMaxPasses = 8;

constexpr unsigned kFlagAggregate = 1u << 30;
constexpr unsigned kFlagPrefix = 2u << 30;
constexpr unsigned kValueMask = (1u << 30) - 1u;

// Device-resident bookkeeping of one sort (zeroed by the chain's first kernel).
struct SortState {
    unsigned cur;                       // buffer (0/1) that holds the current data
    unsigned tile_counter[kMaxPasses];  // next tile to claim, per pass
    unsigned done_counter[kMaxPasses];  // blocks that finished, per pass
    unsigned pad[15];
    unsigned hist[kMaxPasses * kBins];  // global digit histograms, all passes
};

template <typename KeyT>
__device__ __forceinline__ unsigned digit_of(KeyT k, int shift) { return (unsigned)(k >> shift) & 0xFFu; }

__device__ __forceinline__ unsigned ld_volatile(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }
__device__ __forceinline__ void st_volatile(unsigned* p, unsigned v) { *reinterpret_cast<volatile unsigned*>(p) = v; }

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

// ---- one pass -----------------------------------------------------------------------------------------------
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) k_onesweep(KeyT* keys0, KeyT* keys1, unsigned* vals0, unsigned* vals1,
                                                       const unsigned* __restrict__ n_ptr, SortState* st,
                                                       unsigned* lookback0, unsigned* lookback1, int pass) {
    __shared__ unsigned whist[kWarps][kBins];  // per-warp digit counters -> exclusive warp offsets
    __shared__ unsigned gbase[kBins];
    __shared__ unsigned wsum[kWarps];
    __shared__ unsigned s_tile;
    const unsigned n = *n_ptr;
    const unsigned ntiles = (n + kTile - 1) / kTile;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int shift = 8 * pass;
    unsigned* lb = (pass & 1) ? lookback1 : lookback0;
    unsigned* lb_next = (pass & 1) ? lookback0 : lookback1;

    // exclusive scan over the digits of this pass's global histogram; a pass whose keys all share one digit is the
    // identity permutation and is skipped
    const unsigned cnt_d = st->hist[pass * kBins + tid];
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
        const unsigned cur = st->cur;
        const KeyT* __restrict__ keys_in = cur ? keys1 : keys0;
        const unsigned* __restrict__ vals_in = cur ? vals1 : vals0;
        KeyT* __restrict__ keys_out = cur ? keys0 : keys1;
        unsigned* __restrict__ vals_out = cur ? vals0 : vals1;

        while (true) {
            __syncthreads();
            if (tid == 0) s_tile = atomicAdd(&st->tile_counter[pass], 1u);
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
                unsigned peers = __match_any_sync(0xffffffffu, dig[k]);
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
                unsigned t = tile - 1;
                while (true) {
                    unsigned v = ld_volatile(&lb[t * kBins + tid]);
                    if ((v >> 30) == 0u) continue;  // predecessor not published yet
                    excl += v & kValueMask;
                    if (v & kFlagPrefix) break;
                    t--;
                }
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
        }
    }
    // clear the look-back array of the next pass (last used two passes ago)
    for (unsigned i = blockIdx.x * kThreads + tid; i < ntiles * kBins; i += gridDim.x * kThreads) lb_next[i] = 0u;
    __syncthreads();
    if (tid == 0 && !trivial) {
        __threadfence();
        if (atomicAdd(&st->done_counter[pass], 1u) == gridDim.x - 1) st->cur ^= 1u;
    }
}

}  // namespace rsort

// Issues the passes of a sort whose producer has already filled buffer 0, the histograms and look-back array 0.
// Returns the number of launches.  The result is in buffer st->cur (device side).
template <typename KeyT>
inline int radix_sort_passes(KeyT* keys[2], unsigned* vals[2], const unsigned* n_dev, rsort::SortState* st,
                             unsigned* lookback[2], int npasses, int grid, cudaStream_t s) {
    for (int p = 0; p < npasses; p++)
        rsort::k_onesweep<KeyT><<<grid, rsort::kThreads, 0, s>>>(keys[0], keys[1], vals[0], vals[1], n_dev, st, lookback[0],
                                                                 lookback[1], p);
    return npasses;
}

}  // namespace gpet
