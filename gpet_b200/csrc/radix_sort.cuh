// Stable LSD radix sort of (key, value) pairs, 8-bit digits, hand-written for sm_100a.
//
// Replaces the reference's D2H -> std::sort -> H2D round trips (quicksort_h / orderevents, detector.cu:354-385).
// The element count lives on the device (no host sync between digitizer stages); the pass count is host-known.
// Per pass: (1) per-tile digit histograms, (2) one exclusive scan of the bin-major histogram table,
// (3) stable scatter.  Ranking inside a tile uses warp match-any: each warp walks its 32-wide rows in memory
// order, lanes with equal digits find each other with __match_any_sync, the lowest lane bumps the warp's
// digit counter, so ranks are stable without per-thread counters.
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

template <typename KeyT>
__device__ __forceinline__ unsigned digit_of(KeyT k, int shift) { return (unsigned)(k >> shift) & 0xFFu; }

// (1) per-tile histograms -> hist[bin * ntiles + tile]
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) k_hist(const KeyT* __restrict__ keys, const unsigned* __restrict__ n_ptr,
                                                   unsigned* __restrict__ hist, int shift) {
    __shared__ unsigned sh[kBins];
    const unsigned n = *n_ptr;
    const unsigned ntiles = (n + kTile - 1) / kTile;
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        sh[threadIdx.x] = 0;
        __syncthreads();
        const unsigned base = tile * kTile;
#pragma unroll
        for (int k = 0; k < kItems; k++) {
            unsigned i = base + k * kThreads + threadIdx.x;
            if (i < n) atomicAdd(&sh[digit_of(keys[i], shift)], 1u);
        }
        __syncthreads();
        hist[threadIdx.x * ntiles + tile] = sh[threadIdx.x];
        __syncthreads();
    }
}

// (2) exclusive scan of m = 256 * ntiles counters, one block (m is small: n / 8)
__global__ void __launch_bounds__(1024) k_scan_table(unsigned* __restrict__ hist, const unsigned* __restrict__ n_ptr) {
    __shared__ unsigned warp_sums[32];
    __shared__ unsigned carry;
    const unsigned n = *n_ptr;
    const unsigned ntiles = (n + kTile - 1) / kTile;
    const unsigned m = ntiles * kBins;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // each thread owns 4 consecutive counters per round
    for (unsigned base = 0; base < m; base += 1024 * 4) {
        unsigned i0 = base + threadIdx.x * 4;
        unsigned v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = (i0 + k < m) ? hist[i0 + k] : 0u;
        unsigned tsum = v[0] + v[1] + v[2] + v[3];
        unsigned x = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= (unsigned)o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= (unsigned)o) w += y;
            }
            warp_sums[lane] = w;  // inclusive
        }
        __syncthreads();
        unsigned excl = carry + (warp ? warp_sums[warp - 1] : 0u) + (x - tsum);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < m) hist[i0 + k] = excl;
            excl += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry += warp_sums[31];
        __syncthreads();
    }
}

// (3) stable scatter
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) k_scatter(const KeyT* __restrict__ keys_in, const unsigned* __restrict__ vals_in,
                                                      KeyT* __restrict__ keys_out, unsigned* __restrict__ vals_out,
                                                      const unsigned* __restrict__ n_ptr,
                                                      const unsigned* __restrict__ hist, int shift) {
    __shared__ unsigned whist[kWarps][kBins];  // per-warp digit counters -> exclusive warp offsets
    __shared__ unsigned gbase[kBins];
    const unsigned n = *n_ptr;
    const unsigned ntiles = (n + kTile - 1) / kTile;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll
        for (int w = 0; w < kWarps; w++) whist[w][threadIdx.x] = 0;
        gbase[threadIdx.x] = hist[threadIdx.x * ntiles + tile];
        __syncthreads();
        const unsigned wbase = tile * kTile + warp * (32 * kItems);
        KeyT key[kItems];
        unsigned val[kItems], rank[kItems], dig[kItems];
#pragma unroll
        for (int k = 0; k < kItems; k++) {
            unsigned i = wbase + k * 32 + lane;
            bool valid = i < n;
            key[k] = valid ? keys_in[i] : (KeyT)~(KeyT)0;
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
        {   // exclusive prefix over warps for digit = threadIdx.x
            unsigned run = 0;
#pragma unroll
            for (int w = 0; w < kWarps; w++) {
                unsigned c = whist[w][threadIdx.x];
                whist[w][threadIdx.x] = run;
                run += c;
            }
        }
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
        __syncthreads();
    }
}

}  // namespace rsort

// Sorts n = *n_dev pairs held in buffer 0 of (keys, vals) over bits [begin_bit, end_bit); returns the index of the
// buffer that holds the result through *result_buffer and the number of launches as return value.
template <typename KeyT>
inline int radix_sort_pairs(KeyT* keys[2], unsigned* vals[2], unsigned* tile_hist, const unsigned* n_dev,
                            int begin_bit, int end_bit, int* result_buffer, int grid, cudaStream_t s) {
    int cur = 0, launches = 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        rsort::k_hist<KeyT><<<grid, rsort::kThreads, 0, s>>>(keys[cur], n_dev, tile_hist, shift);
        rsort::k_scan_table<<<1, 1024, 0, s>>>(tile_hist, n_dev);
        rsort::k_scatter<KeyT><<<grid, rsort::kThreads, 0, s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1],
                                                                n_dev, tile_hist, shift);
        cur ^= 1;
        launches += 3;
    }
    *result_buffer = cur;
    return launches;
}

}  // namespace gpet
