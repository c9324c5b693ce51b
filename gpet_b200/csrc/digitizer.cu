// Digitizer chain on the device: blur -> thresholder -> time sort -> dead time -> energy window -> singles (time
// sorted) [-> coincidence sorter].  Reference: blur/energywindow/setSitenum/deadtime kernels (gPET_kernals.cu:607-698,
// 814-837) and the host orchestration with three CPU sorts + orderevents (gPET.cu:385-424, detector.cu:354-385).
// Here nothing leaves the device between stages: counts stay in `counters`, events are 48-byte records in the file
// layout, and the launch sequence is static (all sizes and decisions live on the device).
//
// Time sort (D5).  Keys are the order-preserving u64 images of the fp64 times.  Times inside a frame are spread over
// the frame, so the sort is a bucket sort over equal slices of the time range (about 8 events per slice): k_prep takes
// each event's arrival rank in its slice with one atomic, k_bucket_scan turns the slice counts into starts,
// k_bucket_scatter places (key, index, site) at start + arrival rank, k_bucket_rank orders every slice in shared
// memory by (key, event index) -- the unique stable order whatever the arrival order was.  The key range comes from the
// caller (the frame's time slice) or from k_range (replay entry); keys outside it are clamped into the end slices, so
// the range only shapes the load, never the result.  If a slice holds more than kBucketLimit events (times clustered
// far below the range: never the case for decay data, but legal input of the replay entry point) k_bucket_scan raises
// counters[5] and ONE persistent cooperative kernel, k_lsd_fallback, runs the stable LSD radix sort of radix_sort.cuh
// (all passes, grid barriers in between); it is always enqueued and returns at once when the flag is clear.
//
// Dead time (D6, D7) needs no sort by site: in the reference's (site, t) order the predecessor of an event is the
// nearest earlier event of the same site in the time order, and it can only matter while t < (float)t_prev + tau, a
// condition that is monotone in t_prev.  So every event walks back through the time order while that holds and stops
// at the first event of its own site (paralyzable: that event kills it) -- O(events inside one dead time) per event.
// The non-paralyzable chain is cut at the events their predecessor cannot kill (k_deadtime_chain).
//
// Launches per frame: [k_range], k_prep, k_bucket_scan, k_bucket_scatter, k_bucket_rank, k_lsd_fallback
// (normally empty), [k_deadtime_chain: non-paralyzable only], k_emit_singles [, k_coinc].
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "kernels.hpp"
#include "philox.cuh"
#include "ktimer.hpp"
#include "radix_sort.cuh"

#include "../../include/gpet_b200.h"

namespace gpet {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxLogBuckets = 20;   // frames of up to ~8 M pairs keep ~6 events per slice (2^19 until r02z: the rank pass grew beyond 4 M pairs)
constexpr unsigned kMaxBuckets = 1u << kMaxLogBuckets;
constexpr unsigned kBucketLimit = 1024;   // a fuller slice sends the time sort to the LSD fallback
constexpr int kFlagLsd = 5;               // counters[kFlagLsd] != 0: time sort by LSD radix passes
constexpr unsigned kEwinBit = 0x80000000u;  // payload bit: the record is inside the final energy window

__device__ __forceinline__ unsigned long long time_key(double t) {
    unsigned long long b = (unsigned long long)__double_as_longlong(t);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ double key_time(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ unsigned warp_sum(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// sum over the block, valid in thread 0 (all threads must call)
__device__ __forceinline__ unsigned block_sum(unsigned v) {
    __shared__ unsigned s_part[kThreads / 32];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned t = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kThreads / 32; w++) t += s_part[w];
    return t;
}

// Equal slices of the TIME range [lo, hi] (not of the key range: the u64 image of a double is logarithmic in t):
// slice = clamp((long long)((t - lo) * inv), 0, nb - 1), monotone in t, which is all the sort needs.
struct BucketMap {
    double lo, inv;
    unsigned nb;
};

__device__ __forceinline__ BucketMap bucket_map(const TimeRange& r, unsigned n) {
    BucketMap m;
    const double lo = r.dev ? key_time(~r.dev[0]) : r.lo, hi = r.dev ? key_time(r.dev[1]) : r.hi;
    int lognb = (n > 1 ? 32 - __clz((int)(n - 1)) : 0) - 3;   // ceil(log2 n) - 3: about 8 events per slice
    lognb = min(max(lognb, 6), kMaxLogBuckets);
    m.nb = 1u << lognb;
    m.lo = lo;
    m.inv = (hi > lo && (hi - lo) < 1e300) ? (double)m.nb / (hi - lo) : 0.0;
    return m;
}

__device__ __forceinline__ unsigned bucket_of(const BucketMap& m, unsigned long long key) {
    const double x = (key_time(key) - m.lo) * m.inv;
    if (!(x > 0.0)) return 0u;
    return x < (double)m.nb ? (unsigned)x : m.nb - 1u;
}

// (double)((float)t_prev + tau): the end of the dead time an event at t_prev opens, with the reference's fp32 `tdead`
// and fp32 sum (gPET_kernals.cu:670-676); monotone in t_prev
__device__ __forceinline__ double dead_until(double t_prev, float tau) { return (double)__fadd_rn((float)t_prev, tau); }

// replay entry only: key range of the records that are not dead on arrival
__global__ void __launch_bounds__(kThreads) k_range(EventBuf ev, unsigned long long* __restrict__ minmax) {
    const unsigned n = min(*ev.count, ev.capacity);
    unsigned long long kmin = ~0ull, kmax = 0ull;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double t = ev.rec[i].t;
        if (t < kMaxT * 0.1) {
            const unsigned long long key = time_key(t);
            kmin = min(kmin, key);
            kmax = max(kmax, key);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((threadIdx.x & 31) == 0 && kmin <= kmax) {   // the smallest key is kept inverted, so that all zeros means "empty"
        atomicMax(&minmax[0], ~kmin);
        atomicMax(&minmax[1], kmax);
    }
}

// ------------------------------------------------------------------------------------------- noise singles
// addnoise (gPET_kernals.cu:699-735): thread `id` owns the time slice [id, id+1) * interval and walks a Poisson process
// of mean gap `lambda` through it; every arrival is an event with E = Emean + sigma * N(0,1), uniform position numbers in
// (0,1] and a uniformly drawn panel / module / crystal, parn = -1, crystal-level siten.  The reference never launches
// this kernel; what is specified here (and mirrored by the oracle): times accumulate in fp64 (the reference's fp32 `t`
// stops advancing once its ulp exceeds the gaps, i.e. after ~17 s of acquisition at us gaps), int(N * u) is clamped to
// N - 1 (u = 1 is possible), eventid = 0x80000000 | (slice << 10 | ordinal in the slice) so that noise events can be
// told apart, one Philox stream per slice, three blocks per arrival, and only arrivals inside [t_lo, t_hi) are kept
// (the frame being digitized).  Logarithm and Box-Muller in fp64, rounded once: the oracle's libm agrees.
__global__ void __launch_bounds__(kThreads) k_noise(EventBuf ev, DigitizerDev p, uint64_t seed, double t_lo, double t_hi,
                                                    long long id0, long long nslices) {
    const double interval = (double)p.noise_interval, lambda = (double)p.noise_gap;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nslices; k += (long long)gridDim.x * blockDim.x) {
        const long long id = id0 + k;
        double t = (double)id * interval;
        const double tend = (double)(id + 1) * interval;
        Philox rng(seed, (uint64_t)id, (uint32_t)kStageNoise << 24);
        for (unsigned ord = 0; ord < (1u << 20); ord++) {
            const uint4 r = rng.next();
            t = __dadd_rn(t, __dmul_rn(-log((double)u01(r.x)), lambda));   // no contraction: the oracle's arithmetic
            if (!(t < tend)) break;
            const uint4 q = rng.next();
            const uint4 w = rng.next();
            if (t < t_lo || !(t < t_hi)) continue;
            EventRec e;
            const double g = sqrt(-2.0 * log((double)u01(r.y))) * cos(6.283185307179586 * (double)u01(r.z));
            e.E = (float)__dadd_rn((double)p.noise_Emean, __dmul_rn((double)p.noise_sigma, g));
            e.x = u01(r.w); e.y = u01(q.x); e.z = u01(q.y);
            e.parn = -1;
            e.pann = min((int)((float)p.npanels * u01(q.z)), p.npanels - 1);
            e.modn = min((int)((float)p.moduleN * u01(q.w)), p.moduleN - 1);
            e.cryn = min((int)((float)p.crystalN * u01(w.x)), p.crystalN - 1);
            e.siten = e.pann * p.moduleN * p.crystalN + e.modn * p.crystalN + e.cryn;
            e.eventid = (int)(0x80000000u | ((((unsigned)id << 10) | (ord & 1023u)) & 0x7fffffffu));
            e.t = t;
            const unsigned slot = atomicAdd(ev.count, 1u);
            if (slot < ev.capacity) store_event_rec(ev.rec + slot, e);
        }
    }
}

// ------------------------------------------------------------------------------------------- stage 1: blur + thresholder + site + slice
// blur (gPET_kernals.cu:814-837) + energywindow(Eth, 2e6) (gPET.cu:393) + setSitenum (gPET_kernals.cu:607-640) fused.
// Per record: time key (all ones for a dead one), site number, arrival rank in its time slice | final-window flag.
__global__ void __launch_bounds__(kThreads, 4) k_prep(EventBuf ev, DigitizerDev p, uint64_t seed, TimeRange range,
                                                   unsigned long long* __restrict__ keys, int* __restrict__ site_of,
                                                   unsigned* __restrict__ aux, unsigned* __restrict__ bcount,
                                                   unsigned* __restrict__ counters) {
    pdl_wait();
    const unsigned n = min(*ev.count, ev.capacity);
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[0] = n;
    const BucketMap m = bucket_map(range, n);
    unsigned alive_cnt = 0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        EventRec* rec = ev.rec + i;
        const int4 a4 = reinterpret_cast<const int4*>(rec)[0];   // parn, pann, modn, cryn
        const int4 b4 = reinterpret_cast<const int4*>(rec)[1];   // siten, eventid, t
        float4 c4 = reinterpret_cast<const float4*>(rec)[2];     // E, x, y, z
        float E = c4.x;
        double t = __longlong_as_double((long long)(((unsigned long long)(unsigned)b4.w << 32) | (unsigned)b4.z));
        float R = 0.f;
        // float / double mix exactly as the reference expression is typed
        if (p.blur_policy == 0) R = __fmul_rn(__fsqrt_rn(__fdiv_rn(p.Eref, E)), p.Rref);
        if (p.blur_policy == 1) {
            // slope == 0 (the shipped setting): the quotient is +-0 and Rref + (+-0) = Rref, without the fp64 divide
            const float dE = __fsub_rn(E, p.Eref);
            if (p.slope == 0.f && fabsf(dE) <= 3.0e38f) R = p.Rref;
            else R = (float)__dadd_rn((double)p.Rref, __ddiv_rn((double)__fmul_rn(p.slope, dE), 1e6));
        }
        if (!(R > 0.f)) R = 0.f;
        // R == 0 leaves E bit-identical (E + 0), so the draw is skipped: this is the deterministic replay mode
        if (R > 0.f || p.sblur > 0.f || p.tblur > 0.f) {
            // stream = the photon; noise events (parn == -1, k_noise) are told apart by their event id
            // (the full 64-bit photon index: with 1e10 decays the 31-bit number of the record repeats every 2^30 pairs)
            const uint64_t who = a4.x == -1 ? ((1ull << 63) | (uint32_t)b4.y) : photon_index(a4.x, p.id_base);
            Philox rng(seed, who, ((uint32_t)kStageBlur << 24) | ((uint32_t)b4.x & 0xFFFFFFu));
            uint4 r = rng.next();
            float rad = sqrtf(-2.0f * logf(u01(r.x)));
            float g0 = rad * cosf(kTwoPi * u01(r.y));
            float nre = __fmul_rn(__fmul_rn(g0, R), E);
            E = (float)__dadd_rn((double)E, __ddiv_rn((double)nre, 2.35482));
            c4.x = E;
            if (p.sblur > 0.f) {
                uint4 q = rng.next();
                float ra = sqrtf(-2.0f * logf(u01(q.x))), rb = sqrtf(-2.0f * logf(u01(q.z)));
                float a0 = kTwoPi * u01(q.y), a1 = kTwoPi * u01(q.w);
                c4.y = __fadd_rn(c4.y, __fmul_rn(p.sblur, ra * cosf(a0)));
                c4.z = __fadd_rn(c4.z, __fmul_rn(p.sblur, ra * sinf(a0)));
                c4.w = __fadd_rn(c4.w, __fmul_rn(p.sblur, rb * cosf(a1)));
            }
            reinterpret_cast<float4*>(rec)[2] = c4;
            if (p.tblur > 0.f) {
                float g1 = rad * sinf(kTwoPi * u01(r.y));
                double tb = t + (double)p.tblur * (double)g1;
                t = tb > 0.0 ? tb : t;
                rec->t = t;
            }
        }
        // energywindow: dead iff E < lo || E > hi (gPET_kernals.cu:648); an already dead record (t >= MAXT) stays dead
        const bool alive = !(E < p.Eth || E > 2000000.0f) && t < kMaxT * 0.1;
        const unsigned long long key = alive ? time_key(t) : ~0ull;
        keys[i] = key;
        if (alive) {
            // setSitenum at the dead-time level; dlevel == 3 keeps what readout left (gPET.cu:402-407)
            int site = b4.x;
            if (p.dlevel == 0) site = 0;
            else if (p.dlevel == 1) site = a4.y;
            else if (p.dlevel == 2) site = a4.y * p.moduleN + a4.z;
            if (site != b4.x) rec->siten = site;
            site_of[i] = site;
            const unsigned arrival = atomicAdd(&bcount[bucket_of(m, key)], 1u);
            aux[i] = arrival | (!(E < p.Ewinmin || E > p.Ewinmax) ? kEwinBit : 0u);
            alive_cnt++;
        }
    }
    alive_cnt = block_sum(alive_cnt);   // one add per block: tallies of a whole grid land on one line and serialise there
    if (threadIdx.x == 0 && alive_cnt) atomicAdd(&counters[1], alive_cnt);
}

// ------------------------------------------------------------------------------------------- single-pass exclusive scan
// Tiles of 2048 elements (thread = 8 consecutive elements), claimed in order (ticket or block index with all blocks
// resident), so every earlier tile belongs to a running or finished block.  A tile publishes its total in one status
// word and then adds up the words of ALL earlier tiles, one word per thread and round: at frame sizes (<= a few
// thousand tiles) that is one L2 round trip, where a chained look-back walks ~40 ns per tile (measured:
// tools/microbench/latency.cu, 23 us for 343 tiles) because all tiles of these short kernels start at the same time
// and find no finished prefix to stop at.  So flagging, scanning and compacting happen in ONE kernel.
constexpr int kScanTile = 2048;
constexpr int kSpecSmemBins = 1024;
constexpr unsigned kPublished = 1u << 31;
// One status word per 128-byte line: every tile polls the words of all earlier tiles (343 tiles: 59 k loads per round),
// and accesses to ONE line are served one after the other (0.67 ns each for atomics, tools/microbench/latency.cu) --
// packed, the polling alone cost 10-15 us per scan (clock64 phase trace); on lines of their own the loads spread over
// the L2 slices.
constexpr unsigned kStatusStride = 32;

struct TileScan {
    unsigned excl[8];    // exclusive prefix of each of the thread's 8 elements (global)
    unsigned tile_total; // sum over the tile
    unsigned tile_excl;  // sum over all earlier tiles
};

__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ TileScan tile_exclusive_scan(const unsigned v[8], unsigned tile, unsigned* __restrict__ status) {
    __shared__ unsigned ws[kThreads / 32];
    __shared__ unsigned wx[kThreads / 32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += v[k];
    unsigned x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (unsigned)o) x += y;
    }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    unsigned wprefix = 0, total = 0;
#pragma unroll
    for (unsigned w = 0; w < kThreads / 32; w++) {
        unsigned c = ws[w];
        if (w < warp) wprefix += c;
        total += c;
    }
    if (threadIdx.x == 0) st_relaxed(&status[(size_t)tile * kStatusStride], kPublished | total);
    // totals of all earlier tiles (a word that is not published yet is read again)
    unsigned before = 0;
    for (unsigned i = threadIdx.x; i < tile; i += kThreads) {
        unsigned w;
        do { w = ld_relaxed(&status[(size_t)i * kStatusStride]); } while (!(w & kPublished));
        before += w & ~kPublished;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (lane == 0) wx[warp] = before;
    __syncthreads();
    unsigned excl_tile = 0;
#pragma unroll
    for (unsigned w = 0; w < kThreads / 32; w++) excl_tile += wx[w];
    TileScan r;
    r.tile_total = total;
    r.tile_excl = excl_tile;
    unsigned e = excl_tile + wprefix + (x - s);
#pragma unroll
    for (int k = 0; k < 8; k++) { r.excl[k] = e; e += v[k]; }
    __syncthreads();  // ws / wx are reused by the next tile
    return r;
}

// ------------------------------------------------------------------------------------------- stage 2: time sort (bucket sort)
// exclusive scan of the slice counters (one tile of 2048 per block, all blocks resident: tile = blockIdx);
// raises the LSD-fallback flag when a slice is overfull
__global__ void __launch_bounds__(kThreads) k_bucket_scan(const unsigned* __restrict__ bcount, unsigned* __restrict__ bstart,
                                                          unsigned* __restrict__ status, unsigned* __restrict__ counters,
                                                          TimeRange range) {
    pdl_wait();
    const unsigned tile = blockIdx.x;
    const unsigned b0 = tile * kScanTile + threadIdx.x * 8;
    unsigned c[8];
    bool over = false;
    {   // issued before the slice count is known (it hangs on a load of its own)
        const uint4 lo = reinterpret_cast<const uint4*>(bcount + b0)[0], hi = reinterpret_cast<const uint4*>(bcount + b0)[1];
        c[0] = lo.x; c[1] = lo.y; c[2] = lo.z; c[3] = lo.w; c[4] = hi.x; c[5] = hi.y; c[6] = hi.z; c[7] = hi.w;
    }
    const BucketMap m = bucket_map(range, counters[0]);
    if (tile * kScanTile >= m.nb) return;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (b0 + k >= m.nb) c[k] = 0u;
        over |= c[k] > kBucketLimit;
    }
    if (over) counters[kFlagLsd] = 1u;
    TileScan sc = tile_exclusive_scan(c, tile, status);
    if (b0 < m.nb) {   // nb is a multiple of 8
        reinterpret_cast<uint4*>(bstart + b0)[0] = make_uint4(sc.excl[0], sc.excl[1], sc.excl[2], sc.excl[3]);
        reinterpret_cast<uint4*>(bstart + b0)[1] = make_uint4(sc.excl[4], sc.excl[5], sc.excl[6], sc.excl[7]);
    }
}

__global__ void __launch_bounds__(kThreads) k_bucket_scatter(const unsigned long long* __restrict__ keys,
                                                             const int* __restrict__ site_of, const unsigned* __restrict__ aux,
                                                             const unsigned* __restrict__ counters, TimeRange range,
                                                             const unsigned* __restrict__ bstart, uint4* __restrict__ bent) {
    pdl_wait();
    if (counters[kFlagLsd]) return;
    const unsigned n = counters[0];
    const BucketMap m = bucket_map(range, n);
    const unsigned nth = gridDim.x * blockDim.x;
    // four independent elements per thread and round: the dependent gather of the slice start is the latency to hide
    for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * nth) {
        unsigned long long key[4];
        unsigned a[4], pos[4];
        int site[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned i = i0 + u * nth;
            key[u] = i < n ? keys[i] : ~0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned i = i0 + u * nth;
            if (key[u] != ~0ull) { a[u] = aux[i]; site[u] = site_of[i]; pos[u] = __ldg(&bstart[bucket_of(m, key[u])]); }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (key[u] == ~0ull) continue;
            // ONE 16-byte entry per event: the scatter is bound by the number of scattered sectors written, and a key array
            // plus a payload array cost two per event (22.6 -> 13.3 us on the 1.1 M-pair frame, 21.2 -> 15.9 us per M pairs at 4.5 M)
            const unsigned i = i0 + u * nth, o = pos[u] + (a[u] & ~kEwinBit);
            bent[o] = make_uint4((unsigned)key[u], (unsigned)(key[u] >> 32), i | (a[u] & kEwinBit), (unsigned)site[u]);
        }
    }
}

// rank of every event inside its slice by (key, event index): the stable time order.  One thread per event, four
// independent events per thread and round with their loads issued together; the events of a slice are contiguous and a
// few cache lines at most, so the walk over the slice runs out of L1 (the warp that ranks them has just loaded them).
// No shared memory and no barrier: at frame sizes this kernel is a latency chain (load, slice bounds, walk, store), and a
// block that staged 256 slices at a time spent 19 us on six dependent rounds whatever the frame size.
__device__ __forceinline__ unsigned long long entry_key(const uint4& e) { return ((unsigned long long)e.y << 32) | e.x; }

__global__ void __launch_bounds__(kThreads) k_bucket_rank(const uint4* __restrict__ bent, const unsigned* __restrict__ bstart,
                                                          const unsigned* __restrict__ counters, TimeRange range,
                                                          unsigned long long* __restrict__ tsort, unsigned* __restrict__ order_t,
                                                          int* __restrict__ site_t, int tie_site) {
    pdl_wait();
    if (counters[kFlagLsd]) return;
    const unsigned n1 = counters[1];
    const BucketMap m = bucket_map(range, counters[0]);
    const unsigned nth = gridDim.x * blockDim.x;
    for (unsigned e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < n1; e0 += 4 * nth) {
        unsigned long long key[4];
        uint2 pay[4];
        unsigned s[4], en[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned e = e0 + u * nth;
            if (e < n1) { const uint4 en4 = bent[e]; key[u] = entry_key(en4); pay[u] = make_uint2(en4.z, en4.w); }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned e = e0 + u * nth;
            if (e < n1) {
                const unsigned b = bucket_of(m, key[u]);
                s[u] = __ldg(&bstart[b]);
                en[u] = b + 1 < m.nb ? __ldg(&bstart[b + 1]) : n1;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned e = e0 + u * nth;
            if (e >= n1) continue;
            const unsigned idx = pay[u].x & ~kEwinBit;
            unsigned rank = 0;
            for (unsigned q = s[u]; q < en[u]; q++) {
                const uint4 eq = bent[q];
                const unsigned long long kq = entry_key(eq);
                if (kq < key[u]) rank++;
                else if (kq == key[u] && q != e) {
                    // equal times: input order -- or, for the events of a run (their order in the buffer is whatever the
                    // detector kernel's atomics made it), site number first, so that a run's singles do not depend on it
                    const uint2 pq = make_uint2(eq.z, eq.w);
                    if (tie_site && pq.y != pay[u].y) rank += (int)pq.y < (int)pay[u].y ? 1u : 0u;
                    else if ((pq.x & ~kEwinBit) < idx) rank++;
                }
            }
            const unsigned pos = s[u] + rank;
            tsort[pos] = key[u];
            order_t[pos] = pay[u].x;
            site_t[pos] = (int)pay[u].y;
        }
    }
}

// Fallback time sort: stable LSD radix sort of (key, index | window flag), all 8 passes in one persistent cooperative
// kernel (grid barriers between the phases), then the same three output arrays as k_bucket_rank.  Returns at once when
// the bucket sort did the job.
__global__ void __launch_bounds__(rsort::kThreads) k_lsd_fallback(unsigned long long* keys0, unsigned long long* keys1,
                                                                  unsigned* vals0, unsigned* vals1,
                                                                  const unsigned* __restrict__ aux, const int* __restrict__ site_of,
                                                                  const unsigned* __restrict__ counters, rsort::SortState* st,
                                                                  unsigned* lookback0, unsigned* lookback1, unsigned* bar,
                                                                  unsigned* __restrict__ order_t, int* __restrict__ site_t) {
    __shared__ rsort::PassSmem sm;
    __shared__ unsigned sh_hist[8 * rsort::kBins];
    if (counters[kFlagLsd] == 0u) return;
    const unsigned n = counters[0], n1 = counters[1];
    const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    {   // phase 0: clean sort state
        constexpr unsigned kWords = sizeof(rsort::SortState) / 4;
        unsigned* w = reinterpret_cast<unsigned*>(st);
        for (unsigned i = gtid; i < kWords; i += nth) w[i] = 0u;
        rsort::clear_lookback(lookback0, n);
    }
    rsort::grid_barrier(bar);
    // phase 1: payload + digit histograms of all passes
    for (int i = threadIdx.x; i < 8 * rsort::kBins; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    for (unsigned i = gtid; i < n; i += nth) {
        const unsigned long long key = keys0[i];
        vals0[i] = key != ~0ull ? (i | (aux[i] & kEwinBit)) : i;
        rsort::hist_add<unsigned long long, 8>(sh_hist, key);
    }
    __syncthreads();
    rsort::hist_flush<8>(sh_hist, st);
    rsort::grid_barrier(bar);
    for (int pass = 0; pass < 8; pass++) {
        rsort::onesweep_pass<unsigned long long>(sm, keys0, keys1, vals0, vals1, n, st, lookback0, lookback1, pass);
        rsort::grid_barrier(bar);
    }
    // the alive records come first (dead keys are all ones); outputs in the layout of k_bucket_rank (tsort = keys0)
    const unsigned cur = rsort::current_buffer(st, 8, n);
    const unsigned* __restrict__ v = cur ? vals1 : vals0;
    for (unsigned j = gtid; j < n1; j += nth) {
        const unsigned pv = v[j];
        order_t[j] = pv;
        site_t[j] = site_of[pv & ~kEwinBit];
        if (cur) keys0[j] = keys1[j];
    }
}

// ------------------------------------------------------------------------------------------- stage 3: dead time
// deadtime (gPET_kernals.cu:657-698) with the snapshot-start semantics of SURVEY 8(a) D7: every decision uses the
// original times.  Paralyzable: is event j inside the dead time of the nearest earlier event of its site?
__device__ __forceinline__ bool killed_by_predecessor(const unsigned long long* __restrict__ tsort, const int* __restrict__ site_t,
                                                      unsigned j, double t, int site, float tau) {
    for (unsigned q = j; q > 0;) {
        q--;
        if (!(t < dead_until(key_time(__ldg(&tsort[q])), tau))) return false;   // and so for every earlier event
        if (__ldg(&site_t[q]) == site) return true;
    }
    return false;
}

// Non-paralyzable: sequential anchor chain per site.  An event its predecessor cannot kill survives any earlier anchor
// as well (fp32 rounding and the fp32 sum are monotone), so it is a guaranteed anchor and the chain can be cut there:
// one thread per such run start walks forward through the time order while the next event of its site is inside the
// dead time of the previous one.
__device__ __forceinline__ void counters_flag_halo(unsigned* counters) { counters[15] = 1u; }

__global__ void __launch_bounds__(kThreads) k_deadtime_chain(DigitizerDev p, const unsigned long long* __restrict__ tsort,
                                                             const int* __restrict__ site_t, unsigned char* __restrict__ kill,
                                                             unsigned* __restrict__ counters) {
    const unsigned n1 = counters[1];
    const float tau = p.dtime;
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < n1; j += gridDim.x * blockDim.x) {
        const double t = key_time(tsort[j]);
        const int site = site_t[j];
        if (killed_by_predecessor(tsort, site_t, j, t, site, tau)) continue;   // member of an earlier start's chain
        kill[j] = 0;
        float tdead = (float)t;
        double until = dead_until(t, tau);   // dead time of the latest event of this site seen so far
        // a chain that starts inside the first dead time of the list may have lost its real start to the cut of the halo
        const bool suspect = p.emit_on && t < p.trust_lo;
        for (unsigned r = j + 1; r < n1; r++) {
            const double tr = key_time(tsort[r]);
            if (!(tr < until)) break;        // the next event of this site, if any, starts its own chain
            if (site_t[r] != site) continue;
            if (suspect && tr >= p.emit_lo) counters_flag_halo(counters);
            if (tr < (double)__fadd_rn(tdead, tau)) {
                kill[r] = 1;
            } else {
                kill[r] = 0;
                tdead = (float)tr;
            }
            until = dead_until(tr, tau);
        }
    }
}

// ------------------------------------------------------------------------------------------- stage 4: energy window + compaction -> singles
// dead time (paralyzable: decided here) + energywindow(Ewinmin, Ewinmax) (gPET.cu:418) over the time order; survivors are
// compacted into the singles list, their times and panels into two side arrays for the coincidence sorter.
// A tile's times and sites are staged in shared memory with a halo of earlier events, so the backward walks of the
// dead time run out of shared memory; the survivors' records are gathered as 16-byte pieces, eight independent loads
// per thread in flight, and leave as consecutive 16-byte stores.
constexpr int kHalo = 64;

__global__ void __launch_bounds__(kThreads, 3) k_emit_singles(EventBuf ev, DigitizerDev p, EventRec* __restrict__ singles,
                                                           unsigned singles_cap, const unsigned long long* __restrict__ tsort,
                                                           const unsigned* __restrict__ order_t, const int* __restrict__ site_t,
                                                           const unsigned char* __restrict__ kill, unsigned* __restrict__ counters,
                                                           unsigned* __restrict__ status, double* __restrict__ stime,
                                                           int* __restrict__ span, int* __restrict__ spar, int* __restrict__ seid,
                                                           unsigned long long* __restrict__ spectrum,
                                                           int nbins, int spec_stride, float emin, float emax, int fallback_skipped,
                                                           unsigned* __restrict__ h_singles_count) {
    pdl_wait();
    __shared__ unsigned s_tile;
    __shared__ double s_t[kHalo + kScanTile];
    __shared__ int s_site[kHalo + kScanTile];
    __shared__ unsigned s_idx[kScanTile];
    __shared__ __align__(16) unsigned char s_flag[kScanTile];
    __shared__ unsigned s_spec[kSpecSmemBins];   // block-private energy histogram (flushed once per block)
    // the caller did not enqueue the LSD fallback and a slice overflowed: there is no time order; leave no singles (the
    // host sees counters[kFlagLsd] and runs again with the fallback)
    if (fallback_skipped && counters[kFlagLsd]) {
        if (h_singles_count && blockIdx.x == 0 && threadIdx.x == 0) *h_singles_count = 0u;
        return;
    }
    const unsigned n1 = counters[1];
    if (h_singles_count && n1 == 0u && blockIdx.x == 0 && threadIdx.x == 0) *h_singles_count = 0u;   // no tile will report
    const unsigned ntiles = (n1 + kScanTile - 1) / kScanTile;
    const bool spec_smem = spectrum && nbins > 0 && nbins <= kSpecSmemBins;
    const float tau = p.dtime;
    if (spec_smem)
        for (int b = threadIdx.x; b < nbins; b += blockDim.x) s_spec[b] = 0;
    unsigned c2 = 0, n_before = 0, n_inside = 0;   // n_*: singles before / inside the emit window
#ifdef GPET_PHASE_TRACE
    long long ph[8]; int nph = 0;
#define PH() do { if (nph < 8) ph[nph++] = clock64(); } while (0)
    PH();
#else
#define PH() do {} while (0)
#endif
    while (true) {
        if (threadIdx.x == 0) s_tile = atomicAdd(&counters[6], 1u);
        __syncthreads();
        const unsigned tile = s_tile;
        if (tile >= ntiles) break;
        PH();
        const unsigned jt = tile * kScanTile;                 // first position of the tile
        const unsigned j0 = jt + threadIdx.x * 8;
        // this thread's 8 consecutive payloads (scan layout); issued first, used after the flags are known
        unsigned pv[8];
#pragma unroll
        for (int k = 0; k < 8; k++) pv[k] = (j0 + k < n1) ? order_t[j0 + k] : 0u;
        // stage [jt - kHalo, jt + kScanTile)
        {   // all loads first, then the shared-memory stores: one round trip instead of one per round of the loop
            constexpr int kRounds = (kHalo + kScanTile + kThreads - 1) / kThreads;
            unsigned long long kv[kRounds];
            int sv[kRounds];
#pragma unroll
            for (int r = 0; r < kRounds; r++) {
                const int e = r * kThreads + threadIdx.x;
                const long long j = (long long)jt - kHalo + e;
                const bool ok = e < kHalo + kScanTile && j >= 0 && j < (long long)n1;
                kv[r] = ok ? tsort[j] : 0ull;
                sv[r] = ok ? site_t[j] : 0;
            }
#pragma unroll
            for (int r = 0; r < kRounds; r++) {
                const int e = r * kThreads + threadIdx.x;
                if (e < kHalo + kScanTile) {
                    s_t[e] = key_time(kv[r]);
                    s_site[e] = sv[r];
                }
            }
        }
        __syncthreads();
        PH();
        // flags, one element per thread and round (conflict-free shared-memory walks)
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int e = kHalo + k * kThreads + threadIdx.x;
            const unsigned j = jt + k * kThreads + threadIdx.x;
            unsigned char f = 2;   // 2: beyond the list, 1: removed by dead time, 0: survives it
            if (j < n1) {
                bool dead;
                if (p.dtype == 0) {
                    const double t = s_t[e];
                    const int site = s_site[e];
                    dead = false;
                    int q = e - 1;
                    const int qmin = jt >= (unsigned)kHalo ? 0 : kHalo - (int)jt;   // first staged entry
                    for (; q >= qmin; q--) {
                        if (!(t < dead_until(s_t[q], tau))) break;
                        if (s_site[q] == site) { dead = true; break; }
                    }
                    if (q < qmin && qmin == 0 && jt > (unsigned)kHalo)   // the walk left the halo: go on in global memory
                        dead = killed_by_predecessor(tsort, site_t, jt - kHalo, t, site, tau);
                } else {
                    dead = kill[j] != 0;
                }
                f = dead ? 1 : 0;
                c2 += dead ? 0u : 1u;
            }
            s_flag[k * kThreads + threadIdx.x] = f;
        }
        __syncthreads();
        PH();
        unsigned flag[8];
        {
            const uint2 f8 = reinterpret_cast<const uint2*>(s_flag)[threadIdx.x];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const unsigned f = ((k < 4 ? f8.x : f8.y) >> (8 * (k & 3))) & 0xffu;
                flag[k] = (f == 0u && (pv[k] & kEwinBit)) ? 1u : 0u;
            }
        }
        TileScan sc = tile_exclusive_scan(flag, tile, status);
        PH();
        if (tile == ntiles - 1 && threadIdx.x == 0) {
            counters[3] = sc.tile_excl + sc.tile_total;
            // pinned host word (zero-copy store): the host starts the D2H copy of the singles from it without a
            // memcpy of its own on this stream, which would queue behind the previous frame's copy on the D2H engine
            if (h_singles_count) *h_singles_count = sc.tile_excl + sc.tile_total;
        }
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (flag[k]) s_idx[sc.excl[k] - sc.tile_excl] = pv[k] & ~kEwinBit;
        __syncthreads();
        // gather: piece l = 3 * r + part of the r-th survivor of the tile
        const unsigned npieces = 3u * sc.tile_total;
        const int4* __restrict__ src = reinterpret_cast<const int4*>(ev.rec);
        int4* __restrict__ dst = reinterpret_cast<int4*>(singles);
        for (unsigned l0 = 0; l0 < npieces; l0 += 8 * kThreads) {
            int4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const unsigned l = l0 + u * kThreads + threadIdx.x;
                if (l < npieces) {
                    const unsigned r = l / 3u, part = l - 3u * r;
                    v[u] = __ldg(src + 3ull * s_idx[r] + part);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const unsigned l = l0 + u * kThreads + threadIdx.x;
                if (l >= npieces) continue;
                const unsigned r = l / 3u, part = l - 3u * r;
                const unsigned o = sc.tile_excl + r;
                if (o >= singles_cap) continue;
                dst[3ull * o + part] = v[u];
                if (part == 0) {
                    span[o] = v[u].y;
                    spar[o] = v[u].x;
                } else if (part == 1) {
                    const double ts = __longlong_as_double((long long)(((unsigned long long)(unsigned)v[u].w << 32) | (unsigned)v[u].z));
                    stime[o] = ts;
                    seid[o] = v[u].y;
                    if (p.emit_on) {
                        if (ts < p.emit_lo) n_before++;
                        else if (ts < p.emit_hi) n_inside++;
                    }
                } else if (spectrum && nbins > 0) {
                    const float f = (__int_as_float(v[u].x) - emin) / (emax - emin) * nbins;
                    if (f >= 0.f && f < (float)nbins) {
                        if (spec_smem) atomicAdd(&s_spec[(int)f], 1u);
                        else atomicAdd(&spectrum[(size_t)(int)f * spec_stride], 1ull);
                    }
                }
            }
        }
        __syncthreads();   // shared arrays are reused by the next tile
        PH();
    }
#ifdef GPET_PHASE_TRACE
    if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 170 || blockIdx.x == 340) && nph >= 6)
        printf("emit block %d: ticket %lld stage %lld flags %lld scan %lld gather %lld cycles\n", blockIdx.x, ph[1] - ph[0], ph[2] - ph[1],
               ph[3] - ph[2], ph[4] - ph[3], ph[5] - ph[4]);
#endif
    if (spec_smem) {
        __syncthreads();
        for (int b = threadIdx.x; b < nbins; b += blockDim.x)
            if (s_spec[b]) atomicAdd(&spectrum[(size_t)b * spec_stride], (unsigned long long)s_spec[b]);
    }
    c2 = block_sum(c2);
    if (threadIdx.x == 0 && c2) atomicAdd(&counters[2], c2);
    if (p.emit_on) {
        __syncthreads();
        n_before = block_sum(n_before);
        if (threadIdx.x == 0 && n_before) atomicAdd(&counters[10], n_before);
        __syncthreads();
        n_inside = block_sum(n_inside);
        if (threadIdx.x == 0 && n_inside) atomicAdd(&counters[11], n_inside);
    }
}

// Singles as 32-byte records for the trip to the host (gpet_single_compact, include/gpet_b200.h): everything a 48-byte
// Event of a source-mode run holds that is not implied by the rest.  Two 16-byte stores per single, consecutive.
__global__ void __launch_bounds__(kThreads) k_pack_singles(const EventRec* __restrict__ singles, const unsigned* __restrict__ counters,
                                                           unsigned singles_cap, int4* __restrict__ out) {
    pdl_wait();
    const unsigned n = min(counters[3], singles_cap);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4* src = reinterpret_cast<const int4*>(singles + i);
        const int4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);   // parn pann modn cryn | siten eventid t | E x y z
        const unsigned ids = ((unsigned)a.y & 0xffu) | (((unsigned)a.z & 0xfffu) << 8) | (((unsigned)a.w & 0x7ffu) << 20) | (((unsigned)a.x & 1u) << 31);
        out[2ull * i] = make_int4(b.z, b.w, c.x, c.y);                       // t, E, x
        out[2ull * i + 1] = make_int4(c.z, c.w, b.y, (int)ids);              // y, z, eventid, ids
    }
}

// ------------------------------------------------------------------------------------------- stage 5: coincidence sorter (extension)
// Windows are opened by the first single that is not inside an earlier window and last cwin us.  A single whose
// predecessor is at least cwin earlier is a guaranteed opener, so every single finds its own role by replaying the
// windows from the nearest guaranteed opener at or before it (a few steps at realistic rates): no cross-thread state,
// so counting, scanning and emitting are ONE kernel.  Works on the side arrays of the singles list (time, panel),
// staged in shared memory with a halo on both sides.
struct SinglesView {
    const double* __restrict__ gt; const int* __restrict__ gp;   // global arrays
    const double* st; const int* sp;                             // staged copy of [lo, hi)
    unsigned lo, hi;
    __device__ __forceinline__ double t(unsigned i) const { return (i >= lo && i < hi) ? st[i - lo] : gt[i]; }
    __device__ __forceinline__ int pan(unsigned i) const { return (i >= lo && i < hi) ? sp[i - lo] : gp[i]; }
};

// Coincidence classes (the reference has neither a sorter nor a scatter flag, F2 / F11):
//   2 random  -- the two singles come from different annihilations (eventid >> pair_shift differ), or one of them is a
//                noise single (parn == -1, k_noise)
//   1 scatter -- same annihilation, and at least one of the two photons interacted in the phantom (its scatter tag,
//                DetectorDev::scat_tag, carries this frame's serial)
//   0 true    -- same annihilation, both photons unscattered in the phantom
// Per single: bit 0 = its photon scattered in the phantom, bit 1 = noise single.
constexpr unsigned kCfScattered = 1u, kCfNoise = 2u;
__device__ __forceinline__ unsigned class_flags(const DigitizerDev& p, int parn) {
    if (parn == -1) return kCfNoise;
    return (p.scat_tag != nullptr && (unsigned)__ldg(p.scat_tag + ((unsigned)parn & p.scat_mask)) == p.scat_serial) ? kCfScattered : 0u;
}
__device__ __forceinline__ unsigned coincidence_class(const DigitizerDev& p, int eid_a, unsigned cf_a, int eid_b, unsigned cf_b) {
    if (((cf_a | cf_b) & kCfNoise) || (eid_a >> p.pair_shift) != (eid_b >> p.pair_shift)) return 2u;
    return ((cf_a | cf_b) & kCfScattered) ? 1u : 0u;
}

__device__ __forceinline__ bool pair_ok(const SinglesView& v, unsigned a, unsigned b, const DigitizerDev& p) {
    if (p.cmindiff <= 0) return true;
    int d = abs(v.pan(a) - v.pan(b));
    if (p.npanels > 0) d = min(d, p.npanels - d);
    return d >= p.cmindiff;
}

// coincidences opened by single a (0 when it sits inside an earlier window)
__device__ __forceinline__ unsigned coincidences_of(const SinglesView& v, unsigned a, unsigned n, double W, const DigitizerDev& p,
                                                    unsigned* __restrict__ counters) {
    if (p.emit_on && !(v.t(a) >= p.emit_lo && v.t(a) < p.emit_hi)) return 0u;   // opened in a neighbour's slice: theirs
    unsigned w = a;
    while (w > 0 && !(v.t(w) >= v.t(w - 1) + W)) w--;   // nearest guaranteed opener
    // the replay must start from a single whose own status does not depend on what the halo cut off
    if (p.emit_on && v.t(w) < p.trust_lo) counters[15] = 1u;
    while (true) {
        const double tend = v.t(w) + W;
        unsigned m = 0;
        while (w + 1 + m < n && v.t(w + 1 + m) < tend) m++;
        if (w == a) {
            unsigned valid = 0;
            for (unsigned b = a + 1; b <= a + m; b++) valid += pair_ok(v, a, b, p) ? 1u : 0u;
            return p.cpolicy == 0 ? ((m == 1 && valid == 1) ? 1u : 0u) : valid;
        }
        if (a <= w + m) return 0u;   // inside w's window
        w += m + 1;
    }
}

// Index pairs into the run's singles list (pair_base = singles of the run's earlier frames, kept on the device) and,
// when `out` is given, the two 48-byte records side by side; `cls` (optional) receives one class byte per coincidence.
// The photon number and annihilation number of every single come from the side arrays k_emit_singles wrote (spar, seid);
// the scatter tags of the staged singles are looked up during staging -- one more round trip there, all lookups of a
// tile in flight together -- so that the emission loop works from shared memory alone (looked up per coincidence inside
// that loop, two dependent round trips per pair and thread, the kernel took 37 instead of 27 us).
__global__ void __launch_bounds__(kThreads) k_coinc(const EventRec* __restrict__ s, const double* __restrict__ stime,
                                                    const int* __restrict__ span, const int* __restrict__ spar,
                                                    const int* __restrict__ seid, DigitizerDev p, unsigned* __restrict__ counters,
                                                    unsigned singles_cap, unsigned* __restrict__ status,
                                                    gpet_coincidence* __restrict__ out, uint2* __restrict__ pairs, unsigned cap,
                                                    const unsigned* __restrict__ base_in, unsigned* __restrict__ base_out,
                                                    unsigned char* __restrict__ cls) {
    pdl_wait();
    __shared__ unsigned s_tile;
    __shared__ unsigned s_cls[3];
    unsigned n_cls[3] = {0u, 0u, 0u};   // this thread's true / scatter / random coincidences
    if (threadIdx.x < 3) s_cls[threadIdx.x] = 0u;   // ordered before its use by the barriers of the tile loop and of the tally
    __shared__ double s_t[kScanTile + 2 * kHalo];
    __shared__ int s_p[kScanTile + 2 * kHalo];
    __shared__ int s_eid[kScanTile + 2 * kHalo];
    __shared__ unsigned char s_cf[kScanTile + 2 * kHalo];
    __shared__ __align__(16) unsigned short s_cnt[kScanTile];
    const unsigned n = min(counters[3], singles_cap);
    const unsigned ntiles = (n + kScanTile - 1) / kScanTile;
    const double W = (double)p.cwin;
    const unsigned pair_base = base_in ? *base_in : 0u;
    if (base_out && blockIdx.x == 0 && threadIdx.x == 0) *base_out = pair_base + n;
#ifdef GPET_PHASE_TRACE
    long long ph[8]; int nph = 0;
    PH();
#endif
    while (true) {
        if (threadIdx.x == 0) s_tile = atomicAdd(&counters[7], 1u);
        __syncthreads();
        const unsigned tile = s_tile;
        if (tile >= ntiles) break;
        PH();
        const unsigned at = tile * kScanTile;
        SinglesView v;
        v.gt = stime; v.gp = span; v.st = s_t; v.sp = s_p;
        v.lo = at >= (unsigned)kHalo ? at - kHalo : 0u;
        v.hi = min(at + kScanTile + kHalo, n);
        {   // all loads first, then the shared-memory stores
            constexpr int kRounds = (kScanTile + 2 * kHalo + kThreads - 1) / kThreads;
            double tv[kRounds];
            int pv[kRounds], ev[kRounds], pn[kRounds];
#pragma unroll
            for (int r = 0; r < kRounds; r++) {
                const unsigned i = v.lo + r * kThreads + threadIdx.x;
                tv[r] = i < v.hi ? stime[i] : 0.0;
                pv[r] = i < v.hi ? span[i] : 0;
                ev[r] = i < v.hi ? seid[i] : 0;
                pn[r] = i < v.hi ? spar[i] : -1;
            }
#pragma unroll
            for (int r = 0; r < kRounds; r++) {
                const unsigned i = v.lo + r * kThreads + threadIdx.x;
                if (i < v.hi) {
                    s_t[i - v.lo] = tv[r];
                    s_p[i - v.lo] = pv[r];
                    s_eid[i - v.lo] = ev[r];
                }
            }
            // second round trip: the scatter tags of all staged singles.  (Looking up only the singles with a neighbour
            // closer than the window -- half of them on the shipped example -- was measured: no difference, 33.1 us.)
            unsigned cf[kRounds];
#pragma unroll
            for (int r = 0; r < kRounds; r++) cf[r] = class_flags(p, pn[r]);
#pragma unroll
            for (int r = 0; r < kRounds; r++) {
                const unsigned i = v.lo + r * kThreads + threadIdx.x;
                if (i < v.hi) s_cf[i - v.lo] = (unsigned char)cf[r];
            }
        }
        __syncthreads();
        PH();
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const unsigned a = at + k * kThreads + threadIdx.x;
            const unsigned c = a < n ? coincidences_of(v, a, n, W, p, counters) : 0u;
            s_cnt[k * kThreads + threadIdx.x] = (unsigned short)min(c, 0xffffu);
        }
        __syncthreads();
        PH();
        const unsigned a0 = at + threadIdx.x * 8;
        unsigned c[8];
        {
            const uint4 c8 = reinterpret_cast<const uint4*>(s_cnt)[threadIdx.x];
            c[0] = c8.x & 0xffffu; c[1] = c8.x >> 16; c[2] = c8.y & 0xffffu; c[3] = c8.y >> 16;
            c[4] = c8.z & 0xffffu; c[5] = c8.z >> 16; c[6] = c8.w & 0xffffu; c[7] = c8.w >> 16;
        }
        TileScan sc = tile_exclusive_scan(c, tile, status);
        PH();
        if (tile == ntiles - 1 && threadIdx.x == 0) counters[4] = sc.tile_excl + sc.tile_total;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (c[k] == 0) continue;
            const unsigned a = a0 + k;
            unsigned o = sc.excl[k], left = c[k];
            const double tend = v.t(a) + W;
            // annihilation number and class flags of a single: staged, or (window beyond the halo) from global memory
            auto info = [&](unsigned i, int& eid, unsigned& cf) {
                if (i >= v.lo && i < v.hi) { eid = s_eid[i - v.lo]; cf = s_cf[i - v.lo]; }
                else { eid = seid[i]; cf = class_flags(p, spar[i]); }
            };
            int eid_a; unsigned cf_a;
            info(a, eid_a, cf_a);
            for (unsigned b = a + 1; left && b < n && v.t(b) < tend; b++) {
                if (!pair_ok(v, a, b, p)) continue;
                left--;
                int eid_b; unsigned cf_b;
                info(b, eid_b, cf_b);
                const unsigned cl = coincidence_class(p, eid_a, cf_a, eid_b, cf_b);
                n_cls[0] += cl == 0u ? 1u : 0u; n_cls[1] += cl == 1u ? 1u : 0u; n_cls[2] += cl == 2u ? 1u : 0u;
                if (o < cap) {
                    if (pairs) pairs[o] = make_uint2(pair_base + a, pair_base + b);
                    if (cls) cls[o] = (unsigned char)cl;
                    if (out) {   // 2 x 48-byte records copied as 6 x 16 B from the singles list
                        const int4* pa = reinterpret_cast<const int4*>(s + a);
                        const int4* pb = reinterpret_cast<const int4*>(s + b);
                        const int4 a0 = __ldg(pa), a1 = __ldg(pa + 1), a2 = __ldg(pa + 2);
                        const int4 b0 = __ldg(pb), b1 = __ldg(pb + 1), b2 = __ldg(pb + 2);
                        int4* po = reinterpret_cast<int4*>(out + o);
                        po[0] = a0; po[1] = a1; po[2] = a2; po[3] = b0; po[4] = b1; po[5] = b2;
                    }
                }
                o++;
            }
        }
        __syncthreads();   // shared arrays are reused by the next tile
        PH();
    }
#ifdef GPET_PHASE_TRACE
    if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 170 || blockIdx.x == 340) && nph >= 6)
        printf("coinc block %d: ticket %lld stage %lld count %lld scan %lld emit %lld cycles\n", blockIdx.x, ph[1] - ph[0], ph[2] - ph[1],
               ph[3] - ph[2], ph[4] - ph[3], ph[5] - ph[4]);
#endif
    // class tallies: one add per block and class (counters[12..14] = trues, scatters, randoms)
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const unsigned w = warp_sum(n_cls[k]);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_cls[k], w);
    }
    __syncthreads();
    if (threadIdx.x < 3 && s_cls[threadIdx.x]) atomicAdd(&counters[12 + threadIdx.x], s_cls[threadIdx.x]);
}

}  // namespace

// ================================================================================================ launchers
size_t sort_state_bytes() { return sizeof(rsort::SortState); }
size_t sort_lookback_words(size_t capacity) { return ((capacity + rsort::kTile - 1) / rsort::kTile) * (size_t)rsort::kBins; }
unsigned scan_tiles(size_t capacity) { return (unsigned)((capacity + kScanTile - 1) / kScanTile); }
unsigned scan_status_stride() { return kStatusStride; }
unsigned bucket_words() { return kMaxBuckets; }

// Programmatic dependent launch: the kernel may be set up and its blocks made resident while its predecessor in the
// stream is still draining; every kernel launched this way starts with pdl_wait() (griddepcontrol.wait), which returns
// when the predecessor has completed and its writes are visible.  Saves the launch latency at each of the digitizer's
// kernel boundaries (short, latency-bound kernels).
template <typename... KArgs, typename... Args>
void launch_pdl(void (*kernel)(KArgs...), int grid, int block, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

TimeRange time_range_us(double t_lo_us, double t_hi_us) {
    TimeRange r;
    r.lo = t_lo_us;
    r.hi = t_hi_us;
    r.dev = nullptr;
    return r;
}

// The frame's counters to a pinned host block by zero-copy stores: 32 counter words, then the hot counters two by two
// (same layout as fetch_counters_async in abi.cu).  A cudaMemcpyAsync on the compute stream would wait for the D2H
// copy engine, which the previous frame's singles keep busy for hundreds of microseconds.
__global__ void k_publish_counters(const unsigned* __restrict__ counters, const unsigned* __restrict__ hot, unsigned* __restrict__ h_dst) {
    pdl_wait();
    const unsigned i = threadIdx.x;
    if (i < 32) h_dst[i] = counters[i];
    else if (i < 32 + 2 * kHotLines) h_dst[i] = hot[((i - 32) >> 1) * kHotStride + ((i - 32) & 1)];
}

int launch_publish_counters(const unsigned* counters, const unsigned* hot, unsigned* h_dst, cudaStream_t s) {
    GPET_LAUNCH("k_publish_counters", s, launch_pdl(k_publish_counters, 1, 64, s, counters, hot, h_dst));
    return 1;
}

// gpet_mark_scattered: scatter tags for a caller's list of photon numbers (replayed events carry no transport history)
__global__ void k_mark_scattered(const int* __restrict__ parn, unsigned n, unsigned char* __restrict__ tag, unsigned mask, unsigned serial) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (parn[i] != -1) tag[(unsigned)parn[i] & mask] = (unsigned char)serial;
}

int launch_mark_scattered(const int* parn, unsigned n, unsigned char* tag, unsigned mask, unsigned serial, cudaStream_t s) {
    if (n == 0 || tag == nullptr) return 0;
    const int grid = (int)std::min<unsigned>((n + kThreads - 1) / kThreads, 1024u);
    GPET_LAUNCH("k_mark_scattered", s, k_mark_scattered<<<grid, kThreads, 0, s>>>(parn, n, tag, mask, serial));
    return 1;
}

int launch_noise(EventBuf ev, const DigitizerDev& p, double t_lo_us, double t_hi_us, uint64_t seed, int num_sms, cudaStream_t s) {
    if (!(p.noise_gap > 0.f) || !(p.noise_interval > 0.f) || !(t_hi_us > t_lo_us)) return 0;
    const double iv = (double)p.noise_interval;
    const long long id0 = (long long)floor(t_lo_us / iv), id1 = (long long)ceil(t_hi_us / iv);
    const long long n = std::max<long long>(id1 - id0, 1);
    const int grid = (int)std::min<long long>((n + kThreads - 1) / kThreads, (long long)num_sms * 8);
    GPET_LAUNCH("k_noise", s, k_noise<<<grid, kThreads, 0, s>>>(ev, p, seed, t_lo_us, t_hi_us, id0, n));
    return 1;
}

int launch_digitize(EventBuf ev, const DigitizerOut& out, const DigitizerDev& p, DigitizerWorkspace& ws, const TimeRange* range,
                    uint64_t seed, int num_sms, cudaStream_t s, bool reset, bool with_fallback) {
    // One wave per kernel: a grid of exactly the resident blocks.  With 592 blocks everywhere, k_emit_singles (104
    // registers, 2 blocks per SM) ran the last 47 of a frame's 343 tiles in a second wave, 28 + 15 us.  GPET_DIGI_GRID
    // overrides the blocks per SM for tuning runs.
    static const int grid_mult = [] { const char* v = getenv("GPET_DIGI_GRID"); return v && *v ? atoi(v) : 0; }();
    auto resident = [&](const void* kernel, int dflt) {
        if (grid_mult > 0) return num_sms * grid_mult;
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1) per_sm = dflt;
        return num_sms * per_sm;
    };
    // launch shapes depend on the device (SM count): cached per device, so that contexts on different GPUs of one process never
    // share a grid sized for another device (k_bucket_scan and the cooperative fallback rely on all their blocks being resident)
    struct Shapes { int prep, scatter, rank, emit, coinc, chain, coop, pack; };
    static Shapes shapes[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    Shapes& sh = shapes[dev >= 0 && dev < 64 ? dev : 0];
    if (!sh.prep) {
        sh.prep = resident((const void*)k_prep, 4);
        sh.scatter = resident((const void*)k_bucket_scatter, 4);
        sh.rank = resident((const void*)k_bucket_rank, 4);
        sh.emit = resident((const void*)k_emit_singles, 3);
        sh.coinc = resident((const void*)k_coinc, 4);
        sh.chain = resident((const void*)k_deadtime_chain, 4);
        sh.pack = 8 * num_sms;
        int per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lsd_fallback, rsort::kThreads, 0);
        sh.coop = num_sms * std::max(1, std::min(per_sm, 2));
    }
    const int g_prep = sh.prep, g_scatter = sh.scatter, g_rank = sh.rank, g_emit = sh.emit, g_coinc = sh.coinc, g_chain = sh.chain;
    const int grid = num_sms * 4;   // k_range
    int launches = 0;
    EventRec* singles = static_cast<EventRec*>(out.singles);
    unsigned long long* keys = ws.tkeys[0];    // by event index; after the sort: the sorted keys (tsort)
    if (reset) {   // the digitizer's share of the frame state: counters[0..7], then everything behind the counter block
        cudaMemsetAsync(ws.counters, 0, 8 * sizeof(unsigned), s);
        cudaMemsetAsync(ws.counters + 10, 0, 6 * sizeof(unsigned), s);   // emit-window counts, coincidence class tallies, halo flag
        cudaMemsetAsync(ws.counters + 64, 0, (size_t)(ws.hot - (ws.counters + 64)) * sizeof(unsigned), s);   // not the hot block: it holds the event count
    }
    TimeRange tr;
    if (range) {
        tr = *range;
    } else {
        GPET_LAUNCH("k_range", s, k_range<<<grid, kThreads, 0, s>>>(ev, ws.minmax));
        launches++;
        tr.lo = 0.0; tr.hi = 0.0; tr.dev = ws.minmax;
    }
    GPET_LAUNCH("k_prep", s, launch_pdl(k_prep, g_prep, kThreads, s, ev, p, seed, tr, keys, ws.site_of, ws.aux, ws.bcount, ws.counters));
    GPET_LAUNCH("k_bucket_scan", s, launch_pdl(k_bucket_scan, (int)(kMaxBuckets / kScanTile), kThreads, s, ws.bcount, ws.bstart, ws.scan_status[2],
                                                ws.counters, tr));
    GPET_LAUNCH("k_bucket_scatter", s, launch_pdl(k_bucket_scatter, g_scatter, kThreads, s, keys, ws.site_of, ws.aux, ws.counters, tr, ws.bstart,
                                                   ws.bent));
    GPET_LAUNCH("k_bucket_rank", s, launch_pdl(k_bucket_rank, g_rank, kThreads, s, ws.bent, ws.bstart, ws.counters, tr, keys, ws.order_t,
                                                ws.site_t, p.tie_site));
    launches += 4;
    if (with_fallback) {
        const int coop_grid = sh.coop;
        void* args[] = {&ws.tkeys[0], &ws.tkeys[1], &ws.tvals[0], &ws.tvals[1], &ws.aux, &ws.site_of, &ws.counters, &ws.st_time,
                        &ws.lookback[0], &ws.lookback[1], &ws.grid_bar, &ws.order_t, &ws.site_t};
        GPET_LAUNCH("k_lsd_fallback", s,
                    cudaLaunchCooperativeKernel((const void*)k_lsd_fallback, dim3(coop_grid), dim3(rsort::kThreads), args, 0, s));
        launches++;
    }
    if (p.dtype != 0) {
        GPET_LAUNCH("k_deadtime_chain", s, k_deadtime_chain<<<g_chain, kThreads, 0, s>>>(p, keys, ws.site_t, ws.kill, ws.counters));
        launches++;
    }
    GPET_LAUNCH("k_emit_singles", s, launch_pdl(k_emit_singles, g_emit, kThreads, s, ev, p, singles, out.singles_cap, keys, ws.order_t, ws.site_t,
                                                ws.kill, ws.counters, ws.scan_status[0], ws.stime, ws.span, ws.spar, ws.seid, ws.spectrum, ws.spectrum_bins,
                                                ws.spectrum_stride, ws.spec_emin, ws.spec_emax, with_fallback ? 0 : 1,
                                                out.ev_after_emit ? out.h_singles_count : (unsigned*)nullptr));
    launches++;
    if (out.singles_compact) {
        GPET_LAUNCH("k_pack_singles", s, launch_pdl(k_pack_singles, sh.pack, kThreads, s, singles, ws.counters, out.singles_cap,
                                                     static_cast<int4*>(out.singles_compact)));
        launches++;
    }
    if (out.ev_after_emit && out.h_singles_count) cudaEventRecord(out.ev_after_emit, s);
    if (p.cwin > 0.f && (out.coinc || out.pairs)) {
        GPET_LAUNCH("k_coinc", s, launch_pdl(k_coinc, g_coinc, kThreads, s, singles, ws.stime, ws.span, ws.spar, ws.seid, p, ws.counters, out.singles_cap, ws.scan_status[1],
                                                                  static_cast<gpet_coincidence*>(out.coinc), static_cast<uint2*>(out.pairs),
                                                                  out.coinc_cap, out.pair_base_in, out.pair_base_out,
                                                                  static_cast<unsigned char*>(out.cls)));
        launches++;
    }
    return launches;
}

}  // namespace gpet
