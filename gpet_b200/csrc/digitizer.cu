// Digitizer chain on the device: blur -> thresholder -> time sort -> site order -> dead time -> energy window
// -> singles (time sorted) [-> coincidence sorter].  Reference: blur/energywindow/setSitenum/deadtime kernels
// (gPET_kernals.cu:607-698, 814-837) and the host orchestration with three CPU sorts (gPET.cu:385-424,
// detector.cu:354-385).  Here nothing leaves the device between stages: counts stay in `counters`, events are 48-byte
// records in the file layout, and the final singles list is produced by one fused flag + scan + compaction of the
// time order (no re-sort after dead time / energy window, since killing keeps the order).  The launch sequence is
// static: all sizes and decisions live on the device.
//
// Time sort (D5).  The keys are the order-preserving u64 images of the fp64 times.  Decay times inside a frame are
// spread over the frame, so the sort is a bucket sort: k_prep finds the key range, k_bucket_count/_scan/_scatter
// distribute the events over 2^ceil(log2(n/8)) equal slices of that range (about 8 events each), k_bucket_sort ranks
// every event inside its slice by (key, event index) -- four short kernels without ping-pong passes or chained scans,
// and the result is the unique stable order whatever the scatter order was.  If some slice holds more than
// kBucketLimit events (times clustered on a scale far below the key range: never the case for decay data, but legal
// input of the replay entry point), k_bucket_scan raises counters[12] and the stable LSD radix sort of radix_sort.cuh
// runs instead; its kernels are always enqueued and return at once when the flag is clear.
//
// Launches per frame: k_begin, k_prep, k_bucket_count, k_bucket_scan, k_bucket_scatter, k_bucket_sort, [k_lsd_hist,
// 8 x k_onesweep<u64>: fallback, normally empty], k_site_keys, 4 x k_onesweep<u32> (passes whose digit is constant
// return at once), k_deadtime, k_emit_singles [, k_coinc_count, k_coinc_emit].
#include <cstdlib>
#include "kernels.hpp"
#include "philox.cuh"
#include "ktimer.hpp"
#include "radix_sort.cuh"

#include "../../include/gpet_b200.h"

namespace gpet {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxLogBuckets = 17;
constexpr unsigned kMaxBuckets = 1u << kMaxLogBuckets;
constexpr unsigned kBucketLimit = 1024;   // a fuller slice sends the time sort to the LSD fallback
constexpr int kFlagLsd = 12;              // counters[kFlagLsd] != 0: time sort by LSD radix passes

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

// Equal slices of the key range [kmin, kmax] of the alive events: slice = (key - kmin) >> shift.
struct BucketMap {
    unsigned long long kmin;
    int shift;
    unsigned nb;
};

__device__ __forceinline__ BucketMap bucket_map(const unsigned long long* __restrict__ minmax, unsigned n_alive) {
    BucketMap m;
    const unsigned long long kmin = minmax[0], kmax = minmax[1];
    m.kmin = kmin;
    const unsigned long long range = (n_alive && kmax >= kmin) ? kmax - kmin : 0ull;
    const int bits = range ? 64 - __clzll((long long)range) : 0;
    int lognb = (n_alive > 1 ? 32 - __clz((int)(n_alive - 1)) : 0) - 3;   // ceil(log2 n) - 3: about 8 events per slice
    lognb = min(max(lognb, 6), kMaxLogBuckets);
    m.shift = max(bits - lognb, 0);
    m.nb = 1u << lognb;
    return m;
}

__device__ __forceinline__ unsigned bucket_of(const BucketMap& m, unsigned long long key) {
    return (unsigned)((key - m.kmin) >> m.shift);
}

// ------------------------------------------------------------------------------------------- stage 0: reset
// counters[0..7] and the fallback flag, both LSD sort states, the scan status words, the slice counters, the key range.
__global__ void __launch_bounds__(kThreads) k_begin(unsigned* __restrict__ counters, rsort::SortState* st_time,
                                                    rsort::SortState* st_site, unsigned* __restrict__ scan_status0,
                                                    unsigned* __restrict__ scan_status1, unsigned* __restrict__ scan_status2,
                                                    unsigned max_tiles, unsigned* __restrict__ bcount,
                                                    unsigned long long* __restrict__ minmax) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (tid < 8) counters[tid] = 0;
    if (tid == 8) counters[kFlagLsd] = 0;
    if (tid == 9) { minmax[0] = ~0ull; minmax[1] = 0ull; }
    constexpr unsigned kWords = sizeof(rsort::SortState) / 4;
    unsigned* a = reinterpret_cast<unsigned*>(st_time);
    unsigned* b = reinterpret_cast<unsigned*>(st_site);
    for (unsigned i = tid; i < kWords; i += nth) { a[i] = 0; b[i] = 0; }
    for (unsigned i = tid; i < max_tiles; i += nth) { scan_status0[i] = 0; scan_status1[i] = 0; }
    for (unsigned i = tid; i < kMaxBuckets / 2048u; i += nth) scan_status2[i] = 0;
    for (unsigned i = tid; i < kMaxBuckets; i += nth) bcount[i] = 0;
}

// ------------------------------------------------------------------------------------------- stage 1: blur + thresholder + time keys
// blur (gPET_kernals.cu:814-837) + energywindow(Eth, 2e6) (gPET.cu:393) fused; writes the time key of every record
// (all ones for a dead one) and the key range of the alive ones.
__global__ void __launch_bounds__(kThreads) k_prep(EventBuf ev, DigitizerDev p, uint64_t seed,
                                                   unsigned long long* __restrict__ keys, unsigned* __restrict__ counters,
                                                   unsigned long long* __restrict__ minmax) {
    const unsigned n = min(*ev.count, ev.capacity);
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[0] = n;
    unsigned alive_cnt = 0;
    unsigned long long kmin = ~0ull, kmax = 0ull;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        EventRec* rec = ev.rec + i;
        const int4 b4 = reinterpret_cast<const int4*>(rec)[1];   // siten, eventid, t
        float4 c4 = reinterpret_cast<const float4*>(rec)[2];     // E, x, y, z
        float E = c4.x;
        double t = __longlong_as_double((long long)(((unsigned long long)(unsigned)b4.w << 32) | (unsigned)b4.z));
        float R = 0.f;
        // float / double mix exactly as the reference expression is typed
        if (p.blur_policy == 0) R = __fmul_rn(__fsqrt_rn(__fdiv_rn(p.Eref, E)), p.Rref);
        if (p.blur_policy == 1)
            R = (float)__dadd_rn((double)p.Rref, __ddiv_rn((double)__fmul_rn(p.slope, __fsub_rn(E, p.Eref)), 1e6));
        if (!(R > 0.f)) R = 0.f;
        // R == 0 leaves E bit-identical (E + 0), so the draw is skipped: this is the deterministic replay mode
        if (R > 0.f || p.sblur > 0.f || p.tblur > 0.f) {
            const int parn = reinterpret_cast<const int*>(rec)[0];
            Philox rng(seed, (uint64_t)(uint32_t)parn, ((uint32_t)kStageBlur << 24) | ((uint32_t)b4.x & 0xFFFFFFu));
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
            alive_cnt++;
            kmin = min(kmin, key);
            kmax = max(kmax, key);
        }
    }
    alive_cnt = warp_sum(alive_cnt);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((threadIdx.x & 31) == 0 && alive_cnt) {
        atomicAdd(&counters[1], alive_cnt);
        atomicMin(&minmax[0], kmin);
        atomicMax(&minmax[1], kmax);
    }
}

// ------------------------------------------------------------------------------------------- single-pass exclusive scan
// Tiles of 2048 elements (thread = 8 consecutive elements) claimed in order from a device counter; the running
// total travels from tile to tile through one status word per tile (aggregate / inclusive-prefix flags, 30-bit
// values), so flagging, scanning and compacting happen in ONE kernel.
constexpr int kScanTile = 2048;
constexpr int kSpecSmemBins = 1024;

struct TileScan {
    unsigned excl[8];    // exclusive prefix of each of the thread's 8 elements (global)
    unsigned tile_total; // sum over the tile
    unsigned tile_excl;  // sum over all earlier tiles
};

__device__ __forceinline__ TileScan tile_exclusive_scan(const unsigned v[8], unsigned tile, unsigned* __restrict__ status) {
    __shared__ unsigned ws[kThreads / 32];
    __shared__ unsigned s_excl;
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
    if (threadIdx.x == 0) {
        unsigned excl = 0;
        if (tile == 0) {
            rsort::st_volatile(&status[0], rsort::kFlagPrefix | total);
        } else {
            rsort::st_volatile(&status[tile], rsort::kFlagAggregate | total);
            excl = rsort::lookback_sum(status, 1, tile);
            rsort::st_volatile(&status[tile], rsort::kFlagPrefix | (excl + total));
        }
        s_excl = excl;
    }
    __syncthreads();
    TileScan r;
    r.tile_total = total;
    r.tile_excl = s_excl;
    unsigned e = s_excl + wprefix + (x - s);
#pragma unroll
    for (int k = 0; k < 8; k++) { r.excl[k] = e; e += v[k]; }
    __syncthreads();  // ws / s_excl are reused by the next tile
    return r;
}

// ------------------------------------------------------------------------------------------- stage 2: time sort (bucket sort)
__global__ void __launch_bounds__(kThreads) k_bucket_count(const unsigned long long* __restrict__ keys,
                                                           const unsigned* __restrict__ counters,
                                                           const unsigned long long* __restrict__ minmax,
                                                           unsigned* __restrict__ bcount) {
    const unsigned n = counters[0];
    const BucketMap m = bucket_map(minmax, counters[1]);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        if (key != ~0ull) atomicAdd(&bcount[bucket_of(m, key)], 1u);
    }
}

// exclusive scan of the slice counters (one tile of 2048 per block, all blocks resident: tile = blockIdx);
// raises the LSD-fallback flag when a slice is overfull
__global__ void __launch_bounds__(kThreads) k_bucket_scan(const unsigned* __restrict__ bcount, unsigned* __restrict__ bstart,
                                                          unsigned* __restrict__ bcur, unsigned* __restrict__ status,
                                                          unsigned* __restrict__ counters,
                                                          const unsigned long long* __restrict__ minmax) {
    const BucketMap m = bucket_map(minmax, counters[1]);
    const unsigned tile = blockIdx.x;
    if (tile * kScanTile >= m.nb) return;
    const unsigned b0 = tile * kScanTile + threadIdx.x * 8;
    unsigned c[8];
    bool over = false;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        c[k] = (b0 + k < m.nb) ? bcount[b0 + k] : 0u;
        over |= c[k] > kBucketLimit;
    }
    if (over) counters[kFlagLsd] = 1u;
    TileScan sc = tile_exclusive_scan(c, tile, status);
#pragma unroll
    for (int k = 0; k < 8; k++)
        if (b0 + k < m.nb) { bstart[b0 + k] = sc.excl[k]; bcur[b0 + k] = sc.excl[k]; }
}

__global__ void __launch_bounds__(kThreads) k_bucket_scatter(const unsigned long long* __restrict__ keys,
                                                             const unsigned* __restrict__ counters,
                                                             const unsigned long long* __restrict__ minmax,
                                                             unsigned* __restrict__ bcur, unsigned long long* __restrict__ bkeys,
                                                             unsigned* __restrict__ bidx) {
    if (counters[kFlagLsd]) return;
    const unsigned n = counters[0];
    const BucketMap m = bucket_map(minmax, counters[1]);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        if (key == ~0ull) continue;
        const unsigned pos = atomicAdd(&bcur[bucket_of(m, key)], 1u);
        bkeys[pos] = key;
        bidx[pos] = i;
    }
}

// rank of every event inside its slice by (key, event index): the stable time order
__global__ void __launch_bounds__(kThreads) k_bucket_sort(const unsigned long long* __restrict__ bkeys,
                                                          const unsigned* __restrict__ bidx, const unsigned* __restrict__ bstart,
                                                          const unsigned* __restrict__ bend, const unsigned* __restrict__ counters,
                                                          const unsigned long long* __restrict__ minmax,
                                                          unsigned* __restrict__ order_t, unsigned long long* __restrict__ tsort) {
    if (counters[kFlagLsd]) return;
    const unsigned n1 = counters[1];
    const BucketMap m = bucket_map(minmax, n1);
    for (unsigned pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n1; pos += gridDim.x * blockDim.x) {
        const unsigned long long key = bkeys[pos];
        const unsigned i = bidx[pos];
        const unsigned b = bucket_of(m, key);
        const unsigned s = bstart[b], e = bend[b];
        unsigned rank = 0;
        for (unsigned q = s; q < e; q++) {
            const unsigned long long kq = bkeys[q];
            if (kq < key || (kq == key && bidx[q] < i)) rank++;
        }
        order_t[s + rank] = i;
        tsort[s + rank] = key;
    }
}

// fallback only: digit histograms of all 8 passes + identity payload for the LSD radix sort of the time keys
__global__ void __launch_bounds__(kThreads) k_lsd_hist(const unsigned long long* __restrict__ keys, unsigned* __restrict__ vals,
                                                       const unsigned* __restrict__ counters, rsort::SortState* st_time,
                                                       unsigned* __restrict__ lookback0) {
    if (counters[kFlagLsd] == 0u) return;
    __shared__ unsigned sh_hist[8 * rsort::kBins];
    const unsigned n = counters[0];
    for (int i = threadIdx.x; i < 8 * rsort::kBins; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        vals[i] = i;
        rsort::hist_add<unsigned long long, 8>(sh_hist, keys[i]);
    }
    __syncthreads();
    rsort::hist_flush<8>(sh_hist, st_time);
    rsort::clear_lookback(lookback0, n);
}

// ------------------------------------------------------------------------------------------- stage 3: site keys
// setSitenum (gPET_kernals.cu:607-640) fused with building the (site) sort keys over the time order.  After this
// kernel the time order is in order_t / tsort whichever sort produced it.
__global__ void __launch_bounds__(kThreads) k_site_keys(EventBuf ev, DigitizerDev p, unsigned long long* __restrict__ tkeys0,
                                                        const unsigned long long* __restrict__ tkeys1,
                                                        const unsigned* __restrict__ tvals0, const unsigned* __restrict__ tvals1,
                                                        const rsort::SortState* st_time, unsigned* __restrict__ order_t,
                                                        unsigned* __restrict__ keys, unsigned* __restrict__ vals,
                                                        const unsigned* __restrict__ counters, rsort::SortState* st_site,
                                                        unsigned* __restrict__ lookback0) {
    __shared__ unsigned sh_hist[4 * rsort::kBins];
    const unsigned n1 = counters[1];
    const bool lsd = counters[kFlagLsd] != 0u;
    const unsigned cur = rsort::current_buffer(st_time, 8, counters[0]);
    const unsigned* __restrict__ t_sorted_vals = cur ? tvals1 : tvals0;
    for (int i = threadIdx.x; i < 4 * rsort::kBins; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < n1; j += gridDim.x * blockDim.x) {
        unsigned i;
        if (lsd) {
            i = t_sorted_vals[j];
            order_t[j] = i;
            if (cur) tkeys0[j] = tkeys1[j];
        } else {
            i = order_t[j];
        }
        EventRec* rec = ev.rec + i;
        int site;
        switch (p.dlevel) {
            case 0: site = 0; break;
            case 1: site = rec->pann; break;
            case 2: site = rec->pann * p.moduleN + rec->modn; break;
            default: site = rec->siten; break;  // dlevel == 3: keep what readout left (gPET.cu:402-407)
        }
        if (p.dlevel >= 0 && p.dlevel <= 2) rec->siten = site;
        // flip the sign bit: std::sort compares siten as signed int (gPET.h:101-106)
        const unsigned key = (unsigned)site ^ 0x80000000u;
        keys[j] = key;
        vals[j] = j;
        rsort::hist_add<unsigned, 4>(sh_hist, key);
    }
    __syncthreads();
    rsort::hist_flush<4>(sh_hist, st_site);
    rsort::clear_lookback(lookback0, n1);
}

// ------------------------------------------------------------------------------------------- stage 4: dead time
// deadtime (gPET_kernals.cu:657-698) with the snapshot-start semantics of SURVEY 8(a) D7: every decision uses the
// original times; `tdead` is fp32 and `tdead + interval` is an fp32 sum, as in the reference.  q runs over the
// (site, t) order; kill flags are stored by position in the time order.
__global__ void __launch_bounds__(kThreads) k_deadtime(DigitizerDev p, const unsigned long long* __restrict__ tsort,
                                                       const unsigned* __restrict__ skeys0, const unsigned* __restrict__ skeys1,
                                                       const unsigned* __restrict__ svals0, const unsigned* __restrict__ svals1,
                                                       const rsort::SortState* st_site, unsigned char* __restrict__ kill,
                                                       const unsigned* __restrict__ counters) {
    const unsigned n1 = counters[1];
    const float tau = p.dtime;
    const unsigned cur = rsort::current_buffer(st_site, 4, n1);
    const unsigned* __restrict__ site_keys = cur ? skeys1 : skeys0;
    const unsigned* __restrict__ order_s = cur ? svals1 : svals0;
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n1; q += gridDim.x * blockDim.x) {
        const unsigned j = order_s[q];
        const double t = key_time(tsort[j]);
        bool same_prev = false;
        double tprev = 0.0;
        if (q > 0 && site_keys[q] == site_keys[q - 1]) {
            same_prev = true;
            tprev = key_time(tsort[order_s[q - 1]]);
        }
        // "killable by its predecessor": t < (float)t_prev + tau with the fp32 sum of the reference (tdead is float)
        const bool killable = same_prev && t < (double)__fadd_rn((float)tprev, tau);
        if (p.dtype == 0) {
            // paralyzable: tdead follows every event, so the predicate is predecessor-local
            kill[j] = killable ? 1 : 0;
        } else {
            // non-paralyzable: sequential anchor chain per site.  An event its predecessor cannot kill survives any
            // earlier anchor as well (fp32 rounding and the fp32 sum are monotone), so it is a guaranteed anchor and
            // the chain can be cut there: one thread per such run start, runs are short at realistic rates.
            if (killable) continue;
            kill[j] = 0;
            float tdead = (float)t;
            unsigned r = q + 1;
            while (r < n1 && site_keys[r] == site_keys[q]) {
                const unsigned jr = order_s[r];
                const double tr = key_time(tsort[jr]);
                const double tr_prev = key_time(tsort[order_s[r - 1]]);
                if (!(tr < (double)__fadd_rn((float)tr_prev, tau))) break;  // next run start
                if (tr < (double)__fadd_rn(tdead, tau)) {
                    kill[jr] = 1;
                } else {
                    kill[jr] = 0;
                    tdead = (float)tr;
                }
                r++;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------- stage 5: energy window + compaction -> singles
// energywindow(Ewinmin, Ewinmax) (gPET.cu:418) over the survivors of the dead time, in time order.
__global__ void __launch_bounds__(kThreads) k_emit_singles(EventBuf ev, DigitizerDev p, EventRec* __restrict__ singles,
                                                           unsigned singles_cap, const unsigned* __restrict__ order_t,
                                                           const unsigned char* __restrict__ kill, unsigned* __restrict__ counters,
                                                           unsigned* __restrict__ status, unsigned long long* __restrict__ spectrum,
                                                           int nbins, float emin, float emax) {
    __shared__ unsigned s_tile;
    __shared__ unsigned s_idx[kScanTile];
    __shared__ unsigned s_spec[kSpecSmemBins];   // block-private energy histogram (flushed once per block)
    const unsigned n1 = counters[1];
    const unsigned ntiles = (n1 + kScanTile - 1) / kScanTile;
    const bool spec_smem = spectrum && nbins > 0 && nbins <= kSpecSmemBins;
    if (spec_smem)
        for (int b = threadIdx.x; b < nbins; b += blockDim.x) s_spec[b] = 0;
    unsigned c2 = 0;
    while (true) {
        if (threadIdx.x == 0) s_tile = atomicAdd(&counters[6], 1u);
        __syncthreads();
        const unsigned tile = s_tile;
        if (tile >= ntiles) break;
        const unsigned j0 = tile * kScanTile + threadIdx.x * 8;
        unsigned idx[8], flag[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const unsigned j = j0 + k;
            flag[k] = 0; idx[k] = 0;
            if (j < n1) {
                const unsigned i = order_t[j];
                idx[k] = i;
                const bool a2 = kill[j] == 0;
                const float E = ev.rec[i].E;
                flag[k] = (a2 && !(E < p.Ewinmin || E > p.Ewinmax)) ? 1u : 0u;
                c2 += a2 ? 1u : 0u;
            }
        }
        TileScan sc = tile_exclusive_scan(flag, tile, status);
        if (tile == ntiles - 1 && threadIdx.x == 0) counters[3] = sc.tile_excl + sc.tile_total;
        // survivors of this tile, in order, through shared memory: the emission below is then one record per thread
        // and iteration, written as three 16-byte vectors to consecutive addresses
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (flag[k]) s_idx[sc.excl[k] - sc.tile_excl] = idx[k];
        __syncthreads();
        for (unsigned r = threadIdx.x; r < sc.tile_total; r += 2 * kThreads) {
            const unsigned r2 = r + kThreads;
            const bool two = r2 < sc.tile_total;
            const unsigned o = sc.tile_excl + r, o2 = sc.tile_excl + r2;
            const EventRec a = load_event_rec(ev.rec + s_idx[r]);
            EventRec b = a;
            if (two) b = load_event_rec(ev.rec + s_idx[r2]);
            if (o < singles_cap) store_event_rec(singles + o, a);
            if (two && o2 < singles_cap) store_event_rec(singles + o2, b);
            if (spectrum && nbins > 0) {
                float f = (a.E - emin) / (emax - emin) * nbins;
                if (f >= 0.f && f < (float)nbins) {
                    if (spec_smem) atomicAdd(&s_spec[(int)f], 1u);
                    else atomicAdd(&spectrum[(int)f], 1ull);
                }
                f = (b.E - emin) / (emax - emin) * nbins;
                if (two && f >= 0.f && f < (float)nbins) {
                    if (spec_smem) atomicAdd(&s_spec[(int)f], 1u);
                    else atomicAdd(&spectrum[(int)f], 1ull);
                }
            }
        }
    }
    if (spec_smem) {
        __syncthreads();
        for (int b = threadIdx.x; b < nbins; b += blockDim.x)
            if (s_spec[b]) atomicAdd(&spectrum[b], (unsigned long long)s_spec[b]);
    }
    c2 = warp_sum(c2);
    if ((threadIdx.x & 31) == 0 && c2) atomicAdd(&counters[2], c2);
}

// ------------------------------------------------------------------------------------------- stage 6: coincidence sorter (extension)
// Windows are opened by the first single that is not inside an earlier window and last cwin us; a thread owns the
// run of windows starting at a single whose predecessor is at least cwin earlier (guaranteed opener).
__device__ __forceinline__ bool pair_ok(const EventRec* __restrict__ s, unsigned a, unsigned b, const DigitizerDev& p) {
    if (p.cmindiff <= 0) return true;
    int d = abs(s[a].pann - s[b].pann);
    if (p.npanels > 0) d = min(d, p.npanels - d);
    return d >= p.cmindiff;
}

__global__ void __launch_bounds__(kThreads) k_coinc_count(const EventRec* __restrict__ s, DigitizerDev p,
                                                          const unsigned* __restrict__ counters, unsigned singles_cap,
                                                          unsigned* __restrict__ cnt) {
    const unsigned n = min(counters[3], singles_cap);
    const double W = (double)p.cwin;
    for (unsigned a0 = blockIdx.x * blockDim.x + threadIdx.x; a0 < n; a0 += gridDim.x * blockDim.x) {
        if (a0 > 0 && !(s[a0].t >= s[a0 - 1].t + W)) continue;
        unsigned a = a0;
        while (true) {
            const double tend = s[a].t + W;
            unsigned m = 0, valid = 0;
            while (a + 1 + m < n && s[a + 1 + m].t < tend) {
                if (pair_ok(s, a, a + 1 + m, p)) valid++;
                cnt[a + 1 + m] = 0;
                m++;
            }
            unsigned c;
            if (p.cpolicy == 0) c = (m == 1 && valid == 1) ? 1u : 0u;
            else c = valid;
            cnt[a] = c;
            a += m + 1;
            if (a >= n) break;
            if (s[a].t >= s[a - 1].t + W) break;  // next guaranteed opener: owned by another thread
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_coinc_emit(const EventRec* __restrict__ s, DigitizerDev p,
                                                         const unsigned* __restrict__ cnt, unsigned* __restrict__ counters,
                                                         unsigned singles_cap, unsigned* __restrict__ status,
                                                         gpet_coincidence* __restrict__ out, unsigned cap) {
    __shared__ unsigned s_tile;
    const unsigned n = min(counters[3], singles_cap);
    const unsigned ntiles = (n + kScanTile - 1) / kScanTile;
    const double W = (double)p.cwin;
    while (true) {
        if (threadIdx.x == 0) s_tile = atomicAdd(&counters[7], 1u);
        __syncthreads();
        const unsigned tile = s_tile;
        if (tile >= ntiles) break;
        const unsigned a0 = tile * kScanTile + threadIdx.x * 8;
        unsigned c[8];
#pragma unroll
        for (int k = 0; k < 8; k++) c[k] = (a0 + k < n) ? cnt[a0 + k] : 0u;
        TileScan sc = tile_exclusive_scan(c, tile, status);
        if (tile == ntiles - 1 && threadIdx.x == 0) counters[4] = sc.tile_excl + sc.tile_total;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (c[k] == 0) continue;
            const unsigned a = a0 + k;
            unsigned o = sc.excl[k];
            const double tend = s[a].t + W;
            for (unsigned b = a + 1; b < n && s[b].t < tend; b++) {
                if (!pair_ok(s, a, b, p)) continue;
                if (o < cap) {  // 2 x 48-byte records copied as 6 x 16 B from the singles list
                    const int4* pa = reinterpret_cast<const int4*>(s + a);
                    const int4* pb = reinterpret_cast<const int4*>(s + b);
                    const int4 a0 = __ldg(pa), a1 = __ldg(pa + 1), a2 = __ldg(pa + 2);
                    const int4 b0 = __ldg(pb), b1 = __ldg(pb + 1), b2 = __ldg(pb + 2);
                    int4* po = reinterpret_cast<int4*>(out + o);
                    po[0] = a0; po[1] = a1; po[2] = a2; po[3] = b0; po[4] = b1; po[5] = b2;
                }
                o++;
            }
        }
    }
}

}  // namespace

// ================================================================================================ launchers
static inline int grid_for(int num_sms) { return num_sms * 2; }

size_t sort_state_bytes() { return sizeof(rsort::SortState); }
size_t sort_lookback_words(size_t capacity) { return ((capacity + rsort::kTile - 1) / rsort::kTile) * (size_t)rsort::kBins; }
unsigned scan_tiles(size_t capacity) { return (unsigned)((capacity + kScanTile - 1) / kScanTile); }
unsigned bucket_words() { return kMaxBuckets; }

int launch_digitize(EventBuf ev, void* singles_aos, unsigned int singles_cap, void* coinc_aos, unsigned int coinc_cap,
                    const DigitizerDev& p, DigitizerWorkspace& ws, uint64_t seed, int num_sms, cudaStream_t s) {
    const int grid = grid_for(num_sms);
    int launches = 0;
    EventRec* singles = static_cast<EventRec*>(singles_aos);
    unsigned long long* bkeys = ws.tkeys[1];   // the LSD ping-pong buffers double as the scatter target: the two sorts
    unsigned* bidx = ws.tvals[1];              // never run in the same frame
    const unsigned* lsd = &ws.counters[kFlagLsd];
    GPET_LAUNCH("k_begin", s, k_begin<<<32, kThreads, 0, s>>>(ws.counters, ws.st_time, ws.st_site, ws.scan_status[0], ws.scan_status[1],
                                                             ws.scan_status[2], ws.max_tiles, ws.bcount, ws.minmax));
    GPET_LAUNCH("k_prep", s, k_prep<<<grid, kThreads, 0, s>>>(ev, p, seed, ws.tkeys[0], ws.counters, ws.minmax));
    // time sort of the alive records (dead ones are left out: they would sink to the tail, like MAXT does)
    GPET_LAUNCH("k_bucket_count", s, k_bucket_count<<<grid, kThreads, 0, s>>>(ws.tkeys[0], ws.counters, ws.minmax, ws.bcount));
    GPET_LAUNCH("k_bucket_scan", s, k_bucket_scan<<<kMaxBuckets / kScanTile, kThreads, 0, s>>>(ws.bcount, ws.bstart, ws.bcur, ws.scan_status[2],
                                                                                           ws.counters, ws.minmax));
    GPET_LAUNCH("k_bucket_scatter", s, k_bucket_scatter<<<grid, kThreads, 0, s>>>(ws.tkeys[0], ws.counters, ws.minmax, ws.bcur, bkeys, bidx));
    GPET_LAUNCH("k_bucket_sort", s, k_bucket_sort<<<2 * grid, kThreads, 0, s>>>(bkeys, bidx, ws.bstart, ws.bcur, ws.counters, ws.minmax, ws.order_t,
                                                                             ws.tkeys[0]));
    GPET_LAUNCH("k_lsd_hist", s, k_lsd_hist<<<grid, kThreads, 0, s>>>(ws.tkeys[0], ws.tvals[0], ws.counters, ws.st_time, ws.lookback[0]));
    launches += 7;
    launches += radix_sort_passes<unsigned long long>(ws.tkeys, ws.tvals, &ws.counters[0], ws.st_time, ws.lookback, 8, grid, s, lsd);
    // site keys + site sort (stable => (site, t) order == orderevents, detector.cu:369-385)
    GPET_LAUNCH("k_site_keys", s, k_site_keys<<<grid, kThreads, 0, s>>>(ev, p, ws.tkeys[0], ws.tkeys[1], ws.tvals[0], ws.tvals[1], ws.st_time,
                                                                      ws.order_t, ws.skeys[0], ws.svals[0], ws.counters, ws.st_site, ws.lookback[0]));
    launches += 1;
    launches += radix_sort_passes<unsigned>(ws.skeys, ws.svals, &ws.counters[1], ws.st_site, ws.lookback, 4, grid, s);
    GPET_LAUNCH("k_deadtime", s, k_deadtime<<<grid, kThreads, 0, s>>>(p, ws.tkeys[0], ws.skeys[0], ws.skeys[1], ws.svals[0], ws.svals[1], ws.st_site,
                                                                    ws.kill, ws.counters));
    GPET_LAUNCH("k_emit_singles", s, k_emit_singles<<<grid, kThreads, 0, s>>>(ev, p, singles, singles_cap, ws.order_t, ws.kill, ws.counters,
                                                                            ws.scan_status[0], ws.spectrum, ws.spectrum_bins, ws.spec_emin,
                                                                            ws.spec_emax));
    launches += 2;
    if (p.cwin > 0.f && coinc_aos) {
        GPET_LAUNCH("k_coinc_count", s, k_coinc_count<<<grid, kThreads, 0, s>>>(singles, p, ws.counters, singles_cap, ws.coinc_cnt));
        GPET_LAUNCH("k_coinc_emit", s, k_coinc_emit<<<grid, kThreads, 0, s>>>(singles, p, ws.coinc_cnt, ws.counters, singles_cap, ws.scan_status[1],
                                                                            static_cast<gpet_coincidence*>(coinc_aos), coinc_cap));
        launches += 2;
    }
    return launches;
}

}  // namespace gpet
