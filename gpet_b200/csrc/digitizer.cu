// Digitizer chain on the device: blur -> thresholder -> time sort -> site order -> dead time -> energy window
// -> singles (time sorted) [-> coincidence sorter].  Reference: blur/energywindow/setSitenum/deadtime kernels
// (gPET_kernals.cu:607-698, 814-837) and the host orchestration with three CPU sorts (gPET.cu:385-424,
// detector.cu:354-385).  Here nothing leaves the device between stages: counts stay in `counters`, the sorts are
// one-kernel-per-digit radix sorts over the order-preserving u64 image of the fp64 time (radix_sort.cuh), and the
// final singles list is produced by one fused flag + scan + compaction of the time order (no re-sort after dead time
// / energy window, since killing keeps the order).  The launch sequence is static (all sizes live on the device), so
// the whole chain replays from a CUDA graph.
//
// Launches per frame: k_begin, k_prep, 8 x k_onesweep<u64>, k_site_keys, 4 x k_onesweep<u32> (passes whose digit is
// constant return at once), k_deadtime, k_emit_singles [, k_coinc_count, k_coinc_emit].
#include <cstdlib>
#include "kernels.hpp"
#include "philox.cuh"
#include "ktimer.hpp"
#include "radix_sort.cuh"

#include "../../include/gpet_b200.h"

namespace gpet {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ unsigned long long time_key(double t) {
    unsigned long long b = (unsigned long long)__double_as_longlong(t);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ unsigned warp_sum(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------- AoS <-> SoA
__global__ void k_aos_to_soa(const gpet_event* __restrict__ aos, EventSoA ev, unsigned n) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *ev.count = n;
    for (; i < n; i += gridDim.x * blockDim.x) {
        const int4* p = reinterpret_cast<const int4*>(aos + i);  // 48 B = 3 x 16 B
        int4 a = p[0], b = p[1], c = p[2];
        ev.parn[i] = a.x; ev.pann[i] = a.y; ev.modn[i] = a.z; ev.cryn[i] = a.w;
        ev.siten[i] = b.x; ev.eventid[i] = b.y;
        ev.t[i] = __longlong_as_double(((long long)(unsigned)b.w << 32) | (unsigned)b.z);
        ev.E[i] = __int_as_float(c.x); ev.x[i] = __int_as_float(c.y);
        ev.y[i] = __int_as_float(c.z); ev.z[i] = __int_as_float(c.w);
    }
}

__device__ __forceinline__ void store_event_aos(gpet_event* dst, const EventSoA& ev, unsigned i) {
    long long tb = __double_as_longlong(ev.t[i]);
    int4 a = make_int4(ev.parn[i], ev.pann[i], ev.modn[i], ev.cryn[i]);
    int4 b = make_int4(ev.siten[i], ev.eventid[i], (int)(unsigned)(tb & 0xffffffffll), (int)(unsigned)((unsigned long long)tb >> 32));
    int4 c = make_int4(__float_as_int(ev.E[i]), __float_as_int(ev.x[i]), __float_as_int(ev.y[i]), __float_as_int(ev.z[i]));
    int4* p = reinterpret_cast<int4*>(dst);
    p[0] = a; p[1] = b; p[2] = c;
}

// One event held in registers: all eleven columns are loaded before anything is stored, so the loads are independent
// (the SoA columns may alias as far as the compiler can tell, which would otherwise serialise load -> store -> load).
struct EventRec {
    int parn, pann, modn, cryn, siten, eventid;
    double t;
    float E, x, y, z;
};

__device__ __forceinline__ EventRec load_event(const EventSoA& ev, unsigned i) {
    EventRec r;
    r.parn = ev.parn[i]; r.pann = ev.pann[i]; r.modn = ev.modn[i]; r.cryn = ev.cryn[i];
    r.siten = ev.siten[i]; r.eventid = ev.eventid[i];
    r.t = ev.t[i];
    r.E = ev.E[i]; r.x = ev.x[i]; r.y = ev.y[i]; r.z = ev.z[i];
    return r;
}

__device__ __forceinline__ void store_event(const EventSoA& ev, unsigned o, const EventRec& r) {
    ev.parn[o] = r.parn; ev.pann[o] = r.pann; ev.modn[o] = r.modn; ev.cryn[o] = r.cryn;
    ev.siten[o] = r.siten; ev.eventid[o] = r.eventid;
    ev.t[o] = r.t;
    ev.E[o] = r.E; ev.x[o] = r.x; ev.y[o] = r.y; ev.z[o] = r.z;
}

__device__ __forceinline__ void store_event_aos(gpet_event* dst, const EventRec& r) {
    long long tb = __double_as_longlong(r.t);
    int4* p = reinterpret_cast<int4*>(dst);
    p[0] = make_int4(r.parn, r.pann, r.modn, r.cryn);
    p[1] = make_int4(r.siten, r.eventid, (int)(unsigned)(tb & 0xffffffffll), (int)(unsigned)((unsigned long long)tb >> 32));
    p[2] = make_int4(__float_as_int(r.E), __float_as_int(r.x), __float_as_int(r.y), __float_as_int(r.z));
}

__global__ void k_soa_to_aos(EventSoA ev, gpet_event* __restrict__ aos) {
    const unsigned n = min(*ev.count, ev.capacity);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        store_event_aos(aos + i, ev, i);
}

// ------------------------------------------------------------------------------------------- stage 0: reset
// counters[0..7], both sort states (histograms, tile counters, buffer selectors) and the scan status words.
__global__ void __launch_bounds__(kThreads) k_begin(unsigned* __restrict__ counters, rsort::SortState* st_time,
                                                    rsort::SortState* st_site, unsigned* __restrict__ scan_status0,
                                                    unsigned* __restrict__ scan_status1, unsigned max_tiles) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (tid < 8) counters[tid] = 0;
    constexpr unsigned kWords = sizeof(rsort::SortState) / 4;
    unsigned* a = reinterpret_cast<unsigned*>(st_time);
    unsigned* b = reinterpret_cast<unsigned*>(st_site);
    for (unsigned i = tid; i < kWords; i += nth) { a[i] = 0; b[i] = 0; }
    for (unsigned i = tid; i < max_tiles; i += nth) { scan_status0[i] = 0; scan_status1[i] = 0; }
}

// ------------------------------------------------------------------------------------------- stage 1: blur + thresholder + time keys
// blur (gPET_kernals.cu:814-837) + energywindow(Eth, 2e6) (gPET.cu:393) fused; writes the sort keys into buffer 0 of
// the time sort, accumulates the digit histograms of all 8 passes and clears the first look-back array.
__global__ void __launch_bounds__(kThreads) k_prep(EventSoA ev, DigitizerDev p, uint64_t seed,
                                                   unsigned long long* __restrict__ keys, unsigned* __restrict__ vals,
                                                   unsigned* __restrict__ counters, rsort::SortState* st_time,
                                                   unsigned* __restrict__ lookback0) {
    __shared__ unsigned sh_hist[8 * rsort::kBins];
    const unsigned n = min(*ev.count, ev.capacity);
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[0] = n;
    for (int i = threadIdx.x; i < 8 * rsort::kBins; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    unsigned alive_cnt = 0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float E = ev.E[i];
        double t = ev.t[i];
        float R = 0.f;
        // float / double mix exactly as the reference expression is typed
        if (p.blur_policy == 0) R = __fmul_rn(__fsqrt_rn(__fdiv_rn(p.Eref, E)), p.Rref);
        if (p.blur_policy == 1)
            R = (float)__dadd_rn((double)p.Rref, __ddiv_rn((double)__fmul_rn(p.slope, __fsub_rn(E, p.Eref)), 1e6));
        if (!(R > 0.f)) R = 0.f;
        // R == 0 leaves E bit-identical (E + 0), so the draw is skipped: this is the deterministic replay mode
        if (R > 0.f || p.sblur > 0.f || p.tblur > 0.f) {
            Philox rng(seed, (uint64_t)(uint32_t)ev.parn[i], ((uint32_t)kStageBlur << 24) | ((uint32_t)ev.siten[i] & 0xFFFFFFu));
            uint4 r = rng.next();
            float rad = sqrtf(-2.0f * logf(u01(r.x)));
            float g0 = rad * cosf(kTwoPi * u01(r.y));
            float nre = __fmul_rn(__fmul_rn(g0, R), E);
            E = (float)__dadd_rn((double)E, __ddiv_rn((double)nre, 2.35482));
            ev.E[i] = E;
            if (p.sblur > 0.f) {
                uint4 q = rng.next();
                float ra = sqrtf(-2.0f * logf(u01(q.x))), rb = sqrtf(-2.0f * logf(u01(q.z)));
                float a0 = kTwoPi * u01(q.y), a1 = kTwoPi * u01(q.w);
                ev.x[i] = __fadd_rn(ev.x[i], __fmul_rn(p.sblur, ra * cosf(a0)));
                ev.y[i] = __fadd_rn(ev.y[i], __fmul_rn(p.sblur, ra * sinf(a0)));
                ev.z[i] = __fadd_rn(ev.z[i], __fmul_rn(p.sblur, rb * cosf(a1)));
            }
            if (p.tblur > 0.f) {
                float g1 = rad * sinf(kTwoPi * u01(r.y));
                double tb = t + (double)p.tblur * (double)g1;
                t = tb > 0.0 ? tb : t;
                ev.t[i] = t;
            }
        }
        // energywindow: dead iff E < lo || E > hi (gPET_kernals.cu:648); an already dead record (t >= MAXT) stays dead
        bool alive = !(E < p.Eth || E > 2000000.0f) && t < kMaxT * 0.1;
        const unsigned long long key = alive ? time_key(t) : ~0ull;
        keys[i] = key;
        vals[i] = i;
        rsort::hist_add<unsigned long long, 8>(sh_hist, key);
        alive_cnt += alive ? 1u : 0u;
    }
    alive_cnt = warp_sum(alive_cnt);
    if ((threadIdx.x & 31) == 0 && alive_cnt) atomicAdd(&counters[1], alive_cnt);
    __syncthreads();
    rsort::hist_flush<8>(sh_hist, st_time);
    rsort::clear_lookback(lookback0, n);
}

// ------------------------------------------------------------------------------------------- stage 2: site keys
// setSitenum (gPET_kernals.cu:607-640) fused with building the (site) sort keys over the time order.
__global__ void __launch_bounds__(kThreads) k_site_keys(EventSoA ev, DigitizerDev p, const unsigned* __restrict__ tvals0,
                                                        const unsigned* __restrict__ tvals1, const rsort::SortState* st_time,
                                                        unsigned* __restrict__ order_t, unsigned* __restrict__ keys,
                                                        unsigned* __restrict__ vals, const unsigned* __restrict__ counters,
                                                        rsort::SortState* st_site, unsigned* __restrict__ lookback0) {
    __shared__ unsigned sh_hist[4 * rsort::kBins];
    const unsigned n1 = counters[1];
    const unsigned* __restrict__ t_sorted_vals = rsort::current_buffer(st_time, 8, counters[0]) ? tvals1 : tvals0;
    for (int i = threadIdx.x; i < 4 * rsort::kBins; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < n1; j += gridDim.x * blockDim.x) {
        unsigned i = t_sorted_vals[j];
        order_t[j] = i;
        int site;
        switch (p.dlevel) {
            case 0: site = 0; break;
            case 1: site = ev.pann[i]; break;
            case 2: site = ev.pann[i] * p.moduleN + ev.modn[i]; break;
            default: site = ev.siten[i]; break;  // dlevel == 3: keep what readout left (gPET.cu:402-407)
        }
        if (p.dlevel >= 0 && p.dlevel <= 2) ev.siten[i] = site;
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

// ------------------------------------------------------------------------------------------- stage 3: dead time
// deadtime (gPET_kernals.cu:657-698) with the snapshot-start semantics of SURVEY 8(a) D7: every decision uses the
// original times; `tdead` is fp32 and `tdead + interval` is an fp32 sum, as in the reference.  q runs over the
// (site, t) order; kill flags are stored by position in the time order.
__global__ void __launch_bounds__(kThreads) k_deadtime(EventSoA ev, DigitizerDev p, const unsigned* __restrict__ order_t,
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
        const double t = ev.t[order_t[j]];
        bool same_prev = false;
        double tprev = 0.0;
        if (q > 0 && site_keys[q] == site_keys[q - 1]) {
            same_prev = true;
            tprev = ev.t[order_t[order_s[q - 1]]];
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
                const double tr = ev.t[order_t[jr]];
                const double tr_prev = ev.t[order_t[order_s[r - 1]]];
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

// ------------------------------------------------------------------------------------------- stage 4: energy window + compaction -> singles
// energywindow(Ewinmin, Ewinmax) (gPET.cu:418) over the survivors of the dead time, in time order.
__global__ void __launch_bounds__(kThreads) k_emit_singles(EventSoA ev, DigitizerDev p, EventSoA singles,
                                                           gpet_event* __restrict__ singles_aos,
                                                           const unsigned* __restrict__ order_t,
                                                           const unsigned char* __restrict__ kill, unsigned* __restrict__ counters,
                                                           unsigned* __restrict__ status, unsigned long long* __restrict__ spectrum,
                                                           int nbins, float emin, float emax) {
    __shared__ unsigned s_tile;
    __shared__ unsigned s_idx[kScanTile];
    __shared__ unsigned s_spec[kSpecSmemBins];   // block-private energy histogram (flushed once per block)
    const unsigned n1 = counters[1];
    const unsigned ntiles = (n1 + kScanTile - 1) / kScanTile;
    if (ntiles == 0 && blockIdx.x == 0 && threadIdx.x == 0) *singles.count = 0;
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
                const float E = ev.E[i];
                flag[k] = (a2 && !(E < p.Ewinmin || E > p.Ewinmax)) ? 1u : 0u;
                c2 += a2 ? 1u : 0u;
            }
        }
        TileScan sc = tile_exclusive_scan(flag, tile, status);
        if (tile == ntiles - 1 && threadIdx.x == 0) {
            counters[3] = sc.tile_excl + sc.tile_total;
            *singles.count = sc.tile_excl + sc.tile_total;
        }
        // survivors of this tile, in order, through shared memory: the emission below is then one single per thread
        // and iteration, with coalesced column stores
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (flag[k]) s_idx[sc.excl[k] - sc.tile_excl] = idx[k];
        __syncthreads();
        for (unsigned r = threadIdx.x; r < sc.tile_total; r += 2 * kThreads) {
            const unsigned r2 = r + kThreads;
            const bool two = r2 < sc.tile_total;
            const unsigned o = sc.tile_excl + r, o2 = sc.tile_excl + r2;
            const EventRec a = load_event(ev, s_idx[r]);
            EventRec b = a;
            if (two) b = load_event(ev, s_idx[r2]);
            if (o < singles.capacity) {
                store_event(singles, o, a);
                if (singles_aos) store_event_aos(singles_aos + o, a);
            }
            if (two && o2 < singles.capacity) {
                store_event(singles, o2, b);
                if (singles_aos) store_event_aos(singles_aos + o2, b);
            }
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

// ------------------------------------------------------------------------------------------- stage 5: coincidence sorter (extension)
// Windows are opened by the first single that is not inside an earlier window and last cwin us; a thread owns the
// run of windows starting at a single whose predecessor is at least cwin earlier (guaranteed opener).
__device__ __forceinline__ bool pair_ok(const EventSoA& s, unsigned a, unsigned b, const DigitizerDev& p) {
    if (p.cmindiff <= 0) return true;
    int d = abs(s.pann[a] - s.pann[b]);
    if (p.npanels > 0) d = min(d, p.npanels - d);
    return d >= p.cmindiff;
}

__global__ void __launch_bounds__(kThreads) k_coinc_count(EventSoA s, DigitizerDev p, unsigned* __restrict__ cnt) {
    const unsigned n = min(*s.count, s.capacity);
    const double W = (double)p.cwin;
    for (unsigned a0 = blockIdx.x * blockDim.x + threadIdx.x; a0 < n; a0 += gridDim.x * blockDim.x) {
        if (a0 > 0 && !(s.t[a0] >= s.t[a0 - 1] + W)) continue;
        unsigned a = a0;
        while (true) {
            const double tend = s.t[a] + W;
            unsigned m = 0, valid = 0;
            while (a + 1 + m < n && s.t[a + 1 + m] < tend) {
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
            if (s.t[a] >= s.t[a - 1] + W) break;  // next guaranteed opener: owned by another thread
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_coinc_emit(EventSoA s, DigitizerDev p, const unsigned* __restrict__ cnt,
                                                         unsigned* __restrict__ counters, unsigned* __restrict__ status,
                                                         const gpet_event* __restrict__ singles_aos,
                                                         gpet_coincidence* __restrict__ out, unsigned cap) {
    __shared__ unsigned s_tile;
    const unsigned n = min(*s.count, s.capacity);
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
            const double tend = s.t[a] + W;
            for (unsigned b = a + 1; b < n && s.t[b] < tend; b++) {
                if (!pair_ok(s, a, b, p)) continue;
                if (o < cap) {  // 2 x 48-byte records copied as 6 x 16 B from the AoS singles list
                    const int4* pa = reinterpret_cast<const int4*>(singles_aos + a);
                    const int4* pb = reinterpret_cast<const int4*>(singles_aos + b);
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

int launch_events_aos_to_soa(const void* aos, EventSoA ev, unsigned int n, cudaStream_t s) {
    unsigned blocks = n ? (n + kThreads - 1) / kThreads : 1;
    if (blocks > 4096) blocks = 4096;
    GPET_LAUNCH("k_aos_to_soa", s, k_aos_to_soa<<<blocks, kThreads, 0, s>>>(static_cast<const gpet_event*>(aos), ev, n));
    return 1;
}

int launch_events_soa_to_aos(EventSoA ev, void* aos, cudaStream_t s) {
    GPET_LAUNCH("k_soa_to_aos", s, k_soa_to_aos<<<1024, kThreads, 0, s>>>(ev, static_cast<gpet_event*>(aos)));
    return 1;
}

int launch_digitize(EventSoA ev, EventSoA singles, void* singles_aos, void* coinc_aos, unsigned int coinc_cap,
                    const DigitizerDev& p, DigitizerWorkspace& ws, uint64_t seed, int num_sms, cudaStream_t s) {
    const int grid = grid_for(num_sms);
    int launches = 0;
    GPET_LAUNCH("k_begin", s, k_begin<<<8, kThreads, 0, s>>>(ws.counters, ws.st_time, ws.st_site, ws.scan_status[0], ws.scan_status[1], ws.max_tiles));
    GPET_LAUNCH("k_prep", s, k_prep<<<grid, kThreads, 0, s>>>(ev, p, seed, ws.tkeys[0], ws.tvals[0], ws.counters, ws.st_time, ws.lookback[0]));
    launches += 2;
    // time sort over all n_in records (dead ones carry the maximal key and sink to the tail, like MAXT does)
    launches += radix_sort_passes<unsigned long long>(ws.tkeys, ws.tvals, &ws.counters[0], ws.st_time, ws.lookback, 8, grid, s);
    // site keys + site sort (stable => (site, t) order == orderevents, detector.cu:369-385)
    GPET_LAUNCH("k_site_keys", s, k_site_keys<<<grid, kThreads, 0, s>>>(ev, p, ws.tvals[0], ws.tvals[1], ws.st_time, ws.order_t, ws.skeys[0], ws.svals[0],
                                          ws.counters, ws.st_site, ws.lookback[0]));
    launches += 1;
    launches += radix_sort_passes<unsigned>(ws.skeys, ws.svals, &ws.counters[1], ws.st_site, ws.lookback, 4, grid, s);
    GPET_LAUNCH("k_deadtime", s, k_deadtime<<<grid, kThreads, 0, s>>>(ev, p, ws.order_t, ws.skeys[0], ws.skeys[1], ws.svals[0], ws.svals[1], ws.st_site,
                                         ws.kill, ws.counters));
    GPET_LAUNCH("k_emit_singles", s, k_emit_singles<<<grid, kThreads, 0, s>>>(ev, p, singles, static_cast<gpet_event*>(singles_aos), ws.order_t, ws.kill,
                                             ws.counters, ws.scan_status[0], ws.spectrum, ws.spectrum_bins, ws.spec_emin,
                                             ws.spec_emax));
    launches += 2;
    if (p.cwin > 0.f && coinc_aos) {
        GPET_LAUNCH("k_coinc_count", s, k_coinc_count<<<grid, kThreads, 0, s>>>(singles, p, ws.coinc_cnt));
        GPET_LAUNCH("k_coinc_emit", s, k_coinc_emit<<<grid, kThreads, 0, s>>>(singles, p, ws.coinc_cnt, ws.counters, ws.scan_status[1],
                                               static_cast<const gpet_event*>(singles_aos),
                                               static_cast<gpet_coincidence*>(coinc_aos), coinc_cap));
        launches += 2;
    }
    return launches;
}

}  // namespace gpet
