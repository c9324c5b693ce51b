// Digitizer chain on the device: blur -> thresholder -> time sort -> site order -> dead time -> energy window
// -> singles (time sorted) [-> coincidence sorter].  Reference: blur/energywindow/setSitenum/deadtime kernels
// (gPET_kernals.cu:607-698, 814-837) and the host orchestration with three CPU sorts (gPET.cu:385-424,
// detector.cu:354-385).  Here nothing leaves the device between stages: counts stay in `counters`, the sorts are
// radix sorts over the order-preserving u64 image of the fp64 time, and the final singles list is produced by one
// compaction of the time order (no re-sort after dead time / energy window, since killing keeps the order).
#include "kernels.hpp"
#include "philox.cuh"
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

__global__ void k_soa_to_aos(EventSoA ev, gpet_event* __restrict__ aos) {
    const unsigned n = min(*ev.count, ev.capacity);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        store_event_aos(aos + i, ev, i);
}

// ------------------------------------------------------------------------------------------- stage 1: blur + thresholder + time keys
// blur (gPET_kernals.cu:814-837) + energywindow(Eth, 2e6) (gPET.cu:393) fused; writes the sort keys.
__global__ void __launch_bounds__(kThreads) k_prep(EventSoA ev, DigitizerDev p, uint64_t seed,
                                                   unsigned long long* __restrict__ keys, unsigned* __restrict__ vals,
                                                   unsigned* __restrict__ counters) {
    const unsigned n = min(*ev.count, ev.capacity);
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[0] = n;
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
        keys[i] = alive ? time_key(t) : ~0ull;
        vals[i] = i;
        alive_cnt += alive ? 1u : 0u;
    }
    alive_cnt = warp_sum(alive_cnt);
    if ((threadIdx.x & 31) == 0 && alive_cnt) atomicAdd(&counters[1], alive_cnt);
}

// ------------------------------------------------------------------------------------------- stage 2: site keys
// setSitenum (gPET_kernals.cu:607-640) fused with building the (site) sort keys over the time order.
__global__ void __launch_bounds__(kThreads) k_site_keys(EventSoA ev, DigitizerDev p, const unsigned* __restrict__ t_sorted_vals,
                                                        unsigned* __restrict__ order_t, unsigned* __restrict__ keys,
                                                        unsigned* __restrict__ vals, const unsigned* __restrict__ counters) {
    const unsigned n1 = counters[1];
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
        keys[j] = (unsigned)site ^ 0x80000000u;
        vals[j] = j;
    }
}

// ------------------------------------------------------------------------------------------- stage 3: dead time
// deadtime (gPET_kernals.cu:657-698) with the snapshot-start semantics of SURVEY 8(a) D7: every decision uses the
// original times; `tdead` is fp32 and `tdead + interval` is an fp32 sum, as in the reference.
__global__ void __launch_bounds__(kThreads) k_deadtime(EventSoA ev, DigitizerDev p, const unsigned* __restrict__ order_t,
                                                       const unsigned* __restrict__ order_s, const unsigned* __restrict__ site_keys,
                                                       unsigned char* __restrict__ kill, unsigned* __restrict__ counters) {
    const unsigned n1 = counters[1];
    const float tau = p.dtime;
    unsigned killed = 0;
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n1; q += gridDim.x * blockDim.x) {
        const unsigned i = order_t[order_s[q]];
        const double t = ev.t[i];
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
            kill[i] = killable ? 1 : 0;
            killed += killable ? 1u : 0u;
        } else {
            // non-paralyzable: sequential anchor chain per site.  An event its predecessor cannot kill survives any
            // earlier anchor as well (fp32 rounding and the fp32 sum are monotone), so it is a guaranteed anchor and
            // the chain can be cut there: one thread per such run start, runs are short at realistic rates.
            if (killable) continue;
            kill[i] = 0;
            float tdead = (float)t;
            unsigned r = q + 1;
            while (r < n1 && site_keys[r] == site_keys[q]) {
                const unsigned ir = order_t[order_s[r]];
                const double tr = ev.t[ir];
                const double tr_prev = ev.t[order_t[order_s[r - 1]]];
                if (!(tr < (double)__fadd_rn((float)tr_prev, tau))) break;  // next run start
                if (tr < (double)__fadd_rn(tdead, tau)) {
                    kill[ir] = 1;
                    killed++;
                } else {
                    kill[ir] = 0;
                    tdead = (float)tr;
                }
                r++;
            }
        }
    }
    killed = warp_sum(killed);
    if ((threadIdx.x & 31) == 0 && killed) atomicAdd(&counters[5], killed);
}

// ------------------------------------------------------------------------------------------- stage 4: final flags over the time order
__global__ void __launch_bounds__(kThreads) k_final_flags(EventSoA ev, DigitizerDev p, const unsigned* __restrict__ order_t,
                                                          const unsigned char* __restrict__ kill, unsigned* __restrict__ flags,
                                                          unsigned* __restrict__ counters) {
    const unsigned n1 = counters[1];
    unsigned c2 = 0, c3 = 0;
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < n1; j += gridDim.x * blockDim.x) {
        unsigned i = order_t[j];
        bool a2 = kill[i] == 0;
        float E = ev.E[i];
        bool a3 = a2 && !(E < p.Ewinmin || E > p.Ewinmax);  // energywindow(Ewinmin, Ewinmax) (gPET.cu:418)
        flags[j] = a3 ? 1u : 0u;
        c2 += a2 ? 1u : 0u;
        c3 += a3 ? 1u : 0u;
    }
    c2 = warp_sum(c2);
    c3 = warp_sum(c3);
    if ((threadIdx.x & 31) == 0) {
        if (c2) atomicAdd(&counters[2], c2);
        if (c3) atomicAdd(&counters[3], c3);
    }
}

// ------------------------------------------------------------------------------------------- exclusive scan (u32), 3 kernels
constexpr int kScanTile = 2048;  // 256 threads x 8

__global__ void __launch_bounds__(kThreads) k_scan_reduce(const unsigned* __restrict__ in, const unsigned* __restrict__ n_ptr,
                                                          unsigned* __restrict__ block_sums) {
    __shared__ unsigned ws[8];
    const unsigned n = *n_ptr;
    const unsigned ntiles = (n + kScanTile - 1) / kScanTile;
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        unsigned s = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            unsigned i = tile * kScanTile + k * kThreads + threadIdx.x;
            if (i < n) s += in[i];
        }
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned t = 0;
            for (int w = 0; w < 8; w++) t += ws[w];
            block_sums[tile] = t;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) k_scan_sums(unsigned* __restrict__ block_sums, const unsigned* __restrict__ n_ptr,
                                                    unsigned* __restrict__ total) {
    __shared__ unsigned warp_sums[32];
    __shared__ unsigned carry;
    const unsigned n = *n_ptr;
    const unsigned m = (n + kScanTile - 1) / kScanTile;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned base = 0; base < m; base += 1024) {
        unsigned i = base + threadIdx.x;
        unsigned v = i < m ? block_sums[i] : 0u;
        unsigned x = v;
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
            warp_sums[lane] = w;
        }
        __syncthreads();
        unsigned excl = carry + (warp ? warp_sums[warp - 1] : 0u) + (x - v);
        if (i < m) block_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry += warp_sums[31];
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

// out[i] = exclusive prefix of in[] (tile-local scan + scanned block sums)
__global__ void __launch_bounds__(kThreads) k_scan_apply(const unsigned* __restrict__ in, const unsigned* __restrict__ n_ptr,
                                                         const unsigned* __restrict__ block_sums, unsigned* __restrict__ out) {
    __shared__ unsigned ws[8];
    const unsigned n = *n_ptr;
    const unsigned ntiles = (n + kScanTile - 1) / kScanTile;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // blocked arrangement: thread owns 8 consecutive elements
        unsigned i0 = tile * kScanTile + threadIdx.x * 8;
        unsigned v[8], s = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { v[k] = (i0 + k < n) ? in[i0 + k] : 0u; s += v[k]; }
        unsigned x = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= (unsigned)o) x += y;
        }
        if (lane == 31) ws[warp] = x;
        __syncthreads();
        unsigned wprefix = 0;
        for (unsigned w = 0; w < warp; w++) wprefix += ws[w];
        unsigned excl = block_sums[tile] + wprefix + (x - s);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (i0 + k < n) out[i0 + k] = excl;
            excl += v[k];
        }
        __syncthreads();
    }
}

int exclusive_scan(const unsigned* in, unsigned* out, const unsigned* n_dev, unsigned* block_sums, unsigned* total,
                   int grid, cudaStream_t s) {
    k_scan_reduce<<<grid, kThreads, 0, s>>>(in, n_dev, block_sums);
    k_scan_sums<<<1, 1024, 0, s>>>(block_sums, n_dev, total);
    k_scan_apply<<<grid, kThreads, 0, s>>>(in, n_dev, block_sums, out);
    return 3;
}

// ------------------------------------------------------------------------------------------- stage 5: singles out
__global__ void __launch_bounds__(kThreads) k_emit_singles(EventSoA ev, EventSoA singles, gpet_event* __restrict__ singles_aos,
                                                           const unsigned* __restrict__ order_t, const unsigned* __restrict__ flags,
                                                           const unsigned* __restrict__ offs, const unsigned* __restrict__ counters,
                                                           unsigned long long* __restrict__ spectrum, int nbins, float emin, float emax) {
    const unsigned n1 = counters[1];
    if (blockIdx.x == 0 && threadIdx.x == 0) *singles.count = counters[3];
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < n1; j += gridDim.x * blockDim.x) {
        if (!flags[j]) continue;
        unsigned i = order_t[j], o = offs[j];
        if (o >= singles.capacity) continue;
        singles.parn[o] = ev.parn[i]; singles.pann[o] = ev.pann[i]; singles.modn[o] = ev.modn[i];
        singles.cryn[o] = ev.cryn[i]; singles.siten[o] = ev.siten[i]; singles.eventid[o] = ev.eventid[i];
        singles.t[o] = ev.t[i]; singles.E[o] = ev.E[i];
        singles.x[o] = ev.x[i]; singles.y[o] = ev.y[i]; singles.z[o] = ev.z[i];
        if (singles_aos) store_event_aos(singles_aos + o, ev, i);
        if (spectrum && nbins > 0) {
            float f = (ev.E[i] - emin) / (emax - emin) * nbins;
            if (f >= 0.f && f < (float)nbins) atomicAdd(&spectrum[(int)f], 1ull);
        }
    }
}

// ------------------------------------------------------------------------------------------- stage 6: coincidence sorter (extension)
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
                                                         const unsigned* __restrict__ offs, gpet_coincidence* __restrict__ out,
                                                         unsigned cap) {
    const unsigned n = min(*s.count, s.capacity);
    const double W = (double)p.cwin;
    for (unsigned a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x) {
        if (cnt[a] == 0) continue;
        unsigned o = offs[a];
        const double tend = s.t[a] + W;
        for (unsigned b = a + 1; b < n && s.t[b] < tend; b++) {
            if (!pair_ok(s, a, b, p)) continue;
            if (o < cap) {
                store_event_aos(&out[o].a, s, a);
                store_event_aos(&out[o].b, s, b);
            }
            o++;
        }
    }
}

__global__ void k_zero_counters(unsigned* counters) {
    if (threadIdx.x < 8) counters[threadIdx.x] = 0;  // [8..] belong to the detector stage
}

}  // namespace

// ================================================================================================ launchers
static inline int grid_for(int num_sms) { return num_sms * 4; }

int launch_events_aos_to_soa(const void* aos, EventSoA ev, unsigned int n, cudaStream_t s) {
    unsigned blocks = n ? (n + kThreads - 1) / kThreads : 1;
    if (blocks > 4096) blocks = 4096;
    k_aos_to_soa<<<blocks, kThreads, 0, s>>>(static_cast<const gpet_event*>(aos), ev, n);
    return 1;
}

int launch_events_soa_to_aos(EventSoA ev, void* aos, cudaStream_t s) {
    k_soa_to_aos<<<1024, kThreads, 0, s>>>(ev, static_cast<gpet_event*>(aos));
    return 1;
}

int launch_radix_sort_pairs(SortWorkspace& ws, const unsigned int* n_dev, int begin_bit, int end_bit, int* result_buffer,
                            int num_sms, cudaStream_t s) {
    return radix_sort_pairs<unsigned long long>(ws.keys, ws.vals, ws.tile_hist, n_dev, begin_bit, end_bit, result_buffer,
                                                grid_for(num_sms), s);
}

int launch_digitize(EventSoA ev, EventSoA singles, void* singles_aos, void* coinc_aos, unsigned int coinc_cap,
                    const DigitizerDev& p, DigitizerWorkspace& ws, uint64_t seed, int num_sms, cudaStream_t s) {
    const int grid = grid_for(num_sms);
    int launches = 0;
    k_zero_counters<<<1, 32, 0, s>>>(ws.counters);
    k_prep<<<grid, kThreads, 0, s>>>(ev, p, seed, ws.sort.keys[0], ws.sort.vals[0], ws.counters);
    launches += 2;
    // time sort over all n_in records (dead ones carry the maximal key and sink to the tail, like MAXT does)
    int rb = 0;
    launches += radix_sort_pairs<unsigned long long>(ws.sort.keys, ws.sort.vals, ws.sort.tile_hist, &ws.counters[0], 0, 64,
                                                     &rb, grid, s);
    // site keys + site sort (stable => (site, t) order == orderevents, detector.cu:369-385)
    unsigned* k32[2] = {reinterpret_cast<unsigned*>(ws.sort.keys[rb ^ 1]),
                        reinterpret_cast<unsigned*>(ws.sort.keys[rb ^ 1]) + ws.sort.capacity};
    unsigned* v32[2] = {ws.order_s, ws.sort.vals[rb ^ 1]};
    k_site_keys<<<grid, kThreads, 0, s>>>(ev, p, ws.sort.vals[rb], ws.order_t, k32[0], v32[0], ws.counters);
    launches += 1;
    int rb2 = 0;
    launches += radix_sort_pairs<unsigned>(k32, v32, ws.sort.tile_hist, &ws.counters[1], 0, 32, &rb2, grid, s);
    // 32 bits = 4 passes -> result back in buffer 0 (k32[0], ws.order_s)
    k_deadtime<<<grid, kThreads, 0, s>>>(ev, p, ws.order_t, v32[rb2], k32[rb2], ws.kill, ws.counters);
    k_final_flags<<<grid, kThreads, 0, s>>>(ev, p, ws.order_t, ws.kill, ws.flags, ws.counters);
    launches += 2;
    unsigned* offs = ws.sort.vals[rb];  // time-sort payload no longer needed (order_t holds it)
    launches += exclusive_scan(ws.flags, offs, &ws.counters[1], ws.scan_tmp, nullptr, grid, s);
    k_emit_singles<<<grid, kThreads, 0, s>>>(ev, singles, static_cast<gpet_event*>(singles_aos), ws.order_t, ws.flags, offs,
                                             ws.counters, ws.spectrum, ws.spectrum_bins, ws.spec_emin, ws.spec_emax);
    launches += 1;
    if (p.cwin > 0.f && coinc_aos) {
        unsigned* cnt = ws.flags;
        k_coinc_count<<<grid, kThreads, 0, s>>>(singles, p, cnt);
        launches += 1;
        launches += exclusive_scan(cnt, offs, singles.count, ws.scan_tmp, &ws.counters[4], grid, s);
        k_coinc_emit<<<grid, kThreads, 0, s>>>(singles, p, cnt, offs, static_cast<gpet_coincidence*>(coinc_aos), coinc_cap);
        launches += 1;
    }
    return launches;
}

}  // namespace gpet
