// extern "C" boundary of libgpet_b200.so (see include/gpet_b200.h for the reference call each entry replaces).
#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/types.h>
#include <unistd.h>

#include "ctx.hpp"
#include "planner.hpp"

using namespace gpet;

namespace gpet { thread_local KernelTimer* g_ktimer = nullptr; }

namespace {

struct ProfScope {   // routes the launchers' event brackets to this context while profiling is on
    explicit ProfScope(gpet_ctx* c) { g_ktimer = (c && c->profiling) ? &c->ktimer : nullptr; }
    ~ProfScope() { g_ktimer = nullptr; }
};

int fail(const gpet_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}

#define CK(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return fail(c, GPET_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));             \
    } while (0)

#define NEED_DEVICE()                                                                                      \
    do {                                                                                                   \
        if (!c) return GPET_ERR_ARG;                                                                       \
        if (!c->has_device)                                                                                \
            return fail(c, GPET_ERR_NO_DEVICE, "compute entry point called on a host-only context (no CPU fallback)"); \
    } while (0)

template <typename T>
int dev_alloc(gpet_ctx* c, T** p, size_t n) {
    void* q = nullptr;
    CK(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
    c->allocs.push_back(q);
    *p = static_cast<T*>(q);
    return GPET_OK;
}

int alloc_queue(gpet_ctx* c, PhotonQueue& q, size_t cap, unsigned* count) {
    int r;
    if ((r = dev_alloc(c, &q.pos_e, cap))) return r;
    if ((r = dev_alloc(c, &q.dir_n, cap))) return r;
    if ((r = dev_alloc(c, &q.t, cap))) return r;
    if ((r = dev_alloc(c, &q.ids, cap))) return r;
    q.count = count;
    q.capacity = (unsigned)cap;
    return GPET_OK;
}

int ensure_buffers(gpet_ctx* c) {
    if (c->dev_buffers) return GPET_OK;
    int r;
    const size_t cp = c->cap_photons, ch = c->cap_hits, ce = c->cap_events;
    DigitizerWorkspace& w = c->ws;
    // Per-frame device state in ONE block, cleared by one memset per frame: the 64-word counter block (every device-side
    // counter, so one small D2H brings all of them back), the status words of the three scans, the key range, the slice
    // counters of the time sort.
    w.capacity = (unsigned)ce;
    w.max_tiles = scan_tiles(ce);
    {
        const size_t mt4 = (((size_t)w.max_tiles + 3) & ~(size_t)3) * scan_status_stride();
        const size_t st2 = ((size_t)bucket_words() / 2048 + 4) * scan_status_stride();
        const size_t words = 64 + 2 * mt4 + st2 + 4 + (size_t)bucket_words() + kHotWords + 32;
        unsigned* p = nullptr;
        if ((r = dev_alloc(c, &p, words))) return r;
        CK(cudaMemset(p, 0, words * sizeof(unsigned)));
        w.frame_state = p;
        w.frame_state_bytes = words * sizeof(unsigned);
        w.counters = p; p += 64;
        w.scan_status[0] = p; p += mt4;
        w.scan_status[1] = p; p += mt4;
        w.scan_status[2] = p; p += st2;
        w.minmax = reinterpret_cast<unsigned long long*>(p); p += 4;
        w.bcount = p; p += bucket_words();
        // hot counters on lines of their own, 128-byte aligned (kernels.hpp HotWord)
        w.hot = reinterpret_cast<unsigned*>((reinterpret_cast<uintptr_t>(p) + 127) & ~(uintptr_t)127);
    }
    if ((r = alloc_queue(c, c->q[0], cp, w.counters + 16))) return r;
    if ((r = alloc_queue(c, c->q[1], cp, w.counters + 17))) return r;
    if ((r = alloc_queue(c, c->q[2], cp, w.hot + kHotQ2))) return r;   // photons that entered a panel (panel-local frame)
    if ((r = dev_alloc(c, &c->hits.id4, ch))) return r;
    if ((r = dev_alloc(c, &c->hits.f4, ch))) return r;
    if ((r = dev_alloc(c, &c->hits.t, ch))) return r;
    if ((r = dev_alloc(c, &c->hits.type, ch))) return r;
    c->hits.count = w.hot + kHotHits;
    c->hits.capacity = (unsigned)ch;
    if ((r = dev_alloc(c, &c->ev.rec, ce))) return r;
    c->ev.count = w.hot + kHotEvents;
    c->ev.capacity = (unsigned)ce;
    for (int k = 0; k < 2; k++) {
        if ((r = dev_alloc(c, &w.tkeys[k], ce))) return r;
        if ((r = dev_alloc(c, &w.tvals[k], ce))) return r;
    }
    if ((r = dev_alloc(c, &w.site_of, ce))) return r;
    if ((r = dev_alloc(c, &w.aux, ce))) return r;
    if ((r = dev_alloc(c, &w.bent, ce))) return r;
    if ((r = dev_alloc(c, &w.site_t, ce))) return r;
    if ((r = dev_alloc(c, &w.stime, ce))) return r;
    if ((r = dev_alloc(c, &w.span, ce))) return r;
    for (int k = 0; k < 2; k++) {
        if ((r = dev_alloc(c, &w.lookback[k], sort_lookback_words(ce)))) return r;
        CK(cudaMemset(w.lookback[k], 0, sort_lookback_words(ce) * sizeof(unsigned)));
    }
    if ((r = dev_alloc(c, &w.bstart, (size_t)bucket_words()))) return r;
    if ((r = dev_alloc(c, &w.grid_bar, 4))) return r;
    CK(cudaMemset(w.grid_bar, 0, 4 * sizeof(unsigned)));
    if ((r = dev_alloc(c, &c->d_pair_base, 2))) return r;
    CK(cudaMemset(c->d_pair_base, 0, 2 * sizeof(unsigned)));
    {
        unsigned char* p = nullptr;
        if ((r = dev_alloc(c, &p, sort_state_bytes()))) return r;
        CK(cudaMemset(p, 0, sort_state_bytes()));
        w.st_time = reinterpret_cast<rsort::SortState*>(p);
    }
    if ((r = dev_alloc(c, &w.order_t, ce))) return r;
    if ((r = dev_alloc(c, &w.kill, ce))) return r;
    if ((r = dev_alloc(c, &w.spar, ce))) return r;
    if ((r = dev_alloc(c, &w.seid, ce))) return r;
    if (w.spectrum_bins > 0) {
        w.spectrum_stride = w.spectrum_bins <= 1024 ? 16 : 1;
        const size_t nw = (size_t)w.spectrum_bins * w.spectrum_stride;
        if ((r = dev_alloc(c, &w.spectrum, nw))) return r;
        CK(cudaMemset(w.spectrum, 0, sizeof(unsigned long long) * nw));
    }
    {
        void* p = nullptr;
        c->coinc_cap = (unsigned)(ce / 2);
        for (int k = 0; k < 2; k++) {
            CK(cudaMalloc(&p, ce * sizeof(gpet_event)));
            c->allocs.push_back(p);
            c->singles_slot[k] = p;
            CK(cudaMalloc(&p, (size_t)c->coinc_cap * sizeof(gpet_coincidence)));
            c->allocs.push_back(p);
            c->coinc_slot[k] = p;
            CK(cudaMalloc(&p, (size_t)c->coinc_cap * sizeof(uint2)));
            c->allocs.push_back(p);
            c->pairs_slot[k] = p;
            CK(cudaMalloc(&p, (size_t)c->coinc_cap));
            c->allocs.push_back(p);
            c->cls_slot[k] = p;
            CK(cudaMallocHost((void**)&c->h_slot_counters[k], 64 * sizeof(unsigned)));
            CK(cudaEventCreateWithFlags(&c->ev_counters[k], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->ev_copied[k], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->ev_emit[k], cudaEventDisableTiming));
        }
        CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        c->out_slot = 0;
        c->singles_aos = c->singles_slot[0];
        c->coinc_aos = c->coinc_slot[0];
        c->cls_aos = c->cls_slot[0];
        c->stage_bytes = std::max(std::max(ce * sizeof(gpet_event), cp * sizeof(gpet_photon)), ch * sizeof(gpet_hit));
        CK(cudaMalloc(&p, c->stage_bytes));
        c->allocs.push_back(p);
        c->stage_aos = p;
    }
    {   // scatter tags, one byte per photon slot: zeroed here and whenever the serial wraps (serial 0 is never used)
        size_t nt = 1024;
        while (nt < cp) nt <<= 1;
        if ((r = dev_alloc(c, &c->d_scat_tag, nt))) return r;
        CK(cudaMemset(c->d_scat_tag, 0, nt));
        c->scat_mask = (unsigned)(nt - 1);
        c->scat_serial = 0;
    }
    if ((r = dev_alloc(c, &c->d_totals, 32))) return r;
    CK(cudaMemset(c->d_totals, 0, 32 * sizeof(unsigned long long)));
    CK(cudaMallocHost((void**)&c->h_counters, 64 * sizeof(unsigned)));
    CK(cudaMallocHost((void**)&c->h_totals, 32 * sizeof(unsigned long long)));
    c->dev_buffers = true;
    return GPET_OK;
}

int upload_tables(gpet_ctx* c) {
    if (c->dev_tables) return GPET_OK;
    if (!c->tab.loaded()) return fail(c, GPET_ERR_ARG, "cross-section tables not loaded");
    const Tables& t = c->tab;
    std::vector<float4> xs((size_t)t.nmat * t.nen);
    for (size_t k = 0; k < xs.size(); k++) xs[k] = make_float4(t.lamph[k], t.compt[k], t.rayle[k], t.phote[k]);
    int r;
    if ((r = dev_alloc(c, &c->d_xs, xs.size()))) return r;
    CK(cudaMemcpy(c->d_xs, xs.data(), xs.size() * sizeof(float4), cudaMemcpyHostToDevice));
    if ((r = dev_alloc(c, &c->d_cmpsf, t.cmpsf.size()))) return r;
    CK(cudaMemcpy(c->d_cmpsf, t.cmpsf.data(), t.cmpsf.size() * 4, cudaMemcpyHostToDevice));
    if ((r = dev_alloc(c, &c->d_rayff, t.rayff.size()))) return r;
    CK(cudaMemcpy(c->d_rayff, t.rayff.data(), t.rayff.size() * 4, cudaMemcpyHostToDevice));
    if ((r = dev_alloc(c, &c->d_maj_ph, (size_t)t.nen))) return r;
    if ((r = dev_alloc(c, &c->d_maj_det, (size_t)t.nen))) return r;
    c->dev_tables = true;
    return GPET_OK;
}

TablesDev tables_dev(const gpet_ctx* c) {
    const Tables& t = c->tab;
    TablesDev d{};
    d.xs = c->d_xs;
    d.maj_phantom = c->d_maj_ph;
    d.maj_detector = c->d_maj_det;
    d.cmpsf = c->d_cmpsf;
    d.rayff = c->d_rayff;
    d.nmat = t.nmat;
    d.nen = t.nen;
    d.e0 = t.energy.empty() ? 0.f : t.energy[0];
    d.ide = t.nen > 1 ? (float)(t.nen - 1) / (t.energy[t.nen - 1] - t.energy[0]) : 0.f;  // idleph (initialize.cu:421)
    d.cm_ncp = t.cm_ncp; d.cm_ne = t.cm_ne; d.rl_ncp = t.rl_ncp; d.rl_ne = t.rl_ne;
    d.cm_idcp = 1.0f / t.cm_dcp; d.cm_ide = 1.0f / t.cm_de;   // initialize.cu:526-527
    d.rl_idcp = 1.0f / t.rl_dcp; d.rl_ide = 1.0f / t.rl_de;   // initialize.cu:707-708
    return d;
}

// Keep the voxel grid resident in L2: a persisting access-policy window on the stream the kernels are launched on, when the
// driver's persisting carve-out allows it (cudaDevAttrMaxPersistingL2CacheSize).  Re-applied whenever the stream changes
// (gpet_set_stream); GPET_NO_L2_WINDOW=1 leaves it off (A/B measurements, profiles/).
void apply_l2_window(gpet_ctx* c) {
    if (!c->has_device || !c->d_vox || !c->vox_bytes || !c->stream) return;
    if (getenv("GPET_NO_L2_WINDOW")) return;
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c->device);
    if (max_persist <= 0 || max_window <= 0) return;
    const size_t bytes = std::min<size_t>(c->vox_bytes, (size_t)max_window);
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>(bytes, (size_t)max_persist));
    cudaStreamAttrValue attr{};
    attr.accessPolicyWindow.base_ptr = c->d_vox;
    attr.accessPolicyWindow.num_bytes = bytes;
    attr.accessPolicyWindow.hitRatio = std::min(1.0f, (float)max_persist / (float)bytes);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();
}

int upload_phantom(gpet_ctx* c) {
    if (c->dev_phantom) return GPET_OK;
    if (!c->have_ph) return fail(c, GPET_ERR_ARG, "phantom not loaded");
    int r;
    if ((r = upload_tables(c))) return r;
    const Phantom& ph = c->ph;
    const size_t n = ph.nvox();
    std::vector<uint32_t> vox(n);
    c->ph_present.assign(16, 0);
    for (size_t k = 0; k < n; k++) {
        if (ph.mat[k] >= 0 && ph.mat[k] < 16) c->ph_present[(size_t)ph.mat[k]] = 1;
        uint32_t b;
        float d = ph.dens[k];
        memcpy(&b, &d, 4);
        vox[k] = (b & ~15u) | ((uint32_t)ph.mat[k] & 15u);
    }
    if ((r = dev_alloc(c, &c->d_vox, n))) return r;
    CK(cudaMemcpy(c->d_vox, vox.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_maj_ph, c->maj_ph.data(), c->maj_ph.size() * 4, cudaMemcpyHostToDevice));
    c->vox_bytes = n * 4;
    apply_l2_window(c);
    c->dev_phantom = true;
    return GPET_OK;
}

PhantomDev phantom_dev(const gpet_ctx* c) {
    PhantomDev d{};
    const Phantom& ph = c->ph;
    d.vox = c->d_vox;
    d.nx = ph.dim[0]; d.ny = ph.dim[1]; d.nz = ph.dim[2];
    d.ox = ph.offset[0]; d.oy = ph.offset[1]; d.oz = ph.offset[2];
    d.idx = 1.0f / ph.d[0]; d.idy = 1.0f / ph.d[1]; d.idz = 1.0f / ph.d[2];  // initialize.cu:846-851
    d.dx = ph.d[0]; d.dy = ph.d[1]; d.dz = ph.d[2];
    d.rec_on = c->tr.record_psf != 0;
    for (int i = 0; i < 4; i++) d.rec[i] = c->tr.record_sphere[i];
    // shared-memory table staging (GPET_SMEM_TABLES): the materials present in the phantom, if they fit three slots, and the
    // energy nodes up to 600 keV (annihilation photons never exceed 511 keV (1 + a few acollinearity sigmas))
    d.tab_nstage = 0;
    d.tab_slot_map = ~0ull;
    if (c->tab.loaded() && c->tab.nen > 1) {
        bool present[16] = {};
        int npresent = 0;
        for (int m = 0; m < 16 && m < (int)c->ph_present.size(); m++)
            if (c->ph_present[m]) { present[m] = true; npresent++; }
        if (npresent >= 1 && npresent <= 3) {
            const float e0 = c->tab.energy.front(), e1 = c->tab.energy.back();
            const float ide = (float)(c->tab.nen - 1) / (e1 - e0);
            d.tab_nstage = std::min(c->tab.nen, (int)(ide * (600.0e3f - e0)) + 2);
            int slot = 0;
            for (int m = 0; m < 16; m++)
                if (present[m]) d.tab_slot_map = (d.tab_slot_map & ~(15ull << (4 * m))) | ((unsigned long long)slot++ << (4 * m));
        }
    }
    return d;
}

int upload_geometry(gpet_ctx* c) {
    if (c->dev_geo) return GPET_OK;
    if (!c->have_geo) return fail(c, GPET_ERR_ARG, "detector geometry not loaded");
    int r;
    if ((r = upload_tables(c))) return r;
    const Geometry& g = c->geo;
    std::vector<PanelDev> pd(g.panels.size());
    for (size_t i = 0; i < pd.size(); i++) {
        const gpet_panel& p = g.panels[i];
        PanelDev& d = pd[i];
        memset(&d, 0, sizeof(d));
        d.ox = p.offsetx; d.oy = p.offsety; d.oz = p.offsetz;
        d.uxx = p.UniXx; d.uxy = p.UniXy; d.uxz = p.UniXz;
        d.uyx = p.UniYx; d.uyy = p.UniYy; d.uyz = p.UniYz;
        d.uzx = p.UniZx; d.uzy = p.UniZy; d.uzz = p.UniZz;
        d.lx = p.lengthx; d.ly = p.lengthy; d.lz = p.lengthz;
        d.dirx = p.directionx;
        d.mody = p.MODy; d.modz = p.MODz; d.mspy = p.Mspacey; d.mspz = p.Mspacez;
        d.lsoy = p.LSOy; d.lsoz = p.LSOz; d.spy = p.spacey; d.spz = p.spacez;
        d.id = p.panel;
        d.r2 = 0.25f * (p.lengthy * p.lengthy + p.lengthz * p.lengthz);
        // reciprocals of crystalSearch's divisors (the fp32 sums the kernel forms), rounded once from extended precision
        auto rcp = [](float a, float b) { const float s = a + b; return (float)(1.0L / (long double)s); };
        d.rcp_my = rcp(p.MODy, p.Mspacey); d.rcp_mz = rcp(p.MODz, p.Mspacez);
        d.rcp_cy = rcp(p.LSOy, p.spacey); d.rcp_cz = rcp(p.LSOz, p.spacez);
    }
    if ((r = dev_alloc(c, &c->d_panels, pd.size()))) return r;
    CK(cudaMemcpy(c->d_panels, pd.data(), pd.size() * sizeof(PanelDev), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_maj_det, c->maj_det.data(), c->maj_det.size() * 4, cudaMemcpyHostToDevice));
    c->dev_geo = true;
    return GPET_OK;
}

DetectorDev detector_dev(const gpet_ctx* c) {
    DetectorDev d{};
    const Geometry& g = c->geo;
    d.panels = c->d_panels;
    d.npanels = (int)g.panels.size();
    d.moduleNy = g.moduleNy; d.crystalNy = g.crystalNy; d.moduleN = g.moduleN; d.crystalN = g.crystalN;
    d.mat[0] = g.mat[0]; d.mat[1] = g.mat[1];
    d.dens[0] = g.dens[0]; d.dens[1] = g.dens[1];
    // quadric surfaces that can never exclude a point (all coefficients zero, constant term >= 0: the shipped
    // "0 0 0 0 0 0 0 0 0 1") are dropped here instead of being evaluated at every step of every photon
    // (crystalSearch, gPET_kernals.cu:1241-1245, returns when the quadric is < 0)
    d.nsurface = 0;
    for (int k = 0; k < c->tr.nsurface && k < GPET_MAX_SURFACES; k++) {
        const float* q = c->tr.surface + 10 * k;
        bool inert = q[9] >= 0.f;
        for (int j = 0; j < 9; j++) inert = inert && q[j] == 0.f;
        if (!inert) memcpy(d.surface + 10 * d.nsurface++, q, sizeof(float) * 10);
    }
    // the bounding-sphere rejection in panel_entry needs Euclidean local coordinates: orthonormal axes on every panel
    d.dirmask = (c->in_run && c->dirmask_on) ? c->d_dirmask : nullptr;
    d.scat_tag = c->d_scat_tag; d.scat_mask = c->scat_mask; d.scat_serial = c->scat_serial;
    d.prefilter = g.panels.size() <= 32;
    for (const gpet_panel& p : g.panels) {
        const double u[3][3] = {{p.UniXx, p.UniXy, p.UniXz}, {p.UniYx, p.UniYy, p.UniYz}, {p.UniZx, p.UniZy, p.UniZz}};
        for (int a = 0; a < 3; a++)
            for (int b = a; b < 3; b++) {
                const double dot = u[a][0] * u[b][0] + u[a][1] * u[b][1] + u[a][2] * u[b][2];
                if (std::fabs(dot - (a == b ? 1.0 : 0.0)) > 1e-4) d.prefilter = 0;
            }
    }
    return d;
}

DigitizerDev digitizer_dev(const gpet_ctx* c) {
    DigitizerDev d{};
    const gpet_digitizer_params& p = c->dig;
    d.readout_depth = p.readout_depth; d.readout_policy = p.readout_policy;
    d.Eth = p.threshold_eV;
    d.blur_policy = p.blur_policy; d.Eref = p.blur_Eref; d.Rref = p.blur_Rref; d.slope = p.blur_slope; d.sblur = p.blur_space;
    d.dlevel = p.dead_level; d.dtype = p.dead_type; d.dtime = p.dead_time_us;
    d.Ewinmin = p.ewin_min; d.Ewinmax = p.ewin_max;
    d.tblur = p.time_blur_sigma_us; d.cwin = p.coinc_window_us; d.cpolicy = p.coinc_policy; d.cmindiff = p.coinc_min_panel_diff;
    d.npanels = (int)c->geo.panels.size();
    d.moduleN = c->geo.moduleN; d.crystalN = c->geo.crystalN;
    d.noise_gap = p.noise_mean_gap_us; d.noise_Emean = p.noise_Emean_eV; d.noise_sigma = p.noise_sigma_eV;
    d.noise_interval = p.noise_interval_us;
    d.id_base = c->id_base;
    d.emit_on = c->emit_on ? 1 : 0;
    d.emit_lo = c->emit_lo; d.emit_hi = c->emit_hi; d.trust_lo = c->emit_trust;
    d.pair_shift = std::min(std::max(p.coinc_pair_shift, 0), 31);
    d.tie_site = c->in_run ? 1 : 0;
    d.scat_tag = c->d_scat_tag; d.scat_mask = c->scat_mask; d.scat_serial = c->scat_serial;
    return d;
}

// New photons are about to reach the panel faces, or new events are put: tags of anything earlier must not match.
void new_scatter_serial(gpet_ctx* c) {
    if (++c->scat_serial > 255u) {   // one-byte tags: every 255 serials the table is cleared (4 MB, stream ordered) and the count restarts
        if (c->d_scat_tag) cudaMemsetAsync(c->d_scat_tag, 0, (size_t)c->scat_mask + 1, c->stream);
        c->scat_serial = 1u;
    }
}

void rebuild_majorants(gpet_ctx* c) {
    if (!c->tab.loaded()) return;
    if (c->have_ph) {
        // largest density per material present in the phantom (initialize.cu:794-802)
        std::vector<float> maxden((size_t)c->tab.nmat, 0.f);
        for (size_t k = 0; k < c->ph.nvox(); k++) {
            int m = c->ph.mat[k];
            if (m >= 0 && m < c->tab.nmat && c->ph.dens[k] > maxden[m]) maxden[m] = c->ph.dens[k];
        }
        c->maj_ph = build_majorant(c->tab, maxden);
        c->dev_phantom = false;
    }
    if (c->have_geo) {
        std::vector<float> maxden((size_t)c->tab.nmat, 0.f);
        for (int i = 0; i < 2; i++) {  // initialize.cu:932-936
            int m = c->geo.mat[i];
            if (m >= 0 && m < c->tab.nmat && c->geo.dens[i] > maxden[m]) maxden[m] = c->geo.dens[i];
        }
        c->maj_det = build_majorant(c->tab, maxden);
        c->dev_geo = false;
    }
}

int validate_materials(gpet_ctx* c) {
    if (!c->tab.loaded()) return GPET_OK;
    if (c->have_ph) {
        for (size_t k = 0; k < c->ph.nvox(); k++)
            if (c->ph.mat[k] < 0 || c->ph.mat[k] >= c->tab.nmat || c->ph.mat[k] > 15)
                return fail(c, GPET_ERR_FORMAT, "phantom material id outside the loaded table set");
    }
    if (c->have_geo) {
        for (int i = 0; i < 2; i++)
            if (c->geo.mat[i] < 0 || c->geo.mat[i] >= c->tab.nmat)
                return fail(c, GPET_ERR_FORMAT, "detector material id outside the loaded table set");
    }
    return GPET_OK;
}

// One frame's counters to the host: the 32-word block plus the hot counters (gathered by one strided copy into words
// 32..), asynchronously.  patch_counters() then puts the hot values where the host code reads them.
int fetch_counters_async(gpet_ctx* c, unsigned* h) {
    CK(cudaMemcpyAsync(h, c->ws.counters, 32 * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpy2DAsync(h + 32, 2 * sizeof(unsigned), c->ws.hot, kHotStride * sizeof(unsigned), 2 * sizeof(unsigned), kHotLines,
                         cudaMemcpyDeviceToHost, c->stream));
    return GPET_OK;
}

void patch_counters(unsigned* h) {
    h[21] = h[32 + 2 * (kHotQ2 / kHotStride)];                 // photons on a panel
    h[18] = h[32 + 2 * (kHotHits / kHotStride)];               // hits
    h[19] = h[32 + 2 * (kHotEvents / kHotStride)];             // events
}

int read_counters(gpet_ctx* c) {  // synchronises the stream
    int r;
    if ((r = fetch_counters_async(c, c->h_counters))) return r;
    CK(cudaStreamSynchronize(c->stream));
    patch_counters(c->h_counters);
    return GPET_OK;
}

}  // namespace

// =================================================================================================== lifecycle
extern "C" {

int gpet_abi_version(void) { return GPET_ABI_VERSION; }

int gpet_create(int device, gpet_ctx** out) {
    if (!out) return GPET_ERR_ARG;
    gpet_ctx* c = new gpet_ctx();
    c->device = device;
    // defaults = the shipped example's digitizer block would be set by gpet_load_config_file; keep neutral values
    c->dig.readout_depth = 2; c->dig.readout_policy = 1;
    c->dig.dead_level = 3; c->dig.ewin_max = 2.0e6f;
    c->tr.eabs_eV = 1.0e3f; c->tr.record_hits = 1;
    if (device >= 0) {
        cudaError_t e = cudaSetDevice(device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            std::string msg = std::string("gpet_create: ") + cudaGetErrorString(e);
            delete c;
            *out = nullptr;
            fprintf(stderr, "%s\n", msg.c_str());
            return GPET_ERR_CUDA;
        }
        c->stream = c->own_stream;
        cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device);
        c->has_device = true;
    }
    *out = c;
    return GPET_OK;
}

void gpet_destroy(gpet_ctx* c) {
    if (!c) return;
    if (c->has_device) {
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();
        for (void* p : c->allocs) cudaFree(p);
        if (c->h_counters) cudaFreeHost(c->h_counters);
        if (c->h_totals) cudaFreeHost(c->h_totals);
        for (int k = 0; k < 2; k++) {
            if (c->h_slot_counters[k]) cudaFreeHost(c->h_slot_counters[k]);
            if (c->ev_counters[k]) cudaEventDestroy(c->ev_counters[k]);
            if (c->ev_emit[k]) cudaEventDestroy(c->ev_emit[k]);
            if (c->ev_run[k]) cudaEventDestroy(c->ev_run[k]);
            if (c->ev_copied[k]) cudaEventDestroy(c->ev_copied[k]);
        }
        if (c->res_singles.p) cudaFreeHost(c->res_singles.p);
        if (c->res_coinc.p) cudaFreeHost(c->res_coinc.p);
        if (c->res_pairs.p) cudaFreeHost(c->res_pairs.p);
        if (c->res_cls.p) cudaFreeHost(c->res_cls.p);
        if (c->res_adder.p) cudaFreeHost(c->res_adder.p);
        if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
        if (c->own_stream) cudaStreamDestroy(c->own_stream);
    }
    delete c;
}

const char* gpet_last_error(const gpet_ctx* c) { return c ? c->err.c_str() : "null context"; }

int gpet_set_stream(gpet_ctx* c, void* s) {
    if (!c) return GPET_ERR_ARG;
    c->stream = s ? static_cast<cudaStream_t>(s) : c->own_stream;
    if (c->has_device) {
        cudaSetDevice(c->device);
        apply_l2_window(c);   // the access-policy window is a stream attribute: it follows the stream the kernels run on
    }
    return GPET_OK;
}

int gpet_set_seed(gpet_ctx* c, uint64_t seed) {
    if (!c) return GPET_ERR_ARG;
    c->seed = seed;
    c->planned = false;
    return GPET_OK;
}

int gpet_set_capacity(gpet_ctx* c, uint64_t max_photons, uint64_t max_hits, uint64_t max_events) {
    if (!c) return GPET_ERR_ARG;
    if (c->dev_buffers) return fail(c, GPET_ERR_ARG, "capacity must be set before the first compute call");
    if (max_photons < 64 || max_hits < 64 || max_events < 64 || max_photons > (1ull << 31) || max_hits > (1ull << 29) ||
        max_events > (1ull << 31))
        return fail(c, GPET_ERR_ARG, "capacity out of range");
    c->cap_photons = max_photons; c->cap_hits = max_hits; c->cap_events = max_events;
    c->planned = false;
    return GPET_OK;
}

// =================================================================================================== loaders
int gpet_load_tables(gpet_ctx* c, const char* prefix) {
    if (!c || !prefix) return GPET_ERR_ARG;
    std::string p(prefix), e;
    if (p.size() > 8 && p.compare(p.size() - 8, 8, ".gpettab") == 0) e = load_tables_packed(p, c->tab);
    else e = load_tables_ascii(p, c->tab);
    if (!e.empty()) { c->tab = Tables(); return fail(c, GPET_ERR_IO, e); }
    if (c->dev_tables) return fail(c, GPET_ERR_ARG, "tables already uploaded; create a new context");
    rebuild_majorants(c);
    return validate_materials(c);
}

int gpet_save_tables_packed(gpet_ctx* c, const char* path) {
    if (!c || !path) return GPET_ERR_ARG;
    if (!c->tab.loaded()) return fail(c, GPET_ERR_ARG, "no tables loaded");
    std::string e = save_tables_packed(path, c->tab);
    return e.empty() ? GPET_OK : fail(c, GPET_ERR_IO, e);
}

int gpet_load_phantom_files(gpet_ctx* c, const char* mat_file, const char* den_file, const int32_t dim[3],
                            const float offset[3], const float size[3]) {
    if (!c || !mat_file || !den_file) return GPET_ERR_ARG;
    if (c->dev_phantom) return fail(c, GPET_ERR_ARG, "phantom already uploaded; create a new context");
    std::string e = load_phantom(mat_file, den_file, dim, offset, size, c->ph);
    if (!e.empty()) { c->have_ph = false; return fail(c, GPET_ERR_IO, e); }
    c->have_ph = true;
    rebuild_majorants(c);
    return validate_materials(c);
}

int gpet_set_phantom(gpet_ctx* c, const int32_t* mat, const float* dens, const int32_t dim[3], const float offset[3],
                     const float size[3]) {
    if (!c || !mat || !dens) return GPET_ERR_ARG;
    if (c->dev_phantom) return fail(c, GPET_ERR_ARG, "phantom already uploaded; create a new context");
    Phantom& ph = c->ph;
    for (int i = 0; i < 3; i++) {
        if (dim[i] < 1) return fail(c, GPET_ERR_ARG, "phantom dimension must be positive");
        ph.dim[i] = dim[i]; ph.offset[i] = offset[i]; ph.size[i] = size[i]; ph.d[i] = size[i] / dim[i];
    }
    ph.mat.assign(mat, mat + ph.nvox());
    ph.dens.assign(dens, dens + ph.nvox());
    c->have_ph = true;
    rebuild_majorants(c);
    return validate_materials(c);
}

int gpet_load_geometry(gpet_ctx* c, const char* geo_file) {
    if (!c || !geo_file) return GPET_ERR_ARG;
    if (c->dev_geo) return fail(c, GPET_ERR_ARG, "geometry already uploaded; create a new context");
    Geometry g;
    std::string e = parse_geometry(geo_file, g);
    if (!e.empty()) return fail(c, GPET_ERR_IO, e);
    c->geo = g;
    c->geo_version++;
    c->have_geo = true;
    rebuild_majorants(c);
    return validate_materials(c);
}

int gpet_load_isotopes(gpet_ctx* c, const char* f) {
    if (!c || !f) return GPET_ERR_ARG;
    std::string e = parse_isotopes(f, c->iso);
    if (!e.empty()) { c->have_iso = false; return fail(c, GPET_ERR_IO, e); }
    c->have_iso = true;
    c->planned = false;
    return GPET_OK;
}

int gpet_load_source(gpet_ctx* c, const char* f) {
    if (!c || !f) return GPET_ERR_ARG;
    std::string e = parse_sources(f, c->src);
    if (!e.empty()) { c->have_src = false; return fail(c, GPET_ERR_IO, e); }
    c->have_src = true;
    c->usepsf = 0;
    c->planned = false;
    return GPET_OK;
}

int gpet_load_psf(gpet_ctx* c, const char* f, int64_t max_particles, int ptype) {
    if (!c || !f) return GPET_ERR_ARG;
    if (ptype != 0 && ptype != 1) return fail(c, GPET_ERR_ARG, "psf particle type must be 0 (positron) or 1 (photon)");
    std::string e = load_psf(f, max_particles, ptype, c->psf);
    if (!e.empty()) { c->have_psf = false; return fail(c, GPET_ERR_IO, e); }
    c->psf_version++;
    c->have_psf = true;
    c->usepsf = 1;
    return GPET_OK;
}

int gpet_set_digitizer(gpet_ctx* c, const gpet_digitizer_params* p) {
    if (!c || !p) return GPET_ERR_ARG;
    if (p->readout_depth < 0 || p->readout_depth > 3 || p->readout_policy < 0 || p->readout_policy > 1 ||
        p->dead_level < 0 || p->dead_level > 3 || p->dead_type < 0 || p->dead_type > 1 || p->blur_policy < 0 ||
        p->blur_policy > 1)
        return fail(c, GPET_ERR_ARG, "digitizer parameter out of range");
    c->dig = *p;
    return GPET_OK;
}

int gpet_get_digitizer(const gpet_ctx* c, gpet_digitizer_params* p) {
    if (!c || !p) return GPET_ERR_ARG;
    *p = c->dig;
    return GPET_OK;
}

int gpet_set_transport(gpet_ctx* c, const gpet_transport_params* p) {
    if (!c || !p) return GPET_ERR_ARG;
    if (p->nsurface < 0 || p->nsurface > GPET_MAX_SURFACES) return fail(c, GPET_ERR_ARG, "too many quadric surfaces (MAXSURFACE)");
    if (p->use_positron_range != c->tr.use_positron_range || p->noncollinearity_rad != c->tr.noncollinearity_rad)
        c->planned = false;   // both are baked into the per-frame source descriptors
    c->tr = *p;
    return GPET_OK;
}

int gpet_get_transport(const gpet_ctx* c, gpet_transport_params* p) {
    if (!c || !p) return GPET_ERR_ARG;
    *p = c->tr;
    return GPET_OK;
}

int gpet_set_time_window(gpet_ctx* c, float t0, float t1) {
    if (!c) return GPET_ERR_ARG;
    if (!(t1 > t0) || t0 < 0.f) return fail(c, GPET_ERR_ARG, "time window must satisfy 0 <= tstart < tend");
    c->tstart = t0; c->tend = t1;
    c->planned = false;
    return GPET_OK;
}

int gpet_set_source_atoms(gpet_ctx* c, int i, uint64_t natom) {
    if (!c || !c->have_src || i < 0 || i >= c->src.n()) return GPET_ERR_ARG;
    c->src.natom[i] = natom;
    c->planned = false;
    return GPET_OK;
}

int gpet_set_emit_window(gpet_ctx* c, double lo_us, double hi_us, double halo_start_us) {
    if (!c) return GPET_ERR_ARG;
    if (!(hi_us > lo_us) || halo_start_us > lo_us) return fail(c, GPET_ERR_ARG, "emit window must satisfy halo_start <= lo < hi");
    c->emit_on = true;
    c->emit_lo = lo_us; c->emit_hi = hi_us;
    // decisions about events of the window may rest on events at least one dead time / one coincidence window after the cut
    const double reach = std::max((double)c->dig.dead_time_us, (double)c->dig.coinc_window_us);
    c->emit_trust = std::isfinite(halo_start_us) ? halo_start_us + reach * 1.0001 + 1e-6 : -HUGE_VAL;
    if (c->emit_trust > lo_us) return fail(c, GPET_ERR_ARG, "halo shorter than the dead time / coincidence window");
    return GPET_OK;
}

int gpet_clear_emit_window(gpet_ctx* c) {
    if (!c) return GPET_ERR_ARG;
    c->emit_on = false;
    return GPET_OK;
}

int gpet_get_emit_counts(gpet_ctx* c, uint64_t out[4]) {
    NEED_DEVICE();
    if (!out) return GPET_ERR_ARG;
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((r = read_counters(c))) return r;
    out[0] = c->h_counters[10]; out[1] = c->h_counters[11]; out[2] = c->h_counters[15]; out[3] = c->h_counters[4];
    return GPET_OK;
}

// events between the device event buffer and caller-owned DEVICE memory (the exchange of gpet_b200/multi.py moves them
// between GPUs with NCCL; nothing goes through the host)
int64_t gpet_copy_events_to_device(gpet_ctx* c, void* dst_device, int64_t cap) {
    NEED_DEVICE();
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((r = read_counters(c))) return r;
    const int64_t n = std::min<int64_t>(std::min<unsigned>(c->h_counters[19], c->ev.capacity), cap);
    if (n > 0) {
        if (!dst_device) return GPET_ERR_ARG;
        CK(cudaMemcpyAsync(dst_device, c->ev.rec, (size_t)n * sizeof(gpet_event), cudaMemcpyDeviceToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return n;
}

int gpet_put_events_device(gpet_ctx* c, const void* src_device, int64_t n) {
    NEED_DEVICE();
    if (n < 0 || (n > 0 && !src_device)) return GPET_ERR_ARG;
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((uint64_t)n > c->cap_events) return fail(c, GPET_ERR_CAPACITY, "event list exceeds capacity");
    c->id_base = 0;
    new_scatter_serial(c);
    const unsigned n32 = (unsigned)n;
    if (n) CK(cudaMemcpyAsync(c->ev.rec, src_device, (size_t)n * sizeof(gpet_event), cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpyAsync(c->ev.count, &n32, sizeof(unsigned), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return GPET_OK;
}

int gpet_set_first_pair(gpet_ctx* c, uint64_t first_pair) {
    if (!c) return GPET_ERR_ARG;
    if (first_pair > (1ull << 62)) return fail(c, GPET_ERR_ARG, "first pair index out of range");
    c->first_pair = first_pair;
    c->planned = false;
    return GPET_OK;
}

int gpet_peek_config_device(const char* input_file) {
    if (!input_file) return GPET_ERR_ARG;
    Config cfg;
    if (!parse_config(input_file, cfg).empty()) return GPET_ERR_IO;
    return cfg.device >= 0 ? cfg.device : GPET_ERR_FORMAT;
}

int gpet_set_shard(gpet_ctx* c, int rank, int world) {
    if (!c || world < 1 || rank < 0 || rank >= world) return GPET_ERR_ARG;
    c->rank = rank; c->world = world;
    return GPET_OK;
}

int gpet_load_config_file(gpet_ctx* c, const char* input_file, const char* base_dir, const char* data_dir) {
    if (!c || !input_file) return GPET_ERR_ARG;
    std::string base = base_dir ? base_dir : "";
    Config cfg;
    std::string e = parse_config(input_file, cfg);
    if (!e.empty()) return fail(c, GPET_ERR_IO, e);
    c->cfg = cfg;
    int r;
    gpet_transport_params tr{};
    tr.noncollinearity_rad = cfg.nonangle;
    tr.use_positron_range = cfg.useprange;
    tr.eabs_eV = cfg.eabsph;
    tr.nsurface = cfg.nsurface;
    for (int i = 0; i < 10 * cfg.nsurface; i++) tr.surface[i] = cfg.surface[i];
    tr.record_hits = 1;
    tr.record_psf = 0;   // RECORDPSF = 1 in the shipped constants.h: the sphere of field 14 is parsed but unused
    for (int i = 0; i < 4; i++) tr.record_sphere[i] = cfg.recordsphere[i];
    if ((r = gpet_set_transport(c, &tr))) return r;
    gpet_digitizer_params d = c->dig;
    d.readout_depth = cfg.rdepth; d.readout_policy = cfg.rpolicy;
    d.threshold_eV = cfg.Eth;
    d.blur_policy = cfg.blurpolicy; d.blur_Eref = cfg.Eref; d.blur_Rref = cfg.Rref; d.blur_slope = cfg.Eslope; d.blur_space = cfg.Sblur;
    d.dead_level = cfg.dlevel; d.dead_type = cfg.dtype; d.dead_time_us = cfg.dtime;
    d.ewin_min = cfg.Ewinmin; d.ewin_max = cfg.Ewinmax;
    if ((r = gpet_set_digitizer(c, &d))) return r;
    std::string dd = data_dir ? data_dir : join_path(base, "data");
    if (!c->tab.loaded()) {
        std::string packed = join_path(dd, "input4gPET.gpettab");
        FILE* f = fopen(packed.c_str(), "rb");
        if (f) { fclose(f); r = gpet_load_tables(c, packed.c_str()); }
        else r = gpet_load_tables(c, join_path(dd, "input4gPET").c_str());
        if (r) return r;
    }
    if (c->tr.eabs_eV < c->tab.eminph) return fail(c, GPET_ERR_ARG, "init: Eabs out of range");  // initialize.cu:898-902
    if ((r = gpet_load_phantom_files(c, join_path(base, cfg.matfile).c_str(), join_path(base, cfg.denfile).c_str(), cfg.pdim,
                                     cfg.poffset, cfg.psize)))
        return r;
    if ((r = gpet_load_geometry(c, join_path(base, cfg.geofile).c_str()))) return r;
    if (cfg.usepsf) {
        if ((r = gpet_load_psf(c, join_path(base, cfg.sourcefile).c_str(), cfg.nhist, cfg.ptype))) return r;
    } else {
        if ((r = gpet_load_isotopes(c, join_path(dd, "isotopes.txt").c_str()))) return r;
        if ((r = gpet_load_source(c, join_path(base, cfg.sourcefile).c_str()))) return r;
        if ((r = gpet_set_time_window(c, cfg.tstart, cfg.tend))) return r;
    }
    return GPET_OK;
}

// =================================================================================================== getters
int gpet_get_num_panels(const gpet_ctx* c) { return c ? (int)c->geo.panels.size() : GPET_ERR_ARG; }

int gpet_get_panels(const gpet_ctx* c, gpet_panel* out, int cap) {
    if (!c || !out) return GPET_ERR_ARG;
    int n = std::min<int>(cap, (int)c->geo.panels.size());
    for (int i = 0; i < n; i++) out[i] = c->geo.panels[i];
    return n;
}

int gpet_get_geometry_counts(const gpet_ctx* c, int32_t counts[4], int32_t mat[2], float dens[2]) {
    if (!c || !c->have_geo) return GPET_ERR_ARG;
    counts[0] = c->geo.moduleNy; counts[1] = c->geo.crystalNy; counts[2] = c->geo.moduleN; counts[3] = c->geo.crystalN;
    mat[0] = c->geo.mat[0]; mat[1] = c->geo.mat[1];
    dens[0] = c->geo.dens[0]; dens[1] = c->geo.dens[1];
    return GPET_OK;
}

int gpet_get_table_dims(const gpet_ctx* c, int32_t* nmat, int32_t* nen, float* e0, float* e1, int32_t cmd[2], float cms[2],
                        int32_t rld[2], float rls[2]) {
    if (!c || !c->tab.loaded()) return GPET_ERR_ARG;
    const Tables& t = c->tab;
    *nmat = t.nmat; *nen = t.nen; *e0 = t.energy.front(); *e1 = t.energy.back();
    cmd[0] = t.cm_ncp; cmd[1] = t.cm_ne; cms[0] = t.cm_dcp; cms[1] = t.cm_de;
    rld[0] = t.rl_ncp; rld[1] = t.rl_ne; rls[0] = t.rl_dcp; rls[1] = t.rl_de;
    return GPET_OK;
}

int64_t gpet_get_table(const gpet_ctx* c, int which, float* out, int64_t cap) {
    if (!c || !out || !c->tab.loaded()) return GPET_ERR_ARG;
    const Tables& t = c->tab;
    const std::vector<float>* v = nullptr;
    switch (which) {
        case 0: v = &t.lamph; break;
        case 1: v = &t.compt; break;
        case 2: v = &t.phote; break;
        case 3: v = &t.rayle; break;
        case 4: v = &t.cmpsf; break;
        case 5: v = &t.rayff; break;
        case 6: v = &c->maj_ph; break;
        case 7: v = &c->maj_det; break;
        case 8: v = &t.energy; break;
        default: return GPET_ERR_ARG;
    }
    int64_t n = std::min<int64_t>(cap, (int64_t)v->size());
    memcpy(out, v->data(), (size_t)n * 4);
    return n;
}

int gpet_get_num_sources(const gpet_ctx* c) { return c ? c->src.n() : GPET_ERR_ARG; }

int gpet_get_source(const gpet_ctx* c, int i, uint64_t* natom, int32_t* type, int32_t* shape, float coeff[6]) {
    if (!c || i < 0 || i >= c->src.n()) return GPET_ERR_ARG;
    *natom = c->src.natom[i]; *type = c->src.type[i]; *shape = c->src.shape[i];
    for (int j = 0; j < 6; j++) coeff[j] = c->src.coeff[6 * i + j];
    return GPET_OK;
}

int gpet_get_num_isotopes(const gpet_ctx* c) { return c ? c->iso.n() : GPET_ERR_ARG; }

int gpet_get_isotope(const gpet_ctx* c, int i, float* hl, float* ratio, float coef[8]) {
    if (!c || i < 0 || i >= c->iso.n()) return GPET_ERR_ARG;
    *hl = c->iso.halftime[i]; *ratio = c->iso.ratio[i];
    for (int j = 0; j < 8; j++) coef[j] = c->iso.coef[8 * i + j];
    return GPET_OK;
}

int64_t gpet_get_num_psf(const gpet_ctx* c) { return c ? (int64_t)c->psf.p.size() : GPET_ERR_ARG; }

// =================================================================================================== planning
int64_t gpet_plan_frames(gpet_ctx* c, uint64_t max_pairs) {
    if (!c) return GPET_ERR_ARG;
    if (!c->have_src || !c->have_iso) return fail(c, GPET_ERR_ARG, "source and isotope files must be loaded before planning");
    for (int i = 0; i < c->src.n(); i++)
        if (c->src.type[i] < 0 || c->src.type[i] >= c->iso.n()) return fail(c, GPET_ERR_FORMAT, "source isotope type out of range");
    uint64_t cap_pairs = c->cap_photons / 2;
    if (max_pairs == 0 || max_pairs > cap_pairs) max_pairs = cap_pairs;
    c->max_pairs_per_frame = max_pairs;
    std::string e = plan_frames(c->src, c->iso, c->tstart, c->tend, max_pairs, c->seed, c->first_pair, c->frames);
    if (!e.empty()) return fail(c, GPET_ERR_ARG, e);
    c->planned = true;
    if (c->has_device) {
        // upload the per-frame source descriptors
        std::vector<SourceDev> fr(c->frames.size());
        for (size_t f = 0; f < fr.size(); f++) fill_source_dev(c->src, c->iso, c->frames[f], c->tr.noncollinearity_rad, c->tr.use_positron_range, fr[f]);
        cudaSetDevice(c->device);
        if (fr.size() > c->d_frames_n) {
            int r;
            if ((r = dev_alloc(c, &c->d_frames, fr.size()))) return r;
            c->d_frames_n = fr.size();
        }
        if (!fr.empty()) CK(cudaMemcpy(c->d_frames, fr.data(), fr.size() * sizeof(SourceDev), cudaMemcpyHostToDevice));
    }
    return (int64_t)c->frames.size();
}

int64_t gpet_frame_pairs(const gpet_ctx* c, int64_t f) {
    if (!c || !c->planned || f < 0 || f >= (int64_t)c->frames.size()) return GPET_ERR_ARG;
    return (int64_t)c->frames[(size_t)f].npairs;
}

int gpet_get_frame(const gpet_ctx* c, int64_t f, double* t0, double* dt, uint64_t* first_pair, uint64_t* pairs) {
    if (!c || !c->planned || f < 0 || f >= (int64_t)c->frames.size()) return GPET_ERR_ARG;
    const FramePlan& fp = c->frames[(size_t)f];
    if (t0) *t0 = fp.t0_s;
    if (dt) *dt = fp.dt_s;
    if (first_pair) *first_pair = fp.first_pair;
    if (pairs) for (size_t i = 0; i < fp.pairs.size(); i++) pairs[i] = fp.pairs[i];
    return GPET_OK;
}

// =================================================================================================== stages
int gpet_stage_source(gpet_ctx* c, int64_t f) {
    NEED_DEVICE();
    ProfScope prof(c);
    if (!c->planned) return fail(c, GPET_ERR_ARG, "gpet_plan_frames must be called first");
    if (f < 0 || f >= (int64_t)c->frames.size()) return fail(c, GPET_ERR_ARG, "frame index out of range");
    int r;
    if ((r = ensure_buffers(c))) return r;
    const FramePlan& fp = c->frames[(size_t)f];
    if (2 * fp.npairs > c->cap_photons) return fail(c, GPET_ERR_CAPACITY, "frame exceeds the photon capacity");
    PhantomDev ph{};
    if (c->tr.use_positron_range) {   // the positron range walks the density grid (gPET_kernals.cu:379-414)
        if ((r = upload_phantom(c))) return r;
        ph = phantom_dev(c);
    }
    c->id_base = 2ull * fp.first_pair;   // the queues now hold this frame's photons
    c->stats.kernel_launches += launch_source(c->d_frames + f, fp.npairs, ph, c->q[0], c->seed, c->num_sms, c->stream);
    CK(cudaGetLastError());
    return GPET_OK;
}

int gpet_stage_psf(gpet_ctx* c, int64_t first, int64_t n) {
    NEED_DEVICE();
    if (!c->have_psf) return fail(c, GPET_ERR_ARG, "no PSF loaded");
    if (first < 0 || n < 0 || first + n > (int64_t)c->psf.p.size()) return fail(c, GPET_ERR_ARG, "PSF range out of bounds");
    if (c->psf.ptype == 1) return gpet_put_photons(c, 0, c->psf.p.data() + first, n);
    // positron phase space: every record becomes an annihilation photon pair (setPositionForPhoton)
    ProfScope prof(c);
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((uint64_t)(2 * n) > c->cap_photons) return fail(c, GPET_ERR_CAPACITY, "positron batch exceeds half the photon capacity");
    c->id_base = 0;
    PhantomDev ph{};
    if (c->tr.use_positron_range) {
        if ((r = upload_phantom(c))) return r;
        ph = phantom_dev(c);
    }
    if (n) CK(cudaMemcpyAsync(c->stage_aos, c->psf.p.data() + first, (size_t)n * sizeof(gpet_photon), cudaMemcpyHostToDevice, c->stream));
    c->stats.kernel_launches += launch_psf_positron(c->stage_aos, c->q[0], (unsigned)n, (unsigned long long)first, ph,
                                                    c->tr.noncollinearity_rad, c->tr.use_positron_range, c->seed, c->num_sms, c->stream);
    CK(cudaGetLastError());
    return GPET_OK;
}

int gpet_stage_phantom(gpet_ctx* c) {
    NEED_DEVICE();
    ProfScope prof(c);
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((r = upload_phantom(c))) return r;
    c->stats.kernel_launches += launch_phantom(c->q[0], c->q[1], phantom_dev(c), tables_dev(c), c->tr.eabs_eV, c->seed, c->id_base, c->num_sms, c->stream);
    CK(cudaGetLastError());
    return GPET_OK;
}

int gpet_stage_detector(gpet_ctx* c) {
    NEED_DEVICE();
    ProfScope prof(c);
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((r = upload_geometry(c))) return r;
    new_scatter_serial(c);
    c->stats.kernel_launches += launch_panel_entry(c->q[1], c->q[2], detector_dev(c), c->ws.counters, c->num_sms, c->stream);
    {
        const int nl = launch_detector(c->q[2], detector_dev(c), tables_dev(c), c->tr.eabs_eV, c->dig.readout_depth, c->dig.readout_policy,
                                       c->tr.record_hits, c->hits, c->ev, c->ws.counters, c->ws.hot, c->seed, c->id_base, c->num_sms, c->stream, !c->in_run);
        c->stats.kernel_launches += nl;
    }
    CK(cudaGetLastError());
    return GPET_OK;
}

int gpet_stage_front(gpet_ctx* c, int64_t f) {
    NEED_DEVICE();
    ProfScope prof(c);
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((r = upload_phantom(c))) return r;
    if ((r = upload_geometry(c))) return r;
    const gpet::SourceDev* fr = nullptr;
    unsigned long long npairs = 0;
    if (f >= 0) {
        if (!c->planned) return fail(c, GPET_ERR_ARG, "gpet_plan_frames must be called first");
        if (f >= (int64_t)c->frames.size()) return fail(c, GPET_ERR_ARG, "frame index out of range");
        const FramePlan& fp = c->frames[(size_t)f];
        if (2 * fp.npairs > c->cap_photons) return fail(c, GPET_ERR_CAPACITY, "frame exceeds the photon capacity");
        fr = c->d_frames + f;
        npairs = fp.npairs;
        c->id_base = 2ull * fp.first_pair;
    }
    new_scatter_serial(c);
    PhantomDev ph = phantom_dev(c);   // the fused front end tags a photon the moment it scatters (transport.cu mark_scattered)
    ph.scat_tag = c->d_scat_tag; ph.scat_mask = c->scat_mask; ph.scat_serial = c->scat_serial;
    c->stats.kernel_launches += launch_front(fr, npairs, c->q[0], c->q[1], c->q[2], ph, tables_dev(c), detector_dev(c),
                                             c->tr.eabs_eV, c->ws.counters, c->ws.hot, c->seed, c->id_base, c->num_sms, c->stream, !c->in_run);
    CK(cudaGetLastError());
    return GPET_OK;
}

int gpet_stage_panel_transport(gpet_ctx* c) {
    NEED_DEVICE();
    ProfScope prof(c);
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((r = upload_geometry(c))) return r;
    {
        const int nl = launch_detector(c->q[2], detector_dev(c), tables_dev(c), c->tr.eabs_eV, c->dig.readout_depth, c->dig.readout_policy,
                                       c->tr.record_hits, c->hits, c->ev, c->ws.counters, c->ws.hot, c->seed, c->id_base, c->num_sms, c->stream, !c->in_run);
        c->stats.kernel_launches += nl;
    }
    CK(cudaGetLastError());
    return GPET_OK;
}

int gpet_stage_digitize(gpet_ctx* c) {
    NEED_DEVICE();
    ProfScope prof(c);
    int r;
    if ((r = ensure_buffers(c))) return r;
    DigitizerDev d = digitizer_dev(c);
    DigitizerOut out{};
    out.singles = c->singles_aos;
    out.singles_cap = (unsigned)c->cap_events;
    out.coinc_cap = c->coinc_cap;
    out.cls = c->cls_aos;
    if (c->in_run && c->run_compact) out.singles_compact = c->compact_slot[c->out_slot];
    if (c->in_run && c->early_copy) {
        out.h_singles_count = c->h_slot_counters[c->out_slot] + 48;   // a spare word of the slot's pinned block
        out.ev_after_emit = c->ev_emit[c->out_slot];
    }
    if (c->in_run && c->coinc_format == GPET_COINC_PAIRS) {
        // index pairs into the run's singles list: the base (singles of the earlier frames) travels on the device
        out.pairs = c->pairs_slot[c->out_slot];
        out.pair_base_in = c->d_pair_base + (c->run_frame & 1);
        out.pair_base_out = c->d_pair_base + ((c->run_frame + 1) & 1);
    } else {
        out.coinc = c->coinc_aos;
    }
    c->stats.kernel_launches += launch_digitize(c->ev, out, d, c->ws, c->have_range ? &c->range : nullptr, c->seed, c->num_sms, c->stream, !c->in_run,
                                                !(c->in_run && c->skip_fallback));
    CK(cudaGetLastError());
    return GPET_OK;
}

int gpet_stage_noise(gpet_ctx* c, double t_lo_us, double t_hi_us) {
    NEED_DEVICE();
    ProfScope prof(c);
    int r;
    if ((r = ensure_buffers(c))) return r;
    if (!c->have_geo) return fail(c, GPET_ERR_ARG, "noise singles need the detector geometry (panel / module / crystal counts)");
    const DigitizerDev d = digitizer_dev(c);
    if (d.noise_gap > 0.f && !(d.noise_interval > 0.f)) return fail(c, GPET_ERR_ARG, "noise_interval_us must be positive");
    c->stats.kernel_launches += launch_noise(c->ev, d, t_lo_us, t_hi_us, c->seed, c->num_sms, c->stream);
    CK(cudaGetLastError());
    return GPET_OK;
}

// =================================================================================================== buffer access
int64_t gpet_queue_size(gpet_ctx* c, int which) {
    NEED_DEVICE();
    if (which < 0 || which > 2) return GPET_ERR_ARG;
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((r = read_counters(c))) return r;
    return (int64_t)std::min<unsigned>(c->h_counters[which == 2 ? 21 : 16 + which], c->q[which].capacity);
}

int gpet_put_photons(gpet_ctx* c, int which, const gpet_photon* in, int64_t n) {
    NEED_DEVICE();
    ProfScope prof(c);
    if (which < 0 || which > 2 || n < 0 || (n > 0 && !in)) return GPET_ERR_ARG;
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((uint64_t)n > c->cap_photons) return fail(c, GPET_ERR_CAPACITY, "photon batch exceeds capacity");
    c->id_base = 0;   // the caller's ids as they are
    if (n) CK(cudaMemcpyAsync(c->stage_aos, in, (size_t)n * sizeof(gpet_photon), cudaMemcpyHostToDevice, c->stream));
    if (which == 2) new_scatter_serial(c);   // photons put on the panel faces directly carry no scatter tags
    c->stats.kernel_launches += launch_photons_aos_to_queue(c->stage_aos, c->q[which], (unsigned)n, c->stream);
    CK(cudaGetLastError());
    return GPET_OK;
}

int64_t gpet_fetch_photons(gpet_ctx* c, int which, gpet_photon* out, int64_t cap) {
    int64_t n = gpet_queue_size(c, which);
    if (n < 0) return n;
    n = std::min(n, cap);
    if (n == 0) return 0;
    c->stats.kernel_launches += launch_queue_to_photons_aos(c->q[which], c->stage_aos, c->stream);
    CK(cudaMemcpyAsync(out, c->stage_aos, (size_t)n * sizeof(gpet_photon), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return n;
}

int gpet_put_events(gpet_ctx* c, const gpet_event* in, int64_t n) {
    NEED_DEVICE();
    ProfScope prof(c);
    if (n < 0 || (n > 0 && !in)) return GPET_ERR_ARG;
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((uint64_t)n > c->cap_events) return fail(c, GPET_ERR_CAPACITY, "event list exceeds capacity");
    c->id_base = 0;
    // the device buffer holds the records in the file layout: a plain copy
    new_scatter_serial(c);   // replayed events: no photon is tagged as scattered unless gpet_mark_scattered says so
    const unsigned n32 = (unsigned)n;
    if (n) CK(cudaMemcpyAsync(c->ev.rec, in, (size_t)n * sizeof(gpet_event), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->ev.count, &n32, sizeof(unsigned), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));   // n32 lives on this stack frame
    return GPET_OK;
}

int64_t gpet_fetch_events(gpet_ctx* c, gpet_event* out, int64_t cap) {
    NEED_DEVICE();
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((r = read_counters(c))) return r;
    int64_t n = std::min<int64_t>(std::min<unsigned>(c->h_counters[19], c->ev.capacity), cap);
    if (n <= 0) return 0;
    CK(cudaMemcpyAsync(out, c->ev.rec, (size_t)n * sizeof(gpet_event), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return n;
}

int64_t gpet_fetch_hits(gpet_ctx* c, gpet_hit* out, int64_t cap) {
    NEED_DEVICE();
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((r = read_counters(c))) return r;
    int64_t n = std::min<int64_t>(std::min<unsigned>(c->h_counters[18], c->hits.capacity), cap);
    if (n <= 0) return 0;
    if ((size_t)n * sizeof(gpet_hit) > c->stage_bytes) return fail(c, GPET_ERR_CAPACITY, "hit list exceeds the staging buffer");
    c->stats.kernel_launches += launch_hits_to_aos(c->hits, (unsigned)n, c->stage_aos, c->stream);
    CK(cudaMemcpyAsync(out, c->stage_aos, (size_t)n * sizeof(gpet_hit), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return n;
}

int64_t gpet_fetch_singles(gpet_ctx* c, gpet_event* out, int64_t cap) {
    NEED_DEVICE();
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((r = read_counters(c))) return r;
    int64_t n = std::min<int64_t>(std::min<unsigned>(c->h_counters[3], (unsigned)c->cap_events), cap);
    if (n <= 0) return 0;
    CK(cudaMemcpyAsync(out, c->singles_aos, (size_t)n * sizeof(gpet_event), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return n;
}

int64_t gpet_fetch_coincidences(gpet_ctx* c, gpet_coincidence* out, int64_t cap) {
    NEED_DEVICE();
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((r = read_counters(c))) return r;
    int64_t n = std::min<int64_t>(std::min<unsigned>(c->h_counters[4], c->coinc_cap), cap);
    if (n <= 0) return 0;
    CK(cudaMemcpyAsync(out, c->coinc_aos, (size_t)n * sizeof(gpet_coincidence), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return n;
}

int64_t gpet_fetch_coincidence_classes(gpet_ctx* c, uint8_t* out, int64_t cap, uint64_t totals[3]) {
    NEED_DEVICE();
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((r = read_counters(c))) return r;
    if (totals)
        for (int k = 0; k < 3; k++) totals[k] = c->h_counters[12 + k];
    int64_t n = std::min<int64_t>(std::min<unsigned>(c->h_counters[4], c->coinc_cap), cap);
    if (n <= 0 || !out) return 0;
    CK(cudaMemcpyAsync(out, c->cls_aos, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return n;
}

int gpet_mark_scattered(gpet_ctx* c, const int32_t* parn, int64_t n) {
    NEED_DEVICE();
    ProfScope prof(c);
    if (n < 0 || (n > 0 && !parn)) return GPET_ERR_ARG;
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((uint64_t)n * sizeof(int32_t) > c->stage_bytes) return fail(c, GPET_ERR_CAPACITY, "scatter list exceeds the staging buffer");
    if (n == 0) return GPET_OK;
    if (c->scat_serial == 0u) new_scatter_serial(c);
    CK(cudaMemcpyAsync(c->stage_aos, parn, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    c->stats.kernel_launches += launch_mark_scattered(static_cast<const int*>(c->stage_aos), (unsigned)n, c->d_scat_tag, c->scat_mask,
                                                      c->scat_serial, c->stream);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));   // the caller's list may go away
    return GPET_OK;
}

int gpet_last_counts(gpet_ctx* c, uint64_t counts[4]) {
    NEED_DEVICE();
    int r;
    if ((r = ensure_buffers(c))) return r;
    if ((r = read_counters(c))) return r;
    for (int k = 0; k < 4; k++) counts[k] = c->h_counters[k];
    return GPET_OK;
}

// =================================================================================================== whole path
int gpet_digitize(gpet_ctx* c, const gpet_event* in, int64_t n, gpet_event* out, int64_t cap, int64_t* n_out,
                  uint64_t counts[4]) {
    NEED_DEVICE();
    int r;
    if ((r = gpet_put_events(c, in, n))) return r;
    if ((r = gpet_stage_digitize(c))) return r;
    if ((r = read_counters(c))) return r;
    if (counts)
        for (int k = 0; k < 4; k++) counts[k] = c->h_counters[k];
    int64_t ns = c->h_counters[3];
    if (n_out) *n_out = ns;
    if (ns > cap) return fail(c, GPET_ERR_CAPACITY, "output buffer too small for the singles list");
    if (ns > 0 && out) {
        CK(cudaMemcpyAsync(out, c->singles_aos, (size_t)ns * sizeof(gpet_event), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return GPET_OK;
}

}  // extern "C"

namespace {

// append `bytes` from a device pointer to a file (outputData / outevents, detector.cu:287-307, 387-408)
int append_device(gpet_ctx* c, const std::string& path, const void* dptr, size_t bytes, std::vector<char>& tmp) {
    if (!bytes) return GPET_OK;
    tmp.resize(bytes);
    CK(cudaMemcpyAsync(tmp.data(), dptr, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    FILE* f = fopen(path.c_str(), "ab");
    if (!f) return fail(c, GPET_ERR_IO, "cannot open " + path + " for appending");
    fwrite(tmp.data(), 1, bytes, f);
    fclose(f);
    return GPET_OK;
}

// PSF triplet of one stage queue (gPET.cu:63-88, 296-351): per live photon 7 x float32 into out<tag>.dat, the event id into
// id<tag>.dat and the fp64 time into time<tag>.dat, appended (readOutput.m:36-54)
int dump_psf_queue(gpet_ctx* c, const std::string& od, int which, const char* tag, std::vector<gpet_photon>& buf) {
    int64_t n = gpet_queue_size(c, which);
    if (n < 0) return (int)n;
    buf.resize((size_t)n);
    if (n) {
        n = gpet_fetch_photons(c, which, buf.data(), n);
        if (n < 0) return (int)n;
    }
    FILE* fo = fopen(join_path(od, std::string("out") + tag + ".dat").c_str(), "ab");
    FILE* fi = fopen(join_path(od, std::string("id") + tag + ".dat").c_str(), "ab");
    FILE* ft = fopen(join_path(od, std::string("time") + tag + ".dat").c_str(), "ab");
    if (!fo || !fi || !ft) {
        if (fo) fclose(fo);
        if (fi) fclose(fi);
        if (ft) fclose(ft);
        return fail(c, GPET_ERR_IO, std::string("cannot open the PSF dump files of stage ") + tag);
    }
    std::vector<float> f7;
    std::vector<int32_t> id;
    std::vector<double> tt;
    f7.reserve((size_t)n * 7); id.reserve((size_t)n); tt.reserve((size_t)n);
    for (int64_t k = 0; k < n; k++) {
        const gpet_photon& p = buf[(size_t)k];
        if (!(p.t > 0)) continue;   // gPET.cu:76
        const float row[7] = {p.x, p.y, p.z, p.vx, p.vy, p.vz, p.E};
        f7.insert(f7.end(), row, row + 7);
        id.push_back(p.eventid);
        tt.push_back(p.t);
    }
    fwrite(f7.data(), sizeof(float), f7.size(), fo);
    fwrite(id.data(), sizeof(int32_t), id.size(), fi);
    fwrite(tt.data(), sizeof(double), tt.size(), ft);
    fclose(fo); fclose(fi); fclose(ft);
    return GPET_OK;
}

std::string writer_wait_idle(void* file_writer);   // FileWriter::wait_idle (defined below)

// grow a pinned arena so that `extra` more bytes fit (contents kept); in-flight copies into it must have completed
int arena_reserve(gpet_ctx* c, PinnedArena& a, size_t extra) {
    if (a.size + extra <= a.cap) return GPET_OK;
    CK(cudaStreamSynchronize(c->copy_stream));
    if (c->file_writer) {   // the writer thread may still read from the block that is about to move
        const std::string e = writer_wait_idle(c->file_writer);
        if (!e.empty()) return fail(c, GPET_ERR_IO, e);
    }
    size_t ncap = std::max<size_t>(std::max<size_t>(a.cap * 2, a.size + extra), 1u << 20);
    char* np = nullptr;
    CK(cudaMallocHost((void**)&np, ncap));
    if (a.size) memcpy(np, a.p, a.size);
    if (a.p) cudaFreeHost(a.p);
    a.p = np;
    a.cap = ncap;
    return GPET_OK;
}

constexpr int kRetryWithFallback = 1;   // internal: run_attempt() asks run_impl() for a second attempt

// File runs append every frame's records to the reference's output files (outevents, detector.cu:387-408; gPET.cu:383,
// 424).  The records reach pinned host memory by asynchronous copies; ONE writer thread appends them to the files in the
// order they were queued, so that a frame's file I/O (63 bytes per pair, the slowest stage of a file run by an order of
// magnitude) overlaps the kernels and copies of the following frames.  A job whose data is still in flight carries the
// event that marks the end of its copy.
struct FileWriter {
    // Appends to the output files in the background: a pool of threads writes 8 MB chunks at their final offsets (pwrite), so one
    // frame's adder.dat / singles.dat go to the page cache at several times the rate of one thread's write() loop.  The first push
    // of a path opens it and takes its size as the offset (append semantics, the reference's ios::app, detector.cu:287-307); the
    // writer is the file's only writer until it is destroyed.
    static constexpr int kWorkers = 4;
    static constexpr size_t kChunk = 8u << 20;
    struct Ready {   // a CUDA event the chunks of one push wait for; destroyed with the last of them
        cudaEvent_t ev;
        explicit Ready(cudaEvent_t e) : ev(e) {}
        ~Ready() { if (ev) cudaEventDestroy(ev); }
    };
    struct Job { int fd; off_t off; const char* p; size_t n; std::shared_ptr<Ready> ready; std::string path; };
    struct File { int fd; off_t off; };
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv, cv_idle;
    std::deque<Job> q;
    std::map<std::string, File> files;
    bool stop = false, started = false;
    int busy = 0;
    int device = -1;
    std::string err;
    void start() {
        if (started) return;
        started = true;
        for (int w = 0; w < kWorkers; w++)
            th.emplace_back([this] {
                if (device >= 0) cudaSetDevice(device);   // the jobs' events live in this device's context
                for (;;) {
                    Job j;
                    {
                        std::unique_lock<std::mutex> lk(m);
                        cv.wait(lk, [this] { return stop || !q.empty(); });
                        if (q.empty()) return;
                        j = std::move(q.front());
                        q.pop_front();
                        busy++;
                    }
                    std::string e;
                    if (j.ready && j.ready->ev && cudaEventSynchronize(j.ready->ev) != cudaSuccess) e = "copy of " + j.path + " failed";
                    size_t done = 0;
                    while (e.empty() && done < j.n) {
                        const ssize_t w = pwrite(j.fd, j.p + done, j.n - done, j.off + (off_t)done);
                        if (w <= 0) e = "short write to " + j.path;
                        else done += (size_t)w;
                    }
                    j.ready.reset();
                    {
                        std::lock_guard<std::mutex> lk(m);
                        if (!e.empty() && err.empty()) err = e;
                        busy--;
                    }
                    cv_idle.notify_all();
                }
            });
    }
    void push(const std::string& path, const char* p, size_t n, cudaEvent_t ready = nullptr) {
        start();
        std::shared_ptr<Ready> r = ready ? std::make_shared<Ready>(ready) : nullptr;
        {
            std::lock_guard<std::mutex> lk(m);
            auto it = files.find(path);
            if (it == files.end()) {
                const int fd = open(path.c_str(), O_WRONLY | O_CREAT, 0644);
                if (fd < 0) {
                    if (err.empty()) err = "cannot open " + path + " for appending";
                    return;
                }
                it = files.emplace(path, File{fd, lseek(fd, 0, SEEK_END)}).first;
            }
            File& f = it->second;
            for (size_t o = 0; o < n || (o == 0 && r); o += kChunk) {   // an empty push still retires its event
                const size_t len = std::min(kChunk, n - o);
                q.push_back(Job{f.fd, f.off + (off_t)o, p + o, len, r, path});
                if (n == 0) break;
            }
            f.off += (off_t)n;
        }
        cv.notify_all();
    }
    // all queued appends are in the files (or failed: the first error is returned)
    std::string wait_idle() {
        if (!started) return err;
        std::unique_lock<std::mutex> lk(m);
        cv_idle.wait(lk, [this] { return q.empty() && busy == 0; });
        return err;
    }
    ~FileWriter() {
        if (started) {
            {
                std::lock_guard<std::mutex> lk(m);
                stop = true;
            }
            cv.notify_all();
            for (auto& t : th) t.join();
        }
        for (auto& kv : files) close(kv.second.fd);
    }
};

std::string writer_wait_idle(void* file_writer) { return static_cast<FileWriter*>(file_writer)->wait_idle(); }

struct RunState {
    gpet_stats st{};
    bool resident = false;
    std::string od;
    std::vector<char> tmp;
    // File runs keep their results in the pinned arenas only while these stay small: past kKeepBytes the arenas become
    // per-frame scratch (what is in the files is dropped from memory), as the reference streams every epoch to disk
    // (gPET.cu:383, 424).  `dropped` = singles of the run no longer in the arena (index pairs are run-global).
    bool streaming = false;
    size_t dropped = 0;
    FileWriter writer;
};
constexpr size_t kKeepBytes = 1ull << 30;

// Take frame results out of slot `slot`: wait for its counters, account, start the D2H copies of the records.
int retire_frame(gpet_ctx* c, int slot, RunState& rs) {
    int r;
    size_t ns_early = 0;
    char* dst_early = nullptr;
    const size_t srec = c->run_compact ? sizeof(gpet_single_compact) : sizeof(gpet_event);   // a single on its way to the host
    const void* singles_src = c->run_compact ? c->compact_slot[slot] : c->singles_slot[slot];
    if (c->early_copy) {
        // the singles are final once k_emit_singles is done: their copy starts now and overlaps the coincidence sorter
        CK(cudaEventSynchronize(c->ev_emit[slot]));
        ns_early = std::min<size_t>(c->h_slot_counters[slot][48], (size_t)c->cap_events);
        if ((r = arena_reserve(c, c->res_singles, ns_early * srec))) return r;
        dst_early = c->res_singles.p + c->res_singles.size;
        if (ns_early)
            CK(cudaMemcpyAsync(dst_early, singles_src, ns_early * srec, cudaMemcpyDeviceToHost, c->copy_stream));
    }
    CK(cudaEventSynchronize(c->ev_counters[slot]));
    patch_counters(c->h_slot_counters[slot]);
    const unsigned* h = c->h_slot_counters[slot];
    if (c->skip_fallback && h[5]) return kRetryWithFallback;   // a slice of the time sort overflowed and no fallback was enqueued
    gpet_stats& st = rs.st;
    const uint64_t n_ev = h[19], n_hits = h[18], n_q1 = h[17];
    st.frames++;
    st.photons_phantom_out += n_q1;
    st.photons_on_panel += h[8];
    st.hits += n_hits;
    st.events_adder += h[0];
    st.events_threshold += h[1];
    st.events_deadtime += h[2];
    st.singles += h[3];
    st.coincidences += h[4];
    st.trues += h[12];
    st.scatters += h[13];
    st.randoms += h[14];
    st.overflow_adder += h[9];
    if (n_hits > c->hits.capacity) st.overflow_hits += n_hits - c->hits.capacity;
    if (n_ev > c->ev.capacity) st.overflow_events += n_ev - c->ev.capacity;
    if (n_q1 > c->q[1].capacity) st.overflow_events += n_q1 - c->q[1].capacity;
    if (h[4] > c->coinc_cap) st.overflow_events += h[4] - c->coinc_cap;
    for (int k = 0; k < 32; k++) c->h_counters[k] = h[k];
    for (int k = 0; k < 4; k++) c->last_counts[k] = h[k];
    if (rs.resident) return GPET_OK;
    const size_t ns = std::min<size_t>(h[3], (size_t)c->cap_events), nc = std::min<size_t>(h[4], c->coinc_cap);
    const bool want_coinc = c->dig.coinc_window_us > 0.f;
    const bool as_pairs = c->coinc_format == GPET_COINC_PAIRS;
    PinnedArena& arena_c = as_pairs ? c->res_pairs : c->res_coinc;
    const size_t rec_c = as_pairs ? 2 * sizeof(uint32_t) : sizeof(gpet_coincidence);
    if (!c->early_copy && (r = arena_reserve(c, c->res_singles, ns * srec))) return r;
    if (want_coinc && (r = arena_reserve(c, arena_c, nc * rec_c))) return r;
    if (want_coinc && (r = arena_reserve(c, c->res_cls, nc))) return r;
    char* dst_s = c->early_copy ? dst_early : c->res_singles.p + c->res_singles.size;
    char* dst_c = want_coinc ? arena_c.p + arena_c.size : nullptr;
    char* dst_k = want_coinc ? c->res_cls.p + c->res_cls.size : nullptr;
    const size_t first_single = c->res_singles.size / srec;
    if (c->early_copy && ns_early != ns) return fail(c, GPET_ERR_CUDA, "singles count changed after the emit kernel (internal error)");
    if (ns && !c->early_copy)
        CK(cudaMemcpyAsync(dst_s, singles_src, ns * srec, cudaMemcpyDeviceToHost, c->copy_stream));
    if (want_coinc && nc) {
        CK(cudaMemcpyAsync(dst_c, as_pairs ? c->pairs_slot[slot] : c->coinc_slot[slot], nc * rec_c, cudaMemcpyDeviceToHost,
                           c->copy_stream));
        CK(cudaMemcpyAsync(dst_k, c->cls_slot[slot], nc, cudaMemcpyDeviceToHost, c->copy_stream));
    }
    CK(cudaEventRecord(c->ev_copied[slot], c->copy_stream));
    c->res_singles.size += ns * srec;
    if (want_coinc) {
        arena_c.size += nc * rec_c;
        c->res_cls.size += nc;
    }
    if (!rs.od.empty()) {
        // file dumps with the reference layouts (gPET.cu:367-383, 424): this path runs frame by frame (no pipelining)
        CK(cudaStreamSynchronize(c->copy_stream));
        const size_t nh = std::min<size_t>(n_hits, c->hits.capacity);
        if (c->tr.record_hits) {
            // the SoA hit buffer in the reference's row layout (gPET.cu:367-376), built in the staging buffer
            int* id5 = static_cast<int*>(c->stage_aos);
            float* f5 = reinterpret_cast<float*>(id5 + 5 * nh);
            if (nh * 40 > c->stage_bytes) return fail(c, GPET_ERR_CAPACITY, "hit list exceeds the staging buffer");
            c->stats.kernel_launches += launch_hits_to_rows(c->hits, (unsigned)nh, id5, f5, c->stream);
            if ((r = append_device(c, join_path(rs.od, "HitsID.dat"), id5, nh * 5 * sizeof(int32_t), rs.tmp))) return r;
            if ((r = append_device(c, join_path(rs.od, "Hits.dat"), f5, nh * 5 * sizeof(float), rs.tmp))) return r;
        }
        // adder.dat was appended by run_attempt before the digitizer ran (blur works in place; gPET.cu:383-388)
        rs.writer.push(join_path(rs.od, "singles.dat"), dst_s, ns * sizeof(gpet_event));
        if (want_coinc) {
            if (!as_pairs) {
                rs.writer.push(join_path(rs.od, "coincidences.dat"), dst_c, nc * sizeof(gpet_coincidence));
            } else {   // same file either way: gather the two singles of every pair (synchronous: an extension format, not the hot path)
                const std::string e = rs.writer.wait_idle();
                if (!e.empty()) return fail(c, GPET_ERR_IO, e);
                FILE* fc = fopen(join_path(rs.od, "coincidences.dat").c_str(), "ab");
                if (!fc) return fail(c, GPET_ERR_IO, "cannot open coincidences.dat for appending");
                const gpet_event* sg = reinterpret_cast<const gpet_event*>(c->res_singles.p);
                const uint32_t* pr = reinterpret_cast<const uint32_t*>(dst_c);
                const size_t n_all = c->res_singles.size / sizeof(gpet_event);
                for (size_t k = 0; k < nc; k++) {
                    const size_t ia = (size_t)pr[2 * k] - rs.dropped, ib = (size_t)pr[2 * k + 1] - rs.dropped;   // run-global -> arena
                    if (pr[2 * k] < rs.dropped || ia < first_single || ib >= n_all) { fclose(fc); return fail(c, GPET_ERR_ARG, "coincidence pair out of range"); }
                    fwrite(sg + ia, sizeof(gpet_event), 1, fc);
                    fwrite(sg + ib, sizeof(gpet_event), 1, fc);
                }
                fclose(fc);
            }
            // one class byte per record of coincidences.dat (0 true, 1 scatter, 2 random)
            rs.writer.push(join_path(rs.od, "coincidences_class.dat"), dst_k, nc);
        }
        if (rs.streaming || c->res_singles.size + c->res_coinc.size + c->res_pairs.size + c->res_cls.size + c->res_adder.size > kKeepBytes) {
            const std::string e = rs.writer.wait_idle();   // the arenas are about to be reused
            if (!e.empty()) return fail(c, GPET_ERR_IO, e);
            c->res_adder.size = 0;
            rs.streaming = true;
            c->results_streamed = true;
            rs.dropped += c->res_singles.size / sizeof(gpet_event);
            c->res_singles.size = 0; c->res_coinc.size = 0; c->res_pairs.size = 0; c->res_cls.size = 0;
        }
    }
    return GPET_OK;
}

int run_attempt(gpet_ctx* c, const char* output_dir, bool resident, gpet_stats* stats_out);

// Direction table of the panel search (DetectorDev::dirmask): for every cell of the direction cube the panels a photon
// flying in such a direction can possibly enter, given that its line passes within `rref` of `o` -- true inside
// gpet_run, where a photon's line goes through its birth point (inside a source shape / a PSF record) or through its
// last interaction (inside the phantom box).  A panel is dropped from a cell only if, for every direction of the cell,
// (a) the photon moves against the panel's growth direction or (b) the line misses the bounding sphere of the face by
// a margin -- both with the cell's half diagonal and rounding slack on the safe side.  Off with positron range (the
// reference's range walk can displace the annihilation point without bound).
// reference sphere (o, rref) of the direction table; false when the table does not apply to the loaded inputs
bool dirmask_reference(const gpet_ctx* c, double o[3], double& rref) {
    const Geometry& g = c->geo;
    if (!c->have_geo || !c->have_ph || g.panels.empty() || g.panels.size() > 32 || c->tr.use_positron_range) return false;
    for (int k = 0; k < 3; k++) o[k] = (double)c->ph.offset[k] + 0.5 * (double)c->ph.size[k];
    rref = 0.5 * std::sqrt((double)c->ph.size[0] * c->ph.size[0] + (double)c->ph.size[1] * c->ph.size[1] + (double)c->ph.size[2] * c->ph.size[2]);
    auto reach = [&](double x, double y, double z, double ext) {
        const double dx = x - o[0], dy = y - o[1], dz = z - o[2];
        rref = std::max(rref, std::sqrt(dx * dx + dy * dy + dz * dz) + ext);
    };
    if (c->usepsf) {
        // millions of records: once per (PSF file, phantom position), not once per run
        if (c->psf_reach_key[0] != o[0] || c->psf_reach_key[1] != o[1] || c->psf_reach_key[2] != o[2] || c->psf_reach_key[3] != (double)c->psf_version) {
            double far2 = 0.0;
            for (const gpet_photon& p : c->psf.p) {
                const double dx = p.x - o[0], dy = p.y - o[1], dz = p.z - o[2];
                far2 = std::max(far2, dx * dx + dy * dy + dz * dz);
            }
            c->psf_reach = std::sqrt(far2);
            c->psf_reach_key[0] = o[0]; c->psf_reach_key[1] = o[1]; c->psf_reach_key[2] = o[2]; c->psf_reach_key[3] = (double)c->psf_version;
        }
        rref = std::max(rref, c->psf_reach);
    } else {
        for (int i = 0; i < c->src.n(); i++) {
            const float* q = c->src.coeff.data() + 6 * i;
            const int shape = c->src.shape[(size_t)i];
            double ext;
            if (shape == 1) ext = std::sqrt((double)q[3] * q[3] + 0.25 * (double)q[4] * q[4]);        // cylinder: radius, height
            else if (shape == 2) ext = std::fabs((double)q[3]);                                         // sphere
            else ext = 0.5 * std::sqrt((double)q[3] * q[3] + (double)q[4] * q[4] + (double)q[5] * q[5]);   // box: full lengths
            reach(q[0], q[1], q[2], ext);
        }
    }
    if (!std::isfinite(rref)) return false;
    rref = rref * 1.001 + 1e-3;
    return true;
}

// the table itself (32 k cells x panels: ~0.4 ms of host time, so only when the reference sphere or the geometry changed)
bool build_dirmask(const gpet_ctx* c, double o[3], double& rref, std::vector<unsigned>& tab) {
    if (!dirmask_reference(c, o, rref)) return false;
    const Geometry& g = c->geo;
    const int nb = gpet::kDirBins;
    const double hd = std::sqrt(3.0) / nb + 2e-5;
    const unsigned all = g.panels.size() >= 32 ? ~0u : ((1u << g.panels.size()) - 1u);
    tab.assign((size_t)nb * nb * nb, all);
    for (int iz = 0; iz < nb; iz++)
        for (int iy = 0; iy < nb; iy++)
            for (int ix = 0; ix < nb; ix++) {
                const double cx = -1.0 + (ix + 0.5) * 2.0 / nb, cy = -1.0 + (iy + 0.5) * 2.0 / nb, cz = -1.0 + (iz + 0.5) * 2.0 / nb;
                const double cn = std::sqrt(cx * cx + cy * cy + cz * cz);
                if (std::fabs(cn - 1.0) > hd + 2e-4) continue;   // no direction of (nearly) unit length falls here
                unsigned m = 0;
                for (size_t i = 0; i < g.panels.size(); i++) {
                    const gpet_panel& p = g.panels[i];
                    const double un = std::sqrt((double)p.UniXx * p.UniXx + (double)p.UniXy * p.UniXy + (double)p.UniXz * p.UniXz);
                    const double lvx = cx * p.UniXx + cy * p.UniXy + cz * p.UniXz;
                    const bool dir_ok = p.directionx == 0 ? true : p.directionx > 0 ? lvx >= -(hd * un + 1e-3) : lvx <= hd * un + 1e-3;
                    const double dx = p.offsetx - o[0], dy = p.offsety - o[1], dz = p.offsetz - o[2];
                    const double dn = std::sqrt(dx * dx + dy * dy + dz * dz);
                    const double kx = dy * cz - dz * cy, ky = dz * cx - dx * cz, kz = dx * cy - dy * cx;
                    const double miss = (std::sqrt(kx * kx + ky * ky + kz * kz) - dn * hd) / 1.0002 - rref;
                    const double rp = 0.5 * std::sqrt((double)p.lengthy * p.lengthy + (double)p.lengthz * p.lengthz);
                    const bool near = miss <= 1.006 * rp + 0.01;
                    if (dir_ok && near) m |= 1u << i;
                }
                tab[((size_t)iz * nb + iy) * nb + ix] = m;
            }
    return true;
}

int prepare_dirmask(gpet_ctx* c) {
    c->dirmask_on = false;
    if (getenv("GPET_NO_DIRMASK")) return GPET_OK;
    double o[3], rref = 0.0;
    // cheap key first: reference sphere and geometry version
    if (!dirmask_reference(c, o, rref)) return GPET_OK;
    const double key[5] = {o[0], o[1], o[2], rref, (double)c->geo_version};
    bool same = c->d_dirmask != nullptr;
    for (int k = 0; k < 5; k++) same = same && key[k] == c->dirmask_key[k];
    if (!same) {
        std::vector<unsigned> tab;
        if (!build_dirmask(c, o, rref, tab)) return GPET_OK;
        int r;
        if (!c->d_dirmask && (r = dev_alloc(c, &c->d_dirmask, tab.size()))) return r;
        CK(cudaMemcpyAsync(c->d_dirmask, tab.data(), tab.size() * sizeof(unsigned), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));   // `tab` is a local
        for (int k = 0; k < 5; k++) c->dirmask_key[k] = key[k];
    }
    c->dirmask_on = true;
    return GPET_OK;
}

// Source-mode decay times are spread over the frame, so the bucket sort of the time keys never overflows a slice there
// and the idle launch of its LSD fallback is skipped; should a frame raise the overflow flag after all (counter-based
// RNG: every frame can be regenerated), the whole run is repeated with the fallback enqueued.  PSF mode and file dumps
// keep the fallback (arbitrary input times; files are appended frame by frame).
int run_impl(gpet_ctx* c, const char* output_dir, bool resident, gpet_stats* stats_out) {
    const bool speculate = c->usepsf == 0 && !(output_dir && *output_dir);
    c->skip_fallback = speculate;
    int r = run_attempt(c, output_dir, resident, stats_out);
    c->skip_fallback = false;
    if (r == kRetryWithFallback) r = run_attempt(c, output_dir, resident, stats_out);
    return r;
}

int run_attempt(gpet_ctx* c, const char* output_dir, bool resident, gpet_stats* stats_out) {
    int r;
    if ((r = ensure_buffers(c))) return r;
    const bool psf_mode = c->usepsf != 0;
    if (!psf_mode && !c->planned) {
        int64_t nf = gpet_plan_frames(c, c->max_pairs_per_frame);
        if (nf < 0) return (int)nf;
    }
    if (psf_mode && !c->have_psf) return fail(c, GPET_ERR_ARG, "no PSF loaded");
    if ((r = prepare_dirmask(c))) return r;
    RunState rs;
    rs.resident = resident;
    rs.od = output_dir ? output_dir : "";
    gpet_stats& st = rs.st;
    const uint64_t launches0 = c->stats.kernel_launches;
    CK(cudaStreamSynchronize(c->copy_stream));
    c->res_singles.size = 0;
    c->res_coinc.size = 0;
    c->res_pairs.size = 0;
    c->res_cls.size = 0;
    c->res_adder.size = 0;
    c->results_streamed = false;
    c->file_writer = &rs.writer;
    rs.writer.device = c->device;
    struct ClearWriter { gpet_ctx* c; ~ClearWriter() { c->file_writer = nullptr; } } clear_writer{c};
    c->coinc_expanded.clear();
    c->singles_expanded.clear();
    c->run_compact = false;
    if (c->singles_format == GPET_SINGLES_COMPACT && !resident && rs.od.empty()) {
        // 32-byte singles: only where gpet_result_singles can rebuild the 48-byte records bit for bit (include/gpet_b200.h)
        if (psf_mode) return fail(c, GPET_ERR_ARG, "compact singles need device-numbered photons (source mode, not PSF input)");
        if (c->dig.noise_mean_gap_us > 0.f) return fail(c, GPET_ERR_ARG, "compact singles cannot carry noise singles");
        int max_id = 0;
        for (const auto& p : c->geo.panels) max_id = std::max(max_id, p.panel < 0 ? 256 : (int)p.panel);
        if (!c->have_geo || max_id > 255 || c->geo.moduleN > 4096 || c->geo.crystalN > 2048)
            return fail(c, GPET_ERR_ARG, "compact singles: panel ids <= 255, <= 4096 modules per panel, <= 2048 crystals per module");
        for (int k = 0; k < 2; k++)
            if (!c->compact_slot[k]) {
                char* p = nullptr;
                if ((r = dev_alloc(c, &p, (size_t)c->cap_events * sizeof(gpet_single_compact)))) return r;
                c->compact_slot[k] = p;
            }
        c->run_compact = true;
    }
    CK(cudaMemsetAsync(c->d_pair_base, 0, 2 * sizeof(unsigned), c->stream));
    struct InRun {   // gpet_stage_digitize reads these while the run is in flight
        gpet_ctx* c;
        explicit InRun(gpet_ctx* c_) : c(c_) { c->in_run = true; c->run_frame = 0; }
        ~InRun() { c->in_run = false; c->have_range = false; c->early_copy = false; }
    } in_run(c);
    if (!c->ev_run[0]) {
        CK(cudaEventCreate(&c->ev_run[0]));
        CK(cudaEventCreate(&c->ev_run[1]));
    }
    cudaEvent_t e0 = c->ev_run[0], e1 = c->ev_run[1];
    CK(cudaEventRecord(e0, c->stream));
    // simulateParticle batches: NPART photons, or NPART/2 positrons that become NPART photons (gPET.cu:33-44)
    const int64_t psf_batch = (psf_mode && c->psf.ptype == 0) ? (int64_t)(c->cap_photons / 2) : (int64_t)c->cap_photons;
    const int64_t nframes = psf_mode ? ((int64_t)c->psf.p.size() + psf_batch - 1) / psf_batch : (int64_t)c->frames.size();
    const bool pipelined = rs.od.empty();   // file dumps read the (single-buffered) hit and event buffers frame by frame
    c->early_copy = pipelined && !resident;
    const int psf_out = rs.od.empty() ? 0 : c->psf_output;
    std::vector<gpet_photon> psf_buf;
    int64_t k = 0;                           // owned frames launched so far
    int rc = GPET_OK;
    for (int64_t f = 0; f < nframes && rc == GPET_OK; f++) {
        if (f % c->world != c->rank) continue;
        // an empty frame still has its noise singles (addnoise covers the frame's time slice, decays or not)
        if (!psf_mode && c->frames[(size_t)f].npairs == 0 && !(c->dig.noise_mean_gap_us > 0.f)) continue;
        const int slot = (int)(k & 1);
        c->out_slot = slot;
        c->singles_aos = c->singles_slot[slot];
        c->coinc_aos = c->coinc_slot[slot];
        c->cls_aos = c->cls_slot[slot];
        if (k >= 2 && !resident) CK(cudaStreamWaitEvent(c->stream, c->ev_copied[slot], 0));
        // one memset clears every counter, ticket, status word and slice counter of the frame
        CK(cudaMemsetAsync(c->ws.frame_state, 0, c->ws.frame_state_bytes, c->stream));
        // fused front end: photons stay in registers from birth (or from queue 0 in PSF mode) to the panel face
        if (psf_mode) {
            int64_t first = f * psf_batch, n = std::min<int64_t>(psf_batch, (int64_t)c->psf.p.size() - first);
            if ((rc = gpet_stage_psf(c, first, n))) break;
            st.pairs += c->psf.ptype == 0 ? (uint64_t)n : (uint64_t)n / 2;
        } else {
            st.pairs += c->frames[(size_t)f].npairs;
        }
        if (psf_out) {
            // phase-space dumps need the stage queues in memory: staged kernels (same photons as the fused front end)
            if (!psf_mode && (rc = gpet_stage_source(c, f))) break;
            if ((psf_mode && psf_out == 1) || (!psf_mode && psf_out == 2))
                if ((rc = dump_psf_queue(c, rs.od, 0, "source", psf_buf))) break;
            if ((rc = gpet_stage_phantom(c))) break;
            if (psf_out == 2 && (rc = dump_psf_queue(c, rs.od, 1, "phantom", psf_buf))) break;
            if ((rc = gpet_stage_detector(c))) break;
        } else {
            if ((rc = gpet_stage_front(c, psf_mode ? -1 : f))) break;
            if ((rc = gpet_stage_panel_transport(c))) break;
        }
        // source mode: event times lie in the frame's slice (plus a flight time far below a slice of the sort)
        c->have_range = !psf_mode;
        if (!psf_mode) {
            const FramePlan& fp = c->frames[(size_t)f];
            c->range = time_range_us(fp.t0_s * 1e6, (fp.t0_s + fp.dt_s) * 1e6);
        }
        if (c->dig.noise_mean_gap_us > 0.f) {
            // addnoise over the time the frame covers (PSF mode: the span of the batch's own times)
            double lo = 0.0, hi = 0.0;
            if (!psf_mode) {
                lo = c->frames[(size_t)f].t0_s * 1e6;
                hi = (c->frames[(size_t)f].t0_s + c->frames[(size_t)f].dt_s) * 1e6;
            } else {
                const int64_t first = f * psf_batch, n = std::min<int64_t>(psf_batch, (int64_t)c->psf.p.size() - first);
                lo = 1e300; hi = -1e300;
                for (int64_t k = first; k < first + n; k++) {
                    const double t = c->psf.p[(size_t)k].t;
                    if (t > 0) { lo = std::min(lo, t); hi = std::max(hi, t); }
                }
                hi = std::nextafter(hi, 1e300);
            }
            if (hi > lo && (rc = gpet_stage_noise(c, lo, hi))) break;
        }
        if (!pipelined) {
            // adder.dat: the post-readout events as the reference writes them, BEFORE blur (gPET.cu:383-388); the digitizer
            // blurs, relabels siten and kills in place
            if ((rc = read_counters(c))) break;
            const size_t nb = std::min<size_t>(c->h_counters[19], c->ev.capacity) * sizeof(gpet_event);
            if ((rc = arena_reserve(c, c->res_adder, nb))) break;
            char* dst = c->res_adder.p + c->res_adder.size;
            c->res_adder.size += nb;
            if (nb) {
                // copied on the compute stream, i.e. before the digitizer touches the records; the writer waits for the event
                cudaEvent_t done = nullptr;
                if (cudaEventCreateWithFlags(&done, cudaEventDisableTiming) != cudaSuccess ||
                    cudaMemcpyAsync(dst, c->ev.rec, nb, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                    cudaEventRecord(done, c->stream) != cudaSuccess) { rc = fail(c, GPET_ERR_CUDA, "adder.dat copy failed"); break; }
                rs.writer.push(join_path(rs.od, "adder.dat"), dst, nb, done);
            }
        }
        c->run_frame = k;
        rc = gpet_stage_digitize(c);
        c->have_range = false;
        if (rc) break;
        c->stats.kernel_launches += launch_publish_counters(c->ws.counters, c->ws.hot, c->h_slot_counters[slot], c->stream);
        CK(cudaEventRecord(c->ev_counters[slot], c->stream));
        k++;
        if (!pipelined) rc = retire_frame(c, slot, rs);
        else if (k >= 2) rc = retire_frame(c, slot ^ 1, rs);
    }
    if (rc == GPET_OK && pipelined && k >= 1) rc = retire_frame(c, (int)((k - 1) & 1), rs);
    {   // the files are complete when the run returns
        const std::string e = rs.writer.wait_idle();
        if (!e.empty() && rc == GPET_OK) rc = fail(c, GPET_ERR_IO, e);
    }
    if (rc != GPET_OK) {
        cudaStreamSynchronize(c->stream);
        cudaStreamSynchronize(c->copy_stream);
        return rc;
    }
    CK(cudaEventRecord(e1, c->stream));
    CK(cudaEventSynchronize(e1));
    CK(cudaStreamSynchronize(c->copy_stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    st.ms_total = ms;
    st.kernel_launches = c->stats.kernel_launches - launches0;
    const uint64_t keep = c->stats.kernel_launches;
    c->stats = st;
    c->stats.kernel_launches = keep;
    if (stats_out) *stats_out = st;
    if (st.overflow_hits || st.overflow_events)
        return fail(c, GPET_ERR_CAPACITY, "a device buffer overflowed; raise gpet_set_capacity");
    return GPET_OK;
}

}  // namespace

extern "C" {

static const char* const kStreamedMsg = "the run's results were streamed to the output files (more than 1 GiB): read them there";

int gpet_run(gpet_ctx* c, const char* output_dir, gpet_stats* stats) {
    NEED_DEVICE();
    return run_impl(c, output_dir, false, stats);
}

int gpet_run_resident(gpet_ctx* c, gpet_stats* stats) {
    NEED_DEVICE();
    return run_impl(c, nullptr, true, stats);
}

// 48-byte records from the 32-byte ones of a GPET_SINGLES_COMPACT run (see include/gpet_b200.h for the identities)
static void expand_compact(const gpet_ctx* c, const gpet_single_compact* in, size_t n, gpet_event* out) {
    {
        const int dlevel = c->dig.dead_level;
        const int rdepth = c->dig.readout_depth, rpolicy = c->dig.readout_policy;
        const int depth = (rdepth != 3 && rpolicy == 1) ? 2 : rdepth;   // the detector kernel's readout level (transport.cu)
        const int level = dlevel == 3 ? depth : dlevel;
        const int moduleN = c->geo.moduleN, crystalN = c->geo.crystalN;
        for (size_t i = 0; i < n; i++) {
            const gpet_single_compact& s = in[i];
            gpet_event e;
            memset(&e, 0, sizeof(e));
            e.pann = (int32_t)(s.ids & 0xffu); e.modn = (int32_t)((s.ids >> 8) & 0xfffu); e.cryn = (int32_t)((s.ids >> 20) & 0x7ffu);
            e.eventid = s.eventid;
            e.parn = (int32_t)((((uint32_t)s.eventid << 1) | (s.ids >> 31)) & 0x7fffffffu);
            e.siten = level <= 0 ? 0 : level == 1 ? e.pann : level == 2 ? e.pann * moduleN + e.modn : (e.pann * moduleN + e.modn) * crystalN + e.cryn;
            e.t = s.t; e.E = s.E; e.x = s.x; e.y = s.y; e.z = s.z;
            out[i] = e;
        }
    }
}

static const gpet_event* expanded_singles(gpet_ctx* c, size_t& n) {
    n = c->res_singles.size / sizeof(gpet_single_compact);
    if (c->singles_expanded.size() != n * sizeof(gpet_event)) {
        c->singles_expanded.resize(n * sizeof(gpet_event));
        expand_compact(c, reinterpret_cast<const gpet_single_compact*>(c->res_singles.p), n, reinterpret_cast<gpet_event*>(c->singles_expanded.data()));
    }
    return reinterpret_cast<const gpet_event*>(c->singles_expanded.data());
}

int gpet_expand_singles(const gpet_ctx* c, const gpet_single_compact* in, int64_t n, gpet_event* out) {
    if (!c || n < 0 || (n > 0 && (!in || !out))) return GPET_ERR_ARG;
    if (!c->have_geo) return fail(c, GPET_ERR_ARG, "gpet_expand_singles needs the detector geometry (module / crystal counts)");
    expand_compact(c, in, (size_t)n, out);
    return GPET_OK;
}

int64_t gpet_result_singles(gpet_ctx* c, const gpet_event** ptr) {
    if (!c || !ptr) return GPET_ERR_ARG;
    if (c->results_streamed) return fail(c, GPET_ERR_CAPACITY, kStreamedMsg);
    if (c->run_compact) {
        size_t n = 0;
        *ptr = expanded_singles(c, n);
        return (int64_t)n;
    }
    *ptr = reinterpret_cast<const gpet_event*>(c->res_singles.p);
    return (int64_t)(c->res_singles.size / sizeof(gpet_event));
}

int64_t gpet_result_singles_compact(gpet_ctx* c, const gpet_single_compact** ptr) {
    if (!c || !ptr) return GPET_ERR_ARG;
    if (!c->run_compact) return fail(c, GPET_ERR_ARG, "the last run did not deliver compact singles (gpet_set_singles_format)");
    *ptr = reinterpret_cast<const gpet_single_compact*>(c->res_singles.p);
    return (int64_t)(c->res_singles.size / sizeof(gpet_single_compact));
}

int gpet_set_singles_format(gpet_ctx* c, int format) {
    if (!c || (format != GPET_SINGLES_RECORDS && format != GPET_SINGLES_COMPACT)) return GPET_ERR_ARG;
    c->singles_format = format;
    return GPET_OK;
}

int64_t gpet_result_coincidences(gpet_ctx* c, const gpet_coincidence** ptr) {
    if (!c || !ptr) return GPET_ERR_ARG;
    if (c->results_streamed) return fail(c, GPET_ERR_CAPACITY, kStreamedMsg);
    if (c->coinc_format == GPET_COINC_PAIRS) {   // records on demand: gather from the singles list
        const size_t np = c->res_pairs.size / (2 * sizeof(uint32_t));
        size_t ns = c->res_singles.size / sizeof(gpet_event);
        const gpet_event* sg = reinterpret_cast<const gpet_event*>(c->res_singles.p);
        if (c->run_compact) sg = expanded_singles(c, ns);
        if (c->coinc_expanded.size() != np * sizeof(gpet_coincidence)) {
            c->coinc_expanded.resize(np * sizeof(gpet_coincidence));
            const uint32_t* pr = reinterpret_cast<const uint32_t*>(c->res_pairs.p);
            gpet_coincidence* out = reinterpret_cast<gpet_coincidence*>(c->coinc_expanded.data());
            for (size_t k = 0; k < np; k++) {
                if (pr[2 * k] >= ns || pr[2 * k + 1] >= ns) return fail(c, GPET_ERR_ARG, "coincidence pair out of range");
                out[k].a = sg[pr[2 * k]];
                out[k].b = sg[pr[2 * k + 1]];
            }
        }
        *ptr = reinterpret_cast<const gpet_coincidence*>(c->coinc_expanded.data());
        return (int64_t)np;
    }
    *ptr = reinterpret_cast<const gpet_coincidence*>(c->res_coinc.p);
    return (int64_t)(c->res_coinc.size / sizeof(gpet_coincidence));
}

int gpet_set_coincidence_format(gpet_ctx* c, int format) {
    if (!c || (format != GPET_COINC_RECORDS && format != GPET_COINC_PAIRS)) return GPET_ERR_ARG;
    c->coinc_format = format;
    return GPET_OK;
}

int64_t gpet_get_direction_table(const gpet_ctx* c, uint32_t* out, int64_t cap, double ref_sphere[4]) {
    if (!c) return GPET_ERR_ARG;
    double o[3], rref = 0.0;
    std::vector<unsigned> tab;
    if (!build_dirmask(c, o, rref, tab)) return 0;
    if (ref_sphere) { ref_sphere[0] = o[0]; ref_sphere[1] = o[1]; ref_sphere[2] = o[2]; ref_sphere[3] = rref; }
    if (out) {
        if (cap < (int64_t)tab.size()) return GPET_ERR_CAPACITY;
        for (size_t k = 0; k < tab.size(); k++) out[k] = tab[k];
    }
    return (int64_t)tab.size();
}

int gpet_set_psf_output(gpet_ctx* c, int mode) {
    if (!c || mode < 0 || mode > 2) return GPET_ERR_ARG;
    c->psf_output = mode;
    return GPET_OK;
}

int64_t gpet_result_coincidence_classes(gpet_ctx* c, const uint8_t** ptr) {
    if (!c || !ptr) return GPET_ERR_ARG;
    if (c->results_streamed) return fail(c, GPET_ERR_CAPACITY, kStreamedMsg);
    *ptr = reinterpret_cast<const uint8_t*>(c->res_cls.p);
    return (int64_t)c->res_cls.size;
}

int64_t gpet_result_coincidence_pairs(gpet_ctx* c, const uint32_t** ptr) {
    if (!c || !ptr) return GPET_ERR_ARG;
    if (c->results_streamed) return fail(c, GPET_ERR_CAPACITY, kStreamedMsg);
    *ptr = reinterpret_cast<const uint32_t*>(c->res_pairs.p);
    return (int64_t)(c->res_pairs.size / (2 * sizeof(uint32_t)));
}

int gpet_get_stats(const gpet_ctx* c, gpet_stats* s) {
    if (!c || !s) return GPET_ERR_ARG;
    *s = c->stats;
    return GPET_OK;
}

int gpet_set_spectrum(gpet_ctx* c, int nbins, float emin, float emax) {
    if (!c || nbins < 1 || nbins > (1 << 20) || !(emax > emin)) return GPET_ERR_ARG;
    if (c->dev_buffers) return fail(c, GPET_ERR_ARG, "spectrum must be configured before the first compute call");
    c->ws.spectrum_bins = nbins; c->ws.spec_emin = emin; c->ws.spec_emax = emax;
    return GPET_OK;
}

int gpet_profile_enable(gpet_ctx* c, int on) {
    NEED_DEVICE();
    CK(cudaStreamSynchronize(c->stream));
    c->ktimer.reset();
    c->profiling = on != 0;
    return GPET_OK;
}

int gpet_profile_count(gpet_ctx* c) {
    NEED_DEVICE();
    CK(cudaStreamSynchronize(c->stream));
    c->ktimer.collect();
    return (int)c->ktimer.order.size();
}

int gpet_profile_get(gpet_ctx* c, int i, char* name, int name_cap, double* total_ms, uint64_t* launches) {
    if (!c || i < 0 || i >= (int)c->ktimer.order.size() || !name || name_cap < 1) return GPET_ERR_ARG;
    const std::string& n = c->ktimer.order[(size_t)i];
    snprintf(name, (size_t)name_cap, "%s", n.c_str());
    const KernelTimer::Acc& a = c->ktimer.acc[n];
    if (total_ms) *total_ms = a.ms;
    if (launches) *launches = a.launches;
    return GPET_OK;
}

int gpet_get_spectrum(gpet_ctx* c, uint64_t* bins, int nbins) {
    NEED_DEVICE();
    if (!bins || nbins != c->ws.spectrum_bins || !c->ws.spectrum) return fail(c, GPET_ERR_ARG, "spectrum not configured");
    CK(cudaMemcpy2DAsync(bins, sizeof(uint64_t), c->ws.spectrum, sizeof(uint64_t) * c->ws.spectrum_stride, sizeof(uint64_t), (size_t)nbins,
                         cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return GPET_OK;
}

}  // extern "C"
