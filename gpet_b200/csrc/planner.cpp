#include "planner.hpp"

#include <cmath>
#include <cstring>

namespace gpet {

void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t x0 = ctr[0], x1 = ctr[1], x2 = ctr[2], x3 = ctr[3], a = key[0], b = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * x0, p1 = (uint64_t)0xCD9E8D57u * x2;
        uint32_t y0 = (uint32_t)(p1 >> 32) ^ x1 ^ a, y1 = (uint32_t)p1, y2 = (uint32_t)(p0 >> 32) ^ x3 ^ b, y3 = (uint32_t)p0;
        x0 = y0; x1 = y1; x2 = y2; x3 = y3;
        a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
}

namespace {
struct HostRng {
    uint32_t key[2], ctr[4], buf[4];
    int have = 0;
    HostRng(uint64_t seed, uint64_t index, uint32_t stream) {
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
        ctr[0] = (uint32_t)index; ctr[1] = (uint32_t)(index >> 32); ctr[2] = stream; ctr[3] = 0;
    }
    double uniform() {  // (0,1), 53 bits
        if (have < 2) { philox4x32_10(ctr, key, buf); ctr[3]++; have = 4; }
        uint32_t a = buf[4 - have], b = buf[5 - have];
        have -= 2;
        uint64_t k = ((uint64_t)a << 21) | (uint64_t)(b >> 11);
        return (double)k * 1.1102230246251565e-16 + 5.551115123125783e-17;
    }
    double normal() {
        double u1 = uniform(), u2 = uniform();
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

uint64_t binomial_rng(uint64_t n, double p, HostRng& g) {
    if (n == 0 || p <= 0.0) return 0;
    if (p >= 1.0) return n;
    if (p > 0.5) return n - binomial_rng(n, 1.0 - p, g);
    const double np = (double)n * p, var = np * (1.0 - p);
    if (var >= 400.0) {
        // Gaussian limit with continuity rounding; skewness error < 1% of a sigma here
        double k = std::floor(np + std::sqrt(var) * g.normal() + 0.5);
        if (k < 0.0) k = 0.0;
        if (k > (double)n) k = (double)n;
        return (uint64_t)k;
    }
    // exact: count successes through geometric waiting times, O(np) draws
    const double lq = std::log1p(-p);
    uint64_t count = 0;
    double pos = 0.0;
    for (;;) {
        pos += std::floor(std::log(g.uniform()) / lq) + 1.0;
        if (pos > (double)n) break;
        count++;
    }
    return count;
}
}  // namespace

uint64_t sample_binomial(uint64_t n, double p, uint64_t seed, uint64_t stream, uint64_t index) {
    HostRng g(seed, index, ((uint32_t)kStagePlan << 24) | (uint32_t)(stream & 0xFFFFFFu));
    return binomial_rng(n, p, g);
}

std::string plan_frames(const Sources& src, const Isotopes& iso, float tstart_s, float tend_s, uint64_t max_pairs,
                        uint64_t seed, uint64_t first_pair0, std::vector<FramePlan>& out) {
    out.clear();
    const int ns = src.n();
    if (ns < 1) return "no sources";
    if (max_pairs < 1) return "max pairs per frame must be positive";
    std::vector<double> natom(ns), thalf(ns), ratio(ns);
    for (int i = 0; i < ns; i++) {
        thalf[i] = iso.halftime[src.type[i]];
        ratio[i] = iso.ratio[src.type[i]];
        // deterministic decay to the start of the acquisition (gPET.cu:204-208)
        natom[i] = std::floor((double)src.natom[i] * std::exp2(-(double)tstart_s / thalf[i]));
    }
    double t = tstart_s;
    const double tend = tend_s;
    uint64_t first_pair = first_pair0;
    auto expected_pairs = [&](double dt) {
        double e = 0.0;
        for (int i = 0; i < ns; i++) e += natom[i] * (1.0 - std::exp2(-dt / thalf[i])) * ratio[i];
        return e;
    };
    const double target = 0.9 * (double)max_pairs;  // reference aims at 0.95-0.98 of NPART/2 (gPET.cu:449-451)
    for (uint64_t frame = 0; t < tend && frame < (1ull << 24); frame++) {
        double dt = tend - t;
        if (expected_pairs(dt) > target) {
            double lo = 0.0, hi = dt;
            for (int it = 0; it < 200; it++) {
                double mid = 0.5 * (lo + hi);
                if (expected_pairs(mid) > target) hi = mid; else lo = mid;
                if (hi - lo <= 1e-12 * dt) break;
            }
            dt = lo > 0.0 ? lo : hi;
        }
        FramePlan fp;
        fp.t0_s = t;
        fp.dt_s = dt;
        fp.pairs.assign(ns, 0);
        fp.first_pair = first_pair;
        for (int attempt = 0; attempt < 64; attempt++) {
            uint64_t total = 0;
            std::vector<uint64_t> dec(ns);
            for (int i = 0; i < ns; i++) {
                double p = 1.0 - std::exp2(-fp.dt_s / thalf[i]);
                HostRng g(seed, frame * 64 + (uint64_t)attempt, ((uint32_t)kStagePlan << 24) | (uint32_t)i);
                dec[i] = binomial_rng((uint64_t)natom[i], p, g);
                fp.pairs[i] = binomial_rng(dec[i], ratio[i], g);
                total += fp.pairs[i];
            }
            if (total <= max_pairs) {
                for (int i = 0; i < ns; i++) natom[i] -= (double)dec[i];
                fp.npairs = total;
                break;
            }
            fp.dt_s *= 0.5;  // statistically (30 sigma) unreachable; keeps the capacity contract anyway
            if (attempt == 63) return "frame planning failed to fit the capacity";
        }
        first_pair += fp.npairs;
        t += fp.dt_s;
        out.push_back(fp);
        if (fp.dt_s <= 0.0) return "frame planning stalled (activity too high for the capacity)";
    }
    return "";
}

void fill_source_dev(const Sources& src, const Isotopes& iso, const FramePlan& fp, float nonangle, int use_prange,
                     SourceDev& d) {
    memset(&d, 0, sizeof(d));
    const int ns = src.n();
    d.nsource = ns;
    unsigned long long cum = 0;
    for (int i = 0; i < ns && i < 64; i++) {
        cum += fp.pairs[i];
        d.cum_pairs[i] = cum;
        d.type[i] = src.type[i];
        d.shape[i] = src.shape[i];
        for (int j = 0; j < 6; j++) d.coeff[6 * i + j] = src.coeff[6 * i + j];
        // mean life exactly as the reference forms it: double(-halftime * 1.442695) (gPET_kernals.cu:519)
        d.tau_s[i] = (double)iso.halftime[src.type[i]] * 1.442695;
        d.frac[i] = -std::expm1(-fp.dt_s / d.tau_s[i]);
    }
    for (int i = ns; i < 64; i++) d.cum_pairs[i] = cum;   // padding of the binary search in source_pair (transport.cu)
    for (int k = 0; k < iso.n() * 8 && k < 128; k++) d.iso_coef[k] = iso.coef[k];
    d.t0_s = fp.t0_s;
    d.first_pair = fp.first_pair;
    d.nonangle = nonangle;
    d.use_prange = use_prange;
}

}  // namespace gpet
