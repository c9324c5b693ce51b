// Optional per-kernel CUDA-event timing (gpet_profile_enable): every launch of this library's kernels is bracketed by
// two events on the launching stream; totals are kept per kernel name.  Off by default -- the bracketing serialises
// nothing but adds two event records per launch, so timed benchmark regions run with it disabled.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

namespace gpet {

struct KernelTimer {
    struct Rec { const char* name; cudaEvent_t a, b; };
    struct Acc { double ms = 0.0; unsigned long long launches = 0; };
    std::vector<Rec> pending;
    std::vector<cudaEvent_t> pool;
    std::map<std::string, Acc> acc;
    std::vector<std::string> order;  // first-seen order of the names

    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    void begin(const char* name, cudaStream_t s) {
        Rec r{name, get(), get()};
        cudaEventRecord(r.a, s);
        pending.push_back(r);
    }
    void end(cudaStream_t s) { cudaEventRecord(pending.back().b, s); }
    // call after the stream has been synchronised
    void collect() {
        for (auto& r : pending) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
                auto it = acc.find(r.name);
                if (it == acc.end()) { order.push_back(r.name); it = acc.emplace(r.name, Acc{}).first; }
                it->second.ms += ms;
                it->second.launches++;
            }
            pool.push_back(r.a); pool.push_back(r.b);
        }
        pending.clear();
    }
    void reset() { collect(); acc.clear(); order.clear(); }
    ~KernelTimer() { for (auto e : pool) cudaEventDestroy(e); for (auto& r : pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); } }
};

extern thread_local KernelTimer* g_ktimer;  // set by the ABI layer around launcher calls when profiling is on

}  // namespace gpet

#define GPET_LAUNCH(name, stream, ...)                                   \
    do {                                                                 \
        if (::gpet::g_ktimer) ::gpet::g_ktimer->begin(name, stream);     \
        __VA_ARGS__;                                                     \
        if (::gpet::g_ktimer) ::gpet::g_ktimer->end(stream);             \
    } while (0)
