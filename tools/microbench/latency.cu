// Calibration micro-benchmarks for the latency-bound digitizer kernels (B200): launch floor, dependent-load latency,
// atomic round trip, decoupled look-back with volatile (.STRONG.SYS) vs relaxed.gpu accesses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/latency tools/microbench/latency.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_empty() {}

__global__ void k_chase(const unsigned* __restrict__ next, unsigned* out, int steps) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    for (int s = 0; s < steps; s++) i = next[i];
    if (i == 0xffffffffu) out[0] = i;
}

__global__ void k_atomic(unsigned* ctr, unsigned* out, int steps) {
    unsigned v = 0;
    for (int s = 0; s < steps; s++) v += atomicAdd(ctr + ((threadIdx.x + v) & 31) * 32, 1u) & 1u;
    if (v == 0xffffffffu) out[0] = v;
}

// what the transport kernels do: ONE lane per warp adds to a shared counter and the warp waits for the old value.
// words = number of distinct counters the warps spread over (stride words apart); 1 = every warp on the same address
__global__ void __launch_bounds__(256) k_atomic_warp(unsigned* ctr, unsigned* out, int steps, int words, int stride) {
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned v = 0;
    for (int s = 0; s < steps; s++) {
        unsigned r = 0;
        if ((threadIdx.x & 31) == 0) r = atomicAdd(ctr + ((warp + v) % words) * stride, 3u);
        v += __shfl_sync(~0u, r, 0) & 1u;
    }
    if (v == 0xffffffffu) out[0] = v;
}

template <int MODE> __device__ __forceinline__ unsigned ld(const unsigned* p) {
    if (MODE == 0) return *reinterpret_cast<const volatile unsigned*>(p);
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
template <int MODE> __device__ __forceinline__ void st(unsigned* p, unsigned v) {
    if (MODE == 0) { *reinterpret_cast<volatile unsigned*>(p) = v; return; }
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

constexpr unsigned kAgg = 1u << 30, kPre = 2u << 30, kMask = kAgg - 1;

// WARP = 0: thread 0 reads a window of 32 predecessors; WARP = 1: warp 0 reads one predecessor per lane
template <int MODE, int WARP>
__global__ void __launch_bounds__(256) k_scan(const unsigned* __restrict__ in, unsigned* __restrict__ out, unsigned* status, unsigned n) {
    __shared__ unsigned ws[8];
    __shared__ unsigned s_excl;
    const unsigned tile = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned b0 = tile * 2048 + threadIdx.x * 8;
    unsigned c[8], s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { c[k] = b0 + k < n ? in[b0 + k] : 0; s += c[k]; }
    unsigned x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(~0u, x, o); if (lane >= o) x += y; }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    unsigned wp = 0, total = 0;
    for (unsigned w = 0; w < 8; w++) { unsigned cc = ws[w]; if (w < warp) wp += cc; total += cc; }
    if (WARP == 0) {
        if (threadIdx.x == 0) {
            unsigned excl = 0;
            if (tile == 0) st<MODE>(&status[0], kPre | total);
            else {
                st<MODE>(&status[tile], kAgg | total);
                int t = (int)tile - 1;
                bool done = false;
                while (t >= 0 && !done) {
                    unsigned v[32];
#pragma unroll
                    for (int k = 0; k < 32; k++) v[k] = (t - k >= 0) ? ld<MODE>(status + (t - k)) : kPre;
                    int k = 0;
#pragma unroll
                    for (; k < 32; k++) {
                        if ((v[k] >> 30) == 0u) break;
                        excl += v[k] & kMask;
                        if (v[k] & kPre) { done = true; break; }
                    }
                    t -= k;
                }
                st<MODE>(&status[tile], kPre | (excl + total));
            }
            s_excl = excl;
        }
    } else {
        if (warp == 0) {
            unsigned excl = 0;
            if (tile == 0) { if (lane == 0) st<MODE>(&status[0], kPre | total); }
            else {
                if (lane == 0) st<MODE>(&status[tile], kAgg | total);
                int t = (int)tile - 1;
                while (true) {
                    unsigned v = (t - (int)lane >= 0) ? ld<MODE>(status + (t - (int)lane)) : kPre;
                    unsigned unpub = __ballot_sync(~0u, (v >> 30) == 0u);
                    unsigned pre = __ballot_sync(~0u, (v & kPre) != 0u);
                    unsigned stop = unpub | pre;                     // first lane that ends the usable run
                    int first = stop ? __ffs(stop) - 1 : 32;
                    bool is_pre = stop && ((pre >> first) & 1u);
                    unsigned take = (lane < (unsigned)first || (is_pre && lane == (unsigned)first)) ? (v & kMask) : 0u;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) take += __shfl_xor_sync(~0u, take, o);
                    excl += take;
                    if (is_pre) break;
                    t -= first;
                }
                if (lane == 0) st<MODE>(&status[tile], kPre | (excl + total));
            }
            if (lane == 0) s_excl = excl;
        }
    }
    __syncthreads();
    unsigned e = s_excl + wp + (x - s);
#pragma unroll
    for (int k = 0; k < 8; k++) { if (b0 + k < n) out[b0 + k] = e; e += c[k]; }
}

template <typename F> float time_us(F f, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; i++) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms * 1e3f / reps;
}

int main() {
    printf("empty kernel, back to back: %.2f us/launch\n", time_us([] { k_empty<<<1, 32>>>(); }, 2000));
    printf("empty kernel 592x256:       %.2f us/launch\n", time_us([] { k_empty<<<592, 256>>>(); }, 2000));
    {   // pointer chase over 64 MB (L2 resident on the second pass) and 1 GB (HBM)
        for (size_t words : {size_t(1) << 24, size_t(1) << 28}) {
            std::vector<unsigned> h(words);
            unsigned long long s = 12345;
            for (size_t i = 0; i < words; i++) { s = s * 6364136223846793005ull + 1442695040888963407ull; h[i] = (unsigned)((s >> 33) % words); }
            unsigned *d, *o;
            CK(cudaMalloc(&d, words * 4)); CK(cudaMalloc(&o, 4));
            CK(cudaMemcpy(d, h.data(), words * 4, cudaMemcpyHostToDevice));
            for (int steps : {1, 9, 33}) {
                float t1 = time_us([&] { k_chase<<<148, 256>>>(d, o, steps); }, 50);
                printf("chase %4zu MB, %2d dependent loads, 148x256 threads: %.2f us\n", words * 4 >> 20, steps, t1);
            }
            cudaFree(d); cudaFree(o);
        }
    }
    {
        unsigned *c, *o;
        CK(cudaMalloc(&c, 32 * 1024 * 4 + 4096)); CK(cudaMalloc(&o, 4));
        for (int words : {1, 2, 4, 32})
            for (int stride : {1, 32, 1024}) {
                if (words == 1 && stride > 1) continue;
                const int steps = 40;
                float us = time_us([&] { k_atomic_warp<<<592, 256>>>(c, o, steps, words, stride); }, 20);
                printf("warp-aggregated atomicAdd with return: 4736 warps x %d dependent, %2d counters %4d words apart: %8.2f us = %.2f ns per atomic\n",
                       steps, words, stride, us, us * 1e3 / (4736.0 * steps));
            }
        for (int steps : {1, 9, 33}) printf("atomicAdd with return, %2d dependent, 148x256 threads on 32 words: %.2f us\n", steps, time_us([&] { k_atomic<<<148, 256>>>(c, o, steps); }, 50));
    }
    for (unsigned n : {131072u, 702464u, 2097152u}) {
        unsigned *in, *out, *st;
        CK(cudaMalloc(&in, n * 4)); CK(cudaMalloc(&out, n * 4)); CK(cudaMalloc(&st, 4096 * 4));
        CK(cudaMemset(in, 1, n * 4));
        const unsigned tiles = (n + 2047) / 2048;
        auto run = [&](int which) {
            cudaMemsetAsync(st, 0, tiles * 4);
            if (which == 0) k_scan<0, 0><<<tiles, 256>>>(in, out, st, n);
            if (which == 1) k_scan<1, 0><<<tiles, 256>>>(in, out, st, n);
            if (which == 2) k_scan<0, 1><<<tiles, 256>>>(in, out, st, n);
            if (which == 3) k_scan<1, 1><<<tiles, 256>>>(in, out, st, n);
        };
        float base = time_us([&] { cudaMemsetAsync(st, 0, tiles * 4); }, 200);
        const char* names[4] = {"volatile, thread 0 window", "relaxed.gpu, thread 0 window", "volatile, warp window", "relaxed.gpu, warp window"};
        for (int w = 0; w < 4; w++) printf("scan n=%7u (%4u tiles) %-30s %.2f us (memset alone %.2f)\n", n, tiles, names[w], time_us([&] { run(w); }, 200), base);
        cudaFree(in); cudaFree(out); cudaFree(st);
    }
    return 0;
}
