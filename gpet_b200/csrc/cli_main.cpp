// gPET-compatible command line: `gpet_b200 input_PET.in [--data DIR] [--out DIR] [--seed N] [--device N] [--coinc-window US
// [--coinc-policy 0|1] [--min-panel-diff N] [--pair-shift 0|1]]`.
// Mirrors main() of the reference (main.cu:26-276): one positional input file, paths relative to the working
// directory, outputs appended under ./output/ with the layouts output/readOutput.m reads.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "gpet_b200.h"

int main(int argc, char** argv) {
    if (argc < 2) {
        printf("Please execute ./gpet_b200 input_file\nThanks.\n\n");  // main.cu:28-33
        return 1;
    }
    std::string input = argv[1], data, out = "output";
    unsigned long long seed = 0x67504554ull;
    float cwin = 0.f;
    int device = -1, cpolicy = -1, cmindiff = -1, pair_shift = -1;   // device: --device, else the GPU index line of the input file
    for (int i = 2; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--data")) data = argv[i + 1];
        else if (!strcmp(argv[i], "--out")) out = argv[i + 1];
        else if (!strcmp(argv[i], "--seed")) seed = strtoull(argv[i + 1], nullptr, 0);
        else if (!strcmp(argv[i], "--coinc-window")) cwin = (float)atof(argv[i + 1]);
        else if (!strcmp(argv[i], "--device")) device = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--coinc-policy")) cpolicy = atoi(argv[i + 1]);      // 0 drop multiples, 1 all pairs with the opener
        else if (!strcmp(argv[i], "--min-panel-diff")) cmindiff = atoi(argv[i + 1]);   // cyclic panel distance a pair needs
        else if (!strcmp(argv[i], "--pair-shift")) pair_shift = atoi(argv[i + 1]);     // 1: photon PSF with records 2k, 2k+1 = one pair
        else { fprintf(stderr, "unknown option %s\n", argv[i]); return 1; }
    }
    auto t0 = std::chrono::steady_clock::now();
    gpet_ctx* ctx = nullptr;
    if (device < 0) {
        // "GPU index" of input_PET.in, as the reference's main() uses it (main.cu:52-56, iniDevice)
        device = gpet_peek_config_device(input.c_str());
        if (device < 0) { fprintf(stderr, "gpet_b200: cannot read the GPU index from %s\n", input.c_str()); return 1; }
    }
    if (gpet_create(device, &ctx) != GPET_OK) return 1;
    gpet_set_seed(ctx, seed);
    int r = gpet_load_config_file(ctx, input.c_str(), nullptr, data.empty() ? nullptr : data.c_str());
    if (r != GPET_OK) { fprintf(stderr, "gpet_b200: %s\n", gpet_last_error(ctx)); gpet_destroy(ctx); return 1; }
    if (cwin > 0.f) {
        gpet_digitizer_params d;
        gpet_get_digitizer(ctx, &d);
        d.coinc_window_us = cwin;
        if (cpolicy >= 0) d.coinc_policy = cpolicy;
        if (cmindiff >= 0) d.coinc_min_panel_diff = cmindiff;
        if (pair_shift >= 0) d.coinc_pair_shift = pair_shift;
        gpet_set_digitizer(ctx, &d);
    }
    auto t1 = std::chrono::steady_clock::now();
    printf("Initialize time: %f s.\n", std::chrono::duration<double>(t1 - t0).count());
    gpet_stats st;
    r = gpet_run(ctx, out.c_str(), &st);
    if (r != GPET_OK) { fprintf(stderr, "gpet_b200: %s\n", gpet_last_error(ctx)); gpet_destroy(ctx); return 1; }
    auto t2 = std::chrono::steady_clock::now();
    printf("emitted pairs %llu\nthere are %llu Hits\ncounts of events after adder is %llu\n"
           "counts of events after thresholder is %llu\ncounts of events after deadtime is %llu\ncounts of singles is %llu\n",
           (unsigned long long)st.pairs, (unsigned long long)st.hits, (unsigned long long)st.events_adder,
           (unsigned long long)st.events_threshold, (unsigned long long)st.events_deadtime, (unsigned long long)st.singles);
    if (cwin > 0.f)
        printf("counts of coincidences is %llu (trues %llu, scatters %llu, randoms %llu)\n", (unsigned long long)st.coincidences,
               (unsigned long long)st.trues, (unsigned long long)st.scatters, (unsigned long long)st.randoms);
    printf("Simulation time: %f s. (device %f ms, %llu frames, %llu kernel launches)\n",
           std::chrono::duration<double>(t2 - t1).count(), st.ms_total, (unsigned long long)st.frames,
           (unsigned long long)st.kernel_launches);
    gpet_destroy(ctx);
    printf("Total time: %f s.\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    return 0;
}
