/*
 * digitize_replay.c -- the digitizer entry point of libgpet_b200.so from plain C (C99).
 *
 * What a gPET maintainer would write to swap only the digitizer chain (gPET.cu:385-424: blur, energywindow,
 * quicksort_h, setSitenum, orderevents, deadtime, ...): read the reference's own post-adder dump output/adder.dat
 * (48-byte Event records, gPET.cu:383), run gpet_digitize, append output/singles.dat in the layout
 * output/readOutput.m:18-34 reads.  With blur off this reproduces the reference's singles.dat byte for byte
 * (tests/test_reference_pin.py).
 *
 *   gcc -std=c99 -Iinclude examples/c/digitize_replay.c -Lgpet_b200 -lgpet_b200 -Wl,-rpath,$PWD/gpet_b200 -o digitize_replay
 *   ./digitize_replay input/config8.geo output/adder.dat output/singles.dat [device]
 *
 * device = -1 creates a host-only context: the call then fails with GPET_ERR_NO_DEVICE (there is no CPU fallback).
 */
#include <stdio.h>
#include <stdlib.h>

#include "gpet_b200.h"

static int die(gpet_ctx* ctx, const char* what, int rc) {
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, ctx ? gpet_last_error(ctx) : "");
    if (ctx) gpet_destroy(ctx);
    return rc == GPET_ERR_NO_DEVICE ? 3 : 1;
}

int main(int argc, char** argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s geometry.geo adder.dat singles.dat [device]\n", argv[0]);
        return 2;
    }
    const int device = argc > 4 ? atoi(argv[4]) : 0;
    FILE* f = fopen(argv[2], "rb");
    if (!f) { perror(argv[2]); return 1; }
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    const int64_t n = (int64_t)(bytes / (long)sizeof(gpet_event));
    gpet_event* in = (gpet_event*)malloc((size_t)(n > 0 ? n : 1) * sizeof(gpet_event));
    gpet_event* out = (gpet_event*)malloc((size_t)(n > 0 ? n : 1) * sizeof(gpet_event));
    if (!in || !out || fread(in, sizeof(gpet_event), (size_t)n, f) != (size_t)n) { fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
    fclose(f);

    gpet_ctx* ctx = NULL;
    int rc = gpet_create(device, &ctx);                                  /* iniDevice (iniDevice.cu:42-58) */
    if (rc) return die(ctx, "gpet_create", rc);
    if ((rc = gpet_load_geometry(ctx, argv[1]))) return die(ctx, "gpet_load_geometry", rc);   /* module / crystal counts for setSitenum */

    gpet_digitizer_params dg;
    if ((rc = gpet_get_digitizer(ctx, &dg))) return die(ctx, "gpet_get_digitizer", rc);
    dg.readout_depth = 2; dg.readout_policy = 1;                          /* input_PET.in fields 18-22 of the shipped example */
    dg.threshold_eV = 50000.f;
    dg.blur_policy = 1; dg.blur_Eref = 662000.f; dg.blur_Rref = 0.f; dg.blur_slope = 0.f; dg.blur_space = 0.f;   /* blur off: deterministic */
    dg.dead_level = 3; dg.dead_type = 0; dg.dead_time_us = 2.2f;
    dg.ewin_min = 350000.f; dg.ewin_max = 650000.f;
    if ((rc = gpet_set_digitizer(ctx, &dg))) return die(ctx, "gpet_set_digitizer", rc);

    int64_t nsingles = 0;
    uint64_t counts[4] = {0, 0, 0, 0};
    if ((rc = gpet_digitize(ctx, in, n, out, n, &nsingles, counts))) return die(ctx, "gpet_digitize", rc);
    printf("events %lld -> after thresholder %llu, after dead time %llu, singles %lld\n", (long long)n, (unsigned long long)counts[1],
           (unsigned long long)counts[2], (long long)nsingles);

    f = fopen(argv[3], "ab");                                              /* the reference appends (detector.cu:296) */
    if (!f) { perror(argv[3]); gpet_destroy(ctx); return 1; }
    fwrite(out, sizeof(gpet_event), (size_t)nsingles, f);
    fclose(f);
    gpet_destroy(ctx);
    free(in); free(out);
    return 0;
}
