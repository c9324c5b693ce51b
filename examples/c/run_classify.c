/*
 * run_classify.c -- a whole acquisition through libgpet_b200.so from plain C (C99): the calls a gPET maintainer puts in
 * place of main.cu:192-229 (iniDevice ... sampleParticle), plus the coincidence sorter and its true / scatter / random
 * classes, which the reference does not have.
 *
 *   gcc -std=c99 -Iinclude examples/c/run_classify.c -Lgpet_b200 -lgpet_b200 -Wl,-rpath,$PWD/gpet_b200 -o run_classify
 *   ./run_classify input_PET.in window_us [device]          (run from the directory input_PET.in's paths are relative to)
 *
 * Prints the counters the reference prints (gPET.cu:293,364,382,398,415,423) and the class totals, then the first
 * coincidences with their classes.  device = -1: host-only context, the run fails with GPET_ERR_NO_DEVICE (exit code 3).
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
    if (argc < 3) {
        fprintf(stderr, "usage: %s input_PET.in coincidence_window_us [device]\n", argv[0]);
        return 2;
    }
    const int device = argc > 3 ? atoi(argv[3]) : 0;
    if (gpet_abi_version() != GPET_ABI_VERSION) {
        fprintf(stderr, "library ABI %d, header ABI %d\n", gpet_abi_version(), GPET_ABI_VERSION);
        return 1;
    }
    gpet_ctx* ctx = NULL;
    int rc = gpet_create(device, &ctx);
    if (rc != GPET_OK) return die(NULL, "gpet_create", rc);
    if ((rc = gpet_load_config_file(ctx, argv[1], NULL, NULL)) != GPET_OK) return die(ctx, "gpet_load_config_file", rc);

    gpet_digitizer_params dg;
    gpet_get_digitizer(ctx, &dg);
    dg.coinc_window_us = (float)atof(argv[2]);
    dg.coinc_policy = 0;         /* windows with exactly two singles */
    dg.coinc_pair_shift = 0;     /* isotope source: both photons carry the pair's eventid */
    if ((rc = gpet_set_digitizer(ctx, &dg)) != GPET_OK) return die(ctx, "gpet_set_digitizer", rc);

    /* coincidences as index pairs and singles as 32-byte records on their way to the host (a run like this one is bound by that
     * copy); gpet_result_singles / gpet_result_coincidences rebuild the reference's 48-byte Event on demand, byte for byte */
    if ((rc = gpet_set_coincidence_format(ctx, GPET_COINC_PAIRS)) != GPET_OK) return die(ctx, "gpet_set_coincidence_format", rc);
    if ((rc = gpet_set_singles_format(ctx, GPET_SINGLES_COMPACT)) != GPET_OK) return die(ctx, "gpet_set_singles_format", rc);

    gpet_stats st;
    if ((rc = gpet_run(ctx, NULL, &st)) != GPET_OK) return die(ctx, "gpet_run", rc);
    printf("pairs %llu, hits %llu, events after adder %llu, thresholder %llu, deadtime %llu, singles %llu\n",
           (unsigned long long)st.pairs, (unsigned long long)st.hits, (unsigned long long)st.events_adder,
           (unsigned long long)st.events_threshold, (unsigned long long)st.events_deadtime, (unsigned long long)st.singles);
    printf("coincidences %llu: trues %llu, scatters %llu, randoms %llu (%.3f ms on the device, %llu frames)\n",
           (unsigned long long)st.coincidences, (unsigned long long)st.trues, (unsigned long long)st.scatters,
           (unsigned long long)st.randoms, st.ms_total, (unsigned long long)st.frames);

    const gpet_single_compact* cs = NULL;
    const gpet_event* sg = NULL;
    const int64_t ncs = gpet_result_singles_compact(ctx, &cs), nsg = gpet_result_singles(ctx, &sg);
    if (ncs != nsg || (uint64_t)nsg != st.singles) return die(ctx, "singles bookkeeping", GPET_ERR_ARG);
    if (nsg > 0) {   /* the same expansion as a call of its own, for compact records a caller kept */
        gpet_event first;
        if ((rc = gpet_expand_singles(ctx, cs, 1, &first)) != GPET_OK) return die(ctx, "gpet_expand_singles", rc);
        printf("first single: t = %.6f us, E = %.0f eV, panel %d module %d crystal %d (photon %d of annihilation %d)\n", first.t,
               (double)first.E, (int)first.pann, (int)first.modn, (int)first.cryn, (int)first.parn, (int)first.eventid);
        if (first.t != sg[0].t || first.parn != sg[0].parn || first.siten != sg[0].siten) return die(ctx, "expansion mismatch", GPET_ERR_ARG);
    }

    const gpet_coincidence* co = NULL;
    const uint8_t* cls = NULL;
    const int64_t nco = gpet_result_coincidences(ctx, &co), ncl = gpet_result_coincidence_classes(ctx, &cls);
    static const char* const name[3] = {"true", "scatter", "random"};
    for (int64_t k = 0; k < nco && k < ncl && k < 5; k++)
        printf("  t = %.6f us, panels %d / %d, E = %.0f / %.0f eV: %s\n", co[k].a.t, (int)co[k].a.pann, (int)co[k].b.pann,
               (double)co[k].a.E, (double)co[k].b.E, cls[k] < 3 ? name[cls[k]] : "?");
    gpet_destroy(ctx);
    return 0;
}
