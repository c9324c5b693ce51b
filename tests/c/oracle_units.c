/* TEST INFRASTRUCTURE.  Reaches the file-local functions of the CPU oracle (oracle/gpet_oracle.c) so that
 * tests/test_oracle_units.py can check them one by one -- against physics (Klein-Nishina, rotation geometry) and against
 * plain-Python restatements of the reference lines (crystalSearch, adder, readout).  Built by the test with gcc; never
 * linked into the product. */
#include "../../oracle/gpet_oracle.c"

void u_rotate(float* d, int64_t n, const float* costh, const float* phi) {
    for (int64_t i = 0; i < n; i++) rotate_dir(d + 3 * i, d + 3 * i + 1, d + 3 * i + 2, costh[i], phi[i]);
}

void u_compton_kn(float E, uint64_t seed, int64_t n, float* efrac, float* costh) {
    for (int64_t i = 0; i < n; i++) {
        orc_rng g;
        rng_init(&g, seed, (uint64_t)i, 0u);
        compton_kn(E, &g, efrac + i, costh + i);
    }
}

void u_crystal_search(const orc_panel* pd, int moduleNy, int crystalNy, int nsurface, const float* surface, int64_t n,
                      const float* xyz, int32_t* ids) {
    for (int64_t i = 0; i < n; i++) {
        int m, M, L;
        crystal_search(pd, moduleNy, crystalNy, nsurface, surface, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], &m, &M, &L);
        ids[3 * i] = m; ids[3 * i + 1] = M; ids[3 * i + 2] = L;
    }
}

/* one photon: its hits (as events with crystal-level siten) through adder, then readout; returns the number of events */
int u_adder_readout(const orc_event* hits, int n, int depth, int policy, int moduleN, orc_event* out, int* overflow) {
    orc_event evs[ORC_MAXEV];
    int cnt = 0;
    *overflow = 0;
    for (int i = 0; i < n; i++)
        if (!adder(evs, &cnt, hits + i)) (*overflow)++;
    int nout = cnt ? readout(evs, cnt, depth, policy, moduleN) : 0;
    for (int k = 0; k < nout; k++) out[k] = evs[k];
    return nout;
}

void u_sample_ek_positron(const float* coef8, uint64_t seed, int64_t n, float* ek_eV) {
    for (int64_t i = 0; i < n; i++) {
        orc_rng g;
        rng_init(&g, seed, (uint64_t)i, 0u);
        ek_eV[i] = sample_ek_positron(coef8, &g);
    }
}

void u_surface_lookup(const float* surf, int mat, int ncp, int ne, int64_t n, const float* xe, const float* xcp, float* out) {
    for (int64_t i = 0; i < n; i++) out[i] = surface_lookup(surf, mat, ncp, ne, xe[i], xcp[i]);
}
