// Host build of csrc/qt_rng.cuh for tests/test_rng.py (no GPU needed): exposes the same
// draws qt_rng_probe produces on the device.
#include "../../lqr_quadcopter_test_b200/csrc/qt_rng.cuh"

extern "C" void rng_probe_host(const uint64_t *seeds, int64_t E, int n_uniform, int n_normal,
                               double *out) {
    for (int64_t e = 0; e < E; e++) {
        qt::Pcg64 g;
        qt::pcg64_seed(g, seeds[e]);
        for (int i = 0; i < n_uniform; i++) out[(int64_t)i * E + e] = qt::pcg64_double(g);
        for (int i = 0; i < n_normal; i++)
            out[(int64_t)(n_uniform + i) * E + e] = qt::rng_standard_normal(g);
    }
}

// Skip `skip` raw outputs with the jump-ahead, then draw n doubles.
extern "C" void rng_resume_host(uint64_t seed, uint64_t skip, int n, double *out) {
    qt::Pcg64 g;
    qt::pcg64_seed(g, seed);
    qt::pcg64_advance(g, skip);
    for (int i = 0; i < n; i++) out[i] = qt::pcg64_double(g);
}
