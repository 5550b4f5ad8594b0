// Host build of csrc/qt_trig.cuh for tests/test_trig.py.
#include "../../lqr_quadcopter_test_b200/csrc/qt_trig.cuh"

extern "C" void trig_full_host(const double *x, long n, double *s, double *c) {
    for (long i = 0; i < n; i++) qt::sincos_pi(x[i], s[i], c[i]);
}

extern "C" void trig_shifted_host(const double *base, const double *arg, long n, int tier, double *s, double *c) {
    for (long i = 0; i < n; i++) {
        double sb, cb;
        qt::sincos_pi(base[i], sb, cb);
        if (tier == 3)
            qt::sincos_shifted<3>(base[i], sb, cb, arg[i], s[i], c[i]);
        else if (tier == 2)
            qt::sincos_shifted<2>(base[i], sb, cb, arg[i], s[i], c[i]);
        else if (tier == 1)
            qt::sincos_shifted<1>(base[i], sb, cb, arg[i], s[i], c[i]);
        else
            qt::sincos_shifted<0>(base[i], sb, cb, arg[i], s[i], c[i]);
    }
}
