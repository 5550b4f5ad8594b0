// qt_trig.cuh — sine/cosine for the rollout's inner loop.
//
// The reference evaluates np.cos/np.sin of roll, pitch and yaw in each of the four RK4 derivative
// evaluations (env/quadcopter_env.py:368-370), i.e. 12 sin/cos pairs per env step: they dominate
// the step.  ncu on the first version of the kernel (profiles/r01_a_*) showed CUDA's sincos()
// costing, per pair, ~27 FP64 instructions plus ~29 UMOVs that re-materialise its polynomial
// coefficients as immediates and ~25 integer/select/branch instructions for quadrant handling and
// the (never taken) huge-argument path.  Two changes, both accurate to ~1 ulp:
//
//  * sincos_pi(): same algorithm as any libm (Cody-Waite reduction by pi/2, fdlibm's minimax
//    kernels on [-pi/4, pi/4]) but with the coefficients in __constant__ memory, so every DFMA takes
//    its coefficient straight from the constant bank.
//  * sincos_shifted(): the stage angles of RK4 are the step's base angle plus a small increment
//    (|delta| <= dt * max_angular_velocity = 0.1 rad by default).  sin/cos(a + delta) follow from
//    sin/cos(a) by the angle-addition formulas with short Taylor polynomials for sin(delta) and
//    cos(delta) - 1: 11-17 FP64 instructions, no integer work.  The increment is taken as
//    fl(a + delta) - a, i.e. the function is evaluated at exactly the rounded stage angle the
//    reference passes to np.sin / np.cos.
//
// Measured against 40-digit arithmetic (Numerics): kernels <= 1.7e-16 relative,
// shifted evaluation <= 1.2e-16 absolute for |delta| <= 0.26.
#pragma once

// Also compiles as plain host C++ so tests/test_trig.py can measure the accuracy without a GPU.
#if defined(__CUDACC__)
#define QT_TRIG_FN __device__ __forceinline__
#define QT_TRIG_CONST __constant__
#define QT_LOINT(x) __double2loint(x)
#else
#include <cmath>
#include <cstring>
#define QT_TRIG_FN static inline
#define QT_TRIG_CONST static const
static inline int qt_loint_host(double x) {
    long long u;
    memcpy(&u, &x, sizeof u);
    return (int)(unsigned)(u & 0xffffffffLL);
}
#define QT_LOINT(x) qt_loint_host(x)
#endif

namespace qt {

// fdlibm k_sin / k_cos minimax coefficients, Cody-Waite split of pi/2, Taylor coefficients
QT_TRIG_CONST double c_trig[32] = {
    /* 0  S1..S6 */ -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    /* 6  C1..C6 */ 4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11,
    /* 12 2/pi, pi/2 hi, mid, lo */ 6.36619772367581382433e-01, 1.57079632679489655800e+00,
    6.12323399573676603587e-17, -1.49738490485916983000e-33,
    /* 16 sin(d)/d - 1 = z*P(z): -1/3!, 1/5!, -1/7!, 1/9!, -1/11! */ -1.66666666666666666667e-01,
    8.33333333333333333333e-03, -1.98412698412698412698e-04, 2.75573192239858906526e-06,
    -2.50521083854417187751e-08,
    /* 21 cos(d) - 1 = z*Q(z): -1/2!, 1/4!, -1/6!, 1/8!, -1/10!, 1/12! */ -0.5, 4.16666666666666666667e-02,
    -1.38888888888888888889e-03, 2.48015873015873015873e-05, -2.75573192239858906526e-07,
    2.08767569878680989792e-09,
    0.0, 0.0, 0.0, 0.0, 0.0};

constexpr double kShiftMax = 0.26;        // full Taylor polynomials: truncation error < 1e-17 up to here
constexpr double kShiftSmall = 0.0078125;  // 2^-7: sin through d^5, cos through d^6 suffice (< 4e-19)
constexpr double kShiftTiny = 0.0009765625;  // 2^-10: sin through d^3, cos through d^4 suffice (< 8e-18)
constexpr double kShiftMid = 0.03125;        // 2^-5: sin through d^7, cos through d^8 suffice (< 8e-20)

// Out-of-line so the (practically never taken) huge-argument path costs the hot loop one
// predicated call instead of an inlined Payne-Hanek reduction.
#if defined(__CUDACC__)
static __device__ __noinline__ void sincos_library(double x, double *s, double *c) { sincos(x, s, c); }
#else
static void sincos_library(double x, double *s, double *c) {
    *s = std::sin(x);
    *c = std::cos(x);
}
#endif

// sin and cos of x for |x| <= 1e5 (3-term Cody-Waite); larger arguments use the library routine.
QT_TRIG_FN void sincos_pi(double x, double &s, double &c) {
    if (!(fabs(x) <= 1.0e5)) {  // also NaN / inf
        double ls, lc;
        sincos_library(x, &ls, &lc);
        s = ls;
        c = lc;
        return;
    }
    const double magic = 6755399441055744.0;  // 1.5 * 2^52: rounds to nearest integer in the low word
    const double t = fma(x, c_trig[12], magic);
    const int k = QT_LOINT(t);
    const double q = t - magic;
    double r = fma(-q, c_trig[13], x);
    r = fma(-q, c_trig[14], r);
    r = fma(-q, c_trig[15], r);
    const double z = r * r;
    double ps = fma(z, c_trig[5], c_trig[4]);
    double pc = fma(z, c_trig[11], c_trig[10]);
    ps = fma(z, ps, c_trig[3]);
    pc = fma(z, pc, c_trig[9]);
    ps = fma(z, ps, c_trig[2]);
    pc = fma(z, pc, c_trig[8]);
    ps = fma(z, ps, c_trig[1]);
    pc = fma(z, pc, c_trig[7]);
    ps = fma(z, ps, c_trig[0]);
    pc = fma(z, pc, c_trig[6]);
    const double sr = fma(z * r, ps, r);            // r + r^3 * S(z)
    const double cr = fma(z, fma(z, pc, -0.5), 1.0);  // 1 - z/2 + z^2 * C(z)
    const double a = (k & 1) ? cr : sr;
    const double b = (k & 1) ? sr : cr;
    s = (k & 2) ? -a : a;
    c = ((k + 1) & 2) ? -b : b;
}

// sin/cos of `arg`, given sin/cos of `base`, where arg = fl(base + small increment) and the
// caller guarantees |arg - base| <= kShiftMax (SMALL = false) or <= kShiftSmall (SMALL = true:
// two Taylor terms fewer per polynomial, truncation error < 4e-19).
// TIER 0: |d| <= kShiftMax, 1: |d| <= kShiftSmall, 2: |d| <= kShiftTiny, 3: |d| <= kShiftMid
template <int TIER>
QT_TRIG_FN void sincos_shifted_d(double sb, double cb, double d, double &s, double &c);

template <int TIER>
QT_TRIG_FN void sincos_shifted(double base, double sb, double cb, double arg, double &s, double &c) {
    sincos_shifted_d<TIER>(sb, cb, arg - base, s, c);  // exact (Sterbenz) or the increment itself when base is tiny
}

// The same with the increment given directly: sin/cos(b + d) from sin/cos(b).
template <int TIER>
QT_TRIG_FN void sincos_shifted_d(double sb, double cb, double d, double &s, double &c) {
    const double z = d * d;
    if (TIER == 2) {
        const double sd = fma(d * z, c_trig[16], d);              // d - d^3/6
        const double cm = z * fma(z, c_trig[22], c_trig[21]);     // -z/2 + z^2/24
        s = fma(cb, sd, fma(sb, cm, sb));
        c = fma(-sb, sd, fma(cb, cm, cb));
        return;
    }
    double p, q;
    if (TIER == 1) {
        p = fma(z, c_trig[17], c_trig[16]);
        q = fma(z, c_trig[23], c_trig[22]);
    } else if (TIER == 3) {
        p = fma(z, c_trig[18], c_trig[17]);
        q = fma(z, c_trig[24], c_trig[23]);
        p = fma(z, p, c_trig[16]);
        q = fma(z, q, c_trig[22]);
    } else {
        p = fma(z, c_trig[20], c_trig[19]);
        q = fma(z, c_trig[26], c_trig[25]);
        p = fma(z, p, c_trig[18]);
        q = fma(z, q, c_trig[24]);
        p = fma(z, p, c_trig[17]);
        q = fma(z, q, c_trig[23]);
        p = fma(z, p, c_trig[16]);
        q = fma(z, q, c_trig[22]);
    }
    q = fma(z, q, c_trig[21]);
    const double sd = fma(d * z, p, d);  // sin(d)
    const double cm = z * q;             // cos(d) - 1
    s = fma(cb, sd, fma(sb, cm, sb));    // sin(b) cos(d) + cos(b) sin(d)
    c = fma(-sb, sd, fma(cb, cm, cb));   // cos(b) cos(d) - sin(b) sin(d)
}

}  // namespace qt
