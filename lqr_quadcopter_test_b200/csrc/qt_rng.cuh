// qt_rng.cuh — numpy's default_rng(seed) on the device.
//
// The reference seeds every episode with np.random.default_rng(seed)
// (env/quadcopter_env.py:123, env/target_motion.py:299) and draws uniform() and
// standard_normal() from it (quadcopter_env.py:137, target_motion.py:320, 329, 353).  To start
// a million episodes from exactly the reference's initial conditions without a 50 us/episode
// host loop, this header restates the published algorithms numpy uses:
//   SeedSequence (O'Neill's seed_seq_fe128 variant, numpy/random/bit_generator.pyx)
//   PCG64 = PCG XSL-RR 128/64 with the default 128-bit multiplier (numpy/random/src/pcg64)
//   Generator.random(): (u64 >> 11) * 2^-53
//   Generator.standard_normal(): 256-layer ziggurat (numpy/random/src/distributions), tables in
//   qt_ziggurat_tables.h
// Pinned against numpy itself in tests/test_rng.py.
#pragma once
#include <stdint.h>

#include "qt_ziggurat_tables.h"

// The header also compiles as plain host C++ (g++), which is how tests/test_rng.py pins the
// algorithm against numpy on a machine without a GPU.
#if defined(__CUDACC__)
#define QT_RNG_FN __host__ __device__ __forceinline__
#include <cstring>
#else
#define QT_RNG_FN static inline
#include <cmath>
#include <cstring>
#endif

namespace qt {

#if defined(__CUDACC__)
__constant__ uint64_t c_zig_wi[256] = {QT_ZIG_WI_VALUES};
__constant__ uint64_t c_zig_ki[256] = {QT_ZIG_KI_VALUES};
__constant__ uint64_t c_zig_fi[256] = {QT_ZIG_FI_VALUES};
#endif
#if !defined(__CUDA_ARCH__)
static const uint64_t h_zig_wi[256] = {QT_ZIG_WI_VALUES};
static const uint64_t h_zig_ki[256] = {QT_ZIG_KI_VALUES};
static const uint64_t h_zig_fi[256] = {QT_ZIG_FI_VALUES};
#endif

#if defined(__CUDA_ARCH__)
#define QT_ZIG_WI(i) __longlong_as_double((long long)c_zig_wi[i])
#define QT_ZIG_KI(i) c_zig_ki[i]
#define QT_ZIG_FI(i) __longlong_as_double((long long)c_zig_fi[i])
// numpy's C code is not FMA-contracted: keep products and sums separately rounded
#define QT_MUL_RN(a, b) __dmul_rn((a), (b))
#define QT_ADD_RN(a, b) __dadd_rn((a), (b))
#else
QT_RNG_FN double bits_to_double(uint64_t u) {
    double d;
    memcpy(&d, &u, sizeof d);
    return d;
}
#define QT_ZIG_WI(i) bits_to_double(h_zig_wi[i])
#define QT_ZIG_KI(i) h_zig_ki[i]
#define QT_ZIG_FI(i) bits_to_double(h_zig_fi[i])
#define QT_MUL_RN(a, b) ((a) * (b)) /* host build uses -ffp-contract=off */
#define QT_ADD_RN(a, b) ((a) + (b))
#endif

struct Pcg64 {
    unsigned __int128 state;
    unsigned __int128 inc;
    uint64_t draws;  // raw 64-bit outputs produced so far (lets a caller resume a stream)
};

QT_RNG_FN uint32_t ss_hashmix(uint32_t value, uint32_t &hash_const) {
    value ^= hash_const;
    hash_const *= 0x931e8875u;
    value *= hash_const;
    value ^= value >> 16;
    return value;
}

QT_RNG_FN uint32_t ss_mix(uint32_t x, uint32_t y) {
    uint32_t r = 0xca01f9ddu * x - 0x4973f715u * y;
    r ^= r >> 16;
    return r;
}

QT_RNG_FN unsigned __int128 pcg_mult() {
    return ((unsigned __int128)2549297995355413924ULL << 64) | 4865540595714422341ULL;
}

// SeedSequence(seed).generate_state(4, uint64) -> PCG64 seeding (pcg64_set_seed / srandom_r).
QT_RNG_FN void pcg64_seed(Pcg64 &g, uint64_t seed) {
    // entropy as little-endian uint32 words; numpy drops high zero words (0 -> [0])
    uint32_t ent[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    int n_ent = ent[1] ? 2 : 1;
    uint32_t pool[4];
    uint32_t hc = 0x43b0d7e5u;
#pragma unroll
    for (int i = 0; i < 4; i++) pool[i] = ss_hashmix(i < n_ent ? ent[i] : 0u, hc);
#pragma unroll
    for (int s = 0; s < 4; s++)
#pragma unroll
        for (int d = 0; d < 4; d++)
            if (s != d) pool[d] = ss_mix(pool[d], ss_hashmix(pool[s], hc));
    uint32_t out[8];
    uint32_t hb = 0x8b51f9ddu;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t v = pool[i & 3];
        v ^= hb;
        hb *= 0x58f38dedu;
        v *= hb;
        v ^= v >> 16;
        out[i] = v;
    }
    uint64_t w0 = (uint64_t)out[0] | ((uint64_t)out[1] << 32);
    uint64_t w1 = (uint64_t)out[2] | ((uint64_t)out[3] << 32);
    uint64_t w2 = (uint64_t)out[4] | ((uint64_t)out[5] << 32);
    uint64_t w3 = (uint64_t)out[6] | ((uint64_t)out[7] << 32);
    unsigned __int128 initstate = ((unsigned __int128)w0 << 64) | w1;
    unsigned __int128 initseq = ((unsigned __int128)w2 << 64) | w3;
    g.state = 0;
    g.inc = (initseq << 1) | 1u;
    g.state = g.state * pcg_mult() + g.inc;
    g.state += initstate;
    g.state = g.state * pcg_mult() + g.inc;
    g.draws = 0;
}

// Skip n raw outputs (pcg_advance_lcg_128: O(log n) jump-ahead of the underlying LCG).
QT_RNG_FN void pcg64_advance(Pcg64 &g, uint64_t n) {
    unsigned __int128 acc_mult = 1, acc_plus = 0, cur_mult = pcg_mult(), cur_plus = g.inc;
    uint64_t delta = n;
    while (delta > 0) {
        if (delta & 1) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    g.state = acc_mult * g.state + acc_plus;
    g.draws += n;
}

QT_RNG_FN uint64_t pcg64_next(Pcg64 &g) {
    g.draws += 1;
    g.state = g.state * pcg_mult() + g.inc;
    uint64_t hi = (uint64_t)(g.state >> 64), lo = (uint64_t)g.state;
    uint64_t x = hi ^ lo;
    unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((-rot) & 63u));
}

QT_RNG_FN double pcg64_double(Pcg64 &g) {
    return (double)(pcg64_next(g) >> 11) * (1.0 / 9007199254740992.0);
}

// random_uniform(state, low, high - low)
QT_RNG_FN double rng_uniform(Pcg64 &g, double low, double high) {
    return QT_ADD_RN(low, QT_MUL_RN(high - low, pcg64_double(g)));
}

QT_RNG_FN double rng_standard_normal(Pcg64 &g) {
    const double zig_r = 3.6541528853610087963519472518;
    const double zig_inv_r = 0.27366123732975827203338247596;
    for (;;) {
        uint64_t r = pcg64_next(g);
        int idx = (int)(r & 0xff);
        r >>= 8;
        int sign = (int)(r & 0x1);
        uint64_t rabs = (r >> 1) & 0x000fffffffffffffULL;
        double x = (double)rabs * QT_ZIG_WI(idx);
        if (sign) x = -x;
        if (rabs < QT_ZIG_KI(idx)) return x;
        if (idx == 0) {
            for (;;) {
                double xx = -zig_inv_r * log1p(-pcg64_double(g));
                double yy = -log1p(-pcg64_double(g));
                if (yy + yy > xx * xx) return ((rabs >> 8) & 0x1) ? -(zig_r + xx) : zig_r + xx;
            }
        } else {
            double f1 = QT_ZIG_FI(idx - 1);
            double f0 = QT_ZIG_FI(idx);
            if (QT_ADD_RN(QT_MUL_RN(f1 - f0, pcg64_double(g)), f0) < exp(-0.5 * x * x)) return x;
        }
    }
}

}  // namespace qt
