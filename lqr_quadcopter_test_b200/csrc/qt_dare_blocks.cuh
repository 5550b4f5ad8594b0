// qt_dare_blocks.cuh — the decoupled path of the batched DARE solver.
//
// The systems a Q/R sweep generates (controllers/riccati_lqr.py:187-316 design matrices, 602-700 diagonal weights;
// controllers/tuning.py:846-857 one DARE per candidate) are block diagonal after a permutation of the states: the
// x / y / z axes are (position, velocity[, integral]) chains that touch each other nowhere — A couples p_i with v_i
// (and the integral with p_i), every control drives one velocity, Q and R are diagonal.  Structure-preserving
// doubling keeps that structure: products, sums and inverses of matrices that are block diagonal over the same
// partition are block diagonal again, so the n x n recursion IS s independent recursions on the blocks
// (three 2 x 2 for the reference's LQR design, three 3 x 3 for LQI) at (1/9) of the flops.  The dense kernel
// (`dare_kernel<N, R>`) multiplies all those zeros; this kernel finds the blocks and solves them in registers, one
// thread per block, with no shared-memory operand traffic at all.  It is the DARE's counterpart of the rollout's
// structure-exploiting loop: the same numbers (the nonzero terms are accumulated in the same order with the same
// fused multiply-adds and the same pivot rule as the dense kernel; an exact zero contributes nothing to an fma
// chain), decided per system from the data, with the dense kernel as the path for everything else.
//
// Per warp: 32 consecutive systems.
//  1. lane = system: read A, B, Q, R; the path applies when Q and R are exactly diagonal, and the connected
//     components of the graph {i ~ j when A_ij != 0 or A_ji != 0, or one control drives both i and j} have at
//     most 3 states and at most one control each.  Diagonal Q / R are validated from their diagonals with the
//     reference's tolerances (riccati_lqr.py:57-116: eigenvalues of Q >= -1e-10, of R > 1e-10, entries finite).
//     Anything else is marked for the dense path (info = kDareMarkDense).
//  2. the components of the warp's systems are appended to three per-warp lists in shared memory (by size),
//     so that the solve phase runs with all 32 lanes busy on blocks of one size (LQI: 96 blocks = 3 full passes).
//  3. lane = block: G0 = b b' / r, H0 = diag(q); doubling on the S x S block in registers (Gauss-Jordan with the
//     dense kernel's pivot rule) until ||dH||_F <= 1e-14 ||H||_F for the block; P block, gain row; results go
//     to a per-warp output tile in shared memory.
//  4. the tile (K [32][m][n], and P [32][n][n] when P is asked for; zero outside the blocks — exact zeros, which is what lets the rollout
//     kernel's structure test accept the gain) is written to HBM with fully coalesced stores.
// Convergence is tested per block (each block meets the tolerance on its own norm, which implies the dense
// kernel's whole-matrix test); `iters` reports the largest count over the blocks of a system.
#pragma once
#include "qt_common.cuh"

namespace qt {

constexpr int kDareMarkDense = -1;  // info: left to the dense path (not yet validated)
constexpr int kDareMarkValid = -2;  // info: validated, to be solved by the dense kernel
constexpr int kBlkMaxS = 3;         // largest block the register solver is instantiated for
constexpr int kBlkMaxTile = 192;    // doubles of P + K per system the per-warp output tile may hold (32 x 192 x 8 = 48 KB)
// Registers per thread (A/B on one B200, profiles/r02_dare_blocks_ab.txt).  The 3 x 3 block solver needs ~166 for a
// spill-free doubling loop (12 warps per SM); systems of up to six states mostly decompose into 2 x 2 blocks, whose
// solver is small, and gain more from 16 resident warps (128 registers; only the 3 x 3 instance spills there).
#ifndef QT_BLK_MAXREG
#define QT_BLK_MAXREG 168
#endif
#ifndef QT_BLK_MAXREG6
#define QT_BLK_MAXREG6 128
#endif

struct DareArgs;  // qt_dare.cu

// pivot key of the dense kernel (gauss_jordan in qt_dare.cu): top 26 bits of |x|'s high word, lowest row wins ties
__device__ __forceinline__ unsigned blk_pivot_key(double x, int row) {
    const unsigned hi = (unsigned)__double2hiint(x) & 0x7fffffffu;
    return (hi & 0xffffffc0u) | (unsigned)(63 - row);
}

// Gauss-Jordan on an S x S block in registers: on exit r1 = W^-1 R1 and (TWO) r2 = W^-1 R2.  Rows are exchanged
// physically (conditional moves), not normalised until the end, multipliers are w * (1 / pivot) — the dense kernel's
// arithmetic, so that the two paths round alike.
template <int S, bool TWO>
__device__ __forceinline__ bool blk_gauss_jordan(double (&w)[S][S], double (&r1)[S][S], double (&r2)[S][S]) {
    bool ok = true;
    double inv[S];
#pragma unroll
    for (int c = 0; c < S; c++) {
        unsigned best = blk_pivot_key(w[c][c], c);
#pragma unroll
        for (int i = c + 1; i < S; i++) best = max(best, blk_pivot_key(w[i][c], i));
        if ((best >> 6) == 0u || best >= 0x7ff00000u) ok = false;  // zero / NaN / inf pivot
        const int prow = 63 - (int)(best & 63u);
#pragma unroll
        for (int i = c + 1; i < S; i++) {
            if (prow == i) {
#pragma unroll
                for (int j = 0; j < S; j++) {
                    double t = w[c][j];
                    w[c][j] = w[i][j];
                    w[i][j] = t;
                    t = r1[c][j];
                    r1[c][j] = r1[i][j];
                    r1[i][j] = t;
                    if (TWO) {
                        t = r2[c][j];
                        r2[c][j] = r2[i][j];
                        r2[i][j] = t;
                    }
                }
            }
        }
        inv[c] = 1.0 / w[c][c];
#pragma unroll
        for (int r = 0; r < S; r++) {
            if (r == c) continue;
            const double f = w[r][c] * inv[c];
            w[r][c] = 0.0;
#pragma unroll
            for (int j = c + 1; j < S; j++) w[r][j] = fma(-f, w[c][j], w[r][j]);
#pragma unroll
            for (int j = 0; j < S; j++) {
                r1[r][j] = fma(-f, r1[c][j], r1[r][j]);
                if (TWO) r2[r][j] = fma(-f, r2[c][j], r2[r][j]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < S; r++)
#pragma unroll
        for (int j = 0; j < S; j++) {
            r1[r][j] *= inv[r];
            if (TWO) r2[r][j] *= inv[r];
        }
    return ok;
}

// c = x y (S x S, accumulated from zero in ascending k like rows_times of the dense kernel)
template <int S>
__device__ __forceinline__ void blk_mul(const double (&x)[S][S], const double (&y)[S][S], double (&c)[S][S]) {
#pragma unroll
    for (int i = 0; i < S; i++)
#pragma unroll
        for (int j = 0; j < S; j++) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < S; k++) acc = fma(x[i][k], y[k][j], acc);
            c[i][j] = acc;
        }
}

// One S x S block: states idx[0..S-1] (ascending) of system `sys`, driven by control `ctrl` (or none: ctrl < 0).
// Writes the block of P and the control's gain row into the warp's output tile; returns (iters, info).
template <int S>
__device__ __forceinline__ void blk_solve(const double *Ap, const double *Bp, const double *Qp, const double *Rp,
                                          int n, int m, unsigned mask, int ctrl, double *tileP, double *tileK,
                                          int &iters_out, int &info_out, int max_iter, double tol) {
    int idx[S];
#pragma unroll
    for (int k = 0; k < S; k++) {
        idx[k] = __ffs(mask) - 1;
        mask &= mask - 1;
    }
    double a[S][S], g[S][S], h[S][S], bv[S];
    double rc = 1.0;
#pragma unroll
    for (int i = 0; i < S; i++) {
#pragma unroll
        for (int j = 0; j < S; j++) {
            a[i][j] = Ap[idx[i] * n + idx[j]];
            h[i][j] = (i == j) ? Qp[idx[i] * n + idx[i]] : 0.0;  // Q is diagonal on this path
        }
        bv[i] = ctrl >= 0 ? Bp[idx[i] * m + ctrl] : 0.0;
    }
    if (ctrl >= 0) rc = Rp[ctrl * m + ctrl];
    // G0 = B R^-1 B' restricted to the block: z = (1 / r) b' (the dense kernel's Gauss-Jordan on the diagonal R
    // multiplies by the reciprocal), g_ij = b_i z_j
    {
        const double rinv = 1.0 / rc;
#pragma unroll
        for (int i = 0; i < S; i++)
#pragma unroll
            for (int j = 0; j < S; j++) g[i][j] = ctrl >= 0 ? fma(bv[i], bv[j] * rinv, 0.0) : 0.0;
    }
    bool ok = true, conv = false;
    int iters = 0;
    for (int it = 0; it < max_iter && !conv; it++) {
        double w[S][S], x[S][S], y[S][S];
        blk_mul<S>(g, h, w);
#pragma unroll
        for (int i = 0; i < S; i++)
#pragma unroll
            for (int j = 0; j < S; j++) {
                if (i == j) w[i][j] += 1.0;  // W = I + G H
                x[i][j] = a[i][j];
                y[i][j] = g[i][j];
            }
        ok &= blk_gauss_jordan<S, true>(w, x, y);  // X = W^-1 A, Y = W^-1 G
        {
            double t[S][S];
            blk_mul<S>(a, y, t);  // A Y
#pragma unroll
            for (int i = 0; i < S; i++)
#pragma unroll
                for (int j = 0; j < S; j++) {
                    double acc = g[i][j];
#pragma unroll
                    for (int k = 0; k < S; k++) acc = fma(t[i][k], a[j][k], acc);  // G + (A Y) A'
                    w[i][j] = acc;  // (w is free: holds G_{k+1} until A has been used below)
                }
        }
        double an[S][S], hx[S][S], hn[S][S];
        blk_mul<S>(a, x, an);  // A_{k+1} = A X
        blk_mul<S>(h, x, hx);  // H X
        double d2 = 0.0, h2 = 0.0;
#pragma unroll
        for (int i = 0; i < S; i++)
#pragma unroll
            for (int j = 0; j < S; j++) {
                double acc = h[i][j];
#pragma unroll
                for (int k = 0; k < S; k++) acc = fma(a[k][i], hx[k][j], acc);  // H + A' (H X)
                hn[i][j] = acc;
                const double d = acc - h[i][j];
                d2 = fma(d, d, d2);
                h2 = fma(acc, acc, h2);
            }
#pragma unroll
        for (int i = 0; i < S; i++)
#pragma unroll
            for (int j = 0; j < S; j++) {
                a[i][j] = an[i][j];
                g[i][j] = w[i][j];
                h[i][j] = hn[i][j];
            }
        iters = it + 1;
        if (d2 <= tol * tol * h2) conv = true;
        if (!(d2 == d2) || !(h2 <= 1.79769313486231570e308)) {  // NaN, or H overflowed (inf <= inf would pass the test above): stop, reported through info
            conv = true;
            ok = false;
        }
    }

    // P = (H + H') / 2
    double p[S][S];
#pragma unroll
    for (int i = 0; i < S; i++)
#pragma unroll
        for (int j = 0; j < S; j++) {
            p[i][j] = 0.5 * (h[i][j] + h[j][i]);
            if (tileP) tileP[idx[i] * n + idx[j]] = p[i][j];
        }
    // K row of the block's control: (r + b' P b)^-1 b' P A with the ORIGINAL A
    if (ctrl >= 0) {
        double pb[S], den = rc;
#pragma unroll
        for (int i = 0; i < S; i++) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < S; k++) acc = fma(p[i][k], bv[k], acc);
            pb[i] = acc;
        }
#pragma unroll
        for (int k = 0; k < S; k++) den = fma(bv[k], pb[k], den);
        const double dinv = 1.0 / den;
        {
            const unsigned hi = (unsigned)__double2hiint(den) & 0x7fffffffu;
            if (hi < 64u || hi >= 0x7ff00000u) ok = false;  // the dense kernel's zero / NaN / inf pivot test
        }
#pragma unroll
        for (int j = 0; j < S; j++) {
            double num = 0.0;
#pragma unroll
            for (int k = 0; k < S; k++) {
                double pa = 0.0;  // (P A)[k][j]
#pragma unroll
                for (int l = 0; l < S; l++) pa = fma(p[k][l], Ap[idx[l] * n + idx[j]], pa);
                num = fma(bv[k], pa, num);
            }
            if (tileK) tileK[ctrl * n + idx[j]] = num * dinv;
        }
    }
    iters_out = iters;
    info_out = !ok ? 2 : (conv ? 0 : 1);
}

// NMAX: compile-time bound of n for the unrolled structure detection (6, 9 or 12).
template <int NMAX>
__global__ void __maxnreg__(NMAX <= 6 ? QT_BLK_MAXREG6 : QT_BLK_MAXREG)
    dare_blocks_kernel(int n, int m, int64_t batch, const double *__restrict__ A, const double *__restrict__ B,
                       const double *__restrict__ Q, const double *__restrict__ R, int64_t sA, int64_t sB, int64_t sQ,
                       int64_t sR, double *__restrict__ P_out, double *__restrict__ K_out, int32_t *__restrict__ iters_out,
                       int32_t *__restrict__ info_out, int max_iter, double tol) {
    extern __shared__ __align__(16) double blk_smem[];
    const int lane = threadIdx.x;
    const int64_t sys0 = (int64_t)blockIdx.x * 32;
    const int64_t sys = sys0 + lane;
    const bool live = sys < batch;
    const int nn = n * n, mn = m * n;

    // shared memory of the warp: output tile (the P part only when P is asked for: the tuner never is, and the tile is
    // what bounds the residency for n = 9) | lists by block size | per-system (iters, info) | list counters
    double *tileK = blk_smem;                                  // [32][m][n]
    double *tileP = P_out ? tileK + 32 * mn : nullptr;         // [32][n][n]
    const int tile_doubles = 32 * (mn + (P_out ? nn : 0));
    uint32_t *list1 = reinterpret_cast<uint32_t *>(blk_smem + tile_doubles);  // capacity 32 n
    uint32_t *list2 = list1 + 32 * n;              // capacity 32 (n / 2)
    uint32_t *list3 = list2 + 32 * (n / 2);        // capacity 32 (n / 3)
    int *s_iters = reinterpret_cast<int *>(list3 + 32 * (n / 3));
    int *s_info = s_iters + 32;
    int *s_cnt = s_info + 32;                      // [3]
    for (int i = lane; i < tile_doubles; i += 32) blk_smem[i] = 0.0;
    s_iters[lane] = 0;
    s_info[lane] = 0;
    if (lane < 3) s_cnt[lane] = 0;

    const double *Ap = A + (live ? sys : 0) * sA, *Bp = B + (live ? sys : 0) * sB;
    const double *Qp = Q + (live ? sys : 0) * sQ, *Rp = R + (live ? sys : 0) * sR;

    // ---- 1. structure and validation, lane = system
    bool fast = live, valid = true;
    uint32_t reach[NMAX], ctl[QT_MAX_M];
#pragma unroll
    for (int i = 0; i < NMAX; i++) reach[i] = 1u << i;
    if (live) {
#pragma unroll
        for (int i = 0; i < NMAX; i++)
#pragma unroll
            for (int j = 0; j < NMAX; j++) {
                if (i < n && j < n) {
                    const double qv = Qp[i * n + j];
                    if (i != j) {
                        if (Ap[i * n + j] != 0.0) {
                            reach[i] |= 1u << j;
                            reach[j] |= 1u << i;
                        }
                        if (qv != 0.0) fast = false;  // (NaN included: the dense path's validation decides)
                    } else {
                        if (!(fabs(qv) <= 1.79769313486231570e308) || qv < -1e-10) valid = false;
                    }
                }
            }
#pragma unroll
        for (int c = 0; c < QT_MAX_M; c++) {
            uint32_t sc = 0u;
            if (c < m) {
#pragma unroll
                for (int i = 0; i < NMAX; i++)
                    if (i < n && Bp[i * m + c] != 0.0) sc |= 1u << i;
                for (int d = 0; d < m; d++) {
                    const double rv = Rp[c * m + d];
                    if (c != d) {
                        if (rv != 0.0) fast = false;
                    } else if (!(fabs(rv) <= 1.79769313486231570e308) || !(rv > 1e-10)) valid = false;
                }
            }
            ctl[c] = sc;
#pragma unroll
            for (int i = 0; i < NMAX; i++)
                if ((sc >> i) & 1u) reach[i] |= sc;
        }
        // transitive closure (Warshall on bit rows)
#pragma unroll
        for (int k = 0; k < NMAX; k++)
#pragma unroll
            for (int i = 0; i < NMAX; i++)
                if ((reach[i] >> k) & 1u) reach[i] |= reach[k];
        // every component: at most kBlkMaxS states and one control
        uint32_t seen = 0u;
#pragma unroll
        for (int i = 0; i < NMAX; i++) {
            if (i < n && !((seen >> i) & 1u)) {
                const uint32_t comp = reach[i];
                seen |= comp;
                int nc = 0;
#pragma unroll
                for (int c = 0; c < QT_MAX_M; c++) nc += (ctl[c] & comp) ? 1 : 0;
                if (__popc(comp) > kBlkMaxS || nc > 1) fast = false;
            }
        }
    }
    __syncwarp();
    if (fast && valid) {
        uint32_t seen = 0u;
#pragma unroll
        for (int i = 0; i < NMAX; i++) {
            if (i < n && !((seen >> i) & 1u)) {
                const uint32_t comp = reach[i];
                seen |= comp;
                int ctrl = -1;
#pragma unroll
                for (int c = 0; c < QT_MAX_M; c++)
                    if (ctl[c] & comp) ctrl = c;
                const int s = __popc(comp);
                const uint32_t entry = (uint32_t)lane | ((uint32_t)(ctrl + 1) << 5) | (comp << 9);
                const int pos = atomicAdd(&s_cnt[s - 1], 1);
                (s == 1 ? list1 : s == 2 ? list2 : list3)[pos] = entry;
            }
        }
    }
    __syncwarp();

    // ---- 2. lane = block, largest blocks first
    auto run_list = [&](auto size_tag, const uint32_t *list, int count) {
        constexpr int S = decltype(size_tag)::value;
        for (int e = lane; e < count; e += 32) {
            const uint32_t entry = list[e];
            const int sl = entry & 31u, ctrl = (int)((entry >> 5) & 15u) - 1;
            const uint32_t comp = entry >> 9;
            const int64_t s2 = sys0 + sl;
            int it, inf;
            blk_solve<S>(A + s2 * sA, B + s2 * sB, Q + s2 * sQ, R + s2 * sR, n, m, comp, ctrl, tileP ? tileP + sl * nn : nullptr,
                         tileK + sl * mn, it, inf, max_iter, tol);
            atomicMax(&s_iters[sl], it);
            atomicMax(&s_info[sl], inf);
        }
    };
    run_list(std::integral_constant<int, 3>{}, list3, s_cnt[2]);
    run_list(std::integral_constant<int, 2>{}, list2, s_cnt[1]);
    run_list(std::integral_constant<int, 1>{}, list1, s_cnt[0]);
    __syncwarp();

    // ---- 3. results
    if (live) {
        if (!fast) {
            info_out[sys] = kDareMarkDense;
            iters_out[sys] = 0;
        } else if (!valid) {
            info_out[sys] = 3;
            iters_out[sys] = 0;
        } else {
            info_out[sys] = s_info[lane];
            iters_out[sys] = s_iters[lane];
        }
    }
    const int64_t n_live = (batch - sys0 < 32) ? (batch - sys0) : 32;
    if (P_out) {
        double *dst = P_out + sys0 * nn;
        for (int64_t i = lane; i < n_live * nn; i += 32) dst[i] = tileP[i];
    }
    if (K_out) {
        double *dst = K_out + sys0 * mn;
        for (int64_t i = lane; i < n_live * mn; i += 32) dst[i] = tileK[i];
    }
}

inline size_t dare_blocks_smem(int n, int m, bool want_P) {
    return (size_t)32 * ((want_P ? n * n : 0) + m * n) * sizeof(double) +
           (size_t)32 * (n + n / 2 + n / 3) * sizeof(uint32_t) + (32 + 32 + 4) * sizeof(int);
}

inline bool dare_blocks_applicable(int n, int m) { return n <= 12 && (n * n + m * n) <= kBlkMaxTile; }

template <int NMAX>
static cudaError_t launch_dare_blocks_n(int n, int m, int64_t batch, const double *A, const double *B, const double *Q,
                                        const double *R, int64_t sA, int64_t sB, int64_t sQ, int64_t sR, double *P,
                                        double *K, int32_t *iters, int32_t *info, int max_iter, double tol,
                                        cudaStream_t st) {
    const size_t smem = dare_blocks_smem(n, m, P != nullptr);
    cudaError_t e = cudaFuncSetAttribute(dare_blocks_kernel<NMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dare_blocks_kernel<NMAX><<<(unsigned)((batch + 31) / 32), 32, smem, st>>>(n, m, batch, A, B, Q, R, sA, sB, sQ, sR, P, K,
                                                                             iters, info, max_iter, tol);
    return cudaGetLastError();
}

}  // namespace qt
