// qt_dare.cu — batched fp64 DARE solver for many small systems.
//
// Replaces solve_dare (controllers/riccati_lqr.py:119-184), i.e. scipy.linalg.solve_discrete_are
// followed by K = (R + B'PB)^-1 B'PA, for the thousands of n = 6 / n = 9 systems a Q/R sweep
// generates (controllers/tuning.py:846-857 builds one controller, hence one DARE, per candidate).
//
// Algorithm: structure-preserving doubling (SDA) on (A_k, G_k, H_k), G_0 = B R^-1 B', H_0 = Q:
//     W = I + G_k H_k;  X = W^-1 A_k;  Y = W^-1 G_k
//     A_{k+1} = A_k X;  G_{k+1} = G_k + A_k Y A_k';  H_{k+1} = H_k + A_k' H_k X
// H_k converges quadratically to P (14-18 doublings on the tuning ranges).  scipy uses an
// ordered QZ decomposition of the 2n x 2n symplectic pencil instead; both compute the same
// stabilising solution, agreement is checked to 1e-9 relative on K (tests/test_dare_gpu.py).
//
// Mapping: these are tiny, non-GEMM solves — no tensor cores.  A "team" of N / R lanes of one warp owns one system,
// each lane holding R consecutive rows of every matrix in registers; 32 / (N / R) systems share a warp (N = 6, R = 2:
// ten systems per warp on 30 lanes).  A product needs the other lanes' rows, which are staged in shared memory and
// read with 16-byte broadcast loads: one load feeds 2 R fused multiply-adds (4 R where two products share an
// operand), so with R = 2 the kernel issues half the shared-memory wavefronts per flop of the one-row-per-lane
// layout of round 1, whose limiter they were (ncu r01_e: LSU pipe busier than the FP64 pipe, 59 % of the stall
// samples on the shared-memory scoreboard / queue).  W is inverted by Gauss-Jordan elimination with partial pivoting
// across the team: pivot search with 32-bit shuffle keys, pivot rows broadcast through shared memory, one
// elimination pass per column over all R rows of a lane (the pivot row itself takes a zero multiplier, so the pass
// is branch free).  n and m are zero-padded to the next instantiated N, which leaves the solution of the original
// system unchanged (block-diagonal embedding).
#include "qt_common.cuh"
#include "qt_dare_blocks.cuh"

namespace qt {

// A/B switches for the larger sizes (defaults = the measured winners)
#ifndef QT_DARE_N9
#define QT_DARE_N9 9     // n = 7..9 (LQI): 3 lanes x 3 rows, 10 systems per warp
#define QT_DARE_R9 3
#endif
#ifndef QT_DARE_BLOCKSYNC
#define QT_DARE_BLOCKSYNC 20  // kernels with R * N >= this value keep the warps of a block in step (see dare_kernel): n = 9 +6 %, n = 12 +3 %, n = 6 -4 % (A/B r02_o)
#endif
#ifndef QT_DARE_R6
#define QT_DARE_R6 3     // n = 5, 6: 2 lanes x 3 rows, 16 systems per warp
#endif
#ifndef QT_DARE_R12
#define QT_DARE_R12 2    // 6 lanes x 2 rows, 5 systems per warp
#endif
#ifndef QT_DARE_R16
#define QT_DARE_R16 1    // 16 lanes x 1 row, 2 systems per warp (2 rows per lane spill)
#endif

constexpr int kDareMaxIter = 64;
constexpr double kDareTol = 1e-14;  // relative Frobenius change of H

struct DareArgs {
    int n, m;
    int64_t batch;
    const double *A, *B, *Q, *R;
    int64_t sA, sB, sQ, sR;
    double *P, *K;
    int32_t *iters, *info;
};

template <int N, int R>
struct Lay {
    static_assert(N % R == 0, "N must be a multiple of the rows per lane");
    static constexpr int TEAM = N / R;             // lanes per system
    static constexpr int LD = (N + 1) & ~1;        // even row stride: rows stay 16-byte aligned
    static constexpr int MAT = N * LD;
    static constexpr int RAW = 4 * MAT + 4 * LD;   // sA sH sX sY | pivot rows (3 LD) + 1/pivot (LD)
    // Teams of one warp issue the same 16-byte broadcast load at the same offset inside their own block; padding
    // the block to 2 (mod 16) doubles = 4 (mod 32) banks puts neighbouring teams on disjoint bank groups (the
    // unpadded N = 6 layout of the first version was a 5-way conflict: ncu, profiles/r01_a).
    static constexpr int BLOCK = RAW + ((2 - RAW % 16) + 16) % 16;
    static constexpr int TEAMS = 32 / TEAM;        // systems per warp
    static constexpr int SLOTS = TEAMS + ((32 % TEAM) ? 1 : 0);  // + a ghost team for leftover lanes
};

// NOTE: none of the shared-memory pointers below is __restrict__-qualified.  The team's lanes communicate through
// these buffers (a row stored by one lane is loaded by the others after a __syncwarp), and with restrict the
// compiler may keep or forward values across that barrier: the first version of this kernel with restrict on the
// helpers returned garbage for every system on sm_100a (nvcc 12.9) while a lane-by-lane CPU emulation of the same
// source was exact.
template <int N>
__device__ __forceinline__ void store_row(double *M, int row, const double (&x)[N]) {
    constexpr int LD = (N + 1) & ~1;
#pragma unroll
    for (int j = 0; j < N; j += 2) {
        double2 v;
        v.x = x[j];
        v.y = (j + 1 < N) ? x[j + 1] : 0.0;
        *reinterpret_cast<double2 *>(M + row * LD + j) = v;
    }
}

template <int N>
__device__ __forceinline__ void load_row(const double *M, int row, double (&x)[N]) {
    constexpr int LD = (N + 1) & ~1;
#pragma unroll
    for (int j = 0; j < N; j += 2) {
        const double2 v = *reinterpret_cast<const double2 *>(M + row * LD + j);
        x[j] = v.x;
        if (j + 1 < N) x[j + 1] = v.y;
    }
}

// c[r][j] (+)= sum_k x[r][k] * Y[k][j] for the R rows of this lane,  Y row-major in shared memory
template <int N, int R, bool ACCUM>
__device__ __forceinline__ void rows_times(const double (&x)[R][N], const double *Y, double (&c)[R][N]) {
    constexpr int LD = (N + 1) & ~1;
    if (!ACCUM) {
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int j = 0; j < N; j++) c[r][j] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < N; k++) {
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            const double2 y = *reinterpret_cast<const double2 *>(Y + k * LD + j);
#pragma unroll
            for (int r = 0; r < R; r++) {
                c[r][j] = fma(x[r][k], y.x, c[r][j]);
                if (j + 1 < N) c[r][j + 1] = fma(x[r][k], y.y, c[r][j + 1]);
            }
        }
    }
}

// c1 = x1 Y and c2 = x2 Y: two products sharing the loads of Y (4 R multiply-adds per 16-byte load)
template <int N, int R>
__device__ __forceinline__ void rows_times2(const double (&x1)[R][N], const double (&x2)[R][N],
                                            const double *Y, double (&c1)[R][N], double (&c2)[R][N]) {
    constexpr int LD = (N + 1) & ~1;
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int j = 0; j < N; j++) c1[r][j] = c2[r][j] = 0.0;
#pragma unroll
    for (int k = 0; k < N; k++) {
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            const double2 y = *reinterpret_cast<const double2 *>(Y + k * LD + j);
#pragma unroll
            for (int r = 0; r < R; r++) {
                c1[r][j] = fma(x1[r][k], y.x, c1[r][j]);
                c2[r][j] = fma(x2[r][k], y.x, c2[r][j]);
                if (j + 1 < N) {
                    c1[r][j + 1] = fma(x1[r][k], y.y, c1[r][j + 1]);
                    c2[r][j + 1] = fma(x2[r][k], y.y, c2[r][j + 1]);
                }
            }
        }
    }
}

// c[r][j] += sum_k x[r][k] * Y[j][k]   (product with the transpose of Y)
template <int N, int R>
__device__ __forceinline__ void rows_times_T(const double (&x)[R][N], const double *Y, double (&c)[R][N]) {
    constexpr int LD = (N + 1) & ~1;
#pragma unroll
    for (int j = 0; j < N; j++) {
        double acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = c[r][j];
#pragma unroll
        for (int k = 0; k < N; k += 2) {
            const double2 y = *reinterpret_cast<const double2 *>(Y + j * LD + k);
#pragma unroll
            for (int r = 0; r < R; r++) {
                acc[r] = fma(x[r][k], y.x, acc[r]);
                if (k + 1 < N) acc[r] = fma(x[r][k + 1], y.y, acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) c[r][j] = acc[r];
    }
}

// Gauss-Jordan with partial pivoting, R matrix rows per lane.  On exit the (lane, slot) that served as pivot for
// column c holds row c of W^-1 [R1 | R2] in (r1 | r2) and has written it to X[c][:] / Y[c][:].  Returns false if a
// pivot was zero / not finite.
//
// Per column: (1) pivot search without shared memory — every lane publishes one 32-bit key (the top 26 bits of the
// high word of |w[c]| of its best unused row, that row's index in the low 6 bits, inverted so the lowest row wins
// ties) and reads its team mates' keys with shuffles; (2) the owner writes the pivot row and 1/pivot to shared
// memory; (3) every lane eliminates in all of its R rows — the pivot row itself with a zero multiplier, which leaves
// it unchanged without a branch.  Rows are not normalised until the end.  The multiplier is w[c] * (1/pivot): the
// matrices of the reference's designs are sparse, and CUDA's double division takes its slow path for a zero
// numerator, which made one third of the first version's instructions.
template <int N, int R, bool TWO>
__device__ __forceinline__ bool gauss_jordan(double (&w)[R][N], double (&r1)[R][N], double (&r2)[R][N],
                                             double *sRow, double *X,
                                             double *Y, int li, int team_base) {
    constexpr int LD = (N + 1) & ~1, TEAM = N / R;
    bool ok = true;
    bool used[R];
    int my_col[R];
    double my_inv[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        used[r] = false;
        my_col[r] = 0;
        my_inv[r] = 1.0;
    }
#pragma unroll
    for (int c = 0; c < N; c++) {
        unsigned key = 0u;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const unsigned hi = (unsigned)__double2hiint(w[r][c]) & 0x7fffffffu;
            const unsigned k = used[r] ? 0u : ((hi & 0xffffffc0u) | (unsigned)(63 - (li * R + r)));
            key = max(key, k);
        }
        unsigned best = 0u;
#pragma unroll
        for (int k = 0; k < TEAM; k++) best = max(best, __shfl_sync(0xffffffffu, key, team_base + k));
#ifdef QT_DARE_EXTRA_SYNC
        __syncwarp();
#endif
        const int prow = 63 - (int)(best & 63u);
        if ((best >> 6) == 0u || best >= 0x7ff00000u) ok = false;  // zero / NaN / inf pivot
        const int pl = prow / R, ps = prow % R;
        if (li == pl) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (ps == r) {
                    used[r] = true;
                    my_col[r] = c;
                    my_inv[r] = 1.0 / w[r][c];
                    // only columns > c of W are still needed by the other rows
#pragma unroll
                    for (int j = c & ~1; j < N; j += 2) {
                        double2 v;
                        v.x = w[r][j];
                        v.y = (j + 1 < N) ? w[r][j + 1] : 0.0;
                        *reinterpret_cast<double2 *>(sRow + j) = v;
                    }
                    store_row<N>(sRow + LD, 0, r1[r]);
                    if (TWO) store_row<N>(sRow + 2 * LD, 0, r2[r]);
                    sRow[3 * LD] = my_inv[r];
                }
            }
        }
        __syncwarp();
        {
            const double inv = sRow[3 * LD];
            double f[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const bool is_pivot = (li == pl) & (ps == r);
                f[r] = is_pivot ? 0.0 : w[r][c] * inv;
                w[r][c] = is_pivot ? w[r][c] : 0.0;
            }
#pragma unroll
            for (int j = (c + 1) & ~1; j < N; j += 2) {
                const double2 y = *reinterpret_cast<const double2 *>(sRow + j);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if (j > c) w[r][j] = fma(-f[r], y.x, w[r][j]);
                    if (j + 1 < N) w[r][j + 1] = fma(-f[r], y.y, w[r][j + 1]);
                }
            }
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                const double2 y = *reinterpret_cast<const double2 *>(sRow + LD + j);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    r1[r][j] = fma(-f[r], y.x, r1[r][j]);
                    if (j + 1 < N) r1[r][j + 1] = fma(-f[r], y.y, r1[r][j + 1]);
                }
            }
            if (TWO) {
#pragma unroll
                for (int j = 0; j < N; j += 2) {
                    const double2 y = *reinterpret_cast<const double2 *>(sRow + 2 * LD + j);
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        r2[r][j] = fma(-f[r], y.x, r2[r][j]);
                        if (j + 1 < N) r2[r][j + 1] = fma(-f[r], y.y, r2[r][j + 1]);
                    }
                }
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
#pragma unroll
        for (int j = 0; j < N; j++) {
            r1[r][j] *= my_inv[r];
            if (TWO) r2[r][j] *= my_inv[r];
        }
        store_row<N>(X, my_col[r], r1[r]);
        if (TWO) store_row<N>(Y, my_col[r], r2[r]);
    }
    __syncwarp();
    return ok;
}

// sum of v over the lanes of the team, the same order in every lane (so all lanes hold identical bits)
template <int TEAM>
__device__ __forceinline__ double team_sum(double v, int team_base) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < TEAM; k++) s += __shfl_sync(0xffffffffu, v, team_base + k);
    return s;
}

template <int N, int R>
struct DareCfg {
    static constexpr int RN = R * N;  // doubles per matrix held by one lane; ~4 RN doubles are live at the widest point
    // warps per block and resident blocks per SM the register allocator must allow (shared memory per team is
    // 4 N^2 + 4 N doubles, which is what bounds the residency of the larger sizes)
    // (measured, profiles/r02_e_dare_ab.txt: n = 6 as 2 lanes x 3 rows runs 1.45e8 solves/s with 3-warp blocks and
    // 1.76e8 with 4-warp blocks at 2 blocks per SM — the register file is allocated per 4 warps)
    static constexpr int WARPS = RN <= 20 ? 4 : 3;
    static constexpr int MIN_BLOCKS = RN <= 12 ? 3 : 2;
};
#if defined(QT_DARE_W16) && defined(QT_DARE_B16)   // A/B: block shape of the n = 13..16 kernel
template <>
struct DareCfg<16, QT_DARE_R16> {
    static constexpr int RN = 16 * QT_DARE_R16, WARPS = QT_DARE_W16, MIN_BLOCKS = QT_DARE_B16;
};
#endif
#if defined(QT_DARE_W12) && defined(QT_DARE_B12)   // A/B: block shape of the n = 10..12 kernel
template <>
struct DareCfg<12, QT_DARE_R12> {
    static constexpr int RN = 12 * QT_DARE_R12, WARPS = QT_DARE_W12, MIN_BLOCKS = QT_DARE_B12;
};
#endif
#if defined(QT_DARE_W6) && defined(QT_DARE_B6)   // A/B: block shape of the n = 6 kernel
template <>
struct DareCfg<6, QT_DARE_R6> {
    static constexpr int RN = 6 * QT_DARE_R6, WARPS = QT_DARE_W6, MIN_BLOCKS = QT_DARE_B6;
};
#endif
#if defined(QT_DARE_W9) && defined(QT_DARE_B9)   // A/B: block shape of the n = 9 kernel
template <>
struct DareCfg<QT_DARE_N9, QT_DARE_R9> {
    static constexpr int RN = QT_DARE_N9 * QT_DARE_R9, WARPS = QT_DARE_W9, MIN_BLOCKS = QT_DARE_B9;
};
#endif

template <int N, int R>
__device__ __forceinline__ void load_rows(const double *M, int row0, double (&x)[R][N]) {
#pragma unroll
    for (int r = 0; r < R; r++) load_row<N>(M, row0 + r, x[r]);
}

template <int N, int R>
__device__ __forceinline__ void store_rows(double *M, int row0, const double (&x)[R][N]) {
#pragma unroll
    for (int r = 0; r < R; r++) store_row<N>(M, row0 + r, x[r]);
}

// The systems [blk * per_block, (blk + 1) * per_block) of the batch, by one thread block.
template <int N, int R>
__device__ __forceinline__ void dare_block(const DareArgs &D, int64_t blk) {
    extern __shared__ __align__(16) double smem[];
    using L = Lay<N, R>;
    constexpr int LD = L::LD, TEAM = L::TEAM, WARPS = DareCfg<N, R>::WARPS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int team = lane / TEAM, li = lane % TEAM, team_base = lane - li;
    const int row0 = li * R;  // first matrix row held by this lane
    const int64_t sys = (blk * WARPS + warp) * L::TEAMS + team;
    const int n = D.n, m = D.m;
    bool ghost = (team >= L::TEAMS) || (sys >= D.batch);
    // only systems the validation kernel passed on: rejected ones (info = 3) and the ones the decoupled path has
    // already solved (qt_dare_blocks.cuh) are skipped
    if (!ghost && D.info[sys] != kDareMarkValid) ghost = true;
    if (__syncthreads_and(ghost)) return;  // nothing for this block to do (the usual case after the decoupled path)

    // Shared memory of one team: four N x N matrices and the pivot-row buffer.  pA holds A_k, pY receives
    // Y = W^-1 G_k and then A_{k+1} (the two pointers are exchanged at the end of a doubling), sH holds H_k,
    // sX receives X = W^-1 A_k and then H_k X.  G_k is only ever needed row-wise by its owner: registers.
    double *tb = smem + (size_t)(warp * L::SLOTS + team) * L::BLOCK;
    double *pA = tb, *sH = tb + L::MAT, *sX = tb + 2 * L::MAT, *pY = tb + 3 * L::MAT, *sRow = tb + 4 * L::MAT;

    const double *Ap = ghost ? nullptr : D.A + sys * D.sA;
    const double *Bp = ghost ? nullptr : D.B + sys * D.sB;
    const double *Qp = ghost ? nullptr : D.Q + sys * D.sQ;
    const double *Rp = ghost ? nullptr : D.R + sys * D.sR;

    auto load_global_rows = [&](const double *M, int rows, int cols, double (&x)[R][N]) {  // zero-padded rows of M
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int j = 0; j < N; j++)
                x[r][j] = (!ghost && row0 + r < rows && j < cols) ? M[(row0 + r) * cols + j] : 0.0;
    };
    auto load_R_pad = [&](double (&w)[R][N]) {  // R_pad = blockdiag(R, I)
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int j = 0; j < N; j++)
                w[r][j] = (!ghost && row0 + r < m && j < m) ? Rp[(row0 + r) * m + j] : (row0 + r == j ? 1.0 : 0.0);
    };

    double g[R][N];
    bool ok = true;
    {
        // G0 = B R^-1 B':  Z = R_pad^-1 B'_pad (rows of B' as right-hand side), then G0 = B Z
        double w[R][N], r1[R][N], r2[R][N];
        load_R_pad(w);
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int j = 0; j < N; j++) {
                r1[r][j] = (!ghost && row0 + r < m && j < n) ? Bp[j * m + row0 + r] : 0.0;
                r2[r][j] = 0.0;
            }
        ok &= gauss_jordan<N, R, false>(w, r1, r2, sRow, sX, pY, li, team_base);
        double brow[R][N];
        load_global_rows(Bp, n, m, brow);
        rows_times<N, R, false>(brow, sX, g);
        double a[R][N], h[R][N];
        load_global_rows(Ap, n, n, a);
        load_global_rows(Qp, n, n, h);
        __syncwarp();  // every lane has read Z
        store_rows<N, R>(pA, row0, a);
        store_rows<N, R>(sH, row0, h);
    }

    bool conv = false;
    int iters = 0;
    for (int it = 0; it < kDareMaxIter; it++) {
#if QT_DARE_BLOCKSYNC
        // The warps of a block walk the doubling body together: it is tens of KB of straight-line SASS (fully unrolled
        // register-array code), and warps at different places in it evict each other's instruction-cache lines (ncu
        // r02 n = 9: 40 % of the stall samples were instruction fetches).  The block leaves the loop when all of its
        // systems have converged; converged systems keep H frozen, so the extra doublings change nothing.
        if (DareCfg<N, R>::RN >= QT_DARE_BLOCKSYNC) {
            if (__syncthreads_and(conv || ghost)) break;
        } else
#endif
        if (__all_sync(0xffffffffu, conv || ghost)) break;
        __syncwarp();  // the rows of A_k / H_k stored by the team are visible
        {
            double w[R][N], r1[R][N], r2[R][N];
            rows_times<N, R, false>(g, sH, w);  // G H
            load_rows<N, R>(pA, row0, r1);
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int j = 0; j < N; j++) {
                    w[r][j] += (j == row0 + r) ? 1.0 : 0.0;  // W = I + G H
                    r2[r][j] = g[r][j];
                }
            const bool gj_ok = gauss_jordan<N, R, true>(w, r1, r2, sRow, sX, pY, li, team_base);  // X -> sX, Y -> pY
            ok &= conv || gj_ok;
        }
        {
            double a[R][N];
            load_rows<N, R>(pA, row0, a);
            {
                double t[R][N];
                rows_times<N, R, false>(a, pY, t);  // A Y
                rows_times_T<N, R>(t, pA, g);       // G_{k+1} = G + (A Y) A'
            }
            double an[R][N];
            rows_times<N, R, false>(a, sX, an);     // A_{k+1} = A X
            __syncwarp();                           // every lane has read Y: pY now receives A_{k+1}
            store_rows<N, R>(pY, row0, an);
        }
        double h[R][N];
        load_rows<N, R>(sH, row0, h);
        {
            double hx[R][N];
            rows_times<N, R, false>(h, sX, hx);     // H X
            __syncwarp();                           // every lane has read X: sX now receives H X
            store_rows<N, R>(sX, row0, hx);
        }
        __syncwarp();
        double d2 = 0.0, h2 = 0.0;
        {
            // H_{k+1} = H + A' (H X): row i needs column i of A
            double acol[R][N], hn[R][N];
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int k = 0; k < N; k++) acol[r][k] = pA[k * LD + row0 + r];
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int j = 0; j < N; j++) hn[r][j] = h[r][j];
            rows_times<N, R, true>(acol, sX, hn);
            // convergence: ||H+ - H||_F <= tol ||H+||_F over the team
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int j = 0; j < N; j++) {
                    const double d = hn[r][j] - h[r][j];
                    d2 = fma(d, d, d2);
                    h2 = fma(hn[r][j], hn[r][j], h2);
                }
            // A team that has converged keeps H frozen while the rest of its warp goes on (its A and G free-run:
            // they are not outputs), so a system's result never depends on its neighbours in the batch
            if (!conv) store_rows<N, R>(sH, row0, hn);
        }
        const double D2 = team_sum<TEAM>(d2, team_base), H2 = team_sum<TEAM>(h2, team_base);
        if (!conv) {
            iters = it + 1;
            if (D2 <= kDareTol * kDareTol * H2) conv = true;
            if (!(D2 == D2) || !(H2 <= 1.79769313486231570e308)) {  // NaN, or H overflowed (inf <= inf would pass the test above): stop, reported through info
                conv = true;
                ok = false;
            }
        }
        double *tmp = pA;
        pA = pY;
        pY = tmp;
    }

    // P = (H + H') / 2
    __syncwarp();
    double p[R][N];
    {
        double h[R][N];
        load_rows<N, R>(sH, row0, h);
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int j = 0; j < N; j++) p[r][j] = 0.5 * (h[r][j] + sH[j * LD + row0 + r]);
    }
    if (!ghost && D.P) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int gi = row0 + r;
            if (gi < n) {
#pragma unroll
                for (int j = 0; j < N; j++)
                    if (j < n) D.P[sys * (int64_t)n * n + gi * n + j] = p[r][j];
            }
        }
    }

    // K = (R + B'PB)^-1 B'PA
    {
        {
            double a0[R][N], brow[R][N];
            load_global_rows(Ap, n, n, a0);
            load_global_rows(Bp, n, m, brow);
            __syncwarp();  // every lane has read sH (transposed) above
            store_rows<N, R>(pA, row0, a0);
            store_rows<N, R>(pY, row0, brow);
        }
        __syncwarp();
        {
            double pb[R][N];
            rows_times<N, R, false>(p, pY, pb);  // P B
            store_rows<N, R>(sX, row0, pb);
        }
        {
            double pa[R][N];
            rows_times<N, R, false>(p, pA, pa);  // P A
            store_rows<N, R>(sH, row0, pa);
        }
        __syncwarp();
        double bcol[R][N], w[R][N], r1[R][N], r2[R][N];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int k = 0; k < N; k++) bcol[r][k] = pY[k * LD + row0 + r];  // column i of B = row i of B'
        load_R_pad(w);
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int j = 0; j < N; j++) r2[r][j] = 0.0;
        rows_times<N, R, true>(bcol, sX, w);    // R + B'(PB)
        rows_times<N, R, false>(bcol, sH, r1);  // B'(PA)
        __syncwarp();
        ok &= gauss_jordan<N, R, false>(w, r1, r2, sRow, pA, pY, li, team_base);  // K rows -> pA
        if (!ghost && D.K) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int gi = row0 + r;
                if (gi < m) {
#pragma unroll
                    for (int j = 0; j < N; j++)
                        if (j < n) D.K[sys * (int64_t)m * n + gi * n + j] = pA[gi * LD + j];
                }
            }
        }
    }
    // the pivot / NaN tests are evaluated identically by every lane of a team: lane 0 reports
    if (!ghost && li == 0) {
        D.iters[sys] = iters;
        D.info[sys] = !ok ? 2 : (conv ? 0 : 1);
    }
}

// Grid-stride over the batch's blocks of systems.  With the decoupled path in front (qt_dare_blocks.cuh) nearly every
// block finds nothing marked for it: a grid of a few blocks per SM that walks the batch replaces tens of thousands of
// 96 KB blocks that start only to exit (85 us of a 1.5 ms launch for 2^20 n = 9 systems).  The dense-only path keeps
// one block per group (grid = number of groups), which lets the hardware balance systems of different iteration counts.
template <int N, int R>
__global__ void __launch_bounds__(DareCfg<N, R>::WARPS * 32, DareCfg<N, R>::MIN_BLOCKS)
    dare_kernel(const __grid_constant__ DareArgs D, int64_t n_blocks) {
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        dare_block<N, R>(D, blk);
        __syncthreads();   // the next group reuses the shared-memory matrices
    }
}

// ---------------------------------------------------------------------------------------------
// Validation (riccati_lqr.py:57-116): Q symmetric (np.allclose atol 1e-8, rtol 1e-5) with
// eigenvalues >= -1e-10, R symmetric with eigenvalues > 1e-10.  One thread per system.  Exactly
// diagonal matrices (what q_pos/q_vel/q_int/r_controls produce, riccati_lqr.py:620-700) are
// decided from the diagonal; anything else runs cyclic Jacobi sweeps.
__device__ double min_eig_sym(const double *M, int n, int ld) {
    double a[QT_MAX_N * QT_MAX_N];
    bool diag = true;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            // eigvalsh reads the lower triangle only
            const double v = (j <= i) ? M[i * ld + j] : M[j * ld + i];
            a[i * n + j] = v;
            if (i != j && v != 0.0) diag = false;
        }
    if (!diag) {
        for (int sweep = 0; sweep < 30; sweep++) {
            double off = 0.0, dg = 0.0;
            for (int i = 0; i < n; i++)
                for (int j = 0; j < n; j++) {
                    if (i != j) off += a[i * n + j] * a[i * n + j];
                    else dg += a[i * n + j] * a[i * n + j];
                }
            if (off <= 1e-32 * dg || off == 0.0) break;
            for (int p = 0; p < n - 1; p++)
                for (int q = p + 1; q < n; q++) {
                    const double apq = a[p * n + q];
                    if (apq == 0.0) continue;
                    const double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
                    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                    const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                    for (int k = 0; k < n; k++) {
                        const double akp = a[k * n + p], akq = a[k * n + q];
                        a[k * n + p] = c * akp - s * akq;
                        a[k * n + q] = s * akp + c * akq;
                    }
                    for (int k = 0; k < n; k++) {
                        const double apk = a[p * n + k], aqk = a[q * n + k];
                        a[p * n + k] = c * apk - s * aqk;
                        a[q * n + k] = s * apk + c * aqk;
                    }
                }
        }
    }
    double mn = a[0];
    for (int i = 1; i < n; i++) mn = fmin(mn, a[i * n + i]);
    for (int i = 0; i < n; i++)
        if (!(a[i * n + i] == a[i * n + i])) mn = a[i * n + i];
    return mn;
}

// Fast certificate for "all eigenvalues of the symmetric matrix (lower triangle of M) exceed -shift": the LDL'
// factorisation of M + shift I runs through with positive pivots.  Backward error analysis: the computed factors are
// exact for M + shift I + E with |E| <= gamma_{n+1} |L||D||L'|, i.e. ||E|| <= ~n^3 u max|M| <= 3e-11 for n <= 16 and
// max|M| <= 64, so a pass proves lambda_min(M) > -shift - 3e-11.  The callers choose shifts 0.5e-10 inside the
// reference's thresholds (riccati_lqr.py:57-116: Q >= -1e-10, R > 1e-10), so a pass implies the reference's eigvalsh test
// passes; anything else — a non-positive pivot, a large norm, a non-finite entry — returns false and the caller decides
// with the eigenvalues themselves (min_eig_sym), exactly as before.  ~n^3 / 3 flops instead of several Jacobi sweeps of
// ~4 n^3 each: for dense 12 x 12 weights the validation took three times as long as the solve (profiles/r02_t launch list).
__device__ bool ldl_certifies(const double *M, int n, double shift) {
    double a[QT_MAX_N * QT_MAX_N], dk[QT_MAX_N];
    double mx = 0.0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            const double v = M[i * n + j];
            a[i * n + j] = v + (i == j ? shift : 0.0);
            mx = fmax(mx, fabs(v));
            if (!(fabs(v) <= 64.0)) return false;   // large, inf or NaN: the eigenvalue path decides
        }
    for (int j = 0; j < n; j++) {
        double d = a[j * n + j];
        for (int k = 0; k < j; k++) d = fma(-a[j * n + k] * a[j * n + k], dk[k], d);
        if (!(d > 0.0)) return false;
        dk[j] = d;
        const double inv = 1.0 / d;
        for (int i = j + 1; i < n; i++) {
            double v = a[i * n + j];
            for (int k = 0; k < j; k++) v = fma(-a[i * n + k] * a[j * n + k], dk[k], v);
            a[i * n + j] = v * inv;
        }
    }
    return true;
}

__device__ bool allclose_sym(const double *M, int n) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            const double x = M[i * n + j], y = M[j * n + i];
            if (!(fabs(x - y) <= 1e-8 + 1e-5 * fabs(y))) return false;  // also rejects NaN
        }
    return true;
}

// only_marked: the decoupled path ran first and left the systems it did not take with info = kDareMarkDense.
// Accepted systems are handed to the dense kernel with info = kDareMarkValid.
__global__ void dare_validate_kernel(const __grid_constant__ DareArgs D, int only_marked) {
    const int64_t sys = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sys >= D.batch) return;
    if (only_marked && D.info[sys] != kDareMarkDense) return;
    // shared Q / R (stride 0) need to be checked only once, but the check is cheap: keep it simple
    const double *Q = D.Q + sys * D.sQ, *R = D.R + sys * D.sR;
    bool ok = allclose_sym(Q, D.n) && allclose_sym(R, D.m);
    if (ok) {
        // eigenvalues of Q >= -1e-10 and of R > 1e-10: certified by an LDL' factorisation where that is conclusive
        // (lambda_min(Q) > -0.8e-10, lambda_min(R) > 1.2e-10), else decided by the eigenvalues themselves
        bool q_ok = ldl_certifies(Q, D.n, 0.5e-10);
        if (!q_ok) {
            const double qmin = min_eig_sym(Q, D.n, D.n);
            q_ok = !(qmin < -1e-10) && (qmin == qmin);
        }
        bool r_ok = q_ok && ldl_certifies(R, D.m, -1.5e-10);
        if (q_ok && !r_ok) r_ok = min_eig_sym(R, D.m, D.m) > 1e-10;
        ok = q_ok && r_ok;
    }
    D.info[sys] = ok ? kDareMarkValid : 3;
    D.iters[sys] = 0;
}

template <int N, int R>
static cudaError_t launch_dare(const DareArgs &D, cudaStream_t st, int64_t max_grid) {
    using L = Lay<N, R>;
    constexpr int WARPS = DareCfg<N, R>::WARPS;
    const size_t smem = (size_t)WARPS * L::SLOTS * L::BLOCK * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(dare_kernel<N, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t per_block = (int64_t)WARPS * L::TEAMS;
    const int64_t blocks = (D.batch + per_block - 1) / per_block;
    const int64_t grid = (max_grid > 0 && blocks > max_grid) ? max_grid : blocks;
    dare_kernel<N, R><<<(unsigned)grid, WARPS * 32, smem, st>>>(D, blocks);
    return cudaGetLastError();
}

// riccati_lqr.py:187-316 and 602-700 for a batch of diagonal weights
__global__ void build_lqr_systems_kernel(int n, int64_t batch, double dt, double mass, double gravity,
                                         const double *__restrict__ q_diag, const double *__restrict__ r_diag,
                                         double *__restrict__ A, double *__restrict__ B,
                                         double *__restrict__ Q, double *__restrict__ R) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        for (int r = 0; r < n; r++)
            for (int c = 0; c < n; c++) A[r * n + c] = (r == c) ? 1.0 : 0.0;
        for (int k = 0; k < 3; k++) A[k * n + 3 + k] = 1.0 * dt;  // A_d = I + A_c dt
        for (int r = 0; r < n; r++)
            for (int c = 0; c < 4; c++) B[r * 4 + c] = 0.0;
        B[5 * 4 + 0] = (1.0 / mass) * dt;
        B[4 * 4 + 1] = -gravity * dt;
        B[3 * 4 + 2] = gravity * dt;
        if (n == 9)
            for (int k = 0; k < 3; k++) A[(6 + k) * n + k] = 1.0 * dt;  // integral rows: C dt
    }
    if (i >= batch) return;
    double *Qi = Q + i * n * n, *Ri = R + i * 16;
    for (int r = 0; r < n; r++)
        for (int c = 0; c < n; c++) Qi[r * n + c] = (r == c) ? q_diag[i * n + r] : 0.0;
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) Ri[r * 4 + c] = (r == c) ? r_diag[i * 4 + r] : 0.0;
}

}  // namespace qt

using namespace qt;

extern "C" int qt_dare_batched(qt_context *ctx, int32_t n, int32_t m, int64_t batch,
                               const double *A_dev, int64_t strideA, const double *B_dev,
                               int64_t strideB, const double *Q_dev, int64_t strideQ,
                               const double *R_dev, int64_t strideR, double *P_out_dev,
                               double *K_out_dev, int32_t *iters_out_dev, int32_t *info_out_dev) {
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, n >= 1 && n <= QT_MAX_N && m >= 1 && m <= QT_MAX_M, "n must be 1..16, m 1..8");
    QT_REQUIRE(ctx, batch >= 0, "batch < 0");
    QT_REQUIRE(ctx, A_dev && B_dev && Q_dev && R_dev && iters_out_dev && info_out_dev, "null pointer");
    if (batch == 0) return QT_OK;
    DareArgs D{n, m, batch, A_dev, B_dev, Q_dev, R_dev, strideA, strideB, strideQ, strideR,
               P_out_dev, K_out_dev, iters_out_dev, info_out_dev};
    DeviceGuard guard(ctx->device);
    cudaError_t err;
    int rc;
    // decoupled path first (qt_dare_blocks.cuh): systems that are block diagonal with blocks of at most three states
    // are solved there; the rest is marked and goes through validation + the dense kernel below
    const bool try_blocks = ctx->dare_path != QT_DARE_PATH_DENSE && dare_blocks_applicable(n, m);
    if (try_blocks) {
        if (n <= 6)
            err = launch_dare_blocks_n<6>(n, m, batch, A_dev, B_dev, Q_dev, R_dev, strideA, strideB, strideQ, strideR,
                                          P_out_dev, K_out_dev, iters_out_dev, info_out_dev, kDareMaxIter, kDareTol, ctx->stream);
        else if (n <= 9)
            err = launch_dare_blocks_n<9>(n, m, batch, A_dev, B_dev, Q_dev, R_dev, strideA, strideB, strideQ, strideR,
                                          P_out_dev, K_out_dev, iters_out_dev, info_out_dev, kDareMaxIter, kDareTol, ctx->stream);
        else
            err = launch_dare_blocks_n<12>(n, m, batch, A_dev, B_dev, Q_dev, R_dev, strideA, strideB, strideQ, strideR,
                                           P_out_dev, K_out_dev, iters_out_dev, info_out_dev, kDareMaxIter, kDareTol, ctx->stream);
        if (err != cudaSuccess) return fail(ctx, QT_ERR_CUDA, "dare_blocks_kernel launch", cudaGetErrorString(err));
        ctx->launches += 1;
    }
    dare_validate_kernel<<<(unsigned)((batch + 127) / 128), 128, 0, ctx->stream>>>(D, try_blocks ? 1 : 0);
    rc = check_launch(ctx, "dare_validate_kernel");
    if (rc) return rc;
    const int need = n > m ? n : m;
    // <padded size N, rows per lane R>: lanes per system N / R, systems per warp 32 / (N / R)
    // after the decoupled path most groups of systems have nothing left for the dense kernel: a few blocks per SM walk
    // the batch (grid-stride) instead of one block per group
    const int64_t max_grid = try_blocks ? (int64_t)ctx->sm_count * 4 : 0;
    if (need <= 4) err = launch_dare<4, 2>(D, ctx->stream, max_grid);          // 2 lanes, 16 systems per warp
    else if (need <= 6) err = launch_dare<6, QT_DARE_R6>(D, ctx->stream, max_grid);
    else if (need <= 9) err = launch_dare<QT_DARE_N9, QT_DARE_R9>(D, ctx->stream, max_grid);
    else if (need <= 12) err = launch_dare<12, QT_DARE_R12>(D, ctx->stream, max_grid);
    else err = launch_dare<16, QT_DARE_R16>(D, ctx->stream, max_grid);
    if (err != cudaSuccess) return fail(ctx, QT_ERR_CUDA, "dare_kernel launch", cudaGetErrorString(err));
    ctx->launches += 1;
    return QT_OK;
}

extern "C" int qt_build_lqr_systems(qt_context *ctx, int32_t n, int64_t batch, double dt, double mass,
                                    double gravity, const double *q_diag_dev, const double *r_diag_dev,
                                    double *A_out_dev, double *B_out_dev, double *Q_out_dev,
                                    double *R_out_dev) {
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, n == 6 || n == 9, "n must be 6 (LQR) or 9 (LQI)");
    QT_REQUIRE(ctx, batch >= 0 && q_diag_dev && r_diag_dev && A_out_dev && B_out_dev && Q_out_dev && R_out_dev,
               "bad argument");
    DeviceGuard guard(ctx->device);
    const int64_t threads = batch > 0 ? batch : 1;
    build_lqr_systems_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(
        n, batch, dt, mass, gravity, q_diag_dev, r_diag_dev, A_out_dev, B_out_dev, Q_out_dev, R_out_dev);
    return check_launch(ctx, "build_lqr_systems_kernel");
}
