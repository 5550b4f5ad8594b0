// qt_reduce.cu — per-candidate and per-evaluation aggregation of the rollout's metric vectors
// (SURVEY.md §8 rows A13, A14).
//
//   qt_reduce_scores   ControllerTuner._evaluate_config, controllers/tuning.py:912-926
//   qt_reduce_summary  compute_evaluation_summary, utils/metrics.py:341-390
//
// Both read the [QT_METRICS_DIM][E] array qt_rollout wrote; nothing leaves the device until the
// caller copies the (tiny) result.  Sums over the episodes of one candidate follow numpy's
// add.reduce order exactly (pairwise blocks of 128 with 8 accumulators), so a
// candidate's score is bit-identical to the reference's whenever its per-episode metrics are.
// The evaluation summary is a deterministic two-level tree (no atomics); its sums are
// re-associated with respect to numpy, the tolerance is stated in DESIGN.md.
#include "qt_common.cuh"
#include "qt_npsum.cuh"

namespace qt {

__global__ void reduce_scores_kernel(const double *__restrict__ metrics, int64_t C, int64_t eps,
                                     double *__restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int64_t E = C * eps;
    const double *steps = metrics + (int64_t)QT_M_STEPS * E + c * eps;
    const double *on = metrics + (int64_t)QT_M_MET_ON_TARGET * E + c * eps;
    const double *serr = metrics + (int64_t)QT_M_SUM_ERR * E + c * eps;
    const double *succ = metrics + (int64_t)QT_M_MET_SUCCESS * E + c * eps;
    auto ratio_f = [=](int64_t k) { return steps[k] > 0.0 ? on[k] / steps[k] : 0.0; };
    auto merr_f = [=](int64_t k) { return steps[k] > 0.0 ? serr[k] / steps[k] : 0.0; };
    auto succ_f = [=](int64_t k) { return succ[k]; };
    const double n = (double)eps;
    const double mean_on = np_sum(ratio_f, eps) / n;
    const double mean_err = np_sum(merr_f, eps) / n;
    const double sr = np_sum(succ_f, eps) / n;
    out[0 * C + c] = __dadd_rn(mean_on, -__dmul_rn(0.1, mean_err));  // mean_on - 0.1 * mean_err, unfused
    out[1 * C + c] = mean_on;
    out[2 * C + c] = mean_err;
    out[3 * C + c] = sr;
}

// ---- evaluation summary -----------------------------------------------------------------
constexpr int kSumThreads = 256;
constexpr int kSumFields = 16;

struct Partial {
    double v[kSumFields];
};

__device__ __forceinline__ void combine(Partial &a, const Partial &b) {
    // additive fields
#pragma unroll
    for (int i = 0; i < kSumFields; i++) {
        if (i == 7 || i == 8 || i == 9 || i == 10) continue;
        a.v[i] += b.v[i];
    }
    // max ratio / first argmax, min ratio / first argmin (np.argmax / np.argmin semantics)
    if (b.v[7] > a.v[7] || (b.v[7] == a.v[7] && b.v[8] < a.v[8])) {
        a.v[7] = b.v[7];
        a.v[8] = b.v[8];
    }
    if (b.v[9] < a.v[9] || (b.v[9] == a.v[9] && b.v[10] < a.v[10])) {
        a.v[9] = b.v[9];
        a.v[10] = b.v[10];
    }
}

__device__ __forceinline__ Partial empty_partial() {
    Partial p;
#pragma unroll
    for (int i = 0; i < kSumFields; i++) p.v[i] = 0.0;
    p.v[7] = -INFINITY;
    p.v[8] = 1e300;
    p.v[9] = INFINITY;
    p.v[10] = 1e300;
    return p;
}

__device__ Partial block_reduce(Partial p) {
    __shared__ Partial sh[kSumThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Partial q;
#pragma unroll
        for (int i = 0; i < kSumFields; i++) q.v[i] = __shfl_down_sync(0xffffffffu, p.v[i], off);
        combine(p, q);
    }
    if (lane == 0) sh[warp] = p;
    __syncthreads();
    if (warp == 0) {
        p = lane < kSumThreads / 32 ? sh[lane] : empty_partial();
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            Partial q;
#pragma unroll
            for (int i = 0; i < kSumFields; i++) q.v[i] = __shfl_down_sync(0xffffffffu, p.v[i], off);
            combine(p, q);
        }
    }
    return p;  // valid in thread 0
}

// pass 0: plain sums and extrema; pass 1: centred second moments around the means in `sums`
template <int PASS>
__global__ void __launch_bounds__(kSumThreads)
    summary_partial_kernel(const double *__restrict__ metrics, int64_t E, const double *__restrict__ sums,
                           Partial *__restrict__ partials) {
    Partial p = empty_partial();
    double mean_ratio = 0.0, mean_err = 0.0;
    if (PASS == 1) {
        const double n = sums[0];
        mean_ratio = n > 0 ? sums[2] / n : 0.0;
        mean_err = n > 0 ? sums[3] / n : 0.0;
    }
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E;
         e += (int64_t)gridDim.x * blockDim.x) {
        const double steps = metrics[(int64_t)QT_M_STEPS * E + e];
        const double ratio = steps > 0 ? metrics[(int64_t)QT_M_MET_ON_TARGET * E + e] / steps : 0.0;
        const double merr = steps > 0 ? metrics[(int64_t)QT_M_SUM_ERR * E + e] / steps : 0.0;
        if (PASS == 0) {
            const double meff = steps > 0 ? metrics[(int64_t)QT_M_SUM_EFFORT * E + e] / steps : 0.0;
            const double env_ratio = steps > 0 ? metrics[(int64_t)QT_M_ENV_ON_TARGET * E + e] / steps : 0.0;
            Partial q = empty_partial();
            q.v[0] = 1.0;
            q.v[1] = metrics[(int64_t)QT_M_MET_SUCCESS * E + e];
            q.v[2] = ratio;
            q.v[3] = merr;
            q.v[4] = meff;
            q.v[7] = ratio;
            q.v[8] = (double)e;
            q.v[9] = ratio;
            q.v[10] = (double)e;
            q.v[11] = metrics[(int64_t)QT_M_ENV_SUCCESS * E + e];
            q.v[12] = env_ratio;
            q.v[13] = steps;
            q.v[14] = metrics[(int64_t)QT_M_SUM_ERR * E + e];
            q.v[15] = metrics[(int64_t)QT_M_VIOLATIONS * E + e];
            combine(p, q);
        } else {
            const double dr = ratio - mean_ratio, de = merr - mean_err;
            p.v[5] += dr * dr;
            p.v[6] += de * de;
        }
    }
    p = block_reduce(p);
    if (threadIdx.x == 0) partials[blockIdx.x] = p;
}

template <int PASS>
__global__ void __launch_bounds__(kSumThreads)
    summary_final_kernel(const Partial *__restrict__ partials, int n_partials, double *__restrict__ sums) {
    Partial p = empty_partial();
    for (int i = threadIdx.x; i < n_partials; i += blockDim.x) combine(p, partials[i]);
    p = block_reduce(p);
    if (threadIdx.x == 0) {
        if (PASS == 0) {
#pragma unroll
            for (int i = 0; i < kSumFields; i++) sums[i] = p.v[i];
            sums[5] = 0.0;
            sums[6] = 0.0;
        } else {
            sums[5] = p.v[5];
            sums[6] = p.v[6];
        }
    }
}

}  // namespace qt

using namespace qt;

extern "C" int qt_reduce_scores(qt_context *ctx, const double *metrics_dev, int64_t C, int64_t eps,
                                double *cand_out_dev) {
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, metrics_dev && cand_out_dev && C >= 0 && eps >= 1, "bad argument");
    if (C == 0) return QT_OK;
    DeviceGuard guard(ctx->device);
    reduce_scores_kernel<<<(unsigned)((C + 127) / 128), 128, 0, ctx->stream>>>(metrics_dev, C, eps, cand_out_dev);
    return check_launch(ctx, "reduce_scores_kernel");
}

extern "C" int qt_reduce_summary(qt_context *ctx, const double *metrics_dev, int64_t E,
                                 double *sums_out_dev) {
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, metrics_dev && sums_out_dev && E >= 0, "bad argument");
    DeviceGuard guard(ctx->device);
    int blocks = (int)((E + kSumThreads - 1) / kSumThreads);
    const int max_blocks = ctx->partials_cap;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    Partial *partials = reinterpret_cast<Partial *>(ctx->partials_dev);
    int rc;
    summary_partial_kernel<0><<<blocks, kSumThreads, 0, ctx->stream>>>(metrics_dev, E, sums_out_dev, partials);
    if ((rc = check_launch(ctx, "summary_partial_kernel<0>"))) return rc;
    summary_final_kernel<0><<<1, kSumThreads, 0, ctx->stream>>>(partials, blocks, sums_out_dev);
    if ((rc = check_launch(ctx, "summary_final_kernel<0>"))) return rc;
    summary_partial_kernel<1><<<blocks, kSumThreads, 0, ctx->stream>>>(metrics_dev, E, sums_out_dev, partials);
    if ((rc = check_launch(ctx, "summary_partial_kernel<1>"))) return rc;
    summary_final_kernel<1><<<1, kSumThreads, 0, ctx->stream>>>(partials, blocks, sums_out_dev);
    return check_launch(ctx, "summary_final_kernel<1>");
}
