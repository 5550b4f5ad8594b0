// qt_api.cu — context management, defaults, the host-buffer Evaluator entry point and the FP64
// peak probe of libqtrack (include/qtrack.h).
#include <new>
#include <string.h>

#include "qt_common.cuh"

using namespace qt;

extern "C" int qt_abi_version(void) { return QT_ABI_VERSION; }
extern "C" int64_t qt_params_size(void) { return (int64_t)sizeof(qt_params); }

// Defaults of env/config.py:11-95, utils/metrics.py:26-39 and controllers/riccati_lqr.py:456-502.
extern "C" void qt_params_default(qt_params *p) {
    memset(p, 0, sizeof(*p));
    p->mass = 1.0;
    p->gravity = 9.81;
    p->drag_coeff_linear = 0.1;
    p->drag_coeff_angular = 0.01;
    p->max_thrust = 20.0;
    p->min_thrust = 0.0;
    p->max_angular_rate = 3.0;
    p->dt = 0.01;
    p->max_episode_time = 30.0;
    p->max_velocity = 50.0;
    p->max_angular_velocity = 10.0;
    p->max_position = 1000.0;
    p->target_speed = 1.0;
    p->target_amplitude = 2.0;
    p->target_frequency = 0.5;
    p->target_radius = 2.0;
    p->target_center[0] = 0.0;
    p->target_center[1] = 0.0;
    p->target_center[2] = 1.0;
    p->target_max_acceleration = 5.0;
    p->env_target_radius = 0.5;
    p->env_min_on_target_ratio = 0.8;
    p->env_min_episode_duration = 30.0;
    p->crit_target_radius = 0.5;
    p->crit_min_on_target_ratio = 0.8;
    p->crit_min_episode_duration = 30.0;
    p->ctrl_dt = 0.01;
    p->ctrl_hover_thrust = 1.0 * 9.81;
    p->ctrl_max_thrust = 20.0;
    p->ctrl_min_thrust = 0.0;
    p->ctrl_max_rate = 3.0;
    p->integral_limit = 10.0;
    p->integral_zero_threshold = 0.01;
    p->ff_max_velocity = 10.0;
    p->ff_max_acceleration = 5.0;
    p->motion_type = QT_MOTION_STATIONARY;
    p->integrator = QT_INTEGRATOR_RK4;
    p->use_lqi = 0;
    p->feedforward_enabled = 0;
    p->overshoot_window = 10;
    p->controller_type = QT_CTRL_RICCATI;
}

extern "C" int qt_create(int device, qt_context **out) {
    if (!out) return QT_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return QT_ERR_NO_DEVICE;
    if (device < 0 || device >= count) return QT_ERR_ARG;
    qt_context *ctx = new (std::nothrow) qt_context();   // value-initialised: every plain member is zero
    if (!ctx) return QT_ERR_ARG;
    ctx->device = device;
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return QT_ERR_CUDA;
    }
    if (prop.major < 10) {  // built for sm_100a only; no other code path exists
        delete ctx;
        return QT_ERR_UNSUPPORTED;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return QT_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    if (cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_inputs, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_aux_done, cudaEventDisableTiming) != cudaSuccess) {
        if (ctx->ev_inputs) cudaEventDestroy(ctx->ev_inputs);
        if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
        cudaStreamDestroy(ctx->own_stream);
        delete ctx;
        return QT_ERR_CUDA;
    }
    ctx->partials_cap = ctx->sm_count * 8;
    if (cudaMalloc(&ctx->partials_dev, (size_t)ctx->partials_cap * 16 * sizeof(double)) != cudaSuccess) {
        cudaStreamDestroy(ctx->own_stream);
        delete ctx;
        return QT_ERR_CUDA;
    }
    *out = ctx;
    return QT_OK;
}

extern "C" void qt_destroy(qt_context *ctx) {
    if (!ctx) return;
    {
        DeviceGuard guard(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->scratch_dev) cudaFree(ctx->scratch_dev);
        if (ctx->sweep_scratch_dev) cudaFree(ctx->sweep_scratch_dev);
        if (ctx->partials_dev) cudaFree(ctx->partials_dev);
        if (ctx->aux_stream) {
            cudaStreamSynchronize(ctx->aux_stream);
            cudaStreamDestroy(ctx->aux_stream);
        }
        if (ctx->ev_inputs) cudaEventDestroy(ctx->ev_inputs);
        if (ctx->ev_aux_done) cudaEventDestroy(ctx->ev_aux_done);
        if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    }
    delete ctx;
}

extern "C" int qt_set_dare_path(qt_context *ctx, int32_t path) {
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, path == QT_DARE_PATH_AUTO || path == QT_DARE_PATH_DENSE, "unknown DARE path");
    ctx->dare_path = path;
    return QT_OK;
}

extern "C" const char *qt_last_error(const qt_context *ctx) { return ctx ? ctx->err : "null context"; }

extern "C" int qt_set_stream(qt_context *ctx, void *cuda_stream) {
    if (!ctx) return QT_ERR_ARG;
    // NULL is a real stream: CUDA's legacy default stream, which is what torch uses unless told otherwise
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return QT_OK;
}

extern "C" void *qt_get_stream(const qt_context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int qt_synchronize(qt_context *ctx) {
    if (!ctx) return QT_ERR_ARG;
    DeviceGuard guard(ctx->device);
    QT_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return QT_OK;
}

extern "C" int64_t qt_launch_count(const qt_context *ctx) { return ctx ? ctx->launches : 0; }

// Evaluator.evaluate with host buffers (eval.py:169-231): seeds -> reset -> rollout -> metrics.
extern "C" int qt_evaluate_host(qt_context *ctx, const qt_params *params, const uint64_t *seeds_host,
                                int64_t E, int64_t max_steps, int64_t chain, const double *K_host,
                                int32_t k_cols, double *metrics_host) {
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, params && seeds_host && K_host && metrics_host, "null pointer");
    QT_REQUIRE(ctx, E >= 0, "E < 0");
    QT_REQUIRE(ctx, k_cols == (params->controller_type == QT_CTRL_PID ? 3 : (params->use_lqi ? 9 : 6)), "k_cols must be 6 (LQR), 9 (LQI) or 3 (PID)");
    if (E == 0) return QT_OK;
    std::lock_guard<std::mutex> lock(ctx->host_call_mutex);
    DeviceGuard guard(ctx->device);
    // scratch layout: K (512 bytes reserved) | seeds | state0 [12][E] | tparams [12][E] | metrics [16][E]; every
    // region starts on a 16-byte boundary whatever the parity of E (the rollout reads K as 16-byte pairs)
    const size_t n = (size_t)E, n2 = (n + 1) & ~(size_t)1;
    const size_t off_K = 0, off_seeds = 512, off_state = off_seeds + n2 * 8, off_tp = off_state + 12 * n2 * 8,
                 off_met = off_tp + 12 * n2 * 8, total = off_met + 16 * n2 * 8;
    if (int rc0 = ensure_scratch(ctx, total)) return rc0;
    char *base = reinterpret_cast<char *>(ctx->scratch_dev);
    uint64_t *seeds_dev = reinterpret_cast<uint64_t *>(base + off_seeds);
    double *state_dev = reinterpret_cast<double *>(base + off_state);
    double *tp_dev = reinterpret_cast<double *>(base + off_tp);
    double *met_dev = reinterpret_cast<double *>(base + off_met);
    double *K_dev = reinterpret_cast<double *>(base + off_K);
    QT_CUDA_TRY(ctx, cudaMemcpyAsync(seeds_dev, seeds_host, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    QT_CUDA_TRY(ctx, cudaMemcpyAsync(K_dev, K_host, 4 * (size_t)k_cols * 8, cudaMemcpyHostToDevice, ctx->stream));

    // Copy/compute pipeline.  The result is 128 bytes per episode (134 MB for 2^20 episodes, ~2.6 ms of
    // PCIe time after a 65 ms kernel): the episodes are cut into chunks of 1/2, 1/4, 1/8 and 1/8 that
    // alternate between two streams, so a chunk's read-back overlaps the next chunks' kernels and only
    // the last eighth's is exposed.  The kernels of both streams share the GPU (no wave-quantisation
    // loss from the smaller grids).  Chains of episodes that share a controller (chain > 1) and small
    // batches run as one chunk.
    int64_t cuts[5] = {0, E, E, E, E};
    int n_chunks = 1;
    if (chain <= 1 && E >= (1 << 17)) {
        const int64_t unit = 128 * (int64_t)ctx->sm_count;   // chunk boundaries on whole blocks (one block per SM each)
        cuts[1] = (E / 2) / unit * unit;
        cuts[2] = cuts[1] + (E / 4) / unit * unit;
        cuts[3] = cuts[2] + (E / 8) / unit * unit;
        cuts[4] = E;
        n_chunks = 4;
    }
    cudaStream_t user_stream = ctx->stream;
    cudaStream_t lanes[2] = {user_stream, ctx->aux_stream};
    int rc = QT_OK;
    if (n_chunks > 1) {
        QT_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_inputs, user_stream));
        QT_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_inputs, 0));
    }
    for (int c = 0; c < n_chunks && rc == QT_OK; c++) {
        const int64_t lo = cuts[c], nc = cuts[c + 1] - cuts[c];
        if (nc <= 0) continue;
        // chunk-local struct-of-arrays blocks inside the same scratch regions: [12][nc] at 12 * lo, etc.
        double *st = state_dev + 12 * lo, *tp = tp_dev + 12 * lo, *mt = met_dev + 16 * lo;
        ctx->stream = lanes[c & 1];
        rc = qt_reset_batched(ctx, params, seeds_dev + lo, nc, nullptr, st, tp);
        if (rc == QT_OK)
            rc = qt_rollout(ctx, params, nc, max_steps, chain, st, tp, K_dev, k_cols, nc, mt, nullptr, nullptr, nullptr, 1,
                            0, nullptr, nullptr, 1, 0, nullptr);
        if (rc == QT_OK) {
            cudaError_t e = cudaMemcpy2DAsync(metrics_host + lo, n * 8, mt, (size_t)nc * 8, (size_t)nc * 8, 16,
                                              cudaMemcpyDeviceToHost, ctx->stream);
            if (e != cudaSuccess) rc = fail(ctx, QT_ERR_CUDA, "cudaMemcpy2DAsync", cudaGetErrorString(e));
        }
    }
    ctx->stream = user_stream;
    if (n_chunks > 1) {   // the caller's stream is the one that is "done" when the call returns
        cudaEventRecord(ctx->ev_aux_done, ctx->aux_stream);
        cudaStreamWaitEvent(user_stream, ctx->ev_aux_done, 0);
    }
    cudaError_t es = cudaStreamSynchronize(user_stream);
    if (rc != QT_OK) return rc;
    if (es != cudaSuccess) return fail(ctx, QT_ERR_CUDA, "cudaStreamSynchronize", cudaGetErrorString(es));
    return QT_OK;
}

// solve_dare (riccati_lqr.py:119-184) for HOST matrices: H2D, validation + solve + gain, D2H.  The batch is cut into
// four chunks that alternate between the context's two streams so the copies of one chunk overlap the solves of
// its neighbours (inputs are 416 B and outputs 200 B per n = 6 system: at 65 536 systems the PCIe time is of the
// same order as the 0.5 ms of arithmetic).
extern "C" int qt_dare_host(qt_context *ctx, int32_t n, int32_t m, int64_t batch, const double *A_host,
                            int64_t strideA, const double *B_host, int64_t strideB, const double *Q_host,
                            int64_t strideQ, const double *R_host, int64_t strideR, double *P_out_host,
                            double *K_out_host, int32_t *iters_out_host, int32_t *info_out_host) {
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, n >= 1 && n <= QT_MAX_N && m >= 1 && m <= QT_MAX_M, "n must be 1..16, m 1..8");
    QT_REQUIRE(ctx, batch >= 0 && A_host && B_host && Q_host && R_host && info_out_host, "bad argument");
    QT_REQUIRE(ctx, (strideA == 0 || strideA == n * n) && (strideB == 0 || strideB == n * m) &&
                        (strideQ == 0 || strideQ == n * n) && (strideR == 0 || strideR == m * m),
               "host strides must be 0 (shared) or dense");
    if (batch == 0) return QT_OK;
    std::lock_guard<std::mutex> lock(ctx->host_call_mutex);
    DeviceGuard guard(ctx->device);
    const size_t nb = (size_t)batch, nn = (size_t)n * n, nm = (size_t)n * m, mm = (size_t)m * m;
    auto count = [&](int64_t stride, size_t each) { return (stride == 0 ? 1 : nb) * each; };
    size_t off = 0;
    auto take = [&](size_t doubles) { size_t o = off; off += ((doubles + 1) & ~(size_t)1) * 8; return o; };
    const size_t oA = take(count(strideA, nn)), oB = take(count(strideB, nm)), oQ = take(count(strideQ, nn)),
                 oR = take(count(strideR, mm)), oP = take(P_out_host ? nb * nn : 0), oK = take(nb * nm),
                 oI = take(nb), oF = take(nb);   // iters / info: int32, half of the reserved doubles
    if (int rc0 = ensure_scratch(ctx, off)) return rc0;
    char *base = reinterpret_cast<char *>(ctx->scratch_dev);
    double *dA = (double *)(base + oA), *dB = (double *)(base + oB), *dQ = (double *)(base + oQ), *dR = (double *)(base + oR);
    double *dP = P_out_host ? (double *)(base + oP) : nullptr, *dK = (double *)(base + oK);
    int32_t *dI = (int32_t *)(base + oI), *dF = (int32_t *)(base + oF);
    cudaStream_t user_stream = ctx->stream;
    cudaStream_t lanes[2] = {user_stream, ctx->aux_stream};
    const int n_chunks = batch >= 4096 ? 4 : 1;
    // shared (stride 0) matrices first, on the user's stream; the second lane waits for them
    if (strideA == 0) QT_CUDA_TRY(ctx, cudaMemcpyAsync(dA, A_host, nn * 8, cudaMemcpyHostToDevice, user_stream));
    if (strideB == 0) QT_CUDA_TRY(ctx, cudaMemcpyAsync(dB, B_host, nm * 8, cudaMemcpyHostToDevice, user_stream));
    if (strideQ == 0) QT_CUDA_TRY(ctx, cudaMemcpyAsync(dQ, Q_host, nn * 8, cudaMemcpyHostToDevice, user_stream));
    if (strideR == 0) QT_CUDA_TRY(ctx, cudaMemcpyAsync(dR, R_host, mm * 8, cudaMemcpyHostToDevice, user_stream));
    if (n_chunks > 1) {
        QT_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_inputs, user_stream));
        QT_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_inputs, 0));
    }
    int rc = QT_OK;
    for (int c = 0; c < n_chunks && rc == QT_OK; c++) {
        const int64_t lo = batch * c / n_chunks, hi = batch * (c + 1) / n_chunks, nc = hi - lo;
        if (nc <= 0) continue;
        cudaStream_t st = lanes[c & 1];
        ctx->stream = st;
        cudaError_t e = cudaSuccess;
        auto h2d = [&](double *d, const double *h, int64_t stride, size_t each) {
            if (stride != 0 && e == cudaSuccess)
                e = cudaMemcpyAsync(d + lo * each, h + lo * each, (size_t)nc * each * 8, cudaMemcpyHostToDevice, st);
        };
        h2d(dA, A_host, strideA, nn);
        h2d(dB, B_host, strideB, nm);
        h2d(dQ, Q_host, strideQ, nn);
        h2d(dR, R_host, strideR, mm);
        if (e != cudaSuccess) { rc = fail(ctx, QT_ERR_CUDA, "cudaMemcpyAsync (H2D)", cudaGetErrorString(e)); break; }
        rc = qt_dare_batched(ctx, n, m, nc, strideA ? dA + lo * nn : dA, strideA, strideB ? dB + lo * nm : dB, strideB,
                             strideQ ? dQ + lo * nn : dQ, strideQ, strideR ? dR + lo * mm : dR, strideR,
                             dP ? dP + lo * nn : nullptr, dK + lo * nm, dI + lo, dF + lo);
        if (rc != QT_OK) break;
        if (K_out_host) e = cudaMemcpyAsync(K_out_host + lo * nm, dK + lo * nm, (size_t)nc * nm * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && P_out_host)
            e = cudaMemcpyAsync(P_out_host + lo * nn, dP + lo * nn, (size_t)nc * nn * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && iters_out_host)
            e = cudaMemcpyAsync(iters_out_host + lo, dI + lo, (size_t)nc * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(info_out_host + lo, dF + lo, (size_t)nc * 4, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) rc = fail(ctx, QT_ERR_CUDA, "cudaMemcpyAsync (D2H)", cudaGetErrorString(e));
    }
    ctx->stream = user_stream;
    if (n_chunks > 1) {
        cudaEventRecord(ctx->ev_aux_done, ctx->aux_stream);
        cudaStreamWaitEvent(user_stream, ctx->ev_aux_done, 0);
    }
    cudaError_t es = cudaStreamSynchronize(user_stream);
    if (rc != QT_OK) return rc;
    if (es != cudaSuccess) return fail(ctx, QT_ERR_CUDA, "cudaStreamSynchronize", cudaGetErrorString(es));
    return QT_OK;
}

// ControllerTuner's candidate batch end to end with HOST buffers (controllers/tuning.py:846-928 for C candidates):
// diagonal weights in -> design matrices -> C DARE solves -> eps episodes per candidate (seeds seed0 + ep, the same for
// every candidate, tuning.py:875) -> scores out.  One H2D of (n + 4) doubles per candidate, one D2H of 4 doubles + info.
extern "C" int qt_sweep_host(qt_context *ctx, const qt_params *params, int32_t n, int64_t C, int64_t eps,
                             uint64_t seed0, int64_t max_steps, const double *q_host, const double *r_host,
                             double ctrl_mass, double ctrl_gravity, double *cand_out_host, double *K_out_host,
                             int32_t *info_out_host) {
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, params && q_host && r_host && cand_out_host && info_out_host, "null pointer");
    QT_REQUIRE(ctx, (n == 6 && !params->use_lqi) || (n == 9 && params->use_lqi), "n must be 6 (LQR) or 9 (LQI, use_lqi set)");
    QT_REQUIRE(ctx, params->controller_type == QT_CTRL_RICCATI, "controller_type must be QT_CTRL_RICCATI");
    QT_REQUIRE(ctx, C >= 0 && eps >= 1 && eps <= (1 << 20), "bad C / eps");
    if (C == 0) return QT_OK;
    std::lock_guard<std::mutex> lock(ctx->host_call_mutex);
    DeviceGuard guard(ctx->device);
    const size_t nc = (size_t)C, ne = (size_t)eps, nn = (size_t)n * n;
    size_t off = 0;
    auto take = [&](size_t doubles) { size_t o = off; off += ((doubles + 1) & ~(size_t)1) * 8; return o; };
    const size_t o_q = take(nc * n), o_r = take(nc * 4), oA = take(nn), oB = take((size_t)n * 4), oQ = take(nc * nn),
                 oR = take(nc * 16), oK = take(nc * 4 * n), oI = take(nc), oF = take(nc), o_seed = take(ne),
                 o_s0 = take(12 * ne), o_tp = take(12 * ne), o_cand = take(4 * nc);
    if (int rc0 = ensure_scratch(ctx, off)) return rc0;
    char *base = reinterpret_cast<char *>(ctx->scratch_dev);
    auto D = [&](size_t o) { return reinterpret_cast<double *>(base + o); };
    cudaStream_t st = ctx->stream;
    QT_CUDA_TRY(ctx, cudaMemcpyAsync(D(o_q), q_host, nc * n * 8, cudaMemcpyHostToDevice, st));
    QT_CUDA_TRY(ctx, cudaMemcpyAsync(D(o_r), r_host, nc * 4 * 8, cudaMemcpyHostToDevice, st));
    int rc = qt_build_lqr_systems(ctx, n, C, params->ctrl_dt, ctrl_mass, ctrl_gravity, D(o_q), D(o_r), D(oA), D(oB), D(oQ), D(oR));
    if (rc) return rc;
    int32_t *dI = reinterpret_cast<int32_t *>(base + oI), *dF = reinterpret_cast<int32_t *>(base + oF);
    rc = qt_dare_batched(ctx, n, 4, C, D(oA), 0, D(oB), 0, D(oQ), (int64_t)nn, D(oR), 16, nullptr, D(oK), dI, dF);
    if (rc) return rc;
    uint64_t *seeds = reinterpret_cast<uint64_t *>(base + o_seed);
    rc = qt_fill_seeds(ctx, seeds, eps, seed0);
    if (rc) return rc;
    rc = qt_reset_batched(ctx, params, seeds, eps, nullptr, D(o_s0), D(o_tp));
    if (rc) return rc;
    rc = qt_rollout_sweep(ctx, params, C, eps, max_steps, D(o_s0), D(o_tp), D(oK), n, D(o_cand), nullptr, nullptr);
    if (rc) return rc;
    QT_CUDA_TRY(ctx, cudaMemcpyAsync(cand_out_host, D(o_cand), 4 * nc * 8, cudaMemcpyDeviceToHost, st));
    QT_CUDA_TRY(ctx, cudaMemcpyAsync(info_out_host, dF, nc * 4, cudaMemcpyDeviceToHost, st));
    if (K_out_host) QT_CUDA_TRY(ctx, cudaMemcpyAsync(K_out_host, D(oK), nc * 4 * n * 8, cudaMemcpyDeviceToHost, st));
    QT_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return QT_OK;
}

// ---- FP64 peak probe ----------------------------------------------------------------------
// 8 independent DFMA chains per thread, operands in registers: the rate this reaches is the
// denominator of the rollout / DARE rooflines (MEASURED_PEAKS.json carries no FP64 figure).
namespace qt {
constexpr int kProbeChains = 8;
__global__ void __launch_bounds__(256) fp64_probe_kernel(int64_t iters, double seed, double *sink) {
    double acc[kProbeChains];
#pragma unroll
    for (int i = 0; i < kProbeChains; i++) acc[i] = seed + (double)(threadIdx.x + i);
    const double a = 1.0000000001, b = 1e-9;
    for (int64_t k = 0; k < iters; k++) {
#pragma unroll
        for (int i = 0; i < kProbeChains; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kProbeChains; i++) s += acc[i];
    if (s == 123.456) sink[0] = s;  // never true; keeps the loop alive
}
}  // namespace qt

// The same for the FP32 pipe (denominator of the policy kernel's roofline): 16 independent FFMA chains per thread.
namespace qt {
constexpr int kProbeChains32 = 16;
__global__ void __launch_bounds__(256) fp32_probe_kernel(int64_t iters, float seed, float *sink) {
    float acc[kProbeChains32];
#pragma unroll
    for (int i = 0; i < kProbeChains32; i++) acc[i] = seed + (float)(threadIdx.x + i);
    const float a = 1.0000001f, b = 1e-7f;
    for (int64_t k = 0; k < iters; k++) {
#pragma unroll
        for (int i = 0; i < kProbeChains32; i++) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kProbeChains32; i++) s += acc[i];
    if (s == 123.456f) sink[0] = s;  // never true; keeps the loop alive
}
}  // namespace qt

extern "C" int qt_fp32_peak_probe(qt_context *ctx, int64_t iters, double *tflops_out_host, double *ms_out_host) {
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, iters > 0 && tflops_out_host, "bad argument");
    DeviceGuard guard(ctx->device);
    const int threads = 256, blocks = ctx->sm_count * 8;
    cudaEvent_t e0, e1;
    QT_CUDA_TRY(ctx, cudaEventCreate(&e0));
    QT_CUDA_TRY(ctx, cudaEventCreate(&e1));
    float *sink = reinterpret_cast<float *>(ctx->partials_dev);
    fp32_probe_kernel<<<blocks, threads, 0, ctx->stream>>>(iters / 8 + 1, 1.0f, sink);  // warm-up
    QT_CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    fp32_probe_kernel<<<blocks, threads, 0, ctx->stream>>>(iters, 1.0f, sink);
    QT_CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
    QT_CUDA_TRY(ctx, cudaEventSynchronize(e1));
    int rc = check_launch(ctx, "fp32_probe_kernel");
    ctx->launches += 1;
    float ms = 0.f;
    QT_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double flops = 2.0 * (double)kProbeChains32 * (double)iters * (double)threads * (double)blocks;
    *tflops_out_host = flops / ((double)ms * 1e-3) / 1e12;
    if (ms_out_host) *ms_out_host = (double)ms;
    return rc;
}

extern "C" int qt_fp64_peak_probe(qt_context *ctx, int64_t iters, double *tflops_out_host,
                                  double *ms_out_host) {
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, iters > 0 && tflops_out_host, "bad argument");
    DeviceGuard guard(ctx->device);
    const int threads = 256, blocks = ctx->sm_count * 8;
    cudaEvent_t e0, e1;
    QT_CUDA_TRY(ctx, cudaEventCreate(&e0));
    QT_CUDA_TRY(ctx, cudaEventCreate(&e1));
    double *sink = reinterpret_cast<double *>(ctx->partials_dev);
    fp64_probe_kernel<<<blocks, threads, 0, ctx->stream>>>(iters / 8 + 1, 1.0, sink);  // warm-up
    QT_CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    fp64_probe_kernel<<<blocks, threads, 0, ctx->stream>>>(iters, 1.0, sink);
    QT_CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
    QT_CUDA_TRY(ctx, cudaEventSynchronize(e1));
    int rc = check_launch(ctx, "fp64_probe_kernel");
    ctx->launches += 1;
    float ms = 0.f;
    QT_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double flops = 2.0 * (double)kProbeChains * (double)iters * (double)threads * (double)blocks;
    *tflops_out_host = flops / ((double)ms * 1e-3) / 1e12;
    if (ms_out_host) *ms_out_host = (double)ms;
    return rc;
}
