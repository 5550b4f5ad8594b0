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
    qt_context *ctx = new (std::nothrow) qt_context();
    if (!ctx) return QT_ERR_ARG;
    memset(ctx, 0, sizeof(*ctx));
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
    DeviceGuard guard(ctx->device);
    // scratch layout: K (512 bytes reserved) | seeds | state0 [12][E] | tparams [12][E] | metrics [16][E]; every
    // region starts on a 16-byte boundary whatever the parity of E (the rollout reads K as 16-byte pairs)
    const size_t n = (size_t)E, n2 = (n + 1) & ~(size_t)1;
    const size_t off_K = 0, off_seeds = 512, off_state = off_seeds + n2 * 8, off_tp = off_state + 12 * n2 * 8,
                 off_met = off_tp + 12 * n2 * 8, total = off_met + 16 * n2 * 8;
    if (ctx->scratch_bytes < total) {
        if (ctx->scratch_dev) QT_CUDA_TRY(ctx, cudaFree(ctx->scratch_dev));
        ctx->scratch_dev = nullptr;
        ctx->scratch_bytes = 0;
        QT_CUDA_TRY(ctx, cudaMalloc(&ctx->scratch_dev, total));
        ctx->scratch_bytes = total;
    }
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
        const int64_t unit = 128 * 148;   // whole waves of blocks per chunk boundary are not needed, whole blocks are
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
