// qt_env.cu — the environment's reset / observe / step and the controller's compute_action as
// stand-alone batched kernels.  These keep the
// reference's step-at-a-time interface alive for callers that bring their own policy
// (BatchedQuadcopterEnv.step(actions)); the fused loop lives in qt_rollout.cu and shares every
// device function with this file through qt_step.cuh.
#include "qt_common.cuh"
#include "qt_rng.cuh"
#include "qt_step.cuh"

namespace qt {

constexpr int kEnvThreads = 128;

// ------------------------------------------------------------------------------ reset -----
// env/quadcopter_env.py:111-150 + env/target_motion.py:291-369.  With reset(seed) the
// reference builds the env generator and the target generator from the SAME seed, so both
// streams are identical: the target pattern consumes its draws from the start of the stream
// and the position offset is the first three uniform doubles of an identical stream.
__global__ void __launch_bounds__(kEnvThreads)
    reset_kernel(const __grid_constant__ qt_params P, const uint64_t *__restrict__ seeds, int64_t E,
                 uint64_t *__restrict__ draws, double *__restrict__ state, double *__restrict__ tparams) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const uint64_t seed = seeds[e];
    Pcg64 g_env, g_tgt;
    pcg64_seed(g_env, seed);
    g_tgt = g_env;
    if (draws) {  // reset(seed=None): resume both streams where the previous reset left them
        pcg64_advance(g_env, draws[e]);
        pcg64_advance(g_tgt, draws[E + e]);
    }

    double tp[QT_TARGET_PARAMS];
#pragma unroll
    for (int i = 0; i < QT_TARGET_PARAMS; i++) tp[i] = 0.0;
    tp[0] = P.target_center[0];
    tp[1] = P.target_center[1];
    tp[2] = P.target_center[2];
    double p0[3] = {tp[0], tp[1], tp[2]};  // target position at t = 0

    switch (P.motion_type) {
    case QT_MOTION_LINEAR: {
        // direction = rng.standard_normal(3); direction /= norm   (target_motion.py:320-321)
        // LinearMotion.__init__ normalises once more and scales by speed (:47-49)
        double d0 = rng_standard_normal(g_tgt), d1 = rng_standard_normal(g_tgt),
               d2 = rng_standard_normal(g_tgt);
        // np.linalg.norm(x) = sqrt(ddot(x, x)); OpenBLAS evaluates the 3-term dot as this FMA
        // chain (pinned against numpy in tests/test_rng.py and oracle/README.md)
        double n1 = sqrt(__fma_rn(d2, d2, __fma_rn(d1, d1, __dmul_rn(d0, d0))));
        d0 = d0 / n1;
        d1 = d1 / n1;
        d2 = d2 / n1;
        double n2 = sqrt(__fma_rn(d2, d2, __fma_rn(d1, d1, __dmul_rn(d0, d0))));
        tp[3] = (d0 / n2) * P.target_speed;
        tp[4] = (d1 / n2) * P.target_speed;
        tp[5] = (d2 / n2) * P.target_speed;
        // position(0) = start + velocity * 0.0
        p0[0] = tp[0] + tp[3] * 0.0;
        p0[1] = tp[1] + tp[4] * 0.0;
        p0[2] = tp[2] + tp[5] * 0.0;
        break;
    }
    case QT_MOTION_CIRCULAR: {
        const double angle0 = rng_uniform(g_tgt, 0.0, kTwoPi);  // uniform(0, 2*math.pi), :329
        tp[3] = P.target_radius;
        tp[4] = P.target_speed / P.target_radius;
        tp[5] = angle0;
        const double angle = angle0 + tp[4] * 0.0;
        p0[0] = tp[0] + tp[3] * cos(angle);
        p0[1] = tp[1] + tp[3] * sin(angle);
        break;
    }
    case QT_MOTION_SINUSOIDAL: {
        // amplitude * [1, .5, .25], frequency * [1, 1.3, .7], phase = uniform(0, 2pi, 3)  (:339-353)
        const double amp[3] = {P.target_amplitude, P.target_amplitude * 0.5, P.target_amplitude * 0.25};
        const double frq[3] = {P.target_frequency, P.target_frequency * 1.3, P.target_frequency * 0.7};
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double phase = rng_uniform(g_tgt, 0.0, kTwoPi);
            tp[3 + i] = amp[i];
            tp[6 + i] = kTwoPi * frq[i];  // 2 * np.pi * frequency  (:139)
            tp[9 + i] = phase;
            const double theta = tp[6 + i] * 0.0 + phase;
            p0[i] = tp[i] + amp[i] * sin(theta);
        }
        break;
    }
    case QT_MOTION_FIGURE8: {
        tp[3] = P.target_amplitude;                   // scale  (:361-366)
        tp[4] = P.target_speed / P.target_amplitude;  // omega  (:172)
        const double theta = tp[4] * 0.0;
        const double c = cos(theta), s = sin(theta);
        const double denom = 1.0 + s * s;
        p0[0] = tp[0] + tp[3] * c / denom;
        p0[1] = tp[1] + tp[3] * s * c / denom;
        break;
    }
    default: break;
    }

    // offset = rng.uniform(-0.5, 0.5, 3)   (quadcopter_env.py:137-139)
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double off = rng_uniform(g_env, -0.5, 0.5);
        state[(int64_t)i * E + e] = p0[i] + off;
    }
#pragma unroll
    for (int i = 3; i < 12; i++) state[(int64_t)i * E + e] = 0.0;
#pragma unroll
    for (int i = 0; i < QT_TARGET_PARAMS; i++) tparams[(int64_t)i * E + e] = tp[i];
    if (draws) {
        draws[e] = g_env.draws;
        draws[E + e] = g_tgt.draws;
    }
}

__global__ void __launch_bounds__(kEnvThreads)
    rng_probe_kernel(const uint64_t *__restrict__ seeds, int64_t E, int n_uniform, int n_normal,
                     double *__restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    Pcg64 g;
    pcg64_seed(g, seeds[e]);
    for (int i = 0; i < n_uniform; i++) out[(int64_t)i * E + e] = pcg64_double(g);
    for (int i = 0; i < n_normal; i++) out[(int64_t)(n_uniform + i) * E + e] = rng_standard_normal(g);
}

// seeds[i] = seed0 + i: Evaluator.evaluate's base_seed + i (eval.py:200) / the tuner's seed + ep (tuning.py:875)
__global__ void __launch_bounds__(kEnvThreads) fill_seeds_kernel(uint64_t *__restrict__ seeds, int64_t E, uint64_t seed0) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) seeds[e] = seed0 + (uint64_t)e;
}

// ------------------------------------------------------------------------------ observe ---
template <int MOTION>
__device__ __forceinline__ void observe_one(const qt_params &P, int64_t E, int64_t e,
                                            const double *state, double t, const double *tparams,
                                            double *obs) {
    TargetGen<MOTION, 1> tg;
    double par[QT_TARGET_PARAMS];
    tg.load(par, tparams, E, e);
    double p[3], v[3], a[3];
    tg.template state<true>(P, t, p, v, a);
#pragma unroll
    for (int i = 0; i < 12; i++) obs[(int64_t)i * E + e] = state[(int64_t)i * E + e];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        obs[(int64_t)(12 + i) * E + e] = p[i];
        obs[(int64_t)(15 + i) * E + e] = v[i];
        obs[(int64_t)(18 + i) * E + e] = a[i];
    }
    obs[(int64_t)21 * E + e] = t;
}

__global__ void __launch_bounds__(kEnvThreads)
    observe_kernel(const __grid_constant__ qt_params P, int64_t E, const double *__restrict__ state,
                   const double *__restrict__ time, const double *__restrict__ tparams,
                   double *__restrict__ obs) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const double t = time[e];
    switch (P.motion_type) {
    case QT_MOTION_LINEAR: observe_one<QT_MOTION_LINEAR>(P, E, e, state, t, tparams, obs); break;
    case QT_MOTION_CIRCULAR: observe_one<QT_MOTION_CIRCULAR>(P, E, e, state, t, tparams, obs); break;
    case QT_MOTION_SINUSOIDAL: observe_one<QT_MOTION_SINUSOIDAL>(P, E, e, state, t, tparams, obs); break;
    case QT_MOTION_FIGURE8: observe_one<QT_MOTION_FIGURE8>(P, E, e, state, t, tparams, obs); break;
    default: observe_one<QT_MOTION_STATIONARY>(P, E, e, state, t, tparams, obs); break;
    }
}

// ------------------------------------------------------------------------------ controller
struct CtrlArgs {
    int64_t E, k_stride;
    const double *obs, *K;
    double *integral;
    int64_t *saturation;
    double *action;
};

// Controller state layout for the stand-alone calls: integral [3][E] (LQI) or [4][E] (PID: row 3 is
// the time of the previous call, NaN = none).
template <int CTRL>
__device__ __forceinline__ void run_controller(const qt_params &P, const double *K, int64_t E, int64_t e,
                                               const double (&s)[12], const double (&tp)[3], const double (&tv)[3],
                                               const double (&ta)[3], double t, double *integral,
                                               int64_t *saturation, double (&act)[4]) {
    constexpr int KC = CtrlCols<CTRL>::value;
    double integ[3] = {0.0, 0.0, 0.0};
    const bool stateful = (CTRL != CK_LQR) && integral;
    if (stateful) {
#pragma unroll
        for (int i = 0; i < 3; i++) integ[i] = integral[(int64_t)i * E + e];
    }
    if constexpr (CTRL == CK_PID) {
        double last_time = stateful ? integral[(int64_t)3 * E + e] : __longlong_as_double(0x7ff8000000000000LL);
        pid_action<true>(P, K, s, tp, tv, ta, t, integ, last_time, act);
        if (stateful) integral[(int64_t)3 * E + e] = last_time;
    } else {
        const double err = norm3d(tp[0] - s[0], tp[1] - s[1], tp[2] - s[2]);
        int sat = 0;
        ctrl_action<CTRL == CK_LQI, true>(P, K, s, tp, tv, ta, err, integ, sat, act);
        if (saturation) saturation[e] += sat;
    }
    if (stateful) {
#pragma unroll
        for (int i = 0; i < 3; i++) integral[(int64_t)i * E + e] = integ[i];
    }
    (void)KC;
}

template <int CTRL>
__global__ void __launch_bounds__(kEnvThreads)
    compute_action_kernel(const __grid_constant__ qt_params P, const __grid_constant__ CtrlArgs A) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t E = A.E;
    if (e >= E) return;
    constexpr int KC = CtrlCols<CTRL>::value;
    double s[12], tp[3], tv[3], ta[3];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = A.obs[(int64_t)i * E + e];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        tp[i] = A.obs[(int64_t)(12 + i) * E + e];
        tv[i] = A.obs[(int64_t)(15 + i) * E + e];
        ta[i] = A.obs[(int64_t)(18 + i) * E + e];
    }
    double act[4];
    run_controller<CTRL>(P, A.K + (e / A.k_stride) * 4 * KC, E, e, s, tp, tv, ta, A.obs[(int64_t)21 * E + e],
                         A.integral, A.saturation, act);
#pragma unroll
    for (int i = 0; i < 4; i++) A.action[(int64_t)i * E + e] = act[i];
}

// ------------------------------------------------------------------------------ step ------
struct StepArgs {
    int64_t E, k_stride;
    double *state, *time;
    const double *tparams;
    int64_t *counters;
    const double *action_in, *K;
    double *integral;
    int64_t *saturation;
    double *obs, *reward;
    int32_t *done;
    double *info, *action_out;
};

template <int MOTION, int CTRL>
__device__ __forceinline__ void step_one(const qt_params &P, const StepArgs &A, int64_t e) {
    const int64_t E = A.E;
    constexpr int KC = CtrlCols<CTRL>::value;
    double s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = A.state[(int64_t)i * E + e];
    TargetGen<MOTION, 1> tg;
    double par[QT_TARGET_PARAMS];
    tg.load(par, A.tparams, E, e);
    double t = A.time[e];
    double tp[3], tv[3], ta[3], act[4], actv[4];
    if (A.action_in) {
#pragma unroll
        for (int i = 0; i < 4; i++) act[i] = A.action_in[(int64_t)i * E + e];
    } else {
        tg.template state<true>(P, t, tp, tv, ta);
        run_controller<CTRL>(P, A.K + (e / A.k_stride) * 4 * KC, E, e, s, tp, tv, ta, t, A.integral, A.saturation, act);
    }
    const int violated = validate_action(P, act, actv);
    DynConsts C;
    C.init(P);
    if (P.integrator == QT_INTEGRATOR_EULER)
        integrate<true>(C, s, actv);
    else
        integrate<false>(C, s, actv);
    constrain_state(P, C, s);
    t += P.dt;
    tg.template state<true>(P, t, tp, tv, ta);
    const double err = norm3d(s[0] - tp[0], s[1] - tp[1], s[2] - tp[2]);
    const int term = termination_code(P, C, s, t);
    const bool on = err <= P.env_target_radius;
    const int64_t steps = A.counters[0 * E + e] + 1;
    const int64_t on_count = A.counters[1 * E + e] + (on ? 1 : 0);
    const int64_t viol = A.counters[2 * E + e] + violated;
    A.counters[0 * E + e] = steps;
    A.counters[1 * E + e] = on_count;
    A.counters[2 * E + e] = viol;
#pragma unroll
    for (int i = 0; i < 12; i++) A.state[(int64_t)i * E + e] = s[i];
    A.time[e] = t;
    if (A.obs) {
#pragma unroll
        for (int i = 0; i < 12; i++) A.obs[(int64_t)i * E + e] = s[i];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            A.obs[(int64_t)(12 + i) * E + e] = tp[i];
            A.obs[(int64_t)(15 + i) * E + e] = tv[i];
            A.obs[(int64_t)(18 + i) * E + e] = ta[i];
        }
        A.obs[(int64_t)21 * E + e] = t;
    }
    if (A.reward) A.reward[e] = -err;
    if (A.done) A.done[e] = term;
    if (A.info) {
        const double ratio = steps > 0 ? (double)on_count / (double)steps : 0.0;
        A.info[0 * E + e] = err;
        A.info[1 * E + e] = on ? 1.0 : 0.0;
        A.info[2 * E + e] = ratio;
        A.info[3 * E + e] =
            (t >= P.env_min_episode_duration && ratio >= P.env_min_on_target_ratio) ? 1.0 : 0.0;
    }
    if (A.action_out) {
#pragma unroll
        for (int i = 0; i < 4; i++) A.action_out[(int64_t)i * E + e] = actv[i];
    }
}

template <int CTRL>
__global__ void __launch_bounds__(kEnvThreads)
    step_kernel(const __grid_constant__ qt_params P, const __grid_constant__ StepArgs A) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A.E) return;
    switch (P.motion_type) {
    case QT_MOTION_LINEAR: step_one<QT_MOTION_LINEAR, CTRL>(P, A, e); break;
    case QT_MOTION_CIRCULAR: step_one<QT_MOTION_CIRCULAR, CTRL>(P, A, e); break;
    case QT_MOTION_SINUSOIDAL: step_one<QT_MOTION_SINUSOIDAL, CTRL>(P, A, e); break;
    case QT_MOTION_FIGURE8: step_one<QT_MOTION_FIGURE8, CTRL>(P, A, e); break;
    default: step_one<QT_MOTION_STATIONARY, CTRL>(P, A, e); break;
    }
}

static inline unsigned grid_for(int64_t E) { return (unsigned)((E + kEnvThreads - 1) / kEnvThreads); }

}  // namespace qt

using namespace qt;

extern "C" int qt_reset_batched(qt_context *ctx, const qt_params *params, const uint64_t *seeds_dev,
                                int64_t E, uint64_t *draws_io_dev, double *state_out_dev,
                                double *tparams_out_dev) {
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, params && seeds_dev && state_out_dev && tparams_out_dev, "null pointer");
    QT_REQUIRE(ctx, E >= 0, "E < 0");
    QT_REQUIRE(ctx, params->motion_type >= 0 && params->motion_type <= QT_MOTION_FIGURE8, "motion_type");
    if (E == 0) return QT_OK;
    DeviceGuard guard(ctx->device);
    reset_kernel<<<grid_for(E), kEnvThreads, 0, ctx->stream>>>(*params, seeds_dev, E, draws_io_dev,
                                                              state_out_dev, tparams_out_dev);
    return check_launch(ctx, "reset_kernel");
}

extern "C" int qt_rng_probe(qt_context *ctx, const uint64_t *seeds_dev, int64_t E, int32_t n_uniform,
                            int32_t n_normal, double *out_dev) {
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, seeds_dev && out_dev && E >= 0 && n_uniform >= 0 && n_normal >= 0, "bad argument");
    if (E == 0) return QT_OK;
    DeviceGuard guard(ctx->device);
    rng_probe_kernel<<<grid_for(E), kEnvThreads, 0, ctx->stream>>>(seeds_dev, E, n_uniform, n_normal, out_dev);
    return check_launch(ctx, "rng_probe_kernel");
}

extern "C" int qt_fill_seeds(qt_context *ctx, uint64_t *seeds_out_dev, int64_t E, uint64_t seed0) {
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, seeds_out_dev && E >= 0, "bad argument");
    if (E == 0) return QT_OK;
    DeviceGuard guard(ctx->device);
    fill_seeds_kernel<<<grid_for(E), kEnvThreads, 0, ctx->stream>>>(seeds_out_dev, E, seed0);
    return check_launch(ctx, "fill_seeds_kernel");
}

extern "C" int qt_observe_batched(qt_context *ctx, const qt_params *params, int64_t E,
                                  const double *state_dev, const double *time_dev,
                                  const double *tparams_dev, double *obs_out_dev) {
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, params && state_dev && time_dev && tparams_dev && obs_out_dev, "null pointer");
    QT_REQUIRE(ctx, E >= 0, "E < 0");
    if (E == 0) return QT_OK;
    DeviceGuard guard(ctx->device);
    observe_kernel<<<grid_for(E), kEnvThreads, 0, ctx->stream>>>(*params, E, state_dev, time_dev,
                                                                tparams_dev, obs_out_dev);
    return check_launch(ctx, "observe_kernel");
}

extern "C" int qt_compute_action_batched(qt_context *ctx, const qt_params *params, int64_t E,
                                         const double *obs_dev, const double *K_dev, int32_t k_cols,
                                         int64_t k_stride, double *integral_io_dev,
                                         int64_t *saturation_io_dev, double *action_out_dev) {
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, params && obs_dev && K_dev && action_out_dev, "null pointer");
    QT_REQUIRE(ctx, E >= 0, "E < 0");
    QT_REQUIRE(ctx, k_cols == expected_k_cols(*params), "k_cols must be 6 (LQR), 9 (LQI) or 3 (PID)");
    QT_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(K_dev) & 15) == 0, "K must be 16-byte aligned");
    if (E == 0) return QT_OK;
    if (k_stride < 1) k_stride = 1;
    CtrlArgs A{E, k_stride, obs_dev, K_dev, integral_io_dev, saturation_io_dev, action_out_dev};
    DeviceGuard guard(ctx->device);
    switch (ctrl_kind_of(*params)) {
    case CK_PID: compute_action_kernel<CK_PID><<<grid_for(E), kEnvThreads, 0, ctx->stream>>>(*params, A); break;
    case CK_LQI: compute_action_kernel<CK_LQI><<<grid_for(E), kEnvThreads, 0, ctx->stream>>>(*params, A); break;
    default: compute_action_kernel<CK_LQR><<<grid_for(E), kEnvThreads, 0, ctx->stream>>>(*params, A); break;
    }
    return check_launch(ctx, "compute_action_kernel");
}

extern "C" int qt_step_batched(qt_context *ctx, const qt_params *params, int64_t E,
                               double *state_io_dev, double *time_io_dev, const double *tparams_dev,
                               int64_t *counters_io_dev, const double *action_in_dev,
                               const double *K_dev, int32_t k_cols, int64_t k_stride,
                               double *integral_io_dev, int64_t *saturation_io_dev,
                               double *obs_out_dev, double *reward_out_dev, int32_t *done_out_dev,
                               double *info_out_dev, double *action_out_dev) {
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, params && state_io_dev && time_io_dev && tparams_dev && counters_io_dev, "null pointer");
    QT_REQUIRE(ctx, E >= 0, "E < 0");
    QT_REQUIRE(ctx, action_in_dev || K_dev, "either an action or a gain matrix is required");
    if (!action_in_dev) {
        QT_REQUIRE(ctx, k_cols == expected_k_cols(*params), "k_cols must be 6 (LQR), 9 (LQI) or 3 (PID)");
        QT_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(K_dev) & 15) == 0, "K must be 16-byte aligned");
    }
    QT_REQUIRE(ctx, params->motion_type >= 0 && params->motion_type <= QT_MOTION_FIGURE8, "motion_type");
    if (E == 0) return QT_OK;
    if (k_stride < 1) k_stride = 1;
    StepArgs A{E, k_stride, state_io_dev, time_io_dev, tparams_dev, counters_io_dev, action_in_dev,
               K_dev, integral_io_dev, saturation_io_dev, obs_out_dev, reward_out_dev, done_out_dev,
               info_out_dev, action_out_dev};
    DeviceGuard guard(ctx->device);
    const int kind = action_in_dev ? CK_LQR : ctrl_kind_of(*params);
    switch (kind) {
    case CK_PID: step_kernel<CK_PID><<<grid_for(E), kEnvThreads, 0, ctx->stream>>>(*params, A); break;
    case CK_LQI: step_kernel<CK_LQI><<<grid_for(E), kEnvThreads, 0, ctx->stream>>>(*params, A); break;
    default: step_kernel<CK_LQR><<<grid_for(E), kEnvThreads, 0, ctx->stream>>>(*params, A); break;
    }
    return check_launch(ctx, "step_kernel");
}

// ---- self-test hook: the paired square root of qt_step.cuh against IEEE sqrt (tests/test_trig.py's GPU twin)
namespace qt {
__global__ void sqrt_pair_probe_kernel(int64_t n, const double *a, const double *b, double *ra, double *rb) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x, y;
    sqrt_pair(a[i], b[i], x, y);
    ra[i] = x;
    rb[i] = y;
}
}  // namespace qt

extern "C" int qt_sqrt_pair_probe(qt_context *ctx, int64_t n, const double *a_dev, const double *b_dev,
                                  double *ra_out_dev, double *rb_out_dev) {
    using namespace qt;
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, a_dev && b_dev && ra_out_dev && rb_out_dev && n >= 0, "null pointer / n < 0");
    if (n == 0) return QT_OK;
    DeviceGuard guard(ctx->device);
    sqrt_pair_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, a_dev, b_dev, ra_out_dev, rb_out_dev);
    return check_launch(ctx, "sqrt_pair_probe_kernel");
}
