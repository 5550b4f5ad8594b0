// qt_policy.cu — the deep tracking policy evaluated inside the closed loop.
//
// Replaces, for --controller deep and for imitation training,
//     features = controller._extract_features(obs); action = network(features); env.step(action)
// (controllers/deep_tracking_policy.py:199-286, eval.py:119-165, train.py:668-721) by one kernel
// launch for all episodes: one thread per episode, the env in fp64 exactly as in qt_rollout.cu (same
// device functions), the network in fp32 like torch.  The packed weights are staged in shared memory
// (they are read by every thread at the same address: broadcast loads).  The 18 features stay in
// registers for the first layer; hidden activations live in the thread's own column of a
// shared-memory tile (two tiles, alternating, only when there are three or more hidden layers); every
// chunk of 16 output units accumulates in registers, so one weight vector load feeds four FFMAs; and
// the last hidden layer is never stored — each chunk of it is multiplied into the four output
// accumulators as soon as it is activated.  For the default 64-64 network that leaves one 32 KB tile
// + 22 KB of weights per block, i.e. four resident blocks per SM.  This is batch-of-vectors work with
// fp32-accuracy requirements (the parity target is torch's fp32 Linear); it stays on the FP32 pipe.
#include "qt_common.cuh"
#include "qt_step.cuh"

namespace qt {

constexpr int kPolicyThreads = 128;

__host__ __device__ inline int pad4(int x) { return (x + 3) & ~3; }

struct MlpDesc {
    int n_layers;
    int rows[QT_MLP_MAX_LAYERS];     // input rows of layer l in the packed layout
    int outp[QT_MLP_MAX_LAYERS];     // padded output width of layer l
    int w_off[QT_MLP_MAX_LAYERS];    // offset (floats) of Wt of layer l; the bias follows Wt
    int activation;
    int total_floats;
    int tile_w[2];                   // widths (floats per thread) of the two activation tiles: hidden layer l
                                     // (0-based) is stored in tile l & 1; the last hidden layer is not stored
    float out_lo[4], out_scale[4];
};

static bool describe(const qt_mlp &m, MlpDesc &d) {
    if (m.n_layers < 1 || m.n_layers > QT_MLP_MAX_LAYERS) return false;
    if (m.dims[0] != QT_FEATURE_DIM || m.dims[m.n_layers] != QT_ACTION_DIM) return false;
    if (m.activation < QT_ACT_RELU || m.activation > QT_ACT_LEAKY_RELU) return false;
    d.n_layers = m.n_layers;
    d.activation = m.activation;
    int off = 0;
    d.tile_w[0] = d.tile_w[1] = 0;
    for (int l = 0; l < m.n_layers; l++) {
        if (m.dims[l + 1] < 1 || m.dims[l + 1] > QT_MLP_MAX_WIDTH) return false;
        d.rows[l] = l == 0 ? m.dims[0] : pad4(m.dims[l]);
        d.outp[l] = pad4(m.dims[l + 1]);
        d.w_off[l] = off;
        off += d.rows[l] * d.outp[l] + d.outp[l];
        if (l < m.n_layers - 2 && d.outp[l] > d.tile_w[l & 1]) d.tile_w[l & 1] = d.outp[l];  // stored hidden layers
    }
    d.total_floats = off;
    for (int i = 0; i < 4; i++) {
        d.out_lo[i] = m.out_lo[i];
        d.out_scale[i] = m.out_scale[i];
    }
    return true;
}

template <int CH>
__device__ __forceinline__ void chunk_bias(const float *__restrict__ bias, int j0, float (&acc)[CH]) {
#pragma unroll
    for (int i = 0; i < CH; i += 4) {
        const float4 b = *reinterpret_cast<const float4 *>(bias + j0 + i);
        acc[i] = b.x;
        acc[i + 1] = b.y;
        acc[i + 2] = b.z;
        acc[i + 3] = b.w;
    }
}

template <int CH>
__device__ __forceinline__ void chunk_fma(const float *__restrict__ w, float xk, float (&acc)[CH]) {
#pragma unroll
    for (int i = 0; i < CH; i += 4) {
        const float4 wk = *reinterpret_cast<const float4 *>(w + i);
        acc[i] = fmaf(xk, wk.x, acc[i]);
        acc[i + 1] = fmaf(xk, wk.y, acc[i + 1]);
        acc[i + 2] = fmaf(xk, wk.z, acc[i + 2]);
        acc[i + 3] = fmaf(xk, wk.w, acc[i + 3]);
    }
}

// One chunk of CH (16 or 4) output units of one layer for this thread; the layer input is this thread's
// column of a shared-memory tile (xin[k * T]).
template <int CH, int T>
__device__ __forceinline__ void mlp_chunk(const float *__restrict__ Wt, const float *__restrict__ bias, int rows,
                                          int outp, int j0, const float *xin, float (&acc)[CH]) {
    chunk_bias<CH>(bias, j0, acc);
    const float *w = Wt + j0;
#pragma unroll 4
    for (int k = 0; k < rows; k++) {
        chunk_fma<CH>(w, xin[k * T], acc);
        w += outp;
    }
}

// The same for the first layer: its 18 inputs are in registers.
template <int CH>
__device__ __forceinline__ void mlp_chunk_first(const float *__restrict__ Wt, const float *__restrict__ bias,
                                                int outp, int j0, const float (&feat)[QT_FEATURE_DIM], float (&acc)[CH]) {
    chunk_bias<CH>(bias, j0, acc);
    const float *w = Wt + j0;
#pragma unroll
    for (int k = 0; k < QT_FEATURE_DIM; k++) chunk_fma<CH>(w + k * outp, feat[k], acc);
}

template <int KIND>
__device__ __forceinline__ float activate_t(float x) {
    if (KIND == QT_ACT_TANH) return tanhf(x);
    if (KIND == QT_ACT_ELU) return x > 0.0f ? x : expm1f(x);
    if (KIND == QT_ACT_LEAKY_RELU) return x > 0.0f ? x : 0.1f * x;
    return x != x ? x : fmaxf(x, 0.0f);   // torch.relu propagates NaN, fmaxf does not
}

// Activate a finished chunk and either store it (hidden layer) or multiply it into the output
// accumulators (last hidden layer, fused output layer).  The activation kind is dispatched once per
// chunk, not per element.
template <int KIND, int CH, int T>
__device__ __forceinline__ void finish_chunk_t(const float (&acc)[CH], bool last_hidden, const float *__restrict__ Wo,
                                               int j0, float *xout, float (&o)[4]) {
    if (last_hidden) {
#pragma unroll
        for (int i = 0; i < CH; i++) chunk_fma<4>(Wo + (j0 + i) * 4, activate_t<KIND>(acc[i]), o);
    } else {
#pragma unroll
        for (int i = 0; i < CH; i++) xout[(j0 + i) * T] = activate_t<KIND>(acc[i]);
    }
}

template <int CH, int T>
__device__ __forceinline__ void finish_chunk(int kind, const float (&acc)[CH], bool last_hidden,
                                             const float *__restrict__ Wo, int j0, float *xout, float (&o)[4]) {
    switch (kind) {
    case QT_ACT_TANH: finish_chunk_t<QT_ACT_TANH, CH, T>(acc, last_hidden, Wo, j0, xout, o); break;
    case QT_ACT_ELU: finish_chunk_t<QT_ACT_ELU, CH, T>(acc, last_hidden, Wo, j0, xout, o); break;
    case QT_ACT_LEAKY_RELU: finish_chunk_t<QT_ACT_LEAKY_RELU, CH, T>(acc, last_hidden, Wo, j0, xout, o); break;
    default: finish_chunk_t<QT_ACT_RELU, CH, T>(acc, last_hidden, Wo, j0, xout, o); break;
    }
}

// PolicyNetwork.forward (deep_tracking_policy.py:114-147) for one feature vector held in registers.
// tile0 / tile1: this thread's columns of the two activation tiles.  W: packed weights (shared or global).
template <int T>
__device__ __forceinline__ void mlp_forward(const MlpDesc &M, const float *__restrict__ W, float *tile0, float *tile1,
                                            const float (&feat)[QT_FEATURE_DIM], float (&out)[4]) {
    const int L = M.n_layers;                      // linear layers; hidden layers are 0 .. L-2
    const float *Wo = W + M.w_off[L - 1];          // output layer: Wt [rows][4], bias [4]
    const int rows_o = M.rows[L - 1];
    float o[4];
    chunk_bias<4>(Wo + rows_o * 4, 0, o);
    if (L == 1) {                                  // no hidden layer: a linear policy
#pragma unroll
        for (int k = 0; k < QT_FEATURE_DIM; k++) chunk_fma<4>(Wo + k * 4, feat[k], o);
    }
    for (int l = 0; l + 1 < L; l++) {
        const float *Wt = W + M.w_off[l];
        const int rows = M.rows[l], outp = M.outp[l];
        const float *bias = Wt + rows * outp;
        const float *xin = ((l - 1) & 1) ? tile1 : tile0;
        float *xout = (l & 1) ? tile1 : tile0;
        const bool last_hidden = (l == L - 2);
        int j0 = 0;
        for (; j0 + 16 <= outp; j0 += 16) {
            float acc[16];
            if (l == 0)
                mlp_chunk_first<16>(Wt, bias, outp, j0, feat, acc);
            else
                mlp_chunk<16, T>(Wt, bias, rows, outp, j0, xin, acc);
            finish_chunk<16, T>(M.activation, acc, last_hidden, Wo, j0, xout, o);
        }
        for (; j0 < outp; j0 += 4) {
            float acc[4];
            if (l == 0)
                mlp_chunk_first<4>(Wt, bias, outp, j0, feat, acc);
            else
                mlp_chunk<4, T>(Wt, bias, rows, outp, j0, xin, acc);
            finish_chunk<4, T>(M.activation, acc, last_hidden, Wo, j0, xout, o);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float sg = 1.0f / (1.0f + expf(-o[i]));  // torch.sigmoid
        out[i] = fmaf(sg, M.out_scale[i], M.out_lo[i]);
    }
}

// deep_tracking_policy.py:241-286: fp64 differences first, then the fp32 cast
__device__ __forceinline__ void make_features(const double (&s)[12], const double (&tp)[3], const double (&tv)[3],
                                              float (&x)[QT_FEATURE_DIM]) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
        x[i] = (float)(tp[i] - s[i]);
        x[3 + i] = (float)(tv[i] - s[3 + i]);
        x[6 + i] = (float)s[6 + i];
        x[9 + i] = (float)s[9 + i];
        x[12 + i] = (float)tp[i];
        x[15 + i] = (float)tv[i];
    }
}

struct PolicyArgs {
    int64_t E, k_stride;
    int max_steps, w_in_smem, teacher;  // teacher: 0 none, else 1 + CtrlKind
    const double *state0, *tparams, *K;
    const float *W;
    double *metrics, *final_state, *traj;
    float *feat, *actf;
    int traj_stride, traj_slots, data_stride, data_slots;
};

// Stage the packed weights at the start of shared memory (when they fit).
__device__ __forceinline__ void stage_weights(const MlpDesc &M, const float *Wg, int in_smem, float *smem_w) {
    if (!in_smem) return;
    for (int i = threadIdx.x * 4; i < M.total_floats; i += blockDim.x * 4)
        *reinterpret_cast<float4 *>(smem_w + i) = *reinterpret_cast<const float4 *>(Wg + i);
    __syncthreads();
}

// MINB = resident blocks per SM the register allocator must allow.  Measured (2^18 episodes): with the
// weights in shared memory three blocks of 168 registers beat two of 255 (64-64 network 2.40e9 against
// 2.19e9 env-steps/s); when the weights are too large to stage and stream from L1/L2 instead, the deeper
// load prefetch that 255 registers allow wins (128-128: 4.0e8 against 2.4e8).
template <int MOTION, int MINB>
__global__ void __launch_bounds__(kPolicyThreads, MINB)
    rollout_policy_kernel(const __grid_constant__ qt_params P, const __grid_constant__ DynConsts C,
                          const __grid_constant__ MlpDesc M, const __grid_constant__ PolicyArgs A) {
    extern __shared__ __align__(16) float smem_f[];
    constexpr int T = kPolicyThreads;
    const int64_t E = A.E;
    const int64_t e = (int64_t)blockIdx.x * T + threadIdx.x;
    const int w_floats = A.w_in_smem ? ((M.total_floats + 3) & ~3) : 0;
    stage_weights(M, A.W, A.w_in_smem, smem_f);
    float *tile0 = smem_f + w_floats + threadIdx.x;
    float *tile1 = tile0 + M.tile_w[0] * T;
    if (e >= E) return;

    double s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = A.state0[(int64_t)i * E + e];
    double par[QT_TARGET_PARAMS];
    TargetGen<MOTION, 1> tg;
    tg.load(par, A.tparams, E, e);

    // supervisor state: reset before every episode (train.py:665-667)
    const int teacher_kind = A.teacher - 1;
    const int KC = teacher_kind == CK_LQI ? 9 : (teacher_kind == CK_PID ? 3 : 6);
    const double *Kp = A.teacher ? A.K + (e / A.k_stride) * 4 * KC : nullptr;
    double integ[3] = {0.0, 0.0, 0.0};
    double last_time = __longlong_as_double(0x7ff8000000000000LL);
    int saturations = 0;

    double t = 0.0;
    int step = 0, env_on = 0, met_on = 0, violations = 0, term = QT_TERM_NONE;
    double err_post = 0.0, sum_err = 0.0, sum_err2 = 0.0, sum_eff = 0.0, max_err = 0.0, ov_cur = 0.0, ov_max = 0.0;
    bool prev_on = false, in_phase = false;
    int streak = 0, ov_count = 0;
    int next_traj = 0, traj_slot = 0, next_data = 0, data_slot = 0;

    double tp[3], tv[3], ta[3];
    tg.template state<true>(P, t, tp, tv, ta);
    double err_pre = norm3d(tp[0] - s[0], tp[1] - s[1], tp[2] - s[2]);

    while (term == QT_TERM_NONE && step < A.max_steps) {
        // ---- the policy acts on the pre-step observation
        float feat[QT_FEATURE_DIM];
        make_features(s, tp, tv, feat);
        const bool log_data = A.feat && step == next_data && data_slot < A.data_slots;
        if (log_data) {
            float *f = A.feat + ((int64_t)data_slot * E + e) * QT_FEATURE_DIM;
#pragma unroll
            for (int i = 0; i < QT_FEATURE_DIM; i += 2) *reinterpret_cast<float2 *>(f + i) = make_float2(feat[i], feat[i + 1]);
        }
        float out[4];
        if (A.w_in_smem)   // two inlined copies: one reads the weights with LDS, the other with LDG
            mlp_forward<T>(M, smem_f, tile0, tile1, feat, out);
        else
            mlp_forward<T>(M, A.W, tile0, tile1, feat, out);
        double act[4] = {(double)out[0], (double)out[1], (double)out[2], (double)out[3]};

        // ---- supervisor label on the same observation (train.py:682-696)
        double lab[4] = {act[0], act[1], act[2], act[3]};
        if (A.teacher) {
            if (teacher_kind == CK_PID)
                pid_action<true>(P, Kp, s, tp, tv, ta, t, integ, last_time, lab);
            else if (teacher_kind == CK_LQI)
                ctrl_action<true, true>(P, Kp, s, tp, tv, ta, err_pre, integ, saturations, lab);
            else
                ctrl_action<false, true>(P, Kp, s, tp, tv, ta, err_pre, integ, saturations, lab);
        }

        // ---- metrics sample k = step (utils/metrics.py:264-338 streamed, as in qt_rollout.cu)
        {
            const bool on = err_pre <= P.crit_target_radius;
            met_on += on ? 1 : 0;
            sum_err += err_pre;
            sum_err2 += err_pre * err_pre;
            if (step == 0 || err_pre > max_err) max_err = err_pre;
            sum_eff += sqrt(act[0] * act[0] + act[1] * act[1] + act[2] * act[2] + act[3] * act[3]);
            if (step > 0) {
                if (prev_on && !on) {
                    in_phase = true;
                    streak = 1;
                    ov_cur = err_pre - P.crit_target_radius;
                } else if (in_phase && !on) {
                    streak += 1;
                    ov_cur = fmax(ov_cur, err_pre - P.crit_target_radius);
                } else if (in_phase && on) {
                    if (streak >= P.overshoot_window) {
                        ov_count += 1;
                        ov_max = fmax(ov_max, ov_cur);
                    }
                    in_phase = false;
                    streak = 0;
                    ov_cur = 0.0;
                }
            }
            prev_on = on;
        }

        bool log_traj = false;
        if (A.traj && step == next_traj && traj_slot < A.traj_slots) {
            log_traj = true;
            double *row = A.traj + (int64_t)traj_slot * QT_TRAJ_FIELDS * E + e;
#pragma unroll
            for (int i = 0; i < 12; i++) row[(int64_t)(QT_TRAJ_STATE + i) * E] = s[i];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                row[(int64_t)(QT_TRAJ_TARGET_P + i) * E] = tp[i];
                row[(int64_t)(QT_TRAJ_TARGET_V + i) * E] = tv[i];
            }
#pragma unroll
            for (int i = 0; i < 4; i++) row[(int64_t)(QT_TRAJ_ACTION + i) * E] = act[i];
        }
        if (log_data) {
            float2 *y = reinterpret_cast<float2 *>(A.actf + ((int64_t)data_slot * E + e) * QT_ACTION_DIM);
            y[0] = make_float2((float)lab[0], (float)lab[1]);
            y[1] = make_float2((float)lab[2], (float)lab[3]);
            next_data += A.data_stride;
            data_slot += 1;
        }

        // ---- env.step (quadcopter_env.py:171-207)
        double actv[4];
        violations += validate_action<true>(P, act, actv);
        if (P.integrator == QT_INTEGRATOR_EULER)
            integrate<true>(C, s, actv);
        else
            integrate<false>(C, s, actv);
        constrain_state(P, C, s);
        t += C.dt;
        tg.template state<true>(P, t, tp, tv, ta);
        err_post = norm3d(s[0] - tp[0], s[1] - tp[1], s[2] - tp[2]);
        term = termination_code(P, C, s, t);
        env_on += (err_post <= P.env_target_radius) ? 1 : 0;
        if (log_traj) {
            double *row = A.traj + (int64_t)traj_slot * QT_TRAJ_FIELDS * E + e;
            row[(int64_t)QT_TRAJ_TIME * E] = t;
            row[(int64_t)QT_TRAJ_ERR * E] = err_post;
            next_traj += A.traj_stride;
            traj_slot += 1;
        }
        step += 1;
        err_pre = err_post;
    }
    if (in_phase && streak >= P.overshoot_window) {
        ov_count += 1;
        ov_max = fmax(ov_max, ov_cur);
    }
    const double n = (double)step;
    const double env_ratio = step > 0 ? (double)env_on / n : 0.0;
    const double met_ratio = step > 0 ? (double)met_on / n : 0.0;
    double *m = A.metrics + e;
    m[(int64_t)QT_M_STEPS * E] = n;
    m[(int64_t)QT_M_FINAL_TIME * E] = t;
    m[(int64_t)QT_M_ENV_ON_TARGET * E] = (double)env_on;
    m[(int64_t)QT_M_MET_ON_TARGET * E] = (double)met_on;
    m[(int64_t)QT_M_SUM_ERR * E] = sum_err;
    m[(int64_t)QT_M_MAX_ERR * E] = max_err;
    m[(int64_t)QT_M_SUM_ERR2 * E] = sum_err2;
    m[(int64_t)QT_M_SUM_EFFORT * E] = sum_eff;
    m[(int64_t)QT_M_OVERSHOOT_COUNT * E] = (double)ov_count;
    m[(int64_t)QT_M_MAX_OVERSHOOT * E] = ov_max;
    m[(int64_t)QT_M_ENV_SUCCESS * E] =
        (t >= P.env_min_episode_duration && env_ratio >= P.env_min_on_target_ratio) ? 1.0 : 0.0;
    m[(int64_t)QT_M_MET_SUCCESS * E] =
        (step > 0 && t >= P.crit_min_episode_duration && met_ratio >= P.crit_min_on_target_ratio) ? 1.0 : 0.0;
    m[(int64_t)QT_M_TERMINATION * E] = (double)term;
    m[(int64_t)QT_M_VIOLATIONS * E] = (double)violations;
    m[(int64_t)QT_M_SATURATIONS * E] = (double)saturations;
    m[(int64_t)QT_M_FINAL_ERR * E] = err_post;
    if (A.final_state) {
#pragma unroll
        for (int i = 0; i < 12; i++) A.final_state[(int64_t)i * E + e] = s[i];
    }
}

struct PolicyActionArgs {
    int64_t E;
    int w_in_smem;
    const double *obs;
    const float *W;
    double *action;
};

__global__ void __launch_bounds__(kPolicyThreads)
    policy_action_kernel(const __grid_constant__ MlpDesc M, const __grid_constant__ PolicyActionArgs A) {
    extern __shared__ __align__(16) float smem_f[];
    constexpr int T = kPolicyThreads;
    const int64_t E = A.E;
    const int64_t e = (int64_t)blockIdx.x * T + threadIdx.x;
    const int w_floats = A.w_in_smem ? ((M.total_floats + 3) & ~3) : 0;
    stage_weights(M, A.W, A.w_in_smem, smem_f);
    float *tile0 = smem_f + w_floats + threadIdx.x;
    float *tile1 = tile0 + M.tile_w[0] * T;
    if (e >= E) return;
    double s[12], tp[3], tv[3];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = A.obs[(int64_t)i * E + e];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        tp[i] = A.obs[(int64_t)(12 + i) * E + e];
        tv[i] = A.obs[(int64_t)(15 + i) * E + e];
    }
    float feat[QT_FEATURE_DIM];
    make_features(s, tp, tv, feat);
    float out[4];
    if (A.w_in_smem)
        mlp_forward<T>(M, smem_f, tile0, tile1, feat, out);
    else
        mlp_forward<T>(M, A.W, tile0, tile1, feat, out);
#pragma unroll
    for (int i = 0; i < 4; i++) A.action[(int64_t)i * E + e] = (double)out[i];
}

// shared memory plan: [weights if they fit][activation tile 0][activation tile 1]
static size_t smem_plan(const MlpDesc &M, int &w_in_smem) {
    const size_t act = (size_t)(M.tile_w[0] + M.tile_w[1]) * kPolicyThreads * sizeof(float);
    const size_t w = (((size_t)M.total_floats + 3) & ~(size_t)3) * sizeof(float);
    const size_t budget = 100 * 1024;  // at least two blocks per SM
    w_in_smem = (act + w <= budget) ? 1 : 0;
    return act + (w_in_smem ? w : 0);
}

template <int MOTION>
static cudaError_t launch_policy(const qt_params &P, const MlpDesc &M, const PolicyArgs &A, size_t smem, cudaStream_t st) {
    DynConsts C;
    C.init(P);
    const int64_t blocks = (A.E + kPolicyThreads - 1) / kPolicyThreads;
    if (A.w_in_smem && smem <= 72 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(rollout_policy_kernel<MOTION, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        rollout_policy_kernel<MOTION, 3><<<(unsigned)blocks, kPolicyThreads, smem, st>>>(P, C, M, A);
    } else {
        cudaError_t e = cudaFuncSetAttribute(rollout_policy_kernel<MOTION, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        rollout_policy_kernel<MOTION, 2><<<(unsigned)blocks, kPolicyThreads, smem, st>>>(P, C, M, A);
    }
    return cudaGetLastError();
}

}  // namespace qt

extern "C" int64_t qt_mlp_packed_floats(const qt_mlp *policy) {
    qt::MlpDesc d;
    if (!policy || !qt::describe(*policy, d)) return -1;
    return d.total_floats;
}

extern "C" int qt_policy_action_batched(qt_context *ctx, const qt_mlp *policy, int64_t E, const double *obs_dev,
                                        double *action_out_dev) {
    using namespace qt;
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, policy && obs_dev && action_out_dev && policy->weights_dev, "null pointer");
    QT_REQUIRE(ctx, E >= 0, "E < 0");
    MlpDesc M;
    QT_REQUIRE(ctx, describe(*policy, M), "unsupported network shape (18 inputs, 4 outputs, <= 6 layers, widths <= 256)");
    QT_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(policy->weights_dev) & 15) == 0, "weights must be 16-byte aligned");
    if (E == 0) return QT_OK;
    PolicyActionArgs A{E, 0, obs_dev, policy->weights_dev, action_out_dev};
    const size_t smem = smem_plan(M, A.w_in_smem);
    QT_REQUIRE(ctx, smem <= 200 * 1024, "network too wide for the shared-memory activation tiles");
    DeviceGuard guard(ctx->device);
    QT_CUDA_TRY(ctx, cudaFuncSetAttribute(policy_action_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t blocks = (E + kPolicyThreads - 1) / kPolicyThreads;
    policy_action_kernel<<<(unsigned)blocks, kPolicyThreads, smem, ctx->stream>>>(M, A);
    return check_launch(ctx, "policy_action_kernel");
}

extern "C" int qt_rollout_policy(qt_context *ctx, const qt_params *params, int64_t E, int64_t max_steps,
                                 const double *state0_dev, const double *tparams_dev, const qt_mlp *policy,
                                 const double *teacher_K_dev, int32_t k_cols, int64_t k_stride,
                                 double *metrics_out_dev, double *final_state_out_dev, double *traj_out_dev,
                                 int64_t traj_stride, int64_t traj_slots, float *feat_out_dev, float *act_out_dev,
                                 int64_t data_stride, int64_t data_slots) {
    using namespace qt;
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, params && state0_dev && tparams_dev && policy && policy->weights_dev && metrics_out_dev, "null pointer");
    QT_REQUIRE(ctx, E >= 0 && max_steps >= 0 && max_steps < (1ll << 31), "bad E / max_steps");
    QT_REQUIRE(ctx, params->motion_type >= 0 && params->motion_type <= QT_MOTION_FIGURE8, "motion_type");
    QT_REQUIRE(ctx, (feat_out_dev == nullptr) == (act_out_dev == nullptr), "feat/act must come together");
    MlpDesc M;
    QT_REQUIRE(ctx, describe(*policy, M), "unsupported network shape (18 inputs, 4 outputs, <= 6 layers, widths <= 256)");
    QT_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(policy->weights_dev) & 15) == 0, "weights must be 16-byte aligned");
    if (teacher_K_dev) {
        QT_REQUIRE(ctx, k_cols == expected_k_cols(*params), "k_cols must be 6 (LQR), 9 (LQI) or 3 (PID)");
        QT_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(teacher_K_dev) & 15) == 0, "K must be 16-byte aligned");
    }
    if (E == 0) return QT_OK;
    if (k_stride < 1) k_stride = 1;
    if (traj_stride < 1) traj_stride = 1;
    if (data_stride < 1) data_stride = 1;
    QT_REQUIRE(ctx, traj_stride < (1ll << 31) && data_stride < (1ll << 31) && traj_slots < (1ll << 31) &&
                        data_slots < (1ll << 31), "log strides/slots");
    PolicyArgs A;
    A.E = E;
    A.k_stride = k_stride;
    A.max_steps = (int)max_steps;
    A.teacher = teacher_K_dev ? 1 + ctrl_kind_of(*params) : 0;
    A.state0 = state0_dev;
    A.tparams = tparams_dev;
    A.K = teacher_K_dev;
    A.W = policy->weights_dev;
    A.metrics = metrics_out_dev;
    A.final_state = final_state_out_dev;
    A.traj = traj_out_dev;
    A.feat = feat_out_dev;
    A.actf = act_out_dev;
    A.traj_stride = (int)traj_stride;
    A.traj_slots = (int)traj_slots;
    A.data_stride = (int)data_stride;
    A.data_slots = (int)data_slots;
    const size_t smem = smem_plan(M, A.w_in_smem);
    QT_REQUIRE(ctx, smem <= 200 * 1024, "network too wide for the shared-memory activation tiles");
    DeviceGuard guard(ctx->device);
    cudaError_t err;
    switch (params->motion_type) {
    case QT_MOTION_LINEAR: err = launch_policy<QT_MOTION_LINEAR>(*params, M, A, smem, ctx->stream); break;
    case QT_MOTION_CIRCULAR: err = launch_policy<QT_MOTION_CIRCULAR>(*params, M, A, smem, ctx->stream); break;
    case QT_MOTION_SINUSOIDAL: err = launch_policy<QT_MOTION_SINUSOIDAL>(*params, M, A, smem, ctx->stream); break;
    case QT_MOTION_FIGURE8: err = launch_policy<QT_MOTION_FIGURE8>(*params, M, A, smem, ctx->stream); break;
    default: err = launch_policy<QT_MOTION_STATIONARY>(*params, M, A, smem, ctx->stream); break;
    }
    if (err != cudaSuccess) return fail(ctx, QT_ERR_CUDA, "rollout_policy_kernel launch", cudaGetErrorString(err));
    ctx->launches += 1;
    return QT_OK;
}
