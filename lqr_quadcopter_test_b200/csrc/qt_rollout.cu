// qt_rollout.cu — the fused closed-loop rollout kernel.
//
// Replaces the reference's hot loop
//     while not done: action = controller.compute_action(obs); obs, r, done, info = env.step(action)
// (eval.py:119-165, controllers/tuning.py:885-906) together with compute_episode_metrics
// (utils/metrics.py:264-338) by ONE kernel launch for all episodes.
//
// Mapping: one thread per chain of episodes (normally one episode).  The 12-element state, the
// LQI integrator, the target-pattern parameters, simulation time and every metric accumulator
// stay in registers for all max_steps iterations; the gain matrix K of the block's episodes is
// staged once in shared memory (reads are warp-uniform when k_stride is a multiple of 32).  HBM
// is touched at the beginning (initial conditions), at the end (16 doubles per episode) and, in
// the FULL variant only, for the optional strided trajectory log / teacher data set, which are
// written struct-of-arrays so a warp's stores are contiguous.  The kernel is bound by the FP64
// pipe; there is no GEMM-shaped work and no use for tensor cores.
#include <type_traits>

#include "qt_common.cuh"
#include "qt_npsum.cuh"
#include "qt_step.cuh"

namespace qt {

#ifndef QT_ROLLOUT_THREADS
#define QT_ROLLOUT_THREADS 128   // A/B hook (profiles/r02_rollout_ab.txt); the fused candidate scores need a divisor-rich block size
#endif
constexpr int kRolloutThreads = QT_ROLLOUT_THREADS;
#ifndef QT_OPT_KREG
#define QT_OPT_KREG 1
#endif
#ifndef QT_OPT_COLDPTR
#define QT_OPT_COLDPTR 1
#endif
#ifndef QT_OPT_DATASTAGE
#define QT_OPT_DATASTAGE 1   // teacher data set: feature rows leave through a per-warp shared-memory tile, fully coalesced
#endif
#ifndef QT_OPT_RARE
#define QT_OPT_RARE 3     // one combined test + branch for the step's rare range events (1: all, 2: tilt separate, 3: 1 for LQR / PID, 2 for LQI)
#endif
#ifndef QT_OPT_FINITE
#define QT_OPT_FINITE 1   // structured loop: no per-step non-finite tests (proved impossible once per episode)
#endif
#ifndef QT_RESYNC_MASK
#define QT_RESYNC_MASK 15  // the carried sin / cos (attitude, target pattern) are re-evaluated directly every 16th step
#endif
#ifndef QT_ROLLOUT_MIN_BLOCKS
#define QT_ROLLOUT_MIN_BLOCKS 4  // resident blocks per SM the register allocator must allow (128 registers)
#endif

struct RolloutArgs {
    int64_t E, n_chains, chain, k_stride;
    int max_steps;
    int k_in_smem;
    int k_smem_doubles;  // doubles reserved for the staged gains at the start of shared memory
    int env_bounds_check;  // 0 when the controller's clamps lie inside the env's action limits
    const double *state0, *tparams, *K;
    const double *ff;  // optional per-group feedforward gains [groups][6], NULL = qt_params'
    double *metrics, *ctrl_state, *final_state, *traj;
    float *feat, *actf;
    int traj_stride, traj_slots, data_stride, data_slots;
    // candidate sweep (qt_rollout_sweep): episode e starts from initial condition e % ic_period (0 = its own),
    // and each block reduces the eps episodes of its candidates to scores[4][n_cand] (NULL = no fused reduction)
    int64_t ic_period;
    double *scores;
    int64_t n_cand;
    int eps;
};

// Per-thread values that are touched once per step (metric accumulators, target-pattern
// parameters) live in this thread's column of a shared-memory tile instead of registers: the
// integrator needs ~110 registers at its widest point, and keeping the cold state out of the
// register file is what lets four 128-thread blocks (16 warps) stay resident per SM without
// spilling.  Column access (row * blockDim + threadIdx) is conflict free.
enum ColdRow { CR_SUM_ERR = 0, CR_SUM_ERR2, CR_SUM_EFF, CR_OV_CUR, CR_OV_MAX, CR_TARGET };
constexpr int CR_TARGET_TRIG = CR_TARGET + QT_TARGET_PARAMS;  // 6 rows: carried sin/cos of the pattern angles
constexpr int kColdRows = CR_TARGET_TRIG + 6;

// MODE 0 (lean): metrics only.  MODE 1 (full): feedforward, Euler integrator, trajectory log, data set.
// MODE 2 (data): the lean loop plus the teacher data set rows (BASELINE config 5) — the full variant's extra
// live state (target acceleration, feedforward gains, integrator switch, trajectory cursor) costs the
// register-bound loop far more than the rows themselves.
template <int MOTION, int CTRL, int MODE>
__global__ void __launch_bounds__(kRolloutThreads, QT_ROLLOUT_MIN_BLOCKS)
    rollout_kernel(const __grid_constant__ qt_params P, const __grid_constant__ DynConsts C,
                   const __grid_constant__ RolloutArgs A) {
    extern __shared__ double smem[];
    constexpr bool FULL = MODE == 1, DATA = MODE != 0;
    constexpr int KC = CtrlCols<CTRL>::value;
    constexpr int T = kRolloutThreads;
    const int64_t E = A.E;
    const int64_t chain_id = (int64_t)blockIdx.x * T + threadIdx.x;
    double *cold = smem + threadIdx.x;   // this thread's column of the cold tile: cold[row * T]
#if QT_OPT_COLDPTR
    // Keep that address in a register.  Left to itself ptxas rebuilds it from S2R TID.X / S2UR CgaCtaId / ULEA / LEA
    // three times per env step (ncu r02_a: those 15 instructions carry the largest stall counts of the loop — they wait
    // on the special-register reads); the empty asm hides where the value came from, the assumption keeps LDS / STS.
    asm volatile("" : "+l"(cold));
    __builtin_assume(__isShared(cold));
#endif
    double *sK = smem + kColdRows * T;   // staged gains follow the tile (compile-time offset)

    int64_t g0 = 0;
    if (A.k_in_smem) {
        // chain == 1 here: the block covers episodes [b*T, b*T+T)
        const int64_t ep0 = (int64_t)blockIdx.x * T;
        int64_t ep1 = ep0 + T - 1;
        if (ep1 > E - 1) ep1 = E - 1;
        g0 = ep0 / A.k_stride;
        const int64_t g1 = ep1 / A.k_stride;
        const int n = (int)(g1 - g0 + 1) * 4 * KC;
        const double *src = A.K + g0 * 4 * KC;
        for (int i = threadIdx.x; i < n; i += T) sK[i] = src[i];
        __syncthreads();
    }
    // threads past the last chain fall through with an empty episode range: the score epilogue below
    // synchronises the whole block
    const bool active = chain_id < A.n_chains;

    const bool euler = FULL && (P.integrator == QT_INTEGRATOR_EULER);

    double integ[3] = {0.0, 0.0, 0.0};
    int saturations = 0;
    double last_time = __longlong_as_double(0x7ff8000000000000LL);  // PID: time of the previous call, NaN = none
    if (A.ctrl_state && A.chain > 1 && active) {  // the controller object's state at the time of the call
#pragma unroll
        for (int i = 0; i < 3; i++) integ[i] = A.ctrl_state[(int64_t)i * A.n_chains + chain_id];
        if (CTRL == CK_PID)
            last_time = A.ctrl_state[3 * A.n_chains + chain_id];
        else
            saturations = (int)A.ctrl_state[3 * A.n_chains + chain_id];
    }

    const int64_t first = active ? chain_id * A.chain : E;
    int64_t last = first + A.chain;
    if (last > E) last = E;
    // initial conditions: [12][E] indexed by the episode, or (sweep) [12][ic_period] shared by all candidates
    const int64_t ic_E = A.ic_period > 0 ? A.ic_period : E;

    for (int64_t e = first; e < last; e++) {
        const int64_t e_ic = A.ic_period > 0 ? e % A.ic_period : e;
        double s[12];
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = A.state0[(int64_t)i * ic_E + e_ic];
        TargetGen<MOTION, T> tg;
        tg.load(cold + CR_TARGET * T, A.tparams, ic_E, e_ic);
        const int64_t grp = e / A.k_stride;
        const double *Kp = A.k_in_smem ? (sK + (grp - g0) * 4 * KC) : (A.K + grp * 4 * KC);
        const double *ffg = (FULL && A.ff) ? A.ff + grp * 6 : nullptr;
        if (A.chain == 1) {  // controller.reset() before every episode (tuning.py:879)
            integ[0] = integ[1] = integ[2] = 0.0;
            saturations = 0;
            last_time = __longlong_as_double(0x7ff8000000000000LL);
        }

        double t = 0.0;
        int step = 0, env_on = 0, met_on = 0, violations = 0, term = QT_TERM_NONE;
        double err_post = 0.0;
        cold[CR_SUM_ERR * T] = 0.0;
        cold[CR_SUM_ERR2 * T] = 0.0;
        cold[CR_SUM_EFF * T] = 0.0;
        cold[CR_OV_CUR * T] = 0.0;
        cold[CR_OV_MAX * T] = 0.0;
        // overshoot detector state (utils/metrics.py:205-261), streamed
        bool prev_on = false, in_phase = false;
        int streak = 0, ov_count = 0;
        int next_traj = 0, traj_slot = 0, next_data = 0, data_slot = 0;

        double tp[3], tv[3], ta[3];
        tg.trig_init(cold + CR_TARGET_TRIG * T, t);
        tg.template state<FULL>(P, t, tp, tv, ta);
        double err_pre = norm3d(tp[0] - s[0], tp[1] - s[1], tp[2] - s[2]);
        // np.max of the pre-step errors (utils/metrics.py:329).  Sample 0 initialises it; later samples are
        // never NaN (a non-finite state ends the episode before it is sampled), so a plain compare suffices
        double max_err = err_pre;
        Trig B;  // sin/cos of the current attitude, carried from step to step (trig_rebase)
        {
            const double e0[3] = {s[6], s[7], s[8]};
            trig_full(e0, B);
        }

        // ---- structure test (once per episode).  Every design the reference itself produces (PID,
        // heuristic LQR, Riccati-LQR/LQI with diagonal Q and R) commands a yaw rate of exactly zero
        // through an axis-decoupled gain, and reset() starts with zero yaw: yaw then stays exactly
        // zero and the products with the zero gain entries vanish.  The FAST instantiation of the
        // step loop skips that arithmetic; it is bit-identical to the general one (up to the sign
        // of exact zeros) and is only entered when the whole warp qualifies, so the two loop bodies
        // never diverge inside a warp.  params.debug_flags bit 0 forces the general path (tests).
        bool fast = !(P.debug_flags & QT_DEBUG_GENERAL_PATH) && s[8] == 0.0 && s[11] == 0.0;
        {
            int hmax = 0;
#pragma unroll
            for (int i = 0; i < 12; i++) hmax = max(hmax, hi_abs(s[i]));
#pragma unroll
            for (int i = 0; i < TargetParamCount<MOTION>::value; i++) hmax = max(hmax, hi_abs(cold[(CR_TARGET + i) * T]));
#pragma unroll
            for (int i = 0; i < 3; i++) hmax = max(hmax, hi_abs(integ[i]));
            // every gain entry as well: with a finite state, target, gain and constants the action is finite (it is a
            // clamped affine function of them), the env never has to repair it, and the state stays finite for the
            // whole episode — the structured loop drops both per-step non-finite tests (QT_OPT_FINITE)
#pragma unroll
            for (int i = 0; i < 4 * KC; i++) hmax = max(hmax, hi_abs(Kp[i]));
            if (ffg) {
#pragma unroll
                for (int i = 0; i < 6; i++) hmax = max(hmax, hi_abs(ffg[i]));
            }
            fast = fast && hmax < 0x7ff00000 && C.all_finite;  // (0 * inf would also differ from the dense products)
            if constexpr (CTRL != CK_PID) fast = fast && gain_is_axis_decoupled<CTRL == CK_LQI>(Kp);
        }
        const bool warp_fast = __all_sync(__activemask(), fast);

        auto step_loop = [&](auto fast_tag) {
        constexpr bool FAST = decltype(fast_tag)::value;
#if QT_OPT_KREG
        // the structured LQR law uses six gain entries: keep them in registers (same layout, so ctrl_action's
        // constant indices resolve to registers) instead of six loads at the top of every step
        double kreg[4 * KC];
        const double *Kc = Kp;
        if constexpr (FAST && (CTRL == CK_LQR || (QT_OPT_KREG > 1 && CTRL == CK_LQI))) {
#pragma unroll
            for (int i = 0; i < 4 * KC; i++) kreg[i] = 0.0;
#pragma unroll
            for (int c = 0; c < KC; c += 3) {
                kreg[0 * KC + c + 2] = Kp[0 * KC + c + 2];
                kreg[1 * KC + c + 1] = Kp[1 * KC + c + 1];
                kreg[2 * KC + c + 0] = Kp[2 * KC + c + 0];
            }
            Kc = kreg;
        }
#else
        const double *Kc = Kp;
#endif
        while (term == QT_TERM_NONE && step < A.max_steps) {
            // ---- controller on the pre-step observation
            double act[4];
            if constexpr (CTRL == CK_PID)
                pid_action<FULL>(P, Kp, s, tp, tv, ta, t, integ, last_time, act, ffg);
            else
                ctrl_action<CTRL == CK_LQI, FULL, FAST>(P, Kc, s, tp, tv, ta, err_pre, integ, saturations, act, ffg);

            // ---- metrics sample k = step (pre-step observation, eval.py:142-159); the control effort's
            // square root is taken together with the tracking error's at the end of the step (sqrt_pair)
            const double effort2 = act[0] * act[0] + act[1] * act[1] + act[2] * act[2] + act[3] * act[3];
            {
                const bool on = err_pre <= P.crit_target_radius;
                met_on += on ? 1 : 0;
                cold[CR_SUM_ERR * T] += err_pre;
                cold[CR_SUM_ERR2 * T] += err_pre * err_pre;
                max_err = err_pre > max_err ? err_pre : max_err;
                // overshoot detector: nothing to do while the vehicle stays on target (or has never been)
                if (in_phase | (prev_on & !on)) {
                    if (!in_phase) {            // prev_on && !on: a departure starts
                        in_phase = true;
                        streak = 1;
                        cold[CR_OV_CUR * T] = err_pre - P.crit_target_radius;
                    } else if (!on) {
                        streak += 1;
                        cold[CR_OV_CUR * T] = fmax(cold[CR_OV_CUR * T], err_pre - P.crit_target_radius);
                    } else {
                        if (streak >= P.overshoot_window) {
                            ov_count += 1;
                            cold[CR_OV_MAX * T] = fmax(cold[CR_OV_MAX * T], cold[CR_OV_CUR * T]);
                        }
                        in_phase = false;
                        streak = 0;
                        cold[CR_OV_CUR * T] = 0.0;
                    }
                }
                prev_on = on;
            }

            [[maybe_unused]] bool log_traj = false;
            if constexpr (FULL) {
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
            }
            if constexpr (DATA) {
                if (A.feat && step == next_data && data_slot < A.data_slots) {
                    // deep_tracking_policy.py:241-286 feature layout, fp32
                    float ft[18];
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        ft[i] = (float)(tp[i] - s[i]);
                        ft[3 + i] = (float)(tv[i] - s[3 + i]);
                        ft[6 + i] = (float)s[6 + i];
                        ft[9 + i] = (float)s[9 + i];
                        ft[12 + i] = (float)tp[i];
                        ft[15 + i] = (float)tv[i];
                    }
                    const int64_t row = (int64_t)data_slot * E + e;
                    float *yrow = A.actf + row * QT_ACTION_DIM;
                    float *frow = A.feat + row * QT_FEATURE_DIM;
#if QT_OPT_DATASTAGE
                    // Coalesced rows.  An action row is 16 bytes: one float4 per lane, 512 contiguous bytes per warp.
                    // A feature row is 72 bytes; written per thread, a warp's store touches 32 different sectors per
                    // instruction (ncu r02_a: 8.8 useful bytes per 32-byte sector).  When all 32 lanes of the warp log
                    // the same slot — always, until episodes of the warp end at different steps — the rows are laid
                    // side by side in a per-warp shared-memory tile and leave as 144 float4 = 2304 contiguous bytes.
                    const int lane = threadIdx.x & 31;
                    const int64_t row0 = row - lane;                     // the warp's first row
                    const bool aligned16 = ((reinterpret_cast<uintptr_t>(A.feat) | reinterpret_cast<uintptr_t>(A.actf)) & 15) == 0;
                    if (aligned16) {
                        *reinterpret_cast<float4 *>(yrow) = make_float4((float)act[0], (float)act[1], (float)act[2], (float)act[3]);
                    } else {
                        reinterpret_cast<float2 *>(yrow)[0] = make_float2((float)act[0], (float)act[1]);
                        reinterpret_cast<float2 *>(yrow)[1] = make_float2((float)act[2], (float)act[3]);
                    }
                    if (__activemask() == 0xffffffffu && aligned16 && (row0 & 1) == 0) {
                        float *tile = reinterpret_cast<float *>(smem + kColdRows * T + A.k_smem_doubles) + (threadIdx.x >> 5) * (32 * QT_FEATURE_DIM);
                        float2 *mine = reinterpret_cast<float2 *>(tile + lane * QT_FEATURE_DIM);
#pragma unroll
                        for (int i = 0; i < 9; i++) mine[i] = make_float2(ft[2 * i], ft[2 * i + 1]);
                        __syncwarp();
                        const float4 *src = reinterpret_cast<const float4 *>(tile);
                        float4 *dst = reinterpret_cast<float4 *>(A.feat + row0 * QT_FEATURE_DIM);
#pragma unroll
                        for (int k = 0; k < 5; k++) {
                            const int idx = lane + 32 * k;
                            if (idx < 32 * QT_FEATURE_DIM / 4) dst[idx] = src[idx];
                        }
                        __syncwarp();
                    } else {
                        float2 *f = reinterpret_cast<float2 *>(frow);
#pragma unroll
                        for (int i = 0; i < 9; i++) f[i] = make_float2(ft[2 * i], ft[2 * i + 1]);
                    }
#else
                    float2 *f = reinterpret_cast<float2 *>(frow);
                    float2 *y = reinterpret_cast<float2 *>(yrow);
#pragma unroll
                    for (int i = 0; i < 9; i++) f[i] = make_float2(ft[2 * i], ft[2 * i + 1]);
                    y[0] = make_float2((float)act[0], (float)act[1]);
                    y[1] = make_float2((float)act[2], (float)act[3]);
#endif
                    next_data += A.data_stride;
                    data_slot += 1;
                }
            }

            // ---- env.step (quadcopter_env.py:171-207)
            double actv[4], eta_old[3];
            // re-basing of the carried attitude trig inside the integrator's stage block (RK4 only)
            constexpr bool kRebaseInline = QT_OPT_REBASE4 && QT_OPT_STAGES3 && QT_OPT_FOLDED_RK4 && QT_OPT_FLAT_RK4 && QT_OPT_RARE;
            [[maybe_unused]] Trig Bnext;
            [[maybe_unused]] bool attitude_changed = false;
            if constexpr (FAST && QT_OPT_FINITE)
                violations += A.env_bounds_check ? validate_action<true, false>(P, act, actv) : validate_action<false, false>(P, act, actv);
            else
                violations += A.env_bounds_check ? validate_action<true>(P, act, actv) : validate_action<false>(P, act, actv);
            if (euler)
                integrate<true, FAST>(C, s, actv, B, eta_old);
            else if (QT_OPT_FLAT_RK4)
                integrate_flat<FAST>(C, s, actv, B, eta_old, kRebaseInline ? &Bnext : nullptr);
            else
                integrate<false, FAST>(C, s, actv, B, eta_old);
#if QT_OPT_RARE
            // one branch for the step's rare range events instead of four (body rates, speed, position bound and —
            // except for the LQI law, whose default gains sit in the tilt limit most of the time — tilt / wrap); the
            // usual step only takes the wrap's roundings (a + pi) - pi
            constexpr bool kTiltFused = (QT_OPT_RARE == 1) || (QT_OPT_RARE == 3 && CTRL != CK_LQI);
            const bool unusual = step_is_unusual<FAST, kTiltFused>(C, s, s[3] * s[3] + s[4] * s[4] + s[5] * s[5]);
            if (!unusual) {
                if constexpr (kTiltFused) {
                    s[6] = (s[6] + kPi) - kPi;
                    s[7] = (s[7] + kPi) - kPi;
                    if (!FAST) s[8] = (s[8] + kPi) - kPi;
                } else {
                    attitude_changed = constrain_attitude<FAST>(s);
                }
            } else {
                attitude_changed = true;
                constrain_state<FAST>(P, C, s);
            }
#else
            constexpr bool unusual = true;
            constrain_state<FAST>(P, C, s);
#endif
            if (kRebaseInline && !euler) {
                // Bnext belongs to the attitude the integrator produced; the constraints usually only re-round it
                // ((a + pi) - pi, <= 2.3e-16).  A wrap, a clip or the periodic re-synchronisation evaluates directly.
                if (attitude_changed || (step & QT_RESYNC_MASK) == QT_RESYNC_MASK)
                    trig_rebase<FAST>(eta_old, s, true, B);
                else
                    B = Bnext;
            } else {
                trig_rebase<FAST>(eta_old, s, (step & QT_RESYNC_MASK) == QT_RESYNC_MASK, B);
            }
            {
                const double t_old = t;
                t += C.dt;
                tg.trig_advance(t_old, t, (step & QT_RESYNC_MASK) == QT_RESYNC_MASK);
            }
            tg.template state<FULL>(P, t, tp, tv, ta);
            {
                const double dx = s[0] - tp[0], dy = s[1] - tp[1], dz = s[2] - tp[2];
                double effort;
                sqrt_pair(dx * dx + dy * dy + dz * dz, effort2, err_post, effort);
                cold[CR_SUM_EFF * T] += effort;
            }
            if (t >= P.max_episode_time)
                term = QT_TERM_TIME_LIMIT;
            else if (unusual || !(FAST && QT_OPT_FINITE))
                term = termination_code<FAST && QT_OPT_FINITE>(P, C, s, t);
            env_on += (err_post <= P.env_target_radius) ? 1 : 0;
            if constexpr (FULL) {
                if (log_traj) {
                    double *row = A.traj + (int64_t)traj_slot * QT_TRAJ_FIELDS * E + e;
                    row[(int64_t)QT_TRAJ_TIME * E] = t;
                    row[(int64_t)QT_TRAJ_ERR * E] = err_post;
                    next_traj += A.traj_stride;
                    traj_slot += 1;
                }
            }
            step += 1;
            err_pre = err_post;
        }
        };
#ifdef QT_EXPERIMENT_FAST_ONLY   // A/B experiment: register budget of the structured loop on its own
        (void)warp_fast;
        step_loop(std::true_type{});
#else
        if (warp_fast)
            step_loop(std::true_type{});
        else
            step_loop(std::false_type{});
#endif
        double ov_max = cold[CR_OV_MAX * T];
        if (in_phase && streak >= P.overshoot_window) {
            ov_count += 1;
            ov_max = fmax(ov_max, cold[CR_OV_CUR * T]);
        }

        const double n = (double)step;
        const double env_ratio = step > 0 ? (double)env_on / n : 0.0;
        const double met_ratio = step > 0 ? (double)met_on / n : 0.0;
        const double met_success =
            (step > 0 && t >= P.crit_min_episode_duration && met_ratio >= P.crit_min_on_target_ratio) ? 1.0 : 0.0;
        if (A.scores) {
            // what reduce_scores_kernel would read back from the metric rows (tuning.py:908-917), parked in
            // this thread's column of the cold tile for the block-level reduction after the loop
            const double mean_err = step > 0 ? cold[CR_SUM_ERR * T] / n : 0.0;
            // (rows of the target-pattern trig: free once the episode is over, and chain == 1 here)
            cold[(CR_TARGET_TRIG + 0) * T] = met_ratio;
            cold[(CR_TARGET_TRIG + 1) * T] = mean_err;
            cold[(CR_TARGET_TRIG + 2) * T] = met_success;
        }
        if (A.metrics) {
        double *m = A.metrics + e;
        m[(int64_t)QT_M_STEPS * E] = n;
        m[(int64_t)QT_M_FINAL_TIME * E] = t;
        m[(int64_t)QT_M_ENV_ON_TARGET * E] = (double)env_on;
        m[(int64_t)QT_M_MET_ON_TARGET * E] = (double)met_on;
        m[(int64_t)QT_M_SUM_ERR * E] = cold[CR_SUM_ERR * T];
        m[(int64_t)QT_M_MAX_ERR * E] = step > 0 ? max_err : 0.0;
        m[(int64_t)QT_M_SUM_ERR2 * E] = cold[CR_SUM_ERR2 * T];
        m[(int64_t)QT_M_SUM_EFFORT * E] = cold[CR_SUM_EFF * T];
        m[(int64_t)QT_M_OVERSHOOT_COUNT * E] = (double)ov_count;
        m[(int64_t)QT_M_MAX_OVERSHOOT * E] = ov_max;
        m[(int64_t)QT_M_ENV_SUCCESS * E] =
            (t >= P.env_min_episode_duration && env_ratio >= P.env_min_on_target_ratio) ? 1.0 : 0.0;
        m[(int64_t)QT_M_MET_SUCCESS * E] = met_success;
        m[(int64_t)QT_M_TERMINATION * E] = (double)term;
        m[(int64_t)QT_M_VIOLATIONS * E] = (double)violations;
        m[(int64_t)QT_M_SATURATIONS * E] = (double)saturations;
        m[(int64_t)QT_M_FINAL_ERR * E] = err_post;
        }
        if (A.final_state) {
#pragma unroll
            for (int i = 0; i < 12; i++) A.final_state[(int64_t)i * E + e] = s[i];
        }
    }
    if (A.ctrl_state && active) {
#pragma unroll
        for (int i = 0; i < 3; i++) A.ctrl_state[(int64_t)i * A.n_chains + chain_id] = integ[i];
        A.ctrl_state[3 * A.n_chains + chain_id] = CTRL == CK_PID ? last_time : (double)saturations;
    }
    if (A.scores) {
        // ControllerTuner._evaluate_config's aggregation (tuning.py:912-926) fused into the rollout: the host
        // guarantees chain == 1 and eps | 128, so a block holds whole candidates; the first thread of each
        // candidate sums its eps episodes in numpy's pairwise order (qt_npsum.cuh) -> bit-identical to
        // qt_reduce_scores on the metric rows, which then never have to be written (16 x 8 B per episode).
        __syncthreads();
        const int eps = A.eps;
        if (active && (threadIdx.x % eps) == 0) {
            const double *col = smem + threadIdx.x;
            auto ratio_f = [=](int64_t k) { return col[(CR_TARGET_TRIG + 0) * T + k]; };
            auto merr_f = [=](int64_t k) { return col[(CR_TARGET_TRIG + 1) * T + k]; };
            auto succ_f = [=](int64_t k) { return col[(CR_TARGET_TRIG + 2) * T + k]; };
            const double ne = (double)eps;
            const double mean_on = np_sum(ratio_f, eps) / ne;
            const double mean_err = np_sum(merr_f, eps) / ne;
            const double sr = np_sum(succ_f, eps) / ne;
            const int64_t c = chain_id / eps, C = A.n_cand;
            A.scores[0 * C + c] = __dadd_rn(mean_on, -__dmul_rn(0.1, mean_err));  // mean_on - 0.1 * mean_err, unfused
            A.scores[1 * C + c] = mean_on;
            A.scores[2 * C + c] = mean_err;
            A.scores[3 * C + c] = sr;
        }
    }
}

template <int MOTION, int CTRL, int MODE>
static cudaError_t launch_one(const qt_params &P, const RolloutArgs &A, size_t smem, cudaStream_t st) {
    const int64_t blocks = (A.n_chains + kRolloutThreads - 1) / kRolloutThreads;
    if (blocks <= 0) return cudaSuccess;
    DynConsts C;
    C.init(P);
    rollout_kernel<MOTION, CTRL, MODE><<<(unsigned)blocks, kRolloutThreads, smem, st>>>(P, C, A);
    return cudaGetLastError();
}

template <int MOTION, int CTRL>
static cudaError_t launch_mode(const qt_params &P, const RolloutArgs &A, size_t smem, int mode, cudaStream_t st) {
    if (mode == 1) return launch_one<MOTION, CTRL, 1>(P, A, smem, st);
    if (mode == 2) return launch_one<MOTION, CTRL, 2>(P, A, smem, st);
    return launch_one<MOTION, CTRL, 0>(P, A, smem, st);
}

template <int MOTION>
static cudaError_t launch_motion(const qt_params &P, const RolloutArgs &A, size_t smem, int mode,
                                 cudaStream_t st) {
    switch (ctrl_kind_of(P)) {
    case CK_PID: return launch_mode<MOTION, CK_PID>(P, A, smem, mode, st);
    case CK_LQI: return launch_mode<MOTION, CK_LQI>(P, A, smem, mode, st);
    default: return launch_mode<MOTION, CK_LQR>(P, A, smem, mode, st);
    }
}

}  // namespace qt

// Shared tail of qt_rollout / qt_rollout_sweep: argument checks, K staging, kernel-mode selection, launch.
// A arrives with the caller-specific fields (ic_period, scores, n_cand, eps, metrics, ...) already set.
static int rollout_launch(qt_context *ctx, const qt_params *params, qt::RolloutArgs &A, int32_t k_cols) {
    using namespace qt;
    QT_REQUIRE(ctx, A.E >= 0 && A.max_steps >= 0, "bad E / max_steps");
    QT_REQUIRE(ctx, k_cols == expected_k_cols(*params), "k_cols must be 6 (LQR), 9 (LQI) or 3 (PID)");
    QT_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(A.K) & 15) == 0, "K must be 16-byte aligned");
    QT_REQUIRE(ctx, params->motion_type >= 0 && params->motion_type <= QT_MOTION_FIGURE8, "motion_type");
    QT_REQUIRE(ctx, (A.feat == nullptr) == (A.actf == nullptr), "feat/act must come together");
    if (A.E == 0) return QT_OK;
    A.n_chains = (A.E + A.chain - 1) / A.chain;
    if (!params->feedforward_enabled) A.ff = nullptr;

    // the env's range checks can only trigger if the controller is allowed to exceed the env's limits
    A.env_bounds_check = !(params->ctrl_min_thrust >= params->min_thrust && params->ctrl_max_thrust <= params->max_thrust &&
                           params->ctrl_max_rate <= params->max_angular_rate && params->ctrl_max_rate >= 0.0);
    // K staging: with one episode per thread a block needs at most ceil(T / k_stride) + 1 groups
    A.k_in_smem = 0;
    A.k_smem_doubles = 0;
    if (A.chain == 1) {
        int64_t groups = (kRolloutThreads + A.k_stride - 1) / A.k_stride + 1;
        size_t doubles = ((size_t)groups * 4 * k_cols + 1) & ~(size_t)1;  // keep the cold tile 16-byte aligned
        if (doubles * sizeof(double) <= 24 * 1024) {
            A.k_in_smem = 1;
            A.k_smem_doubles = (int)doubles;
        }
    }
    // the per-thread cold-state tile (kColdRows x 128 doubles = 17 KB) + staged gains
    size_t smem = ((size_t)A.k_smem_doubles + (size_t)kColdRows * kRolloutThreads) * sizeof(double);
    if (A.feat) smem += (size_t)(kRolloutThreads / 32) * 32 * QT_FEATURE_DIM * sizeof(float);   // data-set staging tiles
    const bool needs_full = params->feedforward_enabled || params->integrator != QT_INTEGRATOR_RK4 || A.traj != nullptr;
    const int full = needs_full ? 1 : (A.feat != nullptr ? 2 : 0);   // kernel MODE

    DeviceGuard guard(ctx->device);
    cudaError_t err;
    switch (params->motion_type) {
    case QT_MOTION_LINEAR: err = launch_motion<QT_MOTION_LINEAR>(*params, A, smem, full, ctx->stream); break;
    case QT_MOTION_CIRCULAR: err = launch_motion<QT_MOTION_CIRCULAR>(*params, A, smem, full, ctx->stream); break;
    case QT_MOTION_SINUSOIDAL: err = launch_motion<QT_MOTION_SINUSOIDAL>(*params, A, smem, full, ctx->stream); break;
    case QT_MOTION_FIGURE8: err = launch_motion<QT_MOTION_FIGURE8>(*params, A, smem, full, ctx->stream); break;
    default: err = launch_motion<QT_MOTION_STATIONARY>(*params, A, smem, full, ctx->stream); break;
    }
    if (err != cudaSuccess) return fail(ctx, QT_ERR_CUDA, "rollout_kernel launch", cudaGetErrorString(err));
    ctx->launches += 1;
    return QT_OK;
}

extern "C" int qt_rollout(qt_context *ctx, const qt_params *params, int64_t E, int64_t max_steps,
                          int64_t chain, const double *state0_dev, const double *tparams_dev,
                          const double *K_dev, int32_t k_cols, int64_t k_stride,
                          double *metrics_out_dev, double *ctrl_state_io_dev,
                          double *final_state_out_dev, double *traj_out_dev, int64_t traj_stride,
                          int64_t traj_slots,
                          float *feat_out_dev, float *act_out_dev, int64_t data_stride,
                          int64_t data_slots, const double *ff_gains_dev) {
    using namespace qt;
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, params && state0_dev && tparams_dev && K_dev && metrics_out_dev, "null pointer");
    QT_REQUIRE(ctx, E >= 0 && max_steps >= 0 && max_steps < (1ll << 31), "bad E / max_steps");
    if (chain < 1) chain = 1;
    if (k_stride < 1) k_stride = 1;
    if (traj_stride < 1) traj_stride = 1;
    if (data_stride < 1) data_stride = 1;
    QT_REQUIRE(ctx, traj_stride < (1ll << 31) && data_stride < (1ll << 31) &&
                        traj_slots < (1ll << 31) && data_slots < (1ll << 31), "log strides/slots");
    RolloutArgs A{};
    A.E = E;
    A.chain = chain;
    A.k_stride = k_stride;
    A.max_steps = (int)max_steps;
    A.state0 = state0_dev;
    A.tparams = tparams_dev;
    A.K = K_dev;
    A.ff = ff_gains_dev;
    A.metrics = metrics_out_dev;
    A.ctrl_state = ctrl_state_io_dev;
    A.final_state = final_state_out_dev;
    A.traj = traj_out_dev;
    A.feat = feat_out_dev;
    A.actf = act_out_dev;
    A.traj_stride = (int)traj_stride;
    A.traj_slots = (int)traj_slots;
    A.data_stride = (int)data_stride;
    A.data_slots = (int)data_slots;
    return rollout_launch(ctx, params, A, k_cols);
}

// ControllerTuner's candidate batch in one launch (controllers/tuning.py:846-928): C candidates x eps episodes,
// episode (c, ep) = global index c * eps + ep starts from initial condition ep under gain K[c].
extern "C" int qt_rollout_sweep(qt_context *ctx, const qt_params *params, int64_t C, int64_t eps,
                                int64_t max_steps, const double *state0_dev, const double *tparams_dev,
                                const double *K_dev, int32_t k_cols, double *cand_out_dev,
                                double *metrics_out_dev, const double *ff_gains_dev) {
    using namespace qt;
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, params && state0_dev && tparams_dev && K_dev, "null pointer");
    QT_REQUIRE(ctx, cand_out_dev || metrics_out_dev, "cand_out and metrics_out are both NULL");
    QT_REQUIRE(ctx, C >= 0 && eps >= 1 && max_steps >= 0 && max_steps < (1ll << 31), "bad C / eps / max_steps");
    QT_REQUIRE(ctx, C <= (1ll << 40) / eps, "C * eps too large");
    if (C == 0) return QT_OK;
    RolloutArgs A{};
    A.E = C * eps;
    A.chain = 1;                // controller.reset() before every episode (tuning.py:879)
    A.k_stride = eps;
    A.max_steps = (int)max_steps;
    A.state0 = state0_dev;
    A.tparams = tparams_dev;
    A.K = K_dev;
    A.ff = ff_gains_dev;
    A.traj_stride = A.data_stride = 1;
    A.ic_period = eps;
    A.n_cand = C;
    A.eps = (int)eps;
    // the fused reduction needs whole candidates per 128-episode block; other episode counts write the metric rows
    // (into library scratch when the caller does not want them) and reduce with qt_reduce_scores
    const bool fused = cand_out_dev && eps <= kRolloutThreads && (kRolloutThreads % eps) == 0;
    A.scores = fused ? cand_out_dev : nullptr;
    A.metrics = metrics_out_dev;
    if (!fused && !metrics_out_dev) {
        DeviceGuard guard(ctx->device);
        const size_t need = (size_t)QT_METRICS_DIM * (size_t)A.E * sizeof(double);
        if (ctx->sweep_scratch_bytes < need) {
            if (ctx->sweep_scratch_dev) QT_CUDA_TRY(ctx, cudaFree(ctx->sweep_scratch_dev));
            ctx->sweep_scratch_dev = nullptr;
            ctx->sweep_scratch_bytes = 0;
            QT_CUDA_TRY(ctx, cudaMalloc(&ctx->sweep_scratch_dev, need));
            ctx->sweep_scratch_bytes = need;
        }
        A.metrics = reinterpret_cast<double *>(ctx->sweep_scratch_dev);
    }
    const double *metrics_for_reduce = A.metrics;
    int rc = rollout_launch(ctx, params, A, k_cols);
    if (rc != QT_OK) return rc;
    if (cand_out_dev && !fused) rc = qt_reduce_scores(ctx, metrics_for_reduce, C, eps, cand_out_dev);
    return rc;
}
