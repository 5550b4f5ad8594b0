/*
 * qtrack.h — C ABI of libqtrack: batched closed-loop quadcopter tracking on one B200.
 *
 * This is the drop-in boundary for the hot path of `quadcopter_tracking` (the reference,
 * cited below as file:line relative to its src/quadcopter_tracking/ directory).  The
 * reference has no FFI of its own (it is pure Python); every entry point here replaces a
 * Python call site, which is named next to it.  The reference-side binding is a ctypes
 * stub — see INTEGRATION.md.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  No torch / Python types.
 *   - Every function returns 0 on success, a negative QT_ERR_* otherwise; the message is
 *     available from qt_last_error(ctx).
 *   - Pointers named *_dev are DEVICE pointers owned by the caller (e.g. torch CUDA tensors),
 *     pointers named *_host are HOST pointers.  The library never frees caller memory.
 *   - All batched arrays are struct-of-arrays with the episode index fastest:
 *     a "[F][E]" array stores field f of episode e at  base[f * E + e].
 *   - Work is enqueued on the context's stream (qt_set_stream / qt_get_stream); the *_dev
 *     entry points do not synchronise, the *_host entry points return after their result
 *     has landed in host memory.
 *   - Threads: a context is one GPU + one stream + library scratch.  The *_host entry points of one
 *     context serialise on an internal mutex (they pipeline over the context's two streams and share
 *     its scratch); the *_dev entry points are not locked — use one context per host thread.
 *   - There is no CPU fallback: without a CUDA device qt_create fails.
 */
#ifndef QTRACK_H
#define QTRACK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QT_ABI_VERSION 6

#define QT_STATE_DIM 12     /* env/quadcopter_env.py:63-70  [p(3) v(3) roll,pitch,yaw  p,q,r] */
#define QT_ACTION_DIM 4     /* thrust, roll_rate, pitch_rate, yaw_rate (controllers/base.py:28) */
#define QT_TARGET_PARAMS 12 /* per-episode target-pattern parameters, layout below */
#define QT_OBS_DIM 22       /* quad state 12 + target p,v,a 9 + time (quadcopter_env.py:472-496) */
#define QT_METRICS_DIM 16   /* per-episode result vector, QT_M_* below */
#define QT_TRAJ_FIELDS 24   /* per logged step, QT_TRAJ_* below */
#define QT_FEATURE_DIM 18   /* controllers/deep_tracking_policy.py:241-286 */
#define QT_INFO_DIM 4       /* tracking_error, on_target, on_target_ratio, success */
#define QT_MAX_N 16         /* DARE: largest state dimension */
#define QT_MAX_M 8          /* DARE: largest input dimension */

/* env/target_motion.py:264-270 */
enum qt_motion {
    QT_MOTION_STATIONARY = 0, /* tparams[0..2] = position                                  */
    QT_MOTION_LINEAR = 1,     /* [0..2] start, [3..5] velocity = unit direction * speed      */
    QT_MOTION_CIRCULAR = 2,   /* [0..2] centre, [3] radius, [4] omega = speed/radius, [5] angle0 */
    QT_MOTION_SINUSOIDAL = 3, /* [0..2] centre, [3..5] amplitude, [6..8] omega, [9..11] phase */
    QT_MOTION_FIGURE8 = 4     /* [0..2] centre, [3] scale, [4] omega = speed/scale           */
};

enum qt_integrator { QT_INTEGRATOR_RK4 = 0, QT_INTEGRATOR_EULER = 1 }; /* quadcopter_env.py:309-312 */

/* Control law evaluated inside the kernels.
 *   QT_CTRL_RICCATI  u = K e (+ K_i I)          controllers/riccati_lqr.py:779-967 (also the heuristic
 *                    LQRController, controllers/__init__.py:576-686: same law, closed-form K)
 *                    gains: K row-major [4][6] (use_lqi = 0) or [4][9] (use_lqi = 1)
 *   QT_CTRL_PID      per-axis P/I/D on the position error, controllers/__init__.py:243-379
 *                    gains: row-major [4][3] = kp | ki | kd | (unused row)                         */
enum qt_controller { QT_CTRL_RICCATI = 0, QT_CTRL_PID = 1 };

/* qt_params.debug_flags.  QT_DEBUG_GENERAL_PATH: qt_rollout always runs the general step loop, never
 * the structure-exploiting one (zero yaw, axis-decoupled gain); results must not change (tests).  */
#define QT_DEBUG_GENERAL_PATH 1
/* QT_DEBUG_POLICY_PER_THREAD: qt_rollout_policy evaluates the network one episode per thread even where
 * the warp-cooperative tiled evaluation applies; results must not change (tests).                   */
#define QT_DEBUG_POLICY_PER_THREAD 2
/* QT_FLAG_POLICY_TENSOR_CORES (ABI v6): qt_rollout_policy evaluates the hidden layers of a network with one to three
 * hidden layers of at most 64 units each (the reference's default 64-64, its "fast" 32-32) on the tensor cores — tcgen05.mma kind::tf32 with
 * the 3xTF32 split (x = hi + lo), fp32 accumulation in tensor memory; the output layer stays on the FP32 pipe.  The
 * result differs from the FP32-pipe evaluation by about 2^-21 relative per product, the size of fp32 summation-order
 * noise; the default (flag clear) is the FP32-pipe evaluation, which the parity fixtures are held to bit for bit.  */
#define QT_FLAG_POLICY_TENSOR_CORES 4

/* quadcopter_env.py:513-535; 0 = the step cap ended the loop, not the environment */
enum qt_termination {
    QT_TERM_NONE = 0,
    QT_TERM_TIME_LIMIT = 1,
    QT_TERM_POSITION_BOUNDS = 2,
    QT_TERM_NUMERICAL_INSTABILITY = 3
};

/* rows of the per-episode result vector written by qt_rollout ([QT_METRICS_DIM][E]) */
enum qt_metric {
    QT_M_STEPS = 0,             /* env steps taken                                            */
    QT_M_FINAL_TIME = 1,        /* info["time"] of the last step (fp64-accumulated)           */
    QT_M_ENV_ON_TARGET = 2,     /* quadcopter_env.py:205-207, post-step samples 1..T          */
    QT_M_MET_ON_TARGET = 3,     /* utils/metrics.py:165-179, pre-step samples 0..T-1          */
    QT_M_SUM_ERR = 4,           /* sum of pre-step tracking errors (metrics.py:300, 328)      */
    QT_M_MAX_ERR = 5,           /* metrics.py:329                                             */
    QT_M_SUM_ERR2 = 6,          /* metrics.py:330                                             */
    QT_M_SUM_EFFORT = 7,        /* sum ||action||_2, metrics.py:182-202                       */
    QT_M_OVERSHOOT_COUNT = 8,   /* metrics.py:205-261                                         */
    QT_M_MAX_OVERSHOOT = 9,
    QT_M_ENV_SUCCESS = 10,      /* quadcopter_env.py:537-553 (only meaningful when terminated) */
    QT_M_MET_SUCCESS = 11,      /* metrics.py:312-316                                         */
    QT_M_TERMINATION = 12,      /* enum qt_termination                                        */
    QT_M_VIOLATIONS = 13,       /* steps whose action the env had to repair (:174-181)        */
    QT_M_SATURATIONS = 14,      /* controllers/riccati_lqr.py:924-931                         */
    QT_M_FINAL_ERR = 15         /* post-step tracking error of the last step                  */
};

/* rows of one logged step ([slots][QT_TRAJ_FIELDS][E]); mirrors the record eval.py:142-159 builds */
enum qt_traj_field {
    QT_TRAJ_TIME = 0,     /* info["time"] after the step                     */
    QT_TRAJ_STATE = 1,    /* 12 rows: quadcopter state BEFORE the step        */
    QT_TRAJ_TARGET_P = 13,/* 3 rows: target position before the step          */
    QT_TRAJ_TARGET_V = 16,/* 3 rows: target velocity before the step          */
    QT_TRAJ_ACTION = 19,  /* 4 rows: controller output                        */
    QT_TRAJ_ERR = 23      /* tracking error after the step (reward = -err)    */
};

enum qt_error {
    QT_OK = 0,
    QT_ERR_ARG = -1,
    QT_ERR_CUDA = -2,
    QT_ERR_NO_DEVICE = -3,
    QT_ERR_UNSUPPORTED = -4
};

/*
 * Everything the fused step needs, flattened from the reference's config dataclasses and the
 * controller's constructor arguments.  All doubles first, then int32s; no padding surprises
 * (qt_params_size() lets a binding verify its mirror of this struct).
 */
typedef struct qt_params {
    /* env/config.py:11-37 QuadcopterParams (inertia, arm length, k_thrust, k_torque are unused
       by the reference dynamics, quadcopter_env.py:329-426) */
    double mass, gravity, drag_coeff_linear, drag_coeff_angular;
    double max_thrust, min_thrust, max_angular_rate;
    /* env/config.py:40-51 SimulationParams */
    double dt, max_episode_time, max_velocity, max_angular_velocity, max_position;
    /* env/config.py:54-65 TargetParams (used by qt_reset_batched and the accel clamp) */
    double target_speed, target_amplitude, target_frequency, target_radius;
    double target_center[3];
    double target_max_acceleration;
    /* env/config.py:68-74: the env's own success criteria (info["success"]) */
    double env_target_radius, env_min_on_target_ratio, env_min_episode_duration;
    /* utils/metrics.py:26-39: the criteria object the driver passes to compute_episode_metrics */
    double crit_target_radius, crit_min_on_target_ratio, crit_min_episode_duration;
    /* controllers/riccati_lqr.py:456-502 */
    double ctrl_dt, ctrl_hover_thrust, ctrl_max_thrust, ctrl_min_thrust, ctrl_max_rate;
    double integral_limit, integral_zero_threshold;
    double ff_velocity_gain[3], ff_acceleration_gain[3];
    double ff_max_velocity, ff_max_acceleration;   /* integral_limit is shared with the PID law (:212) */
    int32_t motion_type;          /* enum qt_motion, one per launch                         */
    int32_t integrator;           /* enum qt_integrator                                     */
    int32_t use_lqi;              /* riccati_lqr.py:474                                     */
    int32_t feedforward_enabled;  /* riccati_lqr.py:494                                     */
    int32_t overshoot_window;     /* metrics.py:208, default 10                             */
    int32_t controller_type;      /* enum qt_controller                                     */
    int32_t debug_flags;          /* QT_DEBUG_* bits; 0 in production                       */
    int32_t reserved2;
} qt_params;

typedef struct qt_context qt_context;

/* ---- context ---------------------------------------------------------------------------- */
int qt_abi_version(void);
int64_t qt_params_size(void);
/* Fill *p with the reference defaults (env/config.py, riccati_lqr.py:456-502, metrics.py:26-39). */
void qt_params_default(qt_params *p);
int qt_create(int device, qt_context **out);
void qt_destroy(qt_context *ctx);
const char *qt_last_error(const qt_context *ctx);
/* Enqueue on an existing CUDA stream (cudaStream_t as void*, e.g. torch.cuda.current_stream().cuda_stream).
   NULL means CUDA's legacy default stream (what torch uses by default).  Until the first call a
   context uses a private non-blocking stream of its own. */
int qt_set_stream(qt_context *ctx, void *cuda_stream);
void *qt_get_stream(const qt_context *ctx);
int qt_synchronize(qt_context *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t qt_launch_count(const qt_context *ctx);
/* Which kernels qt_dare_batched (and the host / sweep calls built on it) may use (ABI v6).  QT_DARE_PATH_AUTO:
   systems that are block diagonal after a permutation of the states, with blocks of at most three states and one
   control, diagonal Q and R — every design of riccati_lqr.py:187-316, 602-700 — are solved block by block in
   registers; all others by the dense kernel.  QT_DARE_PATH_DENSE: the dense kernel for every system (tests and
   measurements).  Both paths return the same solution. */
#define QT_DARE_PATH_AUTO 0
#define QT_DARE_PATH_DENSE 1
int qt_set_dare_path(qt_context *ctx, int32_t path);

/* ---- A1: QuadcopterEnv.reset(seed) + TargetMotion.reset/_create_pattern ------------------
 * quadcopter_env.py:111-150, target_motion.py:291-369.  One seed per episode; reproduces
 * numpy's default_rng(seed) (SeedSequence -> PCG64 -> uniform / ziggurat standard_normal) on
 * the device.  state_out [12][E], tparams_out [12][E].
 * draws_io [2][E] (or NULL) carries, per episode, how many raw 64-bit outputs the env stream
 * (row 0) and the target stream (row 1) have already consumed; it is updated on return.  This
 * reproduces reset(seed=None), which keeps drawing from the existing generators
 * (quadcopter_env.py:122-130): pass the counters back in with the same seeds.  NULL = fresh
 * generators, i.e. reset(seed).                                                              */
int qt_reset_batched(qt_context *ctx, const qt_params *params, const uint64_t *seeds_dev,
                     int64_t E, uint64_t *draws_io_dev, double *state_out_dev,
                     double *tparams_out_dev);

/* seeds_out[i] = seed0 + i, the seed sequence of Evaluator.evaluate (eval.py:200: base_seed + i) and of the tuner
 * (controllers/tuning.py:875: config.seed + ep), written on the device.                       */
int qt_fill_seeds(qt_context *ctx, uint64_t *seeds_out_dev, int64_t E, uint64_t seed0);

/* Raw generator access for tests: for each seed, n_uniform doubles from Generator.random()
 * followed (same stream) by n_normal draws of standard_normal().  out [n_uniform+n_normal][E]. */
int qt_rng_probe(qt_context *ctx, const uint64_t *seeds_dev, int64_t E, int32_t n_uniform,
                 int32_t n_normal, double *out_dev);

/* ---- A7: QuadcopterEnv._get_observation ---------------------------------------------------
 * obs_out [22][E]: state 12, target position 3, velocity 3, acceleration 3 (clamped,
 * target_motion.py:403-405), time.                                                           */
int qt_observe_batched(qt_context *ctx, const qt_params *params, int64_t E,
                       const double *state_dev, const double *time_dev,
                       const double *tparams_dev, double *obs_out_dev);

/* ---- A10: RiccatiLQRController.compute_action ---------------------------------------------
 * riccati_lqr.py:779-967 for a batch of observations.  K is row-major [groups][4][k_cols],
 * 16-byte aligned, k_cols = 6 (LQR) or 9 (LQI: K_pd | K_i); episode e uses group e / k_stride.
 * integral_io [3][E] and saturation_io [E] hold the controller state (NULL when !use_lqi /
 * not wanted).  For QT_CTRL_PID, K is [groups][4][3] (k_cols = 3) and integral_io is [4][E]:
 * rows 0-2 the integral error, row 3 the time of the previous call (NaN = none yet, i.e.
 * PIDController._last_time is None).  action_out [4][E].                                     */
int qt_compute_action_batched(qt_context *ctx, const qt_params *params, int64_t E,
                              const double *obs_dev, const double *K_dev, int32_t k_cols,
                              int64_t k_stride, double *integral_io_dev,
                              int64_t *saturation_io_dev, double *action_out_dev);

/* ---- A2-A8: QuadcopterEnv.step -------------------------------------------------------------
 * quadcopter_env.py:152-232 for all episodes.  If action_in is NULL the Riccati controller
 * (above) supplies the action from the current state.  counters_io [3][E] = step_count,
 * on_target_count, steps-with-violations.  Outputs (any may be NULL): obs_out [22][E]
 * post-step, reward_out [E], done_out [E] (enum qt_termination, 0 = running), info_out [4][E]
 * = tracking_error, on_target, on_target_ratio, success, action_out [4][E] = action after the
 * env's validation.  Episodes whose done flag is already set are NOT skipped (the reference
 * env keeps stepping if asked to).                                                          */
int qt_step_batched(qt_context *ctx, const qt_params *params, int64_t E, double *state_io_dev,
                    double *time_io_dev, const double *tparams_dev, int64_t *counters_io_dev,
                    const double *action_in_dev, const double *K_dev, int32_t k_cols,
                    int64_t k_stride, double *integral_io_dev, int64_t *saturation_io_dev,
                    double *obs_out_dev, double *reward_out_dev, int32_t *done_out_dev,
                    double *info_out_dev, double *action_out_dev);

/* ---- the fused hot loop: eval.py:119-165 / controllers/tuning.py:885-906 -------------------
 * For each episode: while not done and step < max_steps: compute_action; env.step; record.
 * One thread per chain of `chain` consecutive episodes (chain = 1 resets the controller per
 * episode — tuner semantics, tuning.py:879; chain = E reproduces Evaluator.evaluate, which
 * never resets the controller between episodes, eval.py:198-213).
 *   state0 [12][E], tparams [12][E], K [groups][4][k_cols], metrics_out [16][E],
 *   ctrl_state_io [4][n_chains] or NULL (n_chains = ceil(E / chain)): LQI integrator (3 rows)
 *   and saturation count of each chain's controller (PID: integral error and, in row 3, the
 *   time of the previous call, NaN = none); read as the initial controller state
 *   when chain > 1 (Evaluator semantics: whatever state the controller object carries),
 *   ignored on entry when chain == 1 (reset per episode); always written back on exit,
 *   final_state_out [12][E] or NULL,
 *   traj_out [slots][24][E] or NULL: steps k with k % traj_stride == 0, slot = k / traj_stride,
 *   feat_out f32 [slots][E][18] and act_out f32 [slots][E][4] or NULL (teacher data set,
 *   deep_tracking_policy.py:241-286 / train.py:683-696), every data_stride-th step,
 *   ff_gains [groups][6] or NULL: per-group feedforward gains (velocity x,y,z then acceleration
 *   x,y,z) replacing params->ff_velocity_gain / ff_acceleration_gain when feedforward is enabled -
 *   a tuner batch whose candidates differ in those gains (tuning.py:700-712); same grouping as K. */
int qt_rollout(qt_context *ctx, const qt_params *params, int64_t E, int64_t max_steps,
               int64_t chain, const double *state0_dev, const double *tparams_dev,
               const double *K_dev, int32_t k_cols, int64_t k_stride, double *metrics_out_dev,
               double *ctrl_state_io_dev, double *final_state_out_dev, double *traj_out_dev,
               int64_t traj_stride,
               int64_t traj_slots, float *feat_out_dev, float *act_out_dev,
               int64_t data_stride, int64_t data_slots, const double *ff_gains_dev);

/* ---- A14 fused: one ControllerTuner candidate batch (controllers/tuning.py:846-928) ----------
 * C candidates x eps episodes in one launch.  Episode (c, ep) — global index c * eps + ep — starts from initial
 * condition ep: state0 [12][eps], tparams [12][eps] are shared by all candidates, because every candidate sees the
 * same seeds config.seed + ep (tuning.py:875); the controller is reset before every episode (tuning.py:879).
 * K [C][4][k_cols], ff_gains [C][6] or NULL (as qt_rollout).
 *   cand_out [4][C] or NULL: score, mean_on_target_ratio, mean_tracking_error, success_rate — exactly what
 *     qt_reduce_scores computes from the metric rows (numpy's summation order, tuning.py:912-926).  When eps divides
 *     128 the reduction runs in the rollout kernel's epilogue (a 128-episode block holds whole candidates) and the
 *     16 metric rows per episode are never written; otherwise the rows go through qt_reduce_scores.
 *   metrics_out [16][C*eps] or NULL: the per-episode rows, if the caller wants them as well.     */
int qt_rollout_sweep(qt_context *ctx, const qt_params *params, int64_t C, int64_t eps, int64_t max_steps,
                     const double *state0_dev, const double *tparams_dev, const double *K_dev,
                     int32_t k_cols, double *cand_out_dev, double *metrics_out_dev,
                     const double *ff_gains_dev);

/* ---- A11: solve_dare -----------------------------------------------------------------------
 * riccati_lqr.py:119-184 for `batch` systems: P solves A'PA - P - A'PB(R+B'PB)^-1 B'PA + Q = 0,
 * K = (R+B'PB)^-1 B'PA.  Row-major fp64; A [n][n], B [n][m], Q [n][n], R [m][m] per system with
 * strides in doubles between consecutive systems (stride 0 = shared by all systems).
 * P_out [batch][n][n], K_out [batch][m][n] (either may be NULL), iters_out/info_out [batch].
 * info: 0 ok, 1 not converged, 2 singular pivot, 3 Q not symmetric PSD / R not symmetric PD
 * (riccati_lqr.py:57-116 tolerances).  n <= QT_MAX_N, m <= QT_MAX_M.                          */
int qt_dare_batched(qt_context *ctx, int32_t n, int32_t m, int64_t batch, const double *A_dev,
                    int64_t strideA, const double *B_dev, int64_t strideB, const double *Q_dev,
                    int64_t strideQ, const double *R_dev, int64_t strideR, double *P_out_dev,
                    double *K_out_dev, int32_t *iters_out_dev, int32_t *info_out_dev);

/* Build the reference's design matrices on the device from per-candidate diagonal weights
 * (riccati_lqr.py:187-316, 602-700): q [batch][n] (q_pos, q_vel[, q_int]), r [batch][4];
 * writes A [n][n] and B [n][4] once (shared) and Q [batch][n][n], R [batch][4][4].           */
int qt_build_lqr_systems(qt_context *ctx, int32_t n, int64_t batch, double dt, double mass,
                         double gravity, const double *q_diag_dev, const double *r_diag_dev,
                         double *A_out_dev, double *B_out_dev, double *Q_out_dev,
                         double *R_out_dev);

/* ---- A14: ControllerTuner._evaluate_config aggregation (tuning.py:912-926) ------------------
 * metrics [16][C*eps] from qt_rollout with episode index = c*eps + ep.
 * cand_out [4][C] = score, mean_on_target_ratio, mean_tracking_error, success_rate.          */
int qt_reduce_scores(qt_context *ctx, const double *metrics_dev, int64_t C, int64_t eps,
                     double *cand_out_dev);

/* ---- A13: compute_evaluation_summary partial sums (metrics.py:341-390) -----------------------
 * sums_out [16] on the device: n, successes, sum ratio, sum ratio^2, sum mean_err, sum mean_err^2,
 * sum mean_effort, max ratio, argmax (first), min ratio, argmin (first), ...;
 * ratios use the metrics accounting (pre-step samples).                                       */
int qt_reduce_summary(qt_context *ctx, const double *metrics_dev, int64_t E, double *sums_out_dev);

/* ---- host-buffer convenience: Evaluator.evaluate(num_episodes, base_seed) -------------------
 * eval.py:169-231 with HOST buffers: seeds_host [E] in, K_host [4][k_cols] in,
 * metrics_host [16][E] out.  Does H2D, reset, rollout, D2H and returns when metrics_host is valid.
 * Large batches (E >= 2^17, chain = 1) run as four chunks that alternate between the context's
 * stream and a second stream owned by the context, so that a chunk's read-back overlaps the next
 * chunks' kernels; results do not depend on the chunking.  Pinned host buffers make the copies
 * asynchronous; pageable ones work but serialise.                                               */
int qt_evaluate_host(qt_context *ctx, const qt_params *params, const uint64_t *seeds_host,
                     int64_t E, int64_t max_steps, int64_t chain, const double *K_host,
                     int32_t k_cols, double *metrics_host);

/* ---- host-buffer convenience: solve_dare for a batch of HOST matrices (riccati_lqr.py:119-184) --
 * Same arguments as qt_dare_batched with host pointers; strides must be 0 (one matrix shared by all systems) or
 * dense (n*n, n*m, m*m).  P_out_host, K_out_host and iters_out_host may be NULL.  Does H2D, validation, solve,
 * gain, D2H in four chunks alternating between the context's two streams and returns when the outputs are valid. */
int qt_dare_host(qt_context *ctx, int32_t n, int32_t m, int64_t batch, const double *A_host,
                 int64_t strideA, const double *B_host, int64_t strideB, const double *Q_host,
                 int64_t strideQ, const double *R_host, int64_t strideR, double *P_out_host,
                 double *K_out_host, int32_t *iters_out_host, int32_t *info_out_host);

/* ---- host-buffer convenience: a whole Q/R sweep (controllers/tuning.py:846-928 for C candidates) --
 * q_host [C][n] = per-candidate diagonal of Q (n = 6: q_pos, q_vel; n = 9 with params->use_lqi: + q_int),
 * r_host [C][4] = r_controls.  Builds the reference's design matrices with params->ctrl_dt, ctrl_mass and
 * ctrl_gravity (riccati_lqr.py:187-316), solves the C DAREs, resets eps episodes with seeds seed0 + ep, rolls
 * out C x eps episodes of at most max_steps steps and reduces them.  cand_out_host [4][C] (as qt_reduce_scores),
 * K_out_host [C][4][n] or NULL, info_out_host [C] = the DARE info codes: a candidate with info != 0 has no valid
 * row in cand_out — the caller applies the reference's fallback rule to it (riccati_lqr.py:737-777).   */
int qt_sweep_host(qt_context *ctx, const qt_params *params, int32_t n, int64_t C, int64_t eps,
                  uint64_t seed0, int64_t max_steps, const double *q_host, const double *r_host,
                  double ctrl_mass, double ctrl_gravity, double *cand_out_host, double *K_out_host,
                  int32_t *info_out_host);

/* ---- the deep tracking policy in the loop ------------------------------------
 * controllers/deep_tracking_policy.py:24-147 (PolicyNetwork): 18 features -> hidden layers with one
 * activation -> 4 outputs, action_i = sigmoid(raw_i) * (high_i - low_i) + low_i, all in fp32; the
 * features are the fp32 cast of [p_t - p_q, v_t - v_q, attitude, body rates, p_t, v_t]
 * (deep_tracking_policy.py:241-286).
 * Packed weights (fp32, device): for each linear layer l = 0..n_layers-1 the TRANSPOSED weight
 * Wt[rows_l][pad4(dims[l+1])] followed by bias[pad4(dims[l+1])], where pad4 rounds up to a multiple
 * of 4, rows_0 = dims[0] and rows_l = pad4(dims[l]) for l > 0; padding entries are zero (an
 * activation of 0 is 0 for all four activation kinds, so padded units are inert).
 * qt_mlp_packed_floats() returns the buffer length.  weights_dev must be 16-byte aligned.      */
#define QT_MLP_MAX_LAYERS 6
#define QT_MLP_MAX_WIDTH 256
enum qt_activation { QT_ACT_RELU = 0, QT_ACT_TANH = 1, QT_ACT_ELU = 2, QT_ACT_LEAKY_RELU = 3 };
typedef struct qt_mlp {
    int32_t n_layers;                    /* linear layers = hidden layers + 1                  */
    int32_t dims[QT_MLP_MAX_LAYERS + 1]; /* dims[0] = 18, ..., dims[n_layers] = 4              */
    int32_t activation;                  /* enum qt_activation (leaky_relu slope 0.1)          */
    int32_t reserved;
    float out_lo[4], out_scale[4];       /* low_i and (high_i - low_i), cast to fp32           */
    const float *weights_dev;
} qt_mlp;
int64_t qt_mlp_packed_floats(const qt_mlp *policy);

/* DeepTrackingPolicy.compute_action for a batch: obs [22][E] (fp64) -> action_out [4][E] (the fp32
 * network outputs widened to fp64, as `float(action_array[i])` does).                           */
int qt_policy_action_batched(qt_context *ctx, const qt_mlp *policy, int64_t E, const double *obs_dev,
                             double *action_out_dev);

/* The fused loop with the policy network as the acting controller (eval.py:119-165 with
 * --controller deep; train.py:654-735 when a supervisor is given).  Same outputs as qt_rollout.
 * If teacher_K_dev is not NULL, the control law selected by params (Riccati-LQR / LQI / PID, reset
 * before every episode as train.py:665-667 does) is evaluated on the states the policy visits, and
 * act_out receives ITS actions (the imitation labels); otherwise act_out receives the policy's.  */
int qt_rollout_policy(qt_context *ctx, const qt_params *params, int64_t E, int64_t max_steps,
                      const double *state0_dev, const double *tparams_dev, const qt_mlp *policy,
                      const double *teacher_K_dev, int32_t k_cols, int64_t k_stride,
                      double *metrics_out_dev, double *final_state_out_dev, double *traj_out_dev,
                      int64_t traj_stride, int64_t traj_slots, float *feat_out_dev, float *act_out_dev,
                      int64_t data_stride, int64_t data_slots);

/* Self-test hook: the rollout's two-at-a-time square root (csrc/qt_step.cuh, sqrt_pair) on arbitrary
 * arguments; must equal IEEE-754 sqrt bit for bit (tests/test_guards_gpu.py).                  */
int qt_sqrt_pair_probe(qt_context *ctx, int64_t n, const double *a_dev, const double *b_dev,
                       double *ra_out_dev, double *rb_out_dev);

/* ---- measurement helper: register-resident DFMA loop, reports the achieved FP64 rate --------
 * Used as the FP64 roofline denominator by bench.py (MEASURED_PEAKS.json has no FP64 entry). */
int qt_fp64_peak_probe(qt_context *ctx, int64_t iters, double *tflops_out_host,
                       double *ms_out_host);

/* The same for the FP32 pipe (register-resident FFMA chains): denominator of the policy kernel's roofline. */
int qt_fp32_peak_probe(qt_context *ctx, int64_t iters, double *tflops_out_host,
                       double *ms_out_host);

#ifdef __cplusplus
}
#endif
#endif /* QTRACK_H */
