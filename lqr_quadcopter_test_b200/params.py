"""ctypes mirror of ``qt_params`` (include/qtrack.h) and the flattening of the reference-shaped
configuration objects into it."""
from __future__ import annotations

import ctypes as C

from .config import EnvConfig

CTRL_RICCATI, CTRL_PID = 0, 1
MOTION_TYPES = {"stationary": 0, "linear": 1, "circular": 2, "sinusoidal": 3, "figure8": 4}
INTEGRATORS = {"rk4": 0, "euler": 1}
TERMINATION_REASONS = {0: "", 1: "time_limit", 2: "position_bounds", 3: "numerical_instability"}

STATE_DIM, ACTION_DIM, TARGET_PARAMS, OBS_DIM = 12, 4, 12, 22
METRICS_DIM, TRAJ_FIELDS, FEATURE_DIM, INFO_DIM = 16, 24, 18, 4
# qt_params.debug_flags bits (include/qtrack.h)
DEBUG_GENERAL_PATH, DEBUG_POLICY_PER_THREAD, FLAG_POLICY_TENSOR_CORES = 1, 2, 4

# rows of the metric vector (enum qt_metric)
(M_STEPS, M_FINAL_TIME, M_ENV_ON_TARGET, M_MET_ON_TARGET, M_SUM_ERR, M_MAX_ERR, M_SUM_ERR2,
 M_SUM_EFFORT, M_OVERSHOOT_COUNT, M_MAX_OVERSHOOT, M_ENV_SUCCESS, M_MET_SUCCESS, M_TERMINATION,
 M_VIOLATIONS, M_SATURATIONS, M_FINAL_ERR) = range(16)

# rows of a trajectory record (enum qt_traj_field)
TRAJ_TIME, TRAJ_STATE, TRAJ_TARGET_P, TRAJ_TARGET_V, TRAJ_ACTION, TRAJ_ERR = 0, 1, 13, 16, 19, 23


class QtParams(C.Structure):
    _fields_ = [
        ("mass", C.c_double), ("gravity", C.c_double),
        ("drag_coeff_linear", C.c_double), ("drag_coeff_angular", C.c_double),
        ("max_thrust", C.c_double), ("min_thrust", C.c_double), ("max_angular_rate", C.c_double),
        ("dt", C.c_double), ("max_episode_time", C.c_double), ("max_velocity", C.c_double),
        ("max_angular_velocity", C.c_double), ("max_position", C.c_double),
        ("target_speed", C.c_double), ("target_amplitude", C.c_double),
        ("target_frequency", C.c_double), ("target_radius", C.c_double),
        ("target_center", C.c_double * 3), ("target_max_acceleration", C.c_double),
        ("env_target_radius", C.c_double), ("env_min_on_target_ratio", C.c_double),
        ("env_min_episode_duration", C.c_double),
        ("crit_target_radius", C.c_double), ("crit_min_on_target_ratio", C.c_double),
        ("crit_min_episode_duration", C.c_double),
        ("ctrl_dt", C.c_double), ("ctrl_hover_thrust", C.c_double), ("ctrl_max_thrust", C.c_double),
        ("ctrl_min_thrust", C.c_double), ("ctrl_max_rate", C.c_double),
        ("integral_limit", C.c_double), ("integral_zero_threshold", C.c_double),
        ("ff_velocity_gain", C.c_double * 3), ("ff_acceleration_gain", C.c_double * 3),
        ("ff_max_velocity", C.c_double), ("ff_max_acceleration", C.c_double),
        ("motion_type", C.c_int32), ("integrator", C.c_int32), ("use_lqi", C.c_int32),
        ("feedforward_enabled", C.c_int32), ("overshoot_window", C.c_int32),
        ("controller_type", C.c_int32), ("debug_flags", C.c_int32), ("reserved2", C.c_int32),
    ]

    def copy(self) -> "QtParams":
        out = QtParams()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(QtParams))
        return out


MLP_MAX_LAYERS = 6
MLP_MAX_WIDTH = 256
ACTIVATIONS = {"relu": 0, "tanh": 1, "elu": 2, "leaky_relu": 3}


class QtMlp(C.Structure):
    """Mirror of ``struct qt_mlp`` (include/qtrack.h)."""

    _fields_ = [
        ("n_layers", C.c_int32), ("dims", C.c_int32 * (MLP_MAX_LAYERS + 1)), ("activation", C.c_int32),
        ("reserved", C.c_int32), ("out_lo", C.c_float * 4), ("out_scale", C.c_float * 4),
        ("weights_dev", C.c_void_p),
    ]


def default_params() -> QtParams:
    """Reference defaults; identical to qt_params_default() in the library (checked in tests)."""
    return make_params(EnvConfig())


def make_params(env_config: EnvConfig | None = None, controller=None, criteria=None) -> QtParams:
    """Flatten (EnvConfig, controller, SuccessCriteria) into the kernels' parameter block.

    ``controller`` is any object with the attributes RiccatiLQRController exposes
    (controllers/riccati_lqr.py:456-502); ``criteria`` is the metrics SuccessCriteria the driver
    passes to compute_episode_metrics (utils/metrics.py:26-39).  Either may be None, in which case
    the reference defaults apply.
    """
    cfg = env_config or EnvConfig()
    p = QtParams()
    q, s, t, sc = cfg.quadcopter, cfg.simulation, cfg.target, cfg.success_criteria
    p.mass, p.gravity = q.mass, q.gravity
    p.drag_coeff_linear, p.drag_coeff_angular = q.drag_coeff_linear, q.drag_coeff_angular
    p.max_thrust, p.min_thrust, p.max_angular_rate = q.max_thrust, q.min_thrust, q.max_angular_rate
    p.dt, p.max_episode_time = s.dt, s.max_episode_time
    p.max_velocity, p.max_angular_velocity, p.max_position = (
        s.max_velocity, s.max_angular_velocity, s.max_position)
    motion = t.motion_type.lower()
    if motion not in MOTION_TYPES:
        # same message shape as env/target_motion.py:310-314
        raise ValueError(f"Invalid motion type: {motion}. Valid types: {set(MOTION_TYPES)}")
    integrator = s.integrator
    p.motion_type = MOTION_TYPES[motion]
    # quadcopter_env.py:309-312: anything that is not "euler" integrates with RK4
    p.integrator = INTEGRATORS["euler"] if integrator == "euler" else INTEGRATORS["rk4"]
    p.target_speed, p.target_amplitude = t.speed, t.amplitude
    p.target_frequency, p.target_radius = t.frequency, t.radius
    for i in range(3):
        p.target_center[i] = float(t.center[i])
    p.target_max_acceleration = t.max_acceleration
    p.env_target_radius = sc.target_radius
    p.env_min_on_target_ratio = sc.min_on_target_ratio
    p.env_min_episode_duration = sc.min_episode_duration
    if criteria is None:
        p.crit_target_radius, p.crit_min_on_target_ratio, p.crit_min_episode_duration = 0.5, 0.8, 30.0
    else:
        p.crit_target_radius = criteria.target_radius
        p.crit_min_on_target_ratio = criteria.min_on_target_ratio
        p.crit_min_episode_duration = criteria.min_episode_duration
    p.overshoot_window = 10
    # controller block
    p.ctrl_dt, p.ctrl_hover_thrust = 0.01, 1.0 * 9.81
    p.ctrl_max_thrust, p.ctrl_min_thrust, p.ctrl_max_rate = 20.0, 0.0, 3.0
    p.integral_limit, p.integral_zero_threshold = 10.0, 0.01
    p.ff_max_velocity, p.ff_max_acceleration = 10.0, 5.0
    if controller is not None:
        p.controller_type = CTRL_PID if getattr(controller, "name", "") == "pid" else CTRL_RICCATI
        p.ctrl_dt = controller.dt
        p.ctrl_hover_thrust = controller.hover_thrust
        p.ctrl_max_thrust, p.ctrl_min_thrust = controller.max_thrust, controller.min_thrust
        p.ctrl_max_rate = controller.max_rate
        p.use_lqi = 1 if controller.use_lqi else 0
        p.integral_limit = controller.integral_limit
        p.integral_zero_threshold = controller.integral_zero_threshold
        p.feedforward_enabled = 1 if controller.feedforward_enabled else 0
        for i in range(3):
            p.ff_velocity_gain[i] = float(controller.ff_velocity_gain[i])
            p.ff_acceleration_gain[i] = float(controller.ff_acceleration_gain[i])
        p.ff_max_velocity = controller.ff_max_velocity
        p.ff_max_acceleration = controller.ff_max_acceleration
    return p
