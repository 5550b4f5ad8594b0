"""Riccati-LQR / LQI controller backed by the batched CUDA DARE solver and control-law kernels.

Drop-in for ``quadcopter_tracking.controllers.riccati_lqr`` (cited below as riccati_lqr.py:line):
the constructor takes the same configuration dictionary, ``compute_action(obs)`` takes the same
observation dictionary and returns the same action dictionary, the same attributes and getters
exist, the same exceptions are raised.  What differs is where the arithmetic runs: the DARE is
solved by ``qt_dare_batched`` and the control law is evaluated by ``qt_compute_action_batched``
(or, inside rollouts, by the fused ``qt_rollout`` kernel) on the GPU.  Host-side numpy is used
only to assemble the small design matrices and to validate Q/R exactly as the reference does.
"""
from __future__ import annotations

import logging

from dataclasses import dataclass

import numpy as np

from . import _lib
from .params import QtParams, make_params

logger = logging.getLogger(__name__)

ACTION_KEYS = ("thrust", "roll_rate", "pitch_rate", "yaw_rate")  # controllers/base.py:28


# --------------------------------------------------------------------------- design matrices --
def build_linearized_system(dt: float, mass: float = 1.0, gravity: float = 9.81):
    """riccati_lqr.py:187-263: A = I + A_c dt, B = B_c dt for the 6-state double integrator
    (x, y, z, vx, vy, vz) driven by (thrust_delta, roll_rate, pitch_rate, yaw_rate)."""
    a_c = np.zeros((6, 6))
    a_c[[0, 1, 2], [3, 4, 5]] = 1.0
    b_c = np.zeros((6, 4))
    b_c[5, 0] = 1.0 / mass
    b_c[4, 1] = -gravity
    b_c[3, 2] = gravity
    return np.eye(6) + a_c * dt, b_c * dt


def build_augmented_lqi_system(A: np.ndarray, B: np.ndarray, dt: float):
    """riccati_lqr.py:266-316: append three integrator states, int += dt * position_error."""
    n, m = A.shape[0], B.shape[1]
    a_aug = np.zeros((n + 3, n + 3))
    a_aug[:n, :n] = A
    a_aug[n:, :3] = np.eye(3) * dt
    a_aug[n:, n:] = np.eye(3)
    b_aug = np.zeros((n + 3, m))
    b_aug[:n, :] = B
    return a_aug, b_aug


def heuristic_lqr_gain(q_pos, q_vel, r_thrust: float, r_rate: float) -> np.ndarray:
    """Closed-form per-axis gains of the reference's heuristic LQRController
    (controllers/__init__.py:522-574) — the controller RiccatiLQRController falls back to when
    the DARE cannot be solved (riccati_lqr.py:747-777)."""
    q_pos, q_vel = np.asarray(q_pos, dtype=float), np.asarray(q_vel, dtype=float)
    K = np.zeros((4, 6))
    K[0, 2] = np.sqrt(q_pos[2] / r_thrust)
    K[0, 5] = np.sqrt(2 * np.sqrt(q_pos[2] / r_thrust) + q_vel[2] / r_thrust)
    K[1, 1] = -np.sqrt(q_pos[1] / r_rate)
    K[1, 4] = -np.sqrt(2 * np.sqrt(q_pos[1] / r_rate) + q_vel[1] / r_rate)
    K[2, 0] = np.sqrt(q_pos[0] / r_rate)
    K[2, 3] = np.sqrt(2 * np.sqrt(q_pos[0] / r_rate) + q_vel[0] / r_rate)
    return K


def _is_symmetric(matrix: np.ndarray) -> bool:
    return bool(np.allclose(matrix, matrix.T, atol=1e-8))


def _is_positive_semidefinite(matrix: np.ndarray, name: str = "matrix") -> bool:
    """riccati_lqr.py:57-84"""
    if matrix.ndim != 2 or matrix.shape[0] != matrix.shape[1]:
        return False
    if not _is_symmetric(matrix):
        logger.warning("%s is not symmetric", name)
        return False
    eig = np.linalg.eigvalsh(matrix)
    if np.any(eig < -1e-10):
        logger.warning("%s has negative eigenvalues: %s", name, eig)
        return False
    return True


def _is_positive_definite(matrix: np.ndarray, name: str = "matrix") -> bool:
    """riccati_lqr.py:87-116"""
    if matrix.ndim != 2 or matrix.shape[0] != matrix.shape[1]:
        return False
    if not _is_symmetric(matrix):
        logger.warning("%s is not symmetric", name)
        return False
    eig = np.linalg.eigvalsh(matrix)
    if np.any(eig <= 1e-10):
        logger.warning("%s is not positive definite, eigenvalues: %s", name, eig)
        return False
    return True


# --------------------------------------------------------------------------------- DARE ------
_DARE_INFO = {1: "doubling iteration did not converge", 2: "singular pivot", 3: "invalid Q or R"}


def solve_dare_batched(A, B, Q, R, ctx: "_lib.Context | None" = None, want_P: bool = True):
    """Solve ``batch`` DAREs on the GPU (qt_dare_batched).

    A [n,n] or [batch,n,n]; B [n,m] or [batch,n,m]; Q [batch,n,n] or [n,n]; R [batch,m,m] or
    [m,m] — torch CUDA float64 tensors or anything ``torch.as_tensor`` accepts.  Returns
    ``(P [batch,n,n] or None, K [batch,m,n], iters int32 [batch], info int32 [batch])`` as CUDA
    tensors; info: 0 ok, 1 not converged, 2 singular, 3 Q/R rejected.
    """
    import torch

    ctx = ctx or _lib.get_context()
    dev = ctx.device

    def prep(x, inner):
        t = torch.as_tensor(x, dtype=torch.float64, device=dev)
        if t.dim() == 2:
            t = t.unsqueeze(0)
        if t.dim() != 3 or tuple(t.shape[1:]) != inner:
            raise ValueError(f"expected [...,{inner[0]},{inner[1]}], got {tuple(t.shape)}")
        return t.contiguous()

    a0 = torch.as_tensor(A, dtype=torch.float64, device=dev)
    b0 = torch.as_tensor(B, dtype=torch.float64, device=dev)
    n, m = int(a0.shape[-1]), int(b0.shape[-1])
    At, Bt, Qt, Rt = prep(a0, (n, n)), prep(b0, (n, m)), prep(Q, (n, n)), prep(R, (m, m))
    batch = max(t.shape[0] for t in (At, Bt, Qt, Rt))
    for t in (At, Bt, Qt, Rt):
        if t.shape[0] not in (1, batch):
            raise ValueError("batch sizes of A, B, Q, R must be equal or 1")

    def stride(t, size):
        return 0 if (t.shape[0] == 1 and batch > 1) else size

    P = torch.empty((batch, n, n), dtype=torch.float64, device=dev) if want_P else None
    K = torch.empty((batch, m, n), dtype=torch.float64, device=dev)
    iters = torch.zeros(batch, dtype=torch.int32, device=dev)
    info = torch.zeros(batch, dtype=torch.int32, device=dev)
    ctx.dare_batched(n, m, batch, At, stride(At, n * n), Bt, stride(Bt, n * m), Qt,
                     stride(Qt, n * n), Rt, stride(Rt, m * m), P, K, iters, info)
    return P, K, iters, info


def solve_dare(A: np.ndarray, B: np.ndarray, Q: np.ndarray, R: np.ndarray):
    """riccati_lqr.py:119-184: returns ``(P, K)`` as numpy arrays.

    Raises ValueError for incompatible shapes / Q not PSD / R not PD (checked on the host with the
    reference's own numpy tests) and RuntimeError if the device solver reports failure."""
    A, B, Q, R = (np.asarray(x, dtype=np.float64) for x in (A, B, Q, R))
    n = A.shape[0]
    m = B.shape[1]
    if A.shape != (n, n):
        raise ValueError(f"A must be square, got shape {A.shape}")
    if B.shape != (n, m):
        raise ValueError(f"B must have shape ({n}, m), got {B.shape}")
    if Q.shape != (n, n):
        raise ValueError(f"Q must have shape ({n}, {n}), got {Q.shape}")
    if R.shape != (m, m):
        raise ValueError(f"R must have shape ({m}, {m}), got {R.shape}")
    if not _is_positive_semidefinite(Q, "Q"):
        raise ValueError("Q matrix must be positive semi-definite")
    if not _is_positive_definite(R, "R"):
        raise ValueError("R matrix must be positive definite")
    P, K, _, info = solve_dare_batched(A, B, Q, R)
    code = int(info.cpu()[0])
    if code != 0:
        raise RuntimeError(f"DARE solver failed: {_DARE_INFO.get(code, code)}")
    return P[0].cpu().numpy(), K[0].cpu().numpy()


# --------------------------------------------------------------------------- controllers -----
@dataclass
class ActionLimits:
    """controllers/base.py:32-57"""

    min_thrust: float = 0.0
    max_thrust: float = 20.0
    max_rate: float = 3.0

    def clip_action(self, action: dict) -> dict:
        return {"thrust": np.clip(action["thrust"], self.min_thrust, self.max_thrust),
                "roll_rate": np.clip(action["roll_rate"], -self.max_rate, self.max_rate),
                "pitch_rate": np.clip(action["pitch_rate"], -self.max_rate, self.max_rate),
                "yaw_rate": np.clip(action["yaw_rate"], -self.max_rate, self.max_rate)}


DEFAULT_ACTION_LIMITS = ActionLimits()


def validate_action(action: dict) -> None:
    """controllers/base.py:64-81: KeyError for a missing key, TypeError for a non-numeric value."""
    for key in ACTION_KEYS:
        if key not in action:
            raise KeyError(f"Action missing required key: '{key}'")
        if not isinstance(action[key], (int, float)):
            raise TypeError(f"Action['{key}'] must be numeric, got {type(action[key]).__name__}")


class BaseController:
    """controllers/base.py:83-135"""

    def __init__(self, name: str = "base", config: dict | None = None, mass: float = 1.0,
                 gravity: float = 9.81):
        self.name = name
        self.config = config or {}
        self.mass = mass
        self.gravity = gravity

    def compute_action(self, observation: dict) -> dict:
        raise NotImplementedError("Subclasses must implement compute_action")

    def reset(self) -> None:
        pass


def _validate_observation(observation: dict) -> None:
    """riccati_lqr.py:319-342"""
    for key in ("quadcopter", "target"):
        if key not in observation:
            raise KeyError(f"Observation missing required key: '{key}'")
    for key in ("position", "velocity", "attitude", "angular_velocity"):
        if key not in observation["quadcopter"]:
            raise KeyError(f"Observation['quadcopter'] missing required key: '{key}'")
    for key in ("position", "velocity"):
        if key not in observation["target"]:
            raise KeyError(f"Observation['target'] missing required key: '{key}'")


def observation_to_vector(observation: dict) -> np.ndarray:
    """Flatten the reference's nested observation (quadcopter_env.py:472-496) into the 22-vector
    layout of QT_OBS_DIM: quad p, v, attitude, angular velocity; target p, v, a; time."""
    quad, target = observation["quadcopter"], observation["target"]
    acc = target.get("acceleration", None)
    return np.concatenate([
        np.asarray(quad["position"], dtype=np.float64),
        np.asarray(quad["velocity"], dtype=np.float64),
        np.asarray(quad["attitude"], dtype=np.float64),
        np.asarray(quad["angular_velocity"], dtype=np.float64),
        np.asarray(target["position"], dtype=np.float64),
        np.asarray(target["velocity"], dtype=np.float64),
        np.zeros(3) if acc is None else np.asarray(acc, dtype=np.float64),
        [float(observation.get("time", 0.0))],
    ])


class _GainController(BaseController):
    """Shared machinery of the controllers whose law is u = K e (+ K_i I): device-side state for
    the scalar ``compute_action`` interface and the parameter view the kernels need."""

    def _init_limits_and_ff(self, config: dict) -> None:
        self.max_thrust = config.get("max_thrust", 20.0)
        self.min_thrust = config.get("min_thrust", 0.0)
        self.max_rate = config.get("max_rate", 3.0)
        self.hover_thrust = self.mass * self.gravity
        self.feedforward_enabled = config.get("feedforward_enabled", False)
        self.ff_velocity_gain = self._ensure_array(config.get("ff_velocity_gain", [0.0, 0.0, 0.0]))
        self.ff_acceleration_gain = self._ensure_array(config.get("ff_acceleration_gain", [0.0, 0.0, 0.0]))
        self.ff_max_velocity = config.get("ff_max_velocity", 10.0)
        self.ff_max_acceleration = config.get("ff_max_acceleration", 5.0)
        self._saturation_count = 0
        self.last_control_components: dict | None = None
        self._dev = None  # lazily created device scratch for compute_action

    @staticmethod
    def _ensure_array(value, size: int = 3) -> np.ndarray:
        if np.isscalar(value):
            return np.array([value] * size)
        return np.array(value)

    # -- what the kernels need --------------------------------------------------------------
    def effective_gain(self) -> np.ndarray:
        """Row-major [4, 6] (LQR) or [4, 9] (LQI: K_pd | K_i) gain actually applied."""
        raise NotImplementedError

    def effective_use_lqi(self) -> bool:
        return False

    def k_cols(self) -> int:
        """Doubles per row of the 4-row gain record the kernels read (6 LQR, 9 LQI, 3 PID)."""
        return 9 if self.effective_use_lqi() else 6

    def carries_state(self) -> bool:
        """True if state inside the controller object influences its next action."""
        return self.effective_use_lqi()

    def kernel_params(self, env_config=None, criteria=None) -> QtParams:
        p = make_params(env_config, self, criteria)
        p.use_lqi = 1 if self.effective_use_lqi() else 0
        return p

    # controller state as the kernels see it: [integral(3), saturation count | PID last time]
    def state_vector(self) -> np.ndarray:
        integ = getattr(self, "integral_state", None)
        return np.array([*(integ if integ is not None else np.zeros(3)), float(self._saturation_count)])

    def load_state_vector(self, v: np.ndarray, absolute_count: bool) -> None:
        if getattr(self, "integral_state", None) is not None:
            self.integral_state = np.array(v[0:3], dtype=np.float64)
        self._saturation_count = int(v[3]) if absolute_count else self._saturation_count + int(v[3])

    # -- scalar interface --------------------------------------------------------------------
    def _device_state(self):
        import torch

        if self._dev is None:
            ctx = _lib.get_context()
            dev = ctx.device
            self._dev = {
                "ctx": ctx,
                "obs": torch.zeros((22, 1), dtype=torch.float64, device=dev),
                "integral": torch.zeros((4, 1), dtype=torch.float64, device=dev),
                "sat": torch.zeros(1, dtype=torch.int64, device=dev),
                "action": torch.zeros((4, 1), dtype=torch.float64, device=dev),
                "K": torch.as_tensor(np.ascontiguousarray(self.effective_gain()), dtype=torch.float64, device=dev),
                "params": self.kernel_params(),
            }
        return self._dev

    def _invalidate_device_state(self) -> None:
        self._dev = None

    def _launch_scalar(self, obs_vec: np.ndarray, integral: np.ndarray | None):
        import torch

        d = self._device_state()
        ctx = d["ctx"]
        ctx.bind_torch_stream()
        d["obs"].copy_(torch.from_numpy(obs_vec.reshape(22, 1)))
        stateful = integral is not None
        if stateful:
            n = len(integral)
            d["integral"][:n].copy_(torch.from_numpy(np.asarray(integral, dtype=np.float64).reshape(n, 1)))
        d["sat"].zero_()
        ctx.compute_action_batched(d["params"], 1, d["obs"], d["K"], self.k_cols(), 1,
                                   d["integral"] if stateful else None, d["sat"], d["action"])
        action = d["action"].cpu().numpy().reshape(4)
        new_integral = d["integral"].cpu().numpy().reshape(4)[: len(integral)] if stateful else None
        saturated = bool(int(d["sat"].cpu()[0]))
        return action, new_integral, saturated

    def compute_action_batched(self, obs, integral=None, saturation=None, k_stride: int | None = None, K=None):
        """Batched control law: ``obs`` is a CUDA float64 tensor [22, E]; returns actions [4, E].
        ``integral`` ([3, E]) and ``saturation`` (int64 [E]) are updated in place when given."""
        import torch

        ctx = _lib.get_context()
        E = int(obs.shape[1])
        stateful = self.carries_state()
        Kt = K if K is not None else torch.as_tensor(
            np.ascontiguousarray(self.effective_gain()), dtype=torch.float64, device=obs.device)
        action = torch.empty((4, E), dtype=torch.float64, device=obs.device)
        if stateful and integral is None:
            integral = torch.zeros((4, E), dtype=torch.float64, device=obs.device)
            integral[3] = float("nan")
        ctx.compute_action_batched(self.kernel_params(), E, obs.contiguous(), Kt, self.k_cols(),
                                   k_stride or E, integral if stateful else None, saturation, action)
        return action

    def get_control_components(self) -> dict | None:
        return self.last_control_components

    def _feedforward_diagnostics(self, obs_vec: np.ndarray, has_acceleration: bool):
        """Host restatement of the feedforward shaping for the DIAGNOSTICS dictionary only — the action itself comes
        from the kernel (qt_step.cuh ``ctrl_action`` / ``pid_action``) and nothing computed here feeds back into it.
        Returns (velocity-error with the scaled target velocity, clamped target velocity, FF acceleration term):
        the target velocity is clamped to ``ff_max_velocity`` and scaled by 1 + ff_velocity_gain, the target
        acceleration is clamped to ``ff_max_acceleration`` and scaled by ff_acceleration_gain
        (controllers/__init__.py:300-336, riccati_lqr.py:832-862)."""
        tv = obs_vec[15:18].copy()
        ff_acc = np.zeros(3)
        clamped = tv.copy()
        if self.feedforward_enabled:
            mag = float(np.linalg.norm(tv))
            if mag > self.ff_max_velocity and mag > 0.0:
                clamped = tv * (self.ff_max_velocity / mag)
            tv = (1.0 + self.ff_velocity_gain) * clamped
            if has_acceleration:
                acc = obs_vec[18:21].copy()
                amag = float(np.linalg.norm(acc))
                if amag > self.ff_max_acceleration and amag > 0.0:
                    acc = acc / amag * self.ff_max_acceleration
                ff_acc = self.ff_acceleration_gain * acc
        return tv - obs_vec[3:6], clamped, ff_acc

    def get_saturation_count(self) -> int:
        return self._saturation_count


class LQRController(_GainController):
    """Heuristic / arbitrary-gain LQR (controllers/__init__.py:398-700): ``u = K e`` with either a
    user-supplied 4x6 ``K`` or the closed-form per-axis gains.  Shares the Riccati controller's
    kernels; it exists here because it is RiccatiLQRController's fallback (riccati_lqr.py:747-777).
    """

    def __init__(self, config: dict | None = None):
        config = config or {}
        super().__init__(name="lqr", config=config, mass=config.get("mass", 1.0),
                         gravity=config.get("gravity", 9.81))
        self.dt = config.get("dt", 0.01)
        self.use_lqi = False
        self.integral_limit = 10.0
        self.integral_zero_threshold = 0.01
        self._init_limits_and_ff(config)
        if "K" in config and config["K"] is not None:
            self.K = np.array(config["K"], dtype=np.float64)
            if self.K.shape != (4, 6):
                raise ValueError(f"K matrix must have shape (4, 6), got {self.K.shape}")
        else:
            self.K = heuristic_lqr_gain(
                config.get("q_pos", [0.0001, 0.0001, 16.0]), config.get("q_vel", [0.0036, 0.0036, 4.0]),
                config.get("r_thrust", 1.0), config.get("r_rate", 1.0))

    def effective_gain(self) -> np.ndarray:
        return self.K

    def compute_action(self, observation: dict) -> dict:
        _validate_observation(observation)
        obs_vec = observation_to_vector(observation)
        action, _, saturated = self._launch_scalar(obs_vec, None)
        self._saturation_count += int(saturated)
        # diagnostics only (controllers/__init__.py:655-660)
        vel_error, clamped_tv, ff_acc = self._feedforward_diagnostics(
            obs_vec, observation["target"].get("acceleration", None) is not None)
        ff_on = bool(self.feedforward_enabled)
        self.last_control_components = {
            "feedback_u": self.K @ np.concatenate([obs_vec[12:15] - obs_vec[0:3], vel_error]),
            "ff_velocity_term": self.ff_velocity_gain * clamped_tv if ff_on else np.zeros(3),
            "ff_acceleration_term": ff_acc,
            "is_saturated": saturated,
        }
        return dict(zip(ACTION_KEYS, (float(a) for a in action)))

    def reset(self) -> None:
        self.last_control_components = None


class PIDController(_GainController):
    """PID position controller (controllers/__init__.py:116-395): per-axis P/I/D on the position
    error mapped to thrust / pitch rate / -roll rate, optional feedforward.  The gain record the
    kernels read is ``[kp | ki | kd | 0]`` (4 x 3).  State: ``integral_error`` and the time of the
    previous call (the reference derives dt from observation["time"])."""

    def __init__(self, config: dict | None = None):
        config = config or {}
        super().__init__(name="pid", config=config, mass=config.get("mass", 1.0),
                         gravity=config.get("gravity", 9.81))
        self.dt = config.get("dt", 0.01)
        self.use_lqi = False
        self.integral_zero_threshold = 0.01
        self._init_limits_and_ff(config)
        self.kp_pos = self._vec(config.get("kp_pos", config.get("kp", [0.01, 0.01, 4.0])))
        self.ki_pos = self._vec(config.get("ki_pos", config.get("ki", [0.0, 0.0, 0.0])))
        self.kd_pos = self._vec(config.get("kd_pos", config.get("kd", [0.06, 0.06, 2.0])))
        self.integral_limit = config.get("integral_limit", 0.0)
        self.integral_error = np.zeros(3)
        self._last_time: float | None = None

    @staticmethod
    def _vec(value) -> np.ndarray:
        return np.array(value, dtype=np.float64) if hasattr(value, "__len__") else np.array([value] * 3, dtype=np.float64)

    def effective_gain(self) -> np.ndarray:
        return np.stack([self.kp_pos, self.ki_pos, self.kd_pos, np.zeros(3)])

    def k_cols(self) -> int:
        return 3

    def carries_state(self) -> bool:
        # integral_limit == 0 (the default) clips the integral to zero after every update, and a
        # stale _last_time only makes the first dt of the next episode negative instead of zero -
        # both skip the update - so such a controller behaves identically with or without reset().
        return self.integral_limit != 0

    def state_vector(self) -> np.ndarray:
        return np.array([*self.integral_error, float("nan") if self._last_time is None else float(self._last_time)])

    def load_state_vector(self, v: np.ndarray, absolute_count: bool) -> None:
        self.integral_error = np.array(v[0:3], dtype=np.float64)
        self._last_time = None if np.isnan(v[3]) else float(v[3])

    def compute_action(self, observation: dict) -> dict:
        _validate_observation(observation)
        obs_vec = observation_to_vector(observation)
        action, state, _ = self._launch_scalar(obs_vec, self.state_vector())
        self.load_state_vector(state, True)
        # diagnostics only (controllers/__init__.py:296-354): the terms of the correction the kernel applied
        vel_error, clamped_tv, ff_acc = self._feedforward_diagnostics(
            obs_vec, observation["target"].get("acceleration", None) is not None)
        p_term = self.kp_pos * (obs_vec[12:15] - obs_vec[0:3])
        i_term = self.ki_pos * self.integral_error
        d_term = self.kd_pos * vel_error
        ff_on = bool(self.feedforward_enabled)
        self.last_control_components = {
            "p_term": p_term, "i_term": i_term, "d_term": d_term,
            # the share of the D term that the scaled target velocity contributes
            "ff_velocity_term": self.kd_pos * self.ff_velocity_gain * clamped_tv if ff_on else np.zeros(3),
            "ff_acceleration_term": ff_acc,
            "total_correction": p_term + i_term + d_term + ff_acc,
            "integral_error": self.integral_error.copy(),
        }
        return dict(zip(ACTION_KEYS, (float(a) for a in action)))

    def reset(self) -> None:
        self.integral_error = np.zeros(3)
        self._last_time = None
        self.last_control_components = None


class RiccatiLQRController(_GainController):
    """riccati_lqr.py:345-1086.  See the module docstring for what runs where."""

    def __init__(self, config: dict | None = None):
        config = config or {}
        mass = config.get("mass", 1.0)
        gravity = config.get("gravity", 9.81)
        self.dt = config.get("dt", 0.01)
        super().__init__(name="riccati_lqr", config=config, mass=mass, gravity=gravity)
        self._init_limits_and_ff(config)

        self.use_lqi = config.get("use_lqi", False)
        self.q_int = self._validate_q_int(config.get("q_int", [0.0, 0.0, 0.0]))
        self.integral_limit = config.get("integral_limit", 10.0)
        self.integral_zero_threshold = config.get("integral_zero_threshold", 0.01)
        self.integral_state: np.ndarray | None = np.zeros(3) if self.use_lqi else None
        self.K_pd: np.ndarray | None = None
        self.K_i: np.ndarray | None = None

        self.A, self.B = build_linearized_system(dt=self.dt, mass=self.mass, gravity=self.gravity)
        self.A_aug = self.B_aug = None
        if self.use_lqi:
            self.A_aug, self.B_aug = build_augmented_lqi_system(self.A, self.B, self.dt)
        self.Q = self._build_Q_matrix(config)
        self.R = self._build_R_matrix(config)

        self.fallback_on_failure = config.get("fallback_on_failure", True)
        self.fallback_controller: LQRController | None = None
        self._using_fallback = False
        self.P = None
        self.K = None
        self._solve_riccati()

    # -- configuration (riccati_lqr.py:553-700) -------------------------------------------------
    def _validate_q_int(self, q_int) -> np.ndarray:
        arr = np.array([q_int, q_int, q_int]) if np.isscalar(q_int) else np.array(q_int)
        if len(arr) != 3:
            raise ValueError(
                f"q_int must have exactly 3 elements (one per position axis x, y, z), "
                f"got {len(arr)} elements. Provide a 3-element list like [0.01, 0.01, 0.1] or a scalar.")
        if np.any(arr < 0):
            bad = ", ".join(f"{'xyz'[i]}={arr[i]:.4f}" for i in np.where(arr < 0)[0])
            raise ValueError(
                f"q_int weights must be non-negative, but got negative values for: {bad}. "
                f"Use positive weights or zero to disable integral action on an axis.")
        return arr

    def _build_Q_matrix(self, config: dict) -> np.ndarray:
        if config.get("Q") is not None:
            Q = np.array(config["Q"])
            if self.use_lqi:
                if Q.shape == (9, 9):
                    return Q
                if Q.shape == (6, 6):
                    Q_aug = np.zeros((9, 9))
                    Q_aug[:6, :6] = Q
                    Q_aug[6:, 6:] = np.diag(self.q_int)
                    return Q_aug
                raise ValueError(f"Q matrix for LQI must have shape (9, 9) or (6, 6), got {Q.shape}")
            if Q.shape != (6, 6):
                raise ValueError(f"Q matrix must have shape (6, 6), got {Q.shape}")
            return Q
        q_pos = np.array(config.get("q_pos", [0.0001, 0.0001, 16.0]))
        q_vel = np.array(config.get("q_vel", [0.0036, 0.0036, 4.0]))
        if len(q_pos) != 3:
            raise ValueError(f"q_pos must have 3 elements, got {len(q_pos)}")
        if len(q_vel) != 3:
            raise ValueError(f"q_vel must have 3 elements, got {len(q_vel)}")
        parts = [q_pos, q_vel] + ([self.q_int] if self.use_lqi else [])
        return np.diag(np.concatenate(parts))

    def _build_R_matrix(self, config: dict) -> np.ndarray:
        if config.get("R") is not None:
            R = np.array(config["R"])
            if R.shape != (4, 4):
                raise ValueError(f"R matrix must have shape (4, 4), got {R.shape}")
            return R
        r_controls = np.array(config.get("r_controls", [1.0, 1.0, 1.0, 1.0]))
        if len(r_controls) != 4:
            raise ValueError(f"r_controls must have 4 elements, got {len(r_controls)}")
        return np.diag(r_controls)

    # -- gains (riccati_lqr.py:702-777) -----------------------------------------------------------
    def _solve_riccati(self) -> None:
        try:
            if self.use_lqi:
                self.P, self.K = solve_dare(self.A_aug, self.B_aug, self.Q, self.R)
                self.K_pd, self.K_i = self.K[:, :6], self.K[:, 6:]
            else:
                self.P, self.K = solve_dare(self.A, self.B, self.Q, self.R)
                self.K_pd, self.K_i = self.K, np.zeros((self.K.shape[0], 3))
            self._using_fallback = False
        except _lib.QtrackError:
            raise  # a missing library / GPU is never papered over by the heuristic fallback
        except (ValueError, RuntimeError) as e:
            logger.warning("DARE solver failed: %s", e)
            if not self.fallback_on_failure:
                raise
            logger.warning("Falling back to heuristic LQR controller")
            rate_costs = [self.R[1, 1], self.R[2, 2], self.R[3, 3]]
            self.fallback_controller = LQRController(config={
                "mass": self.mass, "gravity": self.gravity,
                "q_pos": [self.Q[0, 0], self.Q[1, 1], self.Q[2, 2]],
                "q_vel": [self.Q[3, 3], self.Q[4, 4], self.Q[5, 5]],
                "r_thrust": self.R[0, 0], "r_rate": sum(rate_costs) / len(rate_costs),
                "max_thrust": self.max_thrust, "min_thrust": self.min_thrust, "max_rate": self.max_rate,
            })
            self._using_fallback = True
        self._invalidate_device_state()

    def set_gain(self, K: np.ndarray, P: np.ndarray | None = None) -> None:
        """Install a gain computed elsewhere (e.g. one row of a batched DARE sweep)."""
        K = np.asarray(K, dtype=np.float64)
        want = (4, 9) if self.use_lqi else (4, 6)
        if K.shape != want:
            raise ValueError(f"K must have shape {want}, got {K.shape}")
        self.K, self.P = K, P
        if self.use_lqi:
            self.K_pd, self.K_i = K[:, :6], K[:, 6:]
        else:
            self.K_pd, self.K_i = K, np.zeros((4, 3))
        self._using_fallback = False
        self._invalidate_device_state()

    def effective_gain(self) -> np.ndarray:
        if self._using_fallback and self.fallback_controller is not None:
            return self.fallback_controller.K
        return np.ascontiguousarray(self.K)

    def effective_use_lqi(self) -> bool:
        return bool(self.use_lqi) and not self._using_fallback

    def kernel_params(self, env_config=None, criteria=None) -> QtParams:
        if self._using_fallback and self.fallback_controller is not None:
            # the fallback LQRController is built without feedforward (riccati_lqr.py:764-776)
            return self.fallback_controller.kernel_params(env_config, criteria)
        return super().kernel_params(env_config, criteria)

    # -- scalar interface (riccati_lqr.py:779-967) -------------------------------------------------
    def compute_action(self, observation: dict) -> dict:
        if self._using_fallback and self.fallback_controller is not None:
            return self.fallback_controller.compute_action(observation)
        _validate_observation(observation)
        obs_vec = observation_to_vector(observation)
        if "acceleration" not in observation["target"]:
            obs_vec[18:21] = 0.0  # acceleration feedforward only when the key exists (:854-861)
        lqi = self.use_lqi and self.integral_state is not None
        action, new_integral, saturated = self._launch_scalar(obs_vec, self.integral_state if lqi else None)
        if lqi:
            self.integral_state = new_integral
        self._saturation_count += int(saturated)
        # diagnostics only (never feeds back into the action): riccati_lqr.py:946-960
        vel_error, clamped_tv, ff_acc = self._feedforward_diagnostics(obs_vec, "acceleration" in observation["target"])
        state_error = np.concatenate([obs_vec[12:15] - obs_vec[0:3], vel_error])
        if lqi:
            feedback_u = self.K_pd @ state_error + self.K_i @ self.integral_state
        else:
            feedback_u = self.K @ state_error
        self.last_control_components = {
            "state_error": state_error,
            "feedback_u": feedback_u,
            "ff_velocity_term": self.ff_velocity_gain * clamped_tv if self.feedforward_enabled else np.zeros(3),
            "ff_acceleration_term": ff_acc,
            "K_matrix": self.K.copy(),
            "is_saturated": saturated,
        }
        if self.use_lqi:
            self.last_control_components["integral_state"] = self.integral_state.copy()
            self.last_control_components["K_pd"] = self.K_pd.copy()
            self.last_control_components["K_i"] = self.K_i.copy()
        return dict(zip(ACTION_KEYS, (float(a) for a in action)))

    # -- getters (riccati_lqr.py:969-1086) -----------------------------------------------------------
    def is_using_fallback(self) -> bool:
        return self._using_fallback

    def get_gain_matrix(self):
        return self.K if not self._using_fallback else None

    def get_riccati_solution(self):
        return self.P if not self._using_fallback else None

    def get_integral_gains(self):
        return self.K_i if self.use_lqi and not self._using_fallback else None

    def get_pd_gains(self):
        return self.K_pd if self.use_lqi and not self._using_fallback else None

    def get_integral_state(self):
        if self.use_lqi and self.integral_state is not None:
            return self.integral_state.copy()
        return None

    def reset_integral_state(self) -> None:
        if self.use_lqi and self.integral_state is not None:
            self.integral_state = np.zeros(3)

    def is_lqi_mode(self) -> bool:
        return self.use_lqi

    def reset(self) -> None:
        self.last_control_components = None
        self._saturation_count = 0
        if self.use_lqi and self.integral_state is not None:
            self.integral_state = np.zeros(3)
        if self.fallback_controller is not None:
            self.fallback_controller.reset()
