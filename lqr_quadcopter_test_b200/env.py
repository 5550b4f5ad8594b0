"""Quadcopter tracking environments on the GPU.

``BatchedQuadcopterEnv`` steps E independent episodes at once (struct-of-arrays CUDA tensors in,
CUDA tensors out).  ``QuadcopterEnv`` is the drop-in for the reference's scalar environment
(``quadcopter_tracking.env.quadcopter_env.QuadcopterEnv``, cited as quadcopter_env.py:line): same
constructor, ``reset(seed) -> obs``, ``step(action) -> (obs, reward, done, info)``, same nested
observation / info dictionaries, same exceptions — it is a batch of one over the same kernels.

Every number an observation, reward, done flag or info entry contains is computed by
``qt_reset_batched`` / ``qt_observe_batched`` / ``qt_step_batched`` (csrc/qt_env.cu); the Python
here moves tensors and builds dictionaries.
"""
from __future__ import annotations

import logging
import os

import numpy as np

from . import _lib
from . import params as P
from .config import EnvConfig
from .target_motion import TargetMotion

logger = logging.getLogger(__name__)


def _as_env_config(config) -> EnvConfig:
    if config is None:
        return EnvConfig()
    if isinstance(config, dict):
        return EnvConfig.from_dict(config)
    return config


class BatchedQuadcopterEnv:
    """E episodes of the reference environment, advanced in lock-step on one GPU.

    State lives in CUDA tensors (float64, episode index fastest): ``state`` [12, E], ``time`` [E],
    ``tparams`` [12, E] (target pattern parameters), ``counters`` int64 [3, E] (steps, on-target
    steps, steps with repaired actions).
    """

    STATE_DIM = 12
    ACTION_DIM = 4

    def __init__(self, config: dict | EnvConfig | None = None, num_envs: int = 1, device: int | None = None):
        import torch

        self.config = _as_env_config(config)
        self.num_envs = int(num_envs)
        self.ctx = _lib.get_context(device)
        self.device = self.ctx.device
        self.params = P.make_params(self.config)
        E = self.num_envs
        f64 = dict(dtype=torch.float64, device=self.device)
        self.state = torch.zeros((12, E), **f64)
        self.time = torch.zeros(E, **f64)
        self.tparams = torch.zeros((12, E), **f64)
        self.counters = torch.zeros((3, E), dtype=torch.int64, device=self.device)
        self.seeds = torch.zeros(E, dtype=torch.int64, device=self.device)
        self.draws = torch.zeros((2, E), dtype=torch.int64, device=self.device)
        self._obs = torch.zeros((22, E), **f64)
        self._reward = torch.zeros(E, **f64)
        self._done = torch.zeros(E, dtype=torch.int32, device=self.device)
        self._info = torch.zeros((4, E), **f64)
        self._action = torch.zeros((4, E), **f64)
        self._initialized = False

    # -- reset -----------------------------------------------------------------------------------
    def reset(self, seeds=None):
        """Reset all episodes.  ``seeds``: int64 tensor / array / list of E seeds, a single int
        (episode i gets seed + i, the convention of Evaluator.evaluate, eval.py:198-199), or None to
        keep drawing from the existing generators (quadcopter_env.py:122-130).  Returns the
        observation tensor [22, E]."""
        import torch

        E = self.num_envs
        self.ctx.bind_torch_stream()
        if seeds is not None:
            if isinstance(seeds, (int, np.integer)):
                s = torch.arange(E, dtype=torch.int64, device=self.device) + int(seeds)
            else:
                s = torch.as_tensor(seeds, dtype=torch.int64, device=self.device).reshape(-1)
                if s.numel() != E:
                    raise ValueError(f"expected {E} seeds, got {s.numel()}")
            self.seeds = s.contiguous()
            self.draws.zero_()
        self.ctx.reset_batched(self.params, self.seeds, self.state, self.tparams, self.draws)
        self.time.zero_()
        self.counters.zero_()
        self._initialized = True
        return self.observe()

    def observe(self):
        self.ctx.bind_torch_stream()
        self.ctx.observe_batched(self.params, self.num_envs, self.state, self.time, self.tparams, self._obs)
        return self._obs

    # -- step ------------------------------------------------------------------------------------
    def step(self, actions=None, controller=None, K=None, integral=None, saturation=None):
        """One environment step for every episode (quadcopter_env.py:152-232).

        ``actions``: CUDA float64 tensor [4, E] (thrust, roll, pitch, yaw rate).  Alternatively pass
        ``controller`` (a RiccatiLQRController / LQRController) and the control law is evaluated
        inside the same kernel from the current state.  Returns ``(obs [22,E], reward [E],
        done int32 [E] (0 = running, else enum qt_termination), info)`` where info holds the tensors
        tracking_error, on_target, on_target_ratio, success, action (after the env's validation)."""
        import torch

        if not self._initialized:
            raise RuntimeError("Environment not initialized. Call reset() first.")
        self.ctx.bind_torch_stream()
        E = self.num_envs
        params, Kt, k_cols = self.params, None, 6
        if actions is not None:
            a = torch.as_tensor(actions, dtype=torch.float64, device=self.device)
            if tuple(a.shape) != (4, E):
                raise ValueError(f"Action tensor must have shape (4, {E}), got {tuple(a.shape)}")
            a = a.contiguous()
        else:
            if controller is None:
                raise ValueError("step() needs either actions or a controller")
            a = None
            params = controller.kernel_params(self.config)
            k_cols = controller.k_cols()
            Kt = K if K is not None else torch.as_tensor(
                np.ascontiguousarray(controller.effective_gain()), dtype=torch.float64, device=self.device)
            if controller.carries_state() and integral is None:
                raise ValueError("a stateful controller inside step() needs its state tensor "
                                 "([3, E] LQI integrator, [4, E] PID integral + previous time)")
        self.ctx.step_batched(params, E, self.state, self.time, self.tparams, self.counters, a, Kt,
                              k_cols, E if K is None else 1, integral, saturation, self._obs, self._reward,
                              self._done, self._info, self._action)
        info = {
            "tracking_error": self._info[0],
            "on_target": self._info[1],
            "on_target_ratio": self._info[2],
            "success": self._info[3],
            "action": self._action,
            "step": self.counters[0],
            "time": self.time,
        }
        return self._obs, self._reward, self._done, info

    @property
    def is_initialized(self) -> bool:
        return self._initialized


class QuadcopterEnv:
    """Scalar drop-in for the reference environment (quadcopter_env.py:45-667)."""

    POS_X, POS_Y, POS_Z = 0, 1, 2
    VEL_X, VEL_Y, VEL_Z = 3, 4, 5
    ROLL, PITCH, YAW = 6, 7, 8
    ROLL_RATE, PITCH_RATE, YAW_RATE = 9, 10, 11
    STATE_DIM = 12
    ACTION_DIM = 4

    def __init__(self, config: dict | EnvConfig | None = None):
        self.config = _as_env_config(config)
        self._batch = BatchedQuadcopterEnv(self.config, num_envs=1)
        seed = self.config.seed
        if seed is None:  # default_rng(None): fresh OS entropy
            seed = int.from_bytes(os.urandom(7), "little")
        self._batch.seeds.fill_(int(seed))
        # quadcopter_env.py:89: the env's target generator is a public attribute; here a host view of the pattern the
        # kernel drew (reset() below hands it the device's parameters)
        self.target = TargetMotion(self.config.target, seed=int(seed))
        self._time = 0.0
        self._step_count = 0
        self._initialized = False
        self._history: list[dict] = []
        self._action_violations: list[dict] = []
        self._obs_vec = np.zeros(22)

    # -- helpers ---------------------------------------------------------------------------------
    def _refresh_params(self) -> None:
        # the reference reads self.config on every call, so edits made after construction count
        self._batch.params = P.make_params(self.config)

    @staticmethod
    def _obs_dict(v: np.ndarray) -> dict:
        return {
            "quadcopter": {
                "position": v[0:3].copy(),
                "velocity": v[3:6].copy(),
                "attitude": v[6:9].copy(),
                "angular_velocity": v[9:12].copy(),
            },
            "target": {
                "position": v[12:15].copy(),
                "velocity": v[15:18].copy(),
                "acceleration": v[18:21].copy(),
            },
            "time": float(v[21]),
        }

    def _get_observation(self) -> dict:
        return self._obs_dict(self._obs_vec)

    # -- reference interface -----------------------------------------------------------------------
    def reset(self, seed: int | None = None) -> dict:
        import torch

        self._refresh_params()
        b = self._batch
        obs = b.reset(None if seed is None else torch.tensor([int(seed)], dtype=torch.int64))
        self._obs_vec = obs.cpu().numpy().reshape(22)
        self.target = TargetMotion(self.config.target, seed=int(b.seeds[0].item()))
        self.target._tparams = b.tparams[:, 0].cpu().numpy().copy()
        self._time = 0.0
        self._step_count = 0
        self._history = []
        self._action_violations = []
        self._initialized = True
        return self._get_observation()

    def _parse_action(self, action) -> np.ndarray:
        if isinstance(action, dict):
            return np.array([action.get("thrust", 0.0), action.get("roll_rate", 0.0),
                             action.get("pitch_rate", 0.0), action.get("yaw_rate", 0.0)], dtype=np.float64)
        vec = np.asarray(action, dtype=np.float64)
        if vec.shape != (4,):
            raise ValueError(f"Action array must have shape (4,), got {vec.shape}")
        return vec

    def _violation_messages(self, vec: np.ndarray) -> list[str]:
        """Text of quadcopter_env.py:265-291 for an action the kernel reported as repaired."""
        q = self.config.quadcopter
        msgs = []
        if not np.all(np.isfinite(vec)):
            msgs.append("Action contains NaN or Inf values")
            logger.warning("Action contains NaN or Inf, replacing with zeros")
            vec = np.nan_to_num(vec, nan=0.0, posinf=0.0, neginf=0.0)
        if vec[0] < q.min_thrust:
            msgs.append(f"Thrust {vec[0]:.2f} below min {q.min_thrust:.2f}")
        elif vec[0] > q.max_thrust:
            msgs.append(f"Thrust {vec[0]:.2f} above max {q.max_thrust:.2f}")
        for i, name in enumerate(("roll_rate", "pitch_rate", "yaw_rate"), start=1):
            if abs(vec[i]) > q.max_angular_rate:
                msgs.append(f"{name} {vec[i]:.2f} exceeds limit {q.max_angular_rate:.2f}")
        return msgs

    def step(self, action) -> tuple[dict, float, bool, dict]:
        import torch

        if not self._initialized:
            raise RuntimeError("Environment not initialized. Call reset() first.")
        vec = self._parse_action(action)
        self._refresh_params()
        b = self._batch
        violations_before = int(b.counters[2, 0])
        act = torch.from_numpy(vec.reshape(4, 1)).to(b.device)
        obs, reward, done, info_t = b.step(actions=act)
        # one device->host transfer for everything the dictionaries need
        packed = torch.cat([obs.reshape(-1), reward.reshape(-1), b._info.reshape(-1), b._action.reshape(-1),
                            done.to(torch.float64), b.counters[:, 0].to(torch.float64)]).cpu().numpy()
        self._obs_vec = packed[0:22].copy()
        reward_f = float(packed[22])
        info_v = packed[23:27]
        action_vec = packed[27:31]
        term = int(packed[31])
        steps, _on_count, violations_after = (int(x) for x in packed[32:35])
        if violations_after > violations_before:
            self._action_violations.append({
                "step": self._step_count, "time": self._time, "violations": self._violation_messages(vec)})
        self._time = float(packed[21])
        self._step_count = steps
        observation = self._get_observation()
        info = {
            "time": self._time,
            "step": self._step_count,
            "tracking_error": float(info_v[0]),
            "on_target": bool(info_v[1] != 0.0),
            "on_target_ratio": float(info_v[2]),
            "action_violations": len(self._action_violations),
        }
        done_b = term != 0
        if done_b:
            info["termination_reason"] = P.TERMINATION_REASONS[term]
            info["episode_length"] = self._time
            info["success"] = bool(info_v[3] != 0.0)
            if term == 3:
                logger.warning("Numerical instability detected in state")
        if self.config.logging.enabled and self._step_count % self.config.logging.log_interval == 0:
            # quadcopter_env.py:555-584
            self._history.append({
                "time": self._time,
                "step": self._step_count,
                "quadcopter_position": observation["quadcopter"]["position"].tolist(),
                "quadcopter_velocity": observation["quadcopter"]["velocity"].tolist(),
                "quadcopter_attitude": observation["quadcopter"]["attitude"].tolist(),
                "target_position": observation["target"]["position"].tolist(),
                "target_velocity": observation["target"]["velocity"].tolist(),
                "action": action_vec.tolist(),
                "reward": reward_f,
                "tracking_error": info["tracking_error"],
                "on_target": info["on_target"],
            })
        return observation, reward_f, done_b, info

    def get_history(self) -> list[dict]:
        return self._history.copy()

    def get_action_violations(self) -> list[dict]:
        return self._action_violations.copy()

    def render(self, mode: str = "dict"):
        return self._get_observation() if mode == "dict" else None

    @property
    def state(self) -> dict:
        return self._get_observation()

    @property
    def time(self) -> float:
        return self._time

    @property
    def dt(self) -> float:
        return self.config.simulation.dt

    @property
    def is_initialized(self) -> bool:
        return self._initialized

    def get_state_vector(self) -> np.ndarray:
        return self._batch.state.cpu().numpy().reshape(12).copy()

    def set_state_vector(self, state: np.ndarray) -> None:
        import torch

        state = np.asarray(state)
        if state.shape != (self.STATE_DIM,):
            raise ValueError(f"State must have shape ({self.STATE_DIM},), got {state.shape}")
        b = self._batch
        b.state.copy_(torch.from_numpy(state.astype(np.float64).reshape(12, 1)))
        self._refresh_params()
        self._obs_vec = b.observe().cpu().numpy().reshape(22)

    @staticmethod
    def hover_action(mass: float = 1.0, gravity: float = 9.81) -> dict:
        return {"thrust": mass * gravity, "roll_rate": 0.0, "pitch_rate": 0.0, "yaw_rate": 0.0}
