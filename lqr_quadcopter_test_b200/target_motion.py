"""Target motion generators — host-side mirror of ``quadcopter_tracking.env.target_motion`` (cited as
target_motion.py:line) over the device code of the rollout.

The reference evaluates five analytic patterns on the host with numpy (target_motion.py:29-248) and wraps them in
``TargetMotion`` (``reset`` draws the random pattern parameters from ``np.random.default_rng(seed)``, ``get_state``
adds the acceleration clamp, target_motion.py:251-429).  Here every number comes from the kernels the fused rollout
itself uses: pattern parameters from ``qt_reset_batched`` (SeedSequence -> PCG64 -> uniform / ziggurat on the device,
bit-exact with numpy), positions / velocities / accelerations from ``qt_observe_batched`` (``TargetGen`` in
csrc/qt_step.cuh).  The classes exist so that code written against the reference's target API — its own test-suite
first of all — runs against the drop-in; they are batches of one and cost a kernel launch per call.
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib
from . import params as P
from .config import EnvConfig, TargetParams

_INF = float("inf")


def _observe_target(params, tparams: np.ndarray, t: float):
    """(position, velocity, acceleration) of the pattern ``tparams`` at time ``t`` through qt_observe_batched."""
    import torch

    ctx = _lib.get_context()
    ctx.bind_torch_stream()
    f64 = dict(dtype=torch.float64, device=ctx.device)
    tp = torch.as_tensor(np.ascontiguousarray(tparams, dtype=np.float64).reshape(12, 1), **f64)
    state = torch.zeros((12, 1), **f64)
    time = torch.full((1,), float(t), **f64)
    obs = torch.empty((P.OBS_DIM, 1), **f64)
    ctx.observe_batched(params, 1, state, time, tp, obs)
    v = obs.cpu().numpy().reshape(-1)
    return v[12:15].copy(), v[15:18].copy(), v[18:21].copy()


class _Pattern:
    """One analytic pattern = a motion code + its 12 pattern parameters (layout: enum qt_motion, include/qtrack.h).
    ``get_state`` returns the raw pattern values: the acceleration clamp belongs to TargetMotion."""

    motion = "stationary"

    def __init__(self):
        self._tparams = np.zeros(12)

    def _params(self):
        env = EnvConfig()
        env.target.motion_type = self.motion
        env.target.max_acceleration = _INF
        return P.make_params(env)

    def get_state(self, t: float):
        return _observe_target(self._params(), self._tparams, t)


class LinearMotion(_Pattern):
    """target_motion.py:29-56"""

    motion = "linear"

    def __init__(self, start: np.ndarray, direction: np.ndarray, speed: float):
        super().__init__()
        self.start = np.array(start, dtype=np.float64)
        self.direction = np.asarray(direction, dtype=np.float64) / np.linalg.norm(direction)
        self.speed = speed
        self.velocity = self.direction * self.speed
        self._tparams[0:3], self._tparams[3:6] = self.start, self.velocity


class CircularMotion(_Pattern):
    """target_motion.py:59-115"""

    motion = "circular"

    def __init__(self, center: np.ndarray, radius: float, speed: float, initial_angle: float = 0.0):
        super().__init__()
        self.center = np.array(center, dtype=np.float64)
        self.radius, self.speed = radius, speed
        self.omega = speed / radius
        self.initial_angle = initial_angle
        self._tparams[0:3] = self.center
        self._tparams[3:6] = radius, self.omega, initial_angle


class SinusoidalMotion(_Pattern):
    """target_motion.py:118-150"""

    motion = "sinusoidal"

    def __init__(self, center: np.ndarray, amplitude: np.ndarray, frequency: np.ndarray, phase: np.ndarray = None):
        super().__init__()
        self.center = np.array(center, dtype=np.float64)
        self.amplitude = np.array(amplitude, dtype=np.float64)
        self.omega = 2 * np.pi * np.asarray(frequency, dtype=np.float64)
        self.phase = np.asarray(phase, dtype=np.float64) if phase is not None else np.zeros(3)
        self._tparams[0:3], self._tparams[3:6] = self.center, self.amplitude
        self._tparams[6:9], self._tparams[9:12] = self.omega, self.phase


class Figure8Motion(_Pattern):
    """target_motion.py:153-231"""

    motion = "figure8"

    def __init__(self, center: np.ndarray, scale: float, speed: float):
        super().__init__()
        self.center = np.array(center, dtype=np.float64)
        self.scale = scale
        self.omega = speed / scale
        self._tparams[0:3] = self.center
        self._tparams[3:5] = scale, self.omega


class StationaryMotion(_Pattern):
    """target_motion.py:234-248"""

    motion = "stationary"

    def __init__(self, position: np.ndarray):
        super().__init__()
        self.position = np.array(position, dtype=np.float64)
        self._tparams[0:3] = self.position


class TargetMotion:
    """target_motion.py:251-429: ``reset`` draws the pattern, ``get_state`` clamps the acceleration."""

    VALID_MOTION_TYPES = {"linear", "circular", "sinusoidal", "figure8", "stationary"}

    def __init__(self, params: TargetParams | None = None, seed: int | None = None):
        self.params = params or TargetParams()
        self._seed = self._fresh_seed() if seed is None else int(seed)
        self._draws = 0          # raw 64-bit outputs the stream has produced so far (reset() without a seed goes on)
        self._tparams: np.ndarray | None = None
        self._time = 0.0

    @staticmethod
    def _fresh_seed() -> int:
        return int.from_bytes(os.urandom(7), "little")   # default_rng(None): fresh OS entropy

    def _kernel_params(self, clamp: bool = True):
        env = EnvConfig()
        env.target = self.params
        prm = P.make_params(env)          # raises ValueError("Invalid motion type: ...") like target_motion.py:310-314
        if not clamp:
            prm.target_max_acceleration = _INF
        return prm

    def reset(self, seed: int | None = None) -> None:
        import torch

        if seed is not None:
            self._seed, self._draws = int(seed), 0
        self._time = 0.0
        prm = self._kernel_params()
        ctx = _lib.get_context()
        ctx.bind_torch_stream()
        f64 = dict(dtype=torch.float64, device=ctx.device)
        # the kernel reads the seed as an unsigned 64-bit integer: keep the bits of seeds above 2^63
        seeds = torch.from_numpy(np.array([self._seed], dtype=np.uint64).view(np.int64)).to(ctx.device)
        # row 0: the env's own stream (unused here), row 1: the target's stream
        draws = torch.tensor([[0], [self._draws]], dtype=torch.int64, device=ctx.device)
        state, tp = torch.empty((12, 1), **f64), torch.empty((12, 1), **f64)
        ctx.reset_batched(prm, seeds, state, tp, draws)
        self._tparams = tp.cpu().numpy().reshape(12)
        self._draws = int(draws[1, 0].item())

    @property
    def _pattern(self):
        return self._tparams

    def get_position(self, time: float) -> tuple[float, float, float]:
        if self._tparams is None:
            self.reset()
        p, _, _ = _observe_target(self._kernel_params(clamp=False), self._tparams, time)
        return tuple(p)

    def get_state(self, time: float) -> dict:
        if self._tparams is None:
            self.reset()
        p, v, a = _observe_target(self._kernel_params(), self._tparams, time)
        return {"position": p, "velocity": v, "acceleration": a}

    def step(self, dt: float) -> dict:
        self._time += dt
        return self.get_state(self._time)

    @property
    def current_time(self) -> float:
        return self._time
