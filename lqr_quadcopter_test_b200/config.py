"""Configuration types of the environment boundary.

Same names, fields and defaults as the reference's ``quadcopter_tracking.env.config``
(env/config.py:11-222) so that configuration dictionaries, YAML files and test fixtures written
for the reference load unchanged.  These are plain host-side containers; the numbers in them are
flattened into the ``qt_params`` block (params.py) that the CUDA kernels read.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field


@dataclass
class QuadcopterParams:
    """env/config.py:11-37.  Inertia, arm length, k_thrust and k_torque are carried for
    compatibility; the reference dynamics never read them (quadcopter_env.py:329-426)."""

    mass: float = 1.0
    arm_length: float = 0.25
    Ixx: float = 0.0082
    Iyy: float = 0.0082
    Izz: float = 0.0149
    k_thrust: float = 1.0e-5
    k_torque: float = 1.0e-7
    max_thrust: float = 20.0
    min_thrust: float = 0.0
    max_angular_rate: float = 3.0
    gravity: float = 9.81
    drag_coeff_linear: float = 0.1
    drag_coeff_angular: float = 0.01


@dataclass
class SimulationParams:
    """env/config.py:40-51"""

    dt: float = 0.01
    max_episode_time: float = 30.0
    integrator: str = "rk4"
    max_velocity: float = 50.0
    max_angular_velocity: float = 10.0
    max_position: float = 1000.0


@dataclass
class TargetParams:
    """env/config.py:54-65"""

    motion_type: str = "stationary"
    speed: float = 1.0
    amplitude: float = 2.0
    frequency: float = 0.5
    radius: float = 2.0
    center: tuple[float, float, float] = (0.0, 0.0, 1.0)
    max_acceleration: float = 5.0
    radius_requirement: float = 0.5


@dataclass
class SuccessCriteria:
    """env/config.py:68-74 (the env's own criteria, reported in ``info['success']``)."""

    min_on_target_ratio: float = 0.8
    min_episode_duration: float = 30.0
    target_radius: float = 0.5


@dataclass
class LoggingParams:
    """env/config.py:77-83"""

    enabled: bool = True
    log_interval: int = 10
    output_dir: str = "experiments"


def _from_mapping(cls, mapping: dict):
    """Build a dataclass from the keys of ``mapping`` it knows; unknown keys are ignored and
    missing keys keep the default, which is what the reference's per-field ``.get`` calls do."""
    known = {f.name for f in dataclasses.fields(cls)}
    return cls(**{k: v for k, v in mapping.items() if k in known})


@dataclass
class EnvConfig:
    """env/config.py:86-222"""

    seed: int = 42
    quadcopter: QuadcopterParams = field(default_factory=QuadcopterParams)
    simulation: SimulationParams = field(default_factory=SimulationParams)
    target: TargetParams = field(default_factory=TargetParams)
    success_criteria: SuccessCriteria = field(default_factory=SuccessCriteria)
    logging: LoggingParams = field(default_factory=LoggingParams)

    @classmethod
    def from_dict(cls, config_dict: dict) -> "EnvConfig":
        sim = dict(config_dict.get("simulation", {}))
        success = dict(config_dict.get("success_criteria", {}))
        target = dict(config_dict.get("target", {}))
        # legacy flat keys (env/config.py:114-118)
        if "dt" in config_dict and "dt" not in sim:
            sim["dt"] = config_dict["dt"]
        if "episode_length" in config_dict:
            sim["max_episode_time"] = config_dict["episode_length"]
        # target.radius_requirement seeds success_criteria.target_radius (env/config.py:120-126)
        if "radius_requirement" in target:
            success.setdefault("target_radius", target["radius_requirement"])
        if "center" in target:
            target["center"] = tuple(target["center"])
        return cls(
            seed=config_dict.get("seed", 42),
            quadcopter=_from_mapping(QuadcopterParams, config_dict.get("quadcopter", {})),
            simulation=_from_mapping(SimulationParams, sim),
            target=_from_mapping(TargetParams, target),
            success_criteria=_from_mapping(SuccessCriteria, success),
            logging=_from_mapping(LoggingParams, config_dict.get("logging", {})),
        )

    def to_dict(self) -> dict:
        out = {"seed": self.seed}
        for name in ("quadcopter", "simulation", "target", "success_criteria", "logging"):
            out[name] = dataclasses.asdict(getattr(self, name))
        for unused in ("arm_length", "Ixx", "Iyy", "Izz", "k_thrust", "k_torque"):
            pass  # kept in the dictionary exactly as the reference does
        return out
