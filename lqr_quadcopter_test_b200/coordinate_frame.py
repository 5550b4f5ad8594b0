"""ENU (East-North-Up) frame conventions and their assertions (drop-in for
``quadcopter_tracking.utils.coordinate_frame``, utils/coordinate_frame.py:44-405).

Host-side checks only: the kernels hard-wire the same conventions (qt_step.cuh: gravity along -z, body thrust along
+z, +pitch rate -> +x, +roll rate -> -y), and the reference's own tests use these helpers to probe them through the
env and controller interfaces.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

AXIS_X, AXIS_Y, AXIS_Z = 0, 1, 2                       # East, North, Up
GRAVITY_DIRECTION_ENU = np.array([0.0, 0.0, -1.0])
THRUST_DIRECTION_BODY = np.array([0.0, 0.0, 1.0])
PITCH_RATE_TO_X_VEL_SIGN = +1.0
ROLL_RATE_TO_Y_VEL_SIGN = -1.0
THRUST_TO_Z_ACCEL_SIGN = +1.0

DIRECTION_DOT_PRODUCT_TOLERANCE = 0.99                 # cosine of ~8 degrees
ZERO_MAGNITUDE_THRESHOLD = 1e-10
DEFAULT_CONTROL_SIGN_ERROR_THRESHOLD = 0.5             # metres of position error before a sign is checked


@dataclass(frozen=True)
class CoordinateFrame:
    name: str
    x_direction: str
    y_direction: str
    z_direction: str
    gravity_axis: str
    gravity_sign: str

    def __str__(self) -> str:
        return f"{self.name}: X={self.x_direction}, Y={self.y_direction}, Z={self.z_direction}"


ENU_FRAME = CoordinateFrame("ENU", "East", "North", "Up", "z", "-")
NED_FRAME = CoordinateFrame("NED", "North", "East", "Down", "z", "+")   # for comparison only, never used


def get_current_frame() -> CoordinateFrame:
    return ENU_FRAME


class ENUFrameError(ValueError):
    """A vector, control sign or observation contradicts the ENU conventions."""


def _unit(v, what: str, allow_zero: bool):
    v = np.asarray(v, dtype=np.float64)
    mag = float(np.linalg.norm(v))
    if mag < ZERO_MAGNITUDE_THRESHOLD:
        if allow_zero:
            return None
        raise ENUFrameError(f"{what} vector has zero magnitude")
    return v / mag


def assert_gravity_direction_enu(gravity_vector) -> None:
    g = _unit(gravity_vector, "Gravity", allow_zero=False)
    if float(g @ GRAVITY_DIRECTION_ENU) < DIRECTION_DOT_PRODUCT_TOLERANCE:
        raise ENUFrameError(f"Gravity should point in -Z direction for the ENU frame, got direction {g}")


def assert_thrust_direction_enu(thrust_body) -> None:
    t = _unit(thrust_body, "Thrust", allow_zero=True)          # zero thrust has no direction to check
    if t is not None and float(t @ THRUST_DIRECTION_BODY) < DIRECTION_DOT_PRODUCT_TOLERANCE:
        raise ENUFrameError(f"Thrust should point in +Z direction in the body frame, got direction {t}")


def assert_control_signs_enu(pos_error, pitch_rate: float, roll_rate: float,
                             error_threshold: float = DEFAULT_CONTROL_SIGN_ERROR_THRESHOLD) -> None:
    """A position error beyond ``error_threshold`` must be answered by a rate of the right sign:
    +X error -> +pitch rate, +Y error -> -roll rate (a zero rate counts as wrong)."""
    e = np.asarray(pos_error, dtype=np.float64)
    checks = ((e[AXIS_X], pitch_rate * PITCH_RATE_TO_X_VEL_SIGN, "X", "pitch_rate"),
              (e[AXIS_Y], roll_rate * ROLL_RATE_TO_Y_VEL_SIGN, "Y", "roll_rate"))
    for err, effect, axis, name in checks:
        if abs(err) <= error_threshold:
            continue
        if err * effect <= 0.0:
            sign = "+" if err > 0 else "-"
            raise ENUFrameError(f"{sign}{axis} error ({err:.3f} m) must be answered by a {name} that moves the vehicle "
                                f"towards {sign}{axis} in ENU, got pitch_rate={pitch_rate:.3f}, roll_rate={roll_rate:.3f}")


def assert_z_up(position, min_altitude: float = -100.0) -> None:
    z = float(np.asarray(position, dtype=np.float64)[AXIS_Z])
    if z < min_altitude:
        raise ENUFrameError(f"Z-axis position {z:.2f} is below minimum altitude {min_altitude:.2f}: Z must point up (ENU), "
                            "a large negative altitude suggests a NED frame")


def validate_observation_frame(observation: dict) -> bool:
    quad = observation.get("quadcopter", {})
    target = observation.get("target", {})
    if "position" in quad:
        assert_z_up(quad["position"])
    if "position" in target:
        tz = float(np.asarray(target["position"], dtype=np.float64)[AXIS_Z])
        if tz < -10.0:
            raise ENUFrameError(f"Target Z position {tz:.2f} is very negative: targets fly above ground in ENU")
    return True


def compute_expected_hover_thrust(mass: float, gravity: float = 9.81) -> float:
    return float(mass) * float(gravity)


def compute_position_error_enu(quad_position, target_position) -> np.ndarray:
    """target - current: a positive component means the target lies ahead along that ENU axis."""
    return np.asarray(target_position, dtype=np.float64) - np.asarray(quad_position, dtype=np.float64)
