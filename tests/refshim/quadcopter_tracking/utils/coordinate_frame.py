from lqr_quadcopter_test_b200.coordinate_frame import *  # noqa: F401,F403
from lqr_quadcopter_test_b200.coordinate_frame import (AXIS_X, AXIS_Y, AXIS_Z, DEFAULT_CONTROL_SIGN_ERROR_THRESHOLD,  # noqa: F401
                                                       DIRECTION_DOT_PRODUCT_TOLERANCE, ENU_FRAME, GRAVITY_DIRECTION_ENU,
                                                       NED_FRAME, PITCH_RATE_TO_X_VEL_SIGN, ROLL_RATE_TO_Y_VEL_SIGN,
                                                       THRUST_DIRECTION_BODY, THRUST_TO_Z_ACCEL_SIGN, ZERO_MAGNITUDE_THRESHOLD,
                                                       CoordinateFrame, ENUFrameError, assert_control_signs_enu,
                                                       assert_gravity_direction_enu, assert_thrust_direction_enu, assert_z_up,
                                                       compute_expected_hover_thrust, compute_position_error_enu,
                                                       get_current_frame, validate_observation_frame)
