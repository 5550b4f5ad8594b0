from lqr_quadcopter_test_b200.controller import (ACTION_KEYS, DEFAULT_ACTION_LIMITS, ActionLimits, BaseController,
                                                 LQRController, PIDController, RiccatiLQRController, validate_action)
from lqr_quadcopter_test_b200.policy import DeepTrackingPolicy, PolicyNetwork
from lqr_quadcopter_test_b200.tuning import ControllerTuner, GainSearchSpace, TuningConfig, TuningResult

VALID_CONTROLLER_TYPES = ("deep", "lqr", "pid", "riccati_lqr")
