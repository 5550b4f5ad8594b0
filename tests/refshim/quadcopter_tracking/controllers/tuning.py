from lqr_quadcopter_test_b200.tuning import *  # noqa: F401,F403
from lqr_quadcopter_test_b200.tuning import ControllerTuner, GainSearchSpace, TuningConfig, TuningResult  # noqa: F401
