from lqr_quadcopter_test_b200.config import (EnvConfig, LoggingParams, QuadcopterParams, SimulationParams, SuccessCriteria,
                                             TargetParams)
from lqr_quadcopter_test_b200.env import QuadcopterEnv
from lqr_quadcopter_test_b200.target_motion import (CircularMotion, Figure8Motion, LinearMotion, SinusoidalMotion,
                                                    StationaryMotion, TargetMotion)
