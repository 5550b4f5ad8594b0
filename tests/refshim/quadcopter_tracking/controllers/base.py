from lqr_quadcopter_test_b200.controller import (ACTION_KEYS, DEFAULT_ACTION_LIMITS, ActionLimits, BaseController,  # noqa: F401
                                                 validate_action)
