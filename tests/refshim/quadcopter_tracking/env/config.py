from lqr_quadcopter_test_b200.config import *  # noqa: F401,F403
