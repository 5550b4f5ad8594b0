from lqr_quadcopter_test_b200.target_motion import *  # noqa: F401,F403
