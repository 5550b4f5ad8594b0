from lqr_quadcopter_test_b200.env import QuadcopterEnv  # noqa: F401
