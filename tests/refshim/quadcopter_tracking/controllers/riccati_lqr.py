from lqr_quadcopter_test_b200.controller import (RiccatiLQRController, build_augmented_lqi_system,  # noqa: F401
                                                 build_linearized_system, solve_dare)
