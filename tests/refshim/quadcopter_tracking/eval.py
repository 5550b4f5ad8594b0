from lqr_quadcopter_test_b200.eval import Evaluator, load_controller, run_hyperparameter_sweep  # noqa: F401
