from lqr_quadcopter_test_b200.train import Trainer, TrainingConfig, load_checkpoint_and_resume  # noqa: F401
