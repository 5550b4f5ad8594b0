from lqr_quadcopter_test_b200.losses import (CombinedLoss, LossLogger, RewardShapingLoss, TrackingLoss,  # noqa: F401
                                             create_loss_from_config)
