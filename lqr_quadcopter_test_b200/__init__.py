"""qtrack-b200: batched closed-loop quadcopter tracking + Riccati-LQR/LQI on NVIDIA B200.

A from-scratch, B200-native implementation of ONE hot path of ``quadcopter_tracking``
(AgentFoundryExamples/lqr-quadcopter-test): environment reset/step, target motion, the
Riccati-LQR/LQI control law, evaluation metrics, the controller auto-tuner's candidate evaluation
and the batched DARE solve that produces the gains — behind the reference's own Python
interfaces.  All arithmetic runs in hand-written sm_100a CUDA kernels (csrc/, C ABI in
include/qtrack.h); this package is the thin host layer.  There is no CPU fallback.
"""
from .config import (EnvConfig, LoggingParams, QuadcopterParams, SimulationParams, TargetParams)
from .config import SuccessCriteria as EnvSuccessCriteria
from .controller import (ACTION_KEYS, DEFAULT_ACTION_LIMITS, ActionLimits, BaseController, LQRController, PIDController,
                         RiccatiLQRController, validate_action,
                         build_augmented_lqi_system, build_linearized_system, heuristic_lqr_gain,
                         solve_dare, solve_dare_batched)
from .env import BatchedQuadcopterEnv, QuadcopterEnv
from .policy import DeepTrackingPolicy, create_controller_from_config, make_policy_network
from .train import Trainer, TrainingConfig, load_checkpoint_and_resume
from .eval import Evaluator, load_controller, rollout_device, run_hyperparameter_sweep
from .metrics import (EpisodeMetrics, EvaluationSummary, SuccessCriteria, compute_control_effort, compute_episode_metrics,
                      compute_evaluation_summary, compute_on_target_ratio, compute_tracking_error, detect_overshoots,
                      format_metrics_report)
from .target_motion import (CircularMotion, Figure8Motion, LinearMotion, SinusoidalMotion, StationaryMotion, TargetMotion)
from .tuning import ControllerTuner, GainSearchSpace, TuningConfig, TuningResult
from .dataset import teacher_dataset

__version__ = "0.1.0"

__all__ = [
    "EnvConfig", "QuadcopterParams", "SimulationParams", "TargetParams", "LoggingParams",
    "EnvSuccessCriteria", "SuccessCriteria", "QuadcopterEnv", "BatchedQuadcopterEnv",
    "RiccatiLQRController", "LQRController", "PIDController", "DeepTrackingPolicy",
    "create_controller_from_config", "make_policy_network", "Trainer", "TrainingConfig",
    "load_checkpoint_and_resume", "run_hyperparameter_sweep", "BaseController", "ACTION_KEYS", "solve_dare",
    "solve_dare_batched", "build_linearized_system", "build_augmented_lqi_system",
    "heuristic_lqr_gain", "Evaluator", "load_controller", "rollout_device", "EpisodeMetrics",
    "EvaluationSummary", "format_metrics_report", "ControllerTuner", "GainSearchSpace",
    "TuningConfig", "TuningResult", "teacher_dataset", "TargetMotion", "LinearMotion", "CircularMotion",
    "SinusoidalMotion", "Figure8Motion", "StationaryMotion", "ActionLimits", "DEFAULT_ACTION_LIMITS", "validate_action",
    "compute_tracking_error", "compute_on_target_ratio", "compute_control_effort", "detect_overshoots",
    "compute_episode_metrics", "compute_evaluation_summary",
]
