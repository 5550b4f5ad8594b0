from lqr_quadcopter_test_b200.metrics import (EpisodeMetrics, EvaluationSummary, SuccessCriteria, compute_control_effort,  # noqa: F401
                                              compute_episode_metrics, compute_evaluation_summary, compute_on_target_ratio,
                                              compute_tracking_error, detect_overshoots, format_metrics_report)
