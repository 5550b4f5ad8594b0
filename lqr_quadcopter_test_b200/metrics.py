"""Evaluation metrics containers (utils/metrics.py:26-433 of the reference).

The numbers themselves are produced on the GPU: the rollout kernel accumulates, per episode, the
16-entry metric vector described in include/qtrack.h (enum qt_metric) and ``qt_reduce_summary``
folds a whole evaluation into 16 sums.  This module only turns those vectors into the
reference's result objects.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import params as P


@dataclass
class SuccessCriteria:
    """utils/metrics.py:26-39"""

    min_on_target_ratio: float = 0.8
    min_episode_duration: float = 30.0
    target_radius: float = 0.5


@dataclass
class EpisodeMetrics:
    """utils/metrics.py:42-90"""

    episode_duration: float = 0.0
    on_target_ratio: float = 0.0
    mean_tracking_error: float = 0.0
    max_tracking_error: float = 0.0
    rms_tracking_error: float = 0.0
    total_control_effort: float = 0.0
    mean_control_effort: float = 0.0
    overshoot_count: int = 0
    max_overshoot: float = 0.0
    success: bool = False
    termination_reason: str = ""
    action_violations: int = 0

    def to_dict(self) -> dict:
        return dict(self.__dict__)


def episode_metrics_from_vector(v) -> EpisodeMetrics:
    """One column of the [16, E] metric array -> EpisodeMetrics (utils/metrics.py:264-338)."""
    n = float(v[P.M_STEPS])
    if n <= 0:
        return EpisodeMetrics(success=False, termination_reason="no_data")
    return EpisodeMetrics(
        episode_duration=float(v[P.M_FINAL_TIME]),
        on_target_ratio=float(v[P.M_MET_ON_TARGET]) / n,
        mean_tracking_error=float(v[P.M_SUM_ERR]) / n,
        max_tracking_error=float(v[P.M_MAX_ERR]),
        rms_tracking_error=math.sqrt(float(v[P.M_SUM_ERR2]) / n),
        total_control_effort=float(v[P.M_SUM_EFFORT]),
        mean_control_effort=float(v[P.M_SUM_EFFORT]) / n,
        overshoot_count=int(v[P.M_OVERSHOOT_COUNT]),
        max_overshoot=float(v[P.M_MAX_OVERSHOOT]),
        success=bool(v[P.M_MET_SUCCESS] != 0),
        termination_reason=P.TERMINATION_REASONS.get(int(v[P.M_TERMINATION]), ""),
        action_violations=int(v[P.M_VIOLATIONS]),
    )


class EpisodeMetricsView:
    """Lazy list of EpisodeMetrics over a host [16, E] array, so a million-episode evaluation does
    not materialise a million Python objects unless somebody iterates over them."""

    def __init__(self, metrics_host: np.ndarray):
        self._m = metrics_host

    def __len__(self) -> int:
        return int(self._m.shape[1])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [episode_metrics_from_vector(self._m[:, j]) for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return episode_metrics_from_vector(self._m[:, i])

    def __iter__(self):
        for j in range(len(self)):
            yield episode_metrics_from_vector(self._m[:, j])

    def __bool__(self) -> bool:
        return len(self) > 0


@dataclass
class EvaluationSummary:
    """utils/metrics.py:93-141"""

    total_episodes: int = 0
    successful_episodes: int = 0
    success_rate: float = 0.0
    mean_on_target_ratio: float = 0.0
    std_on_target_ratio: float = 0.0
    mean_tracking_error: float = 0.0
    std_tracking_error: float = 0.0
    mean_control_effort: float = 0.0
    best_episode_idx: int = 0
    worst_episode_idx: int = 0
    meets_criteria: bool = False
    episode_metrics: list = field(default_factory=list)

    def to_dict(self) -> dict:
        d = {k: v for k, v in self.__dict__.items() if k != "episode_metrics"}
        d["episode_metrics"] = [m.to_dict() for m in self.episode_metrics]
        return d


# layout of the 16 sums written by qt_reduce_summary (csrc/qt_reduce.cu)
(S_N, S_SUCCESSES, S_SUM_RATIO, S_SUM_ERR, S_SUM_EFFORT, S_M2_RATIO, S_M2_ERR, S_MAX_RATIO, S_ARGMAX,
 S_MIN_RATIO, S_ARGMIN, S_ENV_SUCCESSES, S_SUM_ENV_RATIO, S_SUM_STEPS, S_SUM_TOTAL_ERR,
 S_SUM_VIOLATIONS) = range(16)


def merge_summary_sums(parts: list[np.ndarray], offsets: list[int]) -> np.ndarray:
    """Combine the 16-sum vectors of several shards (ranks) into one.  ``offsets[r]`` is the global
    index of shard r's first episode.  Second moments are merged with the pairwise-variance
    formula, extrema keep numpy's first-occurrence rule."""
    out = np.zeros(16)
    n = sum(p[S_N] for p in parts)
    out[S_N] = n
    if n == 0:
        return out
    for k in (S_SUCCESSES, S_SUM_RATIO, S_SUM_ERR, S_SUM_EFFORT, S_ENV_SUCCESSES, S_SUM_ENV_RATIO,
              S_SUM_STEPS, S_SUM_TOTAL_ERR, S_SUM_VIOLATIONS):
        out[k] = sum(p[k] for p in parts)
    for m2, s in ((S_M2_RATIO, S_SUM_RATIO), (S_M2_ERR, S_SUM_ERR)):
        mean = out[s] / n
        out[m2] = sum(p[m2] + p[S_N] * (p[s] / p[S_N] - mean) ** 2 for p in parts if p[S_N] > 0)
    best = (-math.inf, math.inf)
    worst = (math.inf, math.inf)
    for p, off in zip(parts, offsets):
        if p[S_N] <= 0:
            continue
        cand = (p[S_MAX_RATIO], p[S_ARGMAX] + off)
        if cand[0] > best[0] or (cand[0] == best[0] and cand[1] < best[1]):
            best = cand
        cand = (p[S_MIN_RATIO], p[S_ARGMIN] + off)
        if cand[0] < worst[0] or (cand[0] == worst[0] and cand[1] < worst[1]):
            worst = cand
    out[S_MAX_RATIO], out[S_ARGMAX] = best
    out[S_MIN_RATIO], out[S_ARGMIN] = worst
    return out


def summary_from_sums(sums: np.ndarray, criteria: SuccessCriteria | None, episode_metrics) -> EvaluationSummary:
    """compute_evaluation_summary (utils/metrics.py:341-390) from the device-side sums."""
    criteria = criteria or SuccessCriteria()
    n = int(sums[S_N])
    if n == 0:
        return EvaluationSummary()
    mean_ratio = float(sums[S_SUM_RATIO]) / n
    return EvaluationSummary(
        total_episodes=n,
        successful_episodes=int(sums[S_SUCCESSES]),
        success_rate=float(sums[S_SUCCESSES]) / n,
        mean_on_target_ratio=mean_ratio,
        std_on_target_ratio=math.sqrt(max(float(sums[S_M2_RATIO]), 0.0) / n),
        mean_tracking_error=float(sums[S_SUM_ERR]) / n,
        std_tracking_error=math.sqrt(max(float(sums[S_M2_ERR]), 0.0) / n),
        mean_control_effort=float(sums[S_SUM_EFFORT]) / n,
        best_episode_idx=int(sums[S_ARGMAX]),
        worst_episode_idx=int(sums[S_ARGMIN]),
        meets_criteria=mean_ratio >= criteria.min_on_target_ratio,
        episode_metrics=episode_metrics,
    )


# ---- record post-processing (utils/metrics.py:144-390) ---------------------------------------------------------
# The reference's drivers call these on the per-step records they have just collected; here Evaluator / ControllerTuner
# never do — their per-episode metrics are accumulated inside the rollout kernel (the 16 rows of enum qt_metric) and
# reduced by qt_reduce_summary / the fused score epilogue.  The functions below exist for code that already HOLDS an
# episode record on the host (a list of step dictionaries, e.g. from Evaluator.run_episode or from the reference's own
# loop) and wants the reference's numbers for it: plain numpy on host data, the same operations in the same order as
# the reference, nothing of the accelerated path runs through them.
def compute_tracking_error(quad_positions: np.ndarray, target_positions: np.ndarray) -> np.ndarray:
    """utils/metrics.py:144-162"""
    if quad_positions.shape != target_positions.shape:
        raise ValueError(f"Shape mismatch: quad {quad_positions.shape} vs target {target_positions.shape}")
    return np.linalg.norm(target_positions - quad_positions, axis=1)


def compute_on_target_ratio(tracking_errors: np.ndarray, target_radius: float) -> float:
    """utils/metrics.py:165-179"""
    if len(tracking_errors) == 0:
        return 0.0
    return float(np.mean(tracking_errors <= target_radius))


def compute_control_effort(actions: np.ndarray) -> tuple[float, float]:
    """utils/metrics.py:182-202: (sum, mean) of the per-step 2-norm of the action vector."""
    if len(actions) == 0:
        return 0.0, 0.0
    mags = np.linalg.norm(actions, axis=1)
    return float(np.sum(mags)), float(np.mean(mags))


def detect_overshoots(tracking_errors: np.ndarray, target_radius: float, window_size: int = 10) -> tuple[int, float]:
    """utils/metrics.py:205-261: off-target streaks of >= window_size samples that started from an on-target sample
    (the streaming form of this scan is the overshoot detector inside the rollout kernel, csrc/qt_rollout.cu)."""
    if len(tracking_errors) < window_size:
        return 0, 0.0
    on = tracking_errors <= target_radius
    count, worst, in_phase, streak, current = 0, 0.0, False, 0, 0.0
    for i in range(1, len(on)):
        if on[i - 1] and not on[i]:
            in_phase, streak, current = True, 1, tracking_errors[i] - target_radius
        elif in_phase and not on[i]:
            streak += 1
            current = max(current, tracking_errors[i] - target_radius)
        elif in_phase and on[i]:
            if streak >= window_size:
                count += 1
                worst = max(worst, current)
            in_phase, streak, current = False, 0, 0.0
    if in_phase and streak >= window_size:
        count += 1
        worst = max(worst, current)
    return count, float(worst)


def compute_episode_metrics(episode_data: list[dict], criteria: SuccessCriteria | None = None,
                            episode_info: dict | None = None) -> EpisodeMetrics:
    """utils/metrics.py:264-338 for a recorded episode (list of step dictionaries)."""
    criteria = criteria or SuccessCriteria()
    if not episode_data:
        return EpisodeMetrics(success=False, termination_reason="no_data")
    quad = np.array([d["quadcopter_position"] for d in episode_data])
    target = np.array([d["target_position"] for d in episode_data])
    actions = np.array([d["action"] for d in episode_data])
    times = np.array([d["time"] for d in episode_data])
    errors = compute_tracking_error(quad, target)
    ratio = compute_on_target_ratio(errors, criteria.target_radius)
    total_effort, mean_effort = compute_control_effort(actions)
    overshoots, max_overshoot = detect_overshoots(errors, criteria.target_radius)
    duration = float(times[-1]) if len(times) > 0 else 0.0
    info = episode_info or {}
    return EpisodeMetrics(
        episode_duration=duration, on_target_ratio=ratio, mean_tracking_error=float(np.mean(errors)),
        max_tracking_error=float(np.max(errors)), rms_tracking_error=float(np.sqrt(np.mean(errors ** 2))),
        total_control_effort=total_effort, mean_control_effort=mean_effort, overshoot_count=overshoots,
        max_overshoot=max_overshoot,
        success=bool(duration >= criteria.min_episode_duration and ratio >= criteria.min_on_target_ratio),
        termination_reason=info.get("termination_reason", ""), action_violations=info.get("action_violations", 0))


def compute_evaluation_summary(episode_metrics_list, criteria: SuccessCriteria | None = None) -> EvaluationSummary:
    """utils/metrics.py:341-390 for a list of EpisodeMetrics."""
    criteria = criteria or SuccessCriteria()
    if not episode_metrics_list:
        return EvaluationSummary()
    n = len(episode_metrics_list)
    ratios = [m.on_target_ratio for m in episode_metrics_list]
    errors = [m.mean_tracking_error for m in episode_metrics_list]
    efforts = [m.mean_control_effort for m in episode_metrics_list]
    successes = sum(1 for m in episode_metrics_list if m.success)
    mean_ratio = float(np.mean(ratios))
    return EvaluationSummary(
        total_episodes=n, successful_episodes=successes, success_rate=successes / n, mean_on_target_ratio=mean_ratio,
        std_on_target_ratio=float(np.std(ratios)), mean_tracking_error=float(np.mean(errors)),
        std_tracking_error=float(np.std(errors)), mean_control_effort=float(np.mean(efforts)),
        best_episode_idx=int(np.argmax(ratios)), worst_episode_idx=int(np.argmin(ratios)),
        meets_criteria=mean_ratio >= criteria.min_on_target_ratio, episode_metrics=episode_metrics_list)


def format_metrics_report(summary: EvaluationSummary) -> str:
    """Plain-text report in the spirit of utils/metrics.py:393-433."""
    em = summary.episode_metrics
    lines = [
        "=" * 60, "EVALUATION SUMMARY", "=" * 60, "",
        f"Total Episodes: {summary.total_episodes}",
        f"Successful Episodes: {summary.successful_episodes}",
        f"Success Rate: {summary.success_rate:.1%}", "",
        "Tracking Performance:",
        f"  Mean On-Target Ratio: {summary.mean_on_target_ratio:.1%} (± {summary.std_on_target_ratio:.1%})",
        f"  Mean Tracking Error: {summary.mean_tracking_error:.3f}m (± {summary.std_tracking_error:.3f}m)",
        f"  Mean Control Effort: {summary.mean_control_effort:.3f}", "",
    ]
    if em:
        lines.append(f"Best Episode: #{summary.best_episode_idx + 1} "
                     f"({em[summary.best_episode_idx].on_target_ratio:.1%} on-target)")
        lines.append(f"Worst Episode: #{summary.worst_episode_idx + 1} "
                     f"({em[summary.worst_episode_idx].on_target_ratio:.1%} on-target)")
    else:
        lines += ["Best Episode: N/A", "Worst Episode: N/A"]
    lines += ["", f"SUCCESS CRITERIA MET: {'YES' if summary.meets_criteria else 'NO'}", "=" * 60]
    return "\n".join(lines)
