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
