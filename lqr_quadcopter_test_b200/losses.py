"""Training losses of the deep tracking policy (drop-in for ``quadcopter_tracking.utils.losses``, utils/losses.py:25-406).

Host-side torch autograd code for the optimiser step of ``train.Trainer`` — the roll-outs that produce the batches run in
the fused kernel (qt_policy.cu); nothing here is on the accelerated path.  Weights may be scalars (w I) or full
matrices; the three error kinds are  e' W e  ("l2"),  |e W|_1  ("l1")  and Huber on  |e W|_2.
"""
from __future__ import annotations

import logging
import math

import numpy as np

logger = logging.getLogger(__name__)

ERROR_TYPES = ("l2", "l1", "huber")
SMOOTHING = ("none", "exp", "sigmoid")


def _torch():
    import torch

    return torch


def _weight_matrix(weight, dim: int, device):
    torch = _torch()
    if np.isscalar(weight):
        return torch.eye(dim, dtype=torch.float32, device=device) * float(weight)
    w = torch.as_tensor(np.asarray(weight) if not torch.is_tensor(weight) else weight, dtype=torch.float32).to(device)
    if w.dim() == 1:
        w = torch.diag(w)
    if tuple(w.shape) != (dim, dim):
        raise ValueError(f"Weight must be a scalar, a {dim}-vector or a {dim}x{dim} matrix, got shape {tuple(w.shape)}")
    return w


def weighted_error(error, weight, kind: str = "l2", delta: float = 1.0):
    """Mean over the batch of the weighted norm of ``error`` [B, d] under ``weight`` [d, d]."""
    torch = _torch()
    ew = error @ weight
    if kind == "l2":
        return (ew * error).sum(dim=-1).mean()
    if kind == "l1":
        return ew.abs().sum(dim=-1).mean()
    if kind == "huber":
        n = ew.norm(dim=-1)
        return torch.where(n <= delta, 0.5 * n * n, delta * (n - 0.5 * delta)).mean()
    raise ValueError(f"Unknown error type: {kind}")


class _Module:
    """The little of torch.nn.Module the callers use, without parameters of its own."""

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def to(self, device):
        torch = _torch()
        self.device = torch.device(device)
        for k, v in vars(self).items():
            if torch.is_tensor(v):
                setattr(self, k, v.to(self.device))
            elif isinstance(v, _Module):
                v.to(device)
        return self


class TrackingLoss(_Module):
    """position + velocity + control-effort terms; with ``action_target`` the control term is the imitation error."""

    def __init__(self, position_weight=1.0, velocity_weight=0.1, control_weight=0.01, error_type: str = "l2",
                 huber_delta: float = 1.0, device="cpu"):
        torch = _torch()
        self.error_type, self.huber_delta, self.device = error_type, huber_delta, torch.device(device)
        self.position_weight = _weight_matrix(position_weight, 3, self.device)
        self.velocity_weight = _weight_matrix(velocity_weight, 3, self.device)
        self.control_weight = _weight_matrix(control_weight, 4, self.device)

    def forward(self, position_error, velocity_error, action, action_target=None) -> dict:
        torch = _torch()
        terms = {
            "position": weighted_error(position_error, self.position_weight, self.error_type, self.huber_delta),
            "velocity": weighted_error(velocity_error, self.velocity_weight, self.error_type, self.huber_delta),
            "control": weighted_error(action if action_target is None else action - action_target, self.control_weight,
                                      self.error_type, self.huber_delta),
        }
        total = terms["position"] + terms["velocity"] + terms["control"]
        if torch.isnan(total):            # keep the optimiser alive: a NaN term counts as a large finite loss
            logger.warning("NaN detected in loss computation")
            total = sum(torch.nan_to_num(t, nan=10.0) for t in terms.values())
        return {"total": total, **terms}


class RewardShapingLoss(_Module):
    """on-target bonus minus a (optionally smoothed) distance penalty, as a reward to maximise."""

    def __init__(self, target_radius: float = 0.5, on_target_bonus: float = 1.0, distance_penalty: float = 1.0,
                 smoothing: str = "none", device="cpu"):
        self.target_radius, self.on_target_bonus = target_radius, on_target_bonus
        self.distance_penalty, self.smoothing = distance_penalty, smoothing
        self.device = _torch().device(device)

    def forward(self, tracking_error) -> dict:
        torch = _torch()
        e = tracking_error.reshape(1) if tracking_error.dim() == 0 else tracking_error
        on_target = e <= self.target_radius
        bonus = on_target.float() * self.on_target_bonus
        if self.smoothing == "none":
            shaped = e
        elif self.smoothing == "exp":
            shaped = 1.0 - torch.exp(-e)
        elif self.smoothing == "sigmoid":
            shaped = torch.sigmoid(e - self.target_radius)
        else:
            raise ValueError(f"Unknown smoothing: {self.smoothing}")
        penalty = shaped * self.distance_penalty
        return {"reward": (bonus - penalty).mean(), "on_target": on_target, "bonus": bonus.mean(), "penalty": penalty.mean()}


class CombinedLoss(_Module):
    """tracking_weight * tracking loss - reward_weight * shaped reward."""

    def __init__(self, tracking_loss: TrackingLoss | None = None, reward_loss: RewardShapingLoss | None = None,
                 tracking_weight: float = 1.0, reward_weight: float = 0.0, device="cpu"):
        self.device = _torch().device(device)
        self.tracking_loss = tracking_loss or TrackingLoss(device=device)
        self.reward_loss = reward_loss or RewardShapingLoss(device=device)
        self.tracking_weight, self.reward_weight = tracking_weight, reward_weight

    def forward(self, position_error, velocity_error, action, tracking_error, action_target=None) -> dict:
        t = self.tracking_loss(position_error, velocity_error, action, action_target)
        r = self.reward_loss(tracking_error)
        return {
            "total": self.tracking_weight * t["total"] - self.reward_weight * r["reward"],
            "tracking_total": t["total"], "tracking_position": t["position"], "tracking_velocity": t["velocity"],
            "tracking_control": t["control"], "reward": r["reward"], "on_target_ratio": r["on_target"].float().mean(),
        }


def create_loss_from_config(config: dict) -> CombinedLoss:
    """Keys: position_weight, velocity_weight, control_weight, error_type, huber_delta, target_radius, on_target_bonus,
    distance_penalty, smoothing, tracking_weight, reward_weight, device (utils/losses.py:362-406)."""
    device = config.get("device", "cpu")
    tracking = TrackingLoss(config.get("position_weight", 1.0), config.get("velocity_weight", 0.1),
                            config.get("control_weight", 0.01), config.get("error_type", "l2"),
                            config.get("huber_delta", 1.0), device)
    reward = RewardShapingLoss(config.get("target_radius", 0.5), config.get("on_target_bonus", 1.0),
                               config.get("distance_penalty", 1.0), config.get("smoothing", "none"), device)
    return CombinedLoss(tracking, reward, config.get("tracking_weight", 1.0), config.get("reward_weight", 0.0), device)


class LossLogger:
    """Per-epoch means of the logged loss components (utils/losses.py:409-484).  Device tensors stay on the device
    until the epoch ends: one read-back per component and epoch, not one per optimiser step."""

    def __init__(self):
        self.current: dict[str, list] = {}
        self.history: list[dict[str, float]] = []

    def log(self, losses: dict) -> None:
        for k, v in losses.items():
            self.current.setdefault(k, []).append(v)

    def end_epoch(self) -> dict[str, float]:
        out = {}
        for k, vals in self.current.items():
            if not vals:
                continue
            if any(hasattr(v, "detach") for v in vals):
                torch = _torch()
                dev = next(v.device for v in vals if hasattr(v, "detach"))
                host = torch.stack([v.detach().to(torch.float64).reshape(()) if hasattr(v, "detach")
                                    else torch.as_tensor(float(v), dtype=torch.float64, device=dev) for v in vals]).cpu().numpy()
                out[k] = float(np.mean(host))
            else:
                out[k] = float(np.mean(vals))
        self.history.append(out)
        self.current = {}
        return out

    def get_history(self) -> list[dict[str, float]]:
        return self.history

    def get_best_epoch(self, metric: str = "total") -> tuple[int, float]:
        vals = [(i, h[metric]) for i, h in enumerate(self.history) if metric in h and not math.isnan(h[metric])]
        if not vals:
            return -1, float("inf")
        return min(vals, key=lambda t: t[1])

    def reset(self) -> None:
        self.current, self.history = {}, []
