"""Imitation / tracking training of the deep policy with the closed loop on the GPU.

Mirrors ``quadcopter_tracking.train`` (train.py:60-1045): ``TrainingConfig`` with the same fields,
``Trainer.train()`` with the same epoch structure, losses (utils/losses.py:24-370), checkpoints and
experiment logs.  What changes is where the work happens:

* data collection — the reference steps one env at a time through ``network(features)`` and
  ``supervisor.compute_action(obs)`` (train.py:668-721).  Here ALL episodes of an epoch run in one
  ``qt_rollout_policy`` launch: the student's network acts inside the kernel, the supervisor law labels
  every visited state, and features + labels land in CUDA tensors;
* updates — one optimiser step per episode on ``batch_size`` random steps of that episode, exactly
  the reference's schedule (train.py:751-862), evaluated by torch autograd on those tensors.

Two update schedules (``Trainer(config, update_schedule=...)``):

* ``"epoch"`` (default, the fast one): the episodes of one epoch are driven by the weights the epoch started with —
  they run concurrently in one launch — and the per-episode optimiser steps follow.  The reference updates the
  network BETWEEN the episodes of an epoch (train.py:660-735), so its episode i+1 is driven by the weights after
  episode i; the staleness here is below one epoch and vanishes for ``episodes_per_epoch = 1``.
* ``"episode"`` (exact): one launch per episode, the optimiser step right after it, batches drawn with
  ``np.random.choice`` from numpy's global generator — the reference's order of operations, step for step.  This is
  the mode ``tests/test_train_gpu.py`` holds to a loss curve recorded from the reference's own trainer
  (``tests/golden/trainer.npz``, ``oracle/gen_golden.py::gen_trainer``).
"""
from __future__ import annotations

import csv
import datetime
import json
import logging
import math
from pathlib import Path

import numpy as np

from . import _lib
from . import params as P
from .config import EnvConfig
from .controller import LQRController, PIDController, RiccatiLQRController
from .eval import VALID_CONTROLLER_TYPES, rollout_device
from .losses import LossLogger, create_loss_from_config
from .metrics import SuccessCriteria
from .policy import DeepTrackingPolicy

logger = logging.getLogger(__name__)


class TrainingConfig:
    """train.py:60-273 — same fields, defaults and validation."""

    VALID_TRAINING_MODES = ("tracking", "imitation", "reward_weighted")
    _FIELDS = dict(
        controller="deep", epochs=100, episodes_per_epoch=10, max_steps_per_episode=3000, batch_size=32,
        learning_rate=0.001, optimizer="adam", weight_decay=0.0, grad_clip=1.0, hidden_sizes=None, activation="relu",
        position_weight=1.0, velocity_weight=0.1, control_weight=0.01, error_type="l2", tracking_weight=1.0,
        reward_weight=0.0, training_mode="tracking", supervisor_controller="pid", imitation_weight=1.0,
        supervisor_blend_ratio=0.0, env_seed=42, target_motion_type="circular", episode_length=30.0, target_radius=0.5,
        checkpoint_dir="checkpoints", checkpoint_interval=10, save_best=True, log_dir="experiments/logs", log_interval=1,
        use_curriculum=False, curriculum_start_difficulty=0.5, curriculum_end_difficulty=1.0, nan_recovery_attempts=3,
        lr_reduction_factor=0.5, device=None, diagnostics_enabled=False, diagnostics_log_observations=True,
        diagnostics_log_actions=True, diagnostics_log_gradients=True, diagnostics_log_interval=10,
        diagnostics_output_dir=None, diagnostics_generate_plots=True)

    def __init__(self, **kwargs):
        unknown = set(kwargs) - set(self._FIELDS)
        if unknown:
            raise TypeError(f"TrainingConfig.__init__() got unexpected keyword arguments: {sorted(unknown)}")
        for k, default in self._FIELDS.items():
            setattr(self, k, kwargs.get(k, default))
        if self.controller not in VALID_CONTROLLER_TYPES:
            raise ValueError(f"Invalid controller type: '{self.controller}'. "
                             f"Valid choices are: {', '.join(VALID_CONTROLLER_TYPES)}")
        self.hidden_sizes = self.hidden_sizes or [64, 64]
        if self.training_mode not in self.VALID_TRAINING_MODES:
            raise ValueError(f"Invalid training_mode: '{self.training_mode}'. "
                             f"Valid choices are: {', '.join(self.VALID_TRAINING_MODES)}")
        if self.supervisor_controller not in ("pid", "lqr", "riccati_lqr"):
            raise ValueError(f"Invalid supervisor_controller: '{self.supervisor_controller}'. "
                             "Valid choices are: pid, lqr, riccati_lqr")
        if self.device is None:
            self.device = "cuda"
        self.full_config: dict = {}

    @classmethod
    def from_dict(cls, config_dict: dict) -> "TrainingConfig":
        inst = cls(**{k: v for k, v in config_dict.items() if k in cls._FIELDS})
        inst.full_config = config_dict
        return inst

    @classmethod
    def from_file(cls, path: str | Path) -> "TrainingConfig":
        path = Path(path)
        with open(path) as f:
            if path.suffix in (".yaml", ".yml"):
                import yaml

                d = yaml.safe_load(f)
            elif path.suffix == ".json":
                d = json.load(f)
            else:
                raise ValueError(f"Unsupported config format: {path.suffix}")
        return cls.from_dict(d or {})

    def to_dict(self) -> dict:
        return {k: getattr(self, k) for k in self._FIELDS}


def _weighted_error(torch, err, weight: float, kind: str, delta: float = 1.0):
    """TrackingLoss._compute_weighted_error (utils/losses.py) for scalar weights (weight * I)."""
    weighted = err * weight
    if kind == "l2":
        return (weighted * err).sum(dim=-1).mean()
    if kind == "l1":
        return weighted.abs().sum(dim=-1).mean()
    if kind == "huber":
        n = weighted.norm(dim=-1)
        return torch.where(n <= delta, 0.5 * n ** 2, delta * (n - 0.5 * delta)).mean()
    raise ValueError(f"Unknown error type: {kind}")


class Trainer:
    """train.py:275-1028."""

    def __init__(self, config: TrainingConfig, update_schedule: str = "epoch"):
        import torch

        if update_schedule not in ("epoch", "episode"):
            raise ValueError(f"update_schedule must be 'epoch' or 'episode', got {update_schedule!r}")
        self.update_schedule = update_schedule
        self.config = config
        self.full_config = getattr(config, "full_config", {})
        self._ctx = _lib.get_context()           # raises without a GPU: there is no CPU training path
        self.device = self._ctx.device
        self.is_deep_controller = config.controller == "deep"
        self.controller = self._create_controller()
        self.env_config = EnvConfig()
        self.env_config.seed = config.env_seed
        self.env_config.simulation.max_episode_time = config.episode_length
        self.env_config.target.motion_type = config.target_motion_type
        self.env_config.success_criteria.target_radius = config.target_radius
        self.criteria = SuccessCriteria(target_radius=config.target_radius)
        from .env import QuadcopterEnv

        self.env = QuadcopterEnv(config=self.env_config)     # train.py:300: scalar view of the same configuration
        self.optimizer = self._create_optimizer() if self.is_deep_controller else None
        self.loss_logger = LossLogger()
        # the reference's loss object (train.py:318-330), kept for callers that inspect it; the optimiser step below
        # evaluates the same terms on the roll-out batches
        self.loss_fn = create_loss_from_config({
            "position_weight": config.position_weight, "velocity_weight": config.velocity_weight,
            "control_weight": config.control_weight, "error_type": config.error_type,
            "target_radius": config.target_radius, "device": str(self.device)})
        self.experiment_log: list[dict] = []
        self.checkpoint_dir = Path(config.checkpoint_dir)
        if self.is_deep_controller:
            self.checkpoint_dir.mkdir(parents=True, exist_ok=True)
        self.log_dir = Path(config.log_dir)
        self.log_dir.mkdir(parents=True, exist_ok=True)
        self.current_epoch = 0
        self.best_loss = float("inf")
        self.nan_recovery_count = 0
        stamp = datetime.datetime.now(datetime.timezone.utc).strftime("%Y%m%d_%H%M%S_%f")
        self.experiment_id = f"train_{config.controller}_{stamp}_{config.env_seed}"
        self._step_counter = 0
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(int(config.env_seed))
        self.supervisor = None
        if self.is_deep_controller and self._uses_supervisor_mode:
            self.supervisor = self._create_supervisor()

    @property
    def _uses_supervisor_mode(self) -> bool:
        return self.config.training_mode in ("imitation", "reward_weighted")

    def _create_controller(self):
        c = self.config
        if c.controller == "deep":
            return DeepTrackingPolicy(config={
                "hidden_sizes": c.hidden_sizes, "activation": c.activation,
                "output_bounds": {"thrust": (0.0, 20.0), "roll_rate": (-3.0, 3.0), "pitch_rate": (-3.0, 3.0),
                                  "yaw_rate": (-3.0, 3.0)}})
        if c.controller == "pid":
            return PIDController(config=self.full_config.get("pid", {}))
        if c.controller == "lqr":
            return LQRController(config=self.full_config.get("lqr", {}))
        if c.controller == "riccati_lqr":
            return RiccatiLQRController(config=self.full_config.get("riccati_lqr", {}))
        if c.controller == "lqi":
            cfg = dict(self.full_config.get("riccati_lqr", {}), use_lqi=True)
            cfg.setdefault("q_int", [0.01, 0.01, 0.1])
            return RiccatiLQRController(config=cfg)
        raise ValueError(f"Unknown controller type: {c.controller}")

    def _create_supervisor(self):
        kind = self.config.supervisor_controller
        cfg = self.full_config.get(kind, {})
        return {"pid": PIDController, "lqr": LQRController, "riccati_lqr": RiccatiLQRController}[kind](config=cfg)

    def _create_optimizer(self):
        import torch.optim as optim

        c, params = self.config, self.controller.get_parameters()
        if c.optimizer == "adam":
            return optim.Adam(params, lr=c.learning_rate, weight_decay=c.weight_decay)
        if c.optimizer == "sgd":
            return optim.SGD(params, lr=c.learning_rate, weight_decay=c.weight_decay, momentum=0.9)
        if c.optimizer == "adamw":
            return optim.AdamW(params, lr=c.learning_rate, weight_decay=c.weight_decay)
        raise ValueError(f"Unknown optimizer: {c.optimizer}")

    # -- epochs ------------------------------------------------------------------------------------
    def _epoch_seeds(self):
        import torch

        n = self.config.episodes_per_epoch  # train.py:662: env_seed + epoch * 1000 + episode
        return torch.arange(n, dtype=torch.int64, device=self.device) + int(self.config.env_seed + self.current_epoch * 1000)

    def _episode_stats(self, out, first_err=None) -> dict:
        """Per-epoch means the reference logs (train.py:723-749): summed reward, the env's on-target
        ratio and the last tracking error of each episode."""
        M = out["metrics"]
        steps = M[P.M_STEPS].clamp(min=1.0)
        ratio = M[P.M_ENV_ON_TARGET] / steps
        # reward_k = -(post-step error); sum over an episode = -(sum of pre-step errors - first + last)
        if first_err is None:
            first_err = out["traj"][0, P.TRAJ_TARGET_P:P.TRAJ_TARGET_P + 3] - out["traj"][0, P.TRAJ_STATE:P.TRAJ_STATE + 3]
            first_err = first_err.norm(dim=0)
        reward = -(M[P.M_SUM_ERR] - first_err + M[P.M_FINAL_ERR])
        return {"mean_reward": float(reward.mean()), "mean_on_target_ratio": float(ratio.mean()),
                "mean_tracking_error": float(M[P.M_FINAL_ERR].mean())}

    def _evaluate_epoch(self) -> dict:
        """train.py:578-652: classical controllers are only evaluated (reset before every episode)."""
        out = rollout_device(self.controller, self.env_config, self.criteria, self._epoch_seeds(),
                             self.config.max_steps_per_episode, chain=1, record_stride=10 ** 9, ctx=self._ctx)
        stats = self._episode_stats(out)
        for _ in range(self.config.episodes_per_epoch):
            e = stats["mean_tracking_error"]
            self.loss_logger.log({"total": e, "position": e, "velocity": 0.0, "control": 0.0})
        return {**stats, "difficulty": 1.0}

    def _train_epoch(self, difficulty: float = 1.0) -> dict:
        """train.py:654-749 with the data collection fused into one launch."""
        import torch

        c = self.config
        if self.update_schedule == "episode":
            return self._train_epoch_exact(difficulty)
        out = rollout_device(self.controller, self.env_config, self.criteria, self._epoch_seeds(),
                             c.max_steps_per_episode, dataset_stride=1, ctx=self._ctx, supervisor=self.supervisor)
        feats, labels = out["features"], out["actions"]          # [slots, E, 18], [slots, E, 4]
        steps = out["metrics"][P.M_STEPS].to(torch.int64)
        first_err = feats[0, :, 0:3].to(torch.float64).norm(dim=1)
        steps_host = steps.cpu().tolist()
        for ep, n in enumerate(steps_host):
            if n <= 0:
                continue
            self._step_counter += 1
            self._update_from_episode(feats[:n, ep], labels[:n, ep] if self.supervisor is not None else None)
        return {**self._episode_stats(out, first_err), "difficulty": difficulty}

    def _train_epoch_exact(self, difficulty: float = 1.0) -> dict:
        """train.py:654-749 in the reference's own order: episode, optimiser step, next episode with the new weights."""
        import torch

        c = self.config
        seeds = self._epoch_seeds()
        stats = []
        for ep in range(c.episodes_per_epoch):
            out = rollout_device(self.controller, self.env_config, self.criteria, seeds[ep:ep + 1].contiguous(),
                                 c.max_steps_per_episode, dataset_stride=1, ctx=self._ctx, supervisor=self.supervisor)
            n = int(out["metrics"][P.M_STEPS, 0].item())
            if n <= 0:
                continue
            feats, labels = out["features"], out["actions"]
            first_err = feats[0, :, 0:3].to(torch.float64).norm(dim=1)
            stats.append(self._episode_stats(out, first_err))
            self._step_counter += 1
            self._update_from_episode(feats[:n, 0], labels[:n, 0] if self.supervisor is not None else None,
                                      numpy_batches=True)
        keys = ("mean_reward", "mean_on_target_ratio", "mean_tracking_error")
        return {**{k: float(np.mean([s[k] for s in stats])) if stats else 0.0 for k in keys}, "difficulty": difficulty}

    def _update_from_episode(self, features, supervisor_actions, numpy_batches: bool = False) -> None:
        """train.py:751-862: one optimiser step on ``batch_size`` random steps of one episode."""
        import torch
        import torch.nn.functional as F

        c = self.config
        n = int(features.shape[0])
        batch = min(c.batch_size, n)
        if numpy_batches:   # the reference's draw: numpy's global generator, without replacement (train.py:757)
            idx = torch.as_tensor(np.random.choice(n, batch, replace=False), device=self.device)
        else:
            idx = torch.randperm(n, device=self.device, generator=self._gen)[:batch]
        x = features[idx]
        actions = self.controller.network(x)
        pos_error, vel_error = x[:, 0:3], x[:, 3:6]
        tracking_error = pos_error.norm(dim=1)
        target = supervisor_actions[idx] if supervisor_actions is not None else None

        def tracking_losses(action_target=None):
            pos = _weighted_error(torch, pos_error, c.position_weight, c.error_type)
            vel = _weighted_error(torch, vel_error, c.velocity_weight, c.error_type)
            ctl = _weighted_error(torch, actions - action_target if action_target is not None else actions,
                                  c.control_weight, c.error_type)
            tot = pos + vel + ctl
            on_target = (tracking_error <= c.target_radius).float()
            reward = (on_target * 1.0 - tracking_error * 1.0).mean()     # RewardShapingLoss defaults
            return {"total": c.tracking_weight * tot - c.reward_weight * reward, "tracking_total": tot,
                    "tracking_position": pos, "tracking_velocity": vel, "tracking_control": ctl, "reward": reward,
                    "on_target_ratio": on_target.mean()}

        if c.training_mode == "imitation" and target is not None:
            losses = tracking_losses()
            imitation = F.mse_loss(actions, target)
            total = c.imitation_weight * imitation + c.tracking_weight * losses["total"]
            losses = {**losses, "imitation": imitation, "total": total}
        elif c.training_mode == "reward_weighted":
            losses = tracking_losses(target)
            total = losses["total"]
        else:
            losses = tracking_losses()
            total = losses["total"]
        self.loss_logger.log({k: v.detach() for k, v in losses.items()})
        self.optimizer.zero_grad()
        total.backward()
        if c.grad_clip > 0:
            torch.nn.utils.clip_grad_norm_(self.controller.get_parameters(), c.grad_clip)
        self.optimizer.step()

    def _get_curriculum_difficulty(self, epoch: int) -> float:
        c = self.config
        if not c.use_curriculum:
            return 1.0
        progress = epoch / max(1, c.epochs - 1)
        return c.curriculum_start_difficulty + progress * (c.curriculum_end_difficulty - c.curriculum_start_difficulty)

    def _check_training_failure(self, metrics: dict) -> bool:
        total = metrics.get("total", 0.0)
        return math.isnan(total) or math.isinf(total)

    def _recover_from_nan(self) -> bool:
        self.nan_recovery_count += 1
        if self.nan_recovery_count > self.config.nan_recovery_attempts:
            return False
        self.config.learning_rate *= self.config.lr_reduction_factor
        self.optimizer = self._create_optimizer()
        best = self.checkpoint_dir / f"{self.experiment_id}_best.pt"
        if best.exists():
            self.controller.load_checkpoint(best)
        return True

    def train(self) -> dict:
        """train.py:502-576"""
        self._save_config()
        if self.is_deep_controller:
            self.controller.train_mode()
        for epoch in range(self.config.epochs):
            self.current_epoch = epoch
            difficulty = self._get_curriculum_difficulty(epoch)
            metrics = self._train_epoch(difficulty) if self.is_deep_controller else self._evaluate_epoch()
            metrics.update(self.loss_logger.end_epoch())
            if epoch % self.config.log_interval == 0:
                self._log_epoch(epoch, metrics)
            if self.is_deep_controller:
                if (epoch + 1) % self.config.checkpoint_interval == 0:
                    self._save_checkpoint(epoch, metrics)
                current = metrics.get("total", float("inf"))
                if self.config.save_best and current < self.best_loss:
                    self.best_loss = metrics["total"]
                    self._save_checkpoint(epoch, metrics, is_best=True)
                if self._check_training_failure(metrics) and not self._recover_from_nan():
                    logger.error("Training failed, could not recover")
                    break
        if self.is_deep_controller:
            self._save_checkpoint(self.current_epoch, {}, is_final=True)
            self.controller.eval_mode()
        self._save_experiment_log()
        return self._get_training_summary()

    # -- logs and checkpoints ------------------------------------------------------------------------
    def _log_epoch(self, epoch: int, metrics: dict) -> None:
        logger.info("[%s] Epoch %4d | Loss: %.4f | Reward: %.2f | On-target: %.1f%% | Error: %.3fm",
                    self.config.controller.upper(), epoch, metrics.get("total", 0.0), metrics.get("mean_reward", 0.0),
                    100 * metrics.get("mean_on_target_ratio", 0.0), metrics.get("mean_tracking_error", 0.0))
        self.experiment_log.append({"epoch": epoch, "controller": self.config.controller,
                                    "timestamp": datetime.datetime.now(datetime.timezone.utc).isoformat(), **metrics})

    def _save_checkpoint(self, epoch: int, metrics: dict, is_best: bool = False, is_final: bool = False) -> None:
        name = f"{self.experiment_id}_{'best' if is_best else 'final' if is_final else f'epoch{epoch:04d}'}.pt"
        self.controller.save_checkpoint(self.checkpoint_dir / name, metadata={
            "epoch": epoch, "metrics": metrics, "config": self.config.to_dict(),
            "timestamp": datetime.datetime.now(datetime.timezone.utc).isoformat(), "experiment_id": self.experiment_id})

    def _save_config(self) -> None:
        with open(self.log_dir / f"{self.experiment_id}_config.json", "w") as f:
            json.dump(self.config.to_dict(), f, indent=2)

    def _save_experiment_log(self) -> None:
        with open(self.log_dir / f"{self.experiment_id}_log.json", "w") as f:
            json.dump(self.experiment_log, f, indent=2)
        if self.experiment_log:
            keys = list(self.experiment_log[0].keys())
            with open(self.log_dir / f"{self.experiment_id}_log.csv", "w", newline="") as f:
                w = csv.DictWriter(f, fieldnames=keys, extrasaction="ignore")
                w.writeheader()
                w.writerows(self.experiment_log)

    def _get_training_summary(self) -> dict:
        best_epoch, best_loss = self.loss_logger.get_best_epoch("total")
        summary = {"experiment_id": self.experiment_id, "controller": self.config.controller,
                   "epochs_completed": self.current_epoch + 1, "best_epoch": best_epoch, "best_loss": best_loss,
                   "final_loss": self.experiment_log[-1].get("total", float("inf")) if self.experiment_log else float("inf"),
                   "log_dir": str(self.log_dir)}
        if self.is_deep_controller:
            summary["nan_recoveries"] = self.nan_recovery_count
            summary["checkpoint_dir"] = str(self.checkpoint_dir)
        return summary


def load_checkpoint_and_resume(trainer: Trainer, checkpoint_path: str | Path) -> int:
    """train.py:1030-1045"""
    meta = trainer.controller.load_checkpoint(checkpoint_path)
    epoch = int(meta.get("epoch", -1)) + 1
    trainer.current_epoch = epoch
    return epoch
