"""Controller evaluation driver (drop-in for ``quadcopter_tracking.eval.Evaluator``).

The reference runs episodes one after another, one Python ``compute_action`` + ``env.step`` per
10 ms of simulated time (eval.py:95-231).  Here ``evaluate`` is three kernel launches for ALL
episodes: ``qt_reset_batched`` (initial conditions from the seeds), ``qt_rollout`` (the fused
closed loop + per-episode metrics) and ``qt_reduce_summary`` (compute_evaluation_summary).
Interfaces, seeds (base_seed + i), result types and the controller-state semantics follow the
reference; see ``reset_controller_between_episodes`` for the one place where the reference's
behaviour is inherently serial.
"""
from __future__ import annotations

import json
import logging
from collections.abc import Callable
from copy import deepcopy
from datetime import datetime, timezone
from pathlib import Path

import numpy as np

from . import _lib
from . import params as P
from .config import EnvConfig
from .controller import LQRController, PIDController, RiccatiLQRController
from .metrics import (EpisodeMetricsView, EvaluationSummary, SuccessCriteria, format_metrics_report,
                      summary_from_sums)

logger = logging.getLogger(__name__)

VALID_CONTROLLER_TYPES = ("deep", "lqr", "pid", "riccati_lqr", "lqi")  # controllers/__init__.py:71


def rollout_device(controller, env_config: EnvConfig, criteria, seeds, max_steps: int | None = None,
                   chain: int = 1, K=None, k_stride: int | None = None, record_stride: int = 0,
                   dataset_stride: int = 0, ctrl_state=None, want_final_state: bool = False, ctx=None,
                   supervisor=None):
    """Reset + closed-loop rollout for ``seeds`` (CUDA int64 tensor [E]) entirely on the device.

    Returns a dict of CUDA tensors: ``metrics`` [16, E], optionally ``final_state`` [12, E],
    ``traj`` [slots, 24, E] (every ``record_stride``-th step), ``features`` float32 [slots, E, 18] and
    ``actions`` float32 [slots, E, 4] (every ``dataset_stride``-th step), ``state0``, ``tparams``.
    ``K`` overrides the controller's own gain ([groups, 4, k_cols], group = episode // k_stride).
    A ``DeepTrackingPolicy`` controller acts through its network (``qt_rollout_policy``); with a
    ``supervisor`` (PID / LQR / Riccati controller object, reset before every episode as train.py:665-667
    does) ``actions`` holds the supervisor's labels for the states the policy visits.
    """
    import torch

    ctx = ctx or _lib.get_context()
    ctx.bind_torch_stream()
    dev = ctx.device
    E = int(seeds.numel())
    is_policy = getattr(controller, "name", "") == "deep_tracking"
    law = supervisor if is_policy else controller
    if law is not None:
        prm = law.kernel_params(env_config, criteria)
        k_cols = law.k_cols()
        if K is None:
            K = torch.as_tensor(np.ascontiguousarray(law.effective_gain()), dtype=torch.float64, device=dev)
            k_stride = E
        elif k_stride is None:
            k_stride = 1
    else:
        prm, k_cols, K, k_stride = controller.kernel_params(env_config, criteria), 6, None, E
    if max_steps is None:
        # the env ends the episode itself; this is only a safety cap on the loop length
        max_steps = int(np.ceil(env_config.simulation.max_episode_time / env_config.simulation.dt)) + 2
    f64 = dict(dtype=torch.float64, device=dev)
    state0 = torch.empty((12, E), **f64)
    tparams = torch.empty((12, E), **f64)
    ctx.reset_batched(prm, seeds, state0, tparams)
    out = {"state0": state0, "tparams": tparams, "metrics": torch.empty((16, E), **f64)}
    if want_final_state:
        out["final_state"] = torch.empty((12, E), **f64)
    traj_slots = data_slots = 0
    if record_stride > 0:
        traj_slots = (max_steps + record_stride - 1) // record_stride
        out["traj"] = torch.zeros((traj_slots, P.TRAJ_FIELDS, E), **f64)
    if dataset_stride > 0:
        data_slots = (max_steps + dataset_stride - 1) // dataset_stride
        out["features"] = torch.zeros((data_slots, E, P.FEATURE_DIM), dtype=torch.float32, device=dev)
        out["actions"] = torch.zeros((data_slots, E, P.ACTION_DIM), dtype=torch.float32, device=dev)
    if is_policy:
        if getattr(controller, "tensor_cores", False):
            prm.debug_flags |= P.FLAG_POLICY_TENSOR_CORES
        ctx.rollout_policy(prm, E, max_steps, state0, tparams, controller.packed_mlp(),
                           K.contiguous() if K is not None else None, k_cols, k_stride, out["metrics"],
                           out.get("final_state"), out.get("traj"), max(record_stride, 1), traj_slots,
                           out.get("features"), out.get("actions"), max(dataset_stride, 1), data_slots)
        return out
    ctx.rollout(prm, E, max_steps, chain, state0, tparams, K.contiguous(), k_cols, k_stride, out["metrics"],
                ctrl_state, out.get("final_state"), out.get("traj"), max(record_stride, 1), traj_slots,
                out.get("features"), out.get("actions"), max(dataset_stride, 1), data_slots)
    return out


class Evaluator:
    """eval.py:59-269 — same constructor, ``run_episode``, ``evaluate`` and ``save_report``."""

    def __init__(self, controller, env_config: EnvConfig | None = None,
                 criteria: SuccessCriteria | None = None, output_dir: str | Path = "reports",
                 reset_controller_between_episodes: bool = False):
        """``reset_controller_between_episodes``: the reference's ``evaluate`` never calls
        ``controller.reset()`` (eval.py:198-213), so in LQI mode the integrator state flows from
        episode i into episode i+1 and the episodes form one serial chain.  The default (False)
        keeps that behaviour exactly: for an LQI controller the whole evaluation runs as one chain
        in one GPU thread.  Set True for the tuner/trainer semantics (tuning.py:879: reset before
        every episode), which is what makes a million LQI episodes run in parallel.  Controllers
        without integral action have no state that influences actions; they always run in
        parallel."""
        self.controller = controller
        self.env_config = env_config or EnvConfig()
        self.criteria = criteria or SuccessCriteria()
        self.output_dir = Path(output_dir)
        self.output_dir.mkdir(parents=True, exist_ok=True)
        (self.output_dir / "plots").mkdir(exist_ok=True)
        self.reset_controller_between_episodes = reset_controller_between_episodes
        self.episode_data_list: list[list[dict]] = []
        self.episode_info_list: list[dict] = []
        self.last_metrics = None  # host [16, E] array of the most recent evaluate()

    # -- helpers ---------------------------------------------------------------------------------
    def _chain_for(self, n: int) -> int:
        carry = self.controller.carries_state() and not self.reset_controller_between_episodes
        return max(n, 2) if carry else 1  # chain > 1 <=> "do not reset, carry the state in"

    def _ctrl_state_in(self, n_chains: int):
        """Device [4, n_chains] controller state carried in (integrator + saturation count, or the
        PID integral + time of its previous call)."""
        import torch

        ctx = _lib.get_context()
        st = torch.zeros((4, n_chains), dtype=torch.float64, device=ctx.device)
        if n_chains == 1:
            st[:, 0] = torch.from_numpy(np.asarray(self.controller.state_vector(), dtype=np.float64))
        return st

    def _ctrl_state_out(self, st, chain: int) -> None:
        """Write the controller state back so the Python object ends where the reference's would:
        after one serial chain (chain > 1) in that chain's final state, otherwise (reset before every
        episode) in the last episode's state with the saturation counts added up."""
        host = st.cpu().numpy()
        if chain > 1:
            self.controller.load_state_vector(host[:, 0], True)
        else:
            last = host[:, -1].copy()
            if self.controller.name != "pid":
                last[3] = host[3].sum()
            self.controller.load_state_vector(last, False)

    @staticmethod
    def _episode_records(traj: np.ndarray, n_steps: int, target_radius: float) -> list[dict]:
        """[slots, 24] trajectory of one episode -> the per-step dictionaries of eval.py:142-159."""
        rows = []
        for k in range(n_steps):
            r = traj[k]
            err = float(r[P.TRAJ_ERR])
            rows.append({
                "time": float(r[P.TRAJ_TIME]),
                "step": k,
                "quadcopter_position": r[P.TRAJ_STATE:P.TRAJ_STATE + 3].tolist(),
                "quadcopter_velocity": r[P.TRAJ_STATE + 3:P.TRAJ_STATE + 6].tolist(),
                "target_position": r[P.TRAJ_TARGET_P:P.TRAJ_TARGET_P + 3].tolist(),
                "target_velocity": r[P.TRAJ_TARGET_V:P.TRAJ_TARGET_V + 3].tolist(),
                "action": r[P.TRAJ_ACTION:P.TRAJ_ACTION + 4].tolist(),
                "reward": -err,
                "tracking_error": err,
                "on_target": err <= target_radius,
            })
        return rows

    def _info_from_metrics(self, m: np.ndarray) -> dict:
        """Final ``info`` dictionary of an episode (quadcopter_env.py:209-226)."""
        n = int(m[P.M_STEPS])
        term = int(m[P.M_TERMINATION])
        info = {
            "time": float(m[P.M_FINAL_TIME]),
            "step": n,
            "tracking_error": float(m[P.M_FINAL_ERR]),
            "on_target": bool(m[P.M_FINAL_ERR] <= self.env_config.success_criteria.target_radius),
            "on_target_ratio": float(m[P.M_ENV_ON_TARGET]) / n if n > 0 else 0.0,
            "action_violations": int(m[P.M_VIOLATIONS]),
        }
        if term != 0:
            info["termination_reason"] = P.TERMINATION_REASONS[term]
            info["episode_length"] = info["time"]
            info["success"] = bool(m[P.M_ENV_SUCCESS] != 0)
        return info

    # -- reference interface -----------------------------------------------------------------------
    def run_episode(self, seed: int, max_steps: int | None = None,
                    progress_callback: Callable | None = None) -> tuple[list[dict], dict]:
        """eval.py:95-167.  One episode with the full per-step record."""
        import torch

        ctx = _lib.get_context()
        seeds = torch.tensor([int(seed)], dtype=torch.int64, device=ctx.device)
        chain = self._chain_for(1)
        st = self._ctrl_state_in(1)
        out = rollout_device(self.controller, self.env_config, self.criteria, seeds, max_steps,
                             chain=chain, record_stride=1, ctrl_state=st, ctx=ctx)
        m = out["metrics"].cpu().numpy()[:, 0]
        self._ctrl_state_out(st, chain)
        n = int(m[P.M_STEPS])
        traj = out["traj"].cpu().numpy()[:, :, 0]
        data = self._episode_records(traj, n, self.env_config.success_criteria.target_radius)
        info = self._info_from_metrics(m)
        if progress_callback:
            for k in range(100, n + 1, 100):
                progress_callback(k, info)
        return data, info

    def evaluate(self, num_episodes: int = 10, base_seed: int = 42,
                 max_steps_per_episode: int | None = None, verbose: bool = True,
                 record_trajectories: bool | None = None) -> EvaluationSummary:
        """eval.py:169-231.  ``record_trajectories`` (default: only when num_episodes <= 32) fills
        ``episode_data_list`` with the per-step records; per-episode metrics are always produced."""
        import torch

        ctx = _lib.get_context()
        dev = ctx.device
        E = int(num_episodes)
        self.episode_data_list, self.episode_info_list = [], []
        if E <= 0:
            return EvaluationSummary()
        if record_trajectories is None:
            record_trajectories = E <= 32
        seeds = torch.arange(E, dtype=torch.int64, device=dev) + int(base_seed)
        chain = self._chain_for(E)
        n_chains = (E + chain - 1) // chain
        st = self._ctrl_state_in(n_chains)
        out = rollout_device(self.controller, self.env_config, self.criteria, seeds, max_steps_per_episode,
                             chain=chain, record_stride=1 if record_trajectories else 0, ctrl_state=st, ctx=ctx)
        sums = torch.empty(16, dtype=torch.float64, device=dev)
        ctx.reduce_summary(out["metrics"], E, sums)
        metrics_host = out["metrics"].cpu().numpy()
        sums_host = sums.cpu().numpy()
        self._ctrl_state_out(st, chain)
        self.last_metrics = metrics_host
        self.episode_info_list = _LazyInfos(self, metrics_host)
        if record_trajectories:
            traj = out["traj"].cpu().numpy()
            radius = self.env_config.success_criteria.target_radius
            self.episode_data_list = [
                self._episode_records(traj[:, :, i], int(metrics_host[P.M_STEPS, i]), radius) for i in range(E)]
        summary = summary_from_sums(sums_host, self.criteria, EpisodeMetricsView(metrics_host))
        if verbose:
            print("\n" + format_metrics_report(summary))
        return summary

    def save_report(self, summary: EvaluationSummary, experiment_name: str | None = None) -> dict[str, Path]:
        """eval.py:233-268 (metrics.json + text report)."""
        if experiment_name is None:
            stamp = datetime.now(timezone.utc).strftime("%Y%m%d_%H%M%S")
            experiment_name = f"eval_{self.controller.name}_{stamp}"
        saved = {}
        metrics_path = self.output_dir / "metrics.json"
        with open(metrics_path, "w") as f:
            json.dump(summary.to_dict(), f, indent=2)
        saved["metrics"] = metrics_path
        report_path = self.output_dir / f"{experiment_name}_report.txt"
        with open(report_path, "w") as f:
            f.write(format_metrics_report(summary))
        saved["report"] = report_path
        return saved


    # -- plots (eval.py:270-469; drawn by the package's own renderer, plots.py — matplotlib is not required) -------
    def _episode_arrays(self, episode_idx: int):
        n = len(self.episode_data_list)
        if episode_idx >= n:
            raise ValueError(f"Episode {episode_idx} not found (have {n} episodes)")
        data = self.episode_data_list[episode_idx]
        times = np.array([d["time"] for d in data], dtype=np.float64)
        return data, times

    def plot_trajectory(self, episode_idx: int = 0, save_path: str | Path | None = None, show_radius: bool = True):
        """eval.py:270-340: x / y / z of vehicle and target over time (three stacked charts sharing the time axis),
        with the target-radius band.  Returns (figure, axes)."""
        from . import plots

        data, times = self._episode_arrays(episode_idx)
        quad = np.array([d["quadcopter_position"] for d in data], dtype=np.float64).reshape(-1, 3)
        target = np.array([d["target_position"] for d in data], dtype=np.float64).reshape(-1, 3)
        radius = self.criteria.target_radius
        fig, axes = plots.subplots(3, 1, figsize=(12, 10), sharex=True)
        for i, (ax, name, color) in enumerate(zip(axes, "XYZ", ("tab:blue", "tab:orange", "tab:green"))):
            ax.plot(times, quad[:, i], label="Quadcopter", color=color, linewidth=1)
            ax.plot(times, target[:, i], label="Target", color="red", linestyle="--", linewidth=1)
            if show_radius:
                ax.fill_between(times, target[:, i] - radius, target[:, i] + radius, alpha=0.2, color="green",
                                label="Target Radius" if i == 0 else "")
            ax.set_ylabel(f"{name} Position (m)")
            ax.legend(loc="upper right")
            ax.grid(True, alpha=0.3)
        axes[-1].set_xlabel("Time (s)")
        fig.suptitle(f"Position Tracking - Episode {episode_idx + 1}\nController: {self.controller.name}", fontsize=12)
        if save_path:
            fig.savefig(save_path, dpi=150)
            logger.info("Saved trajectory plot: %s", save_path)
        return fig, axes

    def plot_tracking_error(self, episode_idx: int = 0, save_path: str | Path | None = None):
        """eval.py:342-404: tracking error over time, the on-target samples marked, the target radius as a dashed
        line.  Returns (figure, axes)."""
        from . import plots

        data, times = self._episode_arrays(episode_idx)
        errors = np.array([d["tracking_error"] for d in data], dtype=np.float64)
        on_target = np.array([d["on_target"] for d in data], dtype=bool)
        radius = self.criteria.target_radius
        fig, ax = plots.subplots(figsize=(12, 5))
        ax.plot(times, errors, "b-", linewidth=1, label="Tracking Error")
        ax.scatter(times[on_target], errors[on_target], c="green", s=2, alpha=0.5, label="On-Target")
        ax.axhline(y=radius, color="r", linestyle="--", label=f"Target Radius ({radius}m)")
        ax.set_xlabel("Time (s)")
        ax.set_ylabel("Tracking Error (m)")
        ax.set_title(f"Tracking Error - Episode {episode_idx + 1} | Controller: {self.controller.name}")
        ax.legend()
        ax.grid(True, alpha=0.3)
        ax.set_ylim(bottom=0)
        if save_path:
            fig.savefig(save_path, dpi=150)
            logger.info("Saved tracking error plot: %s", save_path)
        return fig, ax

    def generate_all_plots(self, summary: EvaluationSummary | None = None,
                           experiment_name: str | None = None) -> list[Path]:
        """eval.py:406-469: position-tracking and tracking-error plots of the best episode and, when it is another
        one, of the worst episode, under ``<output_dir>/plots``.  Needs the per-step records
        (``evaluate(record_trajectories=True)``, the default up to 32 episodes); returns the paths written."""
        plots_dir = self.output_dir / "plots"
        saved: list[Path] = []
        n = len(self.episode_data_list)
        if n == 0:
            return saved
        best = min(summary.best_episode_idx, n - 1) if summary is not None else 0
        worst = min(summary.worst_episode_idx, n - 1) if summary is not None else 0
        for tag, idx in (("best", best), ("worst", worst)):
            if tag == "worst" and worst == best:
                break
            path = plots_dir / f"position_tracking_{tag}.png"
            self.plot_trajectory(episode_idx=idx, save_path=path)
            saved.append(path)
            path = plots_dir / f"tracking_error_{tag}.png"
            self.plot_tracking_error(episode_idx=idx, save_path=path)
            saved.append(path)
        return saved


class _LazyInfos:
    """Sequence of final ``info`` dictionaries, built on access."""

    def __init__(self, evaluator: Evaluator, metrics_host: np.ndarray):
        self._ev, self._m = evaluator, metrics_host

    def __len__(self) -> int:
        return int(self._m.shape[1])

    def __getitem__(self, i: int) -> dict:
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return self._ev._info_from_metrics(self._m[:, i])

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def load_controller(controller_type: str, checkpoint_path: str | Path | None = None, config: dict | None = None):
    """eval.py:472-513 for the controller types on the accelerated path."""
    config = config or {}
    if controller_type == "riccati_lqr":
        return RiccatiLQRController(config=config)
    if controller_type == "lqi":
        lqi_config = deepcopy(config)
        lqi_config["use_lqi"] = True
        if "q_int" not in lqi_config:
            lqi_config["q_int"] = [0.01, 0.01, 0.1]
        return RiccatiLQRController(config=lqi_config)
    if controller_type == "lqr":
        return LQRController(config=config)
    if controller_type == "pid":
        return PIDController(config=config)
    if controller_type == "deep":
        from .policy import DeepTrackingPolicy

        controller = DeepTrackingPolicy(config=config)
        if checkpoint_path:
            controller.load_checkpoint(checkpoint_path)
        return controller
    raise ValueError(f"Unknown controller type: {controller_type}")


def run_hyperparameter_sweep(sweep_config_path: str | Path, output_dir: str | Path = "reports/sweeps") -> list[dict]:
    """eval.py:516-628: evaluate every entry of a sweep YAML/JSON (``configurations: [{name, controller_type,
    controller_config, checkpoint_path, env_config, criteria, num_episodes, seed}, ...]``) and rank the
    results by mean on-target ratio.  Each configuration is one fused launch over its episodes."""
    path = Path(sweep_config_path)
    with open(path) as f:
        if path.suffix == ".json":
            sweep_config = json.load(f)
        else:
            import yaml

            sweep_config = yaml.safe_load(f)
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    results = []
    for i, config in enumerate((sweep_config or {}).get("configurations", [])):
        name = config.get("name", f"config_{i}")
        try:
            controller = load_controller(controller_type=config.get("controller_type", "deep"),
                                         checkpoint_path=config.get("checkpoint_path"),
                                         config=config.get("controller_config"))
            env_config = EnvConfig.from_dict(config["env_config"]) if "env_config" in config else EnvConfig()
            crit = config.get("criteria", {})
            criteria = SuccessCriteria(min_on_target_ratio=crit.get("min_on_target_ratio", 0.8),
                                       min_episode_duration=crit.get("min_episode_duration", 30.0),
                                       target_radius=crit.get("target_radius", 0.5))
            evaluator = Evaluator(controller=controller, env_config=env_config, criteria=criteria,
                                  output_dir=output_dir / name)
            summary = evaluator.evaluate(num_episodes=config.get("num_episodes", 5), base_seed=config.get("seed", 42),
                                         verbose=False)
            results.append({"name": name, "config": config, "mean_on_target_ratio": summary.mean_on_target_ratio,
                            "success_rate": summary.success_rate, "mean_tracking_error": summary.mean_tracking_error,
                            "meets_criteria": summary.meets_criteria})
        except Exception as e:  # the reference records the failure and carries on (eval.py:590-598)
            logger.error("Configuration %s failed: %s", name, e)
            results.append({"name": name, "config": config, "error": str(e)})
    valid = [r for r in results if "error" not in r]
    valid.sort(key=lambda r: r["mean_on_target_ratio"], reverse=True)
    with open(output_dir / "sweep_results.json", "w") as f:
        json.dump(valid, f, indent=2)
    return valid
