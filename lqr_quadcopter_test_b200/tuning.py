"""Controller gain auto-tuning on the GPU (drop-in for ``quadcopter_tracking.controllers.tuning``).

The reference evaluates one candidate at a time: build a controller (one scipy DARE), run
``evaluation_episodes`` episodes step by step, average (tuning.py:846-928).  Here a whole batch of
candidates is one ``qt_dare_batched`` launch (all Riccati solutions), one ``qt_rollout`` launch
(candidates x episodes threads, the 64 episodes of a candidate sharing its gain through shared
memory) and one ``qt_reduce_scores`` launch.  Candidate *generation* stays on the host and
consumes ``np.random.default_rng(seed)`` in exactly the reference's order (tuning.py:683-735), so
candidate i here is candidate i there.
"""
from __future__ import annotations

import json
import logging
import os
import pickle
from dataclasses import dataclass, field, fields
from datetime import datetime, timezone
from itertools import product
from pathlib import Path

import numpy as np

from . import _lib
from . import params as P
from .config import EnvConfig
from .controller import LQRController, PIDController, RiccatiLQRController, heuristic_lqr_gain
from .metrics import SuccessCriteria

logger = logging.getLogger(__name__)

_VEC3_RANGES = ("kp_pos_range", "ki_pos_range", "kd_pos_range", "ff_velocity_gain_range",
                "ff_acceleration_gain_range", "q_pos_range", "q_vel_range")
_SCALAR_RANGES = ("r_thrust_range", "r_rate_range")


@dataclass
class GainSearchSpace:
    """tuning.py:146-360.  Each ``*_range`` is ``(min_values, max_values)``; None = not searched."""

    kp_pos_range: tuple | None = None
    ki_pos_range: tuple | None = None
    kd_pos_range: tuple | None = None
    ff_velocity_gain_range: tuple | None = None
    ff_acceleration_gain_range: tuple | None = None
    q_pos_range: tuple | None = None
    q_vel_range: tuple | None = None
    r_thrust_range: tuple | None = None
    r_rate_range: tuple | None = None
    r_controls_range: tuple | None = None
    q_int_range: tuple | None = None
    use_lqi: bool = False

    @staticmethod
    def _check_vector(name: str, rng, size: int, labels, what: str) -> None:
        lo_v, hi_v = rng
        if len(lo_v) != size or len(hi_v) != size:
            raise ValueError(f"{name} must have exactly {size} values for [{', '.join(labels)}], "
                             f"got min={len(lo_v)}, max={len(hi_v)}")
        for lab, lo, hi in zip(labels, lo_v, hi_v):
            if lo > hi:
                raise ValueError(f"{name}[{lab}] has inverted range: min={lo} > max={hi}. "
                                 f"Swap values or use equal values for fixed parameter.")
            if lo < 0:
                raise ValueError(f"{name}[{lab}] has negative minimum: {lo}. {what}")

    def validate(self) -> None:
        for name in _VEC3_RANGES:
            rng = getattr(self, name)
            if rng is not None:
                self._check_vector(name, rng, 3, ("x", "y", "z"), "Controller gains should be non-negative.")
        for name in _SCALAR_RANGES:
            rng = getattr(self, name)
            if rng is None:
                continue
            lo, hi = rng
            if lo > hi:
                raise ValueError(f"{name} has inverted range: min={lo} > max={hi}. "
                                 f"Swap values or use equal values for fixed parameter.")
            if lo < 0:
                raise ValueError(f"{name} has negative minimum: {lo}. Cost weights should be non-negative.")
        if self.r_controls_range is not None:
            self._check_vector("r_controls_range", self.r_controls_range, 4,
                               ("thrust", "roll", "pitch", "yaw"), "Cost weights should be non-negative.")
        if self.q_int_range is not None:
            self._check_vector("q_int_range", self.q_int_range, 3, ("ix", "iy", "iz"),
                               "Integral cost weights must be non-negative.")

    def get_active_parameters(self) -> list[str]:
        return [f.name[: -len("_range")] for f in fields(self)
                if f.name.endswith("_range") and getattr(self, f.name) is not None]

    @classmethod
    def from_dict(cls, config: dict) -> "GainSearchSpace":
        kw = {f.name: config.get(f.name) for f in fields(cls) if f.name != "use_lqi"}
        return cls(use_lqi=config.get("use_lqi", False), **kw)

    def to_dict(self) -> dict:
        return {f.name: getattr(self, f.name) for f in fields(self)}


@dataclass
class TuningConfig:
    """tuning.py:363-478"""

    controller_type: str = "pid"
    search_space: GainSearchSpace = field(default_factory=GainSearchSpace)
    strategy: str = "random"
    max_iterations: int = 50
    grid_points_per_dim: int = 3
    evaluation_episodes: int = 5
    evaluation_horizon: int = 3000
    seed: int = 42
    target_motion_type: str = "stationary"
    episode_length: float = 30.0
    target_radius: float = 0.5
    output_dir: str = "reports/tuning"
    resume_from: str | None = None
    feedforward_enabled: bool = False
    cma_sigma0: float = 0.3
    cma_popsize: int | None = None

    def __post_init__(self) -> None:
        if self.controller_type not in ("pid", "lqr", "riccati_lqr"):
            raise ValueError(f"Invalid controller_type: '{self.controller_type}'. "
                             f"Valid choices are: pid, lqr, riccati_lqr")
        if self.strategy not in ("grid", "random", "cma_es"):
            raise ValueError(f"Invalid strategy: '{self.strategy}'. Valid choices are: grid, random, cma_es")
        if self.max_iterations < 1:
            raise ValueError(f"max_iterations must be >= 1, got {self.max_iterations}")
        if self.evaluation_episodes < 1:
            raise ValueError(f"evaluation_episodes must be >= 1, got {self.evaluation_episodes}")
        if self.cma_sigma0 <= 0:
            raise ValueError(f"cma_sigma0 must be > 0, got {self.cma_sigma0}")
        if self.cma_popsize is not None and self.cma_popsize < 2:
            raise ValueError(f"cma_popsize must be >= 2, got {self.cma_popsize}")

    @classmethod
    def from_dict(cls, config: dict) -> "TuningConfig":
        kw = {f.name: config[f.name] for f in fields(cls) if f.name != "search_space" and f.name in config}
        return cls(search_space=GainSearchSpace.from_dict(config.get("search_space", {})), **kw)

    def to_dict(self) -> dict:
        d = {f.name: getattr(self, f.name) for f in fields(self) if f.name != "search_space"}
        d["search_space"] = self.search_space.to_dict()
        return d


@dataclass
class TuningResult:
    """tuning.py:481-578"""

    best_config: dict
    best_score: float
    best_metrics: dict
    all_results: list
    iterations_completed: int
    interrupted: bool
    timestamp: str
    config: dict

    def to_dict(self) -> dict:
        return {f.name: getattr(self, f.name) for f in fields(self)}

    @classmethod
    def from_dict(cls, data: dict) -> "TuningResult":
        return cls(**{f.name: data[f.name] for f in fields(cls)})

    def save(self, path) -> Path:
        path = Path(path)
        _validate_path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        with open(path, "w") as f:
            json.dump(self.to_dict(), f, indent=2)
        return path

    @classmethod
    def load(cls, path) -> "TuningResult":
        path = Path(path)
        _validate_path(path)
        if not path.exists():
            raise FileNotFoundError(f"Tuning results file not found: {path}")
        if not path.is_file():
            raise ValueError(f"Path is not a file: {path}")
        with open(path) as f:
            return cls.from_dict(json.load(f))


def _validate_path(path: Path) -> None:
    """tuning.py:64-106: refuse traversal sequences and NUL bytes."""
    text = str(path)
    if ".." in text:
        raise ValueError(f"Path contains path traversal sequence '..': {path}. "
                         "Use absolute paths or paths without parent directory references.")
    if "\x00" in text:
        raise ValueError(f"Path contains null byte: {path}")


_Q_POS_DEFAULT, _Q_VEL_DEFAULT = [0.0001, 0.0001, 16.0], [0.0036, 0.0036, 4.0]   # riccati_lqr.py:624-625
_Q_INT_DEFAULT, _R_DEFAULT = [0.0, 0.0, 0.0], [1.0, 1.0, 1.0, 1.0]

# the order in which _generate_random_config draws its parameters (tuning.py:683-735) and their widths
_PARAM_ORDER = (("kp_pos", 3), ("ki_pos", 3), ("kd_pos", 3), ("ff_velocity_gain", 3), ("ff_acceleration_gain", 3),
                ("q_pos", 3), ("q_vel", 3), ("r_thrust", 1), ("r_rate", 1), ("r_controls", 4), ("q_int", 3))


@dataclass
class CandidateBlock:
    """A batch of tuner candidates as ONE array instead of a list of dictionaries: ``values[i]`` holds candidate i's
    parameters in the reference's draw order, ``spec`` names the columns.  ``config(i)`` rebuilds exactly the
    dictionary ``_generate_random_config`` would have returned for that candidate."""

    spec: list            # [(name, dim), ...] in column order
    values: np.ndarray    # float64 [C, sum(dim)]
    use_lqi: bool = False
    zero_q_int: bool = False   # use_lqi without a q_int range: q_int = [0, 0, 0] (tuning.py:727-733)

    def __len__(self) -> int:
        return int(self.values.shape[0])

    def column(self, name: str):
        """float64 [C, dim] view of one parameter, or None if it is not part of the search."""
        at = 0
        for key, dim in self.spec:
            if key == name:
                return self.values[:, at:at + dim]
            at += dim
        return None

    @property
    def feedforward_enabled(self) -> bool:
        return any(k in ("ff_velocity_gain", "ff_acceleration_gain") for k, _ in self.spec)

    def config(self, i: int) -> dict:
        row, cfg, at = self.values[i], {}, 0
        for key, dim in self.spec:
            cfg[key] = float(row[at]) if dim == 1 else [float(v) for v in row[at:at + dim]]
            if key in ("ff_velocity_gain", "ff_acceleration_gain"):
                cfg["feedforward_enabled"] = True
            if key == "q_int":
                cfg["use_lqi"] = True
            at += dim
        if self.zero_q_int:
            cfg["use_lqi"] = True
            cfg["q_int"] = [0.0, 0.0, 0.0]
        return cfg

    def configs(self) -> list[dict]:
        return [self.config(i) for i in range(len(self))]

    def slice(self, lo: int, hi: int) -> "CandidateBlock":
        return CandidateBlock(self.spec, self.values[lo:hi], self.use_lqi, self.zero_q_int)

    @classmethod
    def from_configs(cls, configs: list[dict]) -> "CandidateBlock | None":
        """Pack dictionaries that all carry the same keys (what the generators produce) into one array; None if they
        differ in their key sets or in a non-searched entry (those batches are evaluated one by one)."""
        first = configs[0]
        keys = [(k, d) for k, d in _PARAM_ORDER if k in first]
        searched = {k for k, _ in keys}
        fixed = {k: first[k] for k in first if k not in searched}
        for c in configs:
            if c.keys() != first.keys():
                return None
            for k, v in fixed.items():
                same = c[k] == v
                if not (bool(np.all(same)) if not isinstance(same, bool) else same):
                    return None
        try:
            cols = [np.asarray([c[k] for c in configs], dtype=np.float64).reshape(len(configs), d) for k, d in keys]
        except (ValueError, TypeError):
            return None
        values = np.concatenate(cols, axis=1) if cols else np.zeros((len(configs), 0))
        return cls(keys, np.ascontiguousarray(values), bool(first.get("use_lqi", False)), False)


class ControllerTuner:
    """tuning.py:581-1442 for the Riccati-LQR / LQI (and heuristic LQR) controllers."""

    # candidates evaluated per launch; 65 536 x 64 episodes fits comfortably in 180 GB of HBM
    max_batch_candidates = 65536

    def __init__(self, config: TuningConfig):
        self.config = config
        self.config.search_space.validate()
        self.rng = np.random.default_rng(config.seed)
        self.results: list[dict] = []
        self.best_config: dict = {}
        self.best_score: float = float("-inf")
        self.best_metrics: dict = {}
        self.interrupted: bool = False
        out = Path(os.environ.get("TUNING_OUTPUT_DIR", config.output_dir))
        _validate_path(out)
        self.output_dir = out
        # strategy "cma_es": the reference needs the optional third-party `cma` package (tuning.py:636-643);
        # here the ask/tell loop is cmaes.py, so the strategy is always available
        self.cma_es = None
        self._cma_checkpoint_path: Path | None = None
        self._cma_param_spec: list[tuple[str, int, tuple]] = []

    # -- candidate generation: host, identical RNG consumption to tuning.py:662-735 -----------------
    def _active_spec(self) -> list[tuple[str, int, np.ndarray, np.ndarray]]:
        sp, out = self.config.search_space, []
        for key, dim in _PARAM_ORDER:
            rng_def = getattr(sp, key + "_range")
            if rng_def is not None:
                lo, hi = (np.asarray(v, dtype=np.float64).reshape(dim) for v in rng_def)
                out.append((key, dim, lo, hi))
        return out

    def sample_candidates(self, count: int, lo: int = 0, hi: int | None = None) -> CandidateBlock:
        """``count`` random candidates in one call, consuming ``self.rng`` exactly like ``count`` calls of
        ``_generate_random_config`` (tuning.py:683-735): every parameter there is one scalar ``rng.uniform(lo, hi)``,
        i.e. ``lo + (hi - lo) * next_double`` with one generator step per draw, taken in a fixed order — so one
        ``rng.uniform(lo_row, hi_row, size=(count, k))`` (numpy fills it row by row with the same C routine) yields
        the same bits, 65 536 x 14 draws in ~10 ms instead of 1 s of Python calls.

        ``lo, hi`` select candidates [lo, hi) of the ``count`` (a rank's shard): the generator is advanced past the
        others (PCG64.advance, one step per draw), so the rows — and the generator's final state — are what the full
        draw would have given."""
        hi = count if hi is None else hi
        spec = self._active_spec()
        k = sum(d for _, d, _, _ in spec)
        sp = self.config.search_space
        zero_q_int = bool(sp.use_lqi and sp.q_int_range is None)
        use_lqi = bool(sp.use_lqi or sp.q_int_range is not None)
        if k == 0:
            return CandidateBlock([], np.zeros((hi - lo, 0)), use_lqi, zero_q_int)
        lo_row = np.concatenate([l for _, _, l, _ in spec])
        hi_row = np.concatenate([h for _, _, _, h in spec])
        bg = self.rng.bit_generator
        if lo > 0:
            bg.advance(lo * k)
        values = self.rng.uniform(lo_row, hi_row, size=(hi - lo, k))
        if hi < count:
            bg.advance((count - hi) * k)
        return CandidateBlock([(n, d) for n, d, _, _ in spec], values, use_lqi, zero_q_int)

    def _draw(self, rng_def):
        lo_v, hi_v = rng_def
        return [float(self.rng.uniform(lo, hi)) for lo, hi in zip(lo_v, hi_v)]

    def _generate_random_config(self) -> dict:
        sp = self.config.search_space
        cfg: dict = {}
        for key in ("kp_pos", "ki_pos", "kd_pos"):
            if getattr(sp, key + "_range") is not None:
                cfg[key] = self._draw(getattr(sp, key + "_range"))
        for key in ("ff_velocity_gain", "ff_acceleration_gain"):
            if getattr(sp, key + "_range") is not None:
                cfg[key] = self._draw(getattr(sp, key + "_range"))
                cfg["feedforward_enabled"] = True
        for key in ("q_pos", "q_vel"):
            if getattr(sp, key + "_range") is not None:
                cfg[key] = self._draw(getattr(sp, key + "_range"))
        for key in ("r_thrust", "r_rate"):
            rng_def = getattr(sp, key + "_range")
            if rng_def is not None:
                cfg[key] = float(self.rng.uniform(rng_def[0], rng_def[1]))
        if sp.r_controls_range is not None:
            cfg["r_controls"] = self._draw(sp.r_controls_range)
        if sp.q_int_range is not None:
            cfg["q_int"] = self._draw(sp.q_int_range)
            cfg["use_lqi"] = True
        elif sp.use_lqi:
            cfg["use_lqi"] = True
            cfg["q_int"] = [0.0, 0.0, 0.0]
        return cfg

    def _generate_grid_configs(self) -> list[dict]:
        """tuning.py:737-830"""
        sp, n = self.config.search_space, self.config.grid_points_per_dim

        def axis(lo, hi):
            return [lo] if np.isclose(lo, hi) else list(np.linspace(lo, hi, n))

        grids: dict[str, list] = {}
        for key in ("kp_pos", "ki_pos", "kd_pos", "ff_velocity_gain", "ff_acceleration_gain", "q_pos", "q_vel"):
            rng_def = getattr(sp, key + "_range")
            if rng_def is not None:
                grids[key] = [list(c) for c in product(*[axis(lo, hi) for lo, hi in zip(*rng_def)])]
        for key in ("r_thrust", "r_rate"):
            rng_def = getattr(sp, key + "_range")
            if rng_def is not None:
                grids[key] = axis(rng_def[0], rng_def[1])
        if sp.r_controls_range is not None:
            grids["r_controls"] = [list(c) for c in product(*[axis(lo, hi) for lo, hi in zip(*sp.r_controls_range)])]
        if sp.q_int_range is not None:
            grids["q_int"] = [list(c) for c in product(*[axis(lo, hi) for lo, hi in zip(*sp.q_int_range)])]
        elif sp.use_lqi:
            grids["q_int"] = [[0.0, 0.0, 0.0]]
        if not grids:
            return []
        names = list(grids)
        configs = []
        for combo in product(*[grids[k] for k in names]):
            cfg = dict(zip(names, combo))
            if "ff_velocity_gain" in cfg or "ff_acceleration_gain" in cfg:
                cfg["feedforward_enabled"] = True
            if "q_int" in cfg or sp.use_lqi:
                cfg["use_lqi"] = True
            configs.append(cfg)
        return configs

    # -- evaluation ----------------------------------------------------------------------------------
    def _create_controller(self, controller_config: dict):
        if self.config.controller_type == "pid":
            return PIDController(config=controller_config)
        if self.config.controller_type == "lqr":
            return LQRController(config=controller_config)
        if self.config.controller_type == "riccati_lqr":
            return RiccatiLQRController(config=controller_config)
        raise ValueError(f"Unknown controller type: {self.config.controller_type}")

    def _env_and_criteria(self):
        env_config = EnvConfig()
        env_config.simulation.max_episode_time = self.config.episode_length
        env_config.target.motion_type = self.config.target_motion_type
        env_config.success_criteria.target_radius = self.config.target_radius
        criteria = SuccessCriteria(min_on_target_ratio=0.8, min_episode_duration=self.config.episode_length,
                                   target_radius=self.config.target_radius)
        return env_config, criteria

    def _evaluate_config(self, controller_config: dict) -> tuple[float, dict]:
        """tuning.py:846-928 for one candidate (one controller object, hence one DARE)."""
        scores, metrics = self._rows_to_lists(np.asarray(self._evaluate_one(controller_config)).reshape(4, 1))
        return scores[0], metrics[0]

    @staticmethod
    def _batchable(configs: list[dict]) -> bool:
        """A batch shares one qt_params block: LQI mode and the feedforward switch must agree (the
        feedforward gains themselves travel per candidate, qt_rollout's ff_gains)."""
        lqi = {bool(c.get("use_lqi", False)) for c in configs}
        ff = {bool(c.get("feedforward_enabled", False)) for c in configs}
        return len(lqi) == 1 and len(ff) == 1

    def evaluate_configs(self, configs: list[dict]) -> tuple[list[float], list[dict]]:
        """Score every candidate configuration: returns (scores, metrics dictionaries)."""
        if not configs:
            return [], []
        block = None
        if self._batchable(configs) and not any("Q" in c or "R" in c or "K" in c for c in configs):
            block = CandidateBlock.from_configs(configs)
        if block is None:
            parts = [self._evaluate_one(c) for c in configs]
            rows = np.array([p for p in parts], dtype=np.float64).T.reshape(4, len(configs))
        else:
            rows = self.evaluate_block(block, configs)
        return self._rows_to_lists(rows)

    def _rows_to_lists(self, rows: np.ndarray) -> tuple[list[float], list[dict]]:
        eps = int(self.config.evaluation_episodes)
        out = [{"mean_on_target_ratio": a, "mean_tracking_error": b, "success_rate": c_, "episodes_evaluated": eps}
               for a, b, c_ in zip(rows[1].tolist(), rows[2].tolist(), rows[3].tolist())]
        return rows[0].tolist(), out

    def evaluate_block(self, block: CandidateBlock, configs: list[dict] | None = None) -> np.ndarray:
        """Array-native candidate evaluation: float64 [4, C] = score, mean_on_target_ratio, mean_tracking_error,
        success_rate per candidate (tuning.py:912-926).  No per-candidate Python objects are created."""
        C = len(block)
        rows = np.empty((4, C))
        for i in range(0, C, self.max_batch_candidates):
            j = min(C, i + self.max_batch_candidates)
            rows[:, i:j] = self._evaluate_block_launch(block.slice(i, j), None if configs is None else configs[i:j])
        return rows

    def _evaluate_one(self, controller_config: dict) -> tuple[float, float, float, float]:
        """One candidate through its own controller object (tuning.py:846-928 literally): the constructor solves
        the DARE and applies the reference's rules — ValueError for malformed weights, heuristic-LQR fallback
        WITHOUT feedforward and without integral action when the solve fails (riccati_lqr.py:737-777)."""
        import torch

        ctx = _lib.get_context()
        ctx.bind_torch_stream()
        cfg = self.config
        env_config, criteria = self._env_and_criteria()
        ctrl = self._create_controller(controller_config)
        prm = ctrl.kernel_params(env_config, criteria)
        f64 = dict(dtype=torch.float64, device=ctx.device)
        K = torch.as_tensor(np.ascontiguousarray(ctrl.effective_gain()), **f64).reshape(1, 4, ctrl.k_cols())
        ff = None
        if prm.feedforward_enabled:
            ff = torch.as_tensor(np.concatenate([np.asarray(ctrl.ff_velocity_gain, dtype=np.float64).reshape(3),
                                                 np.asarray(ctrl.ff_acceleration_gain, dtype=np.float64).reshape(3)]),
                                 **f64).reshape(1, 6)
        cand = self._sweep_device(ctx, prm, K, ctrl.k_cols(), ff)
        return tuple(cand.cpu().numpy()[:, 0].tolist())

    def _sweep_device(self, ctx, prm, K, k_cols: int, ff_gains, want_metrics: bool = False):
        """Reset the evaluation episodes once (every candidate sees the same seeds seed + ep, tuning.py:875) and run
        all candidates x episodes in one qt_rollout_sweep launch; returns the device [4, C] score rows."""
        import torch

        cfg = self.config
        C, eps = int(K.shape[0]), int(cfg.evaluation_episodes)
        f64 = dict(dtype=torch.float64, device=ctx.device)
        seeds = torch.empty(eps, dtype=torch.int64, device=ctx.device)
        ctx.fill_seeds(seeds, int(cfg.seed))
        s0, tp = torch.empty((12, eps), **f64), torch.empty((12, eps), **f64)
        ctx.reset_batched(prm, seeds, s0, tp)
        cand = torch.empty((4, C), **f64)
        metrics = torch.empty((16, C * eps), **f64) if want_metrics else None
        ctx.rollout_sweep(prm, C, eps, int(cfg.evaluation_horizon), s0, tp, K.contiguous(), k_cols, cand, metrics, ff_gains)
        self.last_metrics = metrics
        return cand

    keep_episode_metrics = False   # True: evaluate_* also leave the [16, C*eps] per-episode rows in last_metrics

    def _evaluate_block_launch(self, block: CandidateBlock, configs: list[dict] | None) -> np.ndarray:
        import torch

        ctx = _lib.get_context()
        ctx.bind_torch_stream()
        C = len(block)
        env_config, criteria = self._env_and_criteria()
        f64 = dict(dtype=torch.float64, device=ctx.device)
        proto_cfg = configs[0] if configs is not None else block.config(0)
        retry: list[int] = []
        if self.config.controller_type != "riccati_lqr":
            # PID / heuristic LQR: closed-form gains, one small host computation per candidate
            proto = self._create_controller(proto_cfg)
            cfgs = configs if configs is not None else block.configs()
            ks = np.stack([np.ascontiguousarray(self._create_controller(c).effective_gain()) for c in cfgs])
            K = torch.as_tensor(ks, **f64)
        else:
            lqi = block.use_lqi
            n = 9 if lqi else 6
            q = np.empty((C, n))
            for name, at, default in (("q_pos", 0, _Q_POS_DEFAULT), ("q_vel", 3, _Q_VEL_DEFAULT)) + \
                    ((("q_int", 6, _Q_INT_DEFAULT),) if lqi else ()):
                col = block.column(name)
                q[:, at:at + 3] = default if col is None else col
            rc = block.column("r_controls")
            r = np.tile(np.asarray(_R_DEFAULT, dtype=np.float64), (C, 1)) if rc is None else np.ascontiguousarray(rc)
            # the non-tuned controller parameters (dt, mass, limits, ...) come from a prototype built WITHOUT solving
            # anything: default weights always solve, the candidates' own weights are solved in the batch below
            neutral = {k: v for k, v in proto_cfg.items() if k not in ("q_pos", "q_vel", "q_int", "r_controls")}
            if lqi:
                neutral["q_int"] = [0.01, 0.01, 0.1]
            proto = RiccatiLQRController(config=neutral)
            qd, rd = torch.as_tensor(q, **f64), torch.as_tensor(r, **f64)
            A, B = torch.empty((n, n), **f64), torch.empty((n, 4), **f64)
            Q, R = torch.empty((C, n, n), **f64), torch.empty((C, 4, 4), **f64)
            ctx.build_lqr_systems(n, C, proto.dt, proto.mass, proto.gravity, qd, rd, A, B, Q, R)
            K = torch.empty((C, 4, n), **f64)
            iters = torch.zeros(C, dtype=torch.int32, device=ctx.device)
            info = torch.zeros(C, dtype=torch.int32, device=ctx.device)
            ctx.dare_batched(n, 4, C, A, 0, B, 0, Q, n * n, R, 16, None, K, iters, info)
            self.last_dare_info, self.last_dare_iters = info, iters
            retry = torch.nonzero(info != 0).reshape(-1).cpu().tolist()
            if retry:
                K[torch.as_tensor(retry, device=ctx.device)] = 0.0   # placeholder rows, re-evaluated below
        prm = proto.kernel_params(env_config, criteria)
        ff_gains = None
        if proto.feedforward_enabled:   # per-candidate feedforward gains (tuning.py:700-712)
            ffh = np.zeros((C, 6))
            for name, at in (("ff_velocity_gain", 0), ("ff_acceleration_gain", 3)):
                col = block.column(name)
                if col is not None:
                    ffh[:, at:at + 3] = col
            ff_gains = torch.as_tensor(ffh, **f64)
        cand = self._sweep_device(ctx, prm, K, proto.k_cols(), ff_gains, want_metrics=self.keep_episode_metrics)
        rows = cand.cpu().numpy()
        # candidates whose DARE failed (or whose weights the validation rejected) go through their own controller
        # object, one by one, exactly as the reference evaluates them: fallback gain, fallback parameters (no
        # feedforward, no integral action), ValueError for weights the constructor refuses
        for i in retry:
            rows[:, i] = self._evaluate_one(configs[i] if configs is not None else block.config(i))
        return rows

    # -- search loops ----------------------------------------------------------------------------------
    def _record_result(self, config: dict, score: float, metrics: dict) -> None:
        self.results.append({"config": config, "score": score, "metrics": metrics})
        if score > self.best_score:  # strict: the first of equal scores wins (tuning.py:1084)
            self.best_score, self.best_config, self.best_metrics = score, config, metrics

    def tune(self) -> TuningResult:
        """tuning.py:930-982"""
        cfg = self.config
        if cfg.resume_from:
            prev = TuningResult.load(cfg.resume_from)
            self.results, self.best_config = prev.all_results, prev.best_config
            self.best_score, self.best_metrics = prev.best_score, prev.best_metrics
        if cfg.strategy == "grid":
            configs = self._generate_grid_configs()
            configs = configs[: min(len(configs), cfg.max_iterations)]
        elif cfg.strategy == "cma_es":
            self._run_cma_es_search()
            configs = []
        else:
            remaining = max(cfg.max_iterations - len(self.results), 0)
            self._record_block(*self.sweep(remaining))
            configs = []
        # under torchrun the candidates are split over the ranks (one all_gather of the score rows at the end);
        # every rank ends with the same result
        from .dist import evaluate_candidates_sharded, is_primary, barrier

        if configs:
            scores, metrics = evaluate_candidates_sharded(self, configs)
            for c, s, m in zip(configs, scores, metrics):
                self._record_result(c, s, m)
        result = TuningResult(
            best_config=self.best_config, best_score=self.best_score, best_metrics=self.best_metrics,
            all_results=self.results, iterations_completed=len(self.results), interrupted=self.interrupted,
            timestamp=datetime.now(timezone.utc).isoformat(), config=cfg.to_dict())
        if is_primary():          # one writer: N ranks would race on the same second-resolution file names
            self._save_results(result)
            if cfg.strategy == "cma_es":
                self._save_cma_checkpoint()
        barrier()
        return result

    def sweep(self, count: int) -> tuple[CandidateBlock, np.ndarray]:
        """Random search as arrays (BASELINE config 4): draw ``count`` candidates (same generator stream as the
        reference's loop), evaluate them — over all ranks when launched under torchrun, each rank drawing and
        scoring its own contiguous shard — and return ``(block, rows)`` with rows float64 [4, count] = score,
        mean_on_target_ratio, mean_tracking_error, success_rate.  Nothing per candidate is created in Python."""
        from .dist import gather_candidate_rows, shard_range, world

        rank, size, _ = world()
        if size == 1:
            block = self.sample_candidates(count)
            return block, self.evaluate_block(block)
        lo, hi = shard_range(count, rank, size)
        local = self.sample_candidates(count, lo, hi)
        rows = self.evaluate_block(local) if hi > lo else np.zeros((4, 0))
        k = local.values.shape[1]
        packed = np.concatenate([rows, local.values.T.reshape(k, hi - lo)], axis=0)     # [4 + k, C_local]
        full = gather_candidate_rows(packed, count)
        block = CandidateBlock(local.spec, np.ascontiguousarray(full[4:].T), local.use_lqi, local.zero_q_int)
        return block, np.ascontiguousarray(full[:4])

    def _record_block(self, block: CandidateBlock, rows: np.ndarray) -> None:
        """_record_result for a whole block: the result records are built in one pass, the best candidate is the
        first of the highest score (strict '>' scan of tuning.py:1084, which also skips NaN scores)."""
        from .dist import argmax_first

        scores, metrics = self._rows_to_lists(rows)
        configs = block.configs()
        self.results.extend({"config": c, "score": s, "metrics": m} for c, s, m in zip(configs, scores, metrics))
        if len(scores):
            i = argmax_first(rows[0])
            if i >= 0 and scores[i] > self.best_score:
                self.best_score, self.best_config, self.best_metrics = scores[i], configs[i], metrics[i]

    # -- CMA-ES (tuning.py:1156-1428): specification, bounds, x0 and sigma0 as in the reference; ask/tell
    # from cmaes.py; one generation = one batched evaluation ------------------------------------------------
    def _build_cma_param_spec(self) -> list[tuple[str, int, tuple]]:
        sp = self.config.search_space
        spec = []
        for name, dim in (("kp_pos", 3), ("ki_pos", 3), ("kd_pos", 3), ("ff_velocity_gain", 3),
                          ("ff_acceleration_gain", 3), ("q_pos", 3), ("q_vel", 3), ("r_thrust", 1), ("r_rate", 1),
                          ("r_controls", 4), ("q_int", 3)):
            rng_def = getattr(sp, name + "_range")
            if rng_def is not None:
                spec.append((name, dim, rng_def))
        return spec

    def _get_cma_bounds(self) -> tuple[list[float], list[float]]:
        lower, upper = [], []
        for _name, dim, bounds in self._cma_param_spec:
            if dim == 1:
                lower.append(bounds[0])
                upper.append(bounds[1])
            else:
                lower.extend(bounds[0])
                upper.extend(bounds[1])
        return lower, upper

    def _get_cma_x0(self) -> list[float]:
        lower, upper = self._get_cma_bounds()
        x0, idx = [], 0
        for name, dim, _ in self._cma_param_spec:
            log_scale = name in {"q_pos", "q_vel", "r_controls", "q_int"}
            for _ in range(dim):
                lo, hi = lower[idx], upper[idx]
                x0.append(float(np.sqrt(lo * hi)) if (log_scale and lo > 0 and hi > 0) else (lo + hi) / 2.0)
                idx += 1
        return x0

    def _vector_to_config(self, x) -> dict:
        cfg: dict = {}
        idx = 0
        for name, dim, _ in self._cma_param_spec:
            if dim == 1:
                cfg[name] = float(x[idx])
                idx += 1
            else:
                cfg[name] = [float(x[idx + i]) for i in range(dim)]
                idx += dim
        if "ff_velocity_gain" in cfg or "ff_acceleration_gain" in cfg:
            cfg["feedforward_enabled"] = True
        if "q_int" in cfg or self.config.search_space.use_lqi:
            cfg["use_lqi"] = True
        return cfg

    def _run_cma_es_search(self) -> None:
        from .cmaes import CMAEvolutionStrategy

        cfg = self.config
        self._cma_param_spec = self._build_cma_param_spec()
        if not self._cma_param_spec:
            logger.warning("No parameters to tune. CMA-ES requires at least one range.")
            return
        lower, upper = self._get_cma_bounds()
        x0 = self._get_cma_x0()
        ranges = [hi - lo for lo, hi in zip(lower, upper) if hi > lo]
        sigma0 = 1e-5 if not ranges else cfg.cma_sigma0 * float(np.exp(np.mean(np.log(ranges))))
        # resume: the optimiser's state (mean, step size, covariance, generator) from the checkpoint that sits next to
        # the results file being resumed (tuning.py:1357-1375)
        self.cma_es = None
        if cfg.resume_from:
            checkpoint = Path(cfg.resume_from).parent / "cma_checkpoint.pkl"
            if checkpoint.exists():
                try:
                    with open(checkpoint, "rb") as f:
                        state = pickle.load(f)
                    if isinstance(state, CMAEvolutionStrategy) and state.n == len(x0):
                        self.cma_es = state
                        self.cma_es.maxfevals = max(self.cma_es.maxfevals, self.cma_es.counteval + cfg.max_iterations)
                        logger.info("Resumed CMA-ES state from %s", checkpoint)
                except Exception as e:  # a damaged checkpoint only costs the optimiser's memory, not the run
                    logger.warning("Failed to load CMA-ES checkpoint: %s", e)
        if self.cma_es is None:
            self.cma_es = CMAEvolutionStrategy(x0, sigma0, lower, upper, seed=cfg.seed, popsize=cfg.cma_popsize,
                                               tolfun=1e-11, tolx=1e-12, maxfevals=cfg.max_iterations)
        self._cma_checkpoint_path = self.output_dir / "cma_checkpoint.pkl"
        num_evals = len(self.results)
        while not self.cma_es.stop() and num_evals < cfg.max_iterations and not self.interrupted:
            solutions = self.cma_es.ask()[: cfg.max_iterations - num_evals]
            repaired = [self.cma_es.repair(x) for x in solutions]
            configs = [self._vector_to_config(x) for x in repaired]
            from .dist import evaluate_candidates_sharded

            scores, metrics = evaluate_candidates_sharded(self, configs)   # the whole generation: one launch per rank
            for c, s, m in zip(configs, scores, metrics):
                self._record_result(c, s, m)
            num_evals += len(solutions)
            if len(solutions) >= self.cma_es.mu:
                self.cma_es.tell(solutions, [self.cma_es.penalised(x, -s) for x, s in zip(solutions, scores)])

    def _save_cma_checkpoint(self) -> None:
        """tuning.py:1430-1442: the optimiser object, pickled, for ``resume_from``."""
        if self.cma_es is None or self._cma_checkpoint_path is None:
            return
        try:
            self.output_dir.mkdir(parents=True, exist_ok=True)
            with open(self._cma_checkpoint_path, "wb") as f:
                pickle.dump(self.cma_es, f)
            logger.info("Saved CMA-ES checkpoint to %s", self._cma_checkpoint_path)
        except Exception as e:
            logger.warning("Failed to save CMA-ES checkpoint: %s", e)

    def _save_results(self, result: TuningResult) -> None:
        """tuning.py:1103-1150: timestamped results + a stable "latest" copy."""
        self.output_dir.mkdir(parents=True, exist_ok=True)
        stamp = datetime.now(timezone.utc).strftime("%Y%m%d_%H%M%S")
        result.save(self.output_dir / f"tuning_{self.config.controller_type}_{stamp}_results.json")
        best = {"controller_type": self.config.controller_type,
                self.config.controller_type: result.best_config,
                "metrics": result.best_metrics, "score": result.best_score}
        with open(self.output_dir / f"tuning_{self.config.controller_type}_{stamp}_best_config.json", "w") as f:
            json.dump(best, f, indent=2)
