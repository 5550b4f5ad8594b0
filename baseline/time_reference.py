"""Time the UNMODIFIED Python reference (baseline/_ref/quadcopter_tracking) on this machine's host cores — the CPU
baseline BASELINE.md §3 asks for, next to the C port that bench.py times as its `--impl reference` arm.

    python baseline/time_reference.py [--procs N] [--episodes 10] [--dare-samples 1000]

Prints ONE JSON object:
  evaluate_1core      Evaluator.evaluate(num_episodes=10, base_seed=42) with load_controller("riccati_lqr"), stationary
                      target (eval.py:169-231; BASELINE config 1): env-steps/s on one core
  evaluate_nproc      the same call in N processes on disjoint seeds (the reference itself is single-threaded)
  solve_dare_1core    solve_dare over sampled (Q, R) of the tuning ranges (riccati_lqr.py:119-184): solves/s
  solve_dare_nproc    the same spread over N processes
  evaluate_config     ControllerTuner._evaluate_config for 8 candidates x 5 episodes (tuning.py:846-928), one candidate
                      per process
Every worker runs with OMP/OpenBLAS threads pinned to 1 so N processes do not oversubscribe.  Nothing of this repo's
product code is imported.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import pathlib
import sys
import time

HERE = pathlib.Path(__file__).resolve().parent
for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[var] = "1"


def _use_reference():
    sys.path.insert(0, str(HERE.parent))
    from baseline.install_reference import reference_paths

    paths = reference_paths()
    if not paths:
        raise SystemExit("baseline/_ref is not installed: run `python baseline/install_reference.py` where /root/reference exists")
    sys.path[:0] = paths
    import logging

    logging.disable(logging.CRITICAL)


def _evaluate(args):
    episodes, seed = args
    _use_reference()
    import tempfile

    from quadcopter_tracking.env import EnvConfig
    from quadcopter_tracking.eval import Evaluator, load_controller
    from quadcopter_tracking.utils.metrics import SuccessCriteria

    cfg = EnvConfig()
    cfg.target.motion_type = "stationary"
    with tempfile.TemporaryDirectory() as out:
        ev = Evaluator(load_controller("riccati_lqr", config={}), cfg, SuccessCriteria(), output_dir=out)
        t0 = time.perf_counter()
        summary = ev.evaluate(num_episodes=episodes, base_seed=seed, verbose=False)
        dt = time.perf_counter() - t0
        steps = sum(len(d) for d in ev.episode_data_list)
    return steps, dt, summary.mean_on_target_ratio


def _dare(args):
    samples, seed = args
    _use_reference()
    import numpy as np

    from quadcopter_tracking.controllers.riccati_lqr import build_linearized_system, solve_dare

    A, B = build_linearized_system(dt=0.01, mass=1.0, gravity=9.81)
    rng = np.random.default_rng(seed)
    q = rng.uniform([5e-5, 5e-5, 10, 1e-3, 1e-3, 2], [5e-4, 5e-4, 25, 1e-2, 1e-2, 8], (samples, 6))
    r = rng.uniform(0.5, 2.0, (samples, 4))
    solve_dare(A, B, np.diag(q[0]), np.diag(r[0]))        # import scipy outside the timed region
    t0 = time.perf_counter()
    for i in range(samples):
        solve_dare(A, B, np.diag(q[i]), np.diag(r[i]))
    return samples, time.perf_counter() - t0


def _candidate(args):
    index, episodes = args
    _use_reference()
    import tempfile

    from quadcopter_tracking.controllers.tuning import ControllerTuner, GainSearchSpace, TuningConfig

    space = GainSearchSpace(q_pos_range=([5e-5, 5e-5, 10.0], [5e-4, 5e-4, 25.0]), q_vel_range=([1e-3, 1e-3, 2.0], [1e-2, 1e-2, 8.0]),
                            r_controls_range=([0.5] * 4, [2.0] * 4))
    with tempfile.TemporaryDirectory() as out:
        tuner = ControllerTuner(TuningConfig(controller_type="riccati_lqr", search_space=space, strategy="random",
                                             max_iterations=8, evaluation_episodes=episodes, evaluation_horizon=3000,
                                             seed=42, output_dir=out))
        cfgs = [tuner._generate_random_config() for _ in range(index + 1)]
        t0 = time.perf_counter()
        score, metrics = tuner._evaluate_config(cfgs[index])
        dt = time.perf_counter() - t0
    return score, dt, episodes * 3000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--episodes", type=int, default=10)
    ap.add_argument("--dare-samples", type=int, default=1000)
    ap.add_argument("--candidates", type=int, default=8)
    args = ap.parse_args()
    n = max(1, args.procs)
    ctx = mp.get_context("spawn")
    out = {"cores": n, "python": sys.version.split()[0]}
    with ctx.Pool(n) as pool:
        pool.map(_dare, [(1, 0)] * n)                           # warm every worker: imports happen here
        steps, dt, ratio = pool.apply(_evaluate, ((args.episodes, 42),))
        out["evaluate_1core"] = {"env_steps_per_s": steps / dt, "seconds": dt, "episodes": args.episodes,
                                 "mean_on_target_ratio": ratio}
        t0 = time.perf_counter()
        res = pool.map(_evaluate, [(args.episodes, 42 + args.episodes * p) for p in range(n)])
        wall = time.perf_counter() - t0
        out["evaluate_nproc"] = {"env_steps_per_s": sum(r[0] for r in res) / wall, "seconds": wall, "processes": n,
                                 "episodes_per_process": args.episodes}
        s, dt = pool.apply(_dare, ((args.dare_samples, 1),))
        out["solve_dare_1core"] = {"solves_per_s": s / dt, "seconds": dt, "samples": s}
        per = max(50, args.dare_samples // 4)
        t0 = time.perf_counter()
        res = pool.map(_dare, [(per, 2 + p) for p in range(n)])
        wall = time.perf_counter() - t0
        out["solve_dare_nproc"] = {"solves_per_s": sum(r[0] for r in res) / wall, "seconds": wall, "processes": n,
                                   "samples_per_process": per}
        t0 = time.perf_counter()
        res = pool.map(_candidate, [(i, 5) for i in range(args.candidates)])
        wall = time.perf_counter() - t0
        out["evaluate_config"] = {"candidates": args.candidates, "episodes_per_candidate": 5, "seconds_wall": wall,
                                  "processes": min(n, args.candidates),
                                  "seconds_per_candidate_1core": sum(r[1] for r in res) / len(res),
                                  "env_steps_per_s_1core": sum(r[2] for r in res) / sum(r[1] for r in res),
                                  "env_steps_per_s_wall": sum(r[2] for r in res) / wall}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
