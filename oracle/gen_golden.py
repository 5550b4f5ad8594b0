#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the REFERENCE itself.

Run in the build container, where /root/reference exists:

    python oracle/gen_golden.py

The reference (quadcopter_tracking 0.5.0) is imported read-only from /root/reference/src with a
two-file matplotlib stub (matplotlib is absent from the image and only needed at import time by
quadcopter_tracking.eval).  Nothing of the reference is copied: only numbers it computes are
stored.  Versions used for the committed fixtures: Python 3.12.3, numpy 2.3.5, scipy 1.18.1
(the reference's lock file pins numpy 1.26.4 / scipy 1.16.3).

These fixtures pin (a) the CPU oracle (tests/test_oracle_golden.py, CPU) and (b) the CUDA path
(tests/test_golden_gpu.py, GPU); /root/reference is never read at test time.
"""
from __future__ import annotations

import logging
import os
import pathlib
import sys
import tempfile

import numpy as np

REF = pathlib.Path("/root/reference/src")
OUT = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden"


def _import_reference():
    stub = pathlib.Path(tempfile.mkdtemp()) / "matplotlib"
    stub.mkdir()
    (stub / "__init__.py").write_text("def use(*a, **k): pass\n")
    (stub / "pyplot.py").write_text("class Figure: pass\nclass Axes: pass\ndef close(*a, **k): pass\n")
    sys.path.insert(0, str(stub.parent))
    sys.path.insert(0, str(REF))
    logging.disable(logging.CRITICAL)
    import quadcopter_tracking  # noqa: F401


MOTIONS = ["stationary", "linear", "circular", "sinusoidal", "figure8"]


def pattern_params(env) -> np.ndarray:
    """Target pattern of a reset reference env -> the 12-vector layout of enum qt_motion."""
    pat = env.target._pattern
    tp = np.zeros(12)
    name = type(pat).__name__
    if name == "StationaryMotion":
        tp[0:3] = pat.position
    elif name == "LinearMotion":
        tp[0:3], tp[3:6] = pat.start, pat.velocity
    elif name == "CircularMotion":
        tp[0:3] = pat.center
        tp[3], tp[4], tp[5] = pat.radius, pat.omega, pat.initial_angle
    elif name == "SinusoidalMotion":
        tp[0:3], tp[3:6], tp[6:9], tp[9:12] = pat.center, pat.amplitude, pat.omega, pat.phase
    elif name == "Figure8Motion":
        tp[0:3] = pat.center
        tp[3], tp[4] = pat.scale, pat.omega
    return tp


def gen_reset():
    from quadcopter_tracking.env import EnvConfig, QuadcopterEnv

    seeds = np.concatenate([np.arange(42, 42 + 32), [0, 1, 2**31 - 1, 2**32 - 1, 2**32, 2**40 + 17, 2**63 - 1]])
    out = {"seeds": seeds.astype(np.uint64)}
    for motion in MOTIONS:
        cfg = EnvConfig()
        cfg.target.motion_type = motion
        st, tp = [], []
        for s in seeds:
            env = QuadcopterEnv(config=cfg)
            env.reset(seed=int(s))
            st.append(env.get_state_vector())
            tp.append(pattern_params(env))
        out[f"{motion}_state0"] = np.array(st).T
        out[f"{motion}_tparams"] = np.array(tp).T
    # reset(seed=None) keeps drawing from the existing generators: three consecutive resets
    for motion in ("linear", "sinusoidal"):
        cfg = EnvConfig()
        cfg.target.motion_type = motion
        cfg.seed = 7
        env = QuadcopterEnv(config=cfg)
        st, tp = [], []
        for _ in range(3):
            env.reset()
            st.append(env.get_state_vector())
            tp.append(pattern_params(env))
        out[f"{motion}_noseed_state0"] = np.array(st).T
        out[f"{motion}_noseed_tparams"] = np.array(tp).T
    np.savez_compressed(OUT / "reset.npz", **out)


def gen_dare():
    from quadcopter_tracking.controllers.riccati_lqr import (RiccatiLQRController, build_augmented_lqi_system,
                                                             build_linearized_system, solve_dare)

    out = {}
    for name, cfg in (("lqr_default", {}), ("lqi_default", {"use_lqi": True, "q_int": [0.01, 0.01, 0.1]}),
                      ("lqi_zero_qint", {"use_lqi": True})):
        c = RiccatiLQRController(config=cfg)
        A = c.A_aug if c.use_lqi else c.A
        B = c.B_aug if c.use_lqi else c.B
        out[f"{name}_A"], out[f"{name}_B"], out[f"{name}_Q"], out[f"{name}_R"] = A, B, c.Q, c.R
        out[f"{name}_P"], out[f"{name}_K"] = c.P, c.K
    rng = np.random.default_rng(2024)
    A6, B6 = build_linearized_system(0.01)
    A9, B9 = build_augmented_lqi_system(A6, B6, 0.01)
    lo6 = np.array([5e-5, 5e-5, 10.0, 1e-3, 1e-3, 2.0])
    hi6 = np.array([5e-4, 5e-4, 25.0, 1e-2, 1e-2, 8.0])
    lo_i, hi_i = np.array([1e-3, 1e-3, 1e-2]), np.array([0.1, 0.1, 1.0])
    for n, A, B, tag in ((6, A6, B6, "sweep6"), (9, A9, B9, "sweep9")):
        q, r, Ps, Ks = [], [], [], []
        for _ in range(48):
            qd = rng.uniform(lo6, hi6)
            if n == 9:
                qd = np.concatenate([qd, rng.uniform(lo_i, hi_i)])
            rd = rng.uniform(0.5, 2.0, 4)
            P, K = solve_dare(A, B, np.diag(qd), np.diag(rd))
            q.append(qd), r.append(rd), Ps.append(P), Ks.append(K)
        out[f"{tag}_q"], out[f"{tag}_r"] = np.array(q), np.array(r)
        out[f"{tag}_P"], out[f"{tag}_K"] = np.array(Ps), np.array(Ks)
    # dense symmetric positive definite Q / R on the reference's own A, B
    Qs, Rs, Ps, Ks = [], [], [], []
    for _ in range(12):
        M = rng.standard_normal((6, 6))
        Q = M @ M.T * 0.05 + np.diag(rng.uniform(lo6, hi6))
        N = rng.standard_normal((4, 4))
        R = N @ N.T * 0.1 + np.eye(4)
        P, K = solve_dare(A6, B6, Q, R)
        Qs.append(Q), Rs.append(R), Ps.append(P), Ks.append(K)
    out["dense6_Q"], out["dense6_R"], out["dense6_P"], out["dense6_K"] = map(np.array, (Qs, Rs, Ps, Ks))
    # generic random systems of other sizes through the same public solve_dare
    for n, m in ((2, 1), (3, 2), (4, 4), (5, 3), (7, 2), (8, 8), (12, 4), (15, 4), (16, 8)):
        As, Bs, Qs, Rs, Ps, Ks = [], [], [], [], [], []
        for _ in range(4):
            A = rng.standard_normal((n, n))
            A *= 1.05 / max(abs(np.linalg.eigvals(A)))
            B = rng.standard_normal((n, m))
            M = rng.standard_normal((n, n))
            Q = M @ M.T / n + 0.1 * np.eye(n)
            N = rng.standard_normal((m, m))
            R = N @ N.T / m + np.eye(m)
            P, K = solve_dare(A, B, Q, R)
            As.append(A), Bs.append(B), Qs.append(Q), Rs.append(R), Ps.append(P), Ks.append(K)
        for nm, arr in zip("ABQRPK", (As, Bs, Qs, Rs, Ps, Ks)):
            out[f"gen{n}x{m}_{nm}"] = np.array(arr)
    # validation: matrices solve_dare must reject (ValueError)
    bad = {
        "bad_Q_negative": (np.diag([1.0, 1, 1, 1, 1, -1e-6]), np.eye(4)),
        "bad_Q_asym": (np.eye(6) + np.triu(np.full((6, 6), 1e-3), 1), np.eye(4)),
        "bad_R_singular": (np.eye(6), np.diag([1.0, 1, 1, 0.0])),
        "ok_Q_tiny_negative": (np.diag([1.0, 1, 1, 1, 1, -1e-12]), np.eye(4)),
    }
    for k, (Q, R) in bad.items():
        try:
            solve_dare(A6, B6, Q, R)
            ok = 1
        except ValueError:
            ok = 0
        out[f"{k}_Q"], out[f"{k}_R"], out[f"{k}_accepted"] = Q, R, np.array(ok)
    np.savez_compressed(OUT / "dare.npz", **out)


SAMPLE_EVERY = 250


def run_reference_episode(env_cfg, controller, seed, criteria, max_steps=None, reset_controller=True):
    """The driver loop of eval.py:95-167 / tuning.py:871-910, recording what the fixtures need."""
    from quadcopter_tracking.env import QuadcopterEnv
    from quadcopter_tracking.utils.metrics import compute_episode_metrics

    env = QuadcopterEnv(config=env_cfg)
    obs = env.reset(seed=seed)
    if reset_controller:
        controller.reset()
    state0, tp = env.get_state_vector(), pattern_params(env)
    data, samples = [], []
    done, step, info = False, 0, {}
    while not done:
        if max_steps is not None and step >= max_steps:
            break
        if step % SAMPLE_EVERY == 0:
            samples.append(env.get_state_vector())
        action = controller.compute_action(obs)
        nxt, reward, done, info = env.step(action)
        data.append({
            "time": info["time"], "quadcopter_position": obs["quadcopter"]["position"].tolist(),
            "target_position": obs["target"]["position"].tolist(),
            "action": [action["thrust"], action["roll_rate"], action["pitch_rate"], action["yaw_rate"]],
        })
        obs = nxt
        step += 1
    m = compute_episode_metrics(data, criteria, info)
    term = {"": 0, "time_limit": 1, "position_bounds": 2, "numerical_instability": 3}[info.get("termination_reason", "")]
    vec = np.array([
        step, info["time"], env._on_target_count, round(m.on_target_ratio * step), m.mean_tracking_error * step,
        m.max_tracking_error, m.rms_tracking_error ** 2 * step, m.total_control_effort, m.overshoot_count,
        m.max_overshoot, float(info.get("success", env._evaluate_success())), float(m.success), term,
        info["action_violations"], getattr(controller, "get_saturation_count", lambda: 0)(), info["tracking_error"]])
    extra = np.array([m.on_target_ratio, m.mean_tracking_error, m.rms_tracking_error, m.mean_control_effort])
    if hasattr(controller, "integral_error"):     # PID: integral + time of the previous call
        integ = controller.integral_error.copy()
    else:
        integ = getattr(controller, "integral_state", None)
        integ = integ.copy() if integ is not None else np.zeros(3)
    return dict(state0=state0, tparams=tp, metrics=vec, extra=extra, final_state=env.get_state_vector(),
                samples=np.array(samples), integral=integ)


def gen_rollouts():
    from quadcopter_tracking.controllers import RiccatiLQRController
    from quadcopter_tracking.env import EnvConfig
    from quadcopter_tracking.utils.metrics import SuccessCriteria

    cases = []

    def env_cfg(motion, **kw):
        c = EnvConfig()
        c.target.motion_type = motion
        for k, v in kw.items():
            sect, name = k.split("__")
            setattr(getattr(c, sect), name, v)
        return c

    lqi = {"use_lqi": True, "q_int": [0.01, 0.01, 0.1]}
    for motion in MOTIONS:
        cases.append((f"lqr_{motion}", env_cfg(motion), {}, list(range(42, 50)), None, True, SuccessCriteria()))
    for motion in ("stationary", "circular", "sinusoidal"):
        cases.append((f"lqi_{motion}", env_cfg(motion), lqi, list(range(42, 46)), None, True, SuccessCriteria()))
    # Evaluator semantics: the controller is never reset, the integrator carries over (eval.py:198-213)
    cases.append(("lqi_carry_stationary", env_cfg("stationary"), lqi, [42, 43, 44], None, False, SuccessCriteria()))
    # a well-behaved integral gain (the eval default diverges): exercises the anti-windup path mildly
    cases.append(("lqi_small_linear", env_cfg("linear"),
                  {"use_lqi": True, "q_int": [1e-6, 1e-6, 1e-3], "integral_limit": 0.5}, [42, 43, 44, 45],
                  None, True, SuccessCriteria()))
    cases.append(("lqr_ff_circular", env_cfg("circular"),
                  {"feedforward_enabled": True, "ff_velocity_gain": [0.1, 0.1, 0.0],
                   "ff_acceleration_gain": [0.05, 0.05, 0.0]}, [42, 43, 44, 45], None, True, SuccessCriteria()))
    cases.append(("lqr_ff_sinusoidal", env_cfg("sinusoidal"),
                  {"feedforward_enabled": True, "ff_velocity_gain": 0.2, "ff_acceleration_gain": 0.1,
                   "ff_max_velocity": 1.5, "ff_max_acceleration": 2.0}, [42, 43], 1500, True, SuccessCriteria()))
    cases.append(("lqr_euler_linear", env_cfg("linear", simulation__integrator="euler"), {}, [42, 43, 44, 45],
                  1000, True, SuccessCriteria()))
    cases.append(("lqr_custom_circular",
                  env_cfg("circular", quadcopter__mass=1.3, quadcopter__drag_coeff_linear=0.2,
                          simulation__dt=0.02, simulation__max_episode_time=10.0, target__speed=2.0,
                          target__radius=3.0, target__center=(1.0, -1.0, 2.0),
                          success_criteria__target_radius=0.3),
                  {"mass": 1.3, "dt": 0.02, "q_pos": [2e-4, 2e-4, 12.0], "r_controls": [0.8, 1.2, 1.2, 1.0]},
                  [42, 43, 44, 45], None, True,
                  SuccessCriteria(min_on_target_ratio=0.5, min_episode_duration=10.0, target_radius=0.3)))
    cases.append(("lqr_bounds_linear", env_cfg("linear", simulation__max_position=4.0), {}, [42, 43, 44, 45],
                  None, True, SuccessCriteria()))
    cases.append(("lqr_cap_stationary", env_cfg("stationary"), {}, [42, 43], 100, True, SuccessCriteria()))
    cases.append(("lqr_short_stationary", env_cfg("stationary"), {}, [42], 5, True, SuccessCriteria()))

    out = {}
    for name, cfg, ctrl_cfg, seeds, max_steps, reset_ctrl, crit in cases:
        controller = RiccatiLQRController(config=ctrl_cfg)
        recs = [run_reference_episode(cfg, controller, s, crit, max_steps, reset_ctrl) for s in seeds]
        out[f"{name}_seeds"] = np.array(seeds, dtype=np.uint64)
        out[f"{name}_K"] = controller.K
        for key in ("state0", "tparams", "metrics", "extra", "final_state", "integral"):
            out[f"{name}_{key}"] = np.array([r[key] for r in recs]).T
        out[f"{name}_samples"] = np.array([r["samples"] for r in recs])  # [E, n_samples, 12]
        print("  rollout", name, "done", flush=True)
    np.savez_compressed(OUT / "rollouts.npz", **out)


def gen_eval_summary():
    from quadcopter_tracking.env import EnvConfig
    from quadcopter_tracking.eval import Evaluator, load_controller
    from quadcopter_tracking.utils.metrics import SuccessCriteria

    out = {}
    with tempfile.TemporaryDirectory() as td:
        ev = Evaluator(load_controller("riccati_lqr", config={}), EnvConfig(), SuccessCriteria(), output_dir=td)
        s = ev.evaluate(num_episodes=10, base_seed=42, verbose=False)
    out["config1_summary"] = np.array([
        s.total_episodes, s.successful_episodes, s.success_rate, s.mean_on_target_ratio, s.std_on_target_ratio,
        s.mean_tracking_error, s.std_tracking_error, s.mean_control_effort, s.best_episode_idx,
        s.worst_episode_idx, float(s.meets_criteria)])
    out["config1_ratios"] = np.array([m.on_target_ratio for m in s.episode_metrics])
    out["config1_mean_err"] = np.array([m.mean_tracking_error for m in s.episode_metrics])
    out["config1_env_ratios"] = np.array([i["on_target_ratio"] for i in ev.episode_info_list])
    # first record of episode 0 (eval.py:142-159) to pin the per-step record layout
    d0 = ev.episode_data_list[0]
    for k in (0, 1, 2999):
        r = d0[k]
        out[f"config1_rec{k}"] = np.array([r["time"], *r["quadcopter_position"], *r["quadcopter_velocity"],
                                           *r["target_position"], *r["target_velocity"], *r["action"],
                                           r["reward"], r["tracking_error"], float(r["on_target"])])
    np.savez_compressed(OUT / "eval_summary.npz", **out)


def gen_step_edges():
    """Single env.step calls that exercise action repair, constraints and termination."""
    from quadcopter_tracking.env import EnvConfig, QuadcopterEnv

    rng = np.random.default_rng(7)
    states, actions, motions, times = [], [], [], []
    base = np.array([0.2, -0.1, 1.2, 0.3, -0.2, 0.1, 0.05, -0.04, 0.3, 0.1, -0.2, 0.05])
    specials = [
        (base, [9.81, 0, 0, 0]),
        (base, [25.0, 5.0, -7.0, 3.5]),            # everything above limits
        (base, [-3.0, 0.5, 0.5, 0.5]),             # thrust below min
        (base, [np.nan, 1.0, np.inf, -np.inf]),    # non-finite entries -> zeros
        (base, [np.nan, np.nan, np.nan, np.nan]),
        (base + np.array([0, 0, 0, 40, 30, 20, 0, 0, 0, 0, 0, 0.0]), [9.81, 0, 0, 0]),       # |v| > 50
        (base + np.array([0, 0, 0, 0, 0, 0, 0, 0, 0, 14, -12, 11.0]), [9.81, 0, 0, 0]),      # omega clip
        (base + np.array([0, 0, 0, 0, 0, 0, 0, 0, 3.1, 0, 0, 9.0]), [9.81, 0, 0, 3.0]),      # yaw wraps past +pi
        (base + np.array([0, 0, 0, 0, 0, 0, 0, 0, -3.5, 0, 0, -9.0]), [9.81, 0, 0, -3.0]),   # yaw wraps past -pi
        (base + np.array([0, 0, 0, 0, 0, 0, 1.0, -1.1, 0, 9, -9, 0.0]), [9.81, 3, -3, 0]),   # tilt clip
        (base + np.array([999.99, 0, 0, 45, 0, 0, 0, 0, 0, 0, 0, 0.0]), [9.81, 0, 0, 0]),    # crosses max_position
        (np.array([0, 0, 1, 0, 0, 0, 0, 0, np.pi, 0, 0, 0.0]), [9.81, 0, 0, 0]),             # yaw exactly pi
        (np.array([0, 0, 1, 0, 0, 0, 0, 0, -np.pi, 0, 0, 0.0]), [9.81, 0, 0, 0]),
        (np.array([0, 0, 1, 0, 0, 0, 7.0, -8.0, 12.0, 0, 0, 0.0]), [9.81, 0, 0, 0]),         # several turns
        (np.array([np.nan, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.0]), [9.81, 0, 0, 0]),            # NaN state
        (np.array([0, 0, 1, np.inf, 0, 0, 0, 0, 0, 0, 0, 0.0]), [9.81, 0, 0, 0]),
    ]
    for _ in range(24):
        s = rng.standard_normal(12) * np.array([3, 3, 3, 5, 5, 5, 0.8, 0.8, 2.5, 4, 4, 4])
        a = rng.standard_normal(4) * np.array([8, 2.5, 2.5, 2.5]) + np.array([9.81, 0, 0, 0])
        specials.append((s, a))
    out_rows = []
    for i, (s, a) in enumerate(specials):
        motion = MOTIONS[i % len(MOTIONS)]
        cfg = EnvConfig()
        cfg.target.motion_type = motion
        env = QuadcopterEnv(config=cfg)
        env.reset(seed=100 + i)
        n_pre = i % 4  # a few hover steps first so time / counters are non-trivial
        for _ in range(n_pre):
            env.step(env.hover_action())
        env.set_state_vector(np.asarray(s, dtype=float))
        t_before = env.time
        a = np.array(a, dtype=float)  # keep our copy: env.step repairs array actions in place
        obs, reward, done, info = env.step(a.copy())
        term = {"": 0, "time_limit": 1, "position_bounds": 2, "numerical_instability": 3}[info.get("termination_reason", "")]
        out_rows.append(dict(
            motion=MOTIONS.index(motion), seed=100 + i, n_pre=n_pre, state=np.asarray(s, dtype=float),
            action=np.asarray(a, dtype=float), tparams=pattern_params(env), time_before=t_before,
            next_state=env.get_state_vector(),
            obs=np.concatenate([obs["quadcopter"]["position"], obs["quadcopter"]["velocity"],
                                obs["quadcopter"]["attitude"], obs["quadcopter"]["angular_velocity"],
                                obs["target"]["position"], obs["target"]["velocity"],
                                obs["target"]["acceleration"], [obs["time"]]]),
            reward=reward, done=float(done), term=term, tracking_error=info["tracking_error"],
            on_target=float(info["on_target"]), on_target_ratio=info["on_target_ratio"],
            violations=info["action_violations"], step=info["step"]))
    out = {k: np.array([r[k] for r in out_rows]) for k in out_rows[0]}
    np.savez_compressed(OUT / "step_edges.npz", **out)


def gen_controller():
    """compute_action on hand-made observations, incl. an LQI integrator sequence with anti-windup."""
    from quadcopter_tracking.controllers import RiccatiLQRController

    rng = np.random.default_rng(11)
    out = {}

    def obs_from(v):
        return {"quadcopter": {"position": v[0:3], "velocity": v[3:6], "attitude": v[6:9],
                               "angular_velocity": v[9:12]},
                "target": {"position": v[12:15], "velocity": v[15:18], "acceleration": v[18:21]}, "time": v[21]}

    cfgs = {
        "lqr": {},
        "lqr_ff": {"feedforward_enabled": True, "ff_velocity_gain": [0.3, 0.2, 0.1],
                   "ff_acceleration_gain": [0.1, 0.2, 0.3], "ff_max_velocity": 2.0, "ff_max_acceleration": 1.0},
        "lqi": {"use_lqi": True, "q_int": [0.01, 0.01, 0.1], "integral_limit": 0.05,
                "integral_zero_threshold": 0.2},
    }
    for name, cfg in cfgs.items():
        c = RiccatiLQRController(config=cfg)
        obs_rows, act_rows, int_rows, sat_rows = [], [], [], []
        for k in range(60):
            v = rng.standard_normal(22) * np.array([2] * 3 + [3] * 3 + [0.3] * 6 + [2] * 3 + [3] * 3 + [4] * 3 + [1])
            if name == "lqi":
                # hold the error sign for a while so the integrator saturates, then flip it
                v[12:15] = v[0:3] + (np.array([1.0, -0.8, 0.6]) if (k // 20) % 2 == 0 else np.array([-1.0, 0.8, -0.6]))
                if k % 7 == 0:
                    v[12:15] = v[0:3] + 0.01  # below the zero threshold: integration pauses
            a = c.compute_action(obs_from(v))
            obs_rows.append(v)
            act_rows.append([a["thrust"], a["roll_rate"], a["pitch_rate"], a["yaw_rate"]])
            int_rows.append(c.integral_state.copy() if c.integral_state is not None else np.zeros(3))
            sat_rows.append(c.get_saturation_count())
        out[f"{name}_obs"], out[f"{name}_act"] = np.array(obs_rows).T, np.array(act_rows).T
        out[f"{name}_integral"], out[f"{name}_sat"] = np.array(int_rows).T, np.array(sat_rows)
        out[f"{name}_K"] = c.K
    np.savez_compressed(OUT / "controller.npz", **out)


def gen_tuning():
    from quadcopter_tracking.controllers.tuning import ControllerTuner, GainSearchSpace, TuningConfig

    out = {}
    spaces = {
        "riccati": GainSearchSpace(q_pos_range=([5e-5, 5e-5, 10.0], [5e-4, 5e-4, 25.0]),
                                   q_vel_range=([1e-3, 1e-3, 2.0], [1e-2, 1e-2, 8.0]),
                                   r_controls_range=([0.5] * 4, [2.0] * 4)),
        "riccati_lqi": GainSearchSpace(q_pos_range=([5e-5, 5e-5, 10.0], [5e-4, 5e-4, 25.0]),
                                       q_vel_range=([1e-3, 1e-3, 2.0], [1e-2, 1e-2, 8.0]),
                                       r_controls_range=([0.5] * 4, [2.0] * 4),
                                       q_int_range=([1e-6, 1e-6, 1e-4], [1e-5, 1e-5, 1e-2])),
    }
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        for name, space in spaces.items():
            for motion in ("stationary", "linear"):
                tc = TuningConfig(controller_type="riccati_lqr", search_space=space, strategy="random",
                                  max_iterations=5, evaluation_episodes=3, evaluation_horizon=400, seed=42,
                                  target_motion_type=motion, episode_length=4.0, output_dir="reports/tuning")
                tuner = ControllerTuner(tc)
                res = tuner.tune()
                key = f"{name}_{motion}"
                n = 9 if "lqi" in name else 6
                q = np.array([r["config"]["q_pos"] + r["config"]["q_vel"] + (r["config"].get("q_int", []) if n == 9 else [])
                              for r in res.all_results])
                out[f"{key}_q"] = q
                out[f"{key}_r"] = np.array([r["config"]["r_controls"] for r in res.all_results])
                out[f"{key}_score"] = np.array([r["score"] for r in res.all_results])
                out[f"{key}_metrics"] = np.array([[r["metrics"]["mean_on_target_ratio"], r["metrics"]["mean_tracking_error"],
                                                   r["metrics"]["success_rate"]] for r in res.all_results])
                out[f"{key}_best"] = np.array(int(np.argmax([r["score"] for r in res.all_results])))
                out[f"{key}_best_score"] = np.array(res.best_score)
                print("  tuning", key, "done", flush=True)
        # grid generation (host logic only)
        tc = TuningConfig(controller_type="riccati_lqr", strategy="grid", grid_points_per_dim=2, max_iterations=10,
                          search_space=GainSearchSpace(q_pos_range=([1e-4, 1e-4, 10.0], [1e-4, 2e-4, 20.0]),
                                                       r_controls_range=([1.0] * 4, [1.0, 1.0, 1.0, 2.0])))
        grid = ControllerTuner(tc)._generate_grid_configs()
        out["grid_q_pos"] = np.array([g["q_pos"] for g in grid])
        out["grid_r_controls"] = np.array([g["r_controls"] for g in grid])
    np.savez_compressed(OUT / "tuning.npz", **out)


def gen_pid_lqr():
    """§8(f).1: PID and the heuristic LQR through the same driver loops (pid_lqr.npz)."""
    from quadcopter_tracking.controllers import LQRController, PIDController
    from quadcopter_tracking.controllers.tuning import ControllerTuner, GainSearchSpace, TuningConfig
    from quadcopter_tracking.env import EnvConfig
    from quadcopter_tracking.eval import Evaluator, load_controller
    from quadcopter_tracking.utils.metrics import SuccessCriteria

    out = {}

    def env_cfg(motion):
        c = EnvConfig()
        c.target.motion_type = motion
        return c

    pid_ki = {"ki_pos": [0.002, 0.002, 0.5], "integral_limit": 0.4}
    pid_ff = {"feedforward_enabled": True, "ff_velocity_gain": [0.1, 0.1, 0.0], "ff_acceleration_gain": [0.05, 0.05, 0.0],
              "ff_max_velocity": 0.9, "ff_max_acceleration": 0.4}
    lqr_ff = {"feedforward_enabled": True, "ff_velocity_gain": 0.15, "ff_acceleration_gain": [0.05, 0.1, 0.0],
              "ff_max_velocity": 0.8}
    cases = [
        ("pid_stationary", PIDController, "stationary", {}, list(range(42, 48)), None, True),
        ("pid_linear", PIDController, "linear", {}, list(range(42, 48)), None, True),
        ("pid_circular", PIDController, "circular", {}, list(range(42, 46)), None, True),
        ("pid_sinusoidal", PIDController, "sinusoidal", {}, [42, 43], 1500, True),
        ("pidki_linear", PIDController, "linear", pid_ki, list(range(42, 46)), None, True),
        ("pidki_carry_stationary", PIDController, "stationary", pid_ki, [42, 43, 44], 1200, False),
        ("pidff_circular", PIDController, "circular", pid_ff, list(range(42, 46)), None, True),
        ("hlqr_linear", LQRController, "linear", {}, list(range(42, 48)), None, True),
        ("hlqr_circular", LQRController, "circular", {"q_pos": [2e-4, 3e-4, 12.0], "r_rate": 1.5, "r_thrust": 0.7},
         list(range(42, 46)), None, True),
        ("hlqrff_sinusoidal", LQRController, "sinusoidal", lqr_ff, [42, 43], 1500, True),
    ]
    for name, cls, motion, ccfg, seeds, max_steps, reset_ctrl in cases:
        controller = cls(config=ccfg)
        recs = [run_reference_episode(env_cfg(motion), controller, s, SuccessCriteria(), max_steps, reset_ctrl)
                for s in seeds]
        out[f"{name}_seeds"] = np.array(seeds, dtype=np.uint64)
        out[f"{name}_K"] = (np.stack([controller.kp_pos, controller.ki_pos, controller.kd_pos, np.zeros(3)]).astype(float)
                            if cls is PIDController else controller.K)
        for key in ("state0", "tparams", "metrics", "extra", "final_state", "integral"):
            out[f"{name}_{key}"] = np.array([r[key] for r in recs]).T
        out[f"{name}_samples"] = np.array([r["samples"] for r in recs])
        if cls is PIDController:
            out[f"{name}_last_time"] = np.array(np.nan if controller._last_time is None else controller._last_time)
        print("  rollout", name, "done", flush=True)

    # PID compute_action sequences on hand-made observations: times that advance, repeat and go back
    rng = np.random.default_rng(12)

    def obs_from(v):
        return {"quadcopter": {"position": v[0:3], "velocity": v[3:6], "attitude": v[6:9], "angular_velocity": v[9:12]},
                "target": {"position": v[12:15], "velocity": v[15:18], "acceleration": v[18:21]}, "time": float(v[21])}

    for name, ccfg in (("seq_pid", {}), ("seq_pidki", {"ki_pos": [0.5, 0.4, 0.3], "integral_limit": 0.05}),
                       ("seq_pidff", dict(pid_ff, ki_pos=0.2, integral_limit=2.0, kp_pos=1.5, kd_pos=[0.5, 0.6, 0.7]))):
        c = PIDController(config=ccfg)
        obs_rows, act_rows, int_rows, lt_rows = [], [], [], []
        t = 0.0
        for k in range(60):
            v = rng.standard_normal(22) * np.array([2] * 3 + [3] * 3 + [0.3] * 6 + [2] * 3 + [3] * 3 + [4] * 3 + [1])
            t = t + (0.01 if k % 11 else (-0.05 if k % 22 else 0.0))   # mostly forward; sometimes equal / backwards
            v[21] = t
            if k % 13 == 5:
                v[15:18] *= 0.01
                v[18:21] *= 0.01
            a = c.compute_action(obs_from(v))
            obs_rows.append(v)
            act_rows.append([a["thrust"], a["roll_rate"], a["pitch_rate"], a["yaw_rate"]])
            int_rows.append(c.integral_error.copy())
            lt_rows.append(c._last_time)
        out[f"{name}_obs"], out[f"{name}_act"] = np.array(obs_rows).T, np.array(act_rows).T
        out[f"{name}_integral"], out[f"{name}_last_time"] = np.array(int_rows).T, np.array(lt_rows)
        out[f"{name}_K"] = np.stack([c.kp_pos, c.ki_pos, c.kd_pos, np.zeros(3)]).astype(float)

    # Evaluator with the PID / heuristic LQR (eval --controller pid|lqr): the controller is never reset
    with tempfile.TemporaryDirectory() as td:
        for ctype in ("pid", "lqr"):
            cfg = EnvConfig()
            cfg.target.motion_type = "linear"
            ev = Evaluator(load_controller(ctype, config={}), cfg, SuccessCriteria(), output_dir=td)
            s = ev.evaluate(num_episodes=4, base_seed=42, verbose=False)
            out[f"eval_{ctype}_summary"] = np.array([
                s.total_episodes, s.successful_episodes, s.success_rate, s.mean_on_target_ratio, s.std_on_target_ratio,
                s.mean_tracking_error, s.std_tracking_error, s.mean_control_effort, s.best_episode_idx,
                s.worst_episode_idx, float(s.meets_criteria)])
            out[f"eval_{ctype}_ratios"] = np.array([m.on_target_ratio for m in s.episode_metrics])
            out[f"eval_{ctype}_mean_err"] = np.array([m.mean_tracking_error for m in s.episode_metrics])

    # tuner sweeps of tuning_pid_linear.yaml / tuning_lqr_linear.yaml (shortened)
    spaces = {
        "pid": GainSearchSpace(kp_pos_range=([0.005, 0.005, 2.0], [0.05, 0.05, 6.0]),
                               kd_pos_range=([0.02, 0.02, 1.0], [0.15, 0.15, 3.0])),
        "pidki": GainSearchSpace(kp_pos_range=([0.005, 0.005, 2.0], [0.05, 0.05, 6.0]),
                                 ki_pos_range=([0.0, 0.0, 0.0], [0.005, 0.005, 0.5]),
                                 kd_pos_range=([0.02, 0.02, 1.0], [0.15, 0.15, 3.0])),
        "lqr": GainSearchSpace(q_pos_range=([5e-5, 5e-5, 10.0], [5e-4, 5e-4, 25.0]),
                               q_vel_range=([1e-3, 1e-3, 2.0], [1e-2, 1e-2, 8.0])),
    }
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        for name, space in spaces.items():
            tc = TuningConfig(controller_type="lqr" if name == "lqr" else "pid", search_space=space, strategy="random",
                              max_iterations=5, evaluation_episodes=3, evaluation_horizon=400, seed=42,
                              target_motion_type="linear", episode_length=4.0, output_dir="reports/tuning")
            res = ControllerTuner(tc).tune()
            keys = ("kp_pos", "ki_pos", "kd_pos") if name != "lqr" else ("q_pos", "q_vel")
            out[f"tune_{name}_cfg"] = np.array([sum((r["config"].get(k, [0.0, 0.0, 0.0]) for k in keys), [])
                                                for r in res.all_results])
            out[f"tune_{name}_score"] = np.array([r["score"] for r in res.all_results])
            out[f"tune_{name}_metrics"] = np.array([[r["metrics"]["mean_on_target_ratio"], r["metrics"]["mean_tracking_error"],
                                                     r["metrics"]["success_rate"]] for r in res.all_results])
            out[f"tune_{name}_best_score"] = np.array(res.best_score)
            print("  tuning", name, "done", flush=True)
    np.savez_compressed(OUT / "pid_lqr.npz", **out)


POLICIES = {
    # name: (hidden_sizes, activation, trained by imitation?, output bounds or None)
    "relu_64_64": ([64, 64], "relu", True, None),
    "tanh_32": ([32], "tanh", True, None),
    "elu_48_40": ([48, 40], "elu", False, None),
    "leaky_20_20_20": ([20, 20, 20], "leaky_relu", False,
                       {"thrust": (2.0, 18.0), "roll_rate": (-2.0, 2.0), "pitch_rate": (-2.5, 2.5), "yaw_rate": (-1.0, 1.0)}),
}


def gen_policy():
    """§8(f).2: the reference's DeepTrackingPolicy (torch, CPU) on fixed weights -> policy.npz.
    Two of the networks are first fitted to the Riccati-LQR law (so the closed loop tracks and is
    contractive); the fit itself is not part of the fixture, only the resulting weights are."""
    import torch

    from quadcopter_tracking.controllers import RiccatiLQRController
    from quadcopter_tracking.controllers.deep_tracking_policy import DeepTrackingPolicy
    from quadcopter_tracking.env import EnvConfig, QuadcopterEnv
    from quadcopter_tracking.utils.metrics import SuccessCriteria

    torch.set_num_threads(1)
    out = {}
    teacher = RiccatiLQRController(config={})
    K = teacher.K
    rng = np.random.default_rng(5)

    def lqr_labels(feat):
        u = feat[:, 0:6] @ K.T
        return np.stack([np.clip(9.81 + u[:, 0], 0, 20), np.clip(u[:, 1], -3, 3), np.clip(u[:, 2], -3, 3),
                         np.clip(u[:, 3], -3, 3)], axis=1)

    def sample_features(n):
        f = rng.standard_normal((n, 18)) * np.array([1.0] * 3 + [1.5] * 3 + [0.2] * 3 + [0.5] * 3 + [3.0] * 3 + [1.0] * 3)
        return f

    for name, (hidden, act, trained, bounds) in POLICIES.items():
        torch.manual_seed(hash(name) % 1000 if False else len(name) * 7 + 1)
        cfg = {"hidden_sizes": hidden, "activation": act}
        if bounds:
            cfg["output_bounds"] = bounds
        pol = DeepTrackingPolicy(config=cfg, device="cpu")
        if trained:
            opt = torch.optim.Adam(pol.network.parameters(), lr=2e-3)
            pol.network.train()
            for it in range(4000):
                f = sample_features(512)
                x = torch.tensor(f, dtype=torch.float32)
                y = torch.tensor(lqr_labels(f), dtype=torch.float32)
                loss = ((pol.network(x) - y) ** 2 * torch.tensor([1.0, 30.0, 30.0, 1.0])).mean()
                opt.zero_grad()
                loss.backward()
                opt.step()
            pol.network.eval()
            print(f"  policy {name}: imitation fit loss {float(loss):.3e}", flush=True)
        lin = [m for m in pol.network.modules() if isinstance(m, torch.nn.Linear)]
        for l, m in enumerate(lin):
            out[f"{name}_W{l}"] = m.weight.detach().numpy().copy()
            out[f"{name}_b{l}"] = m.bias.detach().numpy().copy()
        # compute_action on hand-made observations
        obs_rows, act_rows = [], []
        for k in range(64):
            v = rng.standard_normal(22) * np.array([2] * 3 + [3] * 3 + [0.3] * 6 + [2] * 3 + [3] * 3 + [4] * 3 + [1])
            if k % 9 == 0:
                v *= 25.0          # saturates the sigmoid
            o = {"quadcopter": {"position": v[0:3], "velocity": v[3:6], "attitude": v[6:9], "angular_velocity": v[9:12]},
                 "target": {"position": v[12:15], "velocity": v[15:18], "acceleration": v[18:21]}, "time": float(v[21])}
            a = pol.compute_action(o)
            obs_rows.append(v)
            act_rows.append([a["thrust"], a["roll_rate"], a["pitch_rate"], a["yaw_rate"]])
        out[f"{name}_obs"], out[f"{name}_act"] = np.array(obs_rows).T, np.array(act_rows).T
        # closed loop through the reference's driver loop
        for motion, seeds, max_steps in (("stationary", [42, 43, 44], None if trained else 300),
                                         ("linear", [42, 43], 1200 if trained else 300)):
            cfg_env = EnvConfig()
            cfg_env.target.motion_type = motion
            recs = [run_reference_episode(cfg_env, pol, sd, SuccessCriteria(), max_steps, True) for sd in seeds]
            key = f"{name}_{motion}"
            out[f"{key}_seeds"] = np.array(seeds, dtype=np.uint64)
            out[f"{key}_max_steps"] = np.array(3002 if max_steps is None else max_steps)
            for fld in ("state0", "tparams", "metrics", "final_state"):
                out[f"{key}_{fld}"] = np.array([r[fld] for r in recs]).T
            out[f"{key}_samples"] = np.array([r["samples"] for r in recs])
            print("  policy rollout", key, "steps", [int(r["metrics"][0]) for r in recs], flush=True)
    # supervisor labels on the states the student visits (train.py:654-735), every 10th step
    name = "relu_64_64"
    pol = DeepTrackingPolicy(config={"hidden_sizes": [64, 64], "activation": "relu"}, device="cpu")
    lin = [m for m in pol.network.modules() if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for l, m in enumerate(lin):
            m.weight.copy_(torch.tensor(out[f"{name}_W{l}"]))
            m.bias.copy_(torch.tensor(out[f"{name}_b{l}"]))
    from quadcopter_tracking.controllers import PIDController
    for sup_name, sup in (("riccati", RiccatiLQRController(config={})),
                          ("pid", PIDController(config={"ki_pos": [0.002, 0.002, 0.5], "integral_limit": 0.4}))):
        cfg_env = EnvConfig()
        cfg_env.target.motion_type = "linear"
        env = QuadcopterEnv(config=cfg_env)
        feats, labs = [], []
        for sd in (42, 43):
            obs = env.reset(seed=sd)
            sup.reset()
            f_ep, l_ep = [], []
            for step in range(400):
                a_sup = sup.compute_action(obs)
                if step % 10 == 0:
                    f_ep.append(pol._extract_features(obs))
                    l_ep.append(np.array([a_sup["thrust"], a_sup["roll_rate"], a_sup["pitch_rate"], a_sup["yaw_rate"]],
                                         dtype=np.float32))
                obs, _, done, _ = env.step(pol.compute_action(obs))
                if done:
                    break
            feats.append(f_ep), labs.append(l_ep)
        out[f"dagger_{sup_name}_features"] = np.array(feats).transpose(1, 0, 2)   # [slots, E, 18]
        out[f"dagger_{sup_name}_labels"] = np.array(labs).transpose(1, 0, 2)      # [slots, E, 4]
    np.savez_compressed(OUT / "policy.npz", **out)


TRAINER_CASE = dict(controller="deep", epochs=3, episodes_per_epoch=3, max_steps_per_episode=80, batch_size=16,
                    learning_rate=0.002, hidden_sizes=[32, 32], activation="tanh", training_mode="imitation",
                    supervisor_controller="riccati_lqr", imitation_weight=1.0, tracking_weight=0.05, env_seed=123,
                    target_motion_type="circular", episode_length=30.0, device="cpu", checkpoint_interval=1000,
                    save_best=False)


def gen_trainer():
    """§8(f).2: the reference's imitation trainer (train.py:654-862) on a small run — three epochs of three episodes,
    one optimiser step after EVERY episode on batch_size steps drawn with the global numpy generator — recorded as
    initial weights, per-epoch loss components / rollout statistics and final weights.  The drop-in's trainer in its
    exact update order (Trainer(update_schedule="episode")) is held to this curve in tests/test_train_gpu.py."""
    import torch

    from quadcopter_tracking.train import Trainer, TrainingConfig

    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        torch.manual_seed(0)
        np.random.seed(0)
        cfg = TrainingConfig(**TRAINER_CASE, checkpoint_dir=os.path.join(tmp, "ck"), log_dir=os.path.join(tmp, "logs"))
        tr = Trainer(cfg)
        lin = [m for m in tr.controller.network.modules() if isinstance(m, torch.nn.Linear)]
        for i, l in enumerate(lin):
            out[f"W0_{i}"] = l.weight.detach().numpy().copy()
            out[f"b0_{i}"] = l.bias.detach().numpy().copy()
        tr.train()
        keys = ("total", "imitation", "tracking_total", "mean_reward", "mean_on_target_ratio", "mean_tracking_error")
        for k in keys:
            out[k] = np.array([e[k] for e in tr.experiment_log], dtype=np.float64)
        for i, l in enumerate(lin):
            out[f"W1_{i}"] = l.weight.detach().numpy().copy()
            out[f"b1_{i}"] = l.bias.detach().numpy().copy()
    np.savez_compressed(OUT / "trainer.npz", **out)
    print("trainer.npz", {k: out[k] for k in ("total", "imitation", "mean_tracking_error")})


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    _import_reference()
    cwd = os.getcwd()
    every = (gen_reset, gen_dare, gen_step_edges, gen_controller, gen_eval_summary, gen_tuning, gen_rollouts, gen_pid_lqr, gen_policy, gen_trainer)
    chosen = [fn for fn in every if fn.__name__ in sys.argv[1:]] or every   # e.g. `gen_golden.py gen_pid_lqr`
    for fn in chosen:
        print(fn.__name__, flush=True)
        fn()
        os.chdir(cwd)
    import numpy, scipy
    (OUT / "VERSIONS.txt").write_text(
        f"python {sys.version.split()[0]}\nnumpy {numpy.__version__}\nscipy {scipy.__version__}\n"
        "reference quadcopter_tracking 0.5.0 (/root/reference)\n")


if __name__ == "__main__":
    main()
