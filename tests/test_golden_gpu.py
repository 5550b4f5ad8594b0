"""GPU: the CUDA path (through the C ABI) against fixtures produced by the reference itself."""
import numpy as np
import pytest

import golden_cases as gc
from lqr_quadcopter_test_b200 import params as P
from lqr_quadcopter_test_b200.config import EnvConfig

pytestmark = pytest.mark.gpu

MOTIONS = ["stationary", "linear", "circular", "sinusoidal", "figure8"]


@pytest.fixture(scope="module")
def gb(gpu_ctx):
    import gpu_backend

    return gpu_backend


@pytest.mark.parametrize("motion", MOTIONS)
def test_reset_matches_reference(golden, gb, motion):
    """A1.  The generator is bit-exact; the only non-identical operations are device sin/cos in
    the circular / sinusoidal start position (<= 1 ulp of a 2-metre offset)."""
    g = golden["reset"]
    env = EnvConfig()
    env.target.motion_type = motion
    prm = P.make_params(env)
    state0, tparams = gb.reset(prm, g["seeds"])
    np.testing.assert_array_equal(tparams, g[f"{motion}_tparams"])
    if motion in ("stationary", "linear", "figure8"):
        np.testing.assert_array_equal(state0, g[f"{motion}_state0"])
    else:
        np.testing.assert_allclose(state0, g[f"{motion}_state0"], rtol=0, atol=5e-16)


@pytest.mark.parametrize("motion", ["linear", "sinusoidal"])
def test_reset_without_seed_continues_the_streams(golden, gb, motion):
    """quadcopter_env.py:122-130: reset(seed=None) keeps drawing from the existing generators."""
    import torch

    g = golden["reset"]
    env = EnvConfig()
    env.target.motion_type = motion
    prm = P.make_params(env)
    draws = torch.zeros((2, 1), dtype=torch.int64, device=gb.ctx().device)
    for k in range(3):
        state0, tparams = gb.reset(prm, [7], draws)
        np.testing.assert_array_equal(tparams[:, 0], g[f"{motion}_noseed_tparams"][:, k])
        np.testing.assert_allclose(state0[:, 0], g[f"{motion}_noseed_state0"][:, k], rtol=0, atol=5e-16)


def test_rollouts_match_reference(golden, gb):
    """A2-A10 fused: metrics, final state and sampled trajectory of every fixture case."""
    g = golden["rollouts"]
    names = gc.case_names(g)
    for name in names:
        prm, max_steps, chain, tol, _cfg, _env, _crit = gc.case(name)
        cs = np.zeros((4, 1)) if chain > 1 else None
        out = gb.rollout(prm, g[name + "_state0"], g[name + "_tparams"], g[name + "_K"], max_steps=max_steps,
                         chain=chain, ctrl_state=cs, traj_stride=250)
        gc.assert_rollout_matches(np, out, g, name, tol)
        samples = g[name + "_samples"]
        n_s = samples.shape[1]
        got = np.transpose(out["traj"][:n_s, P.TRAJ_STATE:P.TRAJ_STATE + 12, :], (2, 0, 1))
        np.testing.assert_allclose(got, samples, rtol=0, atol=tol, err_msg=name)
        if chain > 1:
            np.testing.assert_allclose(cs[0:3, 0], g[name + "_integral"][:, -1], rtol=0, atol=tol)
        elif name.startswith("lqi"):
            cs1 = np.zeros((4, g[name + "_integral"].shape[1]))
            gb.rollout(prm, g[name + "_state0"], g[name + "_tparams"], g[name + "_K"], max_steps=max_steps,
                       ctrl_state=cs1)
            np.testing.assert_allclose(cs1[0:3], g[name + "_integral"], rtol=0, atol=tol, err_msg=name)


def test_step_edge_cases_match_reference(golden, gb):
    """A2-A8 one step at a time: action repair, constraints, wrap, termination."""
    g = golden["step_edges"]
    for i in range(len(g["seed"])):
        motion = MOTIONS[int(g["motion"][i])]
        env = EnvConfig()
        env.target.motion_type = motion
        prm = P.make_params(env)
        state = g["state"][i].reshape(12, 1).copy()
        time = np.array([g["time_before"][i]])
        counters = np.zeros((3, 1), dtype=np.int64)
        counters[0, 0] = g["n_pre"][i]
        out = gb.step(prm, state, time, g["tparams"][i].reshape(12, 1), counters, action=g["action"][i].reshape(4, 1))
        np.testing.assert_allclose(state[:, 0], g["next_state"][i], rtol=0, atol=1e-12, equal_nan=True, err_msg=str(i))
        obs_tol = np.full(22, 1e-12)
        if motion == "figure8":
            obs_tol[18:21] = 5e-3  # finite-difference acceleration: cancellation amplifies 1 ulp to ~1e-4
        assert np.all((np.abs(out["obs"][:, 0] - g["obs"][i]) <= obs_tol) | (np.isnan(out["obs"][:, 0]) & np.isnan(g["obs"][i]))), i
        assert int(out["done"][0]) == int(g["term"][i]), i
        if np.isfinite(g["reward"][i]):
            assert abs(out["reward"][0] - g["reward"][i]) < 1e-12
            assert out["info"][1, 0] == g["on_target"][i]
        assert counters[2, 0] == int(g["violations"][i]), i
        assert counters[0, 0] == int(g["step"][i])


@pytest.mark.parametrize("name", ["lqr", "lqr_ff", "lqi"])
def test_controller_matches_reference(golden, gb, name):
    """A10 incl. feedforward clamps and the LQI integrator with anti-windup / zero threshold."""
    g = golden["controller"]
    kw = {}
    if name == "lqr_ff":
        kw = dict(feedforward_enabled=True, ff_velocity_gain=[0.3, 0.2, 0.1], ff_acceleration_gain=[0.1, 0.2, 0.3],
                  ff_max_velocity=2.0, ff_max_acceleration=1.0)
    if name == "lqi":
        kw = dict(use_lqi=True, integral_limit=0.05, integral_zero_threshold=0.2)
    prm = P.make_params(EnvConfig(), gc.controller_ns(**kw))
    obs, K = g[f"{name}_obs"], g[f"{name}_K"]
    integ = np.zeros((3, 1))
    sat = np.zeros(1, dtype=np.int64)
    for k in range(obs.shape[1]):
        act = gb.compute_action(prm, obs[:, k:k + 1], K, integral=integ if name == "lqi" else None, saturation=sat)
        np.testing.assert_allclose(act[:, 0], g[f"{name}_act"][:, k], rtol=0, atol=1e-12)
        if name == "lqi":
            np.testing.assert_allclose(integ[:, 0], g[f"{name}_integral"][:, k], rtol=0, atol=1e-15)
        assert sat[0] == g[f"{name}_sat"][k]


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_dare_matches_reference(golden, gb):
    """A11: K within 1e-9 (norm-wise relative) of the reference's scipy path, P likewise."""
    g = golden["dare"]
    for name in ("lqr_default", "lqi_default", "lqi_zero_qint"):
        P_, K, iters, info = gb.dare(g[name + "_A"], g[name + "_B"], g[name + "_Q"], g[name + "_R"])
        assert info[0] == 0, name
        assert _rel(K[0], g[name + "_K"]) < 1e-9, (name, _rel(K[0], g[name + "_K"]))
        assert _rel(P_[0], g[name + "_P"]) < 1e-9, name
    for tag, ref in (("sweep6", "lqr_default"), ("sweep9", "lqi_default")):
        q, r = g[tag + "_q"], g[tag + "_r"]
        Q = np.stack([np.diag(x) for x in q])
        R = np.stack([np.diag(x) for x in r])
        P_, K, iters, info = gb.dare(g[ref + "_A"], g[ref + "_B"], Q, R)
        assert not info.any()
        worst = max(_rel(K[i], g[tag + "_K"][i]) for i in range(len(q)))
        assert worst < 1e-9, (tag, worst)
        assert max(_rel(P_[i], g[tag + "_P"][i]) for i in range(len(q))) < 1e-9
        assert 8 <= iters.min() and iters.max() <= 30
    P_, K, _, info = gb.dare(g["lqr_default_A"], g["lqr_default_B"], g["dense6_Q"], g["dense6_R"])
    assert not info.any()
    assert max(_rel(K[i], g["dense6_K"][i]) for i in range(len(K))) < 1e-9
    for n, m in ((2, 1), (3, 2), (4, 4), (5, 3), (7, 2), (8, 8), (12, 4), (15, 4), (16, 8)):
        pre = f"gen{n}x{m}_"
        P_, K, _, info = gb.dare(g[pre + "A"], g[pre + "B"], g[pre + "Q"], g[pre + "R"])
        assert not info.any(), (n, m, info)
        assert max(_rel(K[i], g[pre + "K"][i]) for i in range(len(K))) < 1e-9, (n, m)
        assert max(_rel(P_[i], g[pre + "P"][i]) for i in range(len(K))) < 1e-9, (n, m)
    for k in ("bad_Q_negative", "bad_Q_asym", "bad_R_singular", "ok_Q_tiny_negative"):
        _, _, _, info = gb.dare(g["lqr_default_A"], g["lqr_default_B"], g[k + "_Q"], g[k + "_R"])
        assert (info[0] != 3) == bool(g[k + "_accepted"]), k


def test_evaluator_config1_matches_reference(golden, tmp_path):
    """BASELINE config 1 through the public drop-in API (eval.py:169-231)."""
    from lqr_quadcopter_test_b200 import EnvConfig as EC, Evaluator, SuccessCriteria, load_controller

    g = golden["eval_summary"]
    ev = Evaluator(load_controller("riccati_lqr", config={}), EC(), SuccessCriteria(), output_dir=tmp_path)
    s = ev.evaluate(num_episodes=10, base_seed=42, verbose=False)
    ref = g["config1_summary"]
    assert (s.total_episodes, s.successful_episodes) == (10, 10)
    assert s.success_rate == ref[2]
    assert abs(s.mean_on_target_ratio - ref[3]) < 1e-14 and abs(s.std_on_target_ratio - ref[4]) < 1e-13
    assert abs(s.mean_tracking_error - ref[5]) < 1e-11 and abs(s.std_tracking_error - ref[6]) < 1e-11
    assert abs(s.mean_control_effort - ref[7]) < 1e-10
    assert (s.best_episode_idx, s.worst_episode_idx, s.meets_criteria) == (int(ref[8]), int(ref[9]), bool(ref[10]))
    np.testing.assert_array_equal([m.on_target_ratio for m in s.episode_metrics], g["config1_ratios"])
    np.testing.assert_array_equal([i["on_target_ratio"] for i in ev.episode_info_list], g["config1_env_ratios"])
    assert all(i["termination_reason"] == "time_limit" and i["success"] for i in ev.episode_info_list)
    assert len(ev.episode_data_list) == 10 and len(ev.episode_data_list[0]) == 3000
    for k in (0, 1, 2999):
        rec, r = g[f"config1_rec{k}"], ev.episode_data_list[0][k]
        flat = [r["time"], *r["quadcopter_position"], *r["quadcopter_velocity"], *r["target_position"],
                *r["target_velocity"], *r["action"], r["reward"], r["tracking_error"], float(r["on_target"])]
        np.testing.assert_allclose(flat, rec, rtol=0, atol=1e-11)
    saved = ev.save_report(s, "t")
    assert saved["metrics"].exists() and saved["report"].exists()


def test_tuner_matches_reference(golden, tmp_path):
    """A14 end to end: host candidate sampling -> batched DARE -> rollout -> scores."""
    from lqr_quadcopter_test_b200 import ControllerTuner, GainSearchSpace, TuningConfig

    g = golden["tuning"]
    base = dict(q_pos_range=([5e-5, 5e-5, 10.0], [5e-4, 5e-4, 25.0]), q_vel_range=([1e-3, 1e-3, 2.0], [1e-2, 1e-2, 8.0]),
                r_controls_range=([0.5] * 4, [2.0] * 4))
    for key in ("riccati_stationary", "riccati_linear", "riccati_lqi_stationary", "riccati_lqi_linear"):
        kw = dict(base)
        if "lqi" in key:
            kw["q_int_range"] = ([1e-6, 1e-6, 1e-4], [1e-5, 1e-5, 1e-2])
        tc = TuningConfig(controller_type="riccati_lqr", search_space=GainSearchSpace(**kw), strategy="random",
                          max_iterations=5, evaluation_episodes=3, evaluation_horizon=400, seed=42,
                          target_motion_type=key.split("_")[-1], episode_length=4.0, output_dir=str(tmp_path / key))
        res = ControllerTuner(tc).tune()
        scores = np.array([r["score"] for r in res.all_results])
        np.testing.assert_allclose(scores, g[key + "_score"], rtol=0, atol=1e-10, err_msg=key)
        met = np.array([[r["metrics"]["mean_on_target_ratio"], r["metrics"]["mean_tracking_error"],
                         r["metrics"]["success_rate"]] for r in res.all_results])
        np.testing.assert_allclose(met, g[key + "_metrics"], rtol=0, atol=1e-10, err_msg=key)
        assert abs(res.best_score - float(g[key + "_best_score"])) < 1e-10
        assert res.best_config == res.all_results[int(g[key + "_best"])]["config"]
        assert res.iterations_completed == 5 and list((tmp_path / key).glob("*_results.json"))


# ---- §8(f).1: PID and heuristic LQR in the same kernels --------------------------------------------
def test_pid_and_heuristic_lqr_rollouts_match_reference(golden, gb):
    g = golden["pid_lqr"]
    for name in gc.PID_LQR_CASES:
        prm, max_steps, chain, tol, _cfg, _env = gc.pid_lqr_case(name)
        pid = name.startswith("pid")
        cs = np.zeros((4, 1)) if chain > 1 else None
        if cs is not None:
            cs[3] = np.nan
        out = gb.rollout(prm, g[name + "_state0"], g[name + "_tparams"], g[name + "_K"], max_steps=max_steps,
                         chain=chain, ctrl_state=cs, traj_stride=250)
        gc.assert_rollout_matches(np, out, g, name, tol)
        samples = g[name + "_samples"]
        n_s = samples.shape[1]
        got = np.transpose(out["traj"][:n_s, P.TRAJ_STATE:P.TRAJ_STATE + 12, :], (2, 0, 1))
        np.testing.assert_allclose(got, samples, rtol=0, atol=tol, err_msg=name)
        if pid and chain > 1:
            np.testing.assert_allclose(cs[0:3, 0], g[name + "_integral"][:, -1], rtol=0, atol=tol)
            assert cs[3, 0] == float(g[name + "_last_time"])
        elif pid:
            cs1 = np.zeros((4, samples.shape[0]))
            gb.rollout(prm, g[name + "_state0"], g[name + "_tparams"], g[name + "_K"], max_steps=max_steps,
                       ctrl_state=cs1)
            np.testing.assert_allclose(cs1[0:3], g[name + "_integral"], rtol=0, atol=tol, err_msg=name)


@pytest.mark.parametrize("name", ["seq_pid", "seq_pidki", "seq_pidff"])
def test_pid_controller_matches_reference(golden, gb, name):
    """controllers/__init__.py:243-379 one call at a time: dt from observation times (forward,
    equal and backwards), integral clamp, feedforward clamps."""
    g = golden["pid_lqr"]
    kw = {"seq_pid": dict(integral_limit=0.0), "seq_pidki": dict(integral_limit=0.05),
          "seq_pidff": dict(integral_limit=2.0, **gc.PID_FF)}[name]
    prm = P.make_params(EnvConfig(), gc.controller_ns(name="pid", **kw))
    obs, K = g[f"{name}_obs"], g[f"{name}_K"]
    state = np.zeros((4, 1))
    state[3] = np.nan
    for k in range(obs.shape[1]):
        act = gb.compute_action(prm, obs[:, k:k + 1], K, integral=state)
        np.testing.assert_allclose(act[:, 0], g[f"{name}_act"][:, k], rtol=0, atol=1e-12)
        np.testing.assert_allclose(state[0:3, 0], g[f"{name}_integral"][:, k], rtol=0, atol=1e-15)
        assert state[3, 0] == g[f"{name}_last_time"][k]


def test_pid_dropin_class_matches_reference_sequence(golden, gpu_ctx):
    """PIDController.compute_action(obs dict) through the public class, state kept in the object."""
    from lqr_quadcopter_test_b200 import PIDController

    g = golden["pid_lqr"]
    c = PIDController(config=dict(gc.PID_LQR_CASES["pidff_circular"][3], ki_pos=0.2, integral_limit=2.0, kp_pos=1.5,
                                  kd_pos=[0.5, 0.6, 0.7]))
    obs = g["seq_pidff_obs"]
    for k in range(obs.shape[1]):
        v = obs[:, k]
        a = c.compute_action({"quadcopter": {"position": v[0:3], "velocity": v[3:6], "attitude": v[6:9],
                                             "angular_velocity": v[9:12]},
                              "target": {"position": v[12:15], "velocity": v[15:18], "acceleration": v[18:21]},
                              "time": float(v[21])})
        np.testing.assert_allclose([a["thrust"], a["roll_rate"], a["pitch_rate"], a["yaw_rate"]],
                                   g["seq_pidff_act"][:, k], rtol=0, atol=1e-12)
        np.testing.assert_allclose(c.integral_error, g["seq_pidff_integral"][:, k], rtol=0, atol=1e-15)
        assert c._last_time == g["seq_pidff_last_time"][k]
    c.reset()
    assert c._last_time is None and not c.integral_error.any()


@pytest.mark.parametrize("ctype", ["pid", "lqr"])
def test_evaluator_pid_and_lqr_match_reference(golden, tmp_path, ctype):
    """`eval --controller pid|lqr --motion-type linear` (Makefile compare-controllers)."""
    from lqr_quadcopter_test_b200 import EnvConfig as EC, Evaluator, SuccessCriteria, load_controller

    g = golden["pid_lqr"]
    cfg = EC()
    cfg.target.motion_type = "linear"
    ev = Evaluator(load_controller(ctype, config={}), cfg, SuccessCriteria(), output_dir=tmp_path)
    s = ev.evaluate(num_episodes=4, base_seed=42, verbose=False)
    ref = g[f"eval_{ctype}_summary"]
    assert (s.total_episodes, s.successful_episodes, s.success_rate) == (int(ref[0]), int(ref[1]), ref[2])
    assert abs(s.mean_on_target_ratio - ref[3]) < 1e-14 and abs(s.std_on_target_ratio - ref[4]) < 1e-13
    assert abs(s.mean_tracking_error - ref[5]) < 1e-10 and abs(s.std_tracking_error - ref[6]) < 1e-10
    assert abs(s.mean_control_effort - ref[7]) < 1e-9
    assert (s.best_episode_idx, s.worst_episode_idx, s.meets_criteria) == (int(ref[8]), int(ref[9]), bool(ref[10]))
    np.testing.assert_array_equal([m.on_target_ratio for m in s.episode_metrics], g[f"eval_{ctype}_ratios"])
    np.testing.assert_allclose([m.mean_tracking_error for m in s.episode_metrics], g[f"eval_{ctype}_mean_err"],
                               rtol=0, atol=1e-10)


def test_pid_and_lqr_tuner_matches_reference(golden, tmp_path):
    """tuning_pid_linear.yaml / tuning_lqr_linear.yaml (shortened): sampling order, rollout, scores."""
    from lqr_quadcopter_test_b200 import ControllerTuner, GainSearchSpace, TuningConfig

    g = golden["pid_lqr"]
    spaces = {
        "pid": dict(kp_pos_range=([0.005, 0.005, 2.0], [0.05, 0.05, 6.0]), kd_pos_range=([0.02, 0.02, 1.0], [0.15, 0.15, 3.0])),
        "pidki": dict(kp_pos_range=([0.005, 0.005, 2.0], [0.05, 0.05, 6.0]), ki_pos_range=([0.0, 0.0, 0.0], [0.005, 0.005, 0.5]),
                      kd_pos_range=([0.02, 0.02, 1.0], [0.15, 0.15, 3.0])),
        "lqr": dict(q_pos_range=([5e-5, 5e-5, 10.0], [5e-4, 5e-4, 25.0]), q_vel_range=([1e-3, 1e-3, 2.0], [1e-2, 1e-2, 8.0])),
    }
    for key, kw in spaces.items():
        tc = TuningConfig(controller_type="lqr" if key == "lqr" else "pid", search_space=GainSearchSpace(**kw),
                          strategy="random", max_iterations=5, evaluation_episodes=3, evaluation_horizon=400, seed=42,
                          target_motion_type="linear", episode_length=4.0, output_dir=str(tmp_path / key))
        res = ControllerTuner(tc).tune()
        keys = ("kp_pos", "ki_pos", "kd_pos") if key != "lqr" else ("q_pos", "q_vel")
        cfgs = np.array([sum((r["config"].get(k, [0.0, 0.0, 0.0]) for k in keys), []) for r in res.all_results])
        np.testing.assert_array_equal(cfgs, g[f"tune_{key}_cfg"])
        scores = np.array([r["score"] for r in res.all_results])
        np.testing.assert_allclose(scores, g[f"tune_{key}_score"], rtol=0, atol=1e-10, err_msg=key)
        met = np.array([[r["metrics"]["mean_on_target_ratio"], r["metrics"]["mean_tracking_error"],
                         r["metrics"]["success_rate"]] for r in res.all_results])
        np.testing.assert_allclose(met, g[f"tune_{key}_metrics"], rtol=0, atol=1e-10, err_msg=key)
        assert abs(res.best_score - float(g[f"tune_{key}_best_score"])) < 1e-10


# ---- §8(f).2: deep tracking policy in the loop ------------------------------------------------------
@pytest.mark.parametrize("name", list(gc.POLICIES))
def test_policy_network_matches_reference(golden, gb, name):
    """qt_policy_action_batched / qt_rollout_policy against torch's fp32 network driven by the
    reference's own loop (fixture), incl. widths that need padding and custom output bounds."""
    g = golden["policy"]
    pol = gb.make_policy(g, name)
    act = gb.policy_action(pol, g[name + "_obs"])
    np.testing.assert_allclose(act, g[name + "_act"], rtol=0, atol=gc.TOL_POLICY_ACTION)
    v = g[name + "_obs"][:, 3]
    a = pol.compute_action({"quadcopter": {"position": v[0:3], "velocity": v[3:6], "attitude": v[6:9],
                                           "angular_velocity": v[9:12]},
                            "target": {"position": v[12:15], "velocity": v[15:18], "acceleration": v[18:21]},
                            "time": float(v[21])})
    np.testing.assert_allclose([a["thrust"], a["roll_rate"], a["pitch_rate"], a["yaw_rate"]], g[name + "_act"][:, 3],
                               rtol=0, atol=gc.TOL_POLICY_ACTION)
    for motion in ("stationary", "linear"):
        key = f"{name}_{motion}"
        env = EnvConfig()
        env.target.motion_type = motion
        out = gb.rollout_policy(P.make_params(env), g[key + "_state0"], g[key + "_tparams"], pol,
                                max_steps=int(g[key + "_max_steps"]), traj_stride=250)
        gc.assert_policy_rollout_matches(np, out, g, key)
        # the tensor-core evaluation (two hidden layers of at most 64 units; other shapes ignore the flag and stay on
        # the FP32 pipe) is held to the same reference fixture with the same tolerances
        tcp = P.make_params(env)
        tcp.debug_flags = P.FLAG_POLICY_TENSOR_CORES
        out_tc = gb.rollout_policy(tcp, g[key + "_state0"], g[key + "_tparams"], pol,
                                   max_steps=int(g[key + "_max_steps"]), traj_stride=250)
        gc.assert_policy_rollout_matches(np, out_tc, g, key)


@pytest.mark.parametrize("sup", ["riccati", "pid"])
def test_supervisor_labels_on_student_states_match_reference(golden, gb, sup):
    """train.py:654-735 data collection through the public API (teacher_dataset(student=...))."""
    from lqr_quadcopter_test_b200 import PIDController, RiccatiLQRController, teacher_dataset

    g = golden["policy"]
    pol = gb.make_policy(g, "relu_64_64")
    env = EnvConfig()
    env.target.motion_type = "linear"
    teacher = (PIDController(config={"ki_pos": [0.002, 0.002, 0.5], "integral_limit": 0.4}) if sup == "pid"
               else RiccatiLQRController(config={}))
    feats, labels, _ = teacher_dataset(teacher, env, num_episodes=2, base_seed=42, stride=10, max_steps=400,
                                       flatten=False, student=pol)
    np.testing.assert_allclose(feats.cpu().numpy(), g[f"dagger_{sup}_features"], rtol=0, atol=gc.TOL_POLICY_STATE)
    np.testing.assert_allclose(labels.cpu().numpy(), g[f"dagger_{sup}_labels"], rtol=0, atol=gc.TOL_POLICY_STATE)


def test_deep_policy_evaluator_and_checkpoint_round_trip(golden, gb, tmp_path):
    """`eval --controller deep --checkpoint ...`: save in the reference's checkpoint format, load through
    load_controller, evaluate; per-episode metrics equal the fixture's."""
    from lqr_quadcopter_test_b200 import EnvConfig as EC, Evaluator, SuccessCriteria, load_controller

    g = golden["policy"]
    pol = gb.make_policy(g, "relu_64_64")
    ck = tmp_path / "policy.pt"
    pol.save_checkpoint(ck, metadata={"epoch": 3})
    loaded = load_controller("deep", checkpoint_path=ck, config={"hidden_sizes": [64, 64], "activation": "relu"})
    ev = Evaluator(loaded, EC(), SuccessCriteria(), output_dir=tmp_path)
    s = ev.evaluate(num_episodes=3, base_seed=42, verbose=False)
    G = g["relu_64_64_stationary_metrics"]
    ratios = np.array([m.on_target_ratio for m in s.episode_metrics])
    assert np.abs(ratios * 3000 - G[P.M_MET_ON_TARGET]).max() <= 2
    np.testing.assert_allclose([m.mean_tracking_error for m in s.episode_metrics], G[P.M_SUM_ERR] / G[P.M_STEPS],
                               rtol=2e-5)
    assert s.total_episodes == 3 and len(ev.episode_data_list[0]) == 3000
    with pytest.raises(ValueError, match="Hidden sizes mismatch"):
        load_controller("deep", checkpoint_path=ck, config={"hidden_sizes": [32]})
