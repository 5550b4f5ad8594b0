"""CPU: pin the oracle (oracle/) against fixtures produced by the reference itself."""
import numpy as np
import pytest

import oracle
from lqr_quadcopter_test_b200 import params as P
from lqr_quadcopter_test_b200.config import EnvConfig

import golden_cases as gc

MOTIONS = ["stationary", "linear", "circular", "sinusoidal", "figure8"]


def test_numpy_sum_order_is_restated_exactly():
    rng = np.random.default_rng(0)
    for n in list(range(1, 300)) + [1000, 2999, 3000, 3001, 5000]:
        a = rng.standard_normal(n) * rng.uniform(0.1, 10)
        assert oracle.np_sum(a) == float(np.sum(a)), n


@pytest.mark.parametrize("motion", MOTIONS)
def test_reset_matches_reference(golden, motion):
    g = golden["reset"]
    env = EnvConfig()
    env.target.motion_type = motion
    prm = P.make_params(env)
    state0, tparams = oracle.reset_reference(prm, g["seeds"])
    np.testing.assert_array_equal(state0, g[f"{motion}_state0"])
    np.testing.assert_array_equal(tparams, g[f"{motion}_tparams"])


def test_rollouts_match_reference(golden):
    g = golden["rollouts"]
    names = gc.case_names(g)
    assert len(names) >= 17
    for name in names:
        prm, max_steps, chain, _tol, _cfg, _env, _crit = gc.case(name)
        cs = np.zeros((4, 1)) if chain > 1 else None
        out = oracle.rollout(prm, g[name + "_state0"], g[name + "_tparams"], g[name + "_K"],
                             max_steps=max_steps, chain=chain, ctrl_state=cs, traj_stride=250)
        # the C restatement tracks the Python reference to round-off: 1e-11 even where the loop is unstable
        gc.assert_rollout_matches(np, out, g, name, tol=2e-11, sum_rtol=1e-13)
        # sampled trajectory (every 250 steps)
        samples = g[name + "_samples"]            # [E, n_samples, 12]
        n_s = samples.shape[1]
        got = out["traj"][:n_s, P.TRAJ_STATE:P.TRAJ_STATE + 12, :]  # [slots, 12, E]
        np.testing.assert_allclose(np.transpose(got, (2, 0, 1)), samples, rtol=0, atol=2e-11, err_msg=name)
        if chain > 1:
            np.testing.assert_allclose(cs[0:3, 0], g[name + "_integral"][:, -1], rtol=0, atol=1e-12)
        integ = g[name + "_integral"]
        if name.startswith("lqi") and chain == 1:
            cs1 = np.zeros((4, integ.shape[1]))
            oracle.rollout(prm, g[name + "_state0"], g[name + "_tparams"], g[name + "_K"], max_steps=max_steps,
                           ctrl_state=cs1)
            np.testing.assert_allclose(cs1[0:3], integ, rtol=0, atol=1e-11, err_msg=name + " integral")


def test_step_edge_cases_match_reference(golden):
    g = golden["step_edges"]
    n = len(g["seed"])
    for i in range(n):
        env = EnvConfig()
        env.target.motion_type = MOTIONS[int(g["motion"][i])]
        prm = P.make_params(env)
        state = g["state"][i].reshape(12, 1).copy()
        time = np.array([g["time_before"][i]])
        counters = np.zeros((3, 1), dtype=np.int64)
        counters[0, 0] = g["n_pre"][i]
        out = oracle.step(prm, state, time, g["tparams"][i].reshape(12, 1), counters,
                          action=g["action"][i].reshape(4, 1))
        np.testing.assert_allclose(state[:, 0], g["next_state"][i], rtol=0, atol=1e-13, equal_nan=True, err_msg=str(i))
        np.testing.assert_allclose(out["obs"][:, 0], g["obs"][i], rtol=0, atol=2e-9 if env.target.motion_type == "figure8" else 1e-13,
                                   equal_nan=True, err_msg=str(i))
        assert int(out["done"][0]) == int(g["term"][i]), i
        if np.isfinite(g["reward"][i]):
            assert abs(out["reward"][0] - g["reward"][i]) < 1e-13
            assert out["info"][1, 0] == g["on_target"][i]
        assert counters[2, 0] == int(g["violations"][i]), i
        assert counters[0, 0] == int(g["step"][i])


@pytest.mark.parametrize("name", ["lqr", "lqr_ff", "lqi"])
def test_controller_matches_reference(golden, name):
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
        act = oracle.compute_action(prm, obs[:, k:k + 1], K, integral=integ if name == "lqi" else None, saturation=sat)
        np.testing.assert_allclose(act[:, 0], g[f"{name}_act"][:, k], rtol=0, atol=1e-13)
        if name == "lqi":
            np.testing.assert_allclose(integ[:, 0], g[f"{name}_integral"][:, k], rtol=0, atol=1e-15)
        assert sat[0] == g[f"{name}_sat"][k]


def test_dare_restatement_matches_reference(golden):
    g = golden["dare"]
    for name in ("lqr_default", "lqi_default", "lqi_zero_qint"):
        lqi = name.startswith("lqi")
        A, B = oracle.build_systems_reference(lqi=lqi)
        np.testing.assert_array_equal(A, g[name + "_A"])
        np.testing.assert_array_equal(B, g[name + "_B"])
        P_, K = oracle.solve_dare_reference(A, B, g[name + "_Q"], g[name + "_R"])
        np.testing.assert_allclose(K, g[name + "_K"], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(P_, g[name + "_P"], rtol=1e-12, atol=1e-14)
    for k in ("bad_Q_negative", "bad_Q_asym", "bad_R_singular", "ok_Q_tiny_negative"):
        assert (oracle.validate_reference(g[k + "_Q"], g[k + "_R"]) == 0) == bool(g[k + "_accepted"]), k
    # SURVEY.md §8(c) spot values
    assert abs(g["lqr_default_K"][0, 2] - 3.9313145185949483) < 1e-12
    assert abs(g["lqi_default_K"][0, 8] - 0.31067758261816364) < 1e-12


def test_tuner_scores_match_reference(golden):
    """tuning.py:846-928: gains from the reference's solver, rollout + aggregation by the oracle."""
    g = golden["tuning"]
    for key in ("riccati_stationary", "riccati_linear", "riccati_lqi_stationary", "riccati_lqi_linear"):
        lqi = "lqi" in key
        motion = key.split("_")[-1]
        env = EnvConfig()
        env.target.motion_type = motion
        env.simulation.max_episode_time = 4.0
        from lqr_quadcopter_test_b200.metrics import SuccessCriteria
        crit = SuccessCriteria(0.8, 4.0, 0.5)
        prm = P.make_params(env, gc.controller_ns(use_lqi=lqi), crit)
        A, B = oracle.build_systems_reference(lqi=lqi)
        q, r = g[key + "_q"], g[key + "_r"]
        C_, eps = q.shape[0], 3
        Ks = np.array([oracle.solve_dare_reference(A, B, np.diag(q[i]), np.diag(r[i]))[1] for i in range(C_)])
        s0, tp = oracle.reset_reference(prm, [42, 43, 44])
        out = oracle.rollout(prm, np.tile(s0, (1, C_)), np.tile(tp, (1, C_)), Ks, k_stride=eps, max_steps=400)
        cand = oracle.reduce_scores(out["metrics"], C_, eps)
        np.testing.assert_allclose(cand[0], g[key + "_score"], rtol=0, atol=1e-12, err_msg=key)
        np.testing.assert_allclose(cand[1:4].T, g[key + "_metrics"], rtol=0, atol=1e-12, err_msg=key)
        assert int(np.argmax(cand[0])) == int(g[key + "_best"])


def test_config1_summary_matches_reference(golden):
    """BASELINE config 1: riccati_lqr defaults, stationary, seeds 42..51."""
    g = golden["eval_summary"]
    prm = P.make_params(EnvConfig(), gc.controller_ns())
    s0, tp = oracle.reset_reference(prm, np.arange(42, 52))
    K = golden["dare"]["lqr_default_K"]
    out = oracle.rollout(prm, s0, tp, K, max_steps=3002, traj_stride=1)
    M = out["metrics"]
    ratios = M[P.M_MET_ON_TARGET] / M[P.M_STEPS]
    np.testing.assert_array_equal(ratios, g["config1_ratios"])
    np.testing.assert_array_equal(M[P.M_ENV_ON_TARGET] / M[P.M_STEPS], g["config1_env_ratios"])
    np.testing.assert_allclose(M[P.M_SUM_ERR] / M[P.M_STEPS], g["config1_mean_err"], rtol=1e-13)
    s = g["config1_summary"]
    assert s[1] == M[P.M_MET_SUCCESS].sum() == 10
    assert abs(ratios.mean() - s[3]) < 1e-15 and abs(ratios.std() - s[4]) < 1e-15
    assert int(np.argmax(ratios)) == int(s[8]) and int(np.argmin(ratios)) == int(s[9])
    # the per-step record (eval.py:142-159) of episode 0
    for k in (0, 1, 2999):
        rec = g[f"config1_rec{k}"]
        row = out["traj"][k, :, 0]
        np.testing.assert_allclose(row[P.TRAJ_TIME], rec[0], rtol=0, atol=1e-15)
        np.testing.assert_allclose(row[P.TRAJ_STATE:P.TRAJ_STATE + 6], rec[1:7], rtol=0, atol=1e-13)
        np.testing.assert_allclose(row[P.TRAJ_TARGET_P:P.TRAJ_TARGET_P + 6], rec[7:13], rtol=0, atol=1e-13)
        np.testing.assert_allclose(row[P.TRAJ_ACTION:P.TRAJ_ACTION + 4], rec[13:17], rtol=0, atol=1e-13)
        np.testing.assert_allclose(row[P.TRAJ_ERR], rec[18], rtol=0, atol=1e-13)


# ---- §8(f).1: PID and heuristic LQR ----------------------------------------------------------------
def test_pid_and_heuristic_lqr_rollouts_match_reference(golden):
    g = golden["pid_lqr"]
    for name in gc.PID_LQR_CASES:
        prm, max_steps, chain, _tol, _cfg, _env = gc.pid_lqr_case(name)
        pid = name.startswith("pid")
        cs = np.zeros((4, 1)) if chain > 1 else None
        if cs is not None:
            cs[3] = np.nan            # PIDController._last_time is None
        out = oracle.rollout(prm, g[name + "_state0"], g[name + "_tparams"], g[name + "_K"], max_steps=max_steps,
                             chain=chain, ctrl_state=cs, traj_stride=250)
        gc.assert_rollout_matches(np, out, g, name, tol=2e-11, sum_rtol=1e-13)
        samples = g[name + "_samples"]
        n_s = samples.shape[1]
        got = out["traj"][:n_s, P.TRAJ_STATE:P.TRAJ_STATE + 12, :]
        np.testing.assert_allclose(np.transpose(got, (2, 0, 1)), samples, rtol=0, atol=2e-11, err_msg=name)
        if pid and chain > 1:      # Evaluator semantics: integral and _last_time carry over
            np.testing.assert_allclose(cs[0:3, 0], g[name + "_integral"][:, -1], rtol=0, atol=1e-12)
            assert cs[3, 0] == float(g[name + "_last_time"])
        elif pid:
            cs1 = np.zeros((4, samples.shape[0]))
            oracle.rollout(prm, g[name + "_state0"], g[name + "_tparams"], g[name + "_K"], max_steps=max_steps,
                           ctrl_state=cs1)
            np.testing.assert_allclose(cs1[0:3], g[name + "_integral"], rtol=0, atol=1e-12, err_msg=name)


@pytest.mark.parametrize("name", ["seq_pid", "seq_pidki", "seq_pidff"])
def test_pid_controller_matches_reference(golden, name):
    g = golden["pid_lqr"]
    kw = {"seq_pid": dict(integral_limit=0.0), "seq_pidki": dict(integral_limit=0.05),
          "seq_pidff": dict(integral_limit=2.0, **gc.PID_FF)}[name]
    prm = P.make_params(EnvConfig(), gc.controller_ns(name="pid", **kw))
    obs, K = g[f"{name}_obs"], g[f"{name}_K"]
    state = np.zeros((4, 1))
    state[3] = np.nan
    for k in range(obs.shape[1]):
        act = oracle.compute_action(prm, obs[:, k:k + 1], K, integral=state)
        np.testing.assert_allclose(act[:, 0], g[f"{name}_act"][:, k], rtol=0, atol=1e-13)
        np.testing.assert_allclose(state[0:3, 0], g[f"{name}_integral"][:, k], rtol=0, atol=1e-15)
        assert state[3, 0] == g[f"{name}_last_time"][k]


def test_pid_and_lqr_tuner_scores_match_reference(golden):
    """tuning.py:846-928 for controller_type pid / lqr (tuning_pid_linear.yaml, tuning_lqr_linear.yaml)."""
    from lqr_quadcopter_test_b200.controller import heuristic_lqr_gain
    from lqr_quadcopter_test_b200.metrics import SuccessCriteria

    g = golden["pid_lqr"]
    env = EnvConfig()
    env.target.motion_type = "linear"
    env.simulation.max_episode_time = 4.0
    crit = SuccessCriteria(0.8, 4.0, 0.5)
    for key in ("pid", "pidki", "lqr"):
        cfgs = g[f"tune_{key}_cfg"]
        C_, eps = cfgs.shape[0], 3
        if key == "lqr":
            prm = P.make_params(env, gc.controller_ns(), crit)
            Ks = np.array([heuristic_lqr_gain(c[0:3], c[3:6], 1.0, 1.0) for c in cfgs])
        else:
            prm = P.make_params(env, gc.controller_ns(name="pid", integral_limit=0.0), crit)
            Ks = np.array([np.concatenate([c.reshape(3, 3), np.zeros((1, 3))]) for c in cfgs])  # kp | ki | kd | 0
        s0, tp = oracle.reset_reference(prm, [42, 43, 44])
        out = oracle.rollout(prm, np.tile(s0, (1, C_)), np.tile(tp, (1, C_)), Ks, k_stride=eps, max_steps=400)
        cand = oracle.reduce_scores(out["metrics"], C_, eps)
        np.testing.assert_allclose(cand[0], g[f"tune_{key}_score"], rtol=0, atol=1e-12, err_msg=key)
        np.testing.assert_allclose(cand[1:4].T, g[f"tune_{key}_metrics"], rtol=0, atol=1e-12, err_msg=key)


# ---- §8(f).2: deep tracking policy -------------------------------------------------------------------
def _oracle_mlp(g, name):
    hidden, act, bounds = gc.POLICIES[name]
    W, b = gc.policy_arrays(g, name)
    kw = {} if bounds is None else {"bounds": tuple(bounds[k] for k in ("thrust", "roll_rate", "pitch_rate", "yaw_rate"))}
    return oracle.Mlp(W, b, act, **kw)


@pytest.mark.parametrize("name", list(gc.POLICIES))
def test_policy_network_matches_reference(golden, name):
    """DeepTrackingPolicy.compute_action and the closed loop under it (torch fp32 on the CPU)."""
    g = golden["policy"]
    m = _oracle_mlp(g, name)
    act = oracle.policy_action(m, g[name + "_obs"])
    np.testing.assert_allclose(act, g[name + "_act"], rtol=0, atol=gc.TOL_POLICY_ACTION)
    for motion in ("stationary", "linear"):
        key = f"{name}_{motion}"
        env = EnvConfig()
        env.target.motion_type = motion
        out = oracle.rollout_policy(P.make_params(env), g[key + "_state0"], g[key + "_tparams"], m,
                                    max_steps=int(g[key + "_max_steps"]), traj_stride=250)
        gc.assert_policy_rollout_matches(np, out, g, key)


@pytest.mark.parametrize("sup", ["riccati", "pid"])
def test_supervisor_labels_on_student_states_match_reference(golden, sup):
    """train.py:654-735: the student drives, the supervisor labels (every 10th step recorded)."""
    g = golden["policy"]
    m = _oracle_mlp(g, "relu_64_64")
    env = EnvConfig()
    env.target.motion_type = "linear"
    if sup == "pid":
        prm = P.make_params(env, gc.controller_ns(name="pid", integral_limit=0.4))
        K = np.array([[0.01, 0.01, 4.0], [0.002, 0.002, 0.5], [0.06, 0.06, 2.0], [0.0, 0.0, 0.0]])
    else:
        prm, K = P.make_params(env, gc.controller_ns()), golden["dare"]["lqr_default_K"]
    s0, tp = oracle.reset_reference(prm, [42, 43])
    out = oracle.rollout_policy(prm, s0, tp, m, K=K, max_steps=400, data_stride=10)
    np.testing.assert_allclose(out["features"], g[f"dagger_{sup}_features"], rtol=0, atol=gc.TOL_POLICY_STATE)
    np.testing.assert_allclose(out["actions"], g[f"dagger_{sup}_labels"], rtol=0, atol=gc.TOL_POLICY_STATE)
