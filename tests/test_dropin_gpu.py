"""GPU: interface conformance of the drop-in classes (the behaviours the reference's own tests
assert: tests/test_env_dynamics.py, tests/test_eval.py, tests/test_tuning.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _ctx(gpu_ctx):
    return gpu_ctx


def test_scalar_env_and_controller_loop_reproduces_the_reference(golden):
    """The reference's hot loop, call for call, with the drop-in classes (quadcopter_env.py:152-232)."""
    from lqr_quadcopter_test_b200 import EnvConfig, QuadcopterEnv, RiccatiLQRController

    g = golden["rollouts"]
    cfg = EnvConfig()
    cfg.target.motion_type = "circular"
    env = QuadcopterEnv(config=cfg)
    ctrl = RiccatiLQRController(config={})
    obs = env.reset(seed=42)
    assert set(obs) == {"quadcopter", "target", "time"}
    assert set(obs["quadcopter"]) == {"position", "velocity", "attitude", "angular_velocity"}
    assert set(obs["target"]) == {"position", "velocity", "acceleration"}
    np.testing.assert_allclose(env.get_state_vector(), g["lqr_circular_state0"][:, 0], rtol=0, atol=5e-16)
    done, step = False, 0
    while not done and step < 251:
        if step == 250:
            np.testing.assert_allclose(env.get_state_vector(), g["lqr_circular_samples"][0, 1], rtol=0, atol=1e-10)
        action = ctrl.compute_action(obs)
        assert set(action) == {"thrust", "roll_rate", "pitch_rate", "yaw_rate"}
        assert all(isinstance(v, float) for v in action.values())
        obs, reward, done, info = env.step(action)
        step += 1
    assert set(info) == {"time", "step", "tracking_error", "on_target", "on_target_ratio", "action_violations"}
    assert info["step"] == 251 and abs(env.time - 2.51) < 1e-9 and isinstance(reward, float)
    assert reward == -info["tracking_error"]
    hist = env.get_history()
    assert len(hist) == 25 and hist[0]["step"] == 10 and set(hist[0]) >= {"quadcopter_attitude", "action", "reward"}


def test_env_error_behaviour_and_action_repair():
    from lqr_quadcopter_test_b200 import EnvConfig, QuadcopterEnv

    env = QuadcopterEnv()
    assert not env.is_initialized and env.dt == 0.01
    with pytest.raises(RuntimeError, match="not initialized"):
        env.step(env.hover_action())
    env.reset(seed=1)
    with pytest.raises(ValueError, match=r"shape \(4,\)"):
        env.step(np.zeros(3))
    with pytest.raises(ValueError, match=r"shape \(12,\)"):
        env.set_state_vector(np.zeros(11))
    _, _, _, info = env.step({"thrust": 50.0, "roll_rate": -9.0, "pitch_rate": 0.0, "yaw_rate": float("nan")})
    assert info["action_violations"] == 1
    v = env.get_action_violations()[0]
    assert v["step"] == 0 and any("above max" in m for m in v["violations"]) and any("NaN" in m for m in v["violations"])
    _, _, _, info = env.step(env.hover_action())
    assert info["action_violations"] == 1
    # termination by the time limit, success only reported when done
    cfg = EnvConfig.from_dict({"episode_length": 0.05})
    env = QuadcopterEnv(cfg)
    env.reset(seed=3)
    for k in range(5):
        obs, r, done, info = env.step(env.hover_action())
        assert done == (k == 4)
    assert info["termination_reason"] == "time_limit" and info["success"] is False and info["episode_length"] == info["time"]
    # numerical instability is detected in the step that produces it
    env = QuadcopterEnv()
    env.reset(seed=3)
    bad = env.get_state_vector()
    bad[7] = np.nan
    env.set_state_vector(bad)
    _, _, done, info = env.step(env.hover_action())
    assert done and info["termination_reason"] == "numerical_instability"


def test_reset_seed_semantics():
    from lqr_quadcopter_test_b200 import EnvConfig, QuadcopterEnv

    cfg = EnvConfig()
    cfg.target.motion_type = "linear"
    a, b = QuadcopterEnv(cfg), QuadcopterEnv(cfg)
    oa, ob = a.reset(seed=5), b.reset(seed=5)
    np.testing.assert_array_equal(oa["quadcopter"]["position"], ob["quadcopter"]["position"])
    np.testing.assert_array_equal(oa["target"]["velocity"], ob["target"]["velocity"])
    assert not np.array_equal(a.reset(seed=6)["quadcopter"]["position"], oa["quadcopter"]["position"])
    # seed=None: first reset equals reset(config.seed), the next one continues the streams
    c = QuadcopterEnv(cfg)
    first = c.reset()
    np.testing.assert_array_equal(first["quadcopter"]["position"], QuadcopterEnv(cfg).reset(seed=cfg.seed)["quadcopter"]["position"])
    second = c.reset()
    assert not np.array_equal(second["target"]["velocity"], first["target"]["velocity"])
    assert abs(np.linalg.norm(second["target"]["velocity"]) - 1.0) < 1e-12


def test_controller_interface_and_fallback(golden):
    from lqr_quadcopter_test_b200 import RiccatiLQRController, solve_dare

    g = golden["dare"]
    c = RiccatiLQRController()
    assert c.name == "riccati_lqr" and c.K.shape == (4, 6) and c.P.shape == (6, 6) and c.hover_thrust == 9.81
    assert np.linalg.norm(c.K - g["lqr_default_K"]) / np.linalg.norm(g["lqr_default_K"]) < 1e-9
    assert c.get_gain_matrix() is c.K and c.get_riccati_solution() is c.P and not c.is_lqi_mode()
    assert c.get_integral_gains() is None and c.get_integral_state() is None and not c.is_using_fallback()
    np.testing.assert_allclose(c.P, c.P.T, atol=1e-12)
    assert np.all(np.linalg.eigvalsh(c.P) > 0)
    lqi = RiccatiLQRController({"use_lqi": True, "q_int": [0.01, 0.01, 0.1]})
    assert lqi.K.shape == (4, 9) and lqi.K_pd.shape == (4, 6) and lqi.K_i.shape == (4, 3) and lqi.A_aug.shape == (9, 9)
    assert np.linalg.norm(lqi.K - g["lqi_default_K"]) / np.linalg.norm(g["lqi_default_K"]) < 1e-9
    obs = {"quadcopter": {"position": np.zeros(3), "velocity": np.zeros(3), "attitude": np.zeros(3),
                          "angular_velocity": np.zeros(3)},
           "target": {"position": np.array([1.0, 0.0, 0.0]), "velocity": np.zeros(3)}, "time": 0.0}
    act = lqi.compute_action(obs)
    np.testing.assert_allclose(lqi.get_integral_state(), [0.01, 0.0, 0.0], atol=1e-15)  # integral = dt * err
    assert act["pitch_rate"] > 0 and lqi.get_control_components()["integral_state"][0] == pytest.approx(0.01)
    lqi.reset()
    assert not lqi.get_integral_state().any() and lqi.get_saturation_count() == 0
    with pytest.raises(KeyError, match="target"):
        c.compute_action({"quadcopter": obs["quadcopter"]})
    # sign conventions (controllers/base.py:13-22)
    o2 = dict(obs, target={"position": np.array([0.0, 1.0, 1.0]), "velocity": np.zeros(3)})
    a2 = c.compute_action(o2)
    assert a2["roll_rate"] < 0 and a2["thrust"] > 9.81
    far = dict(obs, target={"position": np.array([0.0, 0.0, 50.0]), "velocity": np.zeros(3)})
    assert c.compute_action(far)["thrust"] == 20.0 and c.get_saturation_count() == 1
    # invalid Q: fallback to the heuristic LQR, or an error when fallback is disabled
    bad = {"Q": np.diag([1.0, 1, 1, 1, 1, -1.0])}
    fb = RiccatiLQRController(bad)
    assert fb.is_using_fallback() and fb.get_gain_matrix() is None and fb.fallback_controller.K.shape == (4, 6)
    assert set(fb.compute_action(obs)) == {"thrust", "roll_rate", "pitch_rate", "yaw_rate"}
    with pytest.raises(ValueError, match="positive semi-definite"):
        RiccatiLQRController({**bad, "fallback_on_failure": False})
    with pytest.raises(ValueError, match="must be square"):
        solve_dare(np.zeros((2, 3)), np.zeros((2, 1)), np.eye(2), np.eye(1))
    with pytest.raises(ValueError, match="positive definite"):
        solve_dare(np.eye(2), np.ones((2, 1)), np.eye(2), np.zeros((1, 1)))
    P_, K = solve_dare(g["gen5x3_A"][0], g["gen5x3_B"][0], g["gen5x3_Q"][0], g["gen5x3_R"][0])
    assert np.linalg.norm(K - g["gen5x3_K"][0]) / np.linalg.norm(g["gen5x3_K"][0]) < 1e-9


def test_evaluator_lqi_carries_controller_state_like_the_reference(golden, tmp_path):
    """eval.py:198-213: evaluate() never resets the controller."""
    from lqr_quadcopter_test_b200 import EnvConfig, Evaluator, load_controller

    g = golden["rollouts"]
    ctrl = load_controller("lqi")
    ev = Evaluator(ctrl, EnvConfig(), output_dir=tmp_path)
    ev.evaluate(num_episodes=3, base_seed=42, verbose=False)
    ref = g["lqi_carry_stationary_metrics"]
    np.testing.assert_array_equal(ev.last_metrics[[0, 2, 3, 12]], ref[[0, 2, 3, 12]])
    np.testing.assert_allclose(ev.last_metrics[4], ref[4], rtol=1e-9)
    np.testing.assert_allclose(ctrl.integral_state, g["lqi_carry_stationary_integral"][:, -1], rtol=0, atol=1e-8)
    # tuner semantics on request: episode 2 then equals the reset-per-episode fixture
    ev2 = Evaluator(load_controller("lqi"), EnvConfig(), output_dir=tmp_path, reset_controller_between_episodes=True)
    ev2.evaluate(num_episodes=3, base_seed=42, verbose=False)
    np.testing.assert_allclose(ev2.last_metrics[4], g["lqi_stationary_metrics"][4][:3], rtol=1e-7)
    data, info = ev.run_episode(seed=42, max_steps=50)
    assert len(data) == 50 and info["step"] == 50 and "termination_reason" not in info


def test_teacher_dataset_shapes_and_content():
    import torch

    from lqr_quadcopter_test_b200 import EnvConfig, RiccatiLQRController, teacher_dataset

    cfg = EnvConfig()
    cfg.target.motion_type = "linear"
    feats, acts, metrics = teacher_dataset(RiccatiLQRController(), cfg, num_episodes=512, base_seed=42, stride=10,
                                           max_steps=1000)
    assert feats.shape == (100 * 512, 18) and acts.shape == (100 * 512, 4)
    assert feats.dtype == torch.float32 and feats.is_cuda and bool(torch.isfinite(feats).all())
    f0 = feats[:512]  # slot 0 = the reset observation: zero velocity / attitude, |offset| <= 0.5
    assert bool((f0[:, 6:12] == 0).all()) and float(f0[:, 0:3].abs().max()) <= 0.5
    assert bool(((acts[:, 0] >= 0) & (acts[:, 0] <= 20)).all()) and bool((acts[:, 1:].abs() <= 3).all())
    assert bool((metrics[0] == 1000).all())


def test_cma_es_strategy_runs_generations_as_batches_and_improves(gpu_ctx, tmp_path):
    """§8(f).3: strategy="cma_es" (tuning.py:1313-1428) without the `cma` package: every generation is
    one batched DARE + rollout evaluation; the search improves on its own first generation."""
    from lqr_quadcopter_test_b200 import ControllerTuner, GainSearchSpace, TuningConfig

    space = GainSearchSpace(q_pos_range=([5e-5, 5e-5, 10.0], [5e-4, 5e-4, 25.0]), q_vel_range=([1e-3, 1e-3, 2.0], [1e-2, 1e-2, 8.0]),
                            r_controls_range=([0.5] * 4, [2.0] * 4))
    tc = TuningConfig(controller_type="riccati_lqr", search_space=space, strategy="cma_es", max_iterations=120,
                      evaluation_episodes=4, evaluation_horizon=600, seed=42, target_motion_type="linear",
                      episode_length=6.0, output_dir=str(tmp_path), cma_popsize=12)
    tuner = ControllerTuner(tc)
    launches0 = gpu_ctx.launches
    res = tuner.tune()
    assert res.iterations_completed == 120 and len(res.all_results) == 120
    # 10 generations; each is seeds + reset + build + DARE (decoupled blocks, validation and dense kernel for what is
    # left) + rollout + reduce, not 12 separate evaluations
    assert gpu_ctx.launches - launches0 <= 10 * 10
    scores = np.array([r["score"] for r in res.all_results])
    assert res.best_score == scores.max() and scores[-24:].mean() > scores[:12].mean()
    lo = np.array([5e-5, 5e-5, 10.0]), np.array([5e-4, 5e-4, 25.0])
    for r in res.all_results:      # evaluated candidates are always inside the box
        assert np.all(np.array(r["config"]["q_pos"]) >= lo[0]) and np.all(np.array(r["config"]["q_pos"]) <= lo[1])
    pid = TuningConfig(controller_type="pid", strategy="cma_es", max_iterations=24, evaluation_episodes=2,
                       evaluation_horizon=300, episode_length=3.0, target_motion_type="linear", output_dir=str(tmp_path / "p"),
                       search_space=GainSearchSpace(kp_pos_range=([0.005, 0.005, 2.0], [0.05, 0.05, 6.0]),
                                                    kd_pos_range=([0.02, 0.02, 1.0], [0.15, 0.15, 3.0])))
    assert ControllerTuner(pid).tune().iterations_completed == 24


@pytest.mark.parametrize("ctype", ["pid", "riccati_lqr"])
def test_feedforward_candidates_are_batched_and_match_single_evaluation(gpu_ctx, tmp_path, ctype):
    """Candidates that differ in their feedforward gains share one launch (qt_rollout ff_gains); the
    scores equal those of evaluating every candidate on its own, and the oracle's."""
    import oracle
    from lqr_quadcopter_test_b200 import ControllerTuner, GainSearchSpace, TuningConfig

    kw = dict(ff_velocity_gain_range=([0.0, 0.0, 0.0], [0.3, 0.3, 0.1]), ff_acceleration_gain_range=([0.0, 0.0, 0.0], [0.2, 0.2, 0.05]))
    if ctype == "pid":
        kw.update(kp_pos_range=([0.005, 0.005, 2.0], [0.05, 0.05, 6.0]), kd_pos_range=([0.02, 0.02, 1.0], [0.15, 0.15, 3.0]))
    else:
        kw.update(q_pos_range=([5e-5, 5e-5, 10.0], [5e-4, 5e-4, 25.0]), r_controls_range=([0.5] * 4, [2.0] * 4))
    tc = TuningConfig(controller_type=ctype, search_space=GainSearchSpace(**kw), strategy="random", max_iterations=6,
                      evaluation_episodes=3, evaluation_horizon=500, seed=11, target_motion_type="circular",
                      episode_length=5.0, output_dir=str(tmp_path))
    tuner = ControllerTuner(tc)
    configs = [tuner._generate_random_config() for _ in range(6)]
    assert all(c["feedforward_enabled"] for c in configs)
    launches0 = gpu_ctx.launches
    scores, metrics = tuner.evaluate_configs(configs)
    batched_launches = gpu_ctx.launches - launches0
    singles = [tuner._evaluate_config(c) for c in configs]
    assert batched_launches < (gpu_ctx.launches - launches0 - batched_launches) / 3
    np.testing.assert_array_equal(scores, [s for s, _ in singles])
    assert metrics == [m for _, m in singles]
    # oracle: same gains, per-candidate feedforward gains
    proto = tuner._create_controller(configs[0])
    env, crit = tuner._env_and_criteria()
    prm = proto.kernel_params(env, crit)
    Ks = np.stack([np.ascontiguousarray(tuner._create_controller(c).effective_gain()) for c in configs])
    ff = np.array([list(c["ff_velocity_gain"]) + list(c["ff_acceleration_gain"]) for c in configs])
    s0, tp = oracle.reset_reference(prm, [11, 12, 13])
    ref = oracle.rollout(prm, np.tile(s0, (1, 6)), np.tile(tp, (1, 6)), Ks, k_stride=3, max_steps=500, ff_gains=ff)
    cand = oracle.reduce_scores(ref["metrics"], 6, 3)
    np.testing.assert_allclose(scores, cand[0], rtol=0, atol=1e-10)


def test_hyperparameter_sweep_ranks_configurations(gpu_ctx, tmp_path):
    """eval.py:516-628: sweep file -> ranked results; a broken entry is recorded and skipped."""
    import json

    from lqr_quadcopter_test_b200 import run_hyperparameter_sweep

    sweep = {"configurations": [
        {"name": "pid", "controller_type": "pid", "num_episodes": 4, "seed": 42,
         "env_config": {"target": {"motion_type": "linear"}}},
        {"name": "riccati", "controller_type": "riccati_lqr", "num_episodes": 4, "seed": 42,
         "env_config": {"target": {"motion_type": "linear"}}, "criteria": {"min_on_target_ratio": 0.5}},
        {"name": "broken", "controller_type": "no_such_controller"},
    ]}
    path = tmp_path / "sweep.json"
    path.write_text(json.dumps(sweep))
    ranked = run_hyperparameter_sweep(path, tmp_path / "out")
    assert [r["name"] for r in ranked] == ["riccati", "pid"]            # better on-target ratio first
    assert ranked[0]["mean_on_target_ratio"] > ranked[1]["mean_on_target_ratio"]
    saved = json.loads((tmp_path / "out" / "sweep_results.json").read_text())
    assert [r["name"] for r in saved] == ["riccati", "pid"]


def test_teacher_dataset_shards_are_slices_of_the_whole(gpu_ctx, monkeypatch):
    """dist.teacher_dataset_sharded (BASELINE config 5 per rank): the shard of rank r is exactly the rows of
    episodes [lo, hi) of the single-process data set; no collective is involved, the torchrun environment
    variables alone select the slice."""
    import torch

    from lqr_quadcopter_test_b200 import EnvConfig, RiccatiLQRController, teacher_dataset
    from lqr_quadcopter_test_b200 import dist as qd

    env = EnvConfig()
    env.target.motion_type = "linear"
    ctrl = RiccatiLQRController(config={})
    E = 1001
    f, a, m = teacher_dataset(ctrl, env, num_episodes=E, base_seed=7, stride=25, max_steps=500, flatten=False)
    for rank in range(3):
        monkeypatch.setenv("RANK", str(rank))
        monkeypatch.setenv("WORLD_SIZE", "3")
        fs, as_, ms, (lo, hi) = qd.teacher_dataset_sharded(ctrl, env, num_episodes=E, base_seed=7, stride=25,
                                                          max_steps=500, flatten=False)
        assert (lo, hi) == qd.shard_range(E, rank, 3)
        assert torch.equal(fs, f[:, lo:hi]) and torch.equal(as_, a[:, lo:hi]) and torch.equal(ms, m[:, lo:hi])
