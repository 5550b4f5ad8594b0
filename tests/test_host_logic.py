"""CPU: host-side logic of the drop-in layer (no kernel launches)."""
import math

import numpy as np
import pytest

from lqr_quadcopter_test_b200 import params as P
from lqr_quadcopter_test_b200.config import EnvConfig
from lqr_quadcopter_test_b200.controller import (build_augmented_lqi_system, build_linearized_system,
                                                 heuristic_lqr_gain, observation_to_vector)
from lqr_quadcopter_test_b200.dist import shard_range
from lqr_quadcopter_test_b200.metrics import (EpisodeMetricsView, SuccessCriteria, episode_metrics_from_vector,
                                              merge_summary_sums, summary_from_sums)
from lqr_quadcopter_test_b200.tuning import CandidateBlock, ControllerTuner, GainSearchSpace, TuningConfig, TuningResult


def test_env_config_round_trip_and_legacy_keys():
    cfg = EnvConfig.from_dict({"dt": 0.02, "episode_length": 12.0, "seed": 7,
                               "target": {"motion_type": "circular", "radius_requirement": 0.3, "center": [1, 2, 3]},
                               "quadcopter": {"mass": 1.5, "unknown_key": 1}})
    assert cfg.simulation.dt == 0.02 and cfg.simulation.max_episode_time == 12.0 and cfg.seed == 7
    assert cfg.success_criteria.target_radius == 0.3 and cfg.target.center == (1, 2, 3)
    assert cfg.quadcopter.mass == 1.5
    again = EnvConfig.from_dict(cfg.to_dict())
    assert again.to_dict() == cfg.to_dict()
    assert EnvConfig().to_dict()["simulation"]["integrator"] == "rk4"


def test_make_params_flattens_configuration():
    cfg = EnvConfig()
    cfg.target.motion_type = "Sinusoidal"
    cfg.simulation.integrator = "euler"
    p = P.make_params(cfg, None, SuccessCriteria(0.7, 20.0, 0.4))
    assert p.motion_type == 3 and p.integrator == 1
    assert (p.crit_min_on_target_ratio, p.crit_min_episode_duration, p.crit_target_radius) == (0.7, 20.0, 0.4)
    cfg.target.motion_type = "zigzag"
    with pytest.raises(ValueError, match="Invalid motion type"):
        P.make_params(cfg)


def test_design_matrices(golden):
    g = golden["dare"]
    A, B = build_linearized_system(0.01)
    np.testing.assert_array_equal(A, g["lqr_default_A"])
    np.testing.assert_array_equal(B, g["lqr_default_B"])
    A9, B9 = build_augmented_lqi_system(A, B, 0.01)
    np.testing.assert_array_equal(A9, g["lqi_default_A"])
    np.testing.assert_array_equal(B9, g["lqi_default_B"])
    K = heuristic_lqr_gain([1e-4, 1e-4, 16.0], [3.6e-3, 3.6e-3, 4.0], 1.0, 1.0)
    assert K.shape == (4, 6) and K[0, 2] == 4.0 and K[1, 1] == -0.01 and K[2, 0] == 0.01 and not K[3].any()


def test_observation_vector_layout():
    obs = {"quadcopter": {"position": [1, 2, 3], "velocity": [4, 5, 6], "attitude": [7, 8, 9],
                          "angular_velocity": [10, 11, 12]},
           "target": {"position": [13, 14, 15], "velocity": [16, 17, 18]}, "time": 0.5}
    v = observation_to_vector(obs)
    assert v.shape == (22,) and list(v[:18]) == list(range(1, 19)) and not v[18:21].any() and v[21] == 0.5


def test_search_space_validation_messages():
    GainSearchSpace(q_pos_range=([0, 0, 0], [1, 1, 1])).validate()
    with pytest.raises(ValueError, match="inverted range"):
        GainSearchSpace(q_pos_range=([2, 0, 0], [1, 1, 1])).validate()
    with pytest.raises(ValueError, match="negative minimum"):
        GainSearchSpace(q_vel_range=([-1, 0, 0], [1, 1, 1])).validate()
    with pytest.raises(ValueError, match="exactly 4 values"):
        GainSearchSpace(r_controls_range=([1, 1, 1], [2, 2, 2])).validate()
    with pytest.raises(ValueError, match="exactly 3 values"):
        GainSearchSpace(q_int_range=([0, 0], [1, 1])).validate()
    sp = GainSearchSpace(q_pos_range=([0] * 3, [1] * 3), r_controls_range=([1] * 4, [2] * 4))
    assert sp.get_active_parameters() == ["q_pos", "r_controls"]
    assert GainSearchSpace.from_dict(sp.to_dict()) == sp


def test_tuning_config_validation():
    with pytest.raises(ValueError, match="Invalid controller_type"):
        TuningConfig(controller_type="mpc")
    with pytest.raises(ValueError, match="Invalid strategy"):
        TuningConfig(strategy="bayes")
    with pytest.raises(ValueError, match="max_iterations"):
        TuningConfig(max_iterations=0)
    tc = TuningConfig(controller_type="riccati_lqr", max_iterations=3)
    assert TuningConfig.from_dict(tc.to_dict()).to_dict() == tc.to_dict()


def _riccati_space(lqi=False):
    kw = dict(q_pos_range=([5e-5, 5e-5, 10.0], [5e-4, 5e-4, 25.0]), q_vel_range=([1e-3, 1e-3, 2.0], [1e-2, 1e-2, 8.0]),
              r_controls_range=([0.5] * 4, [2.0] * 4))
    if lqi:
        kw["q_int_range"] = ([1e-6, 1e-6, 1e-4], [1e-5, 1e-5, 1e-2])
    return GainSearchSpace(**kw)


@pytest.mark.parametrize("lqi", [False, True])
def test_random_candidates_consume_the_rng_like_the_reference(golden, lqi, tmp_path):
    g = golden["tuning"]
    key = "riccati_lqi_stationary" if lqi else "riccati_stationary"
    tuner = ControllerTuner(TuningConfig(controller_type="riccati_lqr", search_space=_riccati_space(lqi), seed=42,
                                         output_dir=str(tmp_path)))
    cfgs = [tuner._generate_random_config() for _ in range(5)]
    q = np.array([c["q_pos"] + c["q_vel"] + (c["q_int"] if lqi else []) for c in cfgs])
    r = np.array([c["r_controls"] for c in cfgs])
    np.testing.assert_array_equal(q, g[key + "_q"])
    np.testing.assert_array_equal(r, g[key + "_r"])
    assert all(c.get("use_lqi", False) == lqi for c in cfgs)


def _full_space():
    return GainSearchSpace(kp_pos_range=([0.001, 0.001, 1.0], [0.1, 0.1, 8.0]), kd_pos_range=([0.01] * 3, [0.2, 0.2, 4.0]),
                           ff_velocity_gain_range=([0.0] * 3, [0.3] * 3), q_pos_range=([5e-5, 5e-5, 10.0], [5e-4, 5e-4, 25.0]),
                           r_thrust_range=(0.5, 2.0), r_rate_range=(0.1, 1.0), r_controls_range=([0.5] * 4, [2.0] * 4),
                           q_int_range=([1e-6, 1e-6, 1e-4], [1e-5, 1e-5, 1e-2]))


@pytest.mark.parametrize("space", ["riccati", "riccati_lqi", "everything", "use_lqi_flag"])
def test_vectorised_sampling_is_the_scalar_draw_sequence(golden, space, tmp_path):
    """sample_candidates (one rng.uniform block) against _generate_random_config called in a loop (the reference's
    scalar draws, tuning.py:683-735; itself pinned to the reference's fixture above): identical dictionaries,
    identical generator state afterwards, and a rank's shard [lo, hi) equals the same rows of the full draw."""
    sp = {"riccati": _riccati_space(False), "riccati_lqi": _riccati_space(True), "everything": _full_space(),
          "use_lqi_flag": GainSearchSpace(q_pos_range=([5e-5, 5e-5, 10.0], [5e-4, 5e-4, 25.0]), use_lqi=True)}[space]
    mk = lambda: ControllerTuner(TuningConfig(controller_type="riccati_lqr", search_space=sp, seed=42, output_dir=str(tmp_path)))
    C = 257
    a, b = mk(), mk()
    scalar = [a._generate_random_config() for _ in range(C)]
    block = b.sample_candidates(C)
    assert block.configs() == scalar
    assert a.rng.random() == b.rng.random()                     # both generators ended in the same state
    if space == "riccati":
        g = golden["tuning"]
        np.testing.assert_array_equal(np.concatenate([block.column("q_pos"), block.column("q_vel")], axis=1)[:5], g["riccati_stationary_q"])
        np.testing.assert_array_equal(block.column("r_controls")[:5], g["riccati_stationary_r"])
    for lo, hi in ((0, 100), (100, 101), (101, 257), (257, 257)):
        c = mk()
        part = c.sample_candidates(C, lo, hi)
        np.testing.assert_array_equal(part.values, block.values[lo:hi])
        assert c.rng.random() == mk_state_after(sp, tmp_path, C)
    # dictionaries -> block -> dictionaries is the identity
    again = CandidateBlock.from_configs(scalar)
    assert again is not None and again.configs() == scalar and again.use_lqi == block.use_lqi
    assert CandidateBlock.from_configs(scalar[:3] + [{"kd_pos": [0.1, 0.1, 1.0], "mass": 2.0}]) is None   # other keys: not one block


def mk_state_after(sp, tmp_path, C):
    t = ControllerTuner(TuningConfig(controller_type="riccati_lqr", search_space=sp, seed=42, output_dir=str(tmp_path)))
    t.sample_candidates(C)
    return t.rng.random()


def test_argmax_first_follows_the_strict_scan():
    """tuning.py:1084: `if score > best_score` while scanning — first of equal scores, NaN never wins."""
    from lqr_quadcopter_test_b200.dist import argmax_first

    def scan(scores):
        best, idx = float("-inf"), -1
        for i, s in enumerate(scores):
            if s > best:
                best, idx = s, i
        return idx

    for scores in ([0.1, 0.5, 0.5, 0.2], [float("nan"), 0.3, 0.7], [0.3, float("nan"), 0.3], [float("nan")] * 3,
                   [float("-inf"), float("-inf")], [1.0]):
        got = argmax_first(np.array(scores))
        want = scan(scores)
        assert got == want or (want == -1 and (got == -1 or scores[got] == float("-inf"))), (scores, got, want)


def test_evaluation_shard_plan():
    from lqr_quadcopter_test_b200.dist import evaluation_shard_plan

    assert [evaluation_shard_plan(10, r, 4, False) for r in range(4)] == [(0, 3, 1), (3, 6, 1), (6, 8, 1), (8, 10, 1)]
    # a controller that carries state across episodes (LQI under Evaluator semantics, eval.py:198-213): one chain on rank 0
    assert [evaluation_shard_plan(10, r, 4, True) for r in range(4)] == [(0, 10, 10), (0, 0, 1), (0, 0, 1), (0, 0, 1)]
    assert evaluation_shard_plan(1, 0, 2, True) == (0, 1, 2)


def test_grid_candidates_match_the_reference(golden, tmp_path):
    g = golden["tuning"]
    tc = TuningConfig(controller_type="riccati_lqr", strategy="grid", grid_points_per_dim=2, max_iterations=10,
                      output_dir=str(tmp_path),
                      search_space=GainSearchSpace(q_pos_range=([1e-4, 1e-4, 10.0], [1e-4, 2e-4, 20.0]),
                                                   r_controls_range=([1.0] * 4, [1.0, 1.0, 1.0, 2.0])))
    grid = ControllerTuner(tc)._generate_grid_configs()
    np.testing.assert_array_equal(np.array([c["q_pos"] for c in grid]), g["grid_q_pos"])
    np.testing.assert_array_equal(np.array([c["r_controls"] for c in grid]), g["grid_r_controls"])


def test_tuner_rejects_what_is_out_of_scope(tmp_path):
    ControllerTuner(TuningConfig(controller_type="pid", output_dir=str(tmp_path)))   # §8(f).1: in scope now
    ControllerTuner(TuningConfig(controller_type="riccati_lqr", strategy="cma_es", output_dir=str(tmp_path)))   # §8(f).3
    with pytest.raises(ValueError, match="traversal"):
        ControllerTuner(TuningConfig(controller_type="riccati_lqr", output_dir="../evil"))


def test_tuning_result_round_trip(tmp_path):
    r = TuningResult({"a": 1}, 0.5, {"m": 2}, [{"config": {}, "score": 0.5, "metrics": {}}], 1, False, "t", {})
    path = r.save(tmp_path / "res.json")
    assert TuningResult.load(path).to_dict() == r.to_dict()
    with pytest.raises(FileNotFoundError):
        TuningResult.load(tmp_path / "missing.json")


def test_episode_metrics_from_vector_and_view(golden):
    g = golden["rollouts"]
    M = g["lqr_linear_metrics"]
    extra = g["lqr_linear_extra"]
    view = EpisodeMetricsView(M)
    assert len(view) == M.shape[1]
    for i, m in enumerate(view):
        assert m.on_target_ratio == extra[0, i]
        assert abs(m.mean_tracking_error - extra[1, i]) < 1e-15
        assert abs(m.rms_tracking_error - extra[2, i]) < 1e-15
        assert abs(m.mean_control_effort - extra[3, i]) < 1e-13
        assert m.termination_reason == "time_limit" and m.episode_duration >= 30.0
    empty = episode_metrics_from_vector(np.zeros(16))
    assert empty.termination_reason == "no_data" and not empty.success


def _sums_of(ratios, errs, effort, succ, offset=0):
    n = len(ratios)
    s = np.zeros(16)
    s[0], s[1], s[2], s[3], s[4] = n, succ.sum(), ratios.sum(), errs.sum(), effort.sum()
    s[5], s[6] = ((ratios - ratios.mean()) ** 2).sum(), ((errs - errs.mean()) ** 2).sum()
    s[7], s[8], s[9], s[10] = ratios.max(), np.argmax(ratios), ratios.min(), np.argmin(ratios)
    return s


def test_merge_of_shard_sums_equals_numpy_on_the_whole():
    rng = np.random.default_rng(3)
    n = 1000
    ratios = np.round(rng.uniform(0.5, 1.0, n), 2)  # ties on purpose
    errs, effort, succ = rng.uniform(0, 1, n), rng.uniform(9, 11, n), (ratios >= 0.8).astype(float)
    cuts = [0, 137, 600, 601, 1000]
    parts = [_sums_of(ratios[a:b], errs[a:b], effort[a:b], succ[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    merged = merge_summary_sums(parts, cuts[:-1])
    s = summary_from_sums(merged, SuccessCriteria(), [])
    assert s.total_episodes == n and s.successful_episodes == int(succ.sum())
    assert math.isclose(s.mean_on_target_ratio, ratios.mean(), rel_tol=1e-13)
    assert math.isclose(s.std_on_target_ratio, ratios.std(), rel_tol=1e-11)
    assert math.isclose(s.std_tracking_error, errs.std(), rel_tol=1e-11)
    assert s.best_episode_idx == int(np.argmax(ratios)) and s.worst_episode_idx == int(np.argmin(ratios))
    assert s.meets_criteria == (ratios.mean() >= 0.8)


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 64, 65536, 1_000_003):
        for size in (1, 2, 3, 8):
            spans = [shard_range(total, r, size) for r in range(size)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            widths = [hi - lo for lo, hi in spans]
            assert max(widths) - min(widths) <= 1


def test_controller_configuration_errors_come_before_any_gpu_work():
    from lqr_quadcopter_test_b200 import RiccatiLQRController

    with pytest.raises(ValueError, match="q_int must have exactly 3 elements"):
        RiccatiLQRController(config={"use_lqi": True, "q_int": [0.1, 0.2]})
    with pytest.raises(ValueError, match="non-negative"):
        RiccatiLQRController(config={"use_lqi": True, "q_int": [0.1, -0.2, 0.1]})
    with pytest.raises(ValueError, match="q_pos must have 3 elements"):
        RiccatiLQRController(config={"q_pos": [1, 2]})
    with pytest.raises(ValueError, match=r"R matrix must have shape \(4, 4\)"):
        RiccatiLQRController(config={"R": np.eye(3)})
    with pytest.raises(ValueError, match=r"Q matrix must have shape \(6, 6\)"):
        RiccatiLQRController(config={"Q": np.eye(5)})


def test_training_config_mirrors_the_reference(tmp_path):
    """train.py:60-273: field names, defaults, validation messages, dict / file round trip."""
    import json

    from lqr_quadcopter_test_b200 import TrainingConfig

    c = TrainingConfig()
    assert (c.controller, c.epochs, c.episodes_per_epoch, c.batch_size, c.training_mode) == ("deep", 100, 10, 32, "tracking")
    assert c.hidden_sizes == [64, 64] and c.supervisor_controller == "pid" and c.target_motion_type == "circular"
    with pytest.raises(ValueError, match="Invalid training_mode"):
        TrainingConfig(training_mode="dagger")
    with pytest.raises(ValueError, match="Invalid supervisor_controller"):
        TrainingConfig(supervisor_controller="mpc")
    with pytest.raises(ValueError, match="Invalid controller type"):
        TrainingConfig(controller="mpc")
    d = {"training_mode": "imitation", "supervisor_controller": "riccati_lqr", "epochs": 3, "pid": {"kp_pos": [1, 1, 1]},
         "not_a_field": 7}
    path = tmp_path / "cfg.json"
    path.write_text(json.dumps(d))
    c2 = TrainingConfig.from_file(path)
    assert c2.epochs == 3 and c2.full_config["pid"] == {"kp_pos": [1, 1, 1]} and "not_a_field" not in c2.to_dict()
    assert TrainingConfig.from_dict(c2.to_dict()).to_dict() == c2.to_dict()


def test_cma_es_specification_and_convergence(tmp_path):
    """tuning.py:1156-1262: parameter spec, bounds, x0 (geometric mean for Q/R weights) as in the reference;
    the ask/tell loop (cmaes.py) converges on a bounded quadratic, including an optimum on the boundary."""
    from lqr_quadcopter_test_b200.cmaes import CMAEvolutionStrategy

    sp = GainSearchSpace(q_pos_range=([1e-4, 1e-4, 10.0], [4e-4, 9e-4, 20.0]), r_thrust_range=(0.5, 2.0),
                         kp_pos_range=([0.0, 0.0, 2.0], [0.1, 0.1, 6.0]))
    t = ControllerTuner(TuningConfig(controller_type="riccati_lqr", strategy="cma_es", search_space=sp, output_dir=str(tmp_path)))
    t._cma_param_spec = t._build_cma_param_spec()
    assert [n for n, _, _ in t._cma_param_spec] == ["kp_pos", "q_pos", "r_thrust"]
    lo, hi = t._get_cma_bounds()
    assert lo == [0.0, 0.0, 2.0, 1e-4, 1e-4, 10.0, 0.5] and hi == [0.1, 0.1, 6.0, 4e-4, 9e-4, 20.0, 2.0]
    np.testing.assert_allclose(t._get_cma_x0(), [0.05, 0.05, 4.0, 2e-4, 3e-4, np.sqrt(200.0), 1.25])
    cfg = t._vector_to_config(list(range(7)))
    assert cfg == {"kp_pos": [0.0, 1.0, 2.0], "q_pos": [3.0, 4.0, 5.0], "r_thrust": 6.0}
    lo, hi = np.array(lo), np.array(hi)
    opt = lo + (hi - lo) * np.array([0.3, 0.7, 0.5, 0.2, 0.9, 0.0, 1.0])
    f = lambda x: float(np.sum(((x - opt) / (hi - lo)) ** 2))
    es = CMAEvolutionStrategy(t._get_cma_x0(), 0.3 * float(np.exp(np.mean(np.log(hi - lo)))), lo, hi, seed=42, maxfevals=3000)
    assert es.lam == 4 + int(3 * np.log(7)) and es.mu == es.lam // 2
    while not es.stop():
        X = es.ask()
        assert len(X) == es.lam
        es.tell(X, [es.penalised(x, f(es.repair(x))) for x in X])
    assert es.best_f < 1e-9 and np.all(es.best_x >= lo) and np.all(es.best_x <= hi)


def test_cma_es_checkpoint_round_trip(tmp_path):
    """tuning.py:1357-1375 / 1430-1442: the optimiser object is pickled to <output_dir>/cma_checkpoint.pkl and a
    restored copy continues with the same candidates (state and generator included)."""
    import pickle

    from lqr_quadcopter_test_b200.cmaes import CMAEvolutionStrategy

    lo, hi = np.zeros(3), np.array([1.0, 2.0, 4.0])
    es = CMAEvolutionStrategy([0.5, 1.0, 2.0], 0.3, lo, hi, seed=7, maxfevals=200)
    f = lambda x: float(np.sum((x - 0.25 * hi) ** 2))
    for _ in range(3):
        X = es.ask()
        es.tell(X, [es.penalised(x, f(es.repair(x))) for x in X])
    t = ControllerTuner(TuningConfig(controller_type="pid", strategy="cma_es", output_dir=str(tmp_path / "out")))
    t._save_cma_checkpoint()                                   # nothing to save yet: no file, no error
    assert not (tmp_path / "out" / "cma_checkpoint.pkl").exists()
    t.cma_es, t._cma_checkpoint_path = es, tmp_path / "out" / "cma_checkpoint.pkl"
    t._save_cma_checkpoint()
    with open(tmp_path / "out" / "cma_checkpoint.pkl", "rb") as fh:
        back = pickle.load(fh)
    assert back.generation == es.generation == 3 and back.counteval == es.counteval
    np.testing.assert_array_equal(np.array(back.ask()), np.array(es.ask()))


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs next to ours): one JSON line on stdout with the
    contract's keys; under torchrun only rank 0 prints."""
    import json
    import os
    import pathlib
    import subprocess
    import sys

    ROOT = pathlib.Path(__file__).resolve().parent.parent
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["value"] > 1e6
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and "workload" in d["config"]
    env["RANK"] = "1"
    r1 = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, env=env, timeout=300)
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_coordinate_frame_helpers():
    """utils/coordinate_frame.py:44-405: ENU constants and the assertions built on them."""
    from lqr_quadcopter_test_b200 import coordinate_frame as cf

    assert (cf.AXIS_X, cf.AXIS_Y, cf.AXIS_Z) == (0, 1, 2) and cf.get_current_frame() is cf.ENU_FRAME
    assert str(cf.ENU_FRAME) == "ENU: X=East, Y=North, Z=Up" and cf.NED_FRAME.gravity_sign == "+"
    cf.assert_gravity_direction_enu([0.0, 0.0, -9.81])
    cf.assert_thrust_direction_enu([0.0, 0.0, 0.0])
    with pytest.raises(cf.ENUFrameError, match="should point in -Z"):
        cf.assert_gravity_direction_enu([0.0, 0.0, 9.81])
    with pytest.raises(cf.ENUFrameError, match="zero magnitude"):
        cf.assert_gravity_direction_enu(np.zeros(3))
    with pytest.raises(cf.ENUFrameError, match=r"should point in \+Z"):
        cf.assert_thrust_direction_enu([0.0, 0.0, -1.0])
    cf.assert_control_signs_enu([1.0, -1.0, 0.0], pitch_rate=0.2, roll_rate=0.3)          # +x -> +pitch, -y -> +roll
    cf.assert_control_signs_enu([0.3, 0.3, 0.0], pitch_rate=-1.0, roll_rate=1.0)          # below the 0.5 m threshold
    with pytest.raises(cf.ENUFrameError, match=r"\+Y error"):
        cf.assert_control_signs_enu([0.0, 1.0, 0.0], 0.0, 0.5)
    with pytest.raises(cf.ENUFrameError, match="-X error"):
        cf.assert_control_signs_enu([-0.3, 0.0, 0.0], 0.0, 0.0, error_threshold=0.2)
    with pytest.raises(cf.ENUFrameError, match="Z-axis.*below minimum"):
        cf.assert_z_up([0.0, 0.0, -150.0])
    with pytest.raises(cf.ENUFrameError, match="Target Z position.*very negative"):
        cf.validate_observation_frame({"quadcopter": {"position": [0, 0, 1.0]}, "target": {"position": [0, 0, -50.0]}})
    assert cf.validate_observation_frame({"quadcopter": {"position": [0, 0, 1.0]}, "target": {"position": [1, 1, 2.0]}})
    assert cf.compute_expected_hover_thrust(2.0) == 19.62
    np.testing.assert_array_equal(cf.compute_position_error_enu([0, 0, 1.0], [2, 3, 4.0]), [2.0, 3.0, 3.0])


def test_training_losses_on_cpu_tensors():
    """utils/losses.py:25-406: the three error kinds, matrix weights, imitation target, reward shaping, the combined
    total and the per-epoch logger — against closed forms."""
    import torch

    from lqr_quadcopter_test_b200.losses import (CombinedLoss, LossLogger, RewardShapingLoss, TrackingLoss,
                                                 create_loss_from_config)

    pe = torch.tensor([[1.0, 2.0, 2.0], [0.0, 0.0, 0.0]])
    ve, act = torch.zeros(2, 3), torch.tensor([[1.0, 0.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0]])
    out = TrackingLoss(position_weight=np.diag([2.0, 1.0, 1.0]), velocity_weight=0.1, control_weight=0.5)(pe, ve, act)
    assert float(out["position"]) == pytest.approx((2 * 1 + 4 + 4) / 2) and float(out["control"]) == pytest.approx(0.5)
    assert float(out["total"]) == pytest.approx(float(out["position"] + out["velocity"] + out["control"]))
    assert float(TrackingLoss(error_type="l1")(pe, ve, act)["position"]) == pytest.approx(5.0 / 2)
    hub = TrackingLoss(error_type="huber", huber_delta=1.0)(pe, ve, act)["position"]        # |e| = 3 and 0
    assert float(hub) == pytest.approx((1.0 * (3.0 - 0.5) + 0.0) / 2)
    imit = TrackingLoss(control_weight=1.0)(pe, ve, act, action_target=act)
    assert float(imit["control"]) == 0.0
    with pytest.raises(ValueError, match="Unknown error type"):
        TrackingLoss(error_type="l3")(pe, ve, act)
    r = RewardShapingLoss(target_radius=0.5, on_target_bonus=2.0, distance_penalty=1.0)(torch.tensor([0.25, 1.0]))
    assert r["on_target"].tolist() == [True, False] and float(r["reward"]) == pytest.approx((2.0 - 0.25 - 1.0) / 2)
    assert float(RewardShapingLoss(smoothing="exp")(torch.tensor(0.0))["penalty"]) == 0.0
    c = create_loss_from_config({"position_weight": 2.0, "error_type": "huber", "target_radius": 0.3, "reward_weight": 0.5})
    assert isinstance(c, CombinedLoss) and c.tracking_loss.error_type == "huber" and c.reward_loss.target_radius == 0.3
    full = c(pe, ve, act, tracking_error=pe.norm(dim=-1))
    assert float(full["total"]) == pytest.approx(float(full["tracking_total"]) - 0.5 * float(full["reward"]))
    assert float(full["on_target_ratio"]) == 0.5
    pe.requires_grad_(True)
    CombinedLoss()(pe, ve, act, tracking_error=pe.norm(dim=-1))["total"].backward()
    assert pe.grad is not None and torch.isfinite(pe.grad).all()
    log = LossLogger()
    for v in (1.0, torch.tensor(3.0)):
        log.log({"total": v})
    assert log.end_epoch() == {"total": 2.0} and log.get_history() == [{"total": 2.0}] and log.get_best_epoch() == (0, 2.0)
    log.reset()
    assert log.get_history() == [] and log.get_best_epoch() == (-1, float("inf"))
