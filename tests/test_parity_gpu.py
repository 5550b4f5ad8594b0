"""GPU: CUDA path vs the CPU oracle on seeded inputs at sizes the oracle finishes in seconds, plus
size-independent properties at BASELINE.json's full sizes."""
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


@pytest.fixture(scope="module")
def orc():
    import oracle

    return oracle


def _gains(golden, lqi):
    return golden["dare"]["lqi_default_K" if lqi else "lqr_default_K"]


def _compare_rollouts(got, ref, tol, label):
    """Counters must agree except where a sample sits on the on-target threshold to within the
    trajectory tolerance; returns the number of such borderline episodes (reported, not hidden)."""
    M, G = got["metrics"], ref["metrics"]
    hard = [P.M_STEPS, P.M_TERMINATION, P.M_VIOLATIONS]
    np.testing.assert_array_equal(M[hard], G[hard], err_msg=label)
    soft = [P.M_ENV_ON_TARGET, P.M_MET_ON_TARGET]
    diff = np.abs(M[soft] - G[soft]).max(axis=0)
    borderline = int((diff > 0).sum())
    assert diff.max() <= 2, f"{label}: on-target counts differ by more than borderline flips"
    assert borderline <= max(2, M.shape[1] // 200), f"{label}: {borderline} borderline episodes"
    same = diff == 0
    for row in (P.M_ENV_SUCCESS, P.M_MET_SUCCESS, P.M_OVERSHOOT_COUNT, P.M_SATURATIONS):
        np.testing.assert_array_equal(M[row][same], G[row][same], err_msg=f"{label} row {row}")
    np.testing.assert_allclose(M[gc.SUM_ROWS], G[gc.SUM_ROWS], rtol=max(1e-11, tol), atol=tol * 3000, err_msg=label)
    np.testing.assert_allclose(M[gc.VALUE_ROWS][:, same], G[gc.VALUE_ROWS][:, same], rtol=0, atol=tol, err_msg=label)
    np.testing.assert_allclose(got["final_state"], ref["final_state"], rtol=0, atol=tol, err_msg=label)
    return borderline


@pytest.mark.parametrize("motion", MOTIONS)
@pytest.mark.parametrize("mode", ["lqr", "lqi"])
def test_reset_and_rollout_match_oracle(golden, gb, orc, motion, mode):
    E = 1536
    lqi = mode == "lqi"
    env = EnvConfig()
    env.target.motion_type = motion
    prm = P.make_params(env, gc.controller_ns(use_lqi=lqi))
    seeds = np.arange(100_000, 100_000 + E)
    s0_ref, tp_ref = orc.reset_reference(prm, seeds)
    s0, tp = gb.reset(prm, seeds)
    np.testing.assert_array_equal(tp, tp_ref)
    np.testing.assert_allclose(s0, s0_ref, rtol=0, atol=5e-16)
    K = _gains(golden, lqi)
    tol = gc.TOL_LQI if lqi else (1e-9 if motion == "figure8" else gc.TOL_LQR)
    ref = orc.rollout(prm, s0_ref, tp_ref, K, max_steps=3002, n_threads=8)
    got = gb.rollout(prm, s0_ref, tp_ref, K, max_steps=3002)
    n_border = _compare_rollouts(got, ref, tol, f"{mode}/{motion}")
    print(f"[parity] {mode}/{motion}: E={E}, borderline on-target episodes = {n_border}, "
          f"max |final state diff| = {np.abs(got['final_state'] - ref['final_state']).max():.2e}")


def test_per_group_gains_and_horizon_cap_match_oracle(golden, gb, orc):
    """Autotune layout: 64 episodes share each candidate's K (k_stride = 64), capped horizon."""
    g = golden["dare"]
    C_, eps = 48, 64
    env = EnvConfig()
    prm = P.make_params(env, gc.controller_ns())
    s0e, tpe = orc.reset_reference(prm, np.arange(42, 42 + eps))
    s0, tp = np.tile(s0e, (1, C_)), np.tile(tpe, (1, C_))
    ref = orc.rollout(prm, s0, tp, g["sweep6_K"], k_stride=eps, max_steps=700, n_threads=8)
    got = gb.rollout(prm, s0, tp, g["sweep6_K"], k_stride=eps, max_steps=700)
    _compare_rollouts(got, ref, gc.TOL_LQR, "autotune layout")
    np.testing.assert_array_equal(gb.reduce_scores(ref["metrics"], C_, eps), orc.reduce_scores(ref["metrics"], C_, eps))
    # odd group size: lanes of one warp read different gains
    ref = orc.rollout(prm, s0[:, :45], tp[:, :45], g["sweep6_K"][:9], k_stride=5, max_steps=300)
    got = gb.rollout(prm, s0[:, :45], tp[:, :45], g["sweep6_K"][:9], k_stride=5, max_steps=300)
    _compare_rollouts(got, ref, gc.TOL_LQR, "k_stride 5")


def test_logging_variants_match_oracle(golden, gb, orc):
    """Trajectory log and teacher data set (A12): same rows as the oracle, fp32 rows bit-equal to
    the cast of the fp64 values wherever those agree."""
    env = EnvConfig()
    env.target.motion_type = "circular"
    prm = P.make_params(env, gc.controller_ns())
    s0, tp = orc.reset_reference(prm, np.arange(7, 7 + 96))
    K = _gains(golden, False)
    ref = orc.rollout(prm, s0, tp, K, max_steps=600, traj_stride=7, data_stride=10)
    got = gb.rollout(prm, s0, tp, K, max_steps=600, traj_stride=7, data_stride=10)
    np.testing.assert_allclose(got["traj"], ref["traj"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(got["features"], ref["features"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(got["actions"], ref["actions"], rtol=0, atol=1e-6)
    assert (got["features"] != ref["features"]).mean() < 0.01  # fp32 rounding of values equal to 1e-13
    # stride-10 data rows are the stride-1 trajectory rows cast to fp32 (layout check, exact)
    full = gb.rollout(prm, s0, tp, K, max_steps=600, traj_stride=1, data_stride=10)
    tr = full["traj"][::10]
    st, tpos, tvel = tr[:, P.TRAJ_STATE:P.TRAJ_STATE + 12], tr[:, P.TRAJ_TARGET_P:P.TRAJ_TARGET_P + 3], tr[:, P.TRAJ_TARGET_V:P.TRAJ_TARGET_V + 3]
    feat = np.concatenate([tpos - st[:, 0:3], tvel - st[:, 3:6], st[:, 6:9], st[:, 9:12], tpos, tvel], axis=1)
    np.testing.assert_array_equal(full["features"], np.transpose(feat, (0, 2, 1)).astype(np.float32))
    np.testing.assert_array_equal(full["actions"], np.transpose(tr[:, P.TRAJ_ACTION:P.TRAJ_ACTION + 4], (0, 2, 1)).astype(np.float32))


def test_chains_and_controller_state_match_oracle(golden, gb, orc):
    """Evaluator semantics (one serial chain, state carried in and out) vs reset per episode."""
    env = EnvConfig()
    prm = P.make_params(env, gc.controller_ns(use_lqi=True, integral_limit=0.5))
    s0, tp = orc.reset_reference(prm, np.arange(500, 512))
    K = golden["rollouts"]["lqi_small_linear_K"]
    for chain, n_chains in ((1, 12), (4, 3), (5, 3), (12, 1), (64, 1)):
        cs_ref = np.zeros((4, n_chains))
        cs_ref[0:3] = 0.01 * np.arange(3 * n_chains).reshape(3, n_chains)
        cs_got = cs_ref.copy()
        ref = orc.rollout(prm, s0, tp, K, max_steps=400, chain=chain, ctrl_state=cs_ref)
        got = gb.rollout(prm, s0, tp, K, max_steps=400, chain=chain, ctrl_state=cs_got)
        _compare_rollouts(got, ref, 1e-10, f"chain {chain}")
        np.testing.assert_allclose(cs_got, cs_ref, rtol=0, atol=1e-12)


def test_stepwise_env_matches_fused_rollout(golden, gb, orc):
    """BatchedQuadcopterEnv.step (controller in the loop) walks the same trajectory as qt_rollout."""
    import torch

    from lqr_quadcopter_test_b200 import BatchedQuadcopterEnv, RiccatiLQRController

    for cfg, motion in (({}, "linear"), ({"use_lqi": True, "q_int": [1e-6, 1e-6, 1e-3]}, "sinusoidal")):
        env_cfg = EnvConfig()
        env_cfg.target.motion_type = motion
        ctrl = RiccatiLQRController(config=cfg)
        E, T = 256, 60
        env = BatchedQuadcopterEnv(env_cfg, num_envs=E)
        obs0 = env.reset(1000).clone()
        integral = torch.zeros((3, E), dtype=torch.float64, device=env.device) if ctrl.use_lqi else None
        sat = torch.zeros(E, dtype=torch.int64, device=env.device)
        for _ in range(T):
            obs, reward, done, info = env.step(controller=ctrl, integral=integral, saturation=sat)
        prm = ctrl.kernel_params(env_cfg)
        s0, tp = orc.reset_reference(prm, np.arange(1000, 1000 + E))
        np.testing.assert_allclose(obs0.cpu().numpy()[:12], s0, rtol=0, atol=5e-16)
        fused = gb.rollout(prm, s0, tp, ctrl.effective_gain(), max_steps=T)
        np.testing.assert_allclose(env.state.cpu().numpy(), fused["final_state"], rtol=0, atol=1e-12)
        np.testing.assert_array_equal(env.counters[0].cpu().numpy(), fused["metrics"][P.M_STEPS])
        np.testing.assert_array_equal(env.counters[1].cpu().numpy(), fused["metrics"][P.M_ENV_ON_TARGET])
        np.testing.assert_allclose(-reward.cpu().numpy(), fused["metrics"][P.M_FINAL_ERR], rtol=0, atol=1e-12)
        assert not done.any()


def test_dare_matches_scipy_oracle_on_tuning_ranges(gb, orc):
    """K vs the reference's scipy path for 2 x 1024 candidates drawn from tuning_riccati.yaml's ranges."""
    rng = np.random.default_rng(99)
    lo6, hi6 = np.array([5e-5, 5e-5, 10.0, 1e-3, 1e-3, 2.0]), np.array([5e-4, 5e-4, 25.0, 1e-2, 1e-2, 8.0])
    for lqi in (False, True):
        n = 9 if lqi else 6
        A, B = orc.build_systems_reference(lqi=lqi)
        C_ = 1024
        q = rng.uniform(lo6, hi6, (C_, 6))
        if lqi:
            q = np.concatenate([q, rng.uniform([1e-3, 1e-3, 1e-2], [0.1, 0.1, 1.0], (C_, 3))], axis=1)
        r = rng.uniform(0.5, 2.0, (C_, 4))
        Q = np.stack([np.diag(x) for x in q])
        R = np.stack([np.diag(x) for x in r])
        P_, K, iters, info = gb.dare(A, B, Q, R)
        assert not info.any()
        worst_k = worst_p = 0.0
        for i in range(C_):
            Pr, Kr = orc.solve_dare_reference(A, B, Q[i], R[i])
            worst_k = max(worst_k, np.linalg.norm(K[i] - Kr) / np.linalg.norm(Kr))
            worst_p = max(worst_p, np.linalg.norm(P_[i] - Pr) / np.linalg.norm(Pr))
        print(f"[parity] DARE n={n}: worst rel K {worst_k:.2e}, worst rel P {worst_p:.2e}, iters {iters.min()}..{iters.max()}")
        assert worst_k < 1e-9 and worst_p < 1e-9
        # DARE residual of the device solution (size-independent property)
        for i in range(0, C_, 97):
            X = P_[i]
            G = np.linalg.solve(R[i] + B.T @ X @ B, B.T @ X @ A)
            res = A.T @ X @ A - X - A.T @ X @ B @ G + Q[i]
            assert np.linalg.norm(res) / np.linalg.norm(X) < 1e-11


def test_build_lqr_systems_matches_reference_matrices(golden, gb):
    import torch

    c = gb.ctx()
    g = golden["dare"]
    for n, key in ((6, "lqr_default"), (9, "lqi_default")):
        q = gb.dev(np.array([np.diag(g[key + "_Q"])]))
        r = gb.dev(np.array([[1.0, 1.0, 1.0, 1.0]]))
        f64 = dict(dtype=torch.float64, device=c.device)
        A, B = torch.empty((n, n), **f64), torch.empty((n, 4), **f64)
        Q, R = torch.empty((1, n, n), **f64), torch.empty((1, 4, 4), **f64)
        c.build_lqr_systems(n, 1, 0.01, 1.0, 9.81, q, r, A, B, Q, R)
        np.testing.assert_array_equal(A.cpu().numpy(), g[key + "_A"])
        np.testing.assert_array_equal(B.cpu().numpy(), g[key + "_B"])
        np.testing.assert_array_equal(Q.cpu().numpy()[0], g[key + "_Q"])
        np.testing.assert_array_equal(R.cpu().numpy()[0], g[key + "_R"])


def test_summary_reduction_matches_numpy(gb):
    rng = np.random.default_rng(5)
    E = 100_003
    M = np.zeros((16, E))
    M[P.M_STEPS] = 3000
    M[P.M_MET_ON_TARGET] = rng.integers(2000, 3001, E)
    M[P.M_ENV_ON_TARGET] = M[P.M_MET_ON_TARGET]
    M[P.M_SUM_ERR] = rng.uniform(100, 900, E)
    M[P.M_SUM_EFFORT] = rng.uniform(29000, 30000, E)
    M[P.M_MET_SUCCESS] = M[P.M_MET_ON_TARGET] >= 2400
    s = gb.reduce_summary(M)
    ratios, errs = M[P.M_MET_ON_TARGET] / 3000, M[P.M_SUM_ERR] / 3000
    assert s[0] == E and s[1] == M[P.M_MET_SUCCESS].sum()
    assert abs(s[2] / E - ratios.mean()) < 1e-13 and abs(np.sqrt(s[5] / E) - ratios.std()) < 1e-12
    assert abs(s[3] / E - errs.mean()) < 1e-12 and abs(np.sqrt(s[6] / E) - errs.std()) < 1e-12
    assert int(s[8]) == int(np.argmax(ratios)) and int(s[10]) == int(np.argmin(ratios))
    assert s[7] == ratios.max() and s[9] == ratios.min()


def test_full_size_properties_one_million_linear_episodes(golden, gb):
    """BASELINE config 2 at full size (2^20 episodes x 3000 steps): properties that need no oracle."""
    import torch

    from lqr_quadcopter_test_b200 import RiccatiLQRController, rollout_device
    from lqr_quadcopter_test_b200.metrics import SuccessCriteria

    E = 1 << 20
    env = EnvConfig()
    env.target.motion_type = "linear"
    ctrl = RiccatiLQRController(config={})
    c = gb.ctx()
    seeds = torch.arange(E, dtype=torch.int64, device=c.device) + 42
    out = rollout_device(ctrl, env, SuccessCriteria(), seeds, want_final_state=True)
    M = out["metrics"]
    assert bool((M[P.M_STEPS] == 3000).all()) and bool((M[P.M_TERMINATION] == 1).all())
    assert bool((M[P.M_FINAL_TIME] == M[P.M_FINAL_TIME][0]).all()) and float(M[P.M_FINAL_TIME][0]) >= 30.0
    assert bool(torch.isfinite(out["final_state"]).all())
    # the two on-target accountings differ exactly by the first and the last sample
    s0, tp = out["state0"], out["tparams"]
    err0 = torch.linalg.vector_norm(s0[0:3] - tp[0:3], dim=0)
    shift = (M[P.M_FINAL_ERR] <= 0.5).double() - (err0 <= 0.5).double()
    assert bool((M[P.M_ENV_ON_TARGET] == M[P.M_MET_ON_TARGET] + shift).all())
    # success flag is consistent with the counts (IEEE division on the host: torch's CUDA division by a
    # scalar multiplies by the reciprocal and turns 2400/3000 into 0.7999999999999999)
    Mh = M.cpu().numpy()
    np.testing.assert_array_equal((Mh[P.M_MET_ON_TARGET] / 3000.0 >= 0.8).astype(float), Mh[P.M_MET_SUCCESS])
    # determinism, and independence from batch composition: the first 4096 episodes alone give the same bits
    again = rollout_device(ctrl, env, SuccessCriteria(), seeds, want_final_state=True)
    assert torch.equal(again["metrics"], M) and torch.equal(again["final_state"], out["final_state"])
    small = rollout_device(ctrl, env, SuccessCriteria(), seeds[:4096].contiguous(), want_final_state=True)
    assert torch.equal(small["metrics"], M[:, :4096]) and torch.equal(small["final_state"], out["final_state"][:, :4096])
    # and those 4096 match the reference-generated fixture where they overlap (seeds 42..49)
    g = golden["rollouts"]
    np.testing.assert_allclose(out["final_state"][:, :8].cpu().numpy(), g["lqr_linear_final_state"], rtol=0, atol=gc.TOL_LQR)
    np.testing.assert_array_equal(M[gc.COUNT_ROWS][:, :8].cpu().numpy(), g["lqr_linear_metrics"][gc.COUNT_ROWS])


def _axis_decoupled(K):
    """K with everything outside the z->thrust / y->roll / x->pitch pattern set to exactly zero
    (what the SDA solver returns for the reference's diagonal designs; scipy leaves ~1e-14 there)."""
    K = np.array(K, dtype=np.float64)
    keep = np.zeros_like(K, dtype=bool)
    for c in range(K.shape[1]):
        keep[{2: 0, 1: 1, 0: 2}[c % 3], c] = True
    return np.where(keep, K, 0.0)


@pytest.mark.parametrize("motion", ["linear", "sinusoidal"])
@pytest.mark.parametrize("mode", ["lqr", "lqi", "pid", "lqr_ff_data"])
def test_structure_exploiting_path_is_identical_to_the_general_path(golden, gb, orc, motion, mode):
    """qt_rollout's FAST step loop (zero yaw, axis-decoupled gain) against the general loop forced by
    QT_DEBUG_GENERAL_PATH: every output identical; and against the oracle with the same gain."""
    E = 1024
    env = EnvConfig()
    env.target.motion_type = motion
    kw = {"lqr": {}, "lqi": dict(use_lqi=True), "pid": dict(name="pid", integral_limit=0.3),
          "lqr_ff_data": dict(feedforward_enabled=True, ff_velocity_gain=[0.1, 0.1, 0.0], ff_acceleration_gain=[0.05, 0.05, 0.0])}[mode]
    prm = P.make_params(env, gc.controller_ns(**kw))
    if mode == "pid":
        K = np.array([[0.01, 0.01, 4.0], [0.002, 0.002, 0.3], [0.06, 0.06, 2.0], [0.0, 0.0, 0.0]])
    else:
        K = _axis_decoupled(_gains(golden, mode == "lqi"))
    s0, tp = orc.reset_reference(prm, np.arange(7_000, 7_000 + E))
    extra = dict(traj_stride=500, data_stride=700) if mode == "lqr_ff_data" else dict(traj_stride=500)
    fast = gb.rollout(prm, s0, tp, K, max_steps=3002, **extra)
    gen_prm = prm.copy()
    gen_prm.debug_flags = 1
    general = gb.rollout(gen_prm, s0, tp, K, max_steps=3002, **extra)
    for key in fast:
        np.testing.assert_array_equal(fast[key], general[key], err_msg=f"{mode}/{motion}: {key}")
    ref = orc.rollout(prm, s0, tp, K, max_steps=3002, n_threads=8)
    _compare_rollouts(fast, ref, gc.TOL_LQI if mode == "lqi" else gc.TOL_LQR, f"fast {mode}/{motion}")
    # a warp that holds one episode with a non-zero yaw (or a dense gain) must take the general path
    s0b = s0.copy()
    s0b[8, 5::64] = 0.3
    s0b[11, 9::64] = -0.2
    mixed = gb.rollout(prm, s0b, tp, K, max_steps=600)
    mixed_general = gb.rollout(gen_prm, s0b, tp, K, max_steps=600)
    for key in mixed:
        np.testing.assert_array_equal(mixed[key], mixed_general[key], err_msg=f"{mode}/{motion} mixed: {key}")
    ref_b = orc.rollout(prm, s0b, tp, K, max_steps=600, n_threads=8)
    _compare_rollouts(mixed, ref_b, gc.TOL_LQI if mode == "lqi" else gc.TOL_LQR, f"mixed {mode}/{motion}")


@pytest.mark.parametrize("name,motion,T", [("relu_64_64", "circular", 3002), ("tanh_32", "sinusoidal", 1500),
                                           ("elu_48_40", "figure8", 400), ("leaky_20_20_20", "linear", 400)])
def test_policy_rollout_matches_oracle(golden, gb, orc, name, motion, T):
    """§8(f).2 at a size the fp32 C oracle finishes in seconds: 512 episodes, policy in the loop, the
    Riccati supervisor labelling every 25th step.  The two networks fitted to the LQR law track the
    target (contractive loop) and are compared over long horizons; the two untrained ones fly off and
    amplify fp32 rounding differences exponentially, so they are compared over 4 s."""
    g = golden["policy"]
    E = 512
    pol = gb.make_policy(g, name)
    hidden, act, bounds = gc.POLICIES[name]
    kw = {} if bounds is None else {"bounds": tuple(bounds[k] for k in ("thrust", "roll_rate", "pitch_rate", "yaw_rate"))}
    om = orc.Mlp(*gc.policy_arrays(g, name), act, **kw)
    env = EnvConfig()
    env.target.motion_type = motion
    prm = P.make_params(env, gc.controller_ns())
    s0, tp = orc.reset_reference(prm, np.arange(900, 900 + E))
    K = golden["dare"]["lqr_default_K"]
    ref = orc.rollout_policy(prm, s0, tp, om, K=K, max_steps=T, data_stride=25, n_threads=8)
    got = gb.rollout_policy(prm, s0, tp, pol, K=K, max_steps=T, data_stride=25)
    M, G = got["metrics"], ref["metrics"]
    np.testing.assert_array_equal(M[[P.M_STEPS, P.M_TERMINATION, P.M_VIOLATIONS]], G[[P.M_STEPS, P.M_TERMINATION, P.M_VIOLATIONS]])
    diff = np.abs(M[[P.M_ENV_ON_TARGET, P.M_MET_ON_TARGET]] - G[[P.M_ENV_ON_TARGET, P.M_MET_ON_TARGET]])
    # untrained networks drift through the 0.5 m sphere slowly, so a 1e-6 difference can move a few samples
    assert diff.max() <= 3 and (diff.max(axis=0) > 0).sum() <= E // 20
    # untrained networks do not track: the vehicle drifts 100 m away, where 1e-4 is a few fp64-of-fp32 ulps;
    # the relative term covers that regime, the absolute one the tracking regime
    np.testing.assert_allclose(got["final_state"], ref["final_state"], rtol=1e-5, atol=gc.TOL_POLICY_STATE)
    np.testing.assert_allclose(M[gc.SUM_ROWS], G[gc.SUM_ROWS], rtol=2e-5, atol=1e-3)
    np.testing.assert_allclose(got["features"], ref["features"], rtol=1e-5, atol=gc.TOL_POLICY_STATE)
    np.testing.assert_allclose(got["actions"], ref["actions"], rtol=0, atol=gc.TOL_POLICY_STATE)
    print(f"[parity] policy {name}/{motion}: max |final state diff| = {np.abs(got['final_state'] - ref['final_state']).max():.2e}, "
          f"episodes with an on-target count differing = {(diff.max(axis=0) > 0).sum()}")


@pytest.mark.parametrize("hidden,act,motion", [([64, 64], "relu", "linear"), ([64], "tanh", "circular"),
                                               ([128, 64], "elu", "sinusoidal"), ([64, 64, 64], "leaky_relu", "figure8"),
                                               ([60, 120], "relu", "stationary")])
def test_tiled_policy_evaluation_is_identical_to_per_thread(golden, gb, orc, hidden, act, motion):
    """qt_rollout_policy's warp-cooperative (tiled) network evaluation against the one-episode-per-thread
    evaluation forced through qt_params.debug_flags: same sums in the same order, so every output must have
    the same bits.  1000 episodes (a partial last warp), random networks (episodes leave the box and end at
    different steps, so warps run with finished lanes), supervisor labels and both logs on."""
    import copy

    import torch

    from lqr_quadcopter_test_b200 import DeepTrackingPolicy

    torch.manual_seed(len(hidden) * 1000 + hidden[0])
    pol = DeepTrackingPolicy(config={"hidden_sizes": hidden, "activation": act})
    E = 1000
    env = EnvConfig()
    env.target.motion_type = motion
    env.simulation.max_position = 30.0   # random networks drift away: episodes end at different steps
    prm = P.make_params(env, gc.controller_ns())
    per_thread = copy.copy(prm)
    per_thread.debug_flags = 2
    s0, tp = orc.reset_reference(prm, np.arange(300, 300 + E))
    K = golden["dare"]["lqr_default_K"]
    a = gb.rollout_policy(prm, s0, tp, pol, K=K, max_steps=1200, data_stride=7, traj_stride=50)
    b = gb.rollout_policy(per_thread, s0, tp, pol, K=K, max_steps=1200, data_stride=7, traj_stride=50)
    for key in a:
        np.testing.assert_array_equal(a[key], b[key], err_msg=f"{hidden}/{act}/{motion}: {key}")
    steps = a["metrics"][P.M_STEPS]
    print(f"[parity] tiled policy {hidden}/{act}/{motion}: episode lengths {int(steps.min())}..{int(steps.max())}, "
          f"{int((steps < 1200).sum())} of {E} ended early")


def test_randomised_configurations_match_oracle(golden, gb, orc):
    """Twenty random env / controller parameter sets (mass, drag, dt, limits, target pattern sizes and
    speeds, gains, feedforward, integrator settings, all three control laws) against the oracle: the
    flattened RK4 coefficients, the carried trig and the integer-pipe range tests all depend on these."""
    import os

    # QT_SOAK_TRIALS / QT_SOAK_SEED widen the sweep for an occasional soak run (round 1: 600 sets, no failure)
    rng = np.random.default_rng(int(os.environ.get("QT_SOAK_SEED", 20260101)))
    worst = 0.0
    n_trials = int(os.environ.get("QT_SOAK_TRIALS", 20))
    for trial in range(n_trials):
        motion = MOTIONS[trial % 5]
        law = ("lqr", "lqi", "pid")[trial % 3]
        env = EnvConfig()
        env.target.motion_type = motion
        env.quadcopter.mass = float(rng.uniform(0.6, 1.8))
        env.quadcopter.drag_coeff_linear = float(rng.uniform(0.02, 0.4))
        env.quadcopter.drag_coeff_angular = float(rng.uniform(0.0, 0.05))
        env.simulation.dt = float(rng.choice([0.005, 0.01, 0.02]))
        env.simulation.max_episode_time = float(rng.uniform(3.0, 6.0))
        env.simulation.max_velocity = float(rng.uniform(4.0, 50.0))
        env.simulation.max_angular_velocity = float(rng.uniform(1.0, 10.0))
        env.simulation.max_position = float(rng.choice([6.0, 1000.0]))
        env.target.speed = float(rng.uniform(0.3, 2.5))
        env.target.radius = float(rng.uniform(1.0, 3.0))
        env.target.amplitude = float(rng.uniform(0.5, 2.5))
        env.target.frequency = float(rng.uniform(0.2, 0.9))
        env.target.center = tuple(rng.uniform(-1.0, 2.0, 3))
        env.success_criteria.target_radius = float(rng.uniform(0.2, 0.8))
        ff = bool(trial % 2)
        kw = dict(dt=env.simulation.dt, hover_thrust=env.quadcopter.mass * 9.81, max_rate=float(rng.uniform(1.5, 3.0)),
                  feedforward_enabled=ff, ff_velocity_gain=list(rng.uniform(0, 0.3, 3)),
                  ff_acceleration_gain=list(rng.uniform(0, 0.2, 3)), ff_max_velocity=float(rng.uniform(0.5, 3.0)),
                  ff_max_acceleration=float(rng.uniform(0.3, 3.0)))
        if law == "pid":
            kw.update(name="pid", integral_limit=float(rng.choice([0.0, 0.3])))
            K = np.array([[0.01, 0.01, 4.0], [0.003, 0.003, 0.4], [0.06, 0.06, 2.0], [0, 0, 0]]) * rng.uniform(0.5, 1.5, (4, 3))
        elif law == "lqi":
            kw.update(use_lqi=True, integral_limit=float(rng.uniform(0.05, 2.0)), integral_zero_threshold=float(rng.uniform(0.0, 0.1)))
            K = _axis_decoupled(golden["dare"]["lqi_default_K"]) * rng.uniform(0.6, 1.2, (4, 9))
            K[:, 6:] *= 0.05     # a mild integral gain: the reference's default one diverges
            if trial % 4 == 1:
                K = K + rng.standard_normal((4, 9)) * 1e-3    # dense gain: the general loop
        else:
            K = _axis_decoupled(golden["dare"]["lqr_default_K"]) * rng.uniform(0.6, 1.4, (4, 6))
            if trial % 4 == 0:
                K = K + rng.standard_normal((4, 6)) * 1e-3
        crit = type("C", (), dict(target_radius=env.success_criteria.target_radius, min_on_target_ratio=0.5,
                                   min_episode_duration=env.simulation.max_episode_time))()
        prm = P.make_params(env, gc.controller_ns(**kw), crit)
        E = 96
        s0, tp = orc.reset_reference(prm, np.arange(5000 + 100 * trial, 5000 + 100 * trial + E))
        if trial % 5 == 3:
            s0[8] = rng.uniform(-0.5, 0.5, E)     # non-zero yaw: general loop
        steps = int(np.ceil(env.simulation.max_episode_time / env.simulation.dt)) + 2
        ref = orc.rollout(prm, s0, tp, K, max_steps=steps, n_threads=8, traj_stride=97)
        got = gb.rollout(prm, s0, tp, K, max_steps=steps, traj_stride=97)
        # figure-8 + feedforward: the reference's target acceleration is a finite difference with h = 1e-6
        # (target_motion.py:209-229) that amplifies a 1-ulp sin/cos difference to ~1e-4, and feedforward
        # feeds it into the action - two libm builds of the reference itself would differ as much
        noisy = motion == "figure8" and ff
        tol = 1e-3 if noisy else 1e-9    # (a 600-set soak run saw 2.1e-4 in the noisy case, <= 1e-12 elsewhere)
        label = f"trial {trial} {law}/{motion} ff={ff}"
        _compare_rollouts(got, ref, tol, label)
        np.testing.assert_allclose(got["traj"], ref["traj"], rtol=0, atol=5e-3 if noisy else tol, err_msg=label)
        if not noisy:
            worst = max(worst, float(np.abs(got["final_state"] - ref["final_state"]).max()))
    print(f"[parity] {n_trials} randomised configurations: max |final state diff| = {worst:.2e} (figure-8 + feedforward excluded)")


@pytest.mark.parametrize("hidden,act,motion", [([64, 64], "relu", "linear"), ([48, 64], "tanh", "circular"),
                                               ([64, 20], "elu", "stationary"), ([64, 64], "leaky_relu", "figure8"),
                                               ([32, 32], "relu", "sinusoidal"), ([64], "tanh", "linear"),
                                               ([20, 40, 60], "elu", "circular"), ([16], "relu", "stationary")])
def test_tensor_core_policy_evaluation_is_close_to_the_fp32_pipe(golden, gb, orc, hidden, act, motion):
    """QT_FLAG_POLICY_TENSOR_CORES: the two hidden layers as tcgen05 TF32 MMAs with the 3xTF32 split (x = hi + lo,
    fp32 accumulation in TMEM) against the FP32-pipe evaluation.  Not bit-identical by construction: each product
    carries about 2^-21 relative error.  Bounds: the action on the SAME observation (first step of every episode)
    within 4e-6 of the action range; a 40-step rollout within 1e-5 in every state; the logged teacher labels (which
    depend on the visited states only) within 1e-4.  1000 episodes: a partial last block, finished and shadow lanes."""
    import copy

    import torch

    from lqr_quadcopter_test_b200 import DeepTrackingPolicy

    torch.manual_seed(len(hidden) * 77 + hidden[-1])
    pol = DeepTrackingPolicy(config={"hidden_sizes": hidden, "activation": act})
    E = 1000
    env = EnvConfig()
    env.target.motion_type = motion
    prm = P.make_params(env, gc.controller_ns())
    tcp = copy.copy(prm)
    tcp.debug_flags = 4
    s0, tp = orc.reset_reference(prm, np.arange(900, 900 + E))
    K = golden["dare"]["lqr_default_K"]
    a = gb.rollout_policy(prm, s0, tp, pol, K=K, max_steps=40, data_stride=7, traj_stride=1)
    b = gb.rollout_policy(tcp, s0, tp, pol, K=K, max_steps=40, data_stride=7, traj_stride=1)
    act0_a, act0_b = a["traj"][0, P.TRAJ_ACTION:P.TRAJ_ACTION + 4], b["traj"][0, P.TRAJ_ACTION:P.TRAJ_ACTION + 4]
    scale = np.array([20.0, 6.0, 6.0, 6.0])[:, None]
    d_act = float((np.abs(act0_a - act0_b) / scale).max())
    d_state = float(np.abs(a["final_state"] - b["final_state"]).max())
    print(f"[parity] tensor-core policy {hidden}/{act}/{motion}: first action diff {d_act:.2e} of range, "
          f"final state diff after 40 steps {d_state:.2e}, identical first actions {int((act0_a == act0_b).all(axis=0).sum())}/{E}")
    assert np.isfinite(b["final_state"]).all() and (b["metrics"][P.M_STEPS] == a["metrics"][P.M_STEPS]).all()
    assert 0.0 < d_act <= 4e-6        # > 0: the tensor-core path really ran (the FP32 paths are bit-identical to each other)
    assert d_state <= 1e-5
    np.testing.assert_array_equal(a["features"][0], b["features"][0])
    np.testing.assert_allclose(a["actions"], b["actions"], rtol=0, atol=1e-4)
