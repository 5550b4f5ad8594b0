"""GPU: BASELINE.json configs 3 and 4 at their FULL sizes, checked through properties that need no
oracle (the oracle needs minutes at these sizes): determinism, independence from batch composition,
the exact relation between the env's and the metrics' on-target accounting, success flags against
counts, Riccati residuals and closed-loop stability of every gain, duplicated candidates, and the
reference-generated fixtures wherever the seeds overlap.  Config 2 has its own test in
test_parity_gpu.py; config 5 (teacher data set) is covered in test_dropin_gpu.py and here by a
row-level identity at 2^20 episodes."""
import numpy as np
import pytest

import golden_cases as gc
from lqr_quadcopter_test_b200 import params as P
from lqr_quadcopter_test_b200.config import EnvConfig

pytestmark = pytest.mark.gpu

LQI = {"use_lqi": True, "q_int": [0.01, 0.01, 0.1]}


@pytest.mark.parametrize("motion", ["circular", "sinusoidal"])
def test_full_size_config3_lqi_half_million_episodes_per_motion(golden, gpu_ctx, motion):
    """BASELINE config 3: LQI (9-state, integral clamp), 2^20 episodes, half circular / half sinusoidal."""
    import torch

    from lqr_quadcopter_test_b200 import RiccatiLQRController, rollout_device
    from lqr_quadcopter_test_b200.metrics import SuccessCriteria

    E = 1 << 19
    env = EnvConfig()
    env.target.motion_type = motion
    ctrl = RiccatiLQRController(config=dict(LQI))
    dev = gpu_ctx.device
    seeds = torch.arange(E, dtype=torch.int64, device=dev) + 42
    out = rollout_device(ctrl, env, SuccessCriteria(), seeds, want_final_state=True)
    M = out["metrics"]
    assert bool(torch.isfinite(M).all()) and bool(torch.isfinite(out["final_state"]).all())
    # the default LQI gains saturate but stay inside the +-1000 m box: every episode runs its 30 s
    assert bool((M[P.M_STEPS] == 3000).all()) and bool((M[P.M_TERMINATION] == 1).all())
    assert bool((M[P.M_FINAL_TIME] == M[P.M_FINAL_TIME][0]).all()) and float(M[P.M_FINAL_TIME][0]) >= 30.0
    # reset quirk (quadcopter_env.py:111-150 + target_motion.py:329): both generators start from the
    # same seed, so the circular start angle and the x offset come from the same uniform draw
    s0, tp = out["state0"], out["tparams"]
    t0 = _target_position_at_zero(env, tp, motion)
    off = s0[0:3] - t0
    assert float(off.abs().max()) <= 0.5 + 1e-12
    if motion == "circular":
        u1 = off[0] + 0.5
        angle0 = torch.atan2(t0[1] - env.target.center[1], t0[0] - env.target.center[0]) % (2 * np.pi)
        d = (angle0 - 2 * np.pi * u1).abs()
        assert float(torch.minimum(d, 2 * np.pi - d).max()) < 1e-9
    # env accounting (post-step samples 1..T) against metrics accounting (pre-step samples 0..T-1)
    err0 = torch.linalg.vector_norm(off, dim=0)
    shift = (M[P.M_FINAL_ERR] <= 0.5).double() - (err0 <= 0.5).double()
    assert bool((M[P.M_ENV_ON_TARGET] == M[P.M_MET_ON_TARGET] + shift).all())
    Mh = M.cpu().numpy()
    np.testing.assert_array_equal((Mh[P.M_MET_ON_TARGET] / 3000.0 >= 0.8).astype(float), Mh[P.M_MET_SUCCESS])
    np.testing.assert_array_equal((Mh[P.M_ENV_ON_TARGET] / 3000.0 >= 0.8).astype(float), Mh[P.M_ENV_SUCCESS])
    assert (Mh[P.M_MAX_ERR] * 3000 >= Mh[P.M_SUM_ERR] * (1 - 1e-12)).all()
    assert (Mh[P.M_SUM_ERR2] * 3000 >= Mh[P.M_SUM_ERR] ** 2 * (1 - 1e-12)).all()      # Cauchy-Schwarz
    assert (Mh[P.M_SUM_EFFORT] <= 3000 * np.sqrt(20.0 ** 2 + 3 * 3.0 ** 2) * (1 + 1e-12)).all()
    # determinism and independence from the batch the episode travels in
    again = rollout_device(ctrl, env, SuccessCriteria(), seeds, want_final_state=True)
    assert torch.equal(again["metrics"], M) and torch.equal(again["final_state"], out["final_state"])
    sub = torch.arange(1000, 1000 + 4096, dtype=torch.int64, device=dev) + 42
    small = rollout_device(ctrl, env, SuccessCriteria(), sub, want_final_state=True)
    assert torch.equal(small["metrics"], M[:, 1000:1000 + 4096])
    assert torch.equal(small["final_state"], out["final_state"][:, 1000:1000 + 4096])
    # the reference's own numbers where the seeds overlap (42..45)
    g = golden["rollouts"]
    np.testing.assert_allclose(out["final_state"][:, :4].cpu().numpy(), g[f"lqi_{motion}_final_state"], rtol=0,
                               atol=gc.TOL_LQI)
    np.testing.assert_array_equal(Mh[gc.COUNT_ROWS][:, :4], g[f"lqi_{motion}_metrics"][gc.COUNT_ROWS])


def _target_position_at_zero(env, tp, motion):
    """Target position at t = 0 from the packed pattern parameters (include/qtrack.h, target layout)."""
    import torch

    from lqr_quadcopter_test_b200 import _lib

    ctx = _lib.get_context()
    E = tp.shape[1]
    from lqr_quadcopter_test_b200 import RiccatiLQRController
    from lqr_quadcopter_test_b200.metrics import SuccessCriteria

    prm = RiccatiLQRController(config={}).kernel_params(env, SuccessCriteria())
    obs = torch.empty((P.OBS_DIM, E), dtype=torch.float64, device=tp.device)
    state = torch.zeros((12, E), dtype=torch.float64, device=tp.device)
    time = torch.zeros(E, dtype=torch.float64, device=tp.device)
    ctx.observe_batched(prm, E, state, time, tp, obs)
    return obs[12:15]


def test_full_size_config4_autotune_sweep_65536_candidates(gpu_ctx, tmp_path):
    """BASELINE config 4: 65 536 candidates x 64 episodes x 3000 steps, one DARE per candidate."""
    import torch

    from lqr_quadcopter_test_b200 import ControllerTuner, GainSearchSpace, TuningConfig

    C, eps = 65536, 64
    space = GainSearchSpace(q_pos_range=([5e-5, 5e-5, 10.0], [5e-4, 5e-4, 25.0]),
                            q_vel_range=([1e-3, 1e-3, 2.0], [1e-2, 1e-2, 8.0]), r_controls_range=([0.5] * 4, [2.0] * 4))
    tc = TuningConfig(controller_type="riccati_lqr", search_space=space, strategy="random", max_iterations=C,
                      evaluation_episodes=eps, evaluation_horizon=3000, seed=42, output_dir=str(tmp_path))
    tuner = ControllerTuner(tc)
    half = [tuner._generate_random_config() for _ in range(C // 2)]
    configs = half + half[::-1]                       # every candidate appears twice, mirrored
    fused_scores, fused_metrics = tuner.evaluate_configs(configs)     # production path: reduction fused into the rollout
    assert tuner.last_metrics is None                                 # ... and no per-episode rows written
    tuner.keep_episode_metrics = True                                 # same launch, rows written as well
    scores, metrics = tuner.evaluate_configs(configs)
    assert scores == fused_scores and metrics == fused_metrics       # bit-identical either way
    s = np.array(scores)
    assert len(scores) == C and np.isfinite(s).all()
    np.testing.assert_array_equal(s[:C // 2], s[C // 2:][::-1])
    # the array-native sweep (ControllerTuner.sweep) draws the same candidates and scores them identically
    tuner2 = ControllerTuner(tc)
    block, rows = tuner2.sweep(C // 2)
    assert block.configs()[:3] == half[:3] and block.config(C // 2 - 1) == half[-1]
    np.testing.assert_array_equal(rows[0], s[:C // 2])
    assert bool((tuner.last_dare_info == 0).all())
    it = tuner.last_dare_iters.cpu().numpy()
    assert it.min() >= 8 and it.max() <= 24           # doubling converges quadratically on these ranges
    M = tuner.last_metrics
    assert M.shape == (16, C * eps) and bool((M[P.M_STEPS] == 3000).all())
    # score = mean ratio - 0.1 mean error (tuning.py:910-918), from the per-episode metrics, in numpy
    Mh = M[[P.M_MET_ON_TARGET, P.M_SUM_ERR, P.M_MET_SUCCESS]].cpu().numpy().reshape(3, C, eps)
    ratio = (Mh[0] / 3000.0).mean(axis=1)
    err = (Mh[1] / 3000.0).mean(axis=1)
    np.testing.assert_allclose(s, ratio - 0.1 * err, rtol=0, atol=1e-12)
    np.testing.assert_allclose([m["success_rate"] for m in metrics], Mh[2].mean(axis=1), rtol=0, atol=1e-15)
    # a slice of the sweep evaluated on its own gives the same bits (no dependence on the batch)
    lo = 12345
    s_small, _ = tuner.evaluate_configs(configs[lo:lo + 512])
    np.testing.assert_array_equal(np.array(s_small), s[lo:lo + 512])
    # first-wins tie-break of tuning.py:1084 on the mirrored list: the best index lies in the first half
    best = int(np.argmax(s))
    assert best < C // 2 and s[best] == s.max()

    # every gain solves its Riccati equation and stabilises its loop
    from lqr_quadcopter_test_b200 import solve_dare_batched

    q = np.array([list(c["q_pos"]) + list(c["q_vel"]) for c in half])
    r = np.array([c["r_controls"] for c in half])
    n, dt, g_ = 6, 0.01, 9.81
    A = np.eye(n)
    A[0:3, 3:6] = dt * np.eye(3)
    B = np.zeros((n, 4))
    B[5, 0] = dt / 1.0
    B[4, 1] = -g_ * dt
    B[3, 2] = g_ * dt
    f64 = dict(dtype=torch.float64, device=gpu_ctx.device)
    Ad, Bd = torch.as_tensor(A, **f64), torch.as_tensor(B, **f64)
    Qd, Rd = torch.diag_embed(torch.as_tensor(q, **f64)), torch.diag_embed(torch.as_tensor(r, **f64))
    Pm, K, iters, info = solve_dare_batched(Ad, Bd, Qd, Rd)
    assert bool((info == 0).all())
    res = Ad.T @ Pm @ Ad - Pm - (Ad.T @ Pm @ Bd) @ K + Qd
    rel = torch.linalg.matrix_norm(res) / torch.linalg.matrix_norm(Pm)
    assert float(rel.max()) < 1e-11, float(rel.max())
    assert float((Pm - Pm.transpose(1, 2)).abs().max()) == 0.0
    assert bool((K[:, 3, :] == 0).all())                                   # B[:, 3] = 0: no yaw authority
    Acl = (Ad - Bd @ K).cpu().numpy()
    rho = np.abs(np.linalg.eigvals(Acl)).max(axis=1)
    assert rho.max() < 1.0, rho.max()


def test_full_size_config5_teacher_rows_are_the_controller_on_the_logged_states(gpu_ctx):
    """BASELINE config 5 at 2^20 episodes of the 4M (stride 10, 300 rows per episode, 27.8 GB): every
    emitted action row equals the control law applied to its own feature row (fp32 rounding only)."""
    import torch

    from lqr_quadcopter_test_b200 import RiccatiLQRController, rollout_device
    from lqr_quadcopter_test_b200.metrics import SuccessCriteria

    E = 1 << 20
    env = EnvConfig()
    env.target.motion_type = "linear"
    ctrl = RiccatiLQRController(config={})
    seeds = torch.arange(E, dtype=torch.int64, device=gpu_ctx.device) + 42
    out = rollout_device(ctrl, env, SuccessCriteria(), seeds, max_steps=3000, dataset_stride=10)
    F, Y = out["features"], out["actions"]
    assert F.shape == (300, E, 18) and Y.shape == (300, E, 4) and F.dtype == torch.float32
    assert bool((out["metrics"][P.M_STEPS] == 3000).all())
    K = torch.as_tensor(np.ascontiguousarray(ctrl.effective_gain()), dtype=torch.float64, device=F.device)
    worst = 0.0
    for lo in range(0, 300, 50):                      # 50 slots at a time: keep the fp64 temporaries small
        e = F[lo:lo + 50, :, 0:6].double()
        u = e @ K.T
        u[..., 0] += ctrl.hover_thrust
        u[..., 0].clamp_(0.0, 20.0)
        u[..., 1:].clamp_(-3.0, 3.0)
        worst = max(worst, float((u - Y[lo:lo + 50].double()).abs().max()))
        del e, u
    # features are fp32-rounded errors (relative 6e-8) times gains of order 4: a few 1e-6 at most
    assert worst < 2e-5, worst
    # first row of every episode is the reset observation
    s0, tp = out["state0"], out["tparams"]
    np.testing.assert_array_equal(F[0, :, 6:12].cpu().numpy(), np.zeros((E, 6), dtype=np.float32))
    assert float((F[0, :, 0:3].double() - (F[0, :, 12:15].double() - s0[0:3].T)).abs().max()) < 1e-6
