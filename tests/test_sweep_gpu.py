"""GPU: the candidate-sweep entry points (BASELINE config 4) — qt_rollout_sweep's shared initial conditions and fused
per-candidate reduction, the host-buffer calls qt_sweep_host / qt_dare_host, and the tuner's array-native path with
the reference's per-candidate fallback rule.  Reference: controllers/tuning.py:846-928, riccati_lqr.py:737-777."""
import numpy as np
import pytest
import torch

import golden_cases as gc
from lqr_quadcopter_test_b200 import params as P
from lqr_quadcopter_test_b200.config import EnvConfig

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gb(gpu_ctx):
    import gpu_backend

    return gpu_backend


@pytest.fixture(scope="module")
def orc():
    import oracle

    return oracle


def _candidates(C, lqi=False, seed=7):
    rng = np.random.default_rng(seed)
    lo = [5e-5, 5e-5, 10, 1e-3, 1e-3, 2] + ([1e-6, 1e-6, 1e-4] if lqi else [])
    hi = [5e-4, 5e-4, 25, 1e-2, 1e-2, 8] + ([1e-5, 1e-5, 1e-2] if lqi else [])
    return rng.uniform(lo, hi, (C, len(lo))), rng.uniform(0.5, 2.0, (C, 4))


def _gains(ctx, q, r):
    C, n = q.shape
    f64 = dict(dtype=torch.float64, device=ctx.device)
    A, B = torch.empty((n, n), **f64), torch.empty((n, 4), **f64)
    Q, R = torch.empty((C, n, n), **f64), torch.empty((C, 4, 4), **f64)
    ctx.build_lqr_systems(n, C, 0.01, 1.0, 9.81, torch.as_tensor(q, **f64), torch.as_tensor(r, **f64), A, B, Q, R)
    K = torch.empty((C, 4, n), **f64)
    iters = torch.zeros(C, dtype=torch.int32, device=ctx.device)
    info = torch.zeros(C, dtype=torch.int32, device=ctx.device)
    ctx.dare_batched(n, 4, C, A, 0, B, 0, Q, n * n, R, 16, None, K, iters, info)
    assert not bool(info.any())
    return K


@pytest.mark.parametrize("C,eps,lqi,motion", [(37, 64, False, "stationary"), (3, 64, False, "linear"), (5, 128, False, "circular"),
                                              (300, 8, True, "stationary"), (129, 1, False, "linear"),
                                              (23, 5, False, "stationary"), (4, 200, False, "sinusoidal")])
def test_sweep_equals_rollout_plus_reduction(gb, orc, C, eps, lqi, motion):
    """qt_rollout_sweep (shared initial conditions, reduction fused when eps | 128) against qt_rollout on tiled
    initial conditions followed by qt_reduce_scores: identical bits, metrics rows included; and the scores against
    the oracle's reduction of the oracle's episodes."""
    ctx = gb.ctx()
    env = EnvConfig()
    env.target.motion_type = motion
    prm = P.make_params(env, gc.controller_ns(use_lqi=lqi))
    q, r = _candidates(C, lqi)
    K = _gains(ctx, q, r)
    k_cols = 9 if lqi else 6
    T = 700
    s1, t1 = orc.reset_reference(prm, np.arange(42, 42 + eps))
    f64 = dict(dtype=torch.float64, device=ctx.device)
    s1d, t1d = gb.dev(s1), gb.dev(t1)
    cand = torch.full((4, C), -7.0, **f64)
    ctx.rollout_sweep(prm, C, eps, T, s1d, t1d, K, k_cols, cand)
    cand_m = torch.full((4, C), -7.0, **f64)
    met = torch.zeros((16, C * eps), **f64)
    ctx.rollout_sweep(prm, C, eps, T, s1d, t1d, K, k_cols, cand_m, met)
    # the round-1 formulation: tile the initial conditions, write every metric row, reduce in a second kernel
    tiled = gb.rollout(prm, np.tile(s1, (1, C)), np.tile(t1, (1, C)), K.cpu().numpy(), k_stride=eps, max_steps=T)
    want = gb.reduce_scores(tiled["metrics"], C, eps)
    np.testing.assert_array_equal(met.cpu().numpy(), tiled["metrics"])
    np.testing.assert_array_equal(cand.cpu().numpy(), want)
    np.testing.assert_array_equal(cand_m.cpu().numpy(), want)
    only_rows = torch.zeros((16, C * eps), **f64)
    ctx.rollout_sweep(prm, C, eps, T, s1d, t1d, K, k_cols, None, only_rows)
    np.testing.assert_array_equal(only_rows.cpu().numpy(), tiled["metrics"])
    ref = orc.rollout(prm, np.tile(s1, (1, C)), np.tile(t1, (1, C)), K.cpu().numpy(), k_stride=eps, max_steps=T, n_threads=8)
    sc = orc.reduce_scores(ref["metrics"], C, eps)
    np.testing.assert_array_equal(want[[1, 3]], sc[[1, 3]])
    np.testing.assert_allclose(want[[0, 2]], sc[[0, 2]], rtol=0, atol=1e-8 if lqi else 1e-11)


def test_sweep_host_matches_the_device_path(gb, orc):
    """qt_sweep_host: host weights in, scores / gains / DARE info out, against the device-resident calls."""
    ctx = gb.ctx()
    C, eps, T = 1000, 64, 400
    env = EnvConfig()
    prm = P.make_params(env, gc.controller_ns())
    q, r = _candidates(C)
    qh, rh = torch.as_tensor(q).pin_memory(), torch.as_tensor(r).pin_memory()
    cand_h = torch.empty((4, C), dtype=torch.float64).pin_memory()
    K_h = torch.empty((C, 4, 6), dtype=torch.float64).pin_memory()
    info_h = torch.empty(C, dtype=torch.int32).pin_memory()
    ctx.sweep_host(prm, 6, C, eps, 42, T, qh, rh, 1.0, 9.81, cand_h, K_h, info_h)
    assert not bool(info_h.any())
    K = _gains(ctx, q, r)
    np.testing.assert_array_equal(K_h.numpy(), K.cpu().numpy())
    s1, t1 = gb.reset(prm, np.arange(42, 42 + eps))
    cand = torch.empty((4, C), dtype=torch.float64, device=ctx.device)
    ctx.rollout_sweep(prm, C, eps, T, gb.dev(s1), gb.dev(t1), K, 6, cand)
    np.testing.assert_array_equal(cand_h.numpy(), cand.cpu().numpy())
    # a rejected candidate is reported, the others are unaffected
    r2 = r.copy()
    r2[17, 3] = 0.0
    ctx.sweep_host(prm, 6, C, eps, 42, T, qh, torch.as_tensor(r2).pin_memory(), 1.0, 9.81, cand_h, None, info_h)
    assert int(info_h[17]) == 3 and int((info_h != 0).sum()) == 1
    keep = np.arange(C) != 17
    np.testing.assert_array_equal(cand_h.numpy()[:, keep], cand.cpu().numpy()[:, keep])


@pytest.mark.parametrize("n,m,batch", [(6, 4, 10000), (9, 4, 5000), (12, 4, 3000), (15, 4, 100), (6, 4, 7)])
def test_dare_host_matches_the_device_path(gb, n, m, batch):
    from lqr_quadcopter_test_b200 import solve_dare_batched

    ctx = gb.ctx()
    rng = np.random.default_rng(n * 100 + batch)
    A = np.eye(n) + 0.01 * np.triu(rng.standard_normal((n, n)), 1)
    B = 0.05 * rng.standard_normal((n, m))
    Mx = rng.standard_normal((batch, n, n))
    Q = Mx @ Mx.transpose(0, 2, 1) * 0.05 + np.eye(n) * rng.uniform(0.1, 2.0, (batch, 1, 1))
    R = np.stack([np.diag(x) for x in rng.uniform(0.5, 2.0, (batch, m))])
    Pd, Kd, itd, infod = solve_dare_batched(A, B, Q, R)
    pin = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).pin_memory()
    Ph, Kh = torch.empty((batch, n, n), dtype=torch.float64).pin_memory(), torch.empty((batch, m, n), dtype=torch.float64).pin_memory()
    ith, infoh = torch.empty(batch, dtype=torch.int32).pin_memory(), torch.empty(batch, dtype=torch.int32).pin_memory()
    ctx.dare_host(n, m, batch, pin(A), 0, pin(B), 0, pin(Q), n * n, pin(R), m * m, Ph, Kh, ith, infoh)
    np.testing.assert_array_equal(Kh.numpy(), Kd.cpu().numpy())
    np.testing.assert_array_equal(Ph.numpy(), Pd.cpu().numpy())
    np.testing.assert_array_equal(ith.numpy(), itd.cpu().numpy())
    np.testing.assert_array_equal(infoh.numpy(), infod.cpu().numpy())
    assert not bool(infoh.any())
    # per-system A and B, no P wanted
    A2 = np.broadcast_to(A, (batch, n, n)).copy()
    B2 = np.broadcast_to(B, (batch, n, m)).copy()
    K2 = torch.empty((batch, m, n), dtype=torch.float64).pin_memory()
    ctx.dare_host(n, m, batch, pin(A2), n * n, pin(B2), n * m, pin(Q), n * n, pin(R), m * m, None, K2, None, infoh)
    np.testing.assert_array_equal(K2.numpy(), Kd.cpu().numpy())


def test_failing_candidate_inside_a_batch_follows_the_reference_fallback(gpu_ctx, tmp_path):
    """a candidate whose DARE cannot be solved must not abort the batch nor keep the batch's parameters:
    the reference evaluates it with the heuristic LQR fallback built WITHOUT feedforward and without integral action
    (riccati_lqr.py:737-777).  The batch result must equal the one-by-one result for every candidate."""
    from lqr_quadcopter_test_b200 import ControllerTuner, GainSearchSpace, LQRController, TuningConfig

    for lqi in (False, True):
        extra = {"use_lqi": True, "q_int": [1e-5, 1e-5, 1e-3]} if lqi else {}
        good1 = {"q_pos": [1e-4, 1e-4, 16.0], "q_vel": [3.6e-3, 3.6e-3, 4.0], "r_controls": [1.0, 1.0, 1.0, 1.0], **extra}
        bad = {"q_pos": [2e-4, 1e-4, 12.0], "q_vel": [3e-3, 3e-3, 5.0], "r_controls": [1.0, 1.5, 0.5, 0.0], **extra}  # R singular
        good2 = {"q_pos": [3e-4, 2e-4, 20.0], "q_vel": [5e-3, 4e-3, 3.0], "r_controls": [0.8, 1.2, 1.1, 0.9], **extra}
        tc = TuningConfig(controller_type="riccati_lqr", search_space=GainSearchSpace(), evaluation_episodes=4,
                          evaluation_horizon=500, seed=42, output_dir=str(tmp_path), target_motion_type="linear")
        tuner = ControllerTuner(tc)
        scores, metrics = tuner.evaluate_configs([good1, bad, good2])
        assert tuner.last_dare_info.cpu().tolist() == [0, 3, 0]      # the batch's DARE info: candidate 1 rejected (R not PD)
        for i, c in enumerate((good1, bad, good2)):
            s1, m1 = ControllerTuner(tc)._evaluate_config(c)
            assert scores[i] == s1 and metrics[i] == m1, (lqi, i)
        # ... and the failing one is the heuristic LQR on the same weights
        fb = LQRController(config={"q_pos": bad["q_pos"], "q_vel": bad["q_vel"], "r_thrust": 1.0,
                                   "r_rate": (1.5 + 0.5 + 0.0) / 3})
        tl = ControllerTuner(TuningConfig(controller_type="lqr", search_space=GainSearchSpace(), evaluation_episodes=4,
                                          evaluation_horizon=500, seed=42, output_dir=str(tmp_path), target_motion_type="linear"))
        s_fb, _ = tl._evaluate_config({"q_pos": bad["q_pos"], "q_vel": bad["q_vel"], "r_thrust": 1.0, "r_rate": (1.5 + 0.5 + 0.0) / 3})
        assert scores[1] == s_fb and np.isfinite(fb.K).all()
