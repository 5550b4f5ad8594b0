"""GPU: the physical and control-law behaviours the reference's own test suite asserts
(tests/test_env_dynamics.py: TestTargetMotion, TestQuadcopterEnv, TestIntegration, TestAxisSignConventions,
TestRiccatiLQRController, TestRiccatiLQRFeedforward, TestRiccatiLQRLQI), restated for the drop-in classes.
Numbers come from first principles or from the reference's documented conventions, not from fixtures: the
fixtures (tests/golden) pin the arithmetic, these pin the meaning."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _ctx(gpu_ctx):
    return gpu_ctx


def _env(motion="stationary", seconds=30.0, **target):
    from lqr_quadcopter_test_b200 import EnvConfig, QuadcopterEnv

    cfg = EnvConfig()
    cfg.target.motion_type = motion
    for k, v in target.items():
        setattr(cfg.target, k, v)
    cfg.simulation.max_episode_time = seconds
    return QuadcopterEnv(config=cfg), cfg


def _obs(err=(0.0, 0.0, 0.0), tvel=(0.0, 0.0, 0.0), tacc=(0.0, 0.0, 0.0), qvel=(0.0, 0.0, 0.0)):
    return {"quadcopter": {"position": np.zeros(3), "velocity": np.array(qvel, float), "attitude": np.zeros(3),
                           "angular_velocity": np.zeros(3)},
            "target": {"position": np.array(err, float), "velocity": np.array(tvel, float),
                       "acceleration": np.array(tacc, float)}, "time": 0.0}


# ---- target generators (env/target_motion.py) ------------------------------------------------------------
def test_target_patterns_follow_their_closed_forms():
    """Linear: constant velocity of the configured speed; circular: radius and angular rate speed / radius in
    the horizontal plane; sinusoidal: bounded by its amplitudes; stationary: the centre, at rest."""
    env, cfg = _env("linear", speed=1.5)
    o0 = env.reset(seed=5)
    p0, v0 = o0["target"]["position"].copy(), o0["target"]["velocity"].copy()
    assert abs(np.linalg.norm(v0) - 1.5) < 1e-12 and np.allclose(o0["target"]["acceleration"], 0.0)
    for _ in range(200):
        o, *_ = env.step(env.hover_action())
    np.testing.assert_allclose(o["target"]["position"], p0 + v0 * o["time"], rtol=0, atol=1e-12)

    env, cfg = _env("circular", radius=2.0, speed=1.0)
    o = env.reset(seed=5)
    c = np.array(cfg.target.center)
    for _ in range(300):
        r = o["target"]["position"] - c
        assert abs(np.hypot(r[0], r[1]) - 2.0) < 1e-12 and abs(r[2]) < 1e-12
        assert abs(np.linalg.norm(o["target"]["velocity"]) - 1.0) < 1e-12
        np.testing.assert_allclose(o["target"]["acceleration"][:2], -(0.5 ** 2) * r[:2], rtol=0, atol=1e-12)
        o, *_ = env.step(env.hover_action())

    env, cfg = _env("sinusoidal", amplitude=2.0, frequency=0.5)
    o = env.reset(seed=5)
    for _ in range(400):
        d = np.abs(o["target"]["position"] - c)
        assert (d <= np.array([2.0, 1.0, 0.5]) + 1e-12).all()
        o, *_ = env.step(env.hover_action())

    env, cfg = _env("stationary")
    o = env.reset(seed=5)
    o, *_ = env.step(env.hover_action())
    np.testing.assert_array_equal(o["target"]["position"], c)
    np.testing.assert_array_equal(o["target"]["velocity"], np.zeros(3))


def test_target_acceleration_is_capped_and_unknown_motion_is_rejected():
    from lqr_quadcopter_test_b200 import EnvConfig, QuadcopterEnv

    env, _ = _env("circular", radius=0.5, speed=3.0)       # centripetal 18 m/s^2 > max_acceleration 5
    o = env.reset(seed=1)
    for _ in range(50):
        assert np.linalg.norm(o["target"]["acceleration"]) <= 5.0 + 1e-12
        o, *_ = env.step(env.hover_action())
    cfg = EnvConfig()
    cfg.target.motion_type = "zigzag"
    with pytest.raises(ValueError, match="zigzag"):
        QuadcopterEnv(config=cfg).reset(seed=0)


def test_same_seed_same_episode_different_seed_different_episode():
    a, _ = _env("sinusoidal")
    b, _ = _env("sinusoidal")
    oa, ob = a.reset(seed=11), b.reset(seed=11)
    for _ in range(50):
        act = {"thrust": 10.5, "roll_rate": 0.2, "pitch_rate": -0.1, "yaw_rate": 0.05}
        oa, ra, *_ = a.step(act)
        ob, rb, *_ = b.step(act)
    np.testing.assert_array_equal(a.get_state_vector(), b.get_state_vector())
    assert ra == rb
    oc = b.reset(seed=12)
    assert not np.array_equal(oc["quadcopter"]["position"], a.reset(seed=11)["quadcopter"]["position"])


# ---- dynamics (env/quadcopter_env.py:295-465) ---------------------------------------------------------------
def test_hover_thrust_holds_altitude_and_free_fall_follows_gravity_with_drag():
    env, cfg = _env()
    env.reset(seed=42)
    s = np.zeros(12)
    s[2] = 5.0
    env.set_state_vector(s)
    for _ in range(100):
        o, *_ = env.step(env.hover_action(mass=cfg.quadcopter.mass, gravity=cfg.quadcopter.gravity))
    np.testing.assert_allclose(env.get_state_vector(), s, rtol=0, atol=1e-12)      # exact equilibrium
    # zero thrust: m dv/dt = -m g - c v  =>  v(t) = -(m g / c)(1 - exp(-c t / m))
    env.set_state_vector(s)
    for _ in range(100):
        o, *_ = env.step({"thrust": 0.0, "roll_rate": 0.0, "pitch_rate": 0.0, "yaw_rate": 0.0})
    m, g, c = cfg.quadcopter.mass, cfg.quadcopter.gravity, cfg.quadcopter.drag_coeff_linear
    t = 1.0
    assert abs(o["quadcopter"]["velocity"][2] + (m * g / c) * (1 - np.exp(-c * t / m))) < 1e-9   # RK4: O(dt^4)


def test_axis_sign_conventions():
    """+pitch rate tilts the thrust towards +X, +roll rate towards -Y (the reference's documented convention,
    tests/test_env_dynamics.py:1748-1830), and the body-rate loop is a first-order lag of 0.1 s."""
    env, _ = _env()
    for key, axis, sign in (("pitch_rate", 0, +1), ("roll_rate", 1, -1)):
        env.reset(seed=42)
        s = np.zeros(12)
        s[2] = 1.0
        env.set_state_vector(s)
        for _ in range(50):
            act = {"thrust": 9.81, "roll_rate": 0.0, "pitch_rate": 0.0, "yaw_rate": 0.0}
            act[key] = 0.5
            o, *_ = env.step(act)
        assert sign * o["quadcopter"]["velocity"][axis] > 0.05
        w = o["quadcopter"]["angular_velocity"][1 if key == "pitch_rate" else 0]
        pole = 1.0 / 0.1 + 0.01                                                    # 1 / tau + c_ang
        assert abs(w - 0.5 * (10.0 / pole) * (1 - np.exp(-pole * 0.5))) < 1e-5


def test_rk4_and_euler_agree_to_first_order_and_state_limits_hold():
    from lqr_quadcopter_test_b200 import EnvConfig, QuadcopterEnv

    finals = {}
    for integ in ("rk4", "euler"):
        cfg = EnvConfig()
        cfg.simulation.integrator = integ
        env = QuadcopterEnv(config=cfg)
        env.reset(seed=9)
        for k in range(200):
            o, *_ = env.step({"thrust": 11.0, "roll_rate": 0.3 * np.sin(0.05 * k), "pitch_rate": 0.2, "yaw_rate": 0.1})
        finals[integ] = env.get_state_vector()
    d = np.abs(finals["rk4"] - finals["euler"]).max()
    assert 1e-6 < d < 0.2
    # attitude clipped to +-60 degrees, body rates to +-10 rad/s, speed to 50 m/s, yaw wrapped to [-pi, pi)
    env = QuadcopterEnv()
    env.reset(seed=9)
    for _ in range(400):
        o, *_ = env.step({"thrust": 20.0, "roll_rate": 3.0, "pitch_rate": -3.0, "yaw_rate": 3.0})
        q = o["quadcopter"]
        assert np.abs(q["attitude"][:2]).max() <= np.pi / 3 + 1e-15 and -np.pi <= q["attitude"][2] < np.pi
        assert np.abs(q["angular_velocity"]).max() <= 10.0 and np.linalg.norm(q["velocity"]) <= 50.0 + 1e-9


def test_leaving_the_position_box_ends_the_episode():
    from lqr_quadcopter_test_b200 import EnvConfig, QuadcopterEnv

    cfg = EnvConfig()
    cfg.simulation.max_position = 3.0
    env = QuadcopterEnv(config=cfg)
    env.reset(seed=2)
    done, steps = False, 0
    while not done and steps < 2000:
        _, _, done, info = env.step({"thrust": 20.0, "roll_rate": 0.0, "pitch_rate": 0.0, "yaw_rate": 0.0})
        steps += 1
    assert done and info["termination_reason"] == "position_bounds" and steps < 200


# ---- Riccati-LQR law (controllers/riccati_lqr.py:779-967) ------------------------------------------------------
def test_riccati_controller_responds_along_the_right_axes():
    from lqr_quadcopter_test_b200 import RiccatiLQRController

    c = RiccatiLQRController(config={})
    hover = c.compute_action(_obs())
    assert abs(hover["thrust"] - 9.81) < 1e-12 and hover["roll_rate"] == 0.0 and hover["pitch_rate"] == 0.0
    assert c.compute_action(_obs(err=(1, 0, 0)))["pitch_rate"] > 0          # target ahead in X: pitch forward
    assert c.compute_action(_obs(err=(0, 1, 0)))["roll_rate"] < 0           # target to +Y: negative roll rate
    assert c.compute_action(_obs(err=(0, 0, 1)))["thrust"] > 9.81           # target above: more thrust
    assert c.compute_action(_obs(qvel=(0, 0, 1)))["thrust"] < 9.81          # climbing past it: less thrust
    for e in ((1e3, -1e3, 1e3), (-1e3, 1e3, -1e3)):                           # outputs stay inside the actuator limits
        a = c.compute_action(_obs(err=e))
        assert 0.0 <= a["thrust"] <= 20.0 and abs(a["roll_rate"]) <= 3.0 and abs(a["pitch_rate"]) <= 3.0
    assert c.compute_action(_obs(err=(1, 0, 0)))["yaw_rate"] == 0.0
    # P solves the Riccati equation, is symmetric positive definite, and A - B K is stable
    A, B, Q, R, Pm, K = c.A, c.B, c.Q, c.R, c.P, c.K
    res = A.T @ Pm @ A - Pm - A.T @ Pm @ B @ np.linalg.solve(R + B.T @ Pm @ B, B.T @ Pm @ A) + Q
    assert np.abs(res).max() < 1e-9 * np.abs(Pm).max() and np.allclose(Pm, Pm.T) and np.linalg.eigvalsh(Pm).min() > 0
    assert np.abs(np.linalg.eigvals(A - B @ K)).max() < 1.0
    # custom mass / gravity move the hover thrust
    c2 = RiccatiLQRController(config={"mass": 2.0, "gravity": 3.7})
    assert abs(c2.compute_action(_obs())["thrust"] - 7.4) < 1e-12


def test_riccati_controller_holds_a_stationary_target_and_stays_bounded_on_moving_ones():
    """What the reference asserts for its default gains (tests/test_env_dynamics.py:3085-3156): success on a
    stationary target; on moving targets only that the loop does not diverge — the default design trails a
    1 m/s target by metres (the fixtures show on-target ratios of 0.1-0.4 on linear and < 0.03 on circular
    targets for the reference itself)."""
    from lqr_quadcopter_test_b200 import Evaluator, RiccatiLQRController

    _, cfg = _env("stationary")
    s = Evaluator(RiccatiLQRController(config={}), cfg).evaluate(num_episodes=8, base_seed=100, verbose=False)
    assert s.mean_on_target_ratio >= 0.9 and s.success_rate == 1.0 and s.meets_criteria
    for motion, bound in (("linear", 5.0), ("circular", 10.0), ("sinusoidal", 15.0)):
        _, cfg = _env(motion)
        s = Evaluator(RiccatiLQRController(config={}), cfg).evaluate(num_episodes=8, base_seed=100, verbose=False)
        assert s.total_episodes == 8 and np.isfinite(s.mean_tracking_error) and s.mean_tracking_error < bound, motion
        assert all(m.episode_duration >= 30.0 for m in s.episode_metrics)          # nobody left the box


def test_feedforward_is_off_by_default_and_helps_on_a_circle():
    from lqr_quadcopter_test_b200 import Evaluator, RiccatiLQRController

    plain = RiccatiLQRController(config={})
    assert plain.feedforward_enabled is False
    o = _obs(err=(0.2, 0.0, 0.0), tvel=(1.0, 0.0, 0.0), tacc=(0.0, 2.0, 0.0))
    ff = RiccatiLQRController(config={"feedforward_enabled": True, "ff_acceleration_gain": [0.1, 0.1, 0.0]})
    a0, a1 = plain.compute_action(o), ff.compute_action(o)
    assert a1["roll_rate"] < a0["roll_rate"] and a1["pitch_rate"] == a0["pitch_rate"]    # +Y acceleration: more negative roll
    explicit_off = RiccatiLQRController(config={"feedforward_enabled": False, "ff_acceleration_gain": [0.1, 0.1, 0.0]})
    assert explicit_off.compute_action(o) == a0
    # the reference's scenario (tests/test_env_dynamics.py:3416-3472): circle of radius 2 at 1 m/s, seed 42, first 5 s
    _, cfg = _env("circular", seconds=10.0, radius=2.0, speed=1.0)
    err = {}
    for name, ctrl in (("plain", plain), ("ff", ff)):
        s = Evaluator(ctrl, cfg).evaluate(num_episodes=1, base_seed=42, max_steps_per_episode=500, verbose=False)
        err[name] = s.mean_tracking_error
    assert err["ff"] < err["plain"], err


def test_lqi_integrator_accumulates_clamps_and_resets():
    from lqr_quadcopter_test_b200 import RiccatiLQRController

    c = RiccatiLQRController(config={"use_lqi": True, "q_int": [0.01, 0.01, 0.1], "integral_limit": 0.05})
    assert c.is_lqi_mode() and c.K.shape == (4, 9) and c.get_integral_gains().shape == (4, 3)
    o = _obs(err=(1.0, -2.0, 0.5))
    c.compute_action(o)
    np.testing.assert_allclose(c.get_integral_state(), 0.01 * np.array([1.0, -2.0, 0.5]), rtol=0, atol=1e-15)
    for _ in range(20):
        c.compute_action(o)
    np.testing.assert_array_equal(np.abs(c.get_integral_state()) <= 0.05, [True, True, True])
    assert abs(c.get_integral_state()[1]) == 0.05                              # clamped at the limit
    before = c.get_integral_state().copy()
    c.compute_action(_obs(err=(0.001, 0.0, 0.0)))                               # inside the 0.01 m dead zone: frozen
    np.testing.assert_array_equal(c.get_integral_state(), before)
    c.compute_action(_obs(err=(0.0, 2.0, 0.0)))                                 # opposite sign: the saturated axis unwinds
    assert abs(c.get_integral_state()[1]) < 0.05
    c.reset()
    np.testing.assert_array_equal(c.get_integral_state(), np.zeros(3))
    # the integral action removes the steady-state offset a constant disturbance leaves under plain LQR
    a_lqr = RiccatiLQRController(config={}).compute_action(o)["thrust"]
    c2 = RiccatiLQRController(config={"use_lqi": True, "q_int": [0.01, 0.01, 0.1]})
    first = c2.compute_action(o)["thrust"]
    for _ in range(50):
        later = c2.compute_action(o)["thrust"]
    assert later > first and np.isfinite(a_lqr)


def test_invalid_cost_matrices_and_observations_raise_like_the_reference():
    from lqr_quadcopter_test_b200 import RiccatiLQRController

    with pytest.raises(ValueError):
        RiccatiLQRController(config={"Q": -np.eye(6), "fallback_on_failure": False})
    with pytest.raises(ValueError):
        RiccatiLQRController(config={"R": np.zeros((4, 4)), "fallback_on_failure": False})
    with pytest.raises(ValueError):
        RiccatiLQRController(config={"Q": np.eye(5), "fallback_on_failure": False})
    Qa = np.diag([1e-4, 1e-4, 16.0, 3.6e-3, 3.6e-3, 4.0])
    Qa[0, 1] = 1e-3                                                              # not symmetric: rejected
    with pytest.raises(ValueError):
        RiccatiLQRController(config={"Q": Qa, "fallback_on_failure": False})
    fb = RiccatiLQRController(config={"Q": Qa})                                  # default: heuristic LQR from diag(Q)
    a = fb.compute_action(_obs(err=(1, 1, 1)))
    assert fb.is_using_fallback() and np.isfinite(a["thrust"]) and a["pitch_rate"] > 0 and a["roll_rate"] < 0
    c = RiccatiLQRController(config={})
    with pytest.raises(KeyError):
        c.compute_action({"quadcopter": {"position": np.zeros(3)}})
