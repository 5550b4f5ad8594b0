"""Shared description of the rollout fixtures in tests/golden/rollouts.npz (made by
oracle/gen_golden.py): how each case configured the reference's env and controller."""
from __future__ import annotations

from types import SimpleNamespace

from lqr_quadcopter_test_b200 import params as P
from lqr_quadcopter_test_b200.config import EnvConfig
from lqr_quadcopter_test_b200.metrics import SuccessCriteria

# trajectory tolerance (absolute, metres / m/s / rad) over the horizon, see DESIGN.md "Numerics":
# closed loops that contract keep fp64 re-association noise at 1e-13; the reference's default LQI
# gains are unstable and saturated and amplify it ~100x.
TOL_LQR = 1e-10
TOL_LQI = 1e-8


def controller_ns(**kw) -> SimpleNamespace:
    base = dict(dt=0.01, hover_thrust=9.81, max_thrust=20.0, min_thrust=0.0, max_rate=3.0, use_lqi=False,
                integral_limit=10.0, integral_zero_threshold=0.01, feedforward_enabled=False,
                ff_velocity_gain=[0.0, 0.0, 0.0], ff_acceleration_gain=[0.0, 0.0, 0.0],
                ff_max_velocity=10.0, ff_max_acceleration=5.0)
    base.update(kw)
    return SimpleNamespace(**base)


def case(name: str):
    """-> (qt_params, max_steps, chain, tolerance, controller config dict, env config, criteria)."""
    motion = name.split("_")[-1]
    env = EnvConfig()
    env.target.motion_type = motion
    ctrl = controller_ns()
    ctrl_cfg: dict = {}
    crit = SuccessCriteria()
    max_steps, chain, tol = 3002, 1, TOL_LQR
    if name.startswith("lqi"):
        ctrl.use_lqi = True
        ctrl_cfg = {"use_lqi": True, "q_int": [0.01, 0.01, 0.1]}
        tol = TOL_LQI
    if name == "lqi_carry_stationary":
        chain = 3
    if name == "lqi_small_linear":
        ctrl.integral_limit = 0.5
        ctrl_cfg = {"use_lqi": True, "q_int": [1e-6, 1e-6, 1e-3], "integral_limit": 0.5}
    if name == "lqr_ff_circular":
        ctrl_cfg = {"feedforward_enabled": True, "ff_velocity_gain": [0.1, 0.1, 0.0],
                    "ff_acceleration_gain": [0.05, 0.05, 0.0]}
        ctrl.feedforward_enabled = True
        ctrl.ff_velocity_gain, ctrl.ff_acceleration_gain = [0.1, 0.1, 0.0], [0.05, 0.05, 0.0]
    if name == "lqr_ff_sinusoidal":
        ctrl_cfg = {"feedforward_enabled": True, "ff_velocity_gain": 0.2, "ff_acceleration_gain": 0.1,
                    "ff_max_velocity": 1.5, "ff_max_acceleration": 2.0}
        ctrl.feedforward_enabled = True
        ctrl.ff_velocity_gain, ctrl.ff_acceleration_gain = [0.2] * 3, [0.1] * 3
        ctrl.ff_max_velocity, ctrl.ff_max_acceleration = 1.5, 2.0
        max_steps = 1500
    if name == "lqr_euler_linear":
        env.simulation.integrator = "euler"
        max_steps = 1000
    if name == "lqr_custom_circular":
        env.quadcopter.mass, env.quadcopter.drag_coeff_linear = 1.3, 0.2
        env.simulation.dt, env.simulation.max_episode_time = 0.02, 10.0
        env.target.speed, env.target.radius, env.target.center = 2.0, 3.0, (1.0, -1.0, 2.0)
        env.success_criteria.target_radius = 0.3
        ctrl.dt, ctrl.hover_thrust = 0.02, 1.3 * 9.81
        ctrl_cfg = {"mass": 1.3, "dt": 0.02, "q_pos": [2e-4, 2e-4, 12.0], "r_controls": [0.8, 1.2, 1.2, 1.0]}
        crit = SuccessCriteria(min_on_target_ratio=0.5, min_episode_duration=10.0, target_radius=0.3)
    if name == "lqr_bounds_linear":
        env.simulation.max_position = 4.0
    if name == "lqr_cap_stationary":
        max_steps = 100
    if name == "lqr_short_stationary":
        max_steps = 5
    if name == "lqr_figure8":
        tol = 1e-9
    return P.make_params(env, ctrl, crit), max_steps, chain, tol, ctrl_cfg, env, crit


def case_names(npz) -> list[str]:
    return sorted({k[: -len("_seeds")] for k in npz.files if k.endswith("_seeds")})


COUNT_ROWS = [P.M_STEPS, P.M_ENV_ON_TARGET, P.M_MET_ON_TARGET, P.M_OVERSHOOT_COUNT, P.M_ENV_SUCCESS,
              P.M_MET_SUCCESS, P.M_TERMINATION, P.M_VIOLATIONS, P.M_SATURATIONS]
SUM_ROWS = [P.M_SUM_ERR, P.M_SUM_ERR2, P.M_SUM_EFFORT]
VALUE_ROWS = [P.M_FINAL_TIME, P.M_MAX_ERR, P.M_MAX_OVERSHOOT, P.M_FINAL_ERR]


def assert_rollout_matches(np, got: dict, npz, name: str, tol: float, sum_rtol: float = 1e-11):
    """Compare a rollout result (metrics [16,E], final_state [12,E]) with the reference fixture."""
    M, G = got["metrics"], npz[name + "_metrics"]
    assert M.shape == G.shape
    # the reference's heuristic LQRController keeps no saturation counter (controllers/__init__.py:398-700)
    count_rows = [r for r in COUNT_ROWS if not (name.startswith("hlqr") and r == P.M_SATURATIONS)]
    np.testing.assert_array_equal(M[count_rows], G[count_rows], err_msg=f"{name}: counters / flags")
    np.testing.assert_allclose(M[SUM_ROWS], G[SUM_ROWS], rtol=max(sum_rtol, tol), atol=tol * 3000,
                               err_msg=f"{name}: per-episode sums")
    np.testing.assert_allclose(M[VALUE_ROWS], G[VALUE_ROWS], rtol=0, atol=tol, err_msg=f"{name}: values")
    np.testing.assert_allclose(got["final_state"], npz[name + "_final_state"], rtol=0, atol=tol,
                               err_msg=f"{name}: final state")


# ---- §8(f).1: PID and heuristic-LQR fixtures (tests/golden/pid_lqr.npz) -----------------------------
PID_KI = dict(integral_limit=0.4)
PID_FF = dict(feedforward_enabled=True, ff_velocity_gain=[0.1, 0.1, 0.0], ff_acceleration_gain=[0.05, 0.05, 0.0],
              ff_max_velocity=0.9, ff_max_acceleration=0.4)
HLQR_FF = dict(feedforward_enabled=True, ff_velocity_gain=[0.15] * 3, ff_acceleration_gain=[0.05, 0.1, 0.0],
               ff_max_velocity=0.8)
PID_LQR_CASES = {
    # name: (controller kwargs for controller_ns, max_steps, chain, constructor config of the drop-in class)
    "pid_stationary": (dict(name="pid", integral_limit=0.0), 3002, 1, {}),
    "pid_linear": (dict(name="pid", integral_limit=0.0), 3002, 1, {}),
    "pid_circular": (dict(name="pid", integral_limit=0.0), 3002, 1, {}),
    "pid_sinusoidal": (dict(name="pid", integral_limit=0.0), 1500, 1, {}),
    "pidki_linear": (dict(name="pid", **PID_KI), 3002, 1, {"ki_pos": [0.002, 0.002, 0.5], "integral_limit": 0.4}),
    "pidki_carry_stationary": (dict(name="pid", **PID_KI), 1200, 3,
                               {"ki_pos": [0.002, 0.002, 0.5], "integral_limit": 0.4}),
    "pidff_circular": (dict(name="pid", integral_limit=0.0, **PID_FF), 3002, 1,
                       {"feedforward_enabled": True, "ff_velocity_gain": [0.1, 0.1, 0.0],
                        "ff_acceleration_gain": [0.05, 0.05, 0.0], "ff_max_velocity": 0.9, "ff_max_acceleration": 0.4}),
    "hlqr_linear": (dict(), 3002, 1, {}),
    "hlqr_circular": (dict(), 3002, 1, {"q_pos": [2e-4, 3e-4, 12.0], "r_rate": 1.5, "r_thrust": 0.7}),
    "hlqrff_sinusoidal": (dict(**HLQR_FF), 1500, 1,
                          {"feedforward_enabled": True, "ff_velocity_gain": 0.15,
                           "ff_acceleration_gain": [0.05, 0.1, 0.0], "ff_max_velocity": 0.8}),
}


def pid_lqr_case(name: str):
    """-> (qt_params, max_steps, chain, tolerance, constructor config, env config)."""
    kw, max_steps, chain, cfg = PID_LQR_CASES[name]
    env = EnvConfig()
    env.target.motion_type = name.split("_")[-1]
    return P.make_params(env, controller_ns(**kw), SuccessCriteria()), max_steps, chain, TOL_LQR, cfg, env


# ---- §8(f).2: deep tracking policy fixtures (tests/golden/policy.npz) --------------------------------
# fp32 network: torch (the reference), the C oracle and the CUDA kernel accumulate the dot products in
# different orders (and the kernel contracts into FMA), so single actions agree to a few fp32 ulps of
# the 20 N thrust range and 30 s closed loops to ~1e-5; stated tolerances carry a 10x margin.
TOL_FIGURE8_FF = 5e-4        # acceleration feedforward on the figure-8 target: see tests/test_soak_gpu.py
TOL_POLICY_ACTION = 5e-5
TOL_POLICY_STATE = 1e-4
POLICIES = {
    "relu_64_64": ([64, 64], "relu", None),
    "tanh_32": ([32], "tanh", None),
    "elu_48_40": ([48, 40], "elu", None),
    "leaky_20_20_20": ([20, 20, 20], "leaky_relu",
                       {"thrust": (2.0, 18.0), "roll_rate": (-2.0, 2.0), "pitch_rate": (-2.5, 2.5), "yaw_rate": (-1.0, 1.0)}),
}


def policy_arrays(npz, name):
    n = len(POLICIES[name][0]) + 1
    return [npz[f"{name}_W{l}"] for l in range(n)], [npz[f"{name}_b{l}"] for l in range(n)]


def assert_policy_rollout_matches(np, got, npz, key):
    M, G = got["metrics"], npz[key + "_metrics"]
    hard = [P.M_STEPS, P.M_TERMINATION, P.M_VIOLATIONS]
    np.testing.assert_array_equal(M[hard], G[hard], err_msg=key)
    soft = [P.M_ENV_ON_TARGET, P.M_MET_ON_TARGET]
    assert np.abs(M[soft] - G[soft]).max() <= 2, f"{key}: on-target counts beyond borderline flips"
    np.testing.assert_allclose(M[SUM_ROWS], G[SUM_ROWS], rtol=2e-5, atol=1e-3, err_msg=key)
    np.testing.assert_allclose(M[[P.M_FINAL_TIME, P.M_MAX_ERR, P.M_FINAL_ERR]], G[[P.M_FINAL_TIME, P.M_MAX_ERR, P.M_FINAL_ERR]],
                               rtol=0, atol=TOL_POLICY_STATE, err_msg=key)
    np.testing.assert_allclose(got["final_state"], npz[key + "_final_state"], rtol=0, atol=TOL_POLICY_STATE, err_msg=key)
    samples = npz[key + "_samples"]
    traj = np.transpose(got["traj"][: samples.shape[1], P.TRAJ_STATE:P.TRAJ_STATE + 12, :], (2, 0, 1))
    np.testing.assert_allclose(traj, samples, rtol=0, atol=TOL_POLICY_STATE, err_msg=key)
