"""CPU ORACLE — test infrastructure, not product code.

Restates the reference's algorithm for the accelerated path so the CUDA kernels can be checked
(and the reference's CPU cost timed) on a machine where the Python reference is absent:

* ``qt_oracle.c``   plain-C closed loop (env step, target motion, Riccati-LQR/LQI law, metrics)
* ``reset_reference``  numpy restatement of QuadcopterEnv.reset + TargetMotion.reset — it calls
  ``numpy.random.default_rng`` itself, the third-party generator the reference uses
* ``solve_dare_reference``  the reference's solve_dare: validation + scipy.linalg.solve_discrete_are
  (scipy is the third-party solver the reference calls; pinned 1.16.3 in its lock file, 1.18.1 here)

Parity pinning: tests/test_oracle_golden.py compares all three against fixtures generated from the
real reference by oracle/gen_golden.py (tests/golden/).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  Nothing in lqr_quadcopter_test_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import math
import pathlib
import subprocess

import numpy as np

from lqr_quadcopter_test_b200.params import QtParams  # the ABI struct is shared boundary data

_DIR = pathlib.Path(__file__).resolve().parent
_LIB_PATH = _DIR / "libqt_oracle.so"
_lib = None


def build(force: bool = False) -> pathlib.Path:
    src = _DIR / "qt_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_DIR), "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(str(_LIB_PATH))
        if l.qto_params_size() != C.sizeof(QtParams):
            raise RuntimeError("oracle/qt_params layout mismatch")
        l.qto_np_sum.restype = C.c_double
        _lib = l
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def k_cols_of(params: QtParams) -> int:
    return 3 if params.controller_type == 1 else (9 if params.use_lqi else 6)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def np_sum(a: np.ndarray) -> float:
    a = _f64(a)
    return float(lib().qto_np_sum(_p(a), C.c_int64(a.size)))


def rollout(params: QtParams, state0, tparams, K, k_stride=None, max_steps=3000, chain=1,
            ctrl_state=None, want_final_state=True, traj_stride=0, data_stride=0, n_threads=1, ff_gains=None):
    """Same contract as qt_rollout (include/qtrack.h).  Arrays are [F, E] float64."""
    state0, tparams, K = _f64(state0), _f64(tparams), _f64(K)
    E = state0.shape[1]
    k_cols = k_cols_of(params)
    K = K.reshape(-1, 4, k_cols)
    if k_stride is None:
        k_stride = E if K.shape[0] == 1 else 1
    out = {"metrics": np.zeros((16, E))}
    if want_final_state:
        out["final_state"] = np.zeros((12, E))
    traj_slots = data_slots = 0
    if traj_stride > 0:
        traj_slots = (max_steps + traj_stride - 1) // traj_stride
        out["traj"] = np.zeros((traj_slots, 24, E))
    if data_stride > 0:
        data_slots = (max_steps + data_stride - 1) // data_stride
        out["features"] = np.zeros((data_slots, E, 18), dtype=np.float32)
        out["actions"] = np.zeros((data_slots, E, 4), dtype=np.float32)
    if ctrl_state is not None:
        assert ctrl_state.dtype == np.float64 and ctrl_state.flags.c_contiguous
    lib().qto_rollout(C.byref(params), C.c_int64(E), C.c_int64(max_steps), C.c_int64(chain), _p(state0),
                      _p(tparams), _p(K), C.c_int32(k_cols), C.c_int64(k_stride), _p(out["metrics"]),
                      _p(ctrl_state), _p(out.get("final_state")), _p(out.get("traj")),
                      C.c_int64(max(traj_stride, 1)), C.c_int64(traj_slots), _p(out.get("features")),
                      _p(out.get("actions")), C.c_int64(max(data_stride, 1)), C.c_int64(data_slots),
                      C.c_int32(n_threads), _p(None if ff_gains is None else _f64(ff_gains)))
    return out


def observe(params: QtParams, state, time, tparams):
    state, time, tparams = _f64(state), _f64(time), _f64(tparams)
    E = state.shape[1]
    obs = np.zeros((22, E))
    lib().qto_observe_batched(C.byref(params), C.c_int64(E), _p(state), _p(time), _p(tparams), _p(obs))
    return obs


def compute_action(params: QtParams, obs, K, k_stride=None, integral=None, saturation=None):
    obs, K = _f64(obs), _f64(K)
    E = obs.shape[1]
    k_cols = k_cols_of(params)
    if k_stride is None:
        k_stride = E
    act = np.zeros((4, E))
    lib().qto_compute_action_batched(C.byref(params), C.c_int64(E), _p(obs), _p(K), C.c_int32(k_cols),
                                     C.c_int64(k_stride), _p(integral), _p(saturation), _p(act))
    return act


def step(params: QtParams, state, time, tparams, counters, action=None, K=None, k_stride=None,
         integral=None, saturation=None):
    """In-place env step on [12,E] state / [E] time / int64 [3,E] counters; returns outputs."""
    E = state.shape[1]
    k_cols = k_cols_of(params)
    out = {"obs": np.zeros((22, E)), "reward": np.zeros(E), "done": np.zeros(E, dtype=np.int32),
           "info": np.zeros((4, E)), "action": np.zeros((4, E))}
    if action is not None:
        action = _f64(action)
    if K is not None:
        K = _f64(K)
    lib().qto_step_batched(C.byref(params), C.c_int64(E), _p(state), _p(time), _p(_f64(tparams)), _p(counters),
                           _p(action), _p(K), C.c_int32(k_cols), C.c_int64(k_stride or E), _p(integral),
                           _p(saturation), _p(out["obs"]), _p(out["reward"]), _p(out["done"]), _p(out["info"]),
                           _p(out["action"]))
    return out


def reduce_scores(metrics, C_, eps):
    metrics = _f64(metrics)
    out = np.zeros((4, C_))
    lib().qto_reduce_scores(_p(metrics), C.c_int64(C_), C.c_int64(eps), _p(out))
    return out


# ------------------------------------------------------------------------------------- reset --
def reset_reference(params: QtParams, seeds, draws=None):
    """QuadcopterEnv.reset(seed) + TargetMotion.reset/_create_pattern
    (env/quadcopter_env.py:111-150, env/target_motion.py:291-369) with numpy's own generator.
    Returns (state0 [12,E], tparams [12,E])."""
    seeds = np.asarray(seeds).reshape(-1)
    E = seeds.size
    state = np.zeros((12, E))
    tp = np.zeros((12, E))
    center = np.array([params.target_center[i] for i in range(3)])
    for e, seed in enumerate(seeds):
        rng_env = np.random.default_rng(int(seed))
        rng_tgt = np.random.default_rng(int(seed))
        tp[0:3, e] = center
        p0 = center.copy()
        mt = params.motion_type
        if mt == 1:  # linear
            d = rng_tgt.standard_normal(3)
            d /= np.linalg.norm(d)
            d = d / np.linalg.norm(d)
            vel = d * params.target_speed
            tp[3:6, e] = vel
            p0 = center + vel * 0.0
        elif mt == 2:  # circular
            a0 = rng_tgt.uniform(0, 2 * math.pi)
            omega = params.target_speed / params.target_radius
            tp[3, e], tp[4, e], tp[5, e] = params.target_radius, omega, a0
            ang = a0 + omega * 0.0
            p0 = np.array([center[0] + params.target_radius * np.cos(ang),
                           center[1] + params.target_radius * np.sin(ang), center[2]])
        elif mt == 3:  # sinusoidal
            amp = np.array([params.target_amplitude, params.target_amplitude * 0.5, params.target_amplitude * 0.25])
            frq = np.array([params.target_frequency, params.target_frequency * 1.3, params.target_frequency * 0.7])
            phase = rng_tgt.uniform(0, 2 * math.pi, 3)
            omega = 2 * np.pi * frq
            tp[3:6, e], tp[6:9, e], tp[9:12, e] = amp, omega, phase
            p0 = center + amp * np.sin(omega * 0.0 + phase)
        elif mt == 4:  # figure8
            scale = params.target_amplitude
            omega = params.target_speed / scale
            tp[3, e], tp[4, e] = scale, omega
            th = omega * 0.0
            c, s = np.cos(th), np.sin(th)
            den = 1 + s ** 2
            p0 = np.array([center[0] + scale * c / den, center[1] + scale * s * c / den, center[2]])
        off = rng_env.uniform(-0.5, 0.5, 3)
        state[0:3, e] = p0 + off
    return state, tp


# -------------------------------------------------------------------------------------- DARE --
def build_systems_reference(dt=0.01, mass=1.0, gravity=9.81, lqi=False):
    """riccati_lqr.py:187-316"""
    A_c = np.zeros((6, 6))
    A_c[0, 3] = A_c[1, 4] = A_c[2, 5] = 1.0
    B_c = np.zeros((6, 4))
    B_c[5, 0] = 1.0 / mass
    B_c[4, 1] = -gravity
    B_c[3, 2] = gravity
    A = np.eye(6) + A_c * dt
    B = B_c * dt
    if not lqi:
        return A, B
    A_aug = np.zeros((9, 9))
    A_aug[:6, :6] = A
    A_aug[6:, :3] = np.eye(3) * dt
    A_aug[6:, 6:] = np.eye(3)
    B_aug = np.zeros((9, 4))
    B_aug[:6, :] = B
    return A_aug, B_aug


def validate_reference(Q, R) -> int:
    """riccati_lqr.py:57-116, 166-172: 0 ok, 3 rejected."""
    def sym(M):
        return np.allclose(M, M.T, atol=1e-8)
    if not sym(Q) or np.any(np.linalg.eigvalsh(Q) < -1e-10):
        return 3
    if not sym(R) or np.any(np.linalg.eigvalsh(R) <= 1e-10):
        return 3
    return 0


def solve_dare_reference(A, B, Q, R):
    """riccati_lqr.py:174-184: P from scipy's ordered-QZ solver, K = (R + B'PB)^-1 B'PA."""
    from scipy.linalg import solve_discrete_are

    P = solve_discrete_are(A, B, Q, R)
    BtP = B.T @ P
    K = np.linalg.solve(R + BtP @ B, BtP @ A)
    return P, K


# ---- §8(f).2: deep tracking policy ------------------------------------------------------------------
class _QtoMlp(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dims", C.c_int32 * 7), ("activation", C.c_int32), ("reserved", C.c_int32),
                ("out_lo", C.c_float * 4), ("out_scale", C.c_float * 4), ("W", C.c_void_p * 6), ("b", C.c_void_p * 6)]


class Mlp:
    """PolicyNetwork parameters in torch's layout: weights[l] [out, in], biases[l] [out], float32."""

    ACT = {"relu": 0, "tanh": 1, "elu": 2, "leaky_relu": 3}

    def __init__(self, weights, biases, activation="relu", bounds=((0.0, 20.0), (-3.0, 3.0), (-3.0, 3.0), (-3.0, 3.0))):
        self.weights = [np.ascontiguousarray(w, dtype=np.float32) for w in weights]
        self.biases = [np.ascontiguousarray(b, dtype=np.float32) for b in biases]
        m = _QtoMlp()
        m.n_layers = len(self.weights)
        m.dims[0] = self.weights[0].shape[1]
        for l, w in enumerate(self.weights):
            m.dims[l + 1] = w.shape[0]
            m.W[l] = w.ctypes.data
            m.b[l] = self.biases[l].ctypes.data
        m.activation = self.ACT[activation]
        for i, (lo, hi) in enumerate(bounds):
            m.out_lo[i] = lo
            m.out_scale[i] = hi - lo
        self.c = m


def policy_action(mlp: Mlp, obs):
    obs = _f64(obs)
    E = obs.shape[1]
    act = np.zeros((4, E))
    lib().qto_policy_action_batched(C.byref(mlp.c), C.c_int64(E), _p(obs), _p(act))
    return act


def rollout_policy(params: QtParams, state0, tparams, mlp: Mlp, K=None, k_stride=None, max_steps=3000,
                   want_final_state=True, traj_stride=0, data_stride=0, n_threads=1):
    """Same contract as qt_rollout_policy (include/qtrack.h)."""
    state0, tparams = _f64(state0), _f64(tparams)
    E = state0.shape[1]
    k_cols = k_cols_of(params)
    if K is not None:
        K = _f64(K).reshape(-1, 4, k_cols)
        if k_stride is None:
            k_stride = E if K.shape[0] == 1 else 1
    out = {"metrics": np.zeros((16, E))}
    if want_final_state:
        out["final_state"] = np.zeros((12, E))
    traj_slots = data_slots = 0
    if traj_stride > 0:
        traj_slots = (max_steps + traj_stride - 1) // traj_stride
        out["traj"] = np.zeros((traj_slots, 24, E))
    if data_stride > 0:
        data_slots = (max_steps + data_stride - 1) // data_stride
        out["features"] = np.zeros((data_slots, E, 18), dtype=np.float32)
        out["actions"] = np.zeros((data_slots, E, 4), dtype=np.float32)
    lib().qto_rollout_policy(C.byref(params), C.c_int64(E), C.c_int64(max_steps), _p(state0), _p(tparams),
                             C.byref(mlp.c), _p(K), C.c_int32(k_cols), C.c_int64(k_stride or E), _p(out["metrics"]),
                             _p(out.get("final_state")), _p(out.get("traj")), C.c_int64(max(traj_stride, 1)),
                             C.c_int64(traj_slots), _p(out.get("features")), _p(out.get("actions")),
                             C.c_int64(max(data_stride, 1)), C.c_int64(data_slots), C.c_int32(n_threads))
    return out
