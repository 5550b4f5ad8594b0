"""Thin numpy-in / numpy-out driver of the C ABI (through the package's ctypes Context), with the
same call shapes as the oracle package so parity tests read symmetrically."""
from __future__ import annotations

import numpy as np
import torch

from lqr_quadcopter_test_b200 import _lib


def ctx():
    return _lib.get_context()


def k_cols_of(params) -> int:
    return 3 if params.controller_type == 1 else (9 if params.use_lqi else 6)


def dev(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=ctx().device).contiguous()


def reset(params, seeds, draws=None):
    c = ctx()
    s = torch.as_tensor(np.asarray(seeds).astype(np.int64), device=c.device)
    E = s.numel()
    state = torch.empty((12, E), dtype=torch.float64, device=c.device)
    tp = torch.empty((12, E), dtype=torch.float64, device=c.device)
    c.reset_batched(params, s, state, tp, draws)
    return state.cpu().numpy(), tp.cpu().numpy()


def rollout(params, state0, tparams, K, k_stride=None, max_steps=3000, chain=1, ctrl_state=None,
            want_final_state=True, traj_stride=0, data_stride=0, ff_gains=None):
    c = ctx()
    E = state0.shape[1]
    k_cols = k_cols_of(params)
    Kt = dev(np.asarray(K).reshape(-1, 4, k_cols))
    if k_stride is None:
        k_stride = E if Kt.shape[0] == 1 else 1
    f64 = dict(dtype=torch.float64, device=c.device)
    out = {"metrics": torch.zeros((16, E), **f64)}
    if want_final_state:
        out["final_state"] = torch.zeros((12, E), **f64)
    ts = dsl = 0
    if traj_stride > 0:
        ts = (max_steps + traj_stride - 1) // traj_stride
        out["traj"] = torch.zeros((ts, 24, E), **f64)
    if data_stride > 0:
        dsl = (max_steps + data_stride - 1) // data_stride
        out["features"] = torch.zeros((dsl, E, 18), dtype=torch.float32, device=c.device)
        out["actions"] = torch.zeros((dsl, E, 4), dtype=torch.float32, device=c.device)
    cs = dev(ctrl_state) if ctrl_state is not None else None
    c.rollout(params, E, max_steps, chain, dev(state0), dev(tparams), Kt, k_cols, k_stride, out["metrics"], cs,
              out.get("final_state"), out.get("traj"), max(traj_stride, 1), ts, out.get("features"),
              out.get("actions"), max(data_stride, 1), dsl, dev(ff_gains) if ff_gains is not None else None)
    res = {k: v.cpu().numpy() for k, v in out.items()}
    if cs is not None:
        ctrl_state[...] = cs.cpu().numpy()
    return res


def observe(params, state, time, tparams):
    c = ctx()
    E = state.shape[1]
    obs = torch.zeros((22, E), dtype=torch.float64, device=c.device)
    c.observe_batched(params, E, dev(state), dev(time), dev(tparams), obs)
    return obs.cpu().numpy()


def compute_action(params, obs, K, k_stride=None, integral=None, saturation=None):
    c = ctx()
    E = obs.shape[1]
    k_cols = k_cols_of(params)
    act = torch.zeros((4, E), dtype=torch.float64, device=c.device)
    it = dev(integral) if integral is not None else None
    st = dev(saturation, torch.int64) if saturation is not None else None
    c.compute_action_batched(params, E, dev(obs), dev(K), k_cols, k_stride or E, it, st, act)
    if integral is not None:
        integral[...] = it.cpu().numpy()
    if saturation is not None:
        saturation[...] = st.cpu().numpy()
    return act.cpu().numpy()


def step(params, state, time, tparams, counters, action=None, K=None, k_stride=None, integral=None,
         saturation=None):
    c = ctx()
    E = state.shape[1]
    k_cols = k_cols_of(params)
    f64 = dict(dtype=torch.float64, device=c.device)
    st, tm, cn = dev(state), dev(time), dev(counters, torch.int64)
    out = {"obs": torch.zeros((22, E), **f64), "reward": torch.zeros(E, **f64),
           "done": torch.zeros(E, dtype=torch.int32, device=c.device), "info": torch.zeros((4, E), **f64),
           "action": torch.zeros((4, E), **f64)}
    it = dev(integral) if integral is not None else None
    sa = dev(saturation, torch.int64) if saturation is not None else None
    c.step_batched(params, E, st, tm, dev(tparams), cn, dev(action) if action is not None else None,
                   dev(K) if K is not None else None, k_cols, k_stride or E, it, sa, out["obs"], out["reward"],
                   out["done"], out["info"], out["action"])
    state[...] = st.cpu().numpy()
    time[...] = tm.cpu().numpy()
    counters[...] = cn.cpu().numpy()
    if integral is not None:
        integral[...] = it.cpu().numpy()
    if saturation is not None:
        saturation[...] = sa.cpu().numpy()
    return {k: v.cpu().numpy() for k, v in out.items()}


def reduce_scores(metrics, C_, eps):
    c = ctx()
    out = torch.zeros((4, C_), dtype=torch.float64, device=c.device)
    c.reduce_scores(dev(metrics), C_, eps, out)
    return out.cpu().numpy()


def reduce_summary(metrics):
    c = ctx()
    out = torch.zeros(16, dtype=torch.float64, device=c.device)
    c.reduce_summary(dev(metrics), metrics.shape[1], out)
    return out.cpu().numpy()


def dare(A, B, Q, R):
    from lqr_quadcopter_test_b200 import solve_dare_batched

    P, K, iters, info = solve_dare_batched(A, B, Q, R)
    return P.cpu().numpy(), K.cpu().numpy(), iters.cpu().numpy(), info.cpu().numpy()


# ---- §8(f).2: deep tracking policy -------------------------------------------------------------------
def make_policy(golden_npz, name):
    """DeepTrackingPolicy drop-in object loaded with the fixture weights."""
    import golden_cases as gc
    from lqr_quadcopter_test_b200 import DeepTrackingPolicy

    hidden, act, bounds = gc.POLICIES[name]
    cfg = {"hidden_sizes": hidden, "activation": act}
    if bounds:
        cfg["output_bounds"] = bounds
    pol = DeepTrackingPolicy(config=cfg)
    pol.load_arrays(*gc.policy_arrays(golden_npz, name))
    return pol


def policy_action(policy, obs):
    return policy.compute_action_batched(dev(obs)).cpu().numpy()


def rollout_policy(params, state0, tparams, policy, K=None, k_stride=None, max_steps=3000, traj_stride=0, data_stride=0):
    c = ctx()
    E = state0.shape[1]
    k_cols = k_cols_of(params)
    Kt = None
    if K is not None:
        Kt = dev(np.asarray(K).reshape(-1, 4, k_cols))
        if k_stride is None:
            k_stride = E if Kt.shape[0] == 1 else 1
    f64 = dict(dtype=torch.float64, device=c.device)
    out = {"metrics": torch.zeros((16, E), **f64), "final_state": torch.zeros((12, E), **f64)}
    ts = dsl = 0
    if traj_stride > 0:
        ts = (max_steps + traj_stride - 1) // traj_stride
        out["traj"] = torch.zeros((ts, 24, E), **f64)
    if data_stride > 0:
        dsl = (max_steps + data_stride - 1) // data_stride
        out["features"] = torch.zeros((dsl, E, 18), dtype=torch.float32, device=c.device)
        out["actions"] = torch.zeros((dsl, E, 4), dtype=torch.float32, device=c.device)
    c.rollout_policy(params, E, max_steps, dev(state0), dev(tparams), policy.packed_mlp(), Kt, k_cols, k_stride or E,
                     out["metrics"], out["final_state"], out.get("traj"), max(traj_stride, 1), ts,
                     out.get("features"), out.get("actions"), max(data_stride, 1), dsl)
    return {k: v.cpu().numpy() for k, v in out.items()}
