"""Soak: the policy kernel (tiled fp32 evaluation, 64-64 network fitted to the LQR law) against the oracle's
fp32 C network over full 30 s episodes.    python tools/soak_policy.py [episodes]"""
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np

import golden_cases as gc
import gpu_backend as gb
import oracle
from lqr_quadcopter_test_b200 import params as P
from lqr_quadcopter_test_b200.config import EnvConfig

E = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = np.load(ROOT / "tests" / "golden" / "policy.npz")
for name, motion in (("relu_64_64", "stationary"), ("relu_64_64", "circular")):
    pol = gb.make_policy(g, name)
    om = oracle.Mlp(*gc.policy_arrays(g, name), gc.POLICIES[name][1])
    env = EnvConfig()
    env.target.motion_type = motion
    prm = P.make_params(env, gc.controller_ns())
    s0, tp = oracle.reset_reference(prm, np.arange(10_000, 10_000 + E))
    ref = oracle.rollout_policy(prm, s0, tp, om, max_steps=3002, n_threads=os.cpu_count() or 8)
    got = gb.rollout_policy(prm, s0, tp, pol, max_steps=3002)
    M, G = got["metrics"], ref["metrics"]
    per_ep = np.abs(got["final_state"] - ref["final_state"]).max(axis=0)
    cnt = ((M[P.M_ENV_ON_TARGET] != G[P.M_ENV_ON_TARGET]) | (M[P.M_MET_ON_TARGET] != G[P.M_MET_ON_TARGET])).sum()
    flags = ((M[P.M_ENV_SUCCESS] != G[P.M_ENV_SUCCESS]) | (M[P.M_MET_SUCCESS] != G[P.M_MET_SUCCESS])).sum()
    hard = sum(int((M[r] != G[r]).sum()) for r in (P.M_STEPS, P.M_TERMINATION, P.M_VIOLATIONS))
    print(f"{name} / {motion}: {E} episodes x {int(M[P.M_STEPS].mean())} steps: max |final state difference| = {per_ep.max():.2e} "
          f"(median {np.median(per_ep):.1e}, 99.9 % {np.percentile(per_ep, 99.9):.1e}); episodes with an on-target count "
          f"differing = {int(cnt)}; success flags differing = {int(flags)}; steps / termination / violations differing = {hard}",
          flush=True)
