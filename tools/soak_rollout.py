"""Soak: full 30 s closed loops of the CUDA path against the CPU oracle on many episodes, with the report the
north star asks for: trajectory difference, and the COUNT of episodes whose on-target counters or success flags
differ (samples sitting on the 0.5 m threshold within round-off).

    python tools/soak_rollout.py [episodes_per_config]"""
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np

import gpu_backend as gb
import oracle
from lqr_quadcopter_test_b200 import RiccatiLQRController, SuccessCriteria
from lqr_quadcopter_test_b200 import params as P
from lqr_quadcopter_test_b200.config import EnvConfig

E = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
LQI = {"use_lqi": True, "q_int": [0.01, 0.01, 0.1]}
for name, cfg, motion in (("config 2: LQR, linear", {}, "linear"), ("LQR, stationary", {}, "stationary"),
                          ("config 3: LQI, circular", LQI, "circular"), ("config 3: LQI, sinusoidal", LQI, "sinusoidal")):
    env = EnvConfig()
    env.target.motion_type = motion
    ctrl = RiccatiLQRController(config=dict(cfg))
    prm = ctrl.kernel_params(env, SuccessCriteria())
    seeds = np.arange(42, 42 + E)
    s0, tp = oracle.reset_reference(prm, seeds)
    K = np.ascontiguousarray(ctrl.effective_gain())
    ref = oracle.rollout(prm, s0, tp, K, max_steps=3002, n_threads=os.cpu_count() or 8)
    got = gb.rollout(prm, s0, tp, K, max_steps=3002)
    M, G = got["metrics"], ref["metrics"]
    per_ep = np.abs(got["final_state"] - ref["final_state"]).max(axis=0)
    d_state = per_ep.max()
    pct = np.percentile(per_ep, [50, 99, 99.9])
    cnt = ((M[P.M_ENV_ON_TARGET] != G[P.M_ENV_ON_TARGET]) | (M[P.M_MET_ON_TARGET] != G[P.M_MET_ON_TARGET])).sum()
    flags = ((M[P.M_ENV_SUCCESS] != G[P.M_ENV_SUCCESS]) | (M[P.M_MET_SUCCESS] != G[P.M_MET_SUCCESS])).sum()
    other = sum(int((M[r] != G[r]).sum()) for r in (P.M_STEPS, P.M_TERMINATION, P.M_VIOLATIONS, P.M_OVERSHOOT_COUNT,
                                                      P.M_SATURATIONS))
    print(f"{name}: {E} episodes x {int(M[P.M_STEPS].mean())} steps: max |final state difference| = {d_state:.2e} "
          f"(median {pct[0]:.1e}, 99 % {pct[1]:.1e}, 99.9 % {pct[2]:.1e}, above 1e-8: {int((per_ep > 1e-8).sum())}); "
          f"episodes with an on-target count differing (threshold-borderline) = {int(cnt)}; success flags differing = "
          f"{int(flags)}; other counters differing = {other}; success rate {M[P.M_MET_SUCCESS].mean():.4f}", flush=True)
