"""Time the policy kernel with the LQR-fitted 64-64 fixture network (tests/golden/policy.npz): every episode
tracks for its full 3000 steps, so no lane of a warp idles — the regime of a trained policy."""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch

import gpu_backend as gb
from lqr_quadcopter_test_b200 import EnvConfig, SuccessCriteria, _lib, rollout_device

ctx = _lib.get_context()
E = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 3 * 128 * 4
pol = gb.make_policy(np.load(ROOT / "tests" / "golden" / "policy.npz"), "relu_64_64")
seeds = torch.arange(E, dtype=torch.int64, device=ctx.device) + 42
for motion in ("stationary", "linear"):
    env = EnvConfig()
    env.target.motion_type = motion
    for _ in range(2):
        rollout_device(pol, env, SuccessCriteria(), seeds)
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    evs[0].record()
    for i in range(3):
        r = rollout_device(pol, env, SuccessCriteria(), seeds)
        evs[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(3))[1]
    steps = float(r["metrics"][0].sum().item())
    print(f"{motion}: {steps / ms * 1e3:.3e} env-steps/s ({ms:.0f} ms, mean episode length {steps / E:.0f}, "
          f"{steps * 5504 * 2 / ms * 1e3 / 1e12:.1f} TFLOP/s fp32)")
