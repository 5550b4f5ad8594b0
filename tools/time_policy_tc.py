"""Policy kernel, FP32-pipe (tiled FFMA2) evaluation against the tensor-core evaluation (tcgen05, 3xTF32) of the same
64-64 fixture network: env-steps/s of both and the difference of their results.

    python tools/time_policy_tc.py [episodes] [hidden sizes, e.g. 32,32: a random-init relu network instead of the fixture]
"""
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
E = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 2 * 128 * 4
if len(sys.argv) > 2:
    from lqr_quadcopter_test_b200 import DeepTrackingPolicy

    torch.manual_seed(0)
    pol = DeepTrackingPolicy(config={"hidden_sizes": [int(x) for x in sys.argv[2].split(",")], "activation": "relu"})
else:
    pol = gb.make_policy(np.load(ROOT / "tests" / "golden" / "policy.npz"), "relu_64_64")
seeds = torch.arange(E, dtype=torch.int64, device=ctx.device) + 42
for motion in ("stationary", "linear"):
    env = EnvConfig()
    env.target.motion_type = motion
    res = {}
    for tcores in (False, True):
        pol.tensor_cores = tcores
        for _ in range(2):
            rollout_device(pol, env, SuccessCriteria(), seeds)
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        evs[0].record()
        for i in range(3):
            r = rollout_device(pol, env, SuccessCriteria(), seeds, want_final_state=True)
            evs[i + 1].record()
        torch.cuda.synchronize()
        ms = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(3))[1]
        steps = float(r["metrics"][0].sum().item())
        res[tcores] = r
        print(f"{motion} {'tensor cores (3xTF32)' if tcores else 'FP32 pipe (FFMA2 tiles)'}: {steps / ms * 1e3:.3e} env-steps/s "
              f"({ms:.1f} ms, mean episode length {steps / E:.0f})", flush=True)
    a, b = res[False], res[True]
    d = (a["final_state"] - b["final_state"]).abs().max().item()
    flags = int((a["metrics"][11] != b["metrics"][11]).sum().item())
    on = (a["metrics"][3] - b["metrics"][3]).abs()
    print(f"{motion}: max |final state diff| {d:.3e}; success flags differing {flags} of {E}; on-target counts differing "
          f"{int((on > 0).sum().item())} (max {int(on.max().item())})", flush=True)
