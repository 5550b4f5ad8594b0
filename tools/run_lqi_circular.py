"""One kind of launch only (for ncu): the lean rollout kernel with the default LQI gains on the circular target.
usage: python tools/run_lqi_circular.py [episodes] [launches]"""
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch

from lqr_quadcopter_test_b200 import EnvConfig, RiccatiLQRController, SuccessCriteria, _lib, rollout_device

E = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = _lib.get_context()
seeds = torch.arange(E, dtype=torch.int64, device=ctx.device) + 42
env = EnvConfig()
env.target.motion_type = "circular"
ctrl = RiccatiLQRController(config={"use_lqi": True, "q_int": [0.01, 0.01, 0.1]})
for _ in range(n):
    out = rollout_device(ctrl, env, SuccessCriteria(), seeds)
torch.cuda.synchronize()
print(float(out["metrics"][0].sum().item()))
