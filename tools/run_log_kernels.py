"""Launch the two LOGGING variants of the rollout kernel once each, sized for an `ncu --set full` capture (ncu replays
a kernel ~40 times, so the launches are kept to a few GB of writes):

    python tools/run_log_kernels.py traj      # MODE 1: trajectory log, every step (24 fp64 fields = 192 B per env step)
    python tools/run_log_kernels.py data      # MODE 2: teacher data set, every 10th step (18 + 4 fp32 = 88 B per row)

Prints the CUDA-event time and the algorithmic write rate of the one launch (a number taken under ncu is not a bench
value; run without ncu for that)."""
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from lqr_quadcopter_test_b200 import EnvConfig, RiccatiLQRController, SuccessCriteria, _lib
from lqr_quadcopter_test_b200 import params as P

which = sys.argv[1] if len(sys.argv) > 1 else "traj"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = _lib.get_context()
dev = ctx.device
env = EnvConfig()
env.target.motion_type = "linear"
ctrl = RiccatiLQRController(config={})
prm = ctrl.kernel_params(env, SuccessCriteria())
f64 = dict(dtype=torch.float64, device=dev)
K = torch.as_tensor(np.ascontiguousarray(ctrl.K), **f64)
if which == "traj":
    E, T, stride = 1 << 16, 300, 1
else:
    E, T, stride = 1 << 18, 3000, 10
seeds = torch.arange(E, dtype=torch.int64, device=dev) + 42
s0, tp = torch.empty((12, E), **f64), torch.empty((12, E), **f64)
ctx.reset_batched(prm, seeds, s0, tp)
met = torch.empty((16, E), **f64)
slots = (T + stride - 1) // stride
if which == "traj":
    log = torch.empty((slots, P.TRAJ_FIELDS, E), **f64)
    run = lambda: ctx.rollout(prm, E, T, 1, s0, tp, K, 6, E, met, None, None, log, stride, slots)
    row_bytes = 192
else:
    feat = torch.empty((slots, E, 18), dtype=torch.float32, device=dev)
    act = torch.empty((slots, E, 4), dtype=torch.float32, device=dev)
    run = lambda: ctx.rollout(prm, E, T, 1, s0, tp, K, 6, E, met, None, None, None, 1, 0, feat, act, stride, slots)
    row_bytes = 88
torch.cuda.synchronize()
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    rows = float(met[0].sum().item()) / stride
    print(f"{which}: E={E} T={T} stride={stride}: {ms:.3f} ms, {rows * row_bytes / ms * 1e3 / 1e9:.0f} GB/s algorithmic writes, "
          f"{float(met[0].sum().item()) / ms * 1e3:.3e} env-steps/s", flush=True)
