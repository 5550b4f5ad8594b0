"""Time the BASELINE.json configurations 2-5 on one GPU (device-resident, CUDA events).

    python tools/bench_configs.py [--scale 1.0]

config 2: 2^20 episodes, linear target, LQR           config 3: 2^20 episodes LQI, circular / sinusoidal
config 4: 65 536 candidates x 64 episodes autotune    config 5: teacher data set (stride 10), scaled to fit
"""
import argparse
import json
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from lqr_quadcopter_test_b200 import (ControllerTuner, EnvConfig, GainSearchSpace, RiccatiLQRController,
                                      SuccessCriteria, TuningConfig, _lib, rollout_device, teacher_dataset)

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0, help="fraction of the full episode counts")
args = ap.parse_args()
ctx = _lib.get_context()
dev = ctx.device
out = {}


def timed(fn, reps=5):
    """Median of `reps` individually timed calls after two warm-up calls (the first launches of a process
    run at ramping clocks; a mean of two once showed 78 ms for a 61 ms kernel)."""
    fn()
    fn()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    evs[0].record()
    for i in range(reps):
        r = fn()
        evs[i + 1].record()
    torch.cuda.synchronize()
    per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(reps))
    return per[len(per) // 2], r


E = int((1 << 20) * args.scale)
seeds = torch.arange(E, dtype=torch.int64, device=dev) + 42
for name, cfg, motion in (("config2_lqr_linear", {}, "linear"), ("lqr_stationary", {}, "stationary"),
                          ("lqr_circular", {}, "circular"), ("lqr_sinusoidal", {}, "sinusoidal"),
                          ("config3_lqi_circular", {"use_lqi": True, "q_int": [0.01, 0.01, 0.1]}, "circular"),
                          ("config3_lqi_sinusoidal", {"use_lqi": True, "q_int": [0.01, 0.01, 0.1]}, "sinusoidal")):
    env = EnvConfig()
    env.target.motion_type = motion
    ctrl = RiccatiLQRController(config=cfg)
    ms, r = timed(lambda: rollout_device(ctrl, env, SuccessCriteria(), seeds))
    steps = float(r["metrics"][0].sum().item())
    out[name] = {"episodes": E, "ms": ms, "env_steps_per_s": steps / ms * 1e3, "includes": "reset + rollout"}
    print(name, out[name], flush=True)

# config 4: autotune sweep, end to end through ControllerTuner.evaluate_configs (host sampling included)
C = int(65536 * args.scale)
space = GainSearchSpace(q_pos_range=([5e-5, 5e-5, 10.0], [5e-4, 5e-4, 25.0]), q_vel_range=([1e-3, 1e-3, 2.0], [1e-2, 1e-2, 8.0]),
                        r_controls_range=([0.5] * 4, [2.0] * 4))
tc = TuningConfig(controller_type="riccati_lqr", search_space=space, strategy="random", max_iterations=C,
                  evaluation_episodes=64, evaluation_horizon=3000, seed=42, output_dir="/tmp/qt_tuning")
tuner = ControllerTuner(tc)
import time

t0 = time.perf_counter()
configs = [tuner._generate_random_config() for _ in range(C)]
t_sample = time.perf_counter() - t0
ms, (scores, _m) = timed(lambda: tuner.evaluate_configs(configs), reps=1)
out["config4_autotune"] = {"candidates": C, "episodes_per_candidate": 64, "host_sampling_s": t_sample, "evaluate_ms": ms,
                           "env_steps_per_s": C * 64 * 3000 / ms * 1e3, "candidates_per_s": C / ms * 1e3,
                           "best_score": float(max(scores)), "includes": "host lists -> DARE -> rollout -> scores -> host"}
print("config4_autotune", out["config4_autotune"], flush=True)

# config 5: teacher data set; 4M episodes x 300 rows x 88 B = 111 GB: run a slice that fits comfortably
E5 = int((1 << 20) * args.scale)
env = EnvConfig()
env.target.motion_type = "linear"
ctrl = RiccatiLQRController(config={})
ms, r = timed(lambda: teacher_dataset(ctrl, env, num_episodes=E5, base_seed=42, stride=10), reps=1)
rows = r[0].shape[0]
out["config5_teacher_dataset"] = {"episodes": E5, "rows": rows, "ms": ms, "rows_per_s": rows / ms * 1e3,
                                  "write_gbs": rows * 88 / ms * 1e3 / 1e9, "env_steps_per_s": E5 * 3000 / ms * 1e3}
print("config5_teacher_dataset", out["config5_teacher_dataset"], flush=True)
# HBM-bound variant: log EVERY step of every episode (24 fp64 fields = 192 B per env step), the layout
# Evaluator.run_episode consumes.  65 536 episodes x 3000 steps x 192 B = 37.7 GB of writes.
E6 = int(65536 * args.scale)
seeds6 = torch.arange(E6, dtype=torch.int64, device=dev) + 42
env = EnvConfig()
env.target.motion_type = "linear"
ctrl = RiccatiLQRController(config={})
peaks = json.load(open(pathlib.Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json")) if (
    pathlib.Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
import numpy as _np
from lqr_quadcopter_test_b200 import params as P
prm = ctrl.kernel_params(env, SuccessCriteria())
f64 = dict(dtype=torch.float64, device=dev)
s0, tp6 = torch.empty((12, E6), **f64), torch.empty((12, E6), **f64)
ctx.reset_batched(prm, seeds6, s0, tp6)
K6 = torch.as_tensor(_np.ascontiguousarray(ctrl.K), **f64)
met6 = torch.empty((16, E6), **f64)
for stride in (1, 2, 4):
    slots = (3000 + stride - 1) // stride
    log = torch.empty((slots, P.TRAJ_FIELDS, E6), **f64)  # preallocated: the kernel alone is timed
    ms, _ = timed(lambda: ctx.rollout(prm, E6, 3000, 1, s0, tp6, K6, 6, E6, met6, None, None, log, stride, slots), reps=2)
    steps = float(met6[0].sum().item())
    gbs = steps / stride * 192 / ms * 1e3 / 1e9
    out[f"trajectory_log_stride{stride}"] = {"episodes": E6, "ms": ms, "env_steps_per_s": steps / ms * 1e3,
                                             "log_write_gbs": gbs, "hbm_peak_gbs": peaks["hbm_gbs"],
                                             "hbm_frac": gbs / peaks["hbm_gbs"],
                                             "note": "kernel only, log preallocated; 192 B (24 fp64 fields) per logged step; "
                                                     "the peak is a copy (read+write) figure, this is write-only traffic"}
    print(f"trajectory_log_stride{stride}", out[f"trajectory_log_stride{stride}"], flush=True)
    del log
    torch.cuda.empty_cache()
print(json.dumps(out))
