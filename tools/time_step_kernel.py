"""Throughput of the one-step interface (qt_step_batched: reset / step for external policies) and its HBM traffic.

Per env step the kernel reads state 96 B + time 8 + target parameters 96 + counters 24 + action 32 (= 256 B; with the
in-kernel control law the action is replaced by a broadcast gain) and writes state 96 + time 8 + counters 24 +
observation 176 + reward 8 + done 4 + info 32 + validated action 32 (= 380 B): 636 B of algorithmic traffic."""
import json
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch

from lqr_quadcopter_test_b200 import BatchedQuadcopterEnv, EnvConfig, RiccatiLQRController

E = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
cfg = EnvConfig()
cfg.target.motion_type = "linear"
env = BatchedQuadcopterEnv(cfg, num_envs=E)
ctrl = RiccatiLQRController(config={})
peak = json.load(open(pathlib.Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json")).get("hbm_gbs", 6554.9) \
    if (pathlib.Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else 6554.9
for name, kw, nbytes in (("external actions", {"actions": torch.full((4, E), 0.1, dtype=torch.float64, device=env.device)}, 636),
                         ("in-kernel Riccati law", {"controller": ctrl}, 604)):
    env.reset(42)
    for _ in range(5):
        env.step(**kw)
    torch.cuda.synchronize()
    n = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        env.step(**kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gbs = E * nbytes / ms * 1e3 / 1e9
    print(f"{name}: {E} envs, {ms:.3f} ms per step call, {E / ms * 1e3:.3e} env-steps/s, {gbs:.0f} GB/s algorithmic "
          f"= {gbs / peak:.2f} of the measured copy peak {peak:.0f} GB/s")
