import sys, torch
sys.path.insert(0, ".")
from lqr_quadcopter_test_b200 import EnvConfig, RiccatiLQRController, SuccessCriteria, _lib, rollout_device
ctx = _lib.get_context()
E = 1 << 20
seeds = torch.arange(E, dtype=torch.int64, device=ctx.device) + 42
for cfg in ({}, {"use_lqi": True, "q_int": [0.01, 0.01, 0.1]}):
    for motion in ("stationary", "linear", "circular", "sinusoidal", "figure8"):
        env = EnvConfig(); env.target.motion_type = motion
        ctrl = RiccatiLQRController(config=cfg)
        for _ in range(2): rollout_device(ctrl, env, SuccessCriteria(), seeds)
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        evs[0].record()
        for i in range(5):
            r = rollout_device(ctrl, env, SuccessCriteria(), seeds); evs[i + 1].record()
        torch.cuda.synchronize()
        ms = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(5))[2]
        print(f"{'lqi' if cfg else 'lqr'} {motion:10s} {float(r['metrics'][0].sum()) / ms * 1e3:.4e} steps/s ({ms:.1f} ms)", flush=True)
