import sys, torch
sys.path.insert(0, ".")
from lqr_quadcopter_test_b200 import EnvConfig, RiccatiLQRController, _lib, teacher_dataset
ctx = _lib.get_context()
env = EnvConfig(); env.target.motion_type = "linear"
ctrl = RiccatiLQRController(config={})
E = 1 << 20
for stride in (10, 1):
    Ee = E if stride == 10 else E // 16
    for _ in range(2): r = teacher_dataset(ctrl, env, num_episodes=Ee, base_seed=42, stride=stride); del r
    torch.cuda.synchronize()
    ts = []
    for i in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = teacher_dataset(ctrl, env, num_episodes=Ee, base_seed=42, stride=stride); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)); rows = r[0].shape[0]; del r
    ms = sorted(ts)[2]
    print(f"stride {stride}: {ms:.1f} ms, {rows / ms * 1e3:.3e} rows/s, {rows * 88 / ms * 1e3 / 1e9:.0f} GB/s, {Ee * 3000 / ms * 1e3:.3e} env-steps/s", flush=True)
