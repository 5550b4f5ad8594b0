import sys, copy, torch, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from lqr_quadcopter_test_b200 import EnvConfig, RiccatiLQRController, SuccessCriteria, _lib
ctx = _lib.get_context(); ctx.bind_torch_stream()
E = 1 << 20
f64 = dict(dtype=torch.float64, device=ctx.device)
for cfg, motion in (({}, "linear"), ({"use_lqi": True, "q_int": [0.01, 0.01, 0.1]}, "circular")):
    env = EnvConfig(); env.target.motion_type = motion
    ctrl = RiccatiLQRController(config=cfg)
    prm = ctrl.kernel_params(env, SuccessCriteria())
    seeds = torch.arange(E, dtype=torch.int64, device=ctx.device) + 42
    s0, tp = torch.empty((12, E), **f64), torch.empty((12, E), **f64)
    ctx.reset_batched(prm, seeds, s0, tp)
    K = torch.as_tensor(np.ascontiguousarray(ctrl.effective_gain()), **f64)
    met = torch.empty((16, E), **f64)
    for flags in (0, 1):
        p = copy.copy(prm); p.debug_flags = flags
        ts = []
        for i in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ctx.rollout(p, E, 3002, 1, s0, tp, K, ctrl.k_cols(), E, met); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = sorted(ts[1:])[2]
        print(f"{motion} {'lqi' if cfg else 'lqr'} {'general' if flags else 'structured'}: {ms:.1f} ms, {float(met[0].sum()) / ms * 1e3:.3e} env-steps/s")
