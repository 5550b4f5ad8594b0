"""GPU: qt_evaluate_host (host buffers in and out, the call bench.py's e2e figure times) against the
device-resident path: identical bits, with the copy/compute pipeline on (large batch, four chunks on two
streams) and off (small batch; chains of episodes sharing one controller)."""
import numpy as np
import pytest

from lqr_quadcopter_test_b200.config import EnvConfig

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("E,chain,mode", [(300_007, 1, "lqr"), (1 << 17, 1, "lqi"), (5000, 1, "lqr"), (200_000, 4, "lqi")])
def test_evaluate_host_equals_the_device_path(gpu_ctx, E, chain, mode):
    import torch

    from lqr_quadcopter_test_b200 import RiccatiLQRController, rollout_device
    from lqr_quadcopter_test_b200.metrics import SuccessCriteria

    env = EnvConfig()
    env.target.motion_type = "circular"
    cfg = {"use_lqi": True, "q_int": [0.01, 0.01, 0.1]} if mode == "lqi" else {}
    ctrl = RiccatiLQRController(config=cfg)
    prm = ctrl.kernel_params(env, SuccessCriteria())
    seeds_host = (torch.arange(E, dtype=torch.int64) * 3 + 7).pin_memory()
    K_host = torch.as_tensor(np.ascontiguousarray(ctrl.effective_gain()), dtype=torch.float64).pin_memory()
    metrics_host = torch.full((16, E), float("nan"), dtype=torch.float64).pin_memory()
    T = 700
    gpu_ctx.bind_torch_stream()
    for _ in range(2):   # the second call reuses the scratch allocation and both streams
        metrics_host.fill_(float("nan"))
        gpu_ctx.evaluate_host(prm, seeds_host, T, chain, K_host, ctrl.k_cols(), metrics_host)
        assert bool(torch.isfinite(metrics_host).all())
    ref = rollout_device(ctrl, env, SuccessCriteria(), seeds_host.to(gpu_ctx.device), max_steps=T, chain=chain,
                         ctrl_state=(torch.zeros((4, (E + chain - 1) // chain), dtype=torch.float64, device=gpu_ctx.device)
                                     if chain > 1 else None))
    assert torch.equal(metrics_host, ref["metrics"].cpu())
