"""GPU: imitation training with the closed loop on the device (train.py:275-1028)."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_imitation_training_learns_the_supervisor(gpu_ctx, tmp_path):
    import torch

    from lqr_quadcopter_test_b200 import Evaluator, SuccessCriteria, Trainer, TrainingConfig, load_controller
    from lqr_quadcopter_test_b200.config import EnvConfig

    torch.manual_seed(0)
    cfg = TrainingConfig(training_mode="imitation", supervisor_controller="riccati_lqr", epochs=8, episodes_per_epoch=192,
                         max_steps_per_episode=500, episode_length=5.0, batch_size=128, learning_rate=3e-3,
                         tracking_weight=0.0, imitation_weight=1.0, target_motion_type="stationary",
                         checkpoint_dir=str(tmp_path / "ck"), log_dir=str(tmp_path / "logs"), checkpoint_interval=4)
    tr = Trainer(cfg)
    summary = tr.train()
    log = tr.experiment_log
    assert summary["epochs_completed"] == 8 and len(log) == 8
    assert log[-1]["imitation"] < 0.1 * log[0]["imitation"], [e["imitation"] for e in log]
    # the student, driving the env itself, ends up tracking far better than at initialisation
    assert log[-1]["mean_tracking_error"] < 0.5 * log[0]["mean_tracking_error"]
    files = sorted(p.name for p in (tmp_path / "ck").iterdir())
    assert any(f.endswith("_best.pt") for f in files) and any(f.endswith("_final.pt") for f in files)
    assert any("_epoch0003" in f for f in files) and any("_epoch0007" in f for f in files)
    assert json.loads(next((tmp_path / "logs").glob("*_log.json")).read_text())[0]["controller"] == "deep"
    # the checkpoint loads through the evaluation entry point and behaves like the trained network
    best = next((tmp_path / "ck").glob("*_final.pt"))
    env = EnvConfig()
    env.target.motion_type = "stationary"
    env.simulation.max_episode_time = 5.0
    ev = Evaluator(load_controller("deep", checkpoint_path=best, config={"hidden_sizes": [64, 64]}), env,
                   SuccessCriteria(min_episode_duration=5.0), output_dir=tmp_path / "rep")
    s = ev.evaluate(num_episodes=64, base_seed=7, verbose=False, record_trajectories=False)
    assert s.total_episodes == 64 and np.isfinite(s.mean_tracking_error)
    assert s.mean_tracking_error < 0.9        # initial offsets are up to 0.87 m; an untrained net drifts metres away


def test_trainer_evaluates_classical_controllers(gpu_ctx, tmp_path):
    """train.py:578-652: controller != deep runs evaluation epochs only; numbers equal the Evaluator's."""
    import torch

    from lqr_quadcopter_test_b200 import RiccatiLQRController, SuccessCriteria, Trainer, TrainingConfig, rollout_device
    from lqr_quadcopter_test_b200 import params as P
    from lqr_quadcopter_test_b200.config import EnvConfig

    cfg = TrainingConfig(controller="riccati_lqr", epochs=2, episodes_per_epoch=5, target_motion_type="linear",
                         log_dir=str(tmp_path / "logs"), checkpoint_dir=str(tmp_path / "ck"))
    tr = Trainer(cfg)
    summary = tr.train()
    assert summary["controller"] == "riccati_lqr" and summary["epochs_completed"] == 2 and "checkpoint_dir" not in summary
    env = EnvConfig()
    env.target.motion_type = "linear"
    seeds = torch.arange(5, dtype=torch.int64, device=gpu_ctx.device) + 42 + 1000
    out = rollout_device(RiccatiLQRController(config={}), env, SuccessCriteria(), seeds)
    M = out["metrics"].cpu().numpy()
    e1 = tr.experiment_log[1]
    assert abs(e1["mean_on_target_ratio"] - (M[P.M_ENV_ON_TARGET] / M[P.M_STEPS]).mean()) < 1e-12
    assert abs(e1["mean_tracking_error"] - M[P.M_FINAL_ERR].mean()) < 1e-12
    assert e1["total"] == pytest.approx(e1["mean_tracking_error"])
