"""Under torchrun: Evaluator.evaluate sharded over the ranks (dist.evaluate_sharded, NCCL all_gather of the 16
summary sums) against the same evaluation done by every rank alone.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_sharded_eval.py"""
import json
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch.distributed as dist

from lqr_quadcopter_test_b200 import EnvConfig, Evaluator, RiccatiLQRController
from lqr_quadcopter_test_b200 import dist as qd

rank, size, local = qd.init_process_group()
out = {}
for motion, cfg in (("stationary", {}), ("linear", {}), ("circular", {"use_lqi": True, "q_int": [0.01, 0.01, 0.1]})):
    env = EnvConfig()
    env.target.motion_type = motion
    ev = Evaluator(RiccatiLQRController(config=dict(cfg)), env, reset_controller_between_episodes=True)
    E = 100_003
    sharded, _ = qd.evaluate_sharded(ev, E, base_seed=42)
    alone = ev.evaluate(num_episodes=E, base_seed=42, verbose=False)
    keys = ("total_episodes", "successful_episodes", "best_episode_idx", "worst_episode_idx", "meets_criteria")
    exact = all(getattr(sharded, k) == getattr(alone, k) for k in keys)
    close = max(abs(getattr(sharded, k) - getattr(alone, k)) / max(abs(getattr(alone, k)), 1e-300)
                for k in ("success_rate", "mean_on_target_ratio", "std_on_target_ratio", "mean_tracking_error",
                          "std_tracking_error", "mean_control_effort"))
    out[motion] = {"counts_and_indices_equal": exact, "max_relative_difference_of_statistics": close,
                   "success_rate": alone.success_rate, "mean_on_target_ratio": alone.mean_on_target_ratio}
    assert exact and close < 1e-12, (motion, exact, close)
# the Evaluator's DEFAULT semantics with a state-carrying controller (LQI, eval.py:198-213: never reset between episodes):
# the episodes are one serial chain, which evaluate_sharded runs on rank 0 and hands to every rank
env = EnvConfig()
ctrl_a = RiccatiLQRController(config={"use_lqi": True, "q_int": [1e-5, 1e-5, 1e-3]})
ctrl_b = RiccatiLQRController(config={"use_lqi": True, "q_int": [1e-5, 1e-5, 1e-3]})
sh, _ = qd.evaluate_sharded(Evaluator(ctrl_a, env), 12, base_seed=42)
al = Evaluator(ctrl_b, env).evaluate(num_episodes=12, base_seed=42, verbose=False)
carry_equal = all(getattr(sh, k) == getattr(al, k) for k in ("total_episodes", "successful_episodes", "best_episode_idx",
                                                             "worst_episode_idx", "mean_on_target_ratio", "mean_tracking_error"))
state_equal = bool((ctrl_a.integral_state == ctrl_b.integral_state).all()) and ctrl_a.get_saturation_count() == ctrl_b.get_saturation_count()
assert carry_equal and state_equal, (carry_equal, state_equal)
out["lqi_carry_over_default"] = {"summary_equal": carry_equal, "controller_state_equal_on_every_rank": state_equal}
if rank == 0:
    print(json.dumps({"n_gpus": size, "episodes": 100_003, "results": out}))
if size > 1:
    dist.destroy_process_group()
