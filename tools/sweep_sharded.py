"""BASELINE config 4 under torchrun: 65 536 Q/R candidates x 64 episodes x 3000 steps split over the ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep_sharded.py [C]

Every rank draws the same candidates, scores its contiguous slice (DARE + rollout + per-candidate means on its GPU)
and one all_gather returns the full score list; rank 0 prints the time (max over ranks, CUDA events around the
evaluation + wall clock around the whole call) and checks a slice against a single-rank evaluation."""
import json
import pathlib
import sys
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist

from lqr_quadcopter_test_b200 import ControllerTuner, GainSearchSpace, TuningConfig
from lqr_quadcopter_test_b200 import dist as qd

rank, size, local = qd.init_process_group()
C = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
space = GainSearchSpace(q_pos_range=([5e-5, 5e-5, 10.0], [5e-4, 5e-4, 25.0]), q_vel_range=([1e-3, 1e-3, 2.0], [1e-2, 1e-2, 8.0]),
                        r_controls_range=([0.5] * 4, [2.0] * 4))
tc = TuningConfig(controller_type="riccati_lqr", search_space=space, strategy="random", max_iterations=C,
                  evaluation_episodes=64, evaluation_horizon=3000, seed=42, output_dir=f"/tmp/qt_sweep_{rank}")
tuner = ControllerTuner(tc)
configs = [tuner._generate_random_config() for _ in range(C)]
qd.evaluate_candidates_sharded(tuner, configs[: 64 * size])          # warm-up: kernels loaded, NCCL communicator up
torch.cuda.synchronize()
if size > 1:
    dist.barrier()
t0 = time.perf_counter()
scores, metrics = qd.evaluate_candidates_sharded(tuner, configs)
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
if size > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
lo = 777
ref, _ = tuner.evaluate_configs(configs[lo:lo + 256])                # this rank alone, whatever slice it owned before
same = scores[lo:lo + 256] == ref
if rank == 0:
    s = float(dt.item())
    print(json.dumps({"candidates": C, "episodes_per_candidate": 64, "n_gpus": size, "seconds": s,
                      "candidates_per_s": C / s, "env_steps_per_s": C * 64 * 3000 / s, "best_score": max(scores),
                      "best_index": int(np.argmax(scores)), "slice_matches_single_rank_evaluation": bool(same)}))
if size > 1:
    dist.destroy_process_group()
