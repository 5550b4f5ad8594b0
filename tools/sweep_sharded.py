"""BASELINE config 4 under torchrun: 65 536 Q/R candidates x 64 episodes x 3000 steps split over the ranks, timed END
TO END INCLUDING CANDIDATE SAMPLING (ControllerTuner.sweep: draw -> DARE -> rollout with fused scores -> all_gather).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep_sharded.py [C] [reps]

Each rank draws only its contiguous shard of the candidates (the generator is advanced past the others, so candidate
i is the reference's candidate i whatever the partition), solves its DAREs, rolls its candidates out and one
all_gather returns every rank the full score and parameter rows.  Rank 0 prints the wall-clock time (max over ranks)
of the median repetition and checks a slice against a single-rank, dictionary-based evaluation."""
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
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
space = GainSearchSpace(q_pos_range=([5e-5, 5e-5, 10.0], [5e-4, 5e-4, 25.0]), q_vel_range=([1e-3, 1e-3, 2.0], [1e-2, 1e-2, 8.0]),
                        r_controls_range=([0.5] * 4, [2.0] * 4))


def make():
    return ControllerTuner(TuningConfig(controller_type="riccati_lqr", search_space=space, strategy="random", max_iterations=C,
                                        evaluation_episodes=64, evaluation_horizon=3000, seed=42,
                                        output_dir=f"/tmp/qt_sweep_{rank}"))


make().sweep(64 * size)                                               # warm-up: kernels loaded, NCCL communicator up
make().sweep(C)
times = []
for _ in range(reps):
    tuner = make()
    torch.cuda.synchronize()
    if size > 1:
        dist.barrier()
    t0 = time.perf_counter()
    block, rows = tuner.sweep(C)                                      # sampling included; rows are host numpy on return
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if size > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    times.append(float(dt.item()))
lo = 777
single = make()
ref, _ = single.evaluate_configs(block.slice(lo, lo + 256).configs())   # this rank alone, through dictionaries
same = rows[0, lo:lo + 256].tolist() == ref
scalar = make()
first = [scalar._generate_random_config() for _ in range(4)]
if rank == 0:
    s = sorted(times)[len(times) // 2]
    print(json.dumps({"candidates": C, "episodes_per_candidate": 64, "n_gpus": size, "seconds": s, "all_seconds": times,
                      "includes": "candidate sampling (host rng), H2D, DARE, reset, rollout + fused scores, D2H, all_gather",
                      "candidates_per_s": C / s, "env_steps_per_s": C * 64 * 3000 / s, "best_score": float(rows[0].max()),
                      "best_index": int(qd.argmax_first(rows[0])), "slice_matches_single_rank_evaluation": bool(same),
                      "candidates_are_the_scalar_draw_sequence": block.configs()[:4] == first}))
if size > 1:
    dist.destroy_process_group()
