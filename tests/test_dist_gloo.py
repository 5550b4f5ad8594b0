"""CPU: the multi-rank reduction path with world_size 2 over gloo."""
import numpy as np
import torch
import torch.multiprocessing as mp

from lqr_quadcopter_test_b200.metrics import SuccessCriteria


def _shard_sums(ratios, errs):
    s = np.zeros(16)
    n = len(ratios)
    s[0], s[1], s[2], s[3], s[4] = n, (ratios >= 0.8).sum(), ratios.sum(), errs.sum(), 10.0 * n
    s[5], s[6] = ((ratios - ratios.mean()) ** 2).sum(), ((errs - errs.mean()) ** 2).sum()
    s[7], s[8], s[9], s[10] = ratios.max(), np.argmax(ratios), ratios.min(), np.argmin(ratios)
    return s


def _worker(rank, size, port, total, out_q):
    import os

    os.environ.update(RANK=str(rank), WORLD_SIZE=str(size), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from lqr_quadcopter_test_b200 import dist as qd

    qd.init_process_group("gloo")
    rng = np.random.default_rng(0)
    ratios, errs = np.round(rng.uniform(0.5, 1, total), 2), rng.uniform(0, 1, total)
    lo, hi = qd.shard_range(total, rank, size)
    local = torch.from_numpy(_shard_sums(ratios[lo:hi], errs[lo:hi]))
    summary, merged = qd.reduce_summary_across_ranks(local, lo, SuccessCriteria())
    # tuning: per-candidate rows of unequal shard widths
    C = 11
    clo, chi = qd.shard_range(C, rank, size)
    cand = torch.arange(4 * C, dtype=torch.float64).reshape(4, C)[:, clo:chi].contiguous()
    full = qd.gather_candidate_scores(cand, C)

    # the tuner's sharded candidate evaluation with a stand-in for the GPU evaluation: candidate order, metrics
    # dictionaries and unequal shards
    class FakeTuner:
        class config:
            evaluation_episodes = 3

        calls = []

        def evaluate_configs(self, cfgs):
            self.calls.append(len(cfgs))
            return ([c["x"] * 0.5 for c in cfgs],
                    [{"mean_on_target_ratio": c["x"] / 100, "mean_tracking_error": 1.0 + c["x"], "success_rate": 1.0,
                      "episodes_evaluated": 3} for c in cfgs])

    ft = FakeTuner()
    cfgs = [{"x": float(i)} for i in range(7)]
    sc, me = qd.evaluate_candidates_sharded(ft, cfgs)
    assert ft.calls == [4 if rank == 0 else 3]
    assert sc == [0.5 * i for i in range(7)] and [m["mean_tracking_error"] for m in me] == [1.0 + i for i in range(7)]
    assert all(m["episodes_evaluated"] == 3 and m["success_rate"] == 1.0 for m in me)
    # a state-carrying controller under Evaluator semantics (LQI, eval.py:198-213): rank 0 owns the whole chain, the
    # other rank contributes an EMPTY shard (n = 0) and receives the summary and the controller's final state
    lo2, hi2, chain = qd.evaluation_shard_plan(total, rank, size, carries_over=True)
    assert (lo2, hi2, chain) == ((0, total, total) if rank == 0 else (0, 0, 1))
    sums2 = torch.from_numpy(_shard_sums(ratios, errs)) if rank == 0 else torch.zeros(16, dtype=torch.float64)
    state = torch.tensor([1.5, -2.5, 3.5, 7.0], dtype=torch.float64) if rank == 0 else torch.zeros(4, dtype=torch.float64)
    summary2, merged2 = qd.reduce_summary_across_ranks(sums2, lo2, SuccessCriteria(), extra=state)
    assert summary2.total_episodes == total and abs(summary2.mean_on_target_ratio - ratios.mean()) < 1e-13
    assert summary2.best_episode_idx == int(np.argmax(ratios))
    assert merged2["extra"][0].tolist() == [1.5, -2.5, 3.5, 7.0]
    # candidate rows with parameter rows riding along (ControllerTuner.sweep): unequal shards, numpy in / out
    rows = np.arange(6 * C, dtype=np.float64).reshape(6, C)
    assert qd.gather_candidate_rows(rows[:, clo:chi], C).tolist() == rows.tolist()
    assert qd.evaluate_candidates_sharded(ft, []) == ([], [])    # nothing to score: returns before any collective
    out_q.put((rank, summary.mean_on_target_ratio, summary.std_on_target_ratio, summary.best_episode_idx,
               summary.worst_episode_idx, summary.total_episodes, full.numpy().tolist()))
    torch.distributed.destroy_process_group()


def test_two_rank_reduction_matches_single_process():
    total, size, port = 1001, 2, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, size, port, total, q)) for r in range(size)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(size)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    ratios = np.round(rng.uniform(0.5, 1, total), 2)
    for _rank, mean, std, best, worst, n, full in results:
        assert n == total
        assert abs(mean - ratios.mean()) < 1e-13 and abs(std - ratios.std()) < 1e-12
        assert best == int(np.argmax(ratios)) and worst == int(np.argmin(ratios))
        assert full == torch.arange(44, dtype=torch.float64).reshape(4, 11).numpy().tolist()
