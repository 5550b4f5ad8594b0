"""Multi-GPU plumbing: one process per GPU, episodes / candidates sharded by rank.

Episodes are independent given (seed, K) and tuning candidates are independent, so the data path
needs no collective: rank r rolls out its contiguous block and the only exchange
is the final metric reduction — an all_gather of one 16-double vector per rank (evaluation) or of
the per-candidate score rows (tuning).  ``torch.distributed`` carries it (NCCL over NVLink on the
GPU box, gloo in the CPU tests); message sizes are bytes to a few MiB, so latency is all that
matters and there is nothing to fuse a collective into.
"""
from __future__ import annotations

import os

import numpy as np

from .metrics import merge_summary_sums, summary_from_sums


def world() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)),
            int(os.environ.get("LOCAL_RANK", 0)))


def init_process_group(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    import torch
    import torch.distributed as dist

    rank, size, local = world()
    if size > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=size, **kw)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, size, local


def shard_range(total: int, rank: int, size: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of ``total`` units owned by ``rank`` (sizes differ by at most 1)."""
    base, rem = divmod(int(total), int(size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_vectors(vec, device=None):
    """Gather one equally-sized 1-D float64 tensor per rank; returns a list ordered by rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [vec]
    parts = [torch.empty_like(vec) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, vec.contiguous())
    return parts


def reduce_summary_across_ranks(local_sums, local_offset: int, criteria=None):
    """All ranks contribute the 16 sums of their shard (qt_reduce_summary) and the global index of
    their first episode; every rank returns the same EvaluationSummary (without per-episode
    objects — those stay on the rank that owns them)."""
    import torch

    payload = torch.cat([local_sums.reshape(-1).to(torch.float64),
                         torch.tensor([float(local_offset)], dtype=torch.float64, device=local_sums.device)])
    parts = [p.cpu().numpy() for p in all_gather_vectors(payload)]
    merged = merge_summary_sums([p[:16] for p in parts], [int(p[16]) for p in parts])
    return summary_from_sums(merged, criteria, []), merged


def evaluate_sharded(evaluator, num_episodes: int, base_seed: int = 42, max_steps: int | None = None):
    """Evaluator.evaluate over all ranks: rank r evaluates episodes [lo, hi) with seeds
    base_seed + index, then the 16-sum vectors are merged.  Returns (summary, local_metrics_host)."""
    import torch

    from . import _lib
    from .eval import rollout_device

    rank, size, _ = world()
    lo, hi = shard_range(num_episodes, rank, size)
    ctx = _lib.get_context()
    seeds = torch.arange(lo, hi, dtype=torch.int64, device=ctx.device) + int(base_seed)
    out = rollout_device(evaluator.controller, evaluator.env_config, evaluator.criteria, seeds, max_steps, ctx=ctx)
    sums = torch.zeros(16, dtype=torch.float64, device=ctx.device)
    ctx.reduce_summary(out["metrics"], hi - lo, sums)
    summary, _ = reduce_summary_across_ranks(sums, lo, evaluator.criteria)
    return summary, out["metrics"]


def teacher_dataset_sharded(controller, env_config=None, num_episodes: int = 1 << 22, base_seed: int = 42,
                            stride: int = 10, **kw):
    """BASELINE config 5 over all ranks: rank r emits the teacher data set of episodes [lo, hi) (seeds
    base_seed + index) into its own GPU's tensors — 4 M episodes at stride 10 are 111 GB in all, 13.8 GB per GPU
    on eight — and no collective is needed: each rank hands its shard to its own trainer process.
    Returns ``(features, actions, metrics, (lo, hi))``."""
    from .dataset import teacher_dataset

    rank, size, _ = world()
    lo, hi = shard_range(num_episodes, rank, size)
    f, a, m = teacher_dataset(controller, env_config, num_episodes=hi - lo, base_seed=int(base_seed) + lo, stride=stride, **kw)
    return f, a, m, (lo, hi)


def gather_candidate_scores(local_cand, total_candidates: int):
    """Tuning: each rank scored candidates [lo, hi); gather the [4, C_local] rows into [4, C] on every
    rank (an all_gather of <= 2 MiB for 65 536 candidates).  Shards may differ by one candidate, so
    rows are padded to the largest shard before the collective."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_cand
    size = dist.get_world_size()
    widths = [shard_range(total_candidates, r, size) for r in range(size)]
    wmax = max(hi - lo for lo, hi in widths)
    pad = torch.zeros((4, wmax), dtype=local_cand.dtype, device=local_cand.device)
    pad[:, : local_cand.shape[1]] = local_cand
    parts = [torch.empty_like(pad) for _ in range(size)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:, : hi - lo] for p, (lo, hi) in zip(parts, widths)], dim=1)


def evaluate_candidates_sharded(tuner, configs: list[dict]) -> tuple[list[float], list[dict]]:
    """ControllerTuner.evaluate_configs over all ranks (BASELINE config 4: the candidates of a Q/R sweep,
    each with its DARE solve and its evaluation episodes, split contiguously over the GPUs).  Rank r scores
    candidates [lo, hi) — every candidate sees the same episode seeds wherever it runs (tuning.py:875), so
    the scores do not depend on the partition — and one all_gather of the [4, C_local] score rows gives
    every rank the full lists, in candidate order."""
    import torch
    import torch.distributed as dist

    rank, size, _ = world()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tuner.evaluate_configs(configs)
    C = len(configs)
    lo, hi = shard_range(C, rank, size)
    scores, metrics = tuner.evaluate_configs(configs[lo:hi]) if hi > lo else ([], [])
    eps = metrics[0]["episodes_evaluated"] if metrics else int(tuner.config.evaluation_episodes)
    rows = np.array([scores, [m["mean_on_target_ratio"] for m in metrics], [m["mean_tracking_error"] for m in metrics],
                     [m["success_rate"] for m in metrics]], dtype=np.float64).reshape(4, hi - lo)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    full = gather_candidate_scores(torch.as_tensor(rows, device=dev), C).cpu().numpy()
    out = [{"mean_on_target_ratio": a, "mean_tracking_error": b, "success_rate": c, "episodes_evaluated": eps}
           for a, b, c in zip(full[1].tolist(), full[2].tolist(), full[3].tolist())]
    return full[0].tolist(), out


def argmax_first(scores: np.ndarray) -> int:
    """Best candidate with the reference's tie rule (strict '>' while scanning, tuning.py:1084)."""
    return int(np.argmax(scores))
