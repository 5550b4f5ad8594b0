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


def is_primary() -> bool:
    """True on the rank that writes files (rank 0; always true for a single process)."""
    return world()[0] == 0


def barrier() -> None:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


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


def reduce_summary_across_ranks(local_sums, local_offset: int, criteria=None, extra=None):
    """All ranks contribute the 16 sums of their shard (qt_reduce_summary) and the global index of
    their first episode; every rank returns the same EvaluationSummary (without per-episode
    objects — those stay on the rank that owns them).  ``extra`` (a small 1-D float64 tensor of the same length on
    every rank) rides along in the same all_gather; then the second return value is a dict
    ``{"sums": merged, "extra": [rank 0's, rank 1's, ...]}`` instead of the merged sums."""
    import torch

    dev = local_sums.device
    parts_in = [local_sums.reshape(-1).to(torch.float64), torch.tensor([float(local_offset)], dtype=torch.float64, device=dev)]
    if extra is not None:
        parts_in.append(extra.reshape(-1).to(torch.float64).to(dev))
    parts = [p.cpu().numpy() for p in all_gather_vectors(torch.cat(parts_in))]
    merged = merge_summary_sums([p[:16] for p in parts], [int(p[16]) for p in parts])
    summary = summary_from_sums(merged, criteria, [])
    if extra is None:
        return summary, merged
    return summary, {"sums": merged, "extra": [p[17:] for p in parts]}


def evaluation_shard_plan(num_episodes: int, rank: int, size: int, carries_over: bool) -> tuple[int, int, int]:
    """(lo, hi, chain) of the episodes rank ``rank`` evaluates.

    Episodes are independent — and shard contiguously — unless the controller carries state from one episode into
    the next: ``Evaluator.evaluate`` never resets the controller (eval.py:198-213), so an LQI integrator (or a PID
    integral with a non-zero limit) makes the episodes of one evaluation a serial chain.  That chain cannot be cut:
    rank 0 runs all of it in one GPU thread (chain = num_episodes), the other ranks contribute an empty shard, and
    the final reduction hands every rank the same summary."""
    if carries_over:
        return (0, int(num_episodes), max(int(num_episodes), 2)) if rank == 0 else (0, 0, 1)
    lo, hi = shard_range(num_episodes, rank, size)
    return lo, hi, 1


def evaluate_sharded(evaluator, num_episodes: int, base_seed: int = 42, max_steps: int | None = None):
    """Evaluator.evaluate over all ranks: rank r evaluates its shard of the episodes (seeds base_seed + index) and
    the 16-sum vectors are merged.  Honours ``evaluator.reset_controller_between_episodes`` exactly like the
    single-process ``Evaluator.evaluate``: a state-carrying controller that is not reset runs as one serial chain on
    rank 0 (see evaluation_shard_plan) and its final state is written back to the controller object on every rank.
    Returns (summary, local_metrics [16, hi - lo])."""
    import torch

    from . import _lib
    from .eval import rollout_device

    rank, size, _ = world()
    ctrl = evaluator.controller
    carries = bool(ctrl.carries_state()) and not evaluator.reset_controller_between_episodes
    lo, hi, chain = evaluation_shard_plan(num_episodes, rank, size, carries)
    ctx = _lib.get_context()
    f64 = dict(dtype=torch.float64, device=ctx.device)
    sums = torch.zeros(16, **f64)
    state = torch.zeros(4, **f64)
    metrics = torch.empty((16, 0), **f64)
    if hi > lo:
        seeds = torch.arange(lo, hi, dtype=torch.int64, device=ctx.device) + int(base_seed)
        st = evaluator._ctrl_state_in((hi - lo + chain - 1) // chain) if carries else None
        out = rollout_device(ctrl, evaluator.env_config, evaluator.criteria, seeds, max_steps, chain=chain,
                             ctrl_state=st, ctx=ctx)
        metrics = out["metrics"]
        ctx.reduce_summary(metrics, hi - lo, sums)
        if carries:
            state = st[:, 0].contiguous()
    summary, merged = reduce_summary_across_ranks(sums, lo, evaluator.criteria, extra=state if carries else None)
    if carries:
        ctrl.load_state_vector(merged["extra"][0], True)
    return summary, metrics


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
    """Tuning: each rank scored candidates [lo, hi); gather the [R, C_local] rows (R = 4 score rows, optionally
    followed by the candidates' parameter rows) into [R, C] on every rank — an all_gather of <= 2 MiB for the scores
    of 65 536 candidates.  Shards may differ by one candidate, so rows are padded to the largest shard before the
    collective."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_cand
    size = dist.get_world_size()
    widths = [shard_range(total_candidates, r, size) for r in range(size)]
    wmax = max(hi - lo for lo, hi in widths)
    rows = local_cand.shape[0]
    pad = torch.zeros((rows, wmax), dtype=local_cand.dtype, device=local_cand.device)
    pad[:, : local_cand.shape[1]] = local_cand
    out = torch.empty((size * rows, wmax), dtype=local_cand.dtype, device=local_cand.device)
    dist.all_gather_into_tensor(out, pad)        # rank r's block = rows [r * rows, (r + 1) * rows)
    out = out.reshape(size, rows, wmax)
    return torch.cat([out[r, :, : hi - lo] for r, (lo, hi) in enumerate(widths)], dim=1)


def _collective_device():
    import torch
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def gather_candidate_rows(local_rows: np.ndarray, total_candidates: int) -> np.ndarray:
    """numpy front of gather_candidate_scores: [R, C_local] in, [R, C] out (same on every rank)."""
    import torch

    t = torch.as_tensor(np.ascontiguousarray(local_rows), dtype=torch.float64, device=_collective_device())
    return gather_candidate_scores(t, total_candidates).cpu().numpy()


def evaluate_candidates_sharded(tuner, configs: list[dict]) -> tuple[list[float], list[dict]]:
    """ControllerTuner.evaluate_configs over all ranks for candidates given as dictionaries (grid search, resumed
    runs): the same list on every rank, rank r scores candidates [lo, hi) — every candidate sees the same episode
    seeds wherever it runs (tuning.py:875), so the scores do not depend on the partition — and one all_gather of the
    [4, C_local] score rows gives every rank the full lists, in candidate order.  (Random search does not come
    through here: ControllerTuner.sweep draws and scores arrays.)"""
    import torch.distributed as dist

    rank, size, _ = world()
    if not configs:
        return [], []          # nothing to score: no collective either
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tuner.evaluate_configs(configs)
    C = len(configs)
    lo, hi = shard_range(C, rank, size)
    scores, metrics = tuner.evaluate_configs(configs[lo:hi]) if hi > lo else ([], [])
    eps = metrics[0]["episodes_evaluated"] if metrics else int(tuner.config.evaluation_episodes)
    rows = np.array([scores, [m["mean_on_target_ratio"] for m in metrics], [m["mean_tracking_error"] for m in metrics],
                     [m["success_rate"] for m in metrics]], dtype=np.float64).reshape(4, hi - lo)
    full = gather_candidate_rows(rows, C)
    out = [{"mean_on_target_ratio": a, "mean_tracking_error": b, "success_rate": c, "episodes_evaluated": eps}
           for a, b, c in zip(full[1].tolist(), full[2].tolist(), full[3].tolist())]
    return full[0].tolist(), out


def argmax_first(scores: np.ndarray) -> int:
    """Index of the best candidate with the reference's rule — a scan that replaces the incumbent only on a
    strict '>' (tuning.py:1084), so the first of equal scores wins and NaN scores never do (np.argmax would return
    the first NaN).  -1 if there is no comparable score at all."""
    scores = np.asarray(scores, dtype=np.float64)
    ok = ~np.isnan(scores)
    if not ok.any():
        return -1
    return int(np.argmax(np.where(ok, scores, -np.inf)))
