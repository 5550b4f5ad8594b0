#!/usr/bin/env python3
"""Headline benchmark: closed-loop env-steps/s of the fused rollout (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's algorithm on host cores

One "step" = one pass of the hot path over one batch: 2^20 episodes per GPU, linear target,
Riccati-LQR default gains, 3000 env steps of dt = 0.01 each (30 s episodes), followed by the
evaluation-summary reduction (and, for N > 1, the one all_gather of 17 doubles per rank).  Weak
scaling: every rank gets its own 2^20 episodes (seeds disjoint).  Prints ONE JSON line.

  value      env-steps/s over all ranks, initial conditions already resident in HBM
  e2e        the same pass through the host-buffer C-ABI call (qt_evaluate_host): pinned host seeds
             in, H2D, reset kernel, rollout kernel, D2H of the [16, E] metric array out
  roofline   rollout kernel vs the FP64 pipe: algorithmic flops per env step of the path that ran (440
             dense; 324 when the gain is axis-decoupled and yaw is zero, see below)
             over the kernel's CUDA-event duration, against a DFMA peak probed in this same run
             (MEASURED_PEAKS.json has no FP64 figure); HBM numbers are given for context
  cpu_baseline  the CPU oracle (C restatement of the reference's loop, all host threads) on a bounded
             sample of the same workload — a reported baseline, not the optimisation target
  dare       batched DARE throughput (second headline metric): 65 536 systems per GPU for n = 6 (LQR), n = 9 (LQI) and
             the 12 x 4 / 15 x 4 sizes the north star names; device-resident, through the host-buffer call qt_dare_host
             (e2e), summed over the ranks, next to scipy's solve_discrete_are + gain on this box (1 core and all cores)
  strong     the literal BASELINE configs at FIXED total size split over the N ranks: config 2 (2^20 episodes), config 4
             (65 536 candidates x 64 episodes end to end INCLUDING candidate sampling), config 5 (2^22-episode teacher
             data set, stride 10)
  policy     SURVEY §8(f).2: the same closed loop driven by the 64-64 deep tracking policy (fp32 network in
             the kernel) on 2^17 episodes, with the Riccati supervisor labelling every 10th step
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

EPISODES_PER_GPU = 1 << 20
STEPS_PER_EPISODE = 3000
FLOPS_PER_ENV_STEP = 440.0           # LQR, linear target, RK4, dense 4x6 gain, 3 angles
# The reference's designs command zero yaw rate through an axis-decoupled gain (6 non-zeros of 24) and
# start at zero yaw; qt_rollout then runs its structure-exploiting loop, which skips the products with
# exact zeros and the identically-zero yaw chain.  Work that path still has to do, itemised like
# mat-vec 44 -> 9, thrust direction 4 x (12 -> 3), rate loop 4 x (12 -> 8), stage
# combinations 156 -> 130 (10 of 12 states), constraints 26 -> 23:  440 - 35 - 36 - 16 - 26 - 3 = 324
# (and 16 instead of 24 sin/cos).  The roofline is quoted on the flops of the path that ran.
FLOPS_PER_ENV_STEP_STRUCTURED = 324.0
DARE_BATCH = 65536
DARE_N = 6
METRIC = "closed-loop env-steps/s (Riccati-LQR, linear target, 30 s episodes at dt=0.01)"


def w_dare(n: int, m: int, k: float) -> float:
    """canonical flop count of one SDA solve with k doubling iterations."""
    return k * (50.0 / 3.0) * n ** 3 + (4 * n * n * m + 2 * n * m * m + (2.0 / 3.0) * m ** 3 + 2 * m * m * n)


def workload_config(n_gpus: int, episodes: int) -> dict:
    return {
        "workload": "BASELINE configs[1]: 2^20 parallel episodes per GPU, linear target motion, Riccati-LQR "
                    "default gains, 30 s at dt=0.01 (3000 env steps per episode), reset(seed=42+i)",
        "episodes_per_gpu": episodes,
        "steps_per_episode": STEPS_PER_EPISODE,
        "parallelism": f"episodes sharded over {n_gpus} GPU(s), no data-path collective",
        "l2_policy": "inputs exceed L2: 24 fp64 initial-condition rows x 2^20 episodes = 201 MB read per pass "
                     "(> 126 MB L2); the kernel is register-resident and FP64-bound, cache state is immaterial",
    }


# ----------------------------------------------------------------------------- clock sampling --
class ClockSampler:
    """nvidia-smi in the background, one line every 20 ms.  It is started BEFORE the warm-up (nvidia-smi needs a few
    hundred ms to come up, longer when eight ranks start theirs at once) and the samples of the timed region are picked
    by their timestamps: mark() right before and right after the region."""

    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None
        self.t0 = self.t1 = None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None
        return self

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.05)      # let the last samples of the region reach the file
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self) -> dict:
        import datetime

        rows = []
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 8:
                    continue
                try:
                    ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(parts[1]), float(parts[2]), float(parts[3]), parts[4:8]))
                except ValueError:
                    continue
        except OSError:
            pass
        finally:
            if self.path and os.path.exists(self.path):
                os.unlink(self.path)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        t0, t1 = self.t0 or rows[0][0], self.t1 or rows[-1][0]
        inside = [r for r in rows if t0 <= r[0] <= t1]
        note = "samples inside the timed region"
        if len(inside) < 3:           # a very short region: the samples closest to it (the GPU is under the same load
            mid = 0.5 * (t0 + t1)     # during the warm-up passes right before)
            inside = sorted(rows, key=lambda r: abs(r[0] - mid))[:5]
            note = "timed region shorter than the sampling period: the five samples closest to it"
        reasons = set()
        for r in inside:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(r[1] for r in inside), "sm_max_mhz": max(r[2] for r in inside),
                "power_w_max": max(r[3] for r in inside), "reasons": sorted(reasons), "samples": len(inside),
                "sampling": f"nvidia-smi -lms 20, {note}"}


# ------------------------------------------------------------------------------ CPU baseline --
def cpu_port_rate(episodes: int, n_threads: int, max_steps: int = STEPS_PER_EPISODE):
    """Time the oracle (oracle/qt_oracle.c) on `episodes` linear-target episodes; returns
    (env-steps/s, seconds).  Initial conditions come from numpy exactly as the reference draws them."""
    import numpy as np

    import oracle
    from lqr_quadcopter_test_b200 import params as P
    from lqr_quadcopter_test_b200.config import EnvConfig

    env = EnvConfig()
    env.target.motion_type = "linear"
    prm = P.make_params(env)
    A, B = oracle.build_systems_reference()
    _, K = oracle.solve_dare_reference(A, B, np.diag([1e-4, 1e-4, 16.0, 3.6e-3, 3.6e-3, 4.0]), np.eye(4))
    s0, tp = oracle.reset_reference(prm, np.arange(42, 42 + episodes))
    t0 = time.perf_counter()
    out = oracle.rollout(prm, s0, tp, K, max_steps=max_steps + 2, n_threads=n_threads, want_final_state=False)
    dt = time.perf_counter() - t0
    steps = float(out["metrics"][0].sum())
    return steps / dt, dt


def sized_cpu_sample(target_seconds: float, n_threads: int) -> int:
    rate, _ = cpu_port_rate(max(2 * n_threads, 16), n_threads, max_steps=600)
    episodes = int(rate * target_seconds / STEPS_PER_EPISODE)
    return max(n_threads, min(episodes, 1 << 16))



def python_reference_timing(cores: int, timeout_s: float = 240.0) -> dict:
    """The UNMODIFIED Python reference (baseline/_ref, see baseline/install_reference.py) timed on this box's host
    cores: Evaluator.evaluate(10, 42), solve_dare x 1000, _evaluate_config 8 x 5 — BASELINE.md §3."""
    script = ROOT / "baseline" / "time_reference.py"
    if not (ROOT / "baseline" / "_ref" / "quadcopter_tracking").is_dir():
        return {"unavailable": "baseline/_ref is not installed (python baseline/install_reference.py needs /root/reference)"}
    try:
        r = subprocess.run([sys.executable, str(script), "--procs", str(cores)], capture_output=True, text=True,
                           timeout=timeout_s, cwd=tempfile.gettempdir())
        if r.returncode != 0:
            return {"unavailable": (r.stderr or r.stdout).strip().splitlines()[-1][:300]}
        out = json.loads(r.stdout.strip().splitlines()[-1])
        out["what"] = ("quadcopter_tracking 0.5.0 imported unmodified from baseline/_ref: Evaluator.evaluate(10 episodes, seed 42, "
                       "riccati_lqr, stationary) on 1 core and in N processes on disjoint seeds; solve_dare over sampled tuning-range "
                       "(Q, R); ControllerTuner._evaluate_config for 8 candidates x 5 episodes (one candidate per process)")
        return out
    except (subprocess.TimeoutExpired, ValueError, IndexError, OSError) as exc:
        return {"unavailable": repr(exc)[:300]}


def _scipy_dare_worker(args):
    """solve_dare (validation + scipy.linalg.solve_discrete_are + gain) on `count` tuning-range systems of size n."""
    n, count, seed = args
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    import numpy as np

    import oracle

    sys.path.insert(0, str(ROOT / "tools"))
    from bench_dare import dare_test_systems

    if n in (6, 9):
        A, B = oracle.build_systems_reference(lqi=(n == 9))
        rng = np.random.default_rng(seed)
        q = rng.uniform([5e-5, 5e-5, 10, 1e-3, 1e-3, 2], [5e-4, 5e-4, 25, 1e-2, 1e-2, 8], (count, 6))
        if n == 9:
            q = np.concatenate([q, rng.uniform([1e-3, 1e-3, 1e-2], [0.1, 0.1, 1.0], (count, 3))], axis=1)
        Q = np.stack([np.diag(x) for x in q])
        R = np.stack([np.diag(x) for x in rng.uniform(0.5, 2.0, (count, 4))])
    else:
        A, B, Q, R = dare_test_systems(n, 4, count, seed)
    oracle.solve_dare_reference(A, B, Q[0], R[0])
    t0 = time.perf_counter()
    for i in range(count):
        oracle.validate_reference(Q[i], R[i])
        oracle.solve_dare_reference(A, B, Q[i], R[i])
    return count, time.perf_counter() - t0


def scipy_dare_rates(n: int, cores: int, count: int = 400) -> dict:
    """The reference's DARE path (scipy) on this box: one core, and all cores (one process per core)."""
    import multiprocessing as mp

    c1, t1 = _scipy_dare_worker((n, count, 0))
    procs = max(1, cores)
    per = max(50, count // 4)
    with mp.get_context("spawn").Pool(procs) as pool:
        pool.map(_scipy_dare_worker, [(n, 2, 0)] * procs)          # imports outside the timed region
        t0 = time.perf_counter()
        res = pool.map(_scipy_dare_worker, [(n, per, 1 + p) for p in range(procs)])
        wall = time.perf_counter() - t0
    return {"value": sum(r[0] for r in res) / wall, "unit": "solves/s", "cores": procs, "kind": "reference",
            "single_core_value": c1 / t1,
            "sample": f"scipy.linalg.solve_discrete_are + validation + gain (the reference's solve_dare, riccati_lqr.py:119-184) on "
                      f"{per} n={n} systems per process x {procs} processes ({wall:.1f} s); single core: {count} systems"}


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    episodes = sized_cpu_sample(6.0, cores)
    for _ in range(args.warmup):
        cpu_port_rate(max(cores, episodes // 8), cores)
    t_total, steps_total = 0.0, 0.0
    for _ in range(args.steps):
        rate, dt = cpu_port_rate(episodes, cores)
        t_total += dt
        steps_total += rate * dt
    value = steps_total / t_total
    sample = (f"{episodes} of the 2^20 episodes per step ({episodes * STEPS_PER_EPISODE} env steps), "
              f"{cores} host threads, C port of the reference's loop (oracle/qt_oracle.c, -O2, no FMA contraction)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, EPISODES_PER_GPU),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample,
                         "python_reference": python_reference_timing(cores)},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "`value` is the C restatement of the reference's loop (oracle/qt_oracle.c) on all host threads: ~1000x faster per "
                "core than the reference itself, i.e. the stronger baseline.  The unmodified Python reference, shipped in "
                "baseline/_ref, is timed in the same run: cpu_baseline.python_reference.",
    }
    emit(line)



DARE_SIZES = ((6, 65536), (9, 65536), (12, 16384), (15, 16384))   # (n, systems per GPU); m = 4 throughout
DARE_LARGE_BATCH = 1 << 20   # second batch size for n = 6 / 9: launch overhead and the partial last wave amortised


def run_dare_section(ctx, rank: int, world: int, fp64_tflops: float, barrier) -> dict:
    """Batched DARE throughput, every rank solving its own batch (weak scaling, no collective in the data path)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT / "tools"))
    from bench_dare import dare_test_systems

    dev = ctx.device
    f64 = dict(dtype=torch.float64, device=dev)
    out = {"metric": "batched DARE solves/s (validation + structure-preserving doubling + gain), summed over the ranks",
           "unit": "solves/s", "sizes": {}}
    for n, batch in DARE_SIZES:
        rng = np.random.default_rng(42 + rank)
        if n <= 9:    # the reference's own designs on the tuning ranges (tuning_riccati.yaml), built on the device
            q = rng.uniform([5e-5, 5e-5, 10, 1e-3, 1e-3, 2], [5e-4, 5e-4, 25, 1e-2, 1e-2, 8], (batch, 6))
            if n == 9:
                q = np.concatenate([q, rng.uniform([1e-3, 1e-3, 1e-2], [0.1, 0.1, 1.0], (batch, 3))], axis=1)
            r = rng.uniform(0.5, 2.0, (batch, 4))
            A, B = torch.empty((n, n), **f64), torch.empty((n, 4), **f64)
            Q, R = torch.empty((batch, n, n), **f64), torch.empty((batch, 4, 4), **f64)
            ctx.build_lqr_systems(n, batch, 0.01, 1.0, 9.81, torch.as_tensor(q, **f64), torch.as_tensor(r, **f64), A, B, Q, R)
            what = "reference design (diagonal Q/R of the tuning ranges)"
        else:         # the 12 x 4 / 15 x 4 sizes the north star names: integrator chains, dense SPD Q
            A, B, Q, R = (torch.as_tensor(x, **f64) for x in dare_test_systems(n, 4, batch, seed=rank))
            what = "synthetic dense system (integrator chain with random coupling, dense SPD Q, diagonal R)"
        K = torch.empty((batch, 4, n), **f64)
        iters = torch.zeros(batch, dtype=torch.int32, device=dev)
        info = torch.zeros(batch, dtype=torch.int32, device=dev)

        def timed(nb, Qx, Rx, Kx, itx, infx):
            for _ in range(3):
                ctx.dare_batched(n, 4, nb, A, 0, B, 0, Qx, n * n, Rx, 16, None, Kx, itx, infx)
            barrier()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            d0.record()
            for _ in range(reps):
                ctx.dare_batched(n, 4, nb, A, 0, B, 0, Qx, n * n, Rx, 16, None, Kx, itx, infx)
            d1.record()
            barrier()
            return d0.elapsed_time(d1) / reps

        ms = timed(batch, Q, R, K, iters, info)          # the shipped path (QT_DARE_PATH_AUTO)
        ms_dense = ms_big = None
        if n <= 9:
            # the reference's designs take the decoupled path (blocks in registers): also time the dense kernel on the
            # same systems, and the shipped path on a batch large enough to amortise launches and the partial last wave
            K_auto = K.clone()
            ctx.set_dare_path(1)
            try:
                ms_dense = timed(batch, Q, R, K, iters, info)
            finally:
                ctx.set_dare_path(0)
            paths_rel = float(((K - K_auto).abs().amax(dim=(1, 2)) / K_auto.abs().amax(dim=(1, 2))).max().item())
            big = DARE_LARGE_BATCH
            Qb, Rb = Q.repeat(big // batch, 1, 1), R.repeat(big // batch, 1, 1)
            Kb = torch.empty((big, 4, n), **f64)
            ib, fb = torch.zeros(big, dtype=torch.int32, device=dev), torch.zeros(big, dtype=torch.int32, device=dev)
            ms_big = timed(big, Qb, Rb, Kb, ib, fb)
            del Qb, Rb, Kb, ib, fb
            ms = timed(batch, Q, R, K, iters, info)      # leaves K / iters / info of the shipped path for the checks below
        # e2e: pinned host matrices in, host gains / info out, through the C ABI (qt_dare_host)
        Ah, Bh, Qh, Rh = (t.cpu().pin_memory() for t in (A, B, Q, R))
        Kh = torch.empty((batch, 4, n), dtype=torch.float64).pin_memory()
        ih, fh = torch.empty(batch, dtype=torch.int32).pin_memory(), torch.empty(batch, dtype=torch.int32).pin_memory()
        ctx.dare_host(n, 4, batch, Ah, 0, Bh, 0, Qh, n * n, Rh, 16, None, Kh, ih, fh)
        barrier()
        e_reps = 5
        w0 = time.perf_counter()
        for _ in range(e_reps):
            ctx.dare_host(n, 4, batch, Ah, 0, Bh, 0, Qh, n * n, Rh, 16, None, Kh, ih, fh)
        barrier()
        e2e_s = (time.perf_counter() - w0) / e_reps
        same = bool(torch.equal(Kh, K.cpu()))
        t = torch.tensor([ms, e2e_s * 1e3, ms_dense or 0.0, ms_big or 0.0], **f64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, ms_dense, ms_big = (float(x) for x in t.tolist())
        k = float(iters.double().mean().item())
        rate = world * batch / (ms * 1e-3)
        decoupled = n <= 9
        # flops of the path that ran: three (n / 3)-state single-input blocks on the decoupled path, the n x n recursion
        # on the dense one (W_dare of SURVEY 8(d) in both cases)
        flops = 3.0 * w_dare(n // 3, 1, k) if decoupled else w_dare(n, 4, k)
        tflops = rate * flops / 1e12
        bytes_per_solve = 8 * (n * n + 16) + 8 * 4 * n + 8      # Q, R in; K, iters, info out (A and B are shared)
        entry_extra = {}
        if decoupled:
            peak = world * fp64_tflops
            dense_rate = world * batch / (ms_dense * 1e-3)
            big_rate = world * DARE_LARGE_BATCH / (ms_big * 1e-3)
            entry_extra = {
                "path": "decoupled (dare_blocks_kernel: 3 blocks of %d states per system, in registers)" % (n // 3),
                "hbm": {"bytes_per_solve": bytes_per_solve, "achieved_gbs": rate * bytes_per_solve / 1e9},
                "large_batch": {"batch_per_gpu": DARE_LARGE_BATCH, "value": big_rate, "ms_per_launch": ms_big,
                                "fp64_tflops": big_rate * flops / 1e12,
                                "fp64_frac": big_rate * flops / 1e12 / peak if peak > 0 else None,
                                "hbm_gbs": big_rate * bytes_per_solve / 1e9},
                "dense_path": {"value": dense_rate, "ms_per_launch": ms_dense, "kernel": "dare_kernel<N, R> (QT_DARE_PATH_DENSE)",
                               "max_rel_diff_of_K_between_paths": paths_rel,
                               "roofline": {"bound": "fp64", "achieved": dense_rate * w_dare(n, 4, k) / 1e12, "peak": peak,
                                            "unit": "TFLOP/s", "flops_per_solve": w_dare(n, 4, k),
                                            "frac": dense_rate * w_dare(n, 4, k) / 1e12 / peak if peak > 0 else None}}}
        else:
            entry_extra = {"path": "dense (dare_kernel<N, R>)"}
        out["sizes"][f"n{n}"] = {
            "n": n, "m": 4, "systems": what, "batch_per_gpu": batch, "value": rate, "ms_per_launch": ms, **entry_extra,
            "mean_doubling_iterations": k, "failed": int((info != 0).sum().item()),
            "e2e": {"value": world * batch / (e2e_ms * 1e-3), "unit": "solves/s", "api": "qt_dare_host (C ABI, pinned host buffers)",
                    "h2d_bytes_per_step": 8 * (n * n + 4 * n + batch * (n * n + 16)), "d2h_bytes_per_step": batch * (8 * 4 * n + 8),
                    "matches_device_resident_result": same},
            "roofline": {"bound": "fp64", "achieved": tflops, "peak": world * fp64_tflops, "unit": "TFLOP/s",
                         "frac": tflops / (world * fp64_tflops) if fp64_tflops > 0 else None,
                         "flops_per_solve": flops}}
    head = out["sizes"]["n6"]
    out.update({"value": head["value"], "batch": head["batch_per_gpu"], "ms_per_launch": head["ms_per_launch"],
                "mean_doubling_iterations": head["mean_doubling_iterations"], "failed": head["failed"], "e2e": head["e2e"],
                "roofline": head["roofline"], "n_gpus": world})
    return out


def run_strong_section(ctx, rank: int, world: int, barrier, args) -> dict:
    """The literal BASELINE configs at fixed TOTAL size, split over the ranks (strong scaling; the driver's SCALE file
    holds this object at every N)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from lqr_quadcopter_test_b200 import ControllerTuner, GainSearchSpace, RiccatiLQRController, SuccessCriteria, TuningConfig
    from lqr_quadcopter_test_b200 import dist as qd
    from lqr_quadcopter_test_b200.config import EnvConfig

    if args.no_strong:
        return {"skipped": "--no-strong"}
    dev = ctx.device
    f64 = dict(dtype=torch.float64, device=dev)
    out = {"scaling": "strong", "n_gpus": world}

    def max_over_ranks(seconds: float) -> float:
        t = torch.tensor([seconds], **f64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def median_wall(fn, reps=3):
        fn()                                                  # warm-up
        ts = []
        for _ in range(reps):
            barrier()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(max_over_ranks(time.perf_counter() - t0))
        return sorted(ts)[len(ts) // 2]

    # config 2: 2^20 episodes in total, linear target, Riccati-LQR defaults: reset + rollout + summary (+ all_gather)
    total2 = int(args.strong_episodes)
    lo, hi = qd.shard_range(total2, rank, world)
    env = EnvConfig()
    env.target.motion_type = "linear"
    ctrl = RiccatiLQRController(config={})
    prm = ctrl.kernel_params(env, SuccessCriteria())
    K = torch.as_tensor(np.ascontiguousarray(ctrl.K), **f64)
    E = hi - lo
    seeds = torch.arange(lo, hi, dtype=torch.int64, device=dev) + 42
    s0, tp, met, sums = torch.empty((12, E), **f64), torch.empty((12, E), **f64), torch.empty((16, E), **f64), torch.zeros(16, **f64)

    def config2():
        ctx.reset_batched(prm, seeds, s0, tp)
        ctx.rollout(prm, E, STEPS_PER_EPISODE + 2, 1, s0, tp, K, 6, E, met)
        ctx.reduce_summary(met, E, sums)
        if world > 1:
            qd.reduce_summary_across_ranks(sums, lo)
        else:
            sums.cpu()

    sec = median_wall(config2)
    out["config2"] = {"workload": f"{total2} episodes in total, linear target, Riccati-LQR, 3000 steps; reset + rollout + summary "
                                  "reduction (+ all_gather of 17 doubles per rank)", "episodes_per_gpu": E, "seconds": sec,
                      "env_steps_per_s": total2 * STEPS_PER_EPISODE / sec}
    del s0, tp, met

    # config 4: 65 536 candidates x 64 episodes, end to end from the generator seed to host score rows
    C = int(args.strong_candidates)
    space = GainSearchSpace(q_pos_range=([5e-5, 5e-5, 10.0], [5e-4, 5e-4, 25.0]), q_vel_range=([1e-3, 1e-3, 2.0], [1e-2, 1e-2, 8.0]),
                            r_controls_range=([0.5] * 4, [2.0] * 4))

    def config4():
        tuner = ControllerTuner(TuningConfig(controller_type="riccati_lqr", search_space=space, strategy="random", max_iterations=C,
                                             evaluation_episodes=64, evaluation_horizon=3000, seed=42,
                                             output_dir=os.path.join(tempfile.gettempdir(), f"qt_bench_tuning_{rank}")))
        block, rows = tuner.sweep(C)
        config4.best = float(rows[0].max())

    sec = median_wall(config4)
    out["config4"] = {"workload": f"{C} Q/R candidates x 64 episodes x 3000 steps (stationary target): candidate sampling (host rng, "
                                  "the reference's stream) + H2D + DARE + reset + rollout with fused per-candidate scores + D2H "
                                  "(+ all_gather of score and parameter rows)", "seconds": sec, "candidates_per_s": C / sec,
                      "dare_solves_per_s_end_to_end": C / sec, "env_steps_per_s": C * 64 * 3000 / sec, "best_score": config4.best}

    # config 5: teacher data set, 2^22 episodes in total at stride 10 (300 rows of 18 + 4 fp32 per episode = 111 GB in
    # all), produced chunk by chunk into buffers the consumer re-uses ("produced in chunks as torch tensors")
    total5 = int(args.strong_dataset_episodes)
    lo5, hi5 = qd.shard_range(total5, rank, world)
    n_chunks = max(1, -(-(hi5 - lo5) // (1 << 20)))
    chunk = -(-(hi5 - lo5) // n_chunks)
    env5 = EnvConfig()
    env5.target.motion_type = "linear"
    slots = (STEPS_PER_EPISODE + 9) // 10
    buffers: dict = {}

    def buffers_for(n):          # one set per chunk size (the last chunk of an uneven shard is shorter)
        if n not in buffers:
            buffers[n] = (torch.empty((slots, n, 18), dtype=torch.float32, device=dev),
                          torch.empty((slots, n, 4), dtype=torch.float32, device=dev),
                          torch.empty((12, n), **f64), torch.empty((12, n), **f64), torch.empty((16, n), **f64))
        return buffers[n]

    def config5():
        for c0 in range(lo5, hi5, chunk):
            n = min(chunk, hi5 - c0)
            feat, act, s0_, tp_, met_ = buffers_for(n)
            sd = torch.arange(c0, c0 + n, dtype=torch.int64, device=dev) + 42
            ctx.reset_batched(prm5, sd, s0_, tp_)
            ctx.rollout(prm5, n, STEPS_PER_EPISODE, 1, s0_, tp_, K, 6, n, met_, None, None, None, 1, 0, feat, act, 10, slots)

    prm5 = ctrl.kernel_params(env5, SuccessCriteria())
    sec = median_wall(config5, reps=2)
    rows = total5 * slots
    out["config5"] = {"workload": f"{total5} episodes in total, teacher data set every 10th step (18 + 4 fp32 per row), chunks of "
                                  f"{chunk} episodes into re-used device buffers", "seconds": sec, "rows_per_s": rows / sec,
                      "write_gbs": rows * 88 / sec / 1e9, "env_steps_per_s": total5 * STEPS_PER_EPISODE / sec}
    return out


def ctx_sm_count(ctx) -> int:
    import torch

    return torch.cuda.get_device_properties(ctx.device).multi_processor_count


# ---------------------------------------------------------------------------------- GPU arm --
def run_gpu_arm(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    from lqr_quadcopter_test_b200 import RiccatiLQRController, SuccessCriteria, _lib
    from lqr_quadcopter_test_b200 import dist as qd
    from lqr_quadcopter_test_b200.config import EnvConfig

    rank, world, local = qd.init_process_group()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: qtrack has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    ctx = _lib.get_context(local)
    dev = ctx.device
    E = args.episodes
    env = EnvConfig()
    env.target.motion_type = "linear"
    ctrl = RiccatiLQRController(config={})
    prm = ctrl.kernel_params(env, SuccessCriteria())
    K = torch.as_tensor(np.ascontiguousarray(ctrl.K), dtype=torch.float64, device=dev)
    f64 = dict(dtype=torch.float64, device=dev)
    seeds = torch.arange(E, dtype=torch.int64, device=dev) + 42 + rank * E
    state0, tparams = torch.empty((12, E), **f64), torch.empty((12, E), **f64)
    metrics, sums = torch.empty((16, E), **f64), torch.zeros(16, **f64)
    ctx.reset_batched(prm, seeds, state0, tparams)
    max_steps = STEPS_PER_EPISODE + 2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kernel_ms: list[float] = []

    def one_pass(timed: bool):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        ctx.rollout(prm, E, max_steps, 1, state0, tparams, K, 6, E, metrics)
        ev1.record()
        ctx.reduce_summary(metrics, E, sums)
        if world > 1:
            qd.reduce_summary_across_ranks(sums, rank * E)
        if timed:
            kernel_events.append((ev0, ev1))

    kernel_events: list = []
    with ClockSampler(local) as clocks:
        for _ in range(args.warmup):
            one_pass(False)
        barrier()
        launches0 = ctx.launches
        clocks.mark_start()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            one_pass(True)
        t1.record()
        barrier()
        clocks.mark_end()
        elapsed_ms = t0.elapsed_time(t1)
    gpu_launches = ctx.launches - launches0
    clock_info = clocks.summary()
    kernel_ms = [a.elapsed_time(b) for a, b in kernel_events]
    steps_per_pass = float(sums.cpu()[13])  # S_SUM_STEPS of qt_reduce_summary: env steps of the last pass
    total_steps_rank = steps_per_pass * args.steps
    el = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([total_steps_rank], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    elapsed_ms, total_steps = float(el.item()), float(tot.item())
    value = total_steps / (elapsed_ms * 1e-3)

    # ---- e2e: host buffers through the C ABI (H2D seeds, reset, rollout, D2H metrics) -------------
    seeds_host = (torch.arange(E, dtype=torch.int64) + 42 + rank * E).pin_memory()
    K_host = torch.as_tensor(np.ascontiguousarray(ctrl.K), dtype=torch.float64).pin_memory()
    metrics_host = torch.empty((16, E), dtype=torch.float64).pin_memory()
    ctx.evaluate_host(prm, seeds_host, max_steps, 1, K_host, 6, metrics_host)  # warm-up (allocates scratch)
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.evaluate_host(prm, seeds_host, max_steps, 1, K_host, 6, metrics_host)
    barrier()
    e2e_s = time.perf_counter() - w0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = float(metrics_host[0].sum().item()) * e2e_steps * world / float(e2e_t.item())
    same = bool(torch.equal(metrics_host, metrics.cpu()))

    # ---- collective secondary sections (every rank takes part; times are max over ranks) ---------------
    # the FP64 roofline denominator with a clock record of its own: the probe repeated under the clock sampler (the first
    # repetitions give nvidia-smi time to come up under the same load)
    with ClockSampler(local) as probe_clocks:
        for _ in range(20):
            ctx.fp64_peak_probe(1 << 17)
        probe_clocks.mark_start()
        probe_vals = [ctx.fp64_peak_probe(1 << 17)[0] for _ in range(20)]
        probe_clocks.mark_end()
    fp64_tflops = statistics.median(probe_vals)
    probe_clock_info = probe_clocks.summary()
    dare_section = run_dare_section(ctx, rank, world, fp64_tflops, barrier)
    strong = run_strong_section(ctx, rank, world, barrier, args)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rank 0 only: roofline of the headline kernel, policy section, CPU baselines --------------------
    k_ms = statistics.mean(kernel_ms)
    steps_per_launch = steps_per_pass
    Kh = np.asarray(ctrl.K)
    pattern = np.zeros((4, 6), dtype=bool)
    pattern[0, [2, 5]] = pattern[1, [1, 4]] = pattern[2, [0, 3]] = True
    structured = bool(np.all(Kh[~pattern] == 0.0))   # the kernel's own test (gain_is_axis_decoupled); yaw0 = 0 after reset
    flops_step = FLOPS_PER_ENV_STEP_STRUCTURED if structured else FLOPS_PER_ENV_STEP
    achieved = steps_per_launch * flops_step / (k_ms * 1e-3) / 1e12
    dense_equiv = steps_per_launch * FLOPS_PER_ENV_STEP / (k_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    algo_bytes = (24 + 16) * 8.0 * E
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture
        tr = json.load(open(ROOT / "profiles" / "ncu_traffic.json"))["rollout_kernel_linear_lqr_2^20x3000"]
        traffic = tr["dram_bytes_per_launch"] if E == EPISODES_PER_GPU else None
    except (OSError, KeyError, ValueError):
        pass
    roofline = {
        "bound": "fp64", "kernel": "rollout_kernel<linear, LQR, lean>" + (" structured loop" if structured else " general loop"),
        "achieved": achieved, "peak": fp64_tflops,
        "unit": "TFLOP/s", "frac": achieved / fp64_tflops if fp64_tflops > 0 else None, "traffic": traffic,
        "peak_source": "register-resident DFMA probe (qt_fp64_peak_probe) in this run, median of 20; MEASURED_PEAKS.json has no FP64 entry",
        "peak_clocks": probe_clock_info, "peak_spread": [min(probe_vals), max(probe_vals)],
        "algorithmic_flops_per_env_step": flops_step, "dense_flops_per_env_step": FLOPS_PER_ENV_STEP,
        "dense_equivalent_tflops": dense_equiv,
        "note": "achieved counts only the flops the executed path needs (324 of the dense 440: products with the "
                "gain's exact zeros and the identically-zero yaw chain are skipped, bit-identically); "
                "dense_equivalent_tflops = env-steps/s x 440 is given for comparison with earlier rounds",
        "kernel_ms": k_ms,
        "kernel_share_of_step": k_ms * args.steps / elapsed_ms if world == 1 else None,
        "hbm": {"algorithmic_bytes_per_launch": algo_bytes, "achieved_gbs": algo_bytes / (k_ms * 1e-3) / 1e9,
                "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
    }

    dare = dare_section

    # deep policy in the loop (random-init 64-64 network, torch seed 0; linear target; labels every 10th step)
    policy = None
    try:
        from lqr_quadcopter_test_b200 import DeepTrackingPolicy, rollout_device

        torch.manual_seed(0)
        pol = DeepTrackingPolicy(config={"hidden_sizes": [64, 64], "activation": "relu"})
        # whole waves of both policy kernels' grids: 3 resident blocks of 128 episodes (FP32-pipe tiles) and 1 block
        # of 512 (tensor cores) per SM: multiples of 1536 per SM, 3 waves / 9 waves (227 328 episodes on a B200)
        Ep = min(E, ctx_sm_count(ctx) * 1536)
        pseeds = torch.arange(Ep, dtype=torch.int64, device=dev) + 42

        def time_policy():
            for _ in range(2):
                rollout_device(pol, env, SuccessCriteria(), pseeds, supervisor=ctrl, dataset_stride=10, ctx=ctx)
            torch.cuda.synchronize()
            pev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            pev[0].record()
            for i in range(5):
                out = rollout_device(pol, env, SuccessCriteria(), pseeds, supervisor=ctrl, dataset_stride=10, ctx=ctx,
                                     want_final_state=True)
                pev[i + 1].record()
            torch.cuda.synchronize()
            ms = statistics.median(pev[i].elapsed_time(pev[i + 1]) for i in range(5))   # secondary figure: median of five
            return ms, out

        p_ms, pout = time_policy()                      # default: FP32 pipe (FFMA2 tiles), the parity path
        pol.tensor_cores = True
        t_ms, tout = time_policy()                      # hidden layers on the tensor cores (tcgen05, 3xTF32)
        pol.tensor_cores = False
        p_steps = float(pout["metrics"][0].sum().item())
        t_steps = float(tout["metrics"][0].sum().item())
        fma_per_step = 18 * 64 + 64 * 64 + 64 * 4
        fp32_nominal = ctx_sm_count(ctx) * 128 * 2 * (clock_info.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
        fp32_peak, _ = ctx.fp32_peak_probe(1 << 17)        # measured: register-resident FFMA chains, this run
        on_diff = (pout["metrics"][3] - tout["metrics"][3]).abs()
        # the two evaluations on the same 40 steps (the random network's episodes leave the box and diverge chaotically
        # over 30 s, so full-length final states say nothing; tests/test_parity_gpu.py bounds this figure at 1e-5)
        short = {}
        for flag in (False, True):
            pol.tensor_cores = flag
            short[flag] = rollout_device(pol, env, SuccessCriteria(), pseeds[:65536], max_steps=40, ctx=ctx,
                                         want_final_state=True)["final_state"]
        pol.tensor_cores = False
        short_diff = float((short[False] - short[True]).abs().max().item())
        policy = {"metric": "closed-loop env-steps/s with the 64-64 deep tracking policy acting (fp32, in-kernel)",
                  "value": p_steps / (p_ms * 1e-3), "unit": "env-steps/s", "episodes": Ep, "ms_per_launch": p_ms,
                  "path": "FP32 pipe (warp-cooperative FFMA2 tiles; bit-identical to a sequential fp32 evaluation; default)",
                  "includes": "reset + rollout + supervisor labels and 18+4 fp32 data-set rows every 10th step",
                  "roofline": {"bound": "fp32", "achieved": p_steps * fma_per_step * 2 / (p_ms * 1e-3) / 1e12,
                               "peak": fp32_peak, "unit": "TFLOP/s",
                               "frac": p_steps * fma_per_step * 2 / (p_ms * 1e-3) / 1e12 / fp32_peak,
                               "peak_source": "register-resident FFMA probe (qt_fp32_peak_probe) in this run",
                               "nominal_peak": fp32_nominal,
                               "algorithmic_flops_per_env_step": 2 * fma_per_step},
                  "tensor_cores": {
                      "path": "hidden layers as tcgen05.mma kind::tf32 (3xTF32 split, fp32 accumulation in TMEM), output layer "
                              "on the FP32 pipe; DeepTrackingPolicy(config={'tensor_cores': True}) / QT_FLAG_POLICY_TENSOR_CORES",
                      "value": t_steps / (t_ms * 1e-3), "unit": "env-steps/s", "ms_per_launch": t_ms,
                      "speedup_over_fp32_pipe": (t_steps / t_ms) / (p_steps / p_ms),
                      "fp32_equivalent_tflops": t_steps * fma_per_step * 2 / (t_ms * 1e-3) / 1e12,
                      "tf32_mma_tflops_issued": t_steps * 3 * 2 * (24 * 64 + 64 * 64) / (t_ms * 1e-3) / 1e12,
                      "vs_fp32_pipe": {"max_abs_state_diff_after_40_steps": short_diff,
                                       "success_flags_differing": int((pout["metrics"][11] != tout["metrics"][11]).sum().item()),
                                       "on_target_counts_differing": int((on_diff > 0).sum().item()),
                                       "max_on_target_count_diff": int(on_diff.max().item()), "episodes": Ep}}}
    except Exception as exc:  # the headline line must not be lost to the secondary section
        policy = {"error": repr(exc)}

    cores = os.cpu_count() or 1
    sample_eps = sized_cpu_sample(10.0, cores)
    cpu_rate, cpu_s = cpu_port_rate(sample_eps, cores)
    cpu1_rate, _ = cpu_port_rate(max(8, sample_eps // (4 * cores)), 1)
    cpu_baseline = {
        "value": cpu_rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
        "sample": f"{sample_eps} of the 2^20 episodes x 3000 steps ({cpu_s:.1f} s), oracle/qt_oracle.c on {cores} threads",
        "single_core_value": cpu1_rate,
        "python_reference": python_reference_timing(cores) if world == 1 else {"skipped": "timed at N=1 only"},
    }
    if world == 1:
        for n_key, entry in dare["sizes"].items():
            try:
                entry["cpu_baseline"] = scipy_dare_rates(int(n_key[1:]), cores)
            except Exception as exc:
                entry["cpu_baseline"] = {"error": repr(exc)[:200]}
        dare["cpu_baseline"] = dare["sizes"]["n6"].get("cpu_baseline")

    line = {
        "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world, E),
        "clocks": clock_info,
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": 8 * E + 4 * 6 * 8,
                "d2h_bytes_per_step": 16 * 8 * E, "api": "qt_evaluate_host (C ABI, pinned host buffers)",
                "matches_device_resident_result": same, "fraction_of_device_timed_value": e2e_value / value},
        "gpu_launches": int(gpu_launches),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "dare": dare, "strong": strong, "policy": policy,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def emit(line: dict) -> None:
    """The one JSON line, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main() -> None:
    # stdout carries exactly one JSON line: everything else that writes to file descriptor 1 (NCCL prints
    # "NCCL version ..." there at communicator creation) is sent to stderr
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["qtrack", "reference"], default="qtrack")
    ap.add_argument("--episodes", type=int, default=EPISODES_PER_GPU, help="episodes per GPU (default 2^20)")
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed-total-size (strong scaling) section")
    ap.add_argument("--strong-episodes", type=int, default=1 << 20, help="config 2: total episodes over all ranks")
    ap.add_argument("--strong-candidates", type=int, default=65536, help="config 4: total candidates over all ranks")
    ap.add_argument("--strong-dataset-episodes", type=int, default=1 << 22, help="config 5: total episodes over all ranks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
