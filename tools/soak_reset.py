"""Soak: the reset kernel's SeedSequence -> PCG64 -> uniform / ziggurat-normal chain against numpy's own
generator (oracle.reset_reference), bit for bit, for many seeds of every motion type.

    python tools/soak_reset.py [seeds_per_motion]"""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np

import gpu_backend as gb
import oracle
from lqr_quadcopter_test_b200 import params as P
from lqr_quadcopter_test_b200.config import EnvConfig

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(123)
seeds = np.concatenate([np.arange(n // 2), rng.integers(0, 2 ** 63 - 1, n - n // 2, dtype=np.int64)])
for motion in ("stationary", "linear", "circular", "sinusoidal", "figure8"):
    env = EnvConfig()
    env.target.motion_type = motion
    prm = P.make_params(env)
    s_ref, t_ref = oracle.reset_reference(prm, seeds)
    s_gpu, t_gpu = gb.reset(prm, seeds)
    bad_t = int((t_ref != t_gpu).any(axis=0).sum())
    ds = np.abs(s_ref - s_gpu).max()
    print(f"{motion}: {len(seeds)} seeds, target parameters differing in {bad_t} episodes, "
          f"max |state0 difference| = {ds:.1e}", flush=True)
