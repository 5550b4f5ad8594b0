"""DARE-only micro benchmark: python tools/bench_dare.py [batch]  (prints solves/s for n = 6 and 9)."""
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from lqr_quadcopter_test_b200 import _lib

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ctx = _lib.get_context()
dev = ctx.device
f64 = dict(dtype=torch.float64, device=dev)
rng = np.random.default_rng(42)
for n in (6, 9):
    q = rng.uniform([5e-5, 5e-5, 10, 1e-3, 1e-3, 2], [5e-4, 5e-4, 25, 1e-2, 1e-2, 8], (batch, 6))
    if n == 9:
        q = np.concatenate([q, rng.uniform([1e-3, 1e-3, 1e-2], [0.1, 0.1, 1.0], (batch, 3))], axis=1)
    r = rng.uniform(0.5, 2.0, (batch, 4))
    qd, rd = torch.as_tensor(q, **f64), torch.as_tensor(r, **f64)
    A, B = torch.empty((n, n), **f64), torch.empty((n, 4), **f64)
    Q, R = torch.empty((batch, n, n), **f64), torch.empty((batch, 4, 4), **f64)
    ctx.build_lqr_systems(n, batch, 0.01, 1.0, 9.81, qd, rd, A, B, Q, R)
    K = torch.empty((batch, 4, n), **f64)
    iters = torch.zeros(batch, dtype=torch.int32, device=dev)
    info = torch.zeros(batch, dtype=torch.int32, device=dev)
    for _ in range(3):
        ctx.dare_batched(n, 4, batch, A, 0, B, 0, Q, n * n, R, 16, None, K, iters, info)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        ctx.dare_batched(n, 4, batch, A, 0, B, 0, Q, n * n, R, 16, None, K, iters, info)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"n={n} batch={batch}: {ms:.3f} ms/launch  {batch / ms * 1e3:.4e} solves/s  "
          f"iters mean {iters.double().mean().item():.2f} max {iters.max().item()}  failed {(info != 0).sum().item()}")
