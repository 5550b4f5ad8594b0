"""DARE-only micro benchmark (device-resident matrices, CUDA events):

    python tools/bench_dare.py [batch] [brief|full] [sizes, e.g. 6,9] [auto|dense]

auto (default): qt_dare_batched as shipped — the reference's designs take the decoupled path (blocks in registers);
dense: QT_DARE_PATH_DENSE, every system through the dense doubling kernel.

n = 6 and n = 9: the reference's LQR / LQI designs on the tuning ranges (qt_build_lqr_systems);
n = 12 and n = 15 (m = 4): the sizes the north star names — chains of integrators with random coupling, dense SPD Q,
diagonal R (dare_test_systems below, also used by bench.py and the tests)."""
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from lqr_quadcopter_test_b200 import _lib


def dare_test_systems(n, m, batch, seed=0):
    """Random stabilisable systems of size n x m: A = I + strictly upper triangular coupling (eigenvalues on the unit
    circle, like the reference's integrator chains), B dense, Q = M M' / 20 + diag, R diagonal."""
    rng = np.random.default_rng(seed + 1000 * n)
    A = np.eye(n) + 0.01 * np.triu(rng.standard_normal((n, n)), 1)
    B = 0.05 * rng.standard_normal((n, m))
    Mx = rng.standard_normal((batch, n, n))
    Q = Mx @ Mx.transpose(0, 2, 1) * 0.05 + np.eye(n) * rng.uniform(0.1, 2.0, (batch, 1, 1))
    R = np.stack([np.diag(x) for x in rng.uniform(0.5, 2.0, (batch, m))])
    return A, B, Q, R


def w_dare(n, m, k):
    """canonical flop count of one SDA solve with k doubling iterations."""
    return k * (50.0 / 3.0) * n ** 3 + (4 * n * n * m + 2 * n * m * m + (2.0 / 3.0) * m ** 3 + 2 * m * m * n)


if __name__ == "__main__":
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    brief = len(sys.argv) > 2 and sys.argv[2] == "brief"
    sizes = tuple(int(x) for x in sys.argv[3].split(",")) if len(sys.argv) > 3 else (6, 9, 12, 15)
    ctx = _lib.get_context()
    path = sys.argv[4] if len(sys.argv) > 4 else "auto"
    ctx.set_dare_path(1 if path == "dense" else 0)
    dev = ctx.device
    f64 = dict(dtype=torch.float64, device=dev)
    rng = np.random.default_rng(42)
    for n in sizes:
        nb = batch if n <= 9 else batch // 4
        if n <= 9:
            q = rng.uniform([5e-5, 5e-5, 10, 1e-3, 1e-3, 2], [5e-4, 5e-4, 25, 1e-2, 1e-2, 8], (nb, 6))
            if n == 9:
                q = np.concatenate([q, rng.uniform([1e-3, 1e-3, 1e-2], [0.1, 0.1, 1.0], (nb, 3))], axis=1)
            r = rng.uniform(0.5, 2.0, (nb, 4))
            A, B = torch.empty((n, n), **f64), torch.empty((n, 4), **f64)
            Q, R = torch.empty((nb, n, n), **f64), torch.empty((nb, 4, 4), **f64)
            ctx.build_lqr_systems(n, nb, 0.01, 1.0, 9.81, torch.as_tensor(q, **f64), torch.as_tensor(r, **f64), A, B, Q, R)
        else:
            A, B, Q, R = (torch.as_tensor(x, **f64) for x in dare_test_systems(n, 4, nb))
        K = torch.empty((nb, 4, n), **f64)
        iters = torch.zeros(nb, dtype=torch.int32, device=dev)
        info = torch.zeros(nb, dtype=torch.int32, device=dev)
        for _ in range(3):
            ctx.dare_batched(n, 4, nb, A, 0, B, 0, Q, n * n, R, 16, None, K, iters, info)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            ctx.dare_batched(n, 4, nb, A, 0, B, 0, Q, n * n, R, 16, None, K, iters, info)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        k = iters.double().mean().item()
        if brief:
            print(f"n={n} [{path}]: {nb / ms * 1e3:.3e}/s", flush=True)
        else:
            print(f"n={n} [{path}] batch={nb}: {ms:.4f} ms/launch  {nb / ms * 1e3:.4e} solves/s  {nb / ms * 1e3 * w_dare(n, 4, k) / 1e12:.2f} TFLOP/s  "
                  f"iters mean {k:.2f} max {iters.max().item()}  failed {(info != 0).sum().item()}", flush=True)
