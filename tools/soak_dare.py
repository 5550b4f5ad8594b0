"""Soak: the batched DARE kernel against scipy (the reference's solve_dare path) on random general systems.

    python tools/soak_dare.py [systems_per_shape] [seed]

Shapes n = 2..16, m = 1..8; A = random normal scaled to a spectral radius drawn from [0.3, 1.6] (stable and
unstable plants), B random normal, Q = M M' * s + diag, R = N N' * s + diag (dense symmetric positive definite)."""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import scipy.linalg

from lqr_quadcopter_test_b200 import solve_dare_batched

per = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst, fails, total = 0.0, 0, 0
for n in range(2, 17):
    for m in sorted({1, min(2, n), min(4, n), min(8, n)}):
        A = rng.standard_normal((per, n, n))
        rho = np.abs(np.linalg.eigvals(A)).max(axis=1)
        A *= (rng.uniform(0.3, 1.6, per) / rho)[:, None, None]
        B = rng.standard_normal((per, n, m))
        M = rng.standard_normal((per, n, n))
        Q = M @ M.transpose(0, 2, 1) * 0.1 + np.eye(n) * rng.uniform(0.01, 2.0, (per, 1, 1))
        N = rng.standard_normal((per, m, m))
        R = N @ N.transpose(0, 2, 1) * 0.1 + np.eye(m) * rng.uniform(0.1, 2.0, (per, 1, 1))
        P_, K, iters, info = solve_dare_batched(A, B, Q, R)
        P_, K, info = P_.cpu().numpy(), K.cpu().numpy(), info.cpu().numpy()
        for i in range(per):
            total += 1
            try:
                Pr = scipy.linalg.solve_discrete_are(A[i], B[i], Q[i], R[i])
            except Exception:
                continue                      # scipy itself gave up: nothing to compare
            Kr = np.linalg.solve(R[i] + B[i].T @ Pr @ B[i], B[i].T @ Pr @ A[i])
            if info[i] != 0:
                fails += 1
                print(f"n={n} m={m} system {i}: info={info[i]} (scipy solved it; cond(P)={np.linalg.cond(Pr):.1e})")
                continue
            rel = np.linalg.norm(K[i] - Kr) / np.linalg.norm(Kr)
            worst = max(worst, rel)
            if rel > 1e-9:
                def resid(P):
                    G = np.linalg.solve(R[i] + B[i].T @ P @ B[i], B[i].T @ P @ A[i])
                    return np.linalg.norm(A[i].T @ P @ A[i] - P - A[i].T @ P @ B[i] @ G + Q[i]) / np.linalg.norm(P)
                print(f"n={n} m={m} system {i}: rel K error {rel:.2e}, iters {int(iters[i])}, cond(P)={np.linalg.cond(Pr):.1e}, "
                      f"Riccati residual kernel {resid(P_[i]):.1e} / scipy {resid(Pr):.1e}")
print(f"{total} systems, {fails} not solved by the kernel, worst relative gain error {worst:.2e}")
