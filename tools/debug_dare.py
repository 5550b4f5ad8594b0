"""Debug helper: which systems of a batch does the DARE kernel get wrong?  python tools/debug_dare.py [batch]"""
import pathlib, sys
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
import oracle
from lqr_quadcopter_test_b200 import solve_dare_batched

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(99)
for lqi in (False, True):
    n = 9 if lqi else 6
    A, B = oracle.build_systems_reference(lqi=lqi)
    q = rng.uniform([5e-5, 5e-5, 10.0, 1e-3, 1e-3, 2.0], [5e-4, 5e-4, 25.0, 1e-2, 1e-2, 8.0], (batch, 6))
    if lqi:
        q = np.concatenate([q, rng.uniform([1e-3, 1e-3, 1e-2], [0.1, 0.1, 1.0], (batch, 3))], axis=1)
    r = rng.uniform(0.5, 2.0, (batch, 4))
    Q = np.stack([np.diag(x) for x in q]); R = np.stack([np.diag(x) for x in r])
    for rep in range(2):
        P_, K, iters, info = solve_dare_batched(A, B, Q, R)
        K, iters, info = K.cpu().numpy(), iters.cpu().numpy(), info.cpu().numpy()
        bad = []
        for i in range(batch):
            _, Kr = oracle.solve_dare_reference(A, B, Q[i], R[i])
            rel = np.linalg.norm(K[i] - Kr) / np.linalg.norm(Kr)
            if not rel < 1e-9:
                bad.append((i, float(rel), int(iters[i]), int(info[i])))
        print(f"n={n} rep {rep}: {len(bad)} bad of {batch}: {bad[:24]}", flush=True)
