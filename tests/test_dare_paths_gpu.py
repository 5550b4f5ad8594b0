"""The two kernels behind qt_dare_batched (riccati_lqr.py:119-184).

QT_DARE_PATH_AUTO solves systems that are block diagonal after a permutation of the states (every design of
riccati_lqr.py:187-316 / 602-700) block by block in registers (csrc/qt_dare_blocks.cuh); QT_DARE_PATH_DENSE sends
everything through the dense doubling kernel.  Both must return the reference's solution: the decoupled path is held
here to the dense path (whose parity with scipy / the reference's solve_dare is what test_golden_gpu, test_parity_gpu
and test_soak_gpu establish — they now run through the decoupled path for the reference's designs, and this file
keeps the dense kernel covered on the same inputs), to scipy directly, and to the DARE residual.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL = 1e-13  # decoupled vs dense on the same system: same fma order, same pivots; only the doubling count may differ


def _both(ctx, A, B, Q, R, want_P=True):
    from lqr_quadcopter_test_b200 import solve_dare_batched

    out = []
    for path in (0, 1):
        ctx.set_dare_path(path)
        try:
            P, K, it, info = solve_dare_batched(A, B, Q, R, ctx=ctx, want_P=want_P)
            torch.cuda.synchronize()
            out.append((None if P is None else P.cpu().numpy(), K.cpu().numpy(), it.cpu().numpy(), info.cpu().numpy()))
        finally:
            ctx.set_dare_path(0)
    return out


def _design(n, dt=0.01, mass=1.0, g=9.81):
    """build_linearized_system / build_augmented_lqi_system of riccati_lqr.py:187-316 (forward Euler)."""
    A = np.eye(n)
    for k in range(3):
        A[k, 3 + k] = dt
    B = np.zeros((n, 4))
    B[5, 0] = dt / mass
    B[4, 1] = -g * dt
    B[3, 2] = g * dt
    if n == 9:
        for k in range(3):
            A[6 + k, k] = dt
    return A, B


def _tuning_weights(rng, n, C):
    q = np.concatenate([rng.uniform(1e-4, 1e-2, (C, 3)), rng.uniform(1e-4, 1e-2, (C, 3))] +
                       ([rng.uniform(1e-5, 1e-2, (C, 3))] if n == 9 else []), axis=1)
    r = rng.uniform(0.1, 5.0, (C, 4))
    Q = np.zeros((C, n, n))
    Q[:, np.arange(n), np.arange(n)] = q
    R = np.zeros((C, 4, 4))
    R[:, np.arange(4), np.arange(4)] = r
    return Q, R


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("n,C", [(6, 4099), (9, 2051)])
def test_decoupled_path_equals_dense_path_on_the_tuning_ranges(gpu_ctx, n, C):
    rng = np.random.default_rng(1000 + n)
    A, B = _design(n)
    Q, R = _tuning_weights(rng, n, C)
    (P0, K0, it0, i0), (P1, K1, it1, i1) = _both(gpu_ctx, A, B, Q, R)
    assert not i0.any() and not i1.any()
    worst_k = max(_rel(K0[i], K1[i]) for i in range(C))
    worst_p = max(_rel(P0[i], P1[i]) for i in range(C))
    print(f"n={n}: decoupled vs dense  K {worst_k:.2e}  P {worst_p:.2e}  bit-identical K: {int((K0 == K1).all(axis=(1, 2)).sum())}/{C}")
    assert worst_k <= REL and worst_p <= REL
    # exact zeros outside the blocks: what the rollout kernel's structure test relies on
    assert ((K0 == 0.0) == (K1 == 0.0)).all() and ((P0 == 0.0) == (P1 == 0.0)).all()
    assert (K0[:, 3, :] == 0.0).all()                      # the yaw control drives nothing
    # per-block convergence: never fewer doublings than the whole-matrix test needs minus one, never more than +2
    assert (it0 >= it1 - 1).all() and (it0 <= it1 + 2).all()


@pytest.mark.parametrize("n", [6, 9])
def test_decoupled_path_matches_scipy(gpu_ctx, n):
    from scipy.linalg import solve_discrete_are

    rng = np.random.default_rng(2000 + n)
    A, B = _design(n)
    C = 96
    Q, R = _tuning_weights(rng, n, C)
    (P0, K0, _, i0), _ = _both(gpu_ctx, A, B, Q, R)
    assert not i0.any()
    for i in range(C):
        Ps = solve_discrete_are(A, B, Q[i], R[i])
        Ks = np.linalg.solve(R[i] + B.T @ Ps @ B, B.T @ Ps @ A)
        assert np.linalg.norm(K0[i] - Ks) / np.linalg.norm(Ks) <= 1e-9      # north_star: K within 1e-9 relative
        assert np.linalg.norm(P0[i] - Ps) / np.linalg.norm(Ps) <= 1e-9


def test_mixed_batch_routes_every_system_to_the_right_kernel(gpu_ctx):
    """One batch with decoupled systems, a dense Q, a coupled A, a non-diagonal R, rejected weights and NaN: info and
    results must not depend on which kernel took a system; systems the dense kernel takes are bit-identical."""
    rng = np.random.default_rng(7)
    n, C = 6, 77
    A0, B0 = _design(n)
    A = np.repeat(A0[None], C, 0)
    B = np.repeat(B0[None], C, 0)
    Q, R = _tuning_weights(rng, n, C)
    M = rng.standard_normal((n, n))
    Q[3] = M @ M.T / n + 0.1 * np.eye(n)          # dense Q
    A[5, 0, 4] = 0.003                            # x couples with vy
    R[8, 0, 1] = R[8, 1, 0] = 0.05                # R not diagonal
    Q[11, 2, 2] = -1.0                            # rejected: Q not PSD
    R[12, 1, 1] = 0.0                             # rejected: R not PD
    Q[13, 0, 0] = np.nan                          # rejected
    R[14, 3, 3] = np.inf                          # rejected
    Q[15, 1, 1] = -1e-11                          # accepted (tolerance -1e-10)
    Q[16, 0, 1] = 1e-9                            # not symmetric within allclose? |1e-9 - 0| <= 1e-8: symmetric enough, dense
    A[20] = np.diag([0.5, 0.6, 0.7, 0.8, 0.9, 0.3])   # six 1 x 1 blocks, three of them driven
    (P0, K0, it0, i0), (P1, K1, it1, i1) = _both(gpu_ctx, A, B, Q, R)
    assert (i0 == i1).all(), (i0, i1)
    assert i0[[11, 12, 13, 14]].tolist() == [3, 3, 3, 3] and i0[15] != 3   # -1e-11 passes validation (the solve may still fail)
    dense_taken = [3, 5, 8, 16]
    for i in dense_taken:
        assert i0[i] == 0 and (K0[i] == K1[i]).all() and (P0[i] == P1[i]).all() and it0[i] == it1[i]
    ok = i0 == 0
    assert max(_rel(K0[i], K1[i]) for i in np.nonzero(ok)[0]) <= REL
    assert max(_rel(P0[i], P1[i]) for i in np.nonzero(ok)[0]) <= REL


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_block_structures_against_the_dense_path(gpu_ctx, seed):
    """Random partitions of up to nine states into blocks of 1-3 states under a random permutation, one control per
    driven block, undriven blocks stable; different structures inside one warp (per-system A and B)."""
    rng = np.random.default_rng(100 + seed)
    n, m, C = [7, 9, 8][seed], [3, 4, 8][seed], 75
    A = np.zeros((C, n, n))
    B = np.zeros((C, n, m))
    Q = np.zeros((C, n, n))
    R = np.zeros((C, m, m))
    for s in range(C):
        perm = rng.permutation(n)
        i, c = 0, 0
        while i < n:
            sz = min(int(rng.integers(1, 4)), n - i)
            blk = perm[i:i + sz]
            M = rng.standard_normal((sz, sz))
            driven = c < m and rng.random() < 0.7
            rad = np.abs(np.linalg.eigvals(M)).max()
            M *= (1.05 if driven else 0.8) / rad          # a driven block may be unstable, an undriven one may not
            A[s][np.ix_(blk, blk)] = M
            if driven:
                B[s, blk, c] = rng.standard_normal(sz)
                c += 1
            i += sz
        Q[s][np.arange(n), np.arange(n)] = rng.uniform(0.01, 2.0, n)
        R[s][np.arange(m), np.arange(m)] = rng.uniform(0.1, 3.0, m)
    (P0, K0, _, i0), (P1, K1, _, i1) = _both(gpu_ctx, A, B, Q, R)
    assert (i0 == i1).all()
    ok = np.nonzero(i0 == 0)[0]
    assert len(ok) >= C * 0.9
    worst = 0.0
    for s in ok:
        worst = max(worst, _rel(K0[s], K1[s]) if np.abs(K1[s]).max() > 0 else float(np.abs(K0[s]).max()), _rel(P0[s], P1[s]))
        G = np.linalg.solve(R[s] + B[s].T @ P0[s] @ B[s], B[s].T @ P0[s] @ A[s])
        res = A[s].T @ P0[s] @ A[s] - P0[s] - A[s].T @ P0[s] @ B[s] @ G + Q[s]
        assert np.linalg.norm(res) / np.linalg.norm(P0[s]) < 1e-10
    print(f"random block structures n={n} m={m}: decoupled vs dense {worst:.2e}")
    assert worst <= 1e-11          # unstable driven blocks: conditioning, not the path


def test_undriven_integrator_is_reported_by_both_paths(gpu_ctx):
    """An integrator nobody drives has no stabilising solution: neither path may report success."""
    A = np.array([[1.0, 0.0], [0.0, 0.5]])
    B = np.array([[0.0], [1.0]])
    (_, _, _, i0), (_, _, _, i1) = _both(gpu_ctx, A, B, np.eye(2)[None], np.eye(1)[None])
    assert i0[0] != 0 and i1[0] != 0


@pytest.mark.parametrize("batch", [1, 31, 32, 33, 1000])
def test_batch_edges_and_optional_outputs(gpu_ctx, batch):
    rng = np.random.default_rng(batch)
    A, B = _design(9)
    Q, R = _tuning_weights(rng, 9, batch)
    (P0, K0, it0, i0), (P1, K1, it1, i1) = _both(gpu_ctx, A, B, Q, R)
    (Pn, Kn, _, _), _ = _both(gpu_ctx, A, B, Q, R, want_P=False)
    assert Pn is None and (Kn == K0).all()
    assert not i0.any() and max(_rel(K0[i], K1[i]) for i in range(batch)) <= REL


def test_failure_behaviour_follows_scipy(gpu_ctx):
    """Where scipy.linalg.solve_discrete_are (the reference's solver, riccati_lqr.py:165) raises, info != 0 — an
    undriven mode on or outside the unit circle; where it returns a solution that is not stabilising without raising
    — an unobservable mode ON the unit circle — both paths return that same solution with info = 0 (no extra
    stability test: the drop-in must not reject what the reference accepts)."""
    from scipy.linalg import solve_discrete_are

    A, B, Q, R = np.diag([1.0, 0.5]), np.array([[1.0], [1.0]]), np.diag([0.0, 1.0]), np.eye(1)
    Ps = solve_discrete_are(A, B, Q, R)
    for (P, K, _, info) in _both(gpu_ctx, A, B, Q[None], R[None]):
        assert info[0] == 0
        np.testing.assert_allclose(P[0], Ps, rtol=0, atol=1e-12)
    for lam in (1.0, 1.2):
        A2, B2 = np.diag([lam, 0.5]), np.array([[0.0], [1.0]])
        with pytest.raises(np.linalg.LinAlgError):
            solve_discrete_are(A2, B2, np.eye(2), np.eye(1))
        for (_, _, _, info) in _both(gpu_ctx, A2, B2, np.eye(2)[None], np.eye(1)[None]):
            assert info[0] != 0


def test_validation_thresholds_follow_numpy_on_dense_weights(gpu_ctx):
    """Q / R validation (riccati_lqr.py:57-116) for DENSE symmetric weights whose smallest eigenvalue sits on either
    side of the reference's thresholds (Q >= -1e-10, R > 1e-10): the kernel's LDL' certificate with its eigenvalue
    fallback must take numpy's decision in every case."""
    rng = np.random.default_rng(5)
    n, m = 6, 4
    A, B = _design(n)
    q_mins = [-3e-10, -1.3e-10, -1.1e-10, -0.9e-10, -0.7e-10, -0.3e-10, 0.0, 1e-11, 1e-6, 0.5]
    r_mins = [-1e-3, 0.0, 0.5e-10, 0.9e-10, 1.1e-10, 1.3e-10, 1.7e-10, 3e-10, 1e-6, 0.5]
    Qs, Rs, expect = [], [], []
    for qm in q_mins:
        for rm in r_mins:
            Vq, _ = np.linalg.qr(rng.standard_normal((n, n)))
            Vr, _ = np.linalg.qr(rng.standard_normal((m, m)))
            Q = (Vq * np.concatenate([[qm], rng.uniform(0.1, 2.0, n - 1)])) @ Vq.T
            R = (Vr * np.concatenate([[rm], rng.uniform(0.5, 2.0, m - 1)])) @ Vr.T
            Q, R = (Q + Q.T) / 2, (R + R.T) / 2
            ok = (not np.any(np.linalg.eigvalsh(Q) < -1e-10)) and (not np.any(np.linalg.eigvalsh(R) <= 1e-10))
            Qs.append(Q), Rs.append(R), expect.append(ok)
    (_, _, _, i0), (_, _, _, i1) = _both(gpu_ctx, A, B, np.array(Qs), np.array(Rs))
    got = i0 != 3
    assert (i0 == 3).tolist() == (i1 == 3).tolist()
    wrong = [(q_mins[k // len(r_mins)], r_mins[k % len(r_mins)]) for k in range(len(expect)) if got[k] != expect[k]]
    assert not wrong, f"validation differs from numpy for (lambda_min(Q), lambda_min(R)) = {wrong}"
    assert sum(expect) >= 30 and sum(expect) <= 70
