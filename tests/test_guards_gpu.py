"""GPU: out-of-bounds write detection with guard bands (compute-sanitizer is closed on this pool).

Every output buffer is carved out of a larger allocation whose surroundings are filled with a
sentinel; after the call the sentinels must be untouched.  Sizes are deliberately awkward (not
multiples of the block / warp / team sizes) so tail handling is exercised."""
import numpy as np
import pytest
import torch

from lqr_quadcopter_test_b200 import params as P
from lqr_quadcopter_test_b200.config import EnvConfig

import golden_cases as gc

pytestmark = pytest.mark.gpu
GUARD = 64


class Guarded:
    def __init__(self, device):
        self.device = device
        self.items = []

    def alloc(self, shape, dtype=torch.float64):
        n = int(np.prod(shape))
        sentinel = {torch.float64: -7.25e77, torch.float32: -7.25e33, torch.int32: -77777, torch.int64: -7777777}[dtype]
        buf = torch.full((n + 2 * GUARD,), sentinel, dtype=dtype, device=self.device)
        view = buf[GUARD:GUARD + n].view(*shape)
        view.zero_()
        self.items.append((buf, n, sentinel))
        return view

    def check(self):
        torch.cuda.synchronize()
        for buf, n, sentinel in self.items:
            assert bool((buf[:GUARD] == sentinel).all()) and bool((buf[GUARD + n:] == sentinel).all())


@pytest.mark.parametrize("E,motion,lqi", [(1, "linear", False), (127, "sinusoidal", True), (129, "circular", False),
                                          (1000, "figure8", True), (4097, "stationary", False)])
def test_rollout_step_reset_reduce_stay_in_bounds(gpu_ctx, golden, E, motion, lqi):
    c = gpu_ctx
    g = Guarded(c.device)
    env = EnvConfig()
    env.target.motion_type = motion
    prm = P.make_params(env, gc.controller_ns(use_lqi=lqi, feedforward_enabled=(E % 2 == 1),
                                              ff_velocity_gain=[0.1] * 3, ff_acceleration_gain=[0.05] * 3))
    K = torch.as_tensor(golden["dare"]["lqi_default_K" if lqi else "lqr_default_K"], dtype=torch.float64, device=c.device)
    seeds = torch.arange(E, dtype=torch.int64, device=c.device) + 5
    state0, tparams, draws = g.alloc((12, E)), g.alloc((12, E)), g.alloc((2, E), torch.int64)
    c.reset_batched(prm, seeds, state0, tparams, draws)
    T, ts, ds = 333, 7, 10
    slots_t, slots_d = (T + ts - 1) // ts, (T + ds - 1) // ds
    metrics, final, cs = g.alloc((16, E)), g.alloc((12, E)), g.alloc((4, E))
    traj = g.alloc((slots_t, 24, E))
    feat, act = g.alloc((slots_d, E, 18), torch.float32), g.alloc((slots_d, E, 4), torch.float32)
    c.rollout(prm, E, T, 1, state0, tparams, K, 9 if lqi else 6, E, metrics, cs, final, traj, ts, slots_t, feat, act, ds, slots_d)
    # chains of 3 episodes, per-episode gains read from global memory
    Kg = K.unsqueeze(0).repeat(E, 1, 1).contiguous()
    n_chains = (E + 2) // 3
    cs3, m3 = g.alloc((4, n_chains)), g.alloc((16, E))
    c.rollout(prm, E, 50, 3, state0, tparams, Kg, 9 if lqi else 6, 1, m3, cs3)
    # single steps, observation, control law
    time, counters = g.alloc((E,)), g.alloc((3, E), torch.int64)
    obs, rew, done, info, aout = g.alloc((22, E)), g.alloc((E,)), g.alloc((E,), torch.int32), g.alloc((4, E)), g.alloc((4, E))
    integ, sat = g.alloc((3, E)), g.alloc((E,), torch.int64)
    st = g.alloc((12, E))
    st.copy_(state0)
    c.observe_batched(prm, E, st, time, tparams, obs)
    c.compute_action_batched(prm, E, obs, K, 9 if lqi else 6, E, integ if lqi else None, sat, aout)
    c.step_batched(prm, E, st, time, tparams, counters, None, K, 9 if lqi else 6, E, integ if lqi else None, sat, obs, rew,
                   done, info, aout)
    c.step_batched(prm, E, st, time, tparams, counters, aout, None, 6, E, None, None, obs, rew, done, info, None)
    sums = g.alloc((16,))
    c.reduce_summary(metrics, E, sums)
    if E % 3 == 1:
        cand = g.alloc((4, E // 3 if E >= 3 else 1))
        if E >= 3:
            c.reduce_scores(metrics[:, : (E // 3) * 3].contiguous(), E // 3, 3, cand)
    g.check()
    assert float(sums[0]) == E and bool((metrics[0] == T).all())


@pytest.mark.parametrize("n,m,batch", [(6, 4, 1), (6, 4, 21), (9, 4, 13), (2, 1, 9), (5, 3, 7), (12, 4, 5), (16, 8, 3), (7, 2, 33)])
def test_dare_stays_in_bounds(gpu_ctx, n, m, batch):
    c = gpu_ctx
    g = Guarded(c.device)
    rng = np.random.default_rng(n * 100 + m)
    A = rng.standard_normal((batch, n, n))
    A *= (1.02 / np.abs(np.linalg.eigvals(A)).max(axis=1))[:, None, None]
    B = rng.standard_normal((batch, n, m))
    M = rng.standard_normal((batch, n, n))
    Q = M @ np.transpose(M, (0, 2, 1)) / n + 0.1 * np.eye(n)
    N_ = rng.standard_normal((batch, m, m))
    R = N_ @ np.transpose(N_, (0, 2, 1)) / m + np.eye(m)
    dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=c.device)
    Pm, K = g.alloc((batch, n, n)), g.alloc((batch, m, n))
    iters, info = g.alloc((batch,), torch.int32), g.alloc((batch,), torch.int32)
    c.dare_batched(n, m, batch, dev(A), n * n, dev(B), n * m, dev(Q), n * n, dev(R), m * m, Pm, K, iters, info)
    g.check()
    assert not bool(info.any())
    # residual of the DARE for the device solution
    X = Pm.cpu().numpy()
    for i in range(batch):
        G = np.linalg.solve(R[i] + B[i].T @ X[i] @ B[i], B[i].T @ X[i] @ A[i])
        res = A[i].T @ X[i] @ A[i] - X[i] - A[i].T @ X[i] @ B[i] @ G + Q[i]
        assert np.linalg.norm(res) / np.linalg.norm(X[i]) < 1e-10
        np.testing.assert_allclose(K.cpu().numpy()[i], G, rtol=1e-8, atol=1e-10)


def test_paired_square_root_is_ieee_sqrt(gpu_ctx):
    """qt_step.cuh sqrt_pair: bit-identical to correctly rounded sqrt, incl. the arguments that take the
    library fallback (0, subnormal, huge, inf, NaN, negative) paired with ordinary ones."""
    import torch

    rng = np.random.default_rng(3)
    a = np.concatenate([rng.uniform(0, 4, 200000), 10.0 ** rng.uniform(-300, 300, 200000), rng.uniform(0, 1e-3, 100000) ** 4])
    b = rng.permutation(a)
    special = np.array([0.0, -0.0, 5e-324, 1e-310, 2.2250738585072014e-308, 1.7976931348623157e308, np.inf, np.nan, -1.0, 1.0])
    a = np.concatenate([a, special, np.ones(10)])
    b = np.concatenate([b, np.full(10, 2.0), special])
    f64 = dict(dtype=torch.float64, device=gpu_ctx.device)
    ta, tb = torch.as_tensor(a, **f64), torch.as_tensor(b, **f64)
    ra, rb = torch.empty_like(ta), torch.empty_like(tb)
    gpu_ctx.sqrt_pair_probe(ta, tb, ra, rb)
    with np.errstate(invalid="ignore"):
        np.testing.assert_array_equal(ra.cpu().numpy(), np.sqrt(a))
        np.testing.assert_array_equal(rb.cpu().numpy(), np.sqrt(b))


@pytest.mark.parametrize("n,batch,want_P", [(6, 1, True), (6, 33, False), (9, 31, True), (9, 65, False), (6, 1000, True)])
def test_decoupled_dare_path_stays_in_bounds(gpu_ctx, n, batch, want_P):
    """The decoupled DARE path (dare_blocks_kernel) writes its per-warp output tile with coalesced stores over whole
    groups of 32 systems: the last, partial group must stop at the batch's end; P is optional."""
    c = gpu_ctx
    g = Guarded(c.device)
    rng = np.random.default_rng(n + batch)
    A = np.eye(n)
    B = np.zeros((n, 4))
    for k in range(3):
        A[k, 3 + k] = 0.01
        if n == 9:
            A[6 + k, k] = 0.01
    B[5, 0], B[4, 1], B[3, 2] = 0.01, -0.0981, 0.0981
    Q = np.zeros((batch, n, n))
    Q[:, np.arange(n), np.arange(n)] = rng.uniform(1e-4, 1e-2, (batch, n))
    R = np.zeros((batch, 4, 4))
    R[:, np.arange(4), np.arange(4)] = rng.uniform(0.1, 5.0, (batch, 4))
    dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=c.device)
    Pm = g.alloc((batch, n, n)) if want_P else None
    K = g.alloc((batch, 4, n))
    iters, info = g.alloc((batch,), torch.int32), g.alloc((batch,), torch.int32)
    c.dare_batched(n, 4, batch, dev(A), 0, dev(B), 0, dev(Q), n * n, dev(R), 16, Pm, K, iters, info)
    g.check()
    assert not bool(info.any()) and bool((iters > 5).all())
    assert bool((K[:, 3, :] == 0).all()) and bool(K[:, 0, 5].abs().min() > 0)


@pytest.mark.parametrize("E,hidden", [(1, [64, 64]), (513, [32, 32]), (1000, [64]), (129, [20, 40, 60])])
def test_tensor_core_policy_rollout_stays_in_bounds(gpu_ctx, golden, E, hidden):
    """qt_rollout_policy with QT_FLAG_POLICY_TENSOR_CORES: blocks of 512 episodes in four sets; partial last block and
    partial last set, all logs on."""
    from lqr_quadcopter_test_b200 import DeepTrackingPolicy

    c = gpu_ctx
    g = Guarded(c.device)
    torch.manual_seed(E)
    pol = DeepTrackingPolicy(config={"hidden_sizes": hidden})
    env = EnvConfig()
    env.target.motion_type = "linear"
    prm = P.make_params(env, gc.controller_ns())
    prm.debug_flags = P.FLAG_POLICY_TENSOR_CORES
    K = torch.as_tensor(golden["dare"]["lqr_default_K"], dtype=torch.float64, device=c.device)
    seeds = torch.arange(E, dtype=torch.int64, device=c.device) + 11
    state0, tparams = g.alloc((12, E)), g.alloc((12, E))
    c.reset_batched(prm, seeds, state0, tparams)
    T, ts, ds = 60, 7, 9
    tslots, dslots = (T + ts - 1) // ts, (T + ds - 1) // ds
    metrics, final = g.alloc((16, E)), g.alloc((12, E))
    traj = g.alloc((tslots, P.TRAJ_FIELDS, E))
    feats, acts = g.alloc((dslots, E, 18), torch.float32), g.alloc((dslots, E, 4), torch.float32)
    c.rollout_policy(prm, E, T, state0, tparams, pol.packed_mlp(), K, 6, E, metrics, final, traj, ts, tslots, feats, acts, ds, dslots)
    g.check()
    assert bool((metrics[0] == T).all()) and bool(torch.isfinite(final).all())
