"""ctypes binding of libqtrack.so (include/qtrack.h).

There is exactly one compute path: the CUDA library.  If it is missing, was not built for this
machine, or no CUDA device is present, importing succeeds (so CPU-only tooling can inspect the
package) but creating a Context raises — nothing silently falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib
import threading

from .params import QtMlp, QtParams

_PKG = pathlib.Path(__file__).resolve().parent
LIB_PATH = pathlib.Path(os.environ.get("QTRACK_LIB", _PKG / "libqtrack.so"))

# every symbol include/qtrack.h declares (tests check the library exports all of them)
ABI_SYMBOLS = (
    "qt_abi_version", "qt_params_size", "qt_params_default", "qt_create", "qt_destroy",
    "qt_last_error", "qt_set_stream", "qt_get_stream", "qt_synchronize", "qt_launch_count",
    "qt_reset_batched", "qt_rng_probe", "qt_observe_batched", "qt_compute_action_batched",
    "qt_step_batched", "qt_rollout", "qt_dare_batched", "qt_build_lqr_systems",
    "qt_reduce_scores", "qt_reduce_summary", "qt_evaluate_host", "qt_fp64_peak_probe",
    "qt_mlp_packed_floats", "qt_policy_action_batched", "qt_rollout_policy", "qt_sqrt_pair_probe",
    "qt_fill_seeds", "qt_rollout_sweep", "qt_dare_host", "qt_sweep_host", "qt_fp32_peak_probe",
    "qt_set_dare_path",
)

_lib = None
_lib_lock = threading.Lock()


class QtrackError(RuntimeError):
    pass


def load_library() -> C.CDLL:
    """dlopen libqtrack.so and declare the prototypes.  Raises if the library is absent."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise QtrackError(
                f"{LIB_PATH} not found: build it with `python -m lqr_quadcopter_test_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback."
            )
        lib = C.CDLL(str(LIB_PATH))
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        PP = C.POINTER(QtParams)
        protos = {
            "qt_abi_version": (C.c_int, []),
            "qt_params_size": (i64, []),
            "qt_params_default": (None, [PP]),
            "qt_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
            "qt_destroy": (None, [vp]),
            "qt_last_error": (C.c_char_p, [vp]),
            "qt_set_stream": (C.c_int, [vp, vp]),
            "qt_get_stream": (vp, [vp]),
            "qt_synchronize": (C.c_int, [vp]),
            "qt_launch_count": (i64, [vp]),
            "qt_set_dare_path": (C.c_int, [vp, i32]),
            "qt_reset_batched": (C.c_int, [vp, PP, vp, i64, vp, vp, vp]),
            "qt_rng_probe": (C.c_int, [vp, vp, i64, i32, i32, vp]),
            "qt_observe_batched": (C.c_int, [vp, PP, i64, vp, vp, vp, vp]),
            "qt_compute_action_batched": (C.c_int, [vp, PP, i64, vp, vp, i32, i64, vp, vp, vp]),
            "qt_step_batched": (C.c_int, [vp, PP, i64, vp, vp, vp, vp, vp, vp, i32, i64, vp, vp,
                                          vp, vp, vp, vp, vp]),
            "qt_rollout": (C.c_int, [vp, PP, i64, i64, i64, vp, vp, vp, i32, i64, vp, vp, vp, vp,
                                     i64, i64, vp, vp, i64, i64, vp]),
            "qt_dare_batched": (C.c_int, [vp, i32, i32, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp,
                                          vp, vp, vp]),
            "qt_build_lqr_systems": (C.c_int, [vp, i32, i64, dbl, dbl, dbl, vp, vp, vp, vp, vp, vp]),
            "qt_reduce_scores": (C.c_int, [vp, vp, i64, i64, vp]),
            "qt_reduce_summary": (C.c_int, [vp, vp, i64, vp]),
            "qt_evaluate_host": (C.c_int, [vp, PP, vp, i64, i64, i64, vp, i32, vp]),
            "qt_fill_seeds": (C.c_int, [vp, vp, i64, C.c_uint64]),
            "qt_rollout_sweep": (C.c_int, [vp, PP, i64, i64, i64, vp, vp, vp, i32, vp, vp, vp]),
            "qt_dare_host": (C.c_int, [vp, i32, i32, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, vp, vp]),
            "qt_sweep_host": (C.c_int, [vp, PP, i32, i64, i64, C.c_uint64, i64, vp, vp, dbl, dbl, vp, vp, vp]),
            "qt_fp64_peak_probe": (C.c_int, [vp, i64, C.POINTER(dbl), C.POINTER(dbl)]),
            "qt_fp32_peak_probe": (C.c_int, [vp, i64, C.POINTER(dbl), C.POINTER(dbl)]),
            "qt_sqrt_pair_probe": (C.c_int, [vp, i64, vp, vp, vp, vp]),
            "qt_mlp_packed_floats": (i64, [C.POINTER(QtMlp)]),
            "qt_policy_action_batched": (C.c_int, [vp, C.POINTER(QtMlp), i64, vp, vp]),
            "qt_rollout_policy": (C.c_int, [vp, PP, i64, i64, vp, vp, C.POINTER(QtMlp), vp, i32, i64, vp, vp, vp,
                                            i64, i64, vp, vp, i64, i64]),
        }
        for name, (res, args) in protos.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.qt_params_size() != C.sizeof(QtParams):
            raise QtrackError(
                f"qt_params layout mismatch: library {lib.qt_params_size()} bytes, "
                f"python mirror {C.sizeof(QtParams)} bytes"
            )
        _lib = lib
        return lib


def _ptr(t):
    """Device (or pinned-host) address of a torch tensor, or NULL."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError("libqtrack needs contiguous tensors")
    return C.c_void_p(t.data_ptr())


class Context:
    """One libqtrack context = one GPU + one stream.  All methods enqueue on ``stream``."""

    def __init__(self, device: int | None = None, use_torch_stream: bool = True):
        import torch

        self.lib = load_library()
        if not torch.cuda.is_available():
            raise QtrackError("no CUDA device: qtrack has no CPU path (the CPU oracle lives in oracle/ "
                              "and is test infrastructure only)")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        handle = C.c_void_p()
        rc = self.lib.qt_create(self.device_index, C.byref(handle))
        if rc != 0 or not handle.value:
            raise QtrackError(f"qt_create(device={self.device_index}) failed with code {rc}")
        self.handle = handle
        if use_torch_stream:
            self.bind_torch_stream()

    # -- plumbing ---------------------------------------------------------------------------
    def bind_torch_stream(self) -> None:
        import torch

        stream = torch.cuda.current_stream(self.device)
        self.check(self.lib.qt_set_stream(self.handle, C.c_void_p(stream.cuda_stream)))

    def check(self, rc: int) -> None:
        if rc != 0:
            msg = self.lib.qt_last_error(self.handle)
            raise QtrackError(f"libqtrack error {rc}: {msg.decode() if msg else ''}")

    def synchronize(self) -> None:
        self.check(self.lib.qt_synchronize(self.handle))

    @property
    def launches(self) -> int:
        return int(self.lib.qt_launch_count(self.handle))

    def close(self) -> None:
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.qt_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    # -- thin wrappers (argument order follows include/qtrack.h) ------------------------------
    def reset_batched(self, params, seeds, state_out, tparams_out, draws_io=None):
        self.check(self.lib.qt_reset_batched(self.handle, C.byref(params), _ptr(seeds),
                                             seeds.numel(), _ptr(draws_io), _ptr(state_out),
                                             _ptr(tparams_out)))

    def rng_probe(self, seeds, n_uniform, n_normal, out):
        self.check(self.lib.qt_rng_probe(self.handle, _ptr(seeds), seeds.numel(), n_uniform,
                                         n_normal, _ptr(out)))

    def observe_batched(self, params, E, state, time, tparams, obs_out):
        self.check(self.lib.qt_observe_batched(self.handle, C.byref(params), E, _ptr(state),
                                               _ptr(time), _ptr(tparams), _ptr(obs_out)))

    def compute_action_batched(self, params, E, obs, K, k_cols, k_stride, integral, saturation,
                               action_out):
        self.check(self.lib.qt_compute_action_batched(
            self.handle, C.byref(params), E, _ptr(obs), _ptr(K), k_cols, k_stride, _ptr(integral),
            _ptr(saturation), _ptr(action_out)))

    def step_batched(self, params, E, state, time, tparams, counters, action_in, K, k_cols,
                     k_stride, integral, saturation, obs_out, reward_out, done_out, info_out,
                     action_out):
        self.check(self.lib.qt_step_batched(
            self.handle, C.byref(params), E, _ptr(state), _ptr(time), _ptr(tparams),
            _ptr(counters), _ptr(action_in), _ptr(K), k_cols, k_stride, _ptr(integral),
            _ptr(saturation), _ptr(obs_out), _ptr(reward_out), _ptr(done_out), _ptr(info_out),
            _ptr(action_out)))

    def rollout(self, params, E, max_steps, chain, state0, tparams, K, k_cols, k_stride,
                metrics_out, ctrl_state_io=None, final_state_out=None, traj_out=None, traj_stride=1,
                traj_slots=0, feat_out=None, act_out=None, data_stride=1, data_slots=0, ff_gains=None):
        self.check(self.lib.qt_rollout(
            self.handle, C.byref(params), E, max_steps, chain, _ptr(state0), _ptr(tparams),
            _ptr(K), k_cols, k_stride, _ptr(metrics_out), _ptr(ctrl_state_io),
            _ptr(final_state_out), _ptr(traj_out),
            traj_stride, traj_slots, _ptr(feat_out), _ptr(act_out), data_stride, data_slots, _ptr(ff_gains)))

    def fill_seeds(self, seeds_out, seed0):
        self.check(self.lib.qt_fill_seeds(self.handle, _ptr(seeds_out), seeds_out.numel(), int(seed0) & (2 ** 64 - 1)))

    def rollout_sweep(self, params, C_, eps, max_steps, state0, tparams, K, k_cols, cand_out, metrics_out=None,
                      ff_gains=None):
        """C_ candidates x eps episodes sharing the eps initial conditions; cand_out [4, C_] (see include/qtrack.h)."""
        self.check(self.lib.qt_rollout_sweep(self.handle, C.byref(params), C_, eps, max_steps, _ptr(state0),
                                             _ptr(tparams), _ptr(K), k_cols, _ptr(cand_out), _ptr(metrics_out),
                                             _ptr(ff_gains)))

    def dare_host(self, n, m, batch, A, strideA, B, strideB, Q, strideQ, R, strideR, P_out, K_out, iters_out, info_out):
        """Host (ideally pinned) tensors in and out; returns after the outputs are valid."""
        self.check(self.lib.qt_dare_host(self.handle, n, m, batch, _ptr(A), strideA, _ptr(B), strideB, _ptr(Q), strideQ,
                                         _ptr(R), strideR, _ptr(P_out), _ptr(K_out), _ptr(iters_out), _ptr(info_out)))

    def sweep_host(self, params, n, C_, eps, seed0, max_steps, q_host, r_host, ctrl_mass, ctrl_gravity, cand_out_host,
                   K_out_host, info_out_host):
        self.check(self.lib.qt_sweep_host(self.handle, C.byref(params), n, C_, eps, int(seed0) & (2 ** 64 - 1), max_steps,
                                          _ptr(q_host), _ptr(r_host), float(ctrl_mass), float(ctrl_gravity),
                                          _ptr(cand_out_host), _ptr(K_out_host), _ptr(info_out_host)))

    def set_dare_path(self, path: int) -> None:
        """0 = automatic (decoupled blocks in registers where the system allows it), 1 = dense kernel only."""
        self.check(self.lib.qt_set_dare_path(self.handle, int(path)))

    def dare_batched(self, n, m, batch, A, strideA, B, strideB, Q, strideQ, R, strideR, P_out,
                     K_out, iters_out, info_out):
        self.check(self.lib.qt_dare_batched(
            self.handle, n, m, batch, _ptr(A), strideA, _ptr(B), strideB, _ptr(Q), strideQ,
            _ptr(R), strideR, _ptr(P_out), _ptr(K_out), _ptr(iters_out), _ptr(info_out)))

    def build_lqr_systems(self, n, batch, dt, mass, gravity, q_diag, r_diag, A_out, B_out, Q_out,
                          R_out):
        self.check(self.lib.qt_build_lqr_systems(
            self.handle, n, batch, dt, mass, gravity, _ptr(q_diag), _ptr(r_diag), _ptr(A_out),
            _ptr(B_out), _ptr(Q_out), _ptr(R_out)))

    def reduce_scores(self, metrics, C_, eps, cand_out):
        self.check(self.lib.qt_reduce_scores(self.handle, _ptr(metrics), C_, eps, _ptr(cand_out)))

    def reduce_summary(self, metrics, E, sums_out):
        self.check(self.lib.qt_reduce_summary(self.handle, _ptr(metrics), E, _ptr(sums_out)))

    def evaluate_host(self, params, seeds_host, max_steps, chain, K_host, k_cols, metrics_host):
        """Host (ideally pinned) tensors in and out; returns after metrics_host is valid."""
        self.check(self.lib.qt_evaluate_host(
            self.handle, C.byref(params), _ptr(seeds_host), seeds_host.numel(), max_steps, chain,
            _ptr(K_host), k_cols, _ptr(metrics_host)))

    def sqrt_pair_probe(self, a, b, ra_out, rb_out):
        self.check(self.lib.qt_sqrt_pair_probe(self.handle, a.numel(), _ptr(a), _ptr(b), _ptr(ra_out), _ptr(rb_out)))

    def policy_action_batched(self, mlp, E, obs, action_out):
        self.check(self.lib.qt_policy_action_batched(self.handle, C.byref(mlp), E, _ptr(obs), _ptr(action_out)))

    def rollout_policy(self, params, E, max_steps, state0, tparams, mlp, teacher_K, k_cols, k_stride, metrics_out,
                       final_state_out=None, traj_out=None, traj_stride=1, traj_slots=0, feat_out=None,
                       act_out=None, data_stride=1, data_slots=0):
        self.check(self.lib.qt_rollout_policy(
            self.handle, C.byref(params), E, max_steps, _ptr(state0), _ptr(tparams), C.byref(mlp),
            _ptr(teacher_K), k_cols, k_stride, _ptr(metrics_out), _ptr(final_state_out), _ptr(traj_out),
            traj_stride, traj_slots, _ptr(feat_out), _ptr(act_out), data_stride, data_slots))

    def fp32_peak_probe(self, iters: int = 1 << 16) -> tuple[float, float]:
        tf, ms = C.c_double(), C.c_double()
        self.check(self.lib.qt_fp32_peak_probe(self.handle, iters, C.byref(tf), C.byref(ms)))
        return tf.value, ms.value

    def fp64_peak_probe(self, iters: int = 1 << 16) -> tuple[float, float]:
        tf, ms = C.c_double(), C.c_double()
        self.check(self.lib.qt_fp64_peak_probe(self.handle, iters, C.byref(tf), C.byref(ms)))
        return tf.value, ms.value


_contexts: dict[int, Context] = {}


def get_context(device: int | None = None) -> Context:
    """Process-wide context per device (created on first use)."""
    import torch

    idx = torch.cuda.current_device() if (device is None and torch.cuda.is_available()) else (device or 0)
    ctx = _contexts.get(idx)
    if ctx is None:
        ctx = Context(idx)
        _contexts[idx] = ctx
    else:
        ctx.bind_torch_stream()
    return ctx
