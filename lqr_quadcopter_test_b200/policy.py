"""Deep tracking policy: the reference's ``PolicyNetwork`` / ``DeepTrackingPolicy``
(controllers/deep_tracking_policy.py:24-415) with inference inside the CUDA closed loop.

The network object is a ``torch.nn.Module`` with the reference's parameter names, so checkpoints
written by either side load in the other (``model_state_dict`` keys ``feature_extractor.{0,2,..}`` and
``output_layer``); torch is the container for the parameters and the autograd engine for training.
Every action the policy produces for an environment — ``compute_action``, ``Evaluator`` roll-outs, the
data-collection passes of imitation training — is computed by libqtrack (``qt_policy_action_batched``,
``qt_rollout_policy``) from the packed weights; there is no CPU inference path.
"""
from __future__ import annotations

import json
import logging
from pathlib import Path

import numpy as np

from . import _lib
from .controller import ACTION_KEYS, BaseController, _validate_observation, observation_to_vector
from .params import ACTIVATIONS, MLP_MAX_LAYERS, MLP_MAX_WIDTH, QtMlp, QtParams, make_params

logger = logging.getLogger(__name__)

DEFAULT_OUTPUT_BOUNDS = {"thrust": (0.0, 20.0), "roll_rate": (-3.0, 3.0), "pitch_rate": (-3.0, 3.0),
                         "yaw_rate": (-3.0, 3.0)}


def _pad4(x: int) -> int:
    return (x + 3) & ~3


def _torch():
    import torch

    return torch


_POLICY_NETWORK_CLASS = None


def _policy_network_class():
    """deep_tracking_policy.py:24-147 as a torch module (parameters + training graph only).  The class is created on
    first use so that importing this package does not import torch."""
    global _POLICY_NETWORK_CLASS
    if _POLICY_NETWORK_CLASS is not None:
        return _POLICY_NETWORK_CLASS
    torch = _torch()
    nn = torch.nn

    class PolicyNetwork(nn.Module):
        def __init__(self, input_dim: int = 18, hidden_sizes=None, activation: str = "relu", output_bounds=None):
            super().__init__()
            self.input_dim = input_dim
            self.hidden_sizes = list(hidden_sizes) if hidden_sizes is not None else [64, 64]
            self.activation_name = activation
            self.output_dim = 4
            self.output_bounds = dict(output_bounds) if output_bounds is not None else dict(DEFAULT_OUTPUT_BOUNDS)
            acts = {"relu": nn.ReLU(), "tanh": nn.Tanh(), "elu": nn.ELU(), "leaky_relu": nn.LeakyReLU(0.1)}
            if activation not in acts:
                raise ValueError(f"Unknown activation: {activation}. Valid: {list(acts)}")
            layers, prev = [], input_dim
            for h in self.hidden_sizes:
                layers += [nn.Linear(prev, h), acts[activation]]
                prev = h
            self.feature_extractor = nn.Sequential(*layers)
            self.output_layer = nn.Linear(prev, self.output_dim)
            for m in self.modules():  # deep_tracking_policy.py:103-109
                if isinstance(m, nn.Linear):
                    nn.init.xavier_uniform_(m.weight)
                    if m.bias is not None:
                        nn.init.zeros_(m.bias)

        def forward(self, x):
            raw = self.output_layer(self.feature_extractor(x))
            sig = torch.sigmoid(raw)
            b = [self.output_bounds[k] for k in ACTION_KEYS]
            lo = torch.tensor([v[0] for v in b], dtype=raw.dtype, device=raw.device)
            sc = torch.tensor([v[1] - v[0] for v in b], dtype=raw.dtype, device=raw.device)
            return sig * sc + lo

        def linear_layers(self):
            return [m for m in self.feature_extractor if isinstance(m, nn.Linear)] + [self.output_layer]

    _POLICY_NETWORK_CLASS = PolicyNetwork
    return PolicyNetwork


def make_policy_network(input_dim: int = 18, hidden_sizes=None, activation: str = "relu", output_bounds=None):
    return _policy_network_class()(input_dim=input_dim, hidden_sizes=hidden_sizes, activation=activation,
                                   output_bounds=output_bounds)


def __getattr__(name: str):
    if name == "PolicyNetwork":          # `from ...policy import PolicyNetwork` (controllers/__init__.py exports it)
        return _policy_network_class()
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


class DeepTrackingPolicy(BaseController):
    """deep_tracking_policy.py:150-415.  ``device`` is accepted for interface compatibility; the
    network's parameters live on the GPU libqtrack runs on."""

    def __init__(self, config: dict | None = None, device: str | None = None):
        super().__init__(name="deep_tracking", config=config)
        config = config or {}
        torch = _torch()
        self._ctx = _lib.get_context()
        self.device = self._ctx.device
        self.network = make_policy_network(18, config.get("hidden_sizes", [64, 64]), config.get("activation", "relu"),
                                           config.get("output_bounds", None)).to(self.device)
        sizes = [18, *self.network.hidden_sizes, 4]
        if len(sizes) - 1 > MLP_MAX_LAYERS or max(sizes[1:]) > MLP_MAX_WIDTH:
            raise ValueError(f"libqtrack supports at most {MLP_MAX_LAYERS} linear layers of width <= {MLP_MAX_WIDTH}")
        self._packed = None
        self._packed_version = None
        self._dev = None
        # extension of the reference's config: evaluate the hidden layers on the tensor cores inside qt_rollout_policy
        # (tcgen05, 3xTF32; networks with one to three hidden layers of at most 64 units; see include/qtrack.h
        # QT_FLAG_POLICY_TENSOR_CORES).  Off by default: the FP32-pipe evaluation is the parity path.
        self.tensor_cores = bool(config.get("tensor_cores", False))
        checkpoint_path = config.get("checkpoint_path")
        if checkpoint_path is not None:
            self.load_checkpoint(checkpoint_path)
        self.network.eval()
        # attributes make_params() reads from a controller: the policy has no control-law block
        self.dt, self.hover_thrust = 0.01, self.mass * self.gravity
        self.max_thrust, self.min_thrust, self.max_rate = 20.0, 0.0, 3.0
        self.use_lqi, self.integral_limit, self.integral_zero_threshold = False, 10.0, 0.01
        self.feedforward_enabled = False
        self.ff_velocity_gain, self.ff_acceleration_gain = np.zeros(3), np.zeros(3)
        self.ff_max_velocity, self.ff_max_acceleration = 10.0, 5.0

    # -- what the kernels need ---------------------------------------------------------------------
    def carries_state(self) -> bool:
        return False

    def kernel_params(self, env_config=None, criteria=None) -> QtParams:
        return make_params(env_config, None, criteria)

    def state_vector(self) -> np.ndarray:      # Evaluator plumbing: the network is stateless
        return np.zeros(4)

    def load_state_vector(self, v, absolute_count: bool) -> None:
        pass

    def _param_version(self):
        return tuple(int(p._version) for p in self.network.parameters())

    def packed_mlp(self) -> QtMlp:
        """Pack the current parameters into the layout of ``struct qt_mlp`` (transposed, rows padded to
        multiples of four, zeros in the padding).  Re-packed only after the parameters changed."""
        torch = _torch()
        ver = self._param_version()
        if self._packed is None or ver != self._packed_version:
            sizes = [18, *self.network.hidden_sizes, 4]
            chunks = []
            with torch.no_grad():
                for l, lin in enumerate(self.network.linear_layers()):
                    rows = sizes[l] if l == 0 else _pad4(sizes[l])
                    outp = _pad4(sizes[l + 1])
                    wt = torch.zeros((rows, outp), dtype=torch.float32, device=self.device)
                    wt[: sizes[l], : sizes[l + 1]] = lin.weight.detach().to(torch.float32).t()
                    b = torch.zeros(outp, dtype=torch.float32, device=self.device)
                    b[: sizes[l + 1]] = lin.bias.detach().to(torch.float32)
                    chunks += [wt.reshape(-1), b]
                self._packed = torch.cat(chunks).contiguous()
            self._packed_version = ver
        m = QtMlp()
        sizes = [18, *self.network.hidden_sizes, 4]
        m.n_layers = len(sizes) - 1
        for i, d in enumerate(sizes):
            m.dims[i] = d
        m.activation = ACTIVATIONS[self.network.activation_name]
        for i, k in enumerate(ACTION_KEYS):
            lo, hi = self.network.output_bounds[k]
            m.out_lo[i] = float(lo)
            m.out_scale[i] = float(hi - lo)
        m.weights_dev = self._packed.data_ptr()
        assert _lib.load_library().qt_mlp_packed_floats(m) == self._packed.numel()
        return m

    # -- reference interface -----------------------------------------------------------------------
    def compute_action(self, observation: dict) -> dict:
        torch = _torch()
        _validate_observation(observation)
        if self._dev is None:
            self._dev = (torch.zeros((22, 1), dtype=torch.float64, device=self.device),
                         torch.zeros((4, 1), dtype=torch.float64, device=self.device))
        obs, act = self._dev
        self._ctx.bind_torch_stream()
        obs.copy_(torch.from_numpy(observation_to_vector(observation).reshape(22, 1)))
        self._ctx.policy_action_batched(self.packed_mlp(), 1, obs, act)
        a = act.cpu().numpy().reshape(4)
        return dict(zip(ACTION_KEYS, (float(x) for x in a)))

    def compute_action_batched(self, obs):
        """``obs`` CUDA float64 [22, E] -> actions [4, E]."""
        torch = _torch()
        E = int(obs.shape[1])
        act = torch.empty((4, E), dtype=torch.float64, device=obs.device)
        self._ctx.bind_torch_stream()
        self._ctx.policy_action_batched(self.packed_mlp(), E, obs.contiguous(), act)
        return act

    def _extract_features(self, observation: dict) -> np.ndarray:
        v = observation_to_vector(observation)
        return np.concatenate([v[12:15] - v[0:3], v[15:18] - v[3:6], v[6:9], v[9:12], v[12:15], v[15:18]]).astype(np.float32)

    def reset(self) -> None:
        pass

    def train_mode(self) -> None:
        self.network.train()

    def eval_mode(self) -> None:
        self.network.eval()

    def get_parameters(self):
        return list(self.network.parameters())

    def save_checkpoint(self, path: str | Path, metadata: dict | None = None) -> None:
        torch = _torch()
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        ckpt = {"model_state_dict": self.network.state_dict(), "config": self.config,
                "input_dim": self.network.input_dim, "hidden_sizes": self.network.hidden_sizes,
                "activation": self.network.activation_name, "output_bounds": self.network.output_bounds}
        if metadata:
            ckpt["metadata"] = metadata
        torch.save(ckpt, path)

    def load_checkpoint(self, path: str | Path) -> dict:
        torch = _torch()
        path = Path(path)
        if not path.exists():
            raise FileNotFoundError(f"Checkpoint not found: {path}")
        ckpt = torch.load(path, map_location=self.device, weights_only=False)
        if ckpt.get("input_dim") != self.network.input_dim:
            raise ValueError(f"Input dim mismatch: checkpoint has {ckpt.get('input_dim')}, "
                             f"network has {self.network.input_dim}")
        if ckpt.get("hidden_sizes") != self.network.hidden_sizes:
            raise ValueError(f"Hidden sizes mismatch: checkpoint has {ckpt.get('hidden_sizes')}, but network is "
                             f"configured with {self.network.hidden_sizes}. Cannot load incompatible model.")
        self.network.load_state_dict(ckpt["model_state_dict"])
        self._packed = None
        return ckpt.get("metadata", {})

    def load_arrays(self, weights: list[np.ndarray], biases: list[np.ndarray]) -> None:
        """Set the parameters from per-layer arrays in torch's layout (W [out, in], b [out])."""
        torch = _torch()
        with torch.no_grad():
            for lin, W, b in zip(self.network.linear_layers(), weights, biases):
                lin.weight.copy_(torch.as_tensor(np.asarray(W), dtype=torch.float32))
                lin.bias.copy_(torch.as_tensor(np.asarray(b), dtype=torch.float32))
        self._packed = None


def create_controller_from_config(config_path: str | Path) -> DeepTrackingPolicy:
    with open(config_path) as f:
        return DeepTrackingPolicy(config=json.load(f))
