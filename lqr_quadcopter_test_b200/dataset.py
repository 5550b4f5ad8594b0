"""Teacher data set emission (BASELINE config 5).

The reference's imitation trainer queries the Riccati-LQR supervisor at every state the env visits
(train.py:683-696) and feeds the policy the 18-D feature vector of
DeepTrackingPolicy._extract_features (controllers/deep_tracking_policy.py:241-286).  Here the
teacher itself is rolled out for E episodes and the fused kernel writes, every ``stride`` steps,
one fp32 feature row and one fp32 action row per episode straight into torch CUDA tensors.
"""
from __future__ import annotations

from . import _lib
from .config import EnvConfig
from .eval import rollout_device
from .metrics import SuccessCriteria


def teacher_dataset(controller, env_config: EnvConfig | None = None, num_episodes: int = 1024,
                    base_seed: int = 42, stride: int = 10, max_steps: int | None = None,
                    reset_controller_between_episodes: bool = True, flatten: bool = True, student=None):
    """Roll out ``controller`` and return ``(features, actions, metrics)`` as CUDA tensors.

    With ``student`` (a ``DeepTrackingPolicy``) the student's network drives the environment and
    ``controller`` only labels the states it visits - the data-collection pass of the reference's
    imitation trainer (train.py:654-735), where ``actions`` are the supervisor's.

    features float32 [N, 18], actions float32 [N, 4] with N = slots * num_episodes, row
    ``slot * num_episodes + episode`` (``flatten=False`` keeps [slots, E, .]); metrics float64
    [16, E].  ``stride`` defaults to the env's own log interval (env/config.py:82)."""
    import torch

    env_config = env_config or EnvConfig()
    ctx = _lib.get_context()
    seeds = torch.arange(num_episodes, dtype=torch.int64, device=ctx.device) + int(base_seed)
    if student is not None:
        out = rollout_device(student, env_config, SuccessCriteria(), seeds, max_steps, dataset_stride=stride,
                             ctx=ctx, supervisor=controller)
    else:
        chain = 1 if (reset_controller_between_episodes or not controller.carries_state()) else max(num_episodes, 2)
        out = rollout_device(controller, env_config, SuccessCriteria(), seeds, max_steps, chain=chain,
                             dataset_stride=stride, ctx=ctx)
    feats, acts = out["features"], out["actions"]
    if flatten:
        feats = feats.reshape(-1, feats.shape[-1])
        acts = acts.reshape(-1, acts.shape[-1])
    return feats, acts, out["metrics"]
