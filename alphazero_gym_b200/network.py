"""Network shape helpers on the host side.

The engine consumes a flat f32 weight vector in state_dict order (include/azg.h azg_set_weights).  A user
of the reference passes `model.state_dict()` of its own policy (alphazero/network/policies.py); for
benchmarks and tests that cannot import the reference, `init_policy_weights` builds the same
default-initialised weights: torch.manual_seed(seed) followed by nn.Linear constructions in the order
policies.py creates them (trunk layers :101-118 / :238-254, value_head :120 / :257, dist_head :434 / :588 /
:259), which reproduces `torch.manual_seed(seed); make_policy(...)` bit for bit.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
from torch import nn


def policy_head_dim(variant: str, num_actions: int = 2, num_components: int = 2, action_dim: int = 1) -> int:
    if variant == "discrete":
        return num_actions
    if num_components > 1:
        return num_components * (2 * action_dim + 1)
    return 2 * action_dim


def init_policy_state_dict(seed: int, state_dim: int, hidden: int, n_hidden: int, head_dim: int) -> "OrderedDict[str, torch.Tensor]":
    torch.manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    dims = [state_dim] + [hidden] * n_hidden
    for i in range(n_hidden):
        lin = nn.Linear(dims[i], dims[i + 1])
        sd[f"trunk.{2 * i}.weight"], sd[f"trunk.{2 * i}.bias"] = lin.weight.detach(), lin.bias.detach()
    v = nn.Linear(hidden, 1)
    sd["value_head.weight"], sd["value_head.bias"] = v.weight.detach(), v.bias.detach()
    d = nn.Linear(hidden, head_dim)
    sd["dist_head.weight"], sd["dist_head.bias"] = d.weight.detach(), d.bias.detach()
    return sd


def init_policy_weights(seed: int, state_dim: int, hidden: int, n_hidden: int, head_dim: int) -> np.ndarray:
    sd = init_policy_state_dict(seed, state_dim, hidden, n_hidden, head_dim)
    return np.concatenate([t.numpy().astype(np.float32).ravel() for t in sd.values()])
