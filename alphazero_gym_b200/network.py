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


class PolicyNet(nn.Module):
    """Same parameter structure (and state_dict keys) as the reference's DiscretePolicy / DiagonalNormalPolicy /
    DiagonalGMMPolicy (alphazero/network/policies.py:163-259, :355-434, :502-588): `trunk` Sequential of
    Linear + activation, `value_head`, `dist_head`.  Used where the reference itself cannot be imported
    (tests and benchmarks on the GPU box); a reference model can be passed to the search classes directly."""

    def __init__(self, state_dim: int, hidden: int, n_hidden: int, head_dim: int, nonlinearity: str = "relu",
                 num_components: int = 1, num_actions: int = 0, action_bound: float = 2.0,
                 log_param_min: float = -5.0, log_param_max: float = 2.0):
        super().__init__()
        act = {"relu": nn.ReLU, "elu": nn.ELU, "leakyrelu": nn.LeakyReLU, "relu6": nn.ReLU6, "swish": nn.SiLU, "silu": nn.SiLU,
               "hardswish": nn.Hardswish}[nonlinearity.lower()]  # the reference's map (alphazero/network/utils.py:5-14)
        layers, d = [], state_dim
        for _ in range(n_hidden):
            layers += [nn.Linear(d, hidden), act()]
            d = hidden
        self.trunk = nn.Sequential(*layers)
        self.value_head = nn.Linear(hidden, 1)
        self.dist_head = nn.Linear(hidden, head_dim)
        self.state_dim, self.num_components, self.num_actions = state_dim, num_components, num_actions
        self.action_bound, self.log_param_min, self.log_param_max = action_bound, log_param_min, log_param_max
        self.layernorm = False

    def load_flat(self, flat: np.ndarray) -> "PolicyNet":
        off = 0
        with torch.no_grad():
            for prm in self.state_dict().values():
                n = prm.numel()
                prm.copy_(torch.from_numpy(np.asarray(flat[off:off + n], np.float32)).reshape(prm.shape))
                off += n
        assert off == len(flat)
        return self


def make_policy(representation_dim: int, action_dim: int, distribution: str, hidden_dimensions, nonlinearity: str,
                num_components: int = None, num_actions: int = None, action_bound: float = None, layernorm: bool = False,
                log_param_min: float = -5, log_param_max: float = 2) -> PolicyNet:
    """Same keyword arguments as the reference's `make_policy` (alphazero/network/policies.py:806-916, the `_target_` of
    config/policy/*.yaml); returns a PolicyNet with the reference's parameter layout (discrete / squashed Normal / GMM heads)."""
    if layernorm:
        raise NotImplementedError("layernorm=True policies are not supported by the CUDA evaluation kernel")
    if len(set(hidden_dimensions)) != 1:
        raise NotImplementedError("all hidden layers must have the same width")
    if distribution == "discrete":
        return PolicyNet(representation_dim, hidden_dimensions[0], len(hidden_dimensions), num_actions, nonlinearity, num_actions=num_actions)
    if distribution == "beta":
        raise NotImplementedError("GeneralizedBetaPolicy is declared broken upstream (README.md:21-22) and is not implemented")
    if distribution != "normal" or action_dim != 1:
        raise NotImplementedError("continuous policies: distribution 'normal' with action_dim 1 (Pendulum)")
    K = int(num_components or 1)
    return PolicyNet(representation_dim, hidden_dimensions[0], len(hidden_dimensions), policy_head_dim("continuous", num_components=K),
                     nonlinearity, num_components=K, action_bound=float(action_bound or 0.0), log_param_min=log_param_min,
                     log_param_max=log_param_max)


def describe_model(model) -> dict:
    """Network shape the engine needs, read off a reference-style policy module (policies.py:806 make_policy
    products or PolicyNet): trunk widths, activation, head size, mixture components, log-std clamp."""
    if getattr(model, "layernorm", False):
        raise NotImplementedError("layernorm=True policies are not supported by the CUDA evaluation kernel")
    cls = type(model).__name__
    if "Beta" in cls:
        # GeneralizedBetaPolicy (policies.py:672) has the same parameter shapes as the squashed Normal (head size 2 * action_dim) but
        # samples from a Beta distribution: evaluating it as a Normal would be silently wrong (and upstream declares it broken)
        raise NotImplementedError(f"{cls}: only DiscretePolicy, DiagonalNormalPolicy and DiagonalGMMPolicy heads are implemented")
    sd = model.state_dict()
    tw = [k for k in sd if k.startswith("trunk.") and k.endswith(".weight")]
    if not tw or "value_head.weight" not in sd or "dist_head.weight" not in sd:
        raise TypeError("model must expose trunk.*, value_head and dist_head parameters (alphazero/network/policies.py)")
    hidden = sd[tw[0]].shape[0]
    for k in tw:
        if sd[k].shape[0] != hidden:
            raise NotImplementedError("all hidden layers must have the same width")
    from ._cabi import ACT_BY_NAME
    act_name = type(model.trunk[1]).__name__.lower()
    if act_name not in ACT_BY_NAME:
        raise NotImplementedError(f"activation {act_name} is not in the reference's map (alphazero/network/utils.py:5-14)")
    return dict(state_dim=sd[tw[0]].shape[1], hidden=hidden, n_hidden=len(tw), activation=ACT_BY_NAME[act_name],
                head_dim=sd["dist_head.weight"].shape[0], num_components=int(getattr(model, "num_components", 1) or 1),
                action_bound=float(getattr(model, "action_bound", None) or 0.0),
                log_std_min=float(getattr(model, "log_param_min", -5.0)), log_std_max=float(getattr(model, "log_param_max", 2.0)))


def weights_version(model) -> int:
    """Cheap change detector: in-place optimizer updates bump every parameter's _version.  Updates that bypass autograd's version
    counters -- a training step replayed from a CUDA graph (train.Trainer(cuda_graph=True)) -- must call `bump_weights_version`
    (Trainer does), or the search classes' `refresh_weights()`."""
    return sum(int(p._version) + (p.data_ptr() % 1000003) for p in model.parameters()) + int(getattr(model, "_azg_weights_epoch", 0))


def bump_weights_version(model) -> None:
    """Tell every search object that shares `model` that its weights changed (see weights_version)."""
    model._azg_weights_epoch = int(getattr(model, "_azg_weights_epoch", 0)) + 1
