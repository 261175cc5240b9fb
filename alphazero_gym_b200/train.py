"""Training step on the device: `ContinuousAgent.update` / `DiscreteAgent.update` for batches that are already
in HBM (DeviceReplayBuffer rows), with the weights handed back to the search engine afterwards.

Reference: alphazero/agent/agents.py:319-389 (discrete), :539-603 (continuous); losses alphazero/agent/losses.py:154-327 (A0CLoss),
:329-500 (A0CLossTuned); policy log-probabilities alphazero/network/policies.py:303-327 (DiscretePolicy.get_train_data), :463-485
(DiagonalNormalPolicy), :633-654 (DiagonalGMMPolicy) with SquashedNormal (alphazero/network/distributions.py:50-109, :205-245).
PyTorch autograd is the plumbing here (SURVEY: "stays PyTorch autograd first"); the formulas are restated explicitly -- no
torch.distributions objects are built per step -- and pinned against the unmodified reference (oracle/gen_train_golden.py ->
tests/golden/train_*.npz, tests/test_train_step.py: losses within 1e-5 relative over three consecutive steps, weights within 2e-6).

Quirks reproduced on purpose:
  * K = 1 (DiagonalNormalPolicy): the squashing correction is `x.shape[-1] * log(bound)` with x of shape [batch, num_actions], i.e. the
    NUMBER OF ROOT ACTIONS multiplies log(bound) (distributions.py:106); for the GMM x is [batch, num_actions, 1] and the factor is 1.
  * atanh(a / (bound + 1e-6)) and the matching correction factor 1 + 1e-6 / bound (distributions.py:84, :105).
  * discrete: counts += 1 before the logarithm (agents.py:362-365); A0CLossTuned: alpha enters the network loss as a detached Python
    float and is then updated by its own Adam step on log_alpha inside the loss call (losses.py:431-456, :486-490).
  * DiscreteAgent + AlphaZeroLoss does not run upstream (agents.py:378-381 passes a Categorical to F.cross_entropy), so only the A0C
    losses exist here.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from .network import PolicyNet, bump_weights_version

_LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


def _normal_log_prob(x: torch.Tensor, mu: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
    # torch.distributions.Normal.log_prob: -((x - mu)^2) / (2 var) - log(sigma) - log(sqrt(2 pi))
    var = sigma ** 2
    return -((x - mu) ** 2) / (2 * var) - sigma.log() - _LOG_SQRT_2PI


def _squash_correction(x: torch.Tensor, bound: float, n_event: int, eps: float = 1e-6) -> torch.Tensor:
    # ScaledTanhTransform.log_abs_det_jacobian (distributions.py:84-109)
    c = 1 + eps / bound
    return n_event * math.log(bound) + 2.0 * (math.log(2.0) - c * x - F.softplus(-2.0 * c * x))


def policy_train_data(net: PolicyNet, states: torch.Tensor, actions: torch.Tensor):
    """(log_probs [B, A], entropy [B] or [B, A], V_hat [B, 1]) as the reference policies' get_train_data return them."""
    h = net.trunk(states)
    V_hat = net.value_head(h)
    out = net.dist_head(h)
    if net.num_actions:  # DiscretePolicy (policies.py:303-327): one Categorical per root action
        logp_all = out - out.logsumexp(dim=-1, keepdim=True)
        log_probs = logp_all.gather(1, actions.long())
        p = logp_all.exp()
        entropy = -(logp_all.clamp(min=torch.finfo(logp_all.dtype).min) * p).sum(-1, keepdim=True).expand(-1, actions.shape[1])
        return log_probs, entropy, V_hat
    K, bound, eps = net.num_components, net.action_bound, 1e-6
    if K > 1:  # DiagonalGMMPolicy (policies.py:590-654), action_dim = 1
        mu, log_std, log_coeff = out[:, :K], out[:, K:2 * K], out[:, 2 * K:]
        sigma = log_std.clamp(min=net.log_param_min, max=net.log_param_max).exp()
        a = actions.unsqueeze(-1)                                  # MixtureSameFamily pads the value: [B, A, 1]
        mu, sigma = mu.unsqueeze(1), sigma.unsqueeze(1)            # [B, 1, K], broadcast over the root actions
        if bound:
            x = torch.atanh(a / (bound + eps))
            comp = _normal_log_prob(x, mu, sigma) - _squash_correction(x, bound, 1)
        else:
            comp = _normal_log_prob(a, mu, sigma)
        log_mix = torch.log_softmax(log_coeff, dim=-1).unsqueeze(1)
        log_probs = torch.logsumexp(comp + log_mix, dim=-1)
    else:  # DiagonalNormalPolicy (policies.py:436-485)
        mu, log_std = out.chunk(2, dim=-1)
        sigma = log_std.clamp(min=net.log_param_min, max=net.log_param_max).exp()
        if bound:
            x = torch.atanh(actions / (bound + eps))
            log_probs = _normal_log_prob(x, mu, sigma) - _squash_correction(x, bound, actions.shape[-1])
        else:
            log_probs = _normal_log_prob(actions, mu, sigma)
    return log_probs, -log_probs.mean(dim=-1), V_hat


@dataclass
class LossConfig:
    """config/loss/A0CLoss.yaml / A0CLossTuned.yaml."""
    tuned: bool = True
    tau: float = 0.1
    policy_coeff: float = 0.1
    value_coeff: float = 1.0
    alpha: float = 1.0          # A0CLoss: fixed temperature; A0CLossTuned: alpha_init
    reduction: str = "mean"
    alpha_lr: float = 0.001     # A0CLossTuned.lr
    action_dim: int = 1         # target entropy = -action_dim
    alpha_grad_clip: float = 0.0


class A0CLoss:
    """A0CLoss / A0CLossTuned (losses.py:154-327, :329-500) on device tensors."""

    def __init__(self, cfg: LossConfig, device, capturable: bool = False):
        self.cfg = cfg
        self.capturable = capturable
        if cfg.tuned:
            self.log_alpha = torch.tensor(np.log(cfg.alpha), requires_grad=True, device=device, dtype=torch.float32)
            self.alpha = self.log_alpha.exp()
            self.alpha_opt = torch.optim.Adam([self.log_alpha], lr=cfg.alpha_lr, capturable=capturable)

    def _reduce(self, x: torch.Tensor) -> torch.Tensor:
        return x.mean() if self.cfg.reduction == "mean" else x.sum()

    def __call__(self, log_probs, counts, entropy, V, V_hat) -> Dict[str, torch.Tensor]:
        c = self.cfg
        with torch.no_grad():
            log_diff = log_probs - c.tau * torch.log(counts)
        policy_loss = c.policy_coeff * self._reduce(torch.einsum("ni, ni -> n", log_diff, log_probs))
        value_loss = c.value_coeff * F.mse_loss(V_hat, V, reduction=c.reduction)
        if not c.tuned:
            entropy_loss = c.alpha * self._reduce(entropy)
            return {"loss": policy_loss + entropy_loss + value_loss, "policy_loss": policy_loss, "entropy_loss": entropy_loss,
                    "value_loss": value_loss}
        # the reference multiplies by alpha.detach().item(): the same f32 value, but a host synchronisation; under CUDA-graph capture
        # the factor stays on the device
        if self.capturable:  # alpha = exp(log_alpha) recomputed from the parameter the optimizer updates in place: no state in Python
            self.alpha = self.log_alpha.exp()
        entropy_loss = (self.alpha.detach() if self.capturable else self.alpha.detach().item()) * self._reduce(entropy)
        loss = policy_loss + entropy_loss + value_loss
        # _update_alpha (losses.py:431-456)
        if self.capturable:
            # d/d log_alpha of mean(alpha * c), c = (entropy + |A|).detach(), is alpha * mean(c): written out, so that the captured step
            # holds one autograd backward (the network's) instead of two
            cm = (entropy - (-c.action_dim)).detach().mean()
            alpha_loss = self.alpha.detach() * cm
            self.log_alpha.grad = alpha_loss.clone()
        else:
            self.log_alpha.grad = None
            alpha_loss = (self.alpha * (entropy - (-c.action_dim)).detach()).mean()
            alpha_loss.backward()
        if c.alpha_grad_clip:
            torch.nn.utils.clip_grad_norm_(self.log_alpha, c.alpha_grad_clip)
        self.alpha_opt.step()
        self.alpha = self.log_alpha.exp()
        return {"loss": loss, "policy_loss": policy_loss, "entropy_loss": entropy_loss, "value_loss": value_loss,
                "alpha_loss": alpha_loss.detach()}


def make_optimizer(name: str, params, lr: float = 0.001, capturable: bool = False):
    """config/optimizer/RMSProp.yaml (default of both run configs) / Adam.yaml."""
    if name.lower() == "rmsprop":
        return torch.optim.RMSprop(params, lr=lr, momentum=0, weight_decay=0, alpha=0.9, eps=1e-10, capturable=capturable)
    if name.lower() == "adam":
        return torch.optim.Adam(params, lr=lr, betas=(0.9, 0.99), weight_decay=0, eps=1e-07, amsgrad=False, capturable=capturable)
    raise ValueError(f"unknown optimizer {name}")


def root_children(c_pw: float, kappa: float, n_rollouts: int) -> int:
    """Children of the root after a finished continuous search: the progressive-widening limit at the last root visit count
    (states.py:252-275 with n = n_rollouts - 1)."""
    # The root holds one action before the first simulation (mcts.py:673) and gains AT MOST one per simulation, when
    # ceil(c_pw * (n + 1)^kappa) exceeds its child count at root visit count n (mcts.py:725-727): with a large c_pw and few
    # rollouts the closed form ceil(c_pw * N^kappa) over-counts (c_pw = 3, N = 2: 3 children, not 5), and the padded columns
    # (count 0) would put log(0) into the loss.
    kids = 1
    for n in range(n_rollouts):
        if int(math.ceil(c_pw * float(n + 1) ** kappa)) > kids:
            kids += 1
    return kids


class Trainer:
    """`Agent.update` on device tensors + hand-over of the new weights to a SearchEngine.

    net: PolicyNet (same state_dict layout as the reference policies, so `flat_weights()` is what azg_set_weights expects);
    batches: dicts with the DeviceReplayBuffer keys (obs, actions, counts, V_target)."""

    def __init__(self, net: PolicyNet, loss: LossConfig, optimizer: str = "rmsprop", lr: float = 0.001, grad_clip: float = 0.0,
                 cuda_graph: bool = False, n_root_actions: Optional[int] = None):
        """n_root_actions: the engine pads root-result rows to azg_cmax columns (count 0, action 0); the reference's rows hold exactly the
        root's children -- ceil(c_pw * n_rollouts^kappa) for the continuous search (`root_children`), the action count for the discrete
        one -- and log(0) would poison the loss, so only the first n_root_actions columns are used when it is given.
        cuda_graph=True (CUDA devices): `update` captures the whole step -- forward, backward, both optimizer steps -- into a CUDA
        graph on its first call for a batch shape and replays it afterwards: the step is ~150 small kernels, i.e. launch-bound
        (2.1 ms eager at any batch size up to 64 K rows, tools/train_bench.py)."""
        self.net = net
        self.device = next(net.parameters()).device
        self.cuda_graph = bool(cuda_graph) and self.device.type == "cuda"
        self.loss = A0CLoss(loss, self.device, capturable=self.cuda_graph)
        # `optimizer`: "rmsprop" / "adam" (the reference's two configs) or an already constructed torch optimizer over net.parameters()
        self.opt = make_optimizer(optimizer, net.parameters(), lr, capturable=self.cuda_graph) if isinstance(optimizer, str) else optimizer
        self.clip = grad_clip
        self.discrete = bool(net.num_actions)
        self.n_root_actions = n_root_actions
        self._graphs: Dict[tuple, tuple] = {}

    def update(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """One gradient step; returns the loss components as 0-d device tensors (no host synchronisation except, without
        cuda_graph, A0CLossTuned's `alpha.item()`, which the reference has as well)."""
        if self.n_root_actions is not None:
            batch = dict(batch, actions=batch["actions"][:, :self.n_root_actions], counts=batch["counts"][:, :self.n_root_actions])
        if not self.cuda_graph:
            return self._step(batch)
        key = tuple((k, tuple(batch[k].shape), batch[k].dtype) for k in ("obs", "actions", "counts", "V_target"))
        if key not in self._graphs:
            self._graphs[key] = self._capture(batch)
        graph, static_in, static_out = self._graphs[key]
        for k, t in static_in.items():
            t.copy_(batch[k], non_blocking=True)
        graph.replay()
        bump_weights_version(self.net)  # graph replays do not touch Parameter._version: drop-in search objects sharing the net must reload
        return {k: v.clone() for k, v in static_out.items()}

    def _capture(self, batch):
        """Warm up on a side stream (the capture rules), restore every bit of state the warm-up steps changed, capture one step."""
        static_in = {k: batch[k].detach().to(self.device).clone() for k in ("obs", "actions", "counts", "V_target")}
        opts = [self.opt] + ([self.loss.alpha_opt] if self.loss.cfg.tuned else [])
        snap_net = {k: v.clone() for k, v in self.net.state_dict().items()}
        snap_opt = [{p: {k: v.clone() for k, v in st.items() if torch.is_tensor(v)} for p, st in o.state.items()} for o in opts]
        snap_alpha = self.loss.log_alpha.detach().clone() if self.loss.cfg.tuned else None

        def restore():  # in place: the graph is captured against these very tensors
            self.net.load_state_dict(snap_net)
            for o, saved in zip(opts, snap_opt):
                for p, st in o.state.items():
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            v.copy_(saved[p][k]) if p in saved else v.zero_()  # a fresh optimizer starts from zeros
            if snap_alpha is not None:
                with torch.no_grad():
                    self.loss.log_alpha.copy_(snap_alpha)

        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                self._step(static_in)
        torch.cuda.current_stream(self.device).wait_stream(side)
        restore()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_out = self._step(static_in)
        return graph, static_in, static_out

    def _step(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        for prm in self.net.parameters():
            prm.grad = None
        states = batch["obs"].to(self.device, torch.float32)
        V = batch["V_target"].to(self.device).unsqueeze(1).float()
        counts = batch["counts"].to(self.device)
        if self.discrete:
            counts = counts + 1  # agents.py:362-365
        counts = counts.float()
        actions = batch["actions"].to(self.device).float()
        log_probs, entropy, V_hat = policy_train_data(self.net, states, actions)
        out = self.loss(log_probs=log_probs, counts=counts, entropy=entropy, V=V, V_hat=V_hat)
        out["loss"].backward()
        if self.clip:
            torch.nn.utils.clip_grad_norm_(self.net.parameters(), self.clip)
        self.opt.step()
        return {k: v.detach() for k, v in out.items()}

    def flat_weights(self) -> torch.Tensor:
        """Flat f32 weights in state_dict order, on the device (azg_set_weights takes a device pointer)."""
        sd = self.net.state_dict()
        keys = [k for k in sd if k.startswith("trunk.")] + ["value_head.weight", "value_head.bias", "dist_head.weight", "dist_head.bias"]
        return torch.cat([sd[k].detach().reshape(-1).float() for k in keys])

    def push_weights(self, engine, broadcast: Optional[bool] = None) -> None:
        """New weights -> search engine(s): NCCL / gloo broadcast from rank 0 when a process group is up (C1), then azg_set_weights."""
        w = self.flat_weights()
        import torch.distributed as dist
        if broadcast if broadcast is not None else (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            dist.broadcast(w, src=0)
        engine.set_weights(w)
