"""Drop-in agents: the reference's constructors, `act()`, `update()` and `train()` over the CUDA search engine.

Mirrors alphazero/agent/agents.py: `Agent.__init__` (:45-90: policy / mcts / loss / optimizer built from Hydra configs with `_target_`),
`DiscreteAgent` (:192-389), `ContinuousAgent` (:395-603), `reset_mcts` (:146-155), `mcts_forward` (:305-317), `train` (:157-184) and
the read-only properties (:92-144).  Point `config/agent/*.yaml:_target_` at these classes (this repo's config/agent/*.yaml do) and the
reference's run scripts construct them unchanged.  Hydra itself is optional: when it is not installed a ten-line resolver of
`_target_` does what `hydra.utils.call / instantiate` do for these configs.

What runs where: `act()` = batched-engine search of one tree (search/mcts.py) + the reference's final-action rule on the host (12
lines); `update()` = train.Trainer on device tensors (pinned against the unmodified reference by tests/test_train_step.py).
Return values, their order and dtypes follow the reference: (action, state, actions, counts, Qs, V).
"""
from __future__ import annotations

import importlib
import random
from collections import defaultdict
from typing import Any, Dict, Mapping, Tuple

import numpy as np
import torch

from ..helpers import stable_normalizer
from ..search.mcts import MCTSContinuous, MCTSDiscrete


def _resolve(cfg: Mapping, **kw):
    """hydra.utils.call / instantiate for a config with `_target_` (what the reference's Agent.__init__ does, agents.py:81-86)."""
    try:
        import hydra  # noqa: F401
        from hydra.utils import instantiate
        return instantiate(cfg, **kw)
    except ImportError:
        c = {k: v for k, v in dict(cfg).items()}
        mod, name = c.pop("_target_").rsplit(".", 1)
        c.update(kw)
        return getattr(importlib.import_module(mod), name)(**c)


class Agent:
    """Common part (reference agents.py:19-184)."""

    def __init__(self, policy_cfg, loss_cfg, mcts_cfg, optimizer_cfg, final_selection: str, train_epochs: int, grad_clip: float,
                 device: str) -> None:
        self.device = torch.device(device)
        self.nn = _resolve(policy_cfg).to(self.device)
        self.mcts = _resolve(mcts_cfg, model=self.nn)
        self.loss = _resolve(loss_cfg).to(self.device)
        self.optimizer = _resolve(optimizer_cfg, params=self.nn.parameters())
        self.final_selection = final_selection
        self.train_epochs = train_epochs
        self.clip = grad_clip
        self._trainer = None

    @classmethod
    def from_modules(cls, nn, mcts, final_selection: str, **kw):
        """An agent around an already built policy module and search object (no configs, no trainer): `act()` only."""
        self = cls.__new__(cls)
        self.device = next(nn.parameters()).device
        self.nn, self.mcts, self.loss, self.optimizer = nn, mcts, None, None
        self.final_selection, self.train_epochs, self.clip, self._trainer = final_selection, 1, 0.0, None
        for k, v in kw.items():
            setattr(self, k, v)
        return self

    # ---- properties of the reference (agents.py:92-144) ----
    @property
    def action_dim(self) -> int:
        return 1

    @property
    def state_dim(self) -> int:
        return self.nn.state_dim

    @property
    def n_hidden_layers(self) -> int:
        return sum(1 for k in self.nn.state_dict() if k.startswith("trunk.") and k.endswith(".weight"))

    @property
    def n_hidden_units(self) -> int:
        return sum(v.shape[0] for k, v in self.nn.state_dict().items() if k.startswith("trunk.") and k.endswith(".weight"))

    @property
    def n_rollouts(self) -> int:
        return self.mcts.n_rollouts

    @property
    def learning_rate(self) -> float:
        return self.optimizer.param_groups[0]["lr"]

    @property
    def c_uct(self) -> float:
        return self.mcts.c_uct

    @property
    def gamma(self) -> float:
        return self.mcts.gamma

    def reset_mcts(self, root_state: np.ndarray) -> None:
        self.mcts.root_node = None
        self.mcts.root_state = root_state

    # ---- training (agents.py:157-184, :319-389, :539-603) ----
    def _get_trainer(self):
        if self._trainer is None:
            from ..train import Trainer
            self._trainer = Trainer(self.nn, self.loss.cfg, optimizer=self.optimizer, grad_clip=self.clip)
        return self._trainer

    def update(self, obs: Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray]) -> Dict[str, float]:
        """One gradient step on a replay batch (states, actions, counts, Qs, V) -- Qs is unused, as upstream."""
        states, actions, counts, _, V = obs
        t = lambda a: a if torch.is_tensor(a) else torch.from_numpy(np.asarray(a))  # noqa: E731
        out = self._get_trainer().update(dict(obs=t(states), actions=t(actions), counts=t(counts), V_target=t(V)))
        return {k: float(v) for k, v in out.items() if k != "alpha_loss"}

    def train(self, buffer) -> Dict[str, Any]:
        buffer.reshuffle()
        running_loss: Dict[str, Any] = defaultdict(float)
        for _ in range(self.train_epochs):
            for obs in buffer:
                if isinstance(obs, dict):  # selfplay.DeviceReplayBuffer rows
                    obs = (obs["obs"], obs["actions"], obs["counts"], obs["Q"], obs["V_target"])
                loss = self.update(obs)
                for key in loss.keys():
                    running_loss[key] += loss[key]
        return running_loss  # (upstream's division by the batch count is a no-op on a loop variable, agents.py:182-183)


class DiscreteAgent(Agent):
    def __init__(self, policy_cfg, mcts_cfg, loss_cfg, optimizer_cfg, final_selection: str, train_epochs: int, grad_clip: float,
                 temperature: float, device: str) -> None:
        super().__init__(policy_cfg=policy_cfg, loss_cfg=loss_cfg, mcts_cfg=mcts_cfg, optimizer_cfg=optimizer_cfg,
                         final_selection=final_selection, train_epochs=train_epochs, grad_clip=grad_clip, device=device)
        assert isinstance(self.mcts, MCTSDiscrete)
        self.temperature = temperature

    @classmethod
    def from_modules(cls, nn, mcts: MCTSDiscrete, final_selection: str = "max_visits", temperature: float = 1.0):
        assert isinstance(mcts, MCTSDiscrete)
        return super().from_modules(nn, mcts, final_selection, temperature=temperature)

    def act(self, Env, deterministic: bool = False) -> Tuple[Any, np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        self.mcts.search(Env=Env)
        state, actions, counts, Qs, V = self.mcts.return_results(self.final_selection)
        if self.final_selection == "max_value":
            pi = stable_normalizer(Qs, self.temperature)
        else:  # any other value means visit counts (DiscreteAgent.yaml:9 says "max_visits")
            pi = stable_normalizer(counts, self.temperature)
        action = pi.argmax() if deterministic else np.random.choice(len(pi), p=pi)
        return action, state, actions, counts, Qs, V

    def mcts_forward(self, action: int, node: np.ndarray) -> None:
        self.mcts.forward(action, node)


class ContinuousAgent(Agent):
    def __init__(self, policy_cfg, mcts_cfg, loss_cfg, optimizer_cfg, final_selection: str, epsilon: float, train_epochs: int,
                 grad_clip: float, device: str) -> None:
        super().__init__(policy_cfg=policy_cfg, loss_cfg=loss_cfg, mcts_cfg=mcts_cfg, optimizer_cfg=optimizer_cfg,
                         final_selection=final_selection, train_epochs=train_epochs, grad_clip=grad_clip, device=device)
        self.epsilon = epsilon

    @classmethod
    def from_modules(cls, nn, mcts: MCTSContinuous, final_selection: str = "max_visit", epsilon: float = 0.0):
        assert isinstance(mcts, MCTSContinuous)
        return super().from_modules(nn, mcts, final_selection, epsilon=epsilon)

    @property
    def action_limit(self) -> float:
        return self.nn.action_bound

    def epsilon_greedy(self, actions: np.ndarray, values: np.ndarray) -> np.ndarray:
        if random.random() < self.epsilon:
            return np.random.choice(actions)[np.newaxis]
        return actions[values.argmax()][np.newaxis]

    def act(self, Env) -> Tuple[Any, np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        self.mcts.search(Env=Env)
        state, actions, counts, Qs, V = self.mcts.return_results(self.final_selection)
        actions = np.atleast_1d(actions)
        values = Qs if self.final_selection == "max_value" else counts
        if self.epsilon == 0:
            action = actions[values.argmax()][np.newaxis]
        else:
            action = self.epsilon_greedy(actions=actions, values=values)
        return action, state, actions, counts, Qs, V
