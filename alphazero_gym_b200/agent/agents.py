"""`act()` over the CUDA search engine -- the caller side of the hot path.

Mirrors DiscreteAgent.act (reference agents.py:257-303), ContinuousAgent.act (:492-537), reset_mcts (:146-155),
mcts_forward (:305-317) and the read-only properties (:126-144).  The reference builds nn / mcts / loss /
optimizer with Hydra inside Agent.__init__ (agents.py:81-86); training is out of scope here, so these agents
take the already-built policy module and search object.  Return values, their order and dtypes follow the
reference: (action, state, actions, counts, Qs, V).
"""
from __future__ import annotations

import random
from typing import Any, Tuple

import numpy as np

from ..helpers import stable_normalizer
from ..search.mcts import MCTSContinuous, MCTSDiscrete


class Agent:
    def __init__(self, nn, mcts, final_selection: str):
        self.nn = nn
        self.mcts = mcts
        self.final_selection = final_selection

    @property
    def n_rollouts(self) -> int:
        return self.mcts.n_rollouts

    @property
    def c_uct(self) -> float:
        return self.mcts.c_uct

    @property
    def gamma(self) -> float:
        return self.mcts.gamma

    def reset_mcts(self, root_state: np.ndarray) -> None:
        self.mcts.root_node = None
        self.mcts.root_state = root_state


class DiscreteAgent(Agent):
    def __init__(self, nn, mcts: MCTSDiscrete, final_selection: str = "max_visits", temperature: float = 1.0):
        assert isinstance(mcts, MCTSDiscrete)
        super().__init__(nn, mcts, final_selection)
        self.temperature = temperature

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
    def __init__(self, nn, mcts: MCTSContinuous, final_selection: str = "max_visit", epsilon: float = 0.0):
        assert isinstance(mcts, MCTSContinuous)
        super().__init__(nn, mcts, final_selection)
        self.epsilon = epsilon

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
