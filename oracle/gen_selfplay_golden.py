#!/usr/bin/env python
"""Golden vectors for the self-play episode loop, produced by the UNMODIFIED reference agents.

    python oracle/gen_selfplay_golden.py          # writes tests/golden/selfplay_*.npz

What runs, unmodified, from /root/reference: `DiscreteAgent.act` / `ContinuousAgent.act` (alphazero/agent/agents.py:257-303,
:492-537) with `helpers.stable_normalizer` (helpers.py:9-27), `Agent.reset_mcts` (:146-155), `DiscreteAgent.mcts_forward`
(:305-317) -> `MCTSDiscrete.forward` (search/mcts.py:495-526), and underneath them the searches and the policy network.  The episode
loop around them is the one of run_discrete.py:95-122 / run_continuous.py:112-142 (act -> store -> Env.step -> reset or
mcts_forward / reset_mcts), written out here for B independent environments that keep going across episode ends (the shape of
azg_selfplay_step, include/azg.h).  hydra / omegaconf / gym are stubbed at import time only (oracle/gen_train_golden.py), the envs
are the restated gym dynamics of oracle/ref_harness.py.

Randomness is injected exactly as the engine and oracle/selfplay.py consume it: step s uses the Philox key
seed + s * 0x9E3779B97F4A7C15; stream 0 replaces Python `random` inside the search, stream 1 torch.multinomial / torch.normal inside
sample_action, stream 2 the uniform that `np.random.choice(len(pi), p=pi)` draws (numpy's legacy algorithm: cdf = cumsum(p) / cdf[-1],
searchsorted(cdf, u, side="right")), stream 3 `Env.reset` (gym: CartPole U(-0.05, 0.05)^4; Pendulum th ~ U(-pi, pi), thdot ~ U(-1, 1)).

Test infrastructure; runs in the build container only (needs /root/reference); the committed .npz files travel.
"""
from __future__ import annotations

import json
import os
import sys
from dataclasses import asdict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import azo, gen_golden as G, gen_train_golden as GT, ref_harness as RH, selfplay as osp  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name: (azo config, agent kwargs, B, steps, max_episode_length, tree_id0, deterministic)
CASES = {
    "selfplay_cartpole_t1": (azo.discrete_config(n_rollouts=16, epsilon=0.1), dict(final_selection="max_visits", temperature=1.0), 6, 9, 4, 7, False),
    "selfplay_cartpole_t05": (azo.discrete_config(n_rollouts=16, epsilon=0.1), dict(final_selection="max_visits", temperature=0.5), 6, 9, 4, 7, False),
    "selfplay_cartpole_t2": (azo.discrete_config(n_rollouts=16, epsilon=0.1), dict(final_selection="max_visits", temperature=2.0), 6, 9, 4, 7, False),
    "selfplay_cartpole_value_t2": (azo.discrete_config(n_rollouts=16, epsilon=0.1), dict(final_selection="max_value", temperature=2.0), 6, 9, 4, 7, False),
    "selfplay_cartpole_det": (azo.discrete_config(n_rollouts=16, epsilon=0.1), dict(final_selection="max_visits", temperature=1.0), 6, 9, 4, 7, True),
    "selfplay_pendulum_k2": (azo.continuous_config(n_rollouts=25), dict(final_selection="max_visit", epsilon=0), 4, 7, 3, 100, False),
    "selfplay_pendulum_value": (azo.continuous_config(n_rollouts=25), dict(final_selection="max_value", epsilon=0), 4, 7, 3, 100, False),
}


def _unit(x: int) -> float:
    return (float(x) + 0.5) * 2.3283064365386963e-10


class ChoiceInjector:
    """np.random.choice(len(pi), p=pi) with the uniform taken from stream 2 (numpy's legacy inverse-CDF algorithm)."""

    def __init__(self):
        self.key, self.tree, self._orig = 0, 0, None

    def choice(self, a, size=None, replace=True, p=None):
        assert isinstance(a, (int, np.integer)) and size is None and p is not None
        u = _unit(azo.rng_u32(self.key, self.tree, 2, 0, 0, 0))
        cdf = np.cumsum(np.asarray(p, np.float64))
        cdf /= cdf[-1]
        return int(np.searchsorted(cdf, u, side="right"))

    def __enter__(self):
        self._orig = np.random.choice
        np.random.choice = self.choice
        return self

    def __exit__(self, *a):
        np.random.choice = self._orig


def make_agent(cfg: azo.Config, agent_kw):
    from alphazero.agent import agents  # type: ignore
    torch.manual_seed(34)
    torch.set_num_threads(1)
    loss = dict(_target_="alphazero.agent.losses.A0CLoss", tau=0.1, policy_coeff=1, alpha=1, value_coeff=1, reduction="mean")
    if cfg.variant == azo.DISCRETE:
        mcts = dict(GT.MCTS_D, n_rollouts=cfg.n_rollouts, epsilon=cfg.epsilon, c_uct=cfg.c_uct)
        return agents.DiscreteAgent(policy_cfg=GT.DISC_POLICY, mcts_cfg=mcts, loss_cfg=loss, optimizer_cfg=GT.RMSPROP, train_epochs=1,
                                    grad_clip=0.0, device="cpu", **agent_kw)
    mcts = dict(GT.MCTS_C, n_rollouts=cfg.n_rollouts, epsilon=cfg.epsilon, c_uct=cfg.c_uct)
    return agents.ContinuousAgent(policy_cfg=GT.CONT_POLICY, mcts_cfg=mcts, loss_cfg=loss, optimizer_cfg=GT.RMSPROP, train_epochs=1,
                                  grad_clip=0.0, device="cpu", **agent_kw)


def run_case(name):
    cfg, agent_kw, B, steps, max_len, tree_id0, deterministic = CASES[name]
    cfg.math_mode = azo.MATH_LIBM
    M, _ = RH.reference_modules()
    import alphazero.helpers as H  # type: ignore
    import random as stock_random
    agent = make_agent(cfg, agent_kw)
    disc = cfg.variant == azo.DISCRETE
    S, cm = cfg.state_dim, cfg.cmax
    states0 = G.cartpole_roots(B, seed=5) if disc else G.pendulum_roots(B, seed=5)
    out = dict(states0=states0.copy(), obs=np.zeros((steps, B, S), np.float32), actions=np.zeros((steps, B, cm), np.float32),
               counts=np.zeros((steps, B, cm), np.int32), Q=np.zeros((steps, B, cm), np.float64), V_target=np.zeros((steps, B), np.float64),
               n_children=np.zeros((steps, B), np.int32), action_taken=np.zeros((steps, B), np.float32),
               reward=np.zeros((steps, B), np.float64), done=np.zeros((steps, B), np.int32),
               env_state=np.zeros((steps, B, 4 if disc else 2), np.float64), ep_step=np.zeros((steps, B), np.int32),
               episode=np.zeros((steps, B), np.int32), root_n=np.zeros((steps, B), np.int32))
    chooser = ChoiceInjector()
    for b in range(B):
        tree = tree_id0 + b
        env = RH.CartPoleEnv(states0[b]) if disc else RH.PendulumEnv(states0[b])
        agent.reset_mcts(root_state=np.array(env.state) if disc else env.obs())
        ep_step = episode = 0
        for s in range(steps):
            key = osp.step_seed(cfg.seed, s)
            H.random = M.random = RH.PhiloxRandom(key, tree)
            chooser.key, chooser.tree = key, tree
            try:
                with RH.NoiseInjector(key, tree, cfg.num_components), chooser:
                    if disc:
                        action, st, actions, counts, Qs, V = agent.act(Env=env, deterministic=deterministic)
                    else:
                        action, st, actions, counts, Qs, V = agent.act(Env=env)
            finally:
                H.random = M.random = stock_random
            C = len(counts)
            out["obs"][s, b] = np.asarray(st, np.float64).reshape(-1).astype(np.float32)
            out["actions"][s, b, :C] = np.asarray(actions, np.float32).reshape(-1)
            out["counts"][s, b, :C] = counts
            out["Q"][s, b, :C] = Qs
            out["V_target"][s, b] = V
            out["n_children"][s, b] = C
            state, r, terminal, _ = env.step(action)  # the true step (run_*.py: Env.step(action))
            ep_step += 1
            done = bool(terminal) or ep_step >= max_len  # run_*.py: `terminal or t == cfg.max_episode_length - 1`
            root_n = 0
            if done:
                episode += 1
                w = [azo.rng_u32(key, tree, 3, episode, 0, k) for k in range(4)]
                if disc:
                    env = RH.CartPoleEnv([-0.05 + (0.05 - -0.05) * _unit(w[k]) for k in range(4)])
                    agent.reset_mcts(root_state=np.array(env.state))
                else:
                    env = RH.PendulumEnv([-np.pi + (np.pi - -np.pi) * _unit(w[0]), -1.0 + (1.0 - -1.0) * _unit(w[1])])
                    agent.reset_mcts(root_state=env.obs())
                ep_step = 0
            elif disc:
                agent.mcts_forward(action, state)  # run_discrete.py:122
                root_n = 0 if agent.mcts.root_node is None else int(agent.mcts.root_node.n)
            else:
                agent.reset_mcts(root_state=state)  # run_continuous.py:142
            out["action_taken"][s, b] = float(np.asarray(action).reshape(-1)[0])
            out["reward"][s, b] = r
            out["done"][s, b] = int(done)
            out["env_state"][s, b] = np.asarray(env.state, np.float64)
            out["ep_step"][s, b], out["episode"][s, b], out["root_n"][s, b] = ep_step, episode, root_n
    out["weights"] = azo.flatten_state_dict(agent.nn.state_dict())
    meta = asdict(cfg)
    meta.update(case=name, B=B, steps=steps, max_episode_length=max_len, tree_id0=tree_id0, deterministic=bool(deterministic),
                final_selection=agent_kw["final_selection"], temperature=float(agent_kw.get("temperature", 1.0)),
                generator="oracle/gen_selfplay_golden.py", reference="timoklein/alphazero-gym", numpy=np.__version__, torch=torch.__version__)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return out


def load(name: str):
    """-> (Config, meta dict, arrays).  Usable without /root/reference."""
    z = np.load(os.path.join(OUT, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    fields = azo.Config.__dataclass_fields__.keys()
    return azo.Config(**{k: v for k, v in meta.items() if k in fields}), meta, {k: z[k] for k in z.files if k != "meta"}


def main():
    GT.install_stubs()
    os.makedirs(OUT, exist_ok=True)
    for name in sys.argv[1:] or list(CASES):
        out = run_case(name)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, "dones", int(out["done"].sum()), "actions", sorted({float(a) for a in out["action_taken"].ravel()})[:4],
              "root_n>0", int((out["root_n"] > 0).sum()))


if __name__ == "__main__":
    main()
