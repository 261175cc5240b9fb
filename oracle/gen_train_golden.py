#!/usr/bin/env python
"""Golden vectors for the training step, produced by the UNMODIFIED reference.

    python oracle/gen_train_golden.py          # writes tests/golden/train_*.npz

Runs `ContinuousAgent.update` / `DiscreteAgent.update` (/root/reference/alphazero/agent/agents.py:319-389, :539-603) with the
reference's own policy (policies.py make_policy), losses (losses.py) and torch optimizers on seeded synthetic replay batches, three
consecutive steps each, on the CPU.  The reference needs `gym`, `hydra` and `omegaconf` only at import / construction time: gym is
stubbed as in ref_harness.py, hydra by a 10-line `_target_` resolver (utils.call / utils.instantiate), omegaconf's DictConfig by
dict.  Nothing under /root/reference is modified; this script is test infrastructure and runs in the build container only (the
committed .npz files travel).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def install_stubs():
    from oracle import ref_harness
    ref_harness._install_gym_stub()
    if "hydra" not in sys.modules:
        def _resolve(cfg, **kw):
            cfg = dict(cfg)
            mod, name = cfg.pop("_target_").rsplit(".", 1)
            cfg.update(kw)
            return getattr(importlib.import_module(mod), name)(**cfg)
        hydra = types.ModuleType("hydra")
        hydra.utils = types.SimpleNamespace(call=_resolve, instantiate=_resolve)
        sys.modules["hydra"] = hydra
        oc = types.ModuleType("omegaconf")
        dc = types.ModuleType("omegaconf.dictconfig")
        dc.DictConfig = dict
        oc.dictconfig = dc
        sys.modules["omegaconf"] = oc
        sys.modules["omegaconf.dictconfig"] = dc
    if REF not in sys.path:
        sys.path.insert(0, REF)


CONT_POLICY = dict(_target_="alphazero.network.policies.make_policy", representation_dim=3, action_dim=1, action_bound=2.0,
                   distribution="normal", num_components=2, hidden_dimensions=[128, 128, 128], nonlinearity="elu", layernorm=False,
                   log_param_min=-5, log_param_max=2)
DISC_POLICY = dict(_target_="alphazero.network.policies.make_policy", representation_dim=4, num_actions=2, action_dim=1,
                   distribution="discrete", hidden_dimensions=[128, 128], nonlinearity="relu", layernorm=False)
MCTS_C = dict(_target_="alphazero.search.mcts.MCTSContinuous", n_rollouts=25, c_uct=0.05, c_pw=1, kappa=0.5, gamma=1, epsilon=0,
              V_target_policy="off_policy", device="cpu", root_state=None)
MCTS_D = dict(_target_="alphazero.search.mcts.MCTSDiscrete", num_actions=2, n_rollouts=8, c_uct=1.5, gamma=1, epsilon=0.1,
              V_target_policy="off_policy", device="cpu", root_state=None)
RMSPROP = dict(_target_="torch.optim.RMSprop", lr=0.001, momentum=0, weight_decay=0, alpha=0.9, eps=1e-10)   # config/optimizer/RMSProp.yaml
ADAM = dict(_target_="torch.optim.Adam", lr=0.001, betas=[0.9, 0.99], weight_decay=0, eps=1e-07, amsgrad=False)  # config/optimizer/Adam.yaml

CASES = {
    # name: (agent class, policy, mcts, loss cfg, optimizer cfg, grad_clip, K override)
    "train_a0c_tuned_rmsprop": ("ContinuousAgent", CONT_POLICY, MCTS_C,
                                dict(_target_="alphazero.agent.losses.A0CLossTuned", action_dim=1, alpha_init=1, lr=0.001, tau=0.1,
                                     policy_coeff=0.1, value_coeff=1, reduction="mean", grad_clip=0, device="cpu"), RMSPROP, 0.0),
    "train_a0c_adam_clip": ("ContinuousAgent", CONT_POLICY, MCTS_C,
                            dict(_target_="alphazero.agent.losses.A0CLoss", tau=0.1, policy_coeff=1, alpha=1, value_coeff=1,
                                 reduction="mean"), ADAM, 0.5),
    "train_a0c_k1_sum": ("ContinuousAgent", dict(CONT_POLICY, num_components=1), MCTS_C,
                         dict(_target_="alphazero.agent.losses.A0CLoss", tau=0.5, policy_coeff=0.3, alpha=0.2, value_coeff=2,
                              reduction="sum"), ADAM, 0.0),
    # (DiscreteAgent + AlphaZeroLoss cannot run upstream: agents.py:381 hands the Categorical returned by DiscretePolicy.forward to
    # F.cross_entropy as logits -> TypeError; run_discrete.yaml ships A0CLoss, which is what is pinned here.)
    "train_a0c_discrete_rmsprop": ("DiscreteAgent", DISC_POLICY, MCTS_D,
                                   dict(_target_="alphazero.agent.losses.A0CLoss", tau=0.1, policy_coeff=1, alpha=1, value_coeff=1,
                                        reduction="mean"), RMSPROP, 0.0),
    "train_a0c_discrete_adam": ("DiscreteAgent", DISC_POLICY, MCTS_D,
                                dict(_target_="alphazero.agent.losses.A0CLoss", tau=0.1, policy_coeff=1, alpha=0.05, value_coeff=1,
                                     reduction="mean"), ADAM, 1.0),
}
STEPS = 3
BATCH = 32  # config/run_*.yaml buffer.batch_size


def make_batch(continuous: bool, step: int, cmax: int = 5):
    """A replay batch shaped like ReplayBuffer.sample (buffers.py:98-127) for the shipped configs."""
    rng = np.random.default_rng(1000 + step)
    if continuous:
        th = rng.uniform(-np.pi, np.pi, BATCH)
        states = np.stack([np.cos(th), np.sin(th), rng.uniform(-8, 8, BATCH)], 1).astype(np.float32)
        actions = rng.uniform(-1.999, 1.999, (BATCH, cmax)).astype(np.float32)   # sampled tanh-squashed actions
        counts = rng.multinomial(25, np.full(cmax, 1.0 / cmax), BATCH).astype(np.int64)
        counts[counts == 0] = 1  # every root child of a finished search has been visited (mcts.py:673-702)
    else:
        states = (rng.uniform(-1, 1, (BATCH, 4)) * np.array([2.4, 3, 0.21, 3])).astype(np.float64)
        actions = np.tile(np.arange(2), (BATCH, 1)).astype(np.int64)
        counts = rng.multinomial(8, [0.5, 0.5], BATCH).astype(np.int64)
    Q = rng.standard_normal(actions.shape)
    V = rng.uniform(-3, 0, BATCH) if continuous else rng.uniform(0, 30, BATCH)
    return states, actions, counts, Q, V


def flat(sd):
    keys = [k for k in sd.keys() if k.startswith("trunk.")] + ["value_head.weight", "value_head.bias", "dist_head.weight", "dist_head.bias"]
    return np.concatenate([sd[k].detach().cpu().numpy().astype(np.float32).ravel() for k in keys])


def run_case(name):
    agent_cls, policy, mcts, loss, opt, clip = CASES[name]
    from alphazero.agent import agents
    torch.manual_seed(34)
    torch.set_num_threads(1)
    continuous = agent_cls == "ContinuousAgent"
    extra = dict(epsilon=0) if continuous else dict(temperature=1.0)  # config/agent/*.yaml
    agent = getattr(agents, agent_cls)(policy_cfg=policy, mcts_cfg=mcts, loss_cfg=loss, optimizer_cfg=opt, final_selection="max_visit",
                                       train_epochs=1, grad_clip=clip, device="cpu", **extra)
    out = {"w0": flat(agent.nn.state_dict())}
    for s in range(STEPS):
        b = make_batch(continuous, s)
        for i, k in enumerate(("states", "actions", "counts", "Q", "V")):
            out[f"{k}_{s}"] = b[i].copy()
        info = agent.update(tuple(a.copy() for a in b))
        for k, v in info.items():
            out[f"info_{k}_{s}"] = np.float64(v)
        if s + 1 == STEPS:
            out[f"w_{s + 1}"] = flat(agent.nn.state_dict())
        if hasattr(agent.loss, "log_alpha"):
            out[f"log_alpha_{s + 1}"] = np.float64(agent.loss.log_alpha.detach())
    return out


def main():
    install_stubs()
    os.makedirs(OUT, exist_ok=True)
    for name in CASES:
        out = run_case(name)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, {k[5:]: float(v) for k, v in out.items() if k.startswith("info_") and k.endswith("_2")})


if __name__ == "__main__":
    main()
