"""Generate tests/golden/*.npz by running the unmodified reference search (see ref_harness.py).

Run in the build container only (needs /root/reference):   python -m oracle.gen_golden
Each file holds: the case config (JSON), flat weights (state_dict order), root states, the reference's
full tree in the oracle's table layout, root results (return_results), and RNG/noise draw counts.
"""
from __future__ import annotations

import json
import os
import sys
from dataclasses import asdict

import numpy as np

from . import azo, ref_harness as RH

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def cartpole_roots(B: int, seed: int = 34) -> np.ndarray:
    """gym CartPole reset: U(-0.05, 0.05)^4."""
    return np.random.default_rng(seed).uniform(-0.05, 0.05, size=(B, 4))


def pendulum_roots(B: int, seed: int = 34) -> np.ndarray:
    """gym Pendulum reset: th ~ U(-pi, pi), thdot ~ U(-1, 1)."""
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(-np.pi, np.pi, B), rng.uniform(-1.0, 1.0, B)], 1)


CASES = {
    # BASELINE config 1: run_discrete.yaml defaults
    "cartpole_n8_eps01": dict(cfg=azo.discrete_config(n_rollouts=8, epsilon=0.1), B=16),
    "cartpole_n8_eps0": dict(cfg=azo.discrete_config(n_rollouts=8, epsilon=0.0), B=16),
    # BASELINE config 3 shape (N=50) at a size the python reference finishes in seconds
    "cartpole_n50_eps01": dict(cfg=azo.discrete_config(n_rollouts=50, epsilon=0.1), B=16),
    "cartpole_n50_eps0": dict(cfg=azo.discrete_config(n_rollouts=50, epsilon=0.0), B=16),
    # deep trees that reach terminal states (pole falls): long search from near-threshold roots
    "cartpole_n200_terminal": dict(cfg=azo.discrete_config(n_rollouts=200, epsilon=0.1), B=6, roots="tilted"),
    "cartpole_n50_onpolicy_g099": dict(cfg=azo.discrete_config(n_rollouts=50, epsilon=0.1, gamma=0.99,
                                                               V_target_policy="on_policy"), B=8),
    # tree reuse quirk: forward() then search again (root.n carried over)
    "cartpole_n16_reuse": dict(cfg=azo.discrete_config(n_rollouts=16, epsilon=0.1), B=8, second_search=True),
    # BASELINE config 2: run_continuous.yaml defaults (GMM K=2)
    "pendulum_n25_k2": dict(cfg=azo.continuous_config(n_rollouts=25), B=12),
    "pendulum_n100_k2": dict(cfg=azo.continuous_config(n_rollouts=100), B=8),
    "pendulum_n200_k2": dict(cfg=azo.continuous_config(n_rollouts=200), B=4),
    # K=1: DiagonalNormalPolicy ("squashed Normal" of the north star)
    "pendulum_n25_k1": dict(cfg=azo.continuous_config(n_rollouts=25, num_components=1), B=12),
    "pendulum_n100_k1": dict(cfg=azo.continuous_config(n_rollouts=100, num_components=1), B=6),
    # non-default knobs: eps-greedy, gamma != 1 (f32 first backup product), c_pw > 1 (unexpanded root edge), on_policy
    "pendulum_n50_eps025_g097": dict(cfg=azo.continuous_config(n_rollouts=50, epsilon=0.25, gamma=0.97), B=8),
    "pendulum_n40_cpw2_onpolicy": dict(cfg=azo.continuous_config(n_rollouts=40, c_pw=2.0, kappa=0.4,
                                                                 V_target_policy="on_policy"), B=8),
    "pendulum_n30_k3": dict(cfg=azo.continuous_config(n_rollouts=30, num_components=3), B=6),
    # V_target_policy "greedy" (mcts.py:133-173; from a non-terminal root its loop never runs, so it is the root's Q.max())
    "cartpole_n8_greedy": dict(cfg=azo.discrete_config(n_rollouts=8, epsilon=0.1, V_target_policy="greedy"), B=8),
    "pendulum_n25_greedy": dict(cfg=azo.continuous_config(n_rollouts=25, V_target_policy="greedy"), B=6),
    # the rest of the reference's activation map (alphazero/network/utils.py:5-14): leakyrelu, relu6, swish / silu, hardswish
    "cartpole_n8_leakyrelu": dict(cfg=azo.discrete_config(n_rollouts=8, epsilon=0.1, activation=azo.ACT_LEAKYRELU), B=8),
    "cartpole_n8_hardswish": dict(cfg=azo.discrete_config(n_rollouts=8, epsilon=0.1, activation=azo.ACT_HARDSWISH), B=8),
    "pendulum_n25_relu6": dict(cfg=azo.continuous_config(n_rollouts=25, activation=azo.ACT_RELU6), B=6),
    "pendulum_n25_silu": dict(cfg=azo.continuous_config(n_rollouts=25, activation=azo.ACT_SILU), B=6),
    # reward wrappers of rl/wrappers.py around the searched env (rl/make_game.py:71-83 name suffixes -v0r / -v0s / -v0rs): the reference's
    # own wrapper classes in the reference's order; the tilted roots reach terminal states, where ReparametrizeWrapper pays -1
    "cartpole_n50_wrap_r": dict(cfg=azo.discrete_config(n_rollouts=50, epsilon=0.1), B=8, modify="r", roots="tilted"),
    "cartpole_n50_wrap_s_g099": dict(cfg=azo.discrete_config(n_rollouts=50, epsilon=0.1, gamma=0.99), B=8, modify="s"),
    "cartpole_n200_wrap_rs_terminal": dict(cfg=azo.discrete_config(n_rollouts=200, epsilon=0.1), B=6, modify="rs", roots="tilted"),
    # TRAINED weights: the reference agent after 80 real `update` steps (agents.py:319-389, :539-603; lr 3e-3 so that the activation
    # range moves well away from the default initialisation), then its search -- the case the per-row block exponent of the
    # tensor-core evaluation has to survive
    "cartpole_n50_trained": dict(cfg=azo.discrete_config(n_rollouts=50, epsilon=0.1), B=8, trained=80),
    "pendulum_n50_trained": dict(cfg=azo.continuous_config(n_rollouts=50), B=8, trained=80),
}


def _mt(cfg: azo.Config) -> azo.Config:
    cfg.rng_mode = azo.RNG_MT19937
    return cfg


# the discrete search of the reference with its STOCK random module (no shim), seeded random.seed(seed + tree)
MT_CASES = {
    "cartpole_mt_n8_eps01": dict(cfg=_mt(azo.discrete_config(n_rollouts=8, epsilon=0.1)), B=16),
    "cartpole_mt_n50_eps0": dict(cfg=_mt(azo.discrete_config(n_rollouts=50, epsilon=0.0)), B=12),
    "cartpole_mt_n200_terminal": dict(cfg=_mt(azo.discrete_config(n_rollouts=200, epsilon=0.1)), B=6, roots="tilted"),
    # ... and the continuous search with the stock `random` module AND torch's own global generator (torch.manual_seed(seed + tree)
    # before the search, torch.multinomial / torch.normal un-wrapped): K = 2 (run_continuous.yaml), K = 1 (the squashed Normal), K = 3
    "pendulum_mt_n25_k2": dict(cfg=_mt(azo.continuous_config(n_rollouts=25)), B=12),
    "pendulum_mt_n100_k2": dict(cfg=_mt(azo.continuous_config(n_rollouts=100)), B=6),
    "pendulum_mt_n50_k1_eps": dict(cfg=_mt(azo.continuous_config(n_rollouts=50, num_components=1, epsilon=0.25)), B=8),
    "pendulum_mt_n30_k3": dict(cfg=_mt(azo.continuous_config(n_rollouts=30, num_components=3)), B=6),
}
CASES_ALL = dict(CASES, **MT_CASES)


def roots_for(case) -> np.ndarray:
    cfg, B = case["cfg"], case["B"]
    if cfg.variant == azo.DISCRETE:
        r = cartpole_roots(B)
        if case.get("roots") == "tilted":
            r[:, 2] = np.linspace(0.12, 0.19, B)  # close to the 0.2095 rad threshold
            r[:, 3] = 0.8
        return r
    return pendulum_roots(B)


def trained_model(cfg: azo.Config, updates: int):
    """The reference's own agent (hydra / omegaconf stubbed at import time only, oracle/gen_train_golden.py) after `updates`
    gradient steps of its unmodified `update` on seeded synthetic replay batches; returns its policy network."""
    from . import gen_train_golden as GT
    import torch
    GT.install_stubs()
    from alphazero.agent import agents  # type: ignore
    torch.manual_seed(34)
    torch.set_num_threads(1)
    cont = cfg.variant == azo.CONTINUOUS
    loss = dict(_target_="alphazero.agent.losses.A0CLoss", tau=0.1, policy_coeff=1, alpha=1, value_coeff=1, reduction="mean")
    opt = dict(GT.RMSPROP, lr=0.003)
    if cont:
        agent = agents.ContinuousAgent(policy_cfg=GT.CONT_POLICY, mcts_cfg=GT.MCTS_C, loss_cfg=loss, optimizer_cfg=opt,
                                       final_selection="max_visit", epsilon=0, train_epochs=1, grad_clip=0.0, device="cpu")
    else:
        agent = agents.DiscreteAgent(policy_cfg=GT.DISC_POLICY, mcts_cfg=GT.MCTS_D, loss_cfg=loss, optimizer_cfg=opt,
                                     final_selection="max_visits", train_epochs=1, grad_clip=0.0, temperature=1.0, device="cpu")
    for s in range(updates):
        agent.update(tuple(a.copy() for a in GT.make_batch(cont, s)))
    return agent.nn


def generate(name: str) -> str:
    case = CASES_ALL[name]
    cfg: azo.Config = case["cfg"]
    cfg.math_mode = azo.MATH_LIBM
    model = trained_model(cfg, case["trained"]) if case.get("trained") else RH.make_model(cfg, weight_seed=34)
    roots = roots_for(case)
    if cfg.variant == azo.DISCRETE:
        modify = case.get("modify", "")
        if modify:  # the constants the engine / oracle are given: the wrappers' own Python expressions (search/mcts.py reward_model)
            from alphazero_gym_b200.search.mcts import reward_model
            cfg.reward_step, cfg.reward_terminal = reward_model(RH.wrap_like_make_game(RH.CartPoleEnv(roots[0]), modify), azo.DISCRETE)
        out = RH.run_discrete(cfg, model, roots, second_search=case.get("second_search", False), modify=modify)
    else:
        out = RH.run_continuous(cfg, model, roots)
    out["weights"] = azo.flatten_state_dict(model.state_dict())
    meta = asdict(cfg)
    meta.update(case=name, weight_seed=34, generator="oracle/gen_golden.py", reference="timoklein/alphazero-gym",
                numpy=np.__version__, torch=__import__("torch").__version__)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    return path


def load(name: str):
    """Load a golden file -> (Config, dict of arrays).  Usable without /root/reference."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    fields = azo.Config.__dataclass_fields__.keys()
    cfg = azo.Config(**{k: v for k, v in meta.items() if k in fields})
    return cfg, {k: z[k] for k in z.files if k != "meta"}


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES_ALL)
    for n in names:
        p = generate(n)
        print(n, "->", p, os.path.getsize(p), "bytes")
