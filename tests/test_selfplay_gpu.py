"""azg_selfplay_step (BASELINE config 5) against the CPU restatement of the reference's episode loop
(oracle/selfplay.py): every array of every step bit-identical -- replay rows, chosen actions, rewards, resets, the
discrete root-count carry-over -- for short episodes that exercise Env.reset and tree reuse."""
import dataclasses

import numpy as np
import pytest

from oracle import azo, selfplay as osp, gen_golden as G

KEYS = ("obs", "actions", "counts", "Q", "V_target", "n_children", "action_taken", "reward", "done")
CARRY = ("env_state", "ep_step", "episode", "root_n")


def _run_engine(cfg, w, states0, steps, max_len, tree_id0=0, **kw):
    import torch
    import enginelib as E
    from alphazero_gym_b200.selfplay import DeviceReplayBuffer, SelfPlayDriver
    B = states0.shape[0]
    eng = E.SearchEngine(E.engine_config(cfg, B))
    try:
        eng.set_weights(w)
        rb = DeviceReplayBuffer(max_size=3 * B, batch_size=16, obs_dim=cfg.state_dim, cmax=eng.cmax, device=eng.device)
        drv = SelfPlayDriver(eng, B, cfg.n_rollouts, max_episode_length=max_len, tree_id0=tree_id0, seed=cfg.seed, replay=rb, **kw)
        drv.set_states(states0)
        out = []
        for _ in range(steps):
            t = drv.step()
            eng.status()
            rec = {k: t[k].cpu().numpy().copy() for k in KEYS + ("ep_step", "episode", "root_n")}
            rec["env_state"] = drv.env_state.cpu().numpy().copy()
            out.append(rec)
        return out, rb, eng.cmax
    finally:
        eng.close()


def _compare(got, ref, cmax_ref):
    assert len(got) == len(ref)
    for s, (g, r) in enumerate(zip(got, ref)):
        for k in KEYS + CARRY:
            a, b = g[k], r[k]
            if a.ndim == 2 and b.ndim == 2 and a.shape[1] != b.shape[1]:
                m = min(a.shape[1], b.shape[1])
                assert not a[:, m:].any() and not b[:, m:].any()
                a, b = a[:, :m], b[:, :m]
            assert np.array_equal(a, b), f"step {s}: {k} differs"


@pytest.mark.gpu
@pytest.mark.parametrize("q8", [False, True])
def test_selfplay_pendulum_equals_oracle(q8):
    cfg, g = G.load("pendulum_n25_k2")
    cfg.math_mode, cfg.use_eval_tape = azo.MATH_DET, 0
    if q8:
        cfg.eval_mode = azo.EVAL_Q8
    B, steps, max_len = 48, 7, 3  # episodes of 3 steps: two resets inside the run
    states0 = G.pendulum_roots(B, seed=5)
    ref = osp.run(cfg, g["weights"], states0, steps, max_len, seed=cfg.seed, tree_id0=100)
    got, rb, cmax = _run_engine(cfg, g["weights"], states0, steps, max_len, tree_id0=100)
    _compare(got, ref, cfg.cmax)
    assert sum(int(r["done"].sum()) for r in ref) == 2 * B
    # the replay buffer holds the last 3 steps' rows (capacity 3 B), oldest overwritten first
    assert len(rb) == 3 * B


@pytest.mark.gpu
@pytest.mark.parametrize("deterministic", [False, True])
def test_selfplay_cartpole_equals_oracle(deterministic):
    cfg, g = G.load("cartpole_n50_eps01")
    cfg.math_mode, cfg.use_eval_tape = azo.MATH_DET, 0
    cfg = dataclasses.replace(cfg, n_rollouts=16)
    B, steps, max_len = 64, 9, 4  # tree reuse inside an episode (root count carried), reset every 4th step
    states0 = G.cartpole_roots(B, seed=5)
    ref = osp.run(cfg, g["weights"], states0, steps, max_len, seed=cfg.seed, tree_id0=7, deterministic=deterministic)
    got, rb, cmax = _run_engine(cfg, g["weights"], states0, steps, max_len, tree_id0=7, deterministic=deterministic)
    _compare(got, ref, cfg.cmax)
    assert any(r["root_n"].any() for r in ref), "the carry-over path was not exercised"
    if not deterministic:
        assert len({int(a) for r in ref for a in r["action_taken"]}) == 2


@pytest.mark.gpu
def test_selfplay_is_shard_invariant():
    """Environments 16..31 of a 32-environment run equal a 16-environment run started at tree_id0 = 16."""
    cfg, g = G.load("pendulum_n25_k2")
    cfg.math_mode, cfg.use_eval_tape = azo.MATH_DET, 0
    states0 = G.pendulum_roots(32, seed=9)
    full, _, _ = _run_engine(cfg, g["weights"], states0, 4, 2, tree_id0=0)
    half, _, _ = _run_engine(cfg, g["weights"], states0[16:], 4, 2, tree_id0=16)
    for f, h in zip(full, half):
        for k in KEYS + CARRY:
            assert np.array_equal(f[k][16:], h[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["selfplay_cartpole_t1", "selfplay_cartpole_t05", "selfplay_cartpole_t2", "selfplay_cartpole_value_t2",
                                  "selfplay_cartpole_det", "selfplay_pendulum_k2", "selfplay_pendulum_value"])
@pytest.mark.parametrize("q8", [False, True])
def test_selfplay_engine_vs_reference_agents(name, q8):
    """azg_selfplay_step against (1) the oracle loop, bit for bit, and (2) the goldens of the UNMODIFIED reference agents
    (DiscreteAgent.act / ContinuousAgent.act + mcts_forward / reset_mcts, oracle/gen_selfplay_golden.py): integers and the chosen
    discrete actions bit-exact, floating point within 1e-5.  Temperature 0.5 / 2, max_value and deterministic selection included --
    the cases that tell x ** temp (helpers.py:26) from x ** (1 / temp)."""
    from oracle import gen_selfplay_golden as GS
    import test_selfplay_golden as TG
    cfg, meta, g = GS.load(name)
    cfg.math_mode, cfg.use_eval_tape = azo.MATH_DET, 0
    if q8:
        cfg.eval_mode = azo.EVAL_Q8
    kw = dict(deterministic=meta["deterministic"], temperature=meta["temperature"])
    ref = osp.run(cfg, g["weights"], g["states0"], meta["steps"], meta["max_episode_length"], seed=cfg.seed, tree_id0=meta["tree_id0"],
                  by_value=meta["final_selection"] == "max_value", **kw)
    got, _, _ = _run_engine(cfg, g["weights"], g["states0"], meta["steps"], meta["max_episode_length"], tree_id0=meta["tree_id0"],
                            final_selection=meta["final_selection"], **kw)
    _compare(got, ref, cfg.cmax)
    TG.compare_with_golden(got, g, cfg.variant == azo.DISCRETE)
