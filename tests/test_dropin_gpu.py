"""The reference-facing classes (MCTSDiscrete / MCTSContinuous / agents) on the GPU, checked against the CPU
oracle and the reference's golden runs.  These read like the tests the reference lacks: build a model and an env,
call search(Env) / return_results / forward / act exactly as alphazero/agent/agents.py does."""
import numpy as np
import pytest
import torch

from oracle import azo, gen_golden as G
from oracle.ref_harness import CartPoleEnv, PendulumEnv
from parity import close

pytestmark = pytest.mark.gpu


def _model(cfg, weights):
    from alphazero_gym_b200.network import PolicyNet
    if cfg.variant == azo.DISCRETE:
        net = PolicyNet(4, cfg.hidden, cfg.n_hidden, cfg.num_actions, "relu", num_actions=cfg.num_actions)
    else:
        hd = 3 * cfg.num_components if cfg.num_components > 1 else 2
        net = PolicyNet(3, cfg.hidden, cfg.n_hidden, hd, "elu", num_components=cfg.num_components, action_bound=cfg.action_bound)
    return net.load_flat(weights)


@pytest.mark.parametrize("eval_q8", [None, False])  # None: the default = tensor-core evaluation (what bench.py times); False: FP32 FMA
@pytest.mark.parametrize("name", ["cartpole_n8_eps01", "cartpole_n50_eps0", "cartpole_n50_onpolicy_g099", "cartpole_n50_trained"])
def test_mcts_discrete_dropin(name, eval_q8):
    from alphazero_gym_b200.search.mcts import MCTSDiscrete
    cfg, g = G.load(name)
    cfg.eval_mode = azo.EVAL_FP32 if eval_q8 is False else azo.EVAL_Q8
    model = _model(cfg, g["weights"])
    env = CartPoleEnv(g["root_state"][0])
    mcts = MCTSDiscrete(model=model, num_actions=2, n_rollouts=cfg.n_rollouts, c_uct=cfg.c_uct, gamma=cfg.gamma, epsilon=cfg.epsilon,
                        V_target_policy=cfg.V_target_policy, device="cuda:0", root_state=np.array(env.state), seed=cfg.seed,
                        **({} if eval_q8 is None else {"eval_q8": eval_q8}))
    assert mcts._engine(1).cfg.is_q8() == (eval_q8 is None)
    mcts.search(Env=env)
    state, actions, counts, Q, V = mcts.return_results("max_visit")
    # first search of the object uses tree id 0 == tree 0 of the golden batch
    assert np.array_equal(state, g["root_state"][0])
    assert actions.tolist() == [0, 1] and counts.dtype == np.int64 and Q.dtype == np.float64
    assert np.array_equal(counts, g["counts"][0]) and counts.sum() == cfg.n_rollouts
    assert close(Q, g["Q"][0]) and close(V, g["V_target"][0])
    cfg.math_mode = azo.MATH_DET
    ref = azo.search(cfg, g["weights"], g["root_state"][:1])
    assert np.array_equal(Q, ref["Q"][0]) and V == ref["V_target"][0]
    mcts.close()


class _Wrapper:  # gym.Wrapper stand-in: the drop-in recognises the reference's wrapper classes by name (rl/wrappers.py)
    def __init__(self, env):
        self.env = env


class ReparametrizeWrapper(_Wrapper):
    pass


class ScaleRewardWrapper(_Wrapper):
    pass


@pytest.mark.parametrize("name, wrap", [("cartpole_n50_wrap_r", lambda e: ReparametrizeWrapper(e)),
                                        ("cartpole_n50_wrap_s_g099", lambda e: ScaleRewardWrapper(e)),
                                        ("cartpole_n200_wrap_rs_terminal", lambda e: ScaleRewardWrapper(ReparametrizeWrapper(e)))])
def test_mcts_discrete_dropin_with_reward_wrappers(name, wrap):
    """Env wrapped like rl/make_game.py does for the -v0r / -v0s / -v0rs names: the search through the drop-in class equals the
    unmodified reference's search through its own wrapper classes (golden) and the oracle bit for bit; then the plain env again."""
    from alphazero_gym_b200.search.mcts import MCTSDiscrete
    cfg, g = G.load(name)
    cfg.eval_mode = azo.EVAL_Q8
    model = _model(cfg, g["weights"])
    base = CartPoleEnv(g["root_state"][0])
    mcts = MCTSDiscrete(model=model, num_actions=2, n_rollouts=cfg.n_rollouts, c_uct=cfg.c_uct, gamma=cfg.gamma, epsilon=cfg.epsilon,
                        V_target_policy=cfg.V_target_policy, device="cuda:0", root_state=np.array(base.state), seed=cfg.seed)
    mcts.search(Env=wrap(base))
    _, _, counts, Q, V = mcts.return_results("max_visit")
    assert np.array_equal(counts, g["counts"][0]) and close(Q, g["Q"][0]) and close(V, g["V_target"][0])
    cfg.math_mode = azo.MATH_DET
    ref = azo.search(cfg, g["weights"], g["root_state"][:1])
    assert np.array_equal(counts, ref["counts"][0]) and np.array_equal(Q, ref["Q"][0]) and V == ref["V_target"][0]
    # the same object on the plain env: the reward model goes back to 1.0 / 1.0 (second search of the object = tree id 1)
    mcts.root_node = None
    mcts.search(Env=base)
    _, _, counts, Q, V = mcts.return_results("max_visit")
    cfg.reward_step = cfg.reward_terminal = 1.0
    ref = azo.search(cfg, g["weights"], g["root_state"][:1], tree_id0=1)
    assert np.array_equal(counts, ref["counts"][0]) and np.array_equal(Q, ref["Q"][0]) and V == ref["V_target"][0]
    mcts.close()


def test_mcts_discrete_forward_reuses_root_count():
    """forward() keeps only root.n; the next search equals the oracle run with that root_n_init."""
    from alphazero_gym_b200.search.mcts import MCTSDiscrete
    cfg, g = G.load("cartpole_n16_reuse")
    cfg.eval_mode = azo.EVAL_Q8  # the drop-in classes run the tensor-core evaluation by default
    model = _model(cfg, g["weights"])
    roots = G.cartpole_roots(8)
    env = CartPoleEnv(roots[0])
    mcts = MCTSDiscrete(model=model, num_actions=2, n_rollouts=16, c_uct=1.5, gamma=1, epsilon=0.1, V_target_policy="off_policy",
                        device="cuda:0", root_state=np.array(env.state), seed=cfg.seed)
    mcts.search(env)
    _, _, counts, _, _ = mcts.return_results("max_visits")
    a = int(np.argmax(counts))
    obs, _, done, _ = env.step(a)
    mcts.forward(a, obs)
    assert mcts.root_node is not None and mcts.root_node.n == g["root_n_init"][0]
    mcts.search(env)
    _, _, counts2, Q2, _ = mcts.return_results("max_visits")
    cfg.math_mode = azo.MATH_DET
    ref = azo.search(cfg, g["weights"], np.array(env.state)[None], np.array([mcts.root_node.n - 16], np.int32), tree_id0=1)
    assert np.array_equal(counts2, ref["counts"][0]) and np.array_equal(Q2, ref["Q"][0])
    # a state that does not match the child's resets the tree (mcts.py:513-524)
    mcts.forward(int(np.argmax(counts2)), obs + 1.0)
    assert mcts.root_node is None
    mcts.close()


@pytest.mark.parametrize("name", ["pendulum_n25_k2", "pendulum_n100_k1", "pendulum_n40_cpw2_onpolicy"])
def test_mcts_continuous_dropin(name):
    from alphazero_gym_b200.search.mcts import MCTSContinuous
    cfg, g = G.load(name)
    model = _model(cfg, g["weights"])
    env = PendulumEnv(g["root_state"][0])
    mcts = MCTSContinuous(model=model, n_rollouts=cfg.n_rollouts, c_uct=cfg.c_uct, c_pw=cfg.c_pw, kappa=cfg.kappa, gamma=cfg.gamma,
                          epsilon=cfg.epsilon, V_target_policy=cfg.V_target_policy, device="cuda:0", root_state=env.obs(), seed=cfg.seed)
    mcts.search(Env=env)
    state, actions, counts, Q, V = mcts.return_results("max_visit")
    C = int(g["n_children"][0])
    assert np.array_equal(state, env.obs()) and actions.shape == (C,) and actions.dtype == np.float32
    assert np.array_equal(counts, g["counts"][0, :C])
    assert close(actions, g["actions"][0, :C]) and close(Q, g["Q"][0, :C]) and close(V, g["V_target"][0])
    # legacy upstream call form search(n_mcts, c, Env, mcts_env)
    mcts2 = MCTSContinuous(model=model, n_rollouts=cfg.n_rollouts, c_uct=cfg.c_uct, c_pw=cfg.c_pw, kappa=cfg.kappa, gamma=cfg.gamma,
                           epsilon=cfg.epsilon, V_target_policy=cfg.V_target_policy, device="cuda:0", root_state=env.obs(), seed=cfg.seed)
    mcts2.search(cfg.n_rollouts, cfg.c_uct, env, None)
    assert np.array_equal(mcts2.return_results("max_visit")[2], counts)
    mcts.close(); mcts2.close()


def test_agents_act_and_weight_refresh():
    from alphazero_gym_b200.agent.agents import ContinuousAgent, DiscreteAgent
    from alphazero_gym_b200.search.mcts import MCTSContinuous, MCTSDiscrete
    cfg, g = G.load("pendulum_n25_k2")
    cfg.eval_mode = azo.EVAL_Q8
    model = _model(cfg, g["weights"])
    env = PendulumEnv(g["root_state"][1])
    agent = ContinuousAgent.from_modules(model, MCTSContinuous(model=model, n_rollouts=25, c_uct=0.05, c_pw=1, kappa=0.5, gamma=1, epsilon=0,
                                                  V_target_policy="off_policy", device="cuda:0", root_state=None, seed=cfg.seed))
    agent.reset_mcts(root_state=env.obs())
    action, state, actions, counts, Qs, V = agent.act(env)
    assert action.shape == (1,) and action[0] == actions[counts.argmax()] and counts.sum() == 25 and V == Qs.max()
    assert agent.n_rollouts == 25 and agent.c_uct == 0.05 and agent.gamma == 1
    # training changes the weights in place; the next search must see them
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(0.5)
    agent.reset_mcts(root_state=env.obs())
    _, _, _, _, Qs2, _ = agent.act(env)
    cfg.math_mode = azo.MATH_DET
    ref = azo.search(cfg, azo.flatten_state_dict(model.state_dict()), g["root_state"][1:2], tree_id0=1)
    assert np.array_equal(Qs2, ref["Q"][0, : len(Qs2)])
    agent.mcts.close()

    dcfg, dg = G.load("cartpole_n8_eps01")
    dmodel = _model(dcfg, dg["weights"])
    denv = CartPoleEnv(dg["root_state"][0])
    dagent = DiscreteAgent.from_modules(dmodel, MCTSDiscrete(model=dmodel, num_actions=2, n_rollouts=8, c_uct=1.5, gamma=1, epsilon=0.1,
                                                V_target_policy="off_policy", device="cuda:0", root_state=np.array(denv.state)))
    np.random.seed(0)
    for _ in range(5):  # the run_discrete.py loop shape: act -> env.step -> mcts_forward
        a, s, acts, counts, Qs, V = dagent.act(denv)
        assert a in (0, 1) and counts.sum() == 8
        obs, r, done, _ = denv.step(int(a))
        if done:
            break
        dagent.mcts_forward(int(a), obs)
    dagent.mcts.close()


def test_search_batch_and_unsupported_env():
    from alphazero_gym_b200.search.mcts import MCTSContinuous
    cfg, g = G.load("pendulum_n25_k2")
    model = _model(cfg, g["weights"])
    mcts = MCTSContinuous(model=model, n_rollouts=25, c_uct=0.05, c_pw=1, kappa=0.5, gamma=1, epsilon=0, V_target_policy="off_policy",
                          device="cuda:0", root_state=None, seed=cfg.seed)
    res = mcts.search_batch(g["root_state"])
    assert np.array_equal(res["counts"][:, : g["counts"].shape[1]], g["counts"])
    with pytest.raises(TypeError):
        mcts.search(CartPoleEnv(np.zeros(4)))
    mcts.close()


def test_raw_ctypes_binding_from_integration_md():
    """The stand-alone ctypes stub shown in INTEGRATION.md section 3 (no alphazero_gym_b200 Python involved)."""
    import ctypes as C
    import os
    cfg, g = G.load("pendulum_n25_k2")

    class AzgConfig(C.Structure):
        _fields_ = [(n, C.c_int32) for n in ("variant", "max_rollouts", "max_trees", "num_actions", "num_components",
                                             "state_dim", "hidden", "n_hidden", "activation", "v_target", "puct_f32", "device")] + \
                   [(n, C.c_double) for n in ("c_uct", "gamma", "epsilon", "c_pw", "kappa")] + \
                   [(n, C.c_float) for n in ("action_bound", "log_std_min", "log_std_max")] + \
                   [("flags", C.c_uint32), ("seed", C.c_uint64)]

    root_dir = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = C.CDLL(os.path.join(root_dir, "alphazero_gym_b200", "lib", "libazg.so"))
    lib.azg_last_error.restype = C.c_char_p
    c = AzgConfig(1, 25, 1, 2, 2, 3, 128, 3, 1, 0, 1, 0, 0.05, 1.0, 0.0, 1.0, 0.5, 2.0, -5.0, 2.0, 0, 34)
    eng = C.c_void_p()
    assert lib.azg_create(C.byref(c), C.byref(eng)) == 0, lib.azg_last_error()
    w = g["weights"].astype(np.float32)
    assert lib.azg_set_weights(eng, w.ctypes.data_as(C.c_void_p), C.c_int64(w.size), None) == 0
    cmax = lib.azg_cmax(eng)
    root = np.ascontiguousarray(g["root_state"][:1], np.float64)
    actions, counts = np.empty((1, cmax), np.float32), np.empty((1, cmax), np.int32)
    Q, Vt, nc = np.empty((1, cmax)), np.empty(1), np.empty(1, np.int32)
    rc = lib.azg_search_host(eng, 1, root.ctypes.data_as(C.c_void_p), None, 25, C.c_int64(0), actions.ctypes.data_as(C.c_void_p),
                             counts.ctypes.data_as(C.c_void_p), Q.ctypes.data_as(C.c_void_p), Vt.ctypes.data_as(C.c_void_p),
                             nc.ctypes.data_as(C.c_void_p))
    assert rc == 0, lib.azg_last_error()
    n = int(nc[0])
    assert np.array_equal(counts[0, :n], g["counts"][0, :n]) and close(Q[0, :n], g["Q"][0, :n])
    lib.azg_destroy(eng)
