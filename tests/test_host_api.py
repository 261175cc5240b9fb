"""CPU-side checks of the boundary: the C-ABI library loads and exports everything include/azg.h declares,
fails loudly without a GPU, and the host mirror of the reference interface keeps the reference's names."""
import ctypes as C
import inspect
import os
import re

import numpy as np
import pytest
import torch

from alphazero_gym_b200 import _cabi
from alphazero_gym_b200.helpers import stable_normalizer
from alphazero_gym_b200.network import PolicyNet, describe_model, init_policy_weights, policy_head_dim, weights_version
from alphazero_gym_b200.parallel import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "azg.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(azg_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = _cabi.load()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/azg.h but not exported by libazg.so"
    assert sorted(_cabi.EXPORTS) == names, "ctypes EXPORTS list and header disagree"
    assert b"azg" in L.azg_version()


def test_config_struct_matches_header_layout():
    # field order and types of azg_config as declared in the header
    txt = open(os.path.join(ROOT, "include", "azg.h")).read()
    body = re.search(r"typedef struct azg_config \{(.*?)\} azg_config;", txt, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ty, names = decl.split(None, 1)
        fields += [(n.strip(), ty) for n in names.split(",")]
    cmap = {"int32_t": C.c_int32, "uint32_t": C.c_uint32, "uint64_t": C.c_uint64, "double": C.c_double, "float": C.c_float}
    got = [(n, t) for n, t in _cabi.AzgConfig._fields_]
    assert [n for n, _ in got] == [n for n, _ in fields]
    for (n, t), (_, ty) in zip(got, fields):
        assert t is cmap[ty], n


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from alphazero_gym_b200.engine import EngineConfig, SearchEngine
    with pytest.raises(RuntimeError):
        SearchEngine(EngineConfig())
    L = _cabi.load()
    h = C.c_void_p()
    rc = L.azg_create(C.byref(EngineConfig().c()), C.byref(h))
    assert rc == _cabi.AZG_ECUDA and b"no CPU fallback" in L.azg_last_error()


def test_reference_signatures_are_kept():
    from alphazero_gym_b200.search.mcts import MCTSContinuous, MCTSDiscrete
    from alphazero_gym_b200.agent.agents import ContinuousAgent, DiscreteAgent
    d = list(inspect.signature(MCTSDiscrete.__init__).parameters)[1:10]
    assert d == ["model", "num_actions", "n_rollouts", "c_uct", "gamma", "epsilon", "V_target_policy", "device", "root_state"]
    c = list(inspect.signature(MCTSContinuous.__init__).parameters)[1:11]
    assert c == ["model", "n_rollouts", "c_uct", "c_pw", "kappa", "gamma", "epsilon", "V_target_policy", "device", "root_state"]
    for cls in (MCTSDiscrete, MCTSContinuous):
        for m in ("search", "return_results", "search_batch"):
            assert callable(getattr(cls, m))
    assert callable(MCTSDiscrete.forward)
    assert list(inspect.signature(DiscreteAgent.act).parameters) == ["self", "Env", "deterministic"]
    assert list(inspect.signature(ContinuousAgent.act).parameters) == ["self", "Env"]
    for a in ("reset_mcts", "n_rollouts", "c_uct", "gamma"):
        assert hasattr(DiscreteAgent, a)


def test_describe_model_and_weight_order():
    net = PolicyNet(3, 128, 3, policy_head_dim("continuous", num_components=2), "elu", num_components=2, action_bound=2.0)
    d = describe_model(net)
    assert d == dict(state_dim=3, hidden=128, n_hidden=3, activation=1, head_dim=6, num_components=2, action_bound=2.0,
                     log_std_min=-5.0, log_std_max=2.0)
    flat = init_policy_weights(34, 3, 128, 3, 6)
    net.load_flat(flat)
    from alphazero_gym_b200.engine import flatten_state_dict
    assert np.array_equal(flatten_state_dict(net.state_dict()), flat)
    v0 = weights_version(net)
    with torch.no_grad():
        net.value_head.bias.add_(1.0)
    assert weights_version(net) != v0
    net.layernorm = True
    with pytest.raises(NotImplementedError):
        describe_model(net)


def test_env_state_extraction_and_unsupported_env():
    from alphazero_gym_b200._cabi import CONTINUOUS, DISCRETE
    from alphazero_gym_b200.search.mcts import env_hidden_state

    class CartPoleEnv:
        state = (0.01, 0.02, 0.03, 0.04)
        unwrapped = property(lambda self: self)

    class Wrapper:
        def __init__(self, env):
            self.env = env

    class MountainCarEnv:
        state = np.zeros(2)
        unwrapped = property(lambda self: self)

    assert np.array_equal(env_hidden_state(Wrapper(CartPoleEnv()), DISCRETE), [0.01, 0.02, 0.03, 0.04])
    with pytest.raises(TypeError):
        env_hidden_state(CartPoleEnv(), CONTINUOUS)
    with pytest.raises(TypeError):
        env_hidden_state(MountainCarEnv(), DISCRETE)
    with pytest.raises(TypeError):
        env_hidden_state(object(), DISCRETE)


def test_stable_normalizer_and_shard_range():
    assert np.allclose(stable_normalizer(np.array([1, 3, 4]), 1.0), [0.125, 0.375, 0.5])
    assert np.allclose(stable_normalizer(np.array([2.0, 2.0]), 0.5), [0.5, 0.5])
    for total, world in ((65536, 8), (10, 3), (7, 8)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_hydra_yaml_drop_ins_instantiate():
    """config/mcts/*.yaml = the reference's files with the engine classes as `_target_`: same keys, and
    `hydra.utils.instantiate(cfg, model=nn)` (agents.py:82) -- resolved here by hand, hydra is not installed -- builds the search
    objects without touching the GPU (the engine is created at the first search)."""
    import importlib
    import os
    import yaml
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref_keys = {"MCTSDiscrete": {"_target_", "num_actions", "n_rollouts", "gamma", "c_uct", "epsilon", "V_target_policy", "device", "root_state"},
                "MCTSContinuous": {"_target_", "n_rollouts", "c_pw", "kappa", "gamma", "c_uct", "epsilon", "V_target_policy", "device", "root_state"}}
    for name, keys in ref_keys.items():
        with open(os.path.join(root, "config", "mcts", name + ".yaml")) as f:
            cfg = yaml.safe_load(f)
        assert set(cfg) == keys  # the keys of /root/reference/config/mcts/<name>.yaml
        cfg = {k: (2 if v == "${policy.num_actions}" else "cuda:0" if v == "${device}" else v) for k, v in cfg.items()}
        mod, cls = cfg.pop("_target_").rsplit(".", 1)
        model = (PolicyNet(4, 128, 2, 2, "relu", num_actions=2) if name == "MCTSDiscrete"
                 else PolicyNet(3, 128, 3, 6, "elu", num_components=2, action_bound=2.0))
        obj = getattr(importlib.import_module(mod), cls)(model=model, **cfg)
        assert obj.n_rollouts == cfg["n_rollouts"] and obj.c_uct == cfg["c_uct"] and obj.gamma == cfg["gamma"]


def test_wrapped_envs_are_carried_or_refused_not_silently_unwrapped():
    """rl/make_game.py:71-83 wraps the env for the -v0n/r/p/s name suffixes (rl/wrappers.py): the search would see other rewards /
    observations than the engine's plain dynamics produce.  The two arithmetic reward wrappers become the engine's reward model,
    evaluated with the wrappers' own expressions; the others are refused; a plain TimeLimit-style wrapper is unwrapped."""
    import numpy as np
    import pytest
    from alphazero_gym_b200._cabi import CONTINUOUS, DISCRETE
    from alphazero_gym_b200.search.mcts import env_hidden_state, reward_model

    class CartPoleEnv:
        state = np.zeros(4)

    class PendulumEnv:
        state = np.zeros(2)

    class TimeLimit:
        def __init__(self, env):
            self.env = env

    class ReparametrizeWrapper(TimeLimit):
        pass

    class ScaleRewardWrapper(TimeLimit):
        pass

    class PILCOWrapper(TimeLimit):
        pass

    class NormalizeWrapper(TimeLimit):
        pass

    assert env_hidden_state(TimeLimit(CartPoleEnv()), DISCRETE).shape == (4,)
    assert reward_model(TimeLimit(CartPoleEnv()), DISCRETE) == (1.0, 1.0)
    assert env_hidden_state(ReparametrizeWrapper(CartPoleEnv()), DISCRETE).shape == (4,)
    assert reward_model(ReparametrizeWrapper(CartPoleEnv()), DISCRETE) == (0.005, -1.0)
    assert reward_model(ScaleRewardWrapper(CartPoleEnv()), DISCRETE) == (1.0 / 250.0, 1.0 / 250.0)
    # prepare_control_env order: Reparametrize inside, ScaleReward outside (rl/make_game.py:75-83)
    assert reward_model(ScaleRewardWrapper(ReparametrizeWrapper(CartPoleEnv())), DISCRETE) == (0.005 / 250.0, -1 / 250.0)
    assert reward_model(ReparametrizeWrapper(PendulumEnv()), CONTINUOUS) == (1.0, 1.0)
    with pytest.raises(NotImplementedError):
        reward_model(ScaleRewardWrapper(PendulumEnv()), CONTINUOUS)
    for bad in (PILCOWrapper, NormalizeWrapper):
        with pytest.raises(NotImplementedError):
            env_hidden_state(bad(CartPoleEnv()), DISCRETE)
        with pytest.raises(NotImplementedError):
            reward_model(ScaleRewardWrapper(bad(CartPoleEnv())), DISCRETE)
