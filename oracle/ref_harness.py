"""Runs the UNMODIFIED reference search (/root/reference/alphazero/search/mcts.py) in this container.

TEST INFRASTRUCTURE ONLY, and only usable where /root/reference exists (the build container):
it is what oracle/gen_golden.py uses to produce tests/golden/*.npz, and what tests marked
`needs_reference` use to cross-check the C oracle live.  Nothing on the GPU box imports this.

How (SURVEY.md section 7 "Oracle recipe", section 8c):
  * `gym` is not installed -> a ~20 line stub module is registered in sys.modules (the reference
    only needs gym.Env / gym.Wrapper / gym.spaces.{Box,Discrete} at import time).
  * CartPole-v0 / Pendulum-v0 dynamics are restated from gym 0.17.2/0.19.0 classic_control
    (third-party, not vendored in the reference).  Squares are written x*x (gym writes x**2; libm
    pow(x,2) is not always the correctly rounded square) and the f32 torque is promoted to f64
    before any arithmetic (numpy 1.x semantics of the reference's pinned numpy).
  * RNG is injected, not emulated: `random` inside alphazero.helpers / alphazero.search.mcts is
    replaced by a shim drawing from the oracle's Philox stream; torch.multinomial / torch.normal
    are wrapped so the squashed-GMM sample consumes the oracle's noise stream.
"""
from __future__ import annotations

import copy
import math
import sys
import types
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from . import azo

REFERENCE_ROOT = "/root/reference"
_M = None
_P = None


def _install_gym_stub() -> None:
    if "gym" in sys.modules:
        return
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class Env:  # noqa: D401
        pass

    class Wrapper(Env):
        def __init__(self, env=None):
            self.env = env

    class Box:
        def __init__(self, low=None, high=None, shape=None, dtype=None):
            self.low, self.high, self.shape = low, high, shape

    class Discrete:
        def __init__(self, n):
            self.n = n

    gym.Env, gym.Wrapper, gym.spaces = Env, Wrapper, spaces
    spaces.Box, spaces.Discrete = Box, Discrete
    sys.modules["gym"] = gym
    sys.modules["gym.spaces"] = spaces


def reference_modules():
    """Import the reference's search and policy modules, unmodified."""
    global _M, _P
    if _M is None:
        _install_gym_stub()
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        import alphazero.network.policies as P  # type: ignore
        import alphazero.search.mcts as M  # type: ignore
        _M, _P = M, P
    return _M, _P


# ---------------------------------------------------------------------------------------------
# Environments (gym classic_control restated; hidden state in `.state`, deepcopy-able)
# ---------------------------------------------------------------------------------------------
class CartPoleEnv:
    gravity, masscart, masspole, length, force_mag, tau = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    total_mass = masspole + masscart
    polemass_length = masspole * length
    theta_threshold_radians = 12 * 2 * math.pi / 360
    x_threshold = 2.4

    class _Spec:
        id = "CartPole-v0"

    spec = _Spec()  # rl/wrappers.py get_name() reads env.spec.id

    def __init__(self, state):
        self.state = tuple(float(v) for v in state)

    @property
    def unwrapped(self):
        return self

    def step(self, action):
        x, x_dot, theta, theta_dot = self.state
        force = self.force_mag if action == 1 else -self.force_mag
        costheta = math.cos(theta)
        sintheta = math.sin(theta)
        temp = (force + self.polemass_length * (theta_dot * theta_dot) * sintheta) / self.total_mass
        thetaacc = (self.gravity * sintheta - costheta * temp) / (
            self.length * (4.0 / 3.0 - self.masspole * (costheta * costheta) / self.total_mass))
        xacc = temp - self.polemass_length * thetaacc * costheta / self.total_mass
        x = x + self.tau * x_dot
        x_dot = x_dot + self.tau * xacc
        theta = theta + self.tau * theta_dot
        theta_dot = theta_dot + self.tau * thetaacc
        self.state = (x, x_dot, theta, theta_dot)
        done = bool(x < -self.x_threshold or x > self.x_threshold
                    or theta < -self.theta_threshold_radians or theta > self.theta_threshold_radians)
        return np.array(self.state, dtype=np.float64), 1.0, done, {}


class PendulumEnv:
    max_speed, max_torque, dt, g, m, l = 8, 2.0, 0.05, 10.0, 1.0, 1.0

    def __init__(self, state):
        self.state = np.array(state, dtype=np.float64)

    @property
    def unwrapped(self):
        return self

    def obs(self):
        th, thdot = float(self.state[0]), float(self.state[1])
        return np.array([math.cos(th), math.sin(th), thdot], dtype=np.float64)

    def step(self, u):
        th, thdot = float(self.state[0]), float(self.state[1])
        g, m, l, dt = self.g, self.m, self.l, self.dt
        uf = np.clip(np.asarray(u, dtype=np.float32), np.float32(-self.max_torque), np.float32(self.max_torque)).ravel()[0]
        u64 = float(uf)  # numpy-1.x promotion: the f32 torque enters f64 arithmetic exactly
        an = ((th + math.pi) % (2 * math.pi)) - math.pi
        costs = an * an + 0.1 * (thdot * thdot) + 0.001 * (u64 * u64)
        newthdot = thdot + (-3 * g / (2 * l) * math.sin(th + math.pi) + 3.0 / (m * (l * l)) * u64) * dt
        newth = th + newthdot * dt
        newthdot = min(max(newthdot, -float(self.max_speed)), float(self.max_speed))
        self.state = np.array([newth, newthdot], dtype=np.float64)
        return self.obs(), np.float64(-costs), False, {}


# ---------------------------------------------------------------------------------------------
# RNG / noise injection
# ---------------------------------------------------------------------------------------------
class PhiloxRandom:
    """Drop-in for the `random` module inside the reference (stream 0 of the oracle's Philox)."""

    def __init__(self, seed: int, tree: int):
        self.seed, self.tree, self.draws = seed, tree, 0

    def _u32(self) -> int:
        x = azo.rng_u32(self.seed, self.tree, 0, self.draws)
        self.draws += 1
        return x

    def random(self) -> float:
        return float(self._u32() >> 8) * 2.0 ** -24

    def choice(self, seq):
        return seq[(self._u32() * len(seq)) >> 32]

    def randint(self, a: int, b: int) -> int:
        return a + ((self._u32() * (b - a + 1)) >> 32)


class StockRandom(__import__("random").Random):
    """CPython's own generator, i.e. the UN-SHIMMED reference: `random.Random(s)` runs exactly what the module-level
    `random.seed(s)` / `random.random` / `random.choice` / `random.randint` run.  The two overrides only count generator outputs
    (random() takes two, getrandbits(k <= 32) one); defining getrandbits keeps `_randbelow_with_getrandbits`, the stock algorithm."""

    def __init__(self, seed: int):
        super().__init__(seed)
        self.draws = 0

    def random(self) -> float:
        self.draws += 2
        return super().random()

    def getrandbits(self, k: int) -> int:
        self.draws += 1
        return super().getrandbits(k)


class NoiseInjector:
    """Wraps torch.multinomial / torch.normal (stream 1): one index per sample_action call."""

    def __init__(self, seed: int, tree: int, K: int):
        self.seed, self.tree, self.K, self.j = seed, tree, K, 0
        self._orig = None

    def multinomial(self, probs, num_samples, replacement=False, **kw):
        u, _ = azo.noise(self.seed, self.tree, self.j, self.K)
        p = probs.detach().numpy().astype(np.float32).reshape(-1)
        k, cum = 0, np.float32(p[0])
        while k < len(p) - 1 and not (np.float32(u) < cum):
            k += 1
            cum = np.float32(cum + p[k])
        return torch.tensor([[k]], dtype=torch.long)

    def normal(self, mean, std, **kw):
        _, z = azo.noise(self.seed, self.tree, self.j, self.K)
        self.j += 1
        zt = torch.from_numpy(z[: mean.numel()].copy()).reshape(mean.shape)
        return zt * std + mean

    def __enter__(self):
        self._orig = (torch.multinomial, torch.normal)
        torch.multinomial, torch.normal = self.multinomial, self.normal
        return self

    def __exit__(self, *a):
        torch.multinomial, torch.normal = self._orig


def make_model(cfg: azo.Config, weight_seed: int = 34):
    """torch.manual_seed(weight_seed); unmodified make_policy (policies.py:806)."""
    _, P = reference_modules()
    torch.manual_seed(weight_seed)
    hidden = [cfg.hidden] * cfg.n_hidden
    act = azo.ACT_NAMES[cfg.activation]
    if cfg.variant == azo.DISCRETE:
        return P.make_policy(cfg.state_dim, 1, "discrete", hidden, act, num_actions=cfg.num_actions)
    return P.make_policy(cfg.state_dim, 1, "normal", hidden, act, num_components=cfg.num_components,
                         action_bound=cfg.action_bound, log_param_min=cfg.log_std_min, log_param_max=cfg.log_std_max)


def _gamma_arg(g: float):
    return int(g) if float(g).is_integer() else g  # config/mcts/*.yaml: `gamma: 1` is a YAML int


# ---------------------------------------------------------------------------------------------
# Running the reference + dumping its pointer-linked tree into the oracle's table layout
# ---------------------------------------------------------------------------------------------
def wrap_like_make_game(env, modify: str):
    """The reference's OWN reward wrappers (rl/wrappers.py, imported unmodified) around `env`, in the order of
    rl/make_game.py:71-83 prepare_control_env: 'r' ReparametrizeWrapper inside, 's' ScaleRewardWrapper outside."""
    if not modify:
        return env
    reference_modules()
    import rl.wrappers as W  # type: ignore
    assert set(modify) <= {"r", "s"}, "only the arithmetic reward wrappers are restated by the oracle"
    if "r" in modify:
        env = W.ReparametrizeWrapper(env)
    if "s" in modify:
        env = W.ScaleRewardWrapper(env)
    return env


def run_discrete(cfg: azo.Config, model, root_states: np.ndarray, tree_id0: int = 0,
                 second_search: bool = False, modify: str = "") -> Dict[str, np.ndarray]:
    """MCTSDiscrete.search per tree.  second_search=True additionally does act-like `forward` on the
    most visited root action and searches again (the root.n carry-over quirk); the dump is
    then of the SECOND search and `root_state`/`root_n_init` describe its root."""
    M, _ = reference_modules()
    B, R, A = len(root_states), cfg.rows, cfg.num_actions
    out = _alloc_discrete(B, R, A, cfg.cmax)
    out["root_state"] = np.zeros((B, 4), np.float64)
    out["root_n_init"] = np.zeros(B, np.int32)
    out["draws"] = np.zeros(B, np.int64)
    import alphazero.helpers as H  # type: ignore
    orig_expansion = M.MCTS.expansion
    for b in range(B):
        # rng_mode MT19937: the stock generator seeded like `random.seed(seed + tree)` right before the search
        rng = StockRandom(cfg.seed + tree_id0 + b) if cfg.rng_mode == azo.RNG_MT19937 else PhiloxRandom(cfg.seed, tree_id0 + b)
        H.random = M.random = rng
        counter = [0]

        def expansion(action, state, reward, terminal, _c=counter):
            node = orig_expansion(action, state, reward, terminal)
            _c[0] += 1
            node._cid = _c[0]
            return node

        M.MCTS.expansion = staticmethod(expansion)
        try:
            base_env = CartPoleEnv(root_states[b])
            env = wrap_like_make_game(base_env, modify)
            mcts = M.MCTSDiscrete(model=model, num_actions=A, n_rollouts=cfg.n_rollouts, c_uct=cfg.c_uct,
                                  gamma=_gamma_arg(cfg.gamma), epsilon=cfg.epsilon,
                                  V_target_policy=cfg.V_target_policy, device="cpu",
                                  root_state=np.array(base_env.state))
            mcts.search(env)
            root_n_init = 0
            if second_search:
                _, _, counts, _, _ = mcts.return_results("max_visit")
                a = int(np.argmax(counts))
                obs, _, done, _ = env.step(a)
                assert not done
                mcts.forward(a, obs)
                assert mcts.root_node is not None, "forward() reset the tree; pick another seed"
                root_n_init = mcts.root_node.n
                rng.draws = 0  # the engine restarts the per-search draw counter
                counter[0] = 0
                mcts.root_node._cid = 0
                mcts.search(env)
            mcts.root_node._cid = 0
            out["root_state"][b] = base_env.state
            out["root_n_init"][b] = root_n_init
            out["draws"][b] = rng.draws
            _dump_discrete(mcts, b, out, A)
            s, actions, counts, Q, Vt = mcts.return_results("max_visit")
            out["n_children"][b] = len(counts)
            out["actions"][b, : len(counts)] = actions
            out["counts"][b, : len(counts)] = counts
            out["Q"][b, : len(counts)] = Q
            out["V_target"][b] = Vt
        finally:
            M.MCTS.expansion = staticmethod(orig_expansion)
            import random as _r
            H.random = M.random = _r
    return out


def _alloc_discrete(B, R, A, cm):
    return dict(n_nodes=np.zeros(B, np.int32), parent=np.zeros((B, R), np.int32), paction=np.zeros((B, R), np.int32),
                node_n=np.zeros((B, R), np.int32), terminal=np.zeros((B, R), np.int32), V=np.zeros((B, R), np.float32),
                r=np.zeros((B, R), np.float64), state=np.zeros((B, R, 4), np.float64),
                prior=np.zeros((B, R, A), np.float32), eW=np.zeros((B, R, A), np.float64),
                en=np.zeros((B, R, A), np.int32), echild=np.zeros((B, R, A), np.int32),
                n_children=np.zeros(B, np.int32), actions=np.zeros((B, cm), np.float32),
                counts=np.zeros((B, cm), np.int32), Q=np.zeros((B, cm), np.float64), V_target=np.zeros(B, np.float64))


def _dump_discrete(mcts, b, out, A):
    stack = [(mcts.root_node, -1, -1, CartPoleEnv(mcts.root_node.state))]
    n = 0
    while stack:
        node, parent, pa, env = stack.pop()
        i = node._cid
        n += 1
        out["parent"][b, i], out["paction"][b, i] = parent, pa
        out["node_n"][b, i], out["terminal"][b, i] = node.n, int(node.terminal)
        out["V"][b, i], out["r"][b, i] = node.V, node.r
        out["state"][b, i] = env.state
        assert np.array_equal(np.asarray(node.state, np.float64), np.asarray(env.state)), "state replay mismatch"
        out["prior"][b, i] = node.priors
        for a, act in enumerate(node.child_actions):
            out["eW"][b, i, a], out["en"][b, i, a] = act.W, act.n
            out["echild"][b, i, a] = -1
            if hasattr(act, "child_node"):
                out["echild"][b, i, a] = act.child_node._cid
                e2 = copy.deepcopy(env)
                e2.step(a)
                stack.append((act.child_node, i, a, e2))
    out["n_nodes"][b] = n


def run_continuous(cfg: azo.Config, model, root_states: np.ndarray, tree_id0: int = 0) -> Dict[str, np.ndarray]:
    """MCTSContinuous.search per tree (root_states: [B,2] hidden th, thdot)."""
    M, _ = reference_modules()
    import alphazero.helpers as H  # type: ignore
    B, R, K = len(root_states), cfg.rows, cfg.num_components
    K3, cm = 3 * K, cfg.cmax
    out = dict(n_rows=np.zeros(B, np.int32), parent=np.zeros((B, R), np.int32), action=np.zeros((B, R), np.float32),
               eW=np.zeros((B, R), np.float64), en=np.zeros((B, R), np.int32), expanded=np.zeros((B, R), np.int32),
               node_n=np.zeros((B, R), np.int32), terminal=np.zeros((B, R), np.int32), V=np.zeros((B, R), np.float32),
               r=np.zeros((B, R), np.float64), state=np.zeros((B, R, 2), np.float64),
               head=np.zeros((B, R, K3), np.float32),
               n_children=np.zeros(B, np.int32), actions=np.zeros((B, cm), np.float32),
               counts=np.zeros((B, cm), np.int32), Q=np.zeros((B, cm), np.float64), V_target=np.zeros(B, np.float64),
               draws=np.zeros(B, np.int64), pw_inserts=np.zeros(B, np.int64),
               root_state=np.array(root_states, np.float64).reshape(B, 2))
    orig_add = M.MCTSContinuous.add_pw_action
    mt = cfg.rng_mode == azo.RNG_MT19937
    for b in range(B):
        # rng_mode MT19937: the UN-SHIMMED reference -- the stock `random` module seeded like random.seed(seed + tree) and torch's own
        # global generator seeded like torch.manual_seed(seed + tree) right before the search; nothing of torch is wrapped
        rng = StockRandom(cfg.seed + tree_id0 + b) if mt else PhiloxRandom(cfg.seed, tree_id0 + b)
        H.random = M.random = rng
        rows = [0]

        def add_pw_action(self, node, _r=rows):
            orig_add(self, node)
            _r[0] += 1
            node.child_actions[-1]._row = _r[0]

        M.MCTSContinuous.add_pw_action = add_pw_action
        try:
            env = PendulumEnv(root_states[b])
            mcts = M.MCTSContinuous(model=model, n_rollouts=cfg.n_rollouts, c_uct=cfg.c_uct, c_pw=cfg.c_pw,
                                    kappa=cfg.kappa, gamma=_gamma_arg(cfg.gamma), epsilon=cfg.epsilon,
                                    V_target_policy=cfg.V_target_policy, device="cpu", root_state=env.obs())
            if mt:
                torch.manual_seed(cfg.seed + tree_id0 + b)
                mcts.search(env)
                out["pw_inserts"][b] = rows[0]
            else:
                with NoiseInjector(cfg.seed, tree_id0 + b, K) as inj:
                    mcts.search(env)
                    out["pw_inserts"][b] = inj.j
            out["draws"][b] = rng.draws
            _dump_continuous(mcts, model, b, out, env, K)
            s, actions, counts, Q, Vt = mcts.return_results("max_visit")
            C = len(counts)
            out["n_children"][b] = C
            out["actions"][b, :C] = np.asarray(actions, np.float32).reshape(-1)
            out["counts"][b, :C] = counts
            out["Q"][b, :C] = Q
            out["V_target"][b] = Vt
        finally:
            M.MCTSContinuous.add_pw_action = orig_add
            import random as _r
            H.random = M.random = _r
    return out


@torch.no_grad()
def _torch_head(model, obs: np.ndarray, K: int) -> np.ndarray:
    """(mu, sigma, mixture probs) exactly as sample_action builds them (policies.py:656-669)."""
    x = torch.from_numpy(obs[None]).float()
    if K > 1:
        mu, sigma, log_coeff, _ = model(x)
        probs = torch.distributions.Categorical(logits=log_coeff).probs
        return torch.cat([mu, sigma, probs], -1).numpy().reshape(-1)
    mu, sigma, _ = model(x)
    return np.array([mu.item(), sigma.item(), 1.0], np.float32)


def _dump_continuous(mcts, model, b, out, root_env, K):
    root = mcts.root_node
    stack = [(root, None, -1, copy.deepcopy(root_env))]
    n = 1
    while stack:
        node, edge, parent_row, env = stack.pop()
        i = 0 if edge is None else edge._row
        out["parent"][b, i] = parent_row
        out["expanded"][b, i] = 1
        out["node_n"][b, i], out["terminal"][b, i] = node.n, int(node.terminal)
        out["V"][b, i] = float(node.V)
        out["r"][b, i] = node.r
        out["state"][b, i] = env.state
        assert np.array_equal(np.asarray(node.state, np.float64).reshape(-1), env.obs()), "obs replay mismatch"
        out["head"][b, i] = _torch_head(model, env.obs(), K)
        for act in node.child_actions:
            j = act._row
            n += 1
            out["parent"][b, j] = i
            out["action"][b, j] = np.asarray(act.action, np.float32).reshape(-1)[0]
            out["eW"][b, j], out["en"][b, j] = act.W, act.n
            if hasattr(act, "child_node"):
                e2 = copy.deepcopy(env)
                e2.step(act.action)
                stack.append((act.child_node, act, i, e2))
    out["n_rows"][b] = n


def time_reference_search(cfg: azo.Config, model, root_states: np.ndarray, seconds: float = 2.0) -> Dict[str, Any]:
    """Throughput of the unmodified reference search on one core (BASELINE.md section 2 style probe)."""
    import time
    M, _ = reference_modules()
    torch.set_num_threads(1)
    sims, t0, i = 0, time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        s = root_states[i % len(root_states)]
        i += 1
        if cfg.variant == azo.DISCRETE:
            env = CartPoleEnv(s)
            m = M.MCTSDiscrete(model=model, num_actions=cfg.num_actions, n_rollouts=cfg.n_rollouts, c_uct=cfg.c_uct,
                               gamma=_gamma_arg(cfg.gamma), epsilon=cfg.epsilon, V_target_policy=cfg.V_target_policy,
                               device="cpu", root_state=np.array(env.state))
        else:
            env = PendulumEnv(s)
            m = M.MCTSContinuous(model=model, n_rollouts=cfg.n_rollouts, c_uct=cfg.c_uct, c_pw=cfg.c_pw, kappa=cfg.kappa,
                                 gamma=_gamma_arg(cfg.gamma), epsilon=cfg.epsilon, V_target_policy=cfg.V_target_policy,
                                 device="cpu", root_state=env.obs())
        m.search(env)
        sims += cfg.n_rollouts
    dt = time.perf_counter() - t0
    return dict(sims=sims, seconds=dt, sims_per_s=sims / dt, searches=i)
