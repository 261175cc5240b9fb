"""Pins the CPU oracle (oracle/azg_oracle.c) against outputs of the reference itself.

tests/golden/*.npz were produced by oracle/gen_golden.py, which runs the UNMODIFIED
/root/reference/alphazero/search/mcts.py (MCTSDiscrete.search :418, MCTSContinuous.search :656).
The reference has no tests of its own, so these files are the golden vectors.
"""
import numpy as np
import pytest

from oracle import azo, gen_golden as G
from parity import assert_tree_equal, close, max_rel

CASES = sorted(G.CASES)


def _tapes(cfg, g):
    t = {"V": g["V"]}
    if cfg.variant == azo.DISCRETE:
        t["prior"] = g["prior"]
    else:
        t["action"] = g["action"]
    return t


@pytest.mark.parametrize("name", CASES)
def test_level_a_tape_bit_exact(name):
    """Parity level A: evaluator injected (V / priors / sampled actions from the reference run), libm math
    => topology, counts, W/Q (f64), rewards and env states all bit-identical to the reference."""
    cfg, g = G.load(name)
    cfg.use_eval_tape, cfg.math_mode = 1, azo.MATH_LIBM
    o = azo.search(cfg, None, g["root_state"], g.get("root_n_init"), tapes=_tapes(cfg, g))
    assert_tree_equal(o, g, cfg.variant == azo.DISCRETE, exact_fp=True, skip=("head",))
    assert o["counters"][5] == g["draws"].sum(), "number of tie-break / eps-greedy RNG draws differs"
    if cfg.variant == azo.CONTINUOUS:
        assert o["counters"][3] == g["pw_inserts"].sum()


@pytest.mark.parametrize("math_mode", [azo.MATH_LIBM, azo.MATH_DET])
@pytest.mark.parametrize("name", CASES)
def test_level_b_end_to_end(name, math_mode):
    """Parity level B: the oracle's own MLP (fixed fmaf order) and own noise -> action path.  Integer
    results bit-exact on these pinned seeds, floating point within 1e-5 relative (parity.close)."""
    cfg, g = G.load(name)
    cfg.use_eval_tape, cfg.math_mode = 0, math_mode
    o = azo.search(cfg, g["weights"], g["root_state"], g.get("root_n_init"))
    assert_tree_equal(o, g, cfg.variant == azo.DISCRETE, exact_fp=False)


@pytest.mark.parametrize("name", ["cartpole_n50_eps01", "pendulum_n100_k2", "pendulum_n100_k1"])
def test_mlp_vs_torch_outputs(name):
    """The oracle MLP alone vs the values torch produced inside the reference run."""
    cfg, g = G.load(name)
    cfg.math_mode = azo.MATH_DET
    if cfg.variant == azo.DISCRETE:
        n = int(g["n_nodes"][0])
        x = g["state"][0, :n].astype(np.float32)
        V, raw = azo.mlp_forward(cfg, g["weights"], x)
        pri = np.stack([azo.head_post(cfg, r) for r in raw])
        nt = g["terminal"][0, :n] == 0
        assert close(V[nt], g["V"][0, :n][nt])
        assert close(pri, g["prior"][0, :n])
    else:
        n = int(g["n_rows"][0])
        ex = g["expanded"][0, :n] == 1
        s = g["state"][0, :n][ex]
        obs = np.stack([np.cos(s[:, 0]), np.sin(s[:, 0]), s[:, 1]], 1).astype(np.float32)
        V, raw = azo.mlp_forward(cfg, g["weights"], obs)
        head = np.stack([azo.head_post(cfg, r) for r in raw])
        assert close(V, g["V"][0, :n][ex])
        assert close(head, g["head"][0, :n][ex]), max_rel(head, g["head"][0, :n][ex])


def test_thread_count_invariance():
    cfg, g = G.load("pendulum_n25_k2")
    a = azo.search(cfg, g["weights"], g["root_state"], n_threads=1)
    b = azo.search(cfg, g["weights"], g["root_state"], n_threads=4)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_tree_id_offset_is_batch_invariant():
    """Tree i's result depends on its global id only, not on batch composition."""
    cfg, g = G.load("pendulum_n25_k2")
    full = azo.search(cfg, g["weights"], g["root_state"])
    part = azo.search(cfg, g["weights"], g["root_state"][5:9], tree_id0=5)
    for k in ("counts", "Q", "eW", "parent", "action"):
        assert np.array_equal(full[k][5:9], part[k]), k


MT_CASES = sorted(G.MT_CASES)


@pytest.mark.parametrize("name", MT_CASES)
def test_mt19937_mode_reproduces_the_unshimmed_reference(name):
    """with rng_mode = MT19937 the selection draws are CPython's (init_by_array seeding, 53-bit random(),
    _randbelow_with_getrandbits for choice / randint), so the goldens of the reference run with its STOCK `random` module --
    `random.Random(seed + tree)`, no shim -- are reproduced: level A (tapes) bit for bit including the number of generator
    outputs consumed, level B (own MLP) with exact integers."""
    cfg, g = G.load(name)
    assert cfg.rng_mode == azo.RNG_MT19937
    disc = cfg.variant == azo.DISCRETE
    cfg.use_eval_tape, cfg.math_mode = 1, azo.MATH_LIBM
    o = azo.search(cfg, None, g["root_state"], g.get("root_n_init"), tapes=_tapes(cfg, g))
    assert_tree_equal(o, g, disc, exact_fp=True, skip=("head",))
    assert o["counters"][5] == g["draws"].sum(), "number of MT19937 outputs consumed differs"
    # own MLP and -- for the continuous search -- own replay of torch's generator: mt19937 seeded like torch.manual_seed(seed + tree),
    # the exponential race of torch.multinomial(probs, 1) and Box-Muller pairs with the cached sine of torch.normal
    for math_mode in (azo.MATH_LIBM, azo.MATH_DET):
        cfg.use_eval_tape, cfg.math_mode = 0, math_mode
        o = azo.search(cfg, g["weights"], g["root_state"], g.get("root_n_init"))
        assert_tree_equal(o, g, disc, exact_fp=False)
        if not disc:
            assert o["counters"][3] == g["pw_inserts"].sum()
