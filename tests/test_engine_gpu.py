"""GPU parity tests: the CUDA engine, called through the C ABI, against the CPU oracle and the golden
vectors from the reference (tests/golden/, see test_oracle_vs_golden.py for how those are pinned).

Bars (BASELINE.json north_star): integer work (counts, child indices, topology) bit-exact; FP within 1e-5
relative (parity.close).  Against the oracle in deterministic-math mode the engine is required to be
bit-identical in EVERYTHING, f64 W/Q and env states included -- same fixed FMA order, same Philox streams,
same elementary functions.
"""
import numpy as np
import pytest
import torch

from oracle import azo, gen_golden as G
from parity import RES_FP, RES_INT, assert_tree_equal, close

pytestmark = pytest.mark.gpu

CASES = sorted(G.CASES)


def _eng():
    import enginelib
    return enginelib


def _fit(out, ref):
    E = _eng()
    for k in RES_INT + RES_FP:
        if k in out and k in ref and out[k].ndim == 2:
            out[k] = E.fit_columns(out[k], ref[k].shape[1])
    return out


def _tapes(cfg, g):
    t = {"V": g["V"]}
    if cfg.variant == azo.DISCRETE:
        t["prior"] = g["prior"]
    else:
        t["action"] = g["action"]
    return t


@pytest.mark.parametrize("name", CASES)
def test_engine_equals_oracle_bit_exact(name):
    cfg, g = G.load(name)
    cfg.math_mode, cfg.use_eval_tape = azo.MATH_DET, 0
    ref = azo.search(cfg, g["weights"], g["root_state"], g.get("root_n_init"))
    out = _fit(_eng().run_engine(cfg, g["weights"], g["root_state"], g.get("root_n_init")), ref)
    assert_tree_equal(out, ref, cfg.variant == azo.DISCRETE, exact_fp=True)
    assert np.array_equal(out["counters"][:7], ref["counters"][:7]), (out["counters"], ref["counters"])


@pytest.mark.parametrize("name", CASES)
def test_engine_tape_mode_vs_reference_golden(name):
    """Evaluator injected from the reference run: integer results must equal the reference's exactly;
    FP differs only through sin/cos (deterministic vs glibc, <= 1 ulp) in the env state."""
    cfg, g = G.load(name)
    out = _fit(_eng().run_engine(cfg, None, g["root_state"], g.get("root_n_init"), tapes=_tapes(cfg, g)), g)
    assert_tree_equal(out, g, cfg.variant == azo.DISCRETE, exact_fp=False, skip=("head",))
    assert out["counters"][5] == g["draws"].sum()
    if cfg.variant == azo.DISCRETE:  # rewards are exactly 1.0 and V is injected: W, Q are bit-exact too
        for k in ("eW", "Q", "V_target", "V", "prior"):
            assert np.array_equal(out[k], g[k]), k


@pytest.mark.parametrize("name", CASES)
def test_engine_end_to_end_vs_reference_golden(name):
    """Own CUDA MLP + own noise->action path vs the reference's torch-CPU run on the same weights/seeds."""
    cfg, g = G.load(name)
    out = _fit(_eng().run_engine(cfg, g["weights"], g["root_state"], g.get("root_n_init")), g)
    assert_tree_equal(out, g, cfg.variant == azo.DISCRETE, exact_fp=False)


@pytest.mark.parametrize("name", ["cartpole_n50_eps01", "pendulum_n100_k2", "pendulum_n25_k1", "pendulum_n30_k3"])
def test_mlp_kernel(name):
    """Evaluation kernel alone: bit-identical to the oracle MLP, within tolerance of torch's outputs."""
    cfg, g = G.load(name)
    cfg.math_mode = azo.MATH_DET
    rng = np.random.default_rng(1)
    n = 1000  # not a multiple of the 128-row tile
    x = rng.uniform(-2, 2, (n, cfg.state_dim)).astype(np.float32)
    E = _eng()
    eng = E.SearchEngine(E.engine_config(cfg, 4))
    try:
        eng.set_weights(g["weights"])
        V, head = eng.mlp_forward(x)
    finally:
        eng.close()
    Vo, raw = azo.mlp_forward(cfg, g["weights"], x)
    ho = np.stack([azo.head_post(cfg, r) for r in raw])
    assert np.array_equal(V, Vo)
    assert np.array_equal(head, ho)
    # torch fp32 reference of the same network (plain nn.Linear stack built from the flat weights)
    H, S, L = cfg.hidden, cfg.state_dim, cfg.n_hidden
    w = torch.from_numpy(g["weights"])
    off, h = 0, torch.from_numpy(x)
    for l in range(L):
        K = S if l == 0 else H
        W = w[off:off + H * K].reshape(H, K); off += H * K
        b = w[off:off + H]; off += H
        h = torch.nn.functional.linear(h, W, b)
        h = torch.relu(h) if cfg.activation == azo.ACT_RELU else torch.nn.functional.elu(h)
    Wv = w[off:off + H].reshape(1, H); off += H
    bv = w[off:off + 1]; off += 1
    Vt = torch.nn.functional.linear(h, Wv, bv).reshape(-1).numpy()
    assert close(V, Vt)


@pytest.mark.parametrize("variant", [azo.DISCRETE, azo.CONTINUOUS])
def test_env_step_kernel(variant):
    cfg = azo.discrete_config() if variant == azo.DISCRETE else azo.continuous_config()
    rng = np.random.default_rng(2)
    n = 4096
    if variant == azo.DISCRETE:
        s = rng.uniform(-1, 1, (n, 4)) * np.array([2.5, 3, 0.25, 3])
        a = rng.integers(0, 2, n).astype(np.float32)
    else:
        s = np.stack([rng.uniform(-30, 30, n), rng.uniform(-8, 8, n)], 1)
        a = rng.uniform(-2.5, 2.5, n).astype(np.float32)
    E = _eng()
    eng = E.SearchEngine(E.engine_config(cfg, 4))
    try:
        nxt, rew, term, obs = eng.env_step(s, a)
    finally:
        eng.close()
    for mode, exact in ((azo.MATH_DET, True), (azo.MATH_LIBM, False)):
        cfg.math_mode = mode
        ref = [azo.env_step(cfg, s[i], float(a[i])) for i in range(n)]
        rn = np.stack([r[0] for r in ref]); rr = np.array([r[1] for r in ref])
        rt = np.array([r[2] for r in ref]); ro = np.stack([r[3] for r in ref])
        assert np.array_equal(term.astype(bool), rt)
        if exact:
            assert np.array_equal(nxt, rn) and np.array_equal(rew, rr) and np.array_equal(obs, ro)
        else:
            assert np.allclose(nxt, rn, rtol=1e-13, atol=1e-13) and np.allclose(rew, rr, rtol=1e-13, atol=1e-13)
            assert np.allclose(obs, ro, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", ["cartpole_n50_eps01", "pendulum_n100_k2"])
def test_host_path_graph_and_batch_invariance(name):
    cfg, g = G.load(name)
    E = _eng()
    a = E.run_engine(cfg, g["weights"], g["root_state"])
    b = E.run_engine(cfg, g["weights"], g["root_state"], host=True)              # pinned-host C-ABI entry point
    c = E.run_engine(cfg, g["weights"], g["root_state"], use_graph=False)        # direct launches
    d = E.run_engine(cfg, g["weights"], g["root_state"][3:6], tree_id0=3)         # same trees in another batch
    for k in a:
        if k == "counters":
            continue
        assert np.array_equal(a[k], b[k]), ("host", k)
        assert np.array_equal(a[k], c[k]), ("nograph", k)
        assert np.array_equal(a[k][3:6], d[k]), ("batch", k)


@pytest.mark.parametrize("variant,B,N", [(azo.DISCRETE, 1024, 50), (azo.CONTINUOUS, 1024, 100), (azo.CONTINUOUS, 300, 25),
                                         (azo.CONTINUOUS, 257, 200)])
def test_random_batches_equal_oracle(variant, B, N):
    """Bigger seeded batches (ragged vs the CTA/tile sizes) at the BASELINE rollouts: everything bit-exact."""
    if variant == azo.DISCRETE:
        cfg = azo.discrete_config(n_rollouts=N, epsilon=0.1)
        roots = G.cartpole_roots(B, seed=7)
    else:
        cfg = azo.continuous_config(n_rollouts=N)
        roots = G.pendulum_roots(B, seed=7)
    rng = np.random.default_rng(5)
    w = (rng.standard_normal(cfg.num_weights) * 0.08).astype(np.float32)
    ref = azo.search(cfg, w, roots, tree_id0=1000, n_threads=8)
    out = _fit(_eng().run_engine(cfg, w, roots, tree_id0=1000), ref)
    assert_tree_equal(out, ref, variant == azo.DISCRETE, exact_fp=True)
    assert np.array_equal(out["counters"][:7], ref["counters"][:7])


def test_errors():
    E = _eng()
    from alphazero_gym_b200 import _cabi
    cfg = azo.continuous_config(n_rollouts=25)
    with pytest.raises(_cabi.AzgError):
        E.SearchEngine(E.engine_config(cfg, 4, hidden=100))
    eng = E.SearchEngine(E.engine_config(cfg, 4))
    try:
        with pytest.raises(_cabi.AzgError):  # weights not set
            eng.search(torch.zeros((2, 2), dtype=torch.float64, device="cuda"), 25)
        with pytest.raises(_cabi.AzgError):
            eng.set_weights(np.zeros(10, np.float32))
        eng.set_weights(np.zeros(eng.num_weights, np.float32))
        with pytest.raises(_cabi.AzgError):  # more rollouts than capacity
            eng.search(torch.zeros((2, 2), dtype=torch.float64, device="cuda"), 26)
        with pytest.raises(_cabi.AzgError):  # more trees than capacity
            eng.search(torch.zeros((5, 2), dtype=torch.float64, device="cuda"), 25)
    finally:
        eng.close()


@pytest.mark.gpu
def test_host_entry_point_with_page_locked_buffers():
    """azg_search_host DMAs straight into page-locked caller buffers; same bits as the staged path for pageable ones."""
    import enginelib as E
    import torch
    cfg, g = G.load("pendulum_n25_k2")
    cfg.math_mode, cfg.use_eval_tape = azo.MATH_DET, 0
    roots = G.pendulum_roots(300, seed=3)
    eng = E.SearchEngine(E.engine_config(cfg, 300))
    try:
        eng.set_weights(g["weights"])
        a = eng.search_host(roots, cfg.n_rollouts, tree_id0=5)
        out = eng.host_buffers(300)
        b = eng.search_host(torch.from_numpy(roots).pin_memory().numpy(), cfg.n_rollouts, tree_id0=5, out=out)
        assert b is out
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    finally:
        eng.close()


@pytest.mark.gpu
def test_pipelined_host_entry_points_equal_the_blocking_call():
    """azg_search_host_begin / _end with two searches in flight return what azg_search_host returns, slot by slot, and refuse
    pageable buffers and a slot that is still in flight."""
    import torch
    import enginelib as E
    from alphazero_gym_b200._cabi import AzgError
    cfg = azo.continuous_config(n_rollouts=25)
    cfg.eval_mode = azo.EVAL_Q8
    B = 700
    w = (np.random.default_rng(4).standard_normal(cfg.num_weights) * 0.08).astype(np.float32)
    eng = E.SearchEngine(E.engine_config(cfg, B))
    try:
        eng.set_weights(w)
        roots = [torch.from_numpy(G.pendulum_roots(B, seed=s)).pin_memory().numpy() for s in (1, 2, 3)]
        want = [{k: v.copy() for k, v in eng.search_host(r, cfg.n_rollouts, tree_id0=9).items()} for r in roots]
        outs = [eng.host_buffers(B), eng.host_buffers(B)]
        eng.search_host_begin(0, roots[0], cfg.n_rollouts, outs[0], tree_id0=9)
        eng.search_host_begin(1, roots[1], cfg.n_rollouts, outs[1], tree_id0=9)
        with pytest.raises(AzgError):
            eng.search_host_begin(1, roots[2], cfg.n_rollouts, outs[1], tree_id0=9)  # slot 1 is in flight
        eng.search_host_end(0)
        for k in want[0]:
            assert np.array_equal(outs[0][k], want[0][k]), k
        eng.search_host_begin(0, roots[2], cfg.n_rollouts, outs[0], tree_id0=9)
        eng.search_host_end(1)
        for k in want[1]:
            assert np.array_equal(outs[1][k], want[1][k]), k
        eng.search_host_end(0)
        for k in want[2]:
            assert np.array_equal(outs[0][k], want[2][k]), k
        with pytest.raises(AzgError):
            eng.search_host_end(0)  # nothing in flight
        pageable = dict(actions=np.empty((B, eng.cmax), np.float32), counts=np.empty((B, eng.cmax), np.int32),
                        Q=np.empty((B, eng.cmax), np.float64), V_target=np.empty(B, np.float64), n_children=np.empty(B, np.int32))
        with pytest.raises(AzgError):
            eng.search_host_begin(0, roots[0], cfg.n_rollouts, pageable, tree_id0=9)
    finally:
        eng.close()
