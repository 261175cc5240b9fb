"""AZG_FLAG_EVAL_Q8 / AZO_EVAL_Q8: the hidden x hidden layers as exact int8-sliced fixed-point products (tcgen05.mma kind::i8
on the GPU, csrc/qmlp.cuh; plain integer loops in the oracle, contract in oracle/azg_oracle.h).

CPU tests pin the contract: accuracy against an f64 evaluation and the reference goldens (integer results bit-exact, FP within
1e-5).  GPU tests require the CUDA kernel to reproduce the oracle bit for bit -- possible because integer products are exact
and order-independent and everything else is a fixed sequence of correctly rounded f32 operations.
"""
import dataclasses

import numpy as np
import pytest

from oracle import azo, gen_golden as G
from parity import RES_FP, RES_INT, assert_tree_equal, close

CASES = sorted(G.CASES)
# the tensor-core kernels serve the two configured activations (relu, elu); the rest of the reference's map runs on the FP32 kernel
ENGINE_CASES = [n for n in CASES if G.CASES[n]["cfg"].activation <= azo.ACT_ELU]


def _q8(cfg):
    cfg.eval_mode = azo.EVAL_Q8
    return cfg


def _f64_value(cfg, w, x):
    H, S, L = cfg.hidden, cfg.state_dim, cfg.n_hidden
    w = w.astype(np.float64)
    a, o = x.astype(np.float64), 0
    for l in range(L):
        K = S if l == 0 else H
        W = w[o:o + H * K].reshape(H, K); o += H * K
        b = w[o:o + H]; o += H
        a = a @ W.T + b
        a = np.maximum(a, 0) if cfg.activation == azo.ACT_RELU else np.where(a > 0, a, np.expm1(np.minimum(a, 0)))
    return a @ w[o:o + H] + w[o + H]


@pytest.mark.parametrize("variant", [azo.DISCRETE, azo.CONTINUOUS])
def test_q8_accuracy_against_f64(variant):
    """V of the fixed-point evaluation vs an f64 evaluation of the same network: within 3x the FP32 path's error and
    below the 1e-6 absolute floor of the parity tolerance."""
    from alphazero_gym_b200.network import init_policy_weights
    rng = np.random.default_rng(1)
    n = 4000
    if variant == azo.CONTINUOUS:
        cfg, w = azo.continuous_config(n_rollouts=25), init_policy_weights(34, 3, 128, 3, 6)
        th = rng.uniform(-np.pi, np.pi, n)
        x = np.stack([np.cos(th), np.sin(th), rng.uniform(-8, 8, n)], 1).astype(np.float32)
    else:
        cfg, w = azo.discrete_config(n_rollouts=8), init_policy_weights(34, 4, 128, 2, 2)
        x = (rng.uniform(-1, 1, (n, 4)) * np.array([2.4, 3, 0.21, 3])).astype(np.float32)
    Vt = _f64_value(cfg, w, x)
    V32, h32 = azo.mlp_forward(cfg, w, x)
    Vq, hq = azo.mlp_forward(_q8(dataclasses.replace(cfg)), w, x)
    e32, eq = np.abs(V32 - Vt), np.abs(Vq - Vt)
    assert eq.max() < 1e-6
    assert eq.mean() < 3.0 * e32.mean() + 1e-9
    assert close(Vq, V32) and close(hq, h32)


@pytest.mark.parametrize("name", CASES)
def test_q8_oracle_vs_reference_golden(name):
    """Parity level B in Q8 mode: integer results bit-exact against the reference run, FP within 1e-5."""
    cfg, g = G.load(name)
    cfg.use_eval_tape, cfg.math_mode = 0, azo.MATH_DET
    o = azo.search(_q8(cfg), g["weights"], g["root_state"], g.get("root_n_init"))
    assert_tree_equal(o, g, cfg.variant == azo.DISCRETE, exact_fp=False)


def test_q8_scale_edge_cases():
    """Zero rows, huge / tiny activations and zero weight rows must not leave the int8 digit range (results stay finite)."""
    cfg = _q8(azo.continuous_config(n_rollouts=25))
    rng = np.random.default_rng(3)
    w = (rng.standard_normal(cfg.num_weights) * 0.08).astype(np.float32)
    H, S = cfg.hidden, cfg.state_dim
    o = S * H + H
    w[o:o + H] = 0.0                      # output 0 of hidden layer 1: all-zero weight row
    w[o + 5 * H:o + 6 * H] *= 1e-30       # tiny weights
    x = np.array([[0, 0, 0], [1e4, -1e4, 1e4], [1e-30, 0, 0], [0.3, -0.7, 7.9]], np.float32)
    V, h = azo.mlp_forward(cfg, w, x)
    V32, h32 = azo.mlp_forward(dataclasses.replace(cfg, eval_mode=azo.EVAL_FP32), w, x)
    assert np.all(np.isfinite(V)) and np.all(np.isfinite(h))
    assert np.allclose(V, V32, rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------------- GPU

def _fit(E, out, ref):
    for k in RES_INT + RES_FP:
        if k in out and k in ref and out[k].ndim == 2:
            out[k] = E.fit_columns(out[k], ref[k].shape[1])
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name,n", [("cartpole_n50_eps01", 1000), ("pendulum_n100_k2", 1000), ("pendulum_n25_k1", 37),
                                    ("pendulum_n30_k3", 128 * 148 + 5), ("pendulum_n100_k2", 65536)])
def test_q8_kernel_equals_oracle_bit_exact(name, n):
    import enginelib as E
    cfg, g = G.load(name)
    cfg.math_mode = azo.MATH_DET
    _q8(cfg)
    rng = np.random.default_rng(1)
    x = rng.uniform(-2, 2, (n, cfg.state_dim)).astype(np.float32)
    x[::7] *= 4.0
    x[3] = 0.0
    eng = E.SearchEngine(E.engine_config(cfg, 4))
    try:
        eng.set_weights(g["weights"])
        V, head = eng.mlp_forward(x)
    finally:
        eng.close()
    Vo, raw = azo.mlp_forward(cfg, g["weights"], x)
    ho = np.stack([azo.head_post(cfg, r) for r in raw[:2000]])
    assert np.array_equal(V, Vo), f"{(V != Vo).sum()} of {n} values differ, max abs {np.abs(V - Vo).max()}"
    assert np.array_equal(head[:2000], ho)


def _fused_kw(cfg, fused):
    """AZG_FLAG_FUSED: the whole search in one persistent kernel, or one launch per kernel and simulation."""
    return {"fused": fused}


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("name", ENGINE_CASES)
def test_q8_engine_equals_oracle_bit_exact(name, fused):
    import enginelib as E
    cfg, g = G.load(name)
    cfg.math_mode, cfg.use_eval_tape = azo.MATH_DET, 0
    _q8(cfg)
    ref = azo.search(cfg, g["weights"], g["root_state"], g.get("root_n_init"))
    out = _fit(E, E.run_engine(cfg, g["weights"], g["root_state"], g.get("root_n_init"), **_fused_kw(cfg, fused)), ref)
    assert_tree_equal(out, ref, cfg.variant == azo.DISCRETE, exact_fp=True)
    assert np.array_equal(out["counters"][:7], ref["counters"][:7])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ENGINE_CASES)
def test_q8_engine_vs_reference_golden(name):
    import enginelib as E
    cfg, g = G.load(name)
    _q8(cfg)
    out = _fit(E, E.run_engine(cfg, g["weights"], g["root_state"], g.get("root_n_init")), g)
    assert_tree_equal(out, g, cfg.variant == azo.DISCRETE, exact_fp=False)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("variant,B,N", [(azo.DISCRETE, 1024, 50), (azo.CONTINUOUS, 1024, 100), (azo.CONTINUOUS, 19000, 25),
                                         (azo.CONTINUOUS, 65536 + 77, 40)])
def test_q8_random_batches_equal_oracle(variant, B, N, fused):
    import enginelib as E
    if variant == azo.DISCRETE:
        cfg, roots = azo.discrete_config(n_rollouts=N, epsilon=0.1), G.cartpole_roots(B, seed=7)
    else:
        cfg, roots = azo.continuous_config(n_rollouts=N), G.pendulum_roots(B, seed=7)
    _q8(cfg)
    rng = np.random.default_rng(5)
    w = (rng.standard_normal(cfg.num_weights) * 0.08).astype(np.float32)
    kw = _fused_kw(cfg, fused)
    ref = azo.search(cfg, w, roots, tree_id0=1000, n_threads=8, dump=False)
    out = _fit(E, E.run_engine(cfg, w, roots, tree_id0=1000, dump=False, **kw), ref)
    for k in RES_INT + RES_FP:
        assert np.array_equal(out[k], ref[k]), k
    assert np.array_equal(out["counters"][:7], ref["counters"][:7])
    if fused:
        assert out["counters"][7] == 1  # one launch for the whole search


@pytest.mark.gpu
def test_fused_search_in_chunks_equals_per_simulation_launches():
    """A batch larger than one launch of the whole-search kernel covers (148 SMs x 16 tiles x 128 trees) is cut into chunks;
    trees, counters and results must not depend on it."""
    import enginelib as E
    B, N = 148 * 16 * 128 + 1000, 12
    cfg, roots = _q8(azo.continuous_config(n_rollouts=N)), G.pendulum_roots(B, seed=11)
    rng = np.random.default_rng(6)
    w = (rng.standard_normal(cfg.num_weights) * 0.08).astype(np.float32)
    a = E.run_engine(cfg, w, roots, tree_id0=5, dump=False, fused=False)
    b = E.run_engine(cfg, w, roots, tree_id0=5, dump=False, fused=True)
    for k in RES_INT + RES_FP:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["counters"][:7], b["counters"][:7])
    assert b["counters"][7] == 2
