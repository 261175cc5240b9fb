"""AZG_FLAG_RNG_MT19937: the discrete search drawing from CPython's generator.

The goldens cartpole_mt_* come from the reference run with its STOCK `random` module (oracle/ref_harness.py StockRandom =
random.Random(seed + tree), no shim).  CPU side: tests/test_oracle_vs_golden.py pins the oracle's MT19937 mode against them bit for
bit.  Here the CUDA engine must equal the oracle bit for bit (trees, results, number of generator outputs consumed) in both launch
schedules and both evaluation modes, and reproduce the reference's integers exactly (floating point within 1e-5)."""
import dataclasses

import numpy as np
import pytest

from oracle import azo, gen_golden as G
from parity import RES_FP, RES_INT, assert_tree_equal

MT_CASES = sorted(G.MT_CASES)


def _fit(E, out, ref):
    for k in RES_INT + RES_FP:
        if k in out and k in ref and out[k].ndim == 2:
            out[k] = E.fit_columns(out[k], ref[k].shape[1])
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("q8", [False, True])
@pytest.mark.parametrize("name", MT_CASES)
def test_mt19937_engine_equals_oracle_and_reference(name, q8):
    import enginelib as E
    cfg, g = G.load(name)
    cfg.math_mode, cfg.use_eval_tape = azo.MATH_DET, 0
    if q8:
        cfg.eval_mode = azo.EVAL_Q8  # tensor-core evaluation -> the whole-search kernel
    ref = azo.search(cfg, g["weights"], g["root_state"], g.get("root_n_init"))
    out = _fit(E, E.run_engine(cfg, g["weights"], g["root_state"], g.get("root_n_init")), ref)
    disc = cfg.variant == azo.DISCRETE
    assert_tree_equal(out, ref, disc, exact_fp=True)
    assert np.array_equal(out["counters"][:7], ref["counters"][:7])  # [5] = MT19937 outputs consumed
    assert_tree_equal(out, g, disc, exact_fp=False)  # = the UN-SHIMMED reference run (stock `random`, un-wrapped torch generator)


@pytest.mark.gpu
def test_mt19937_random_batch_and_sharding():
    """1000 trees, tree ids offset: the generator of tree i is seeded with seed + global id, so a shard equals the slice."""
    import enginelib as E
    cfg = azo.discrete_config(n_rollouts=50, epsilon=0.1)
    cfg.rng_mode = azo.RNG_MT19937
    roots = G.cartpole_roots(1000, seed=3)
    w = (np.random.default_rng(2).standard_normal(cfg.num_weights) * 0.08).astype(np.float32)
    ref = azo.search(cfg, w, roots, tree_id0=77, n_threads=8, dump=False)
    out = _fit(E, E.run_engine(cfg, w, roots, tree_id0=77, dump=False), ref)
    part = _fit(E, E.run_engine(cfg, w, roots[300:400], tree_id0=377, dump=False), ref)
    for k in RES_INT + RES_FP:
        assert np.array_equal(out[k], ref[k]), k
        assert np.array_equal(part[k], ref[k][300:400]), k
    assert np.array_equal(out["counters"][:7], ref["counters"][:7])


@pytest.mark.gpu
def test_mt19937_continuous_batch_and_sharding():
    """The continuous search in un-shimmed mode on 600 trees with offset ids (both generators of tree i are seeded with
    seed + global id): equal to the oracle bit for bit, a shard equals the slice; the whole-search kernel refuses the mode."""
    import enginelib as E
    from alphazero_gym_b200._cabi import AzgError
    cfg = azo.continuous_config(n_rollouts=40, epsilon=0.1)
    cfg.rng_mode = azo.RNG_MT19937
    roots = G.pendulum_roots(600, seed=3)
    w = (np.random.default_rng(2).standard_normal(cfg.num_weights) * 0.08).astype(np.float32)
    ref = azo.search(cfg, w, roots, tree_id0=77, n_threads=8, dump=False)
    out = _fit(E, E.run_engine(cfg, w, roots, tree_id0=77, dump=False), ref)
    part = _fit(E, E.run_engine(cfg, w, roots[300:400], tree_id0=377, dump=False), ref)
    for k in RES_INT + RES_FP:
        assert np.array_equal(out[k], ref[k]), k
        assert np.array_equal(part[k], ref[k][300:400]), k
    assert np.array_equal(out["counters"][:7], ref["counters"][:7])
    with pytest.raises(AzgError):
        E.SearchEngine(dataclasses.replace(E.engine_config(cfg, 4), eval_q8=True, fused=True))
