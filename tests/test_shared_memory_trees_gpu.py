"""The discrete whole-search kernel with the trees resident in shared memory (qmlp2.cuh TSM, tree_discrete.cuh ds_step)
against the CPU oracle -- full tree dump, root results and counters, bit for bit -- and against the same search on rows in HBM
(AZG_NO_TSM=1).  Sizes chosen for the kernel's own edge cases: one tree, a ragged last CTA, a lane group without a tree next to one
with, one full wave (BASELINE config 3 + 5 trees), two waves as two launches, paths longer than a lane group (chunked backup) and more
levels than one batch of draws, tree reuse (root visit counts carried in), the on-policy value target, epsilon = 0 (one draw per level)."""
import os

import numpy as np
import pytest

from oracle import azo
from parity import assert_tree_equal

pytestmark = pytest.mark.gpu

TSM = "two_phase_trees_in_shared_memory"


def _weights(cfg=None):
    from alphazero_gym_b200.network import init_policy_weights
    return init_policy_weights(34, 4, 128, cfg.n_hidden if cfg is not None else 2, 2)


def _roots(B, seed=34):
    return np.random.default_rng(seed).uniform(-0.05, 0.05, size=(B, 4))


def _run(cfg, roots, root_n_init=None, tree_id0=0, no_tsm=False):
    import torch
    import enginelib as E
    from alphazero_gym_b200.engine import SearchEngine
    B = roots.shape[0]
    old = os.environ.pop("AZG_NO_TSM", None)
    if no_tsm:
        os.environ["AZG_NO_TSM"] = "1"
    try:
        eng = SearchEngine(E.engine_config(cfg, B))
        try:
            eng.set_weights(_weights(cfg))
            eng.set_reward_model(cfg.reward_step, cfg.reward_terminal)
            rn = None if root_n_init is None else torch.from_numpy(np.ascontiguousarray(root_n_init, np.int32)).cuda()
            eng.search(torch.from_numpy(np.ascontiguousarray(roots, np.float64)).cuda(), cfg.n_rollouts, rn, tree_id0)
            eng.status()
            out = {k: v.cpu().numpy() for k, v in eng.root_results().items()}
            out.update(eng.dump_tree(B))
            c = eng.counters(B)
            out["counters"] = np.array([c["sims"], c["levels"], c["children_scanned"], c["pw_inserts"], c["evals"], c["rng_draws"],
                                        c["terminal_leaf_sims"]], np.int64)
            out["kernel"] = eng.fused_stats()["kernel"]
            out["launches"] = c["launches"]
            return out
        finally:
            eng.close()
    finally:
        os.environ.pop("AZG_NO_TSM", None)
        if old is not None:
            os.environ["AZG_NO_TSM"] = old


def _fit(out, ref):
    import enginelib as E
    for k in ("counts", "actions", "Q"):
        out[k] = E.fit_columns(out[k], ref[k].shape[1])
    return out


CASES = {
    # name: (trees, oracle config kwargs, tree reuse)
    "one_tree": (1, dict(n_rollouts=50, epsilon=0.1), False),
    "ragged_37": (37, dict(n_rollouts=50, epsilon=0.1), False),
    "partial_warp_600": (600, dict(n_rollouts=20, epsilon=0.1), False),        # 5 trees per CTA: a lane group without a tree
    "config3_plus_5": (4101, dict(n_rollouts=50, epsilon=0.1), False),          # 28 trees per CTA, ragged last CTA
    "two_waves_6000": (6000, dict(n_rollouts=50, epsilon=0.1), False),          # two launches of 3000 trees
    "deep_300x200": (300, dict(n_rollouts=200, epsilon=0.05), False),           # paths > 8 levels, > 16 levels per search step
    "eps0_on_policy_reuse": (257, dict(n_rollouts=64, epsilon=0.0, gamma=0.99, V_target_policy="on_policy"), True),
    "three_layers_elu": (300, dict(n_rollouts=30, epsilon=0.1, n_hidden=3, activation=azo.ACT_ELU), False),  # two tensor-core layers per evaluation
    "reward_wrappers_rs": (300, dict(n_rollouts=60, epsilon=0.1, reward_step=0.005 / 250.0, reward_terminal=-1 / 250.0), False),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_shared_memory_trees_equal_oracle_and_hbm_rows(name):
    B, kw, reuse = CASES[name]
    cfg = azo.discrete_config(**kw)
    cfg.eval_mode, cfg.math_mode = azo.EVAL_Q8, azo.MATH_DET
    roots = _roots(B)
    rn = np.random.default_rng(5).integers(0, 30, B).astype(np.int32) if reuse else None
    tree_id0 = 1000 if reuse else 0
    ref = azo.search(cfg, _weights(cfg), roots, rn, tree_id0=tree_id0, n_threads=os.cpu_count() or 8)
    out = _fit(_run(cfg, roots, rn, tree_id0), ref)
    assert out["kernel"] == TSM, out["kernel"]
    assert out["launches"] == (2 if name.startswith("two_waves") else 1)
    assert_tree_equal(out, ref, True, exact_fp=True)
    assert np.array_equal(out["counters"], ref["counters"][:7]), (out["counters"], ref["counters"])
    if B <= 1000:  # the same search with the rows in HBM: every array of the dump identical
        hbm = _fit(_run(cfg, roots, rn, tree_id0, no_tsm=True), ref)
        assert hbm["kernel"] == "two_phase"
        assert_tree_equal(hbm, out, True, exact_fp=True)
        assert np.array_equal(hbm["counters"], out["counters"])
    if name.startswith("deep"):
        assert int(out["counters"][1]) > 8 * B * cfg.n_rollouts  # mean depth above one lane group
