"""Parity at the EXACT sizes BASELINE.json names, on the path bench.py times (tensor-core evaluation, whole-search kernel):
  config 3   4096 CartPole trees x 50 simulations
  config 4   65536 Pendulum trees x 100 simulations (progressive widening, GMM K = 2)
  config 5   32768 Pendulum trees x 200 simulations per GPU (root fan-out 15 of the 16 slots, 202 rows per tree)
plus a search at the 8-bit row-link limit (253 rollouts -> 255 rows per tree).  Weights and roots are bench.py's
(default initialisation torch.manual_seed(34); numpy default_rng(34)).  Every root result -- visit counts, actions, Q, V_target,
children counts -- and the per-search counters must equal the CPU oracle's bit for bit; the oracle itself is pinned to the reference
by tests/test_oracle_vs_golden.py and tests/test_q8_eval.py (same network family, 16-tree goldens)."""
import os

import numpy as np
import pytest

from oracle import azo
from parity import RES_FP, RES_INT

pytestmark = pytest.mark.gpu


def _weights(variant):
    from alphazero_gym_b200.network import init_policy_weights
    return init_policy_weights(34, 4, 128, 2, 2) if variant == azo.DISCRETE else init_policy_weights(34, 3, 128, 3, 6)


def _roots(variant, B):
    rng = np.random.default_rng(34)
    if variant == azo.DISCRETE:
        return rng.uniform(-0.05, 0.05, size=(B, 4))
    return np.stack([rng.uniform(-np.pi, np.pi, B), rng.uniform(-1.0, 1.0, B)], 1)


CASES = {
    "config3_cartpole_4096x50": (azo.DISCRETE, 4096, dict(n_rollouts=50, epsilon=0.1)),
    "config4_pendulum_65536x100": (azo.CONTINUOUS, 65536, dict(n_rollouts=100)),
    "config5_pendulum_32768x200": (azo.CONTINUOUS, 32768, dict(n_rollouts=200)),
    "row_link_limit_pendulum_4096x253": (azo.CONTINUOUS, 4096, dict(n_rollouts=253, kappa=0.45)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_baseline_size_equals_oracle_bit_exact(name):
    import enginelib as E
    variant, B, kw = CASES[name]
    cfg = azo.discrete_config(**kw) if variant == azo.DISCRETE else azo.continuous_config(**kw)
    cfg.eval_mode, cfg.math_mode = azo.EVAL_Q8, azo.MATH_DET
    w, roots = _weights(variant), _roots(variant, B)
    ref = azo.search(cfg, w, roots, dump=False, n_threads=os.cpu_count() or 8)
    out = E.run_engine(cfg, w, roots, dump=False)
    assert out["counters"][7] == 1, "the whole search is expected to be one launch"
    for k in RES_INT + RES_FP:
        got = E.fit_columns(out[k], ref[k].shape[1]) if out[k].ndim == 2 else out[k]
        assert np.array_equal(got, ref[k]), f"{name}: {k} differs from the oracle"
    assert np.array_equal(out["counters"][:7], ref["counters"][:7]), (out["counters"], ref["counters"])
    assert int(out["counts"].sum()) == B * cfg.n_rollouts
    if name.startswith("config5"):
        assert int(out["n_children"].max()) == 15  # ceil(sqrt(201)): 15 of the 16 root-edge slots
    if name.startswith("row_link"):
        assert cfg.rows == 255
