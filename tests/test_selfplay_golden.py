"""The self-play checker (oracle/selfplay.py) pinned against the UNMODIFIED reference agents: tests/golden/selfplay_*.npz come from
DiscreteAgent.act / ContinuousAgent.act + mcts_forward / reset_mcts of /root/reference run live (oracle/gen_selfplay_golden.py),
with temperature in {0.5, 1, 2}, max_value and deterministic selection.  Integer work (counts, chosen discrete actions, dones,
episode counters, the root-count carry-over) bit-exact; floating point within 1e-5 relative (+1e-6 absolute, tests/parity.py)."""
import numpy as np
import pytest

from oracle import azo, gen_selfplay_golden as GS, selfplay as osp
from parity import close

CASES = list(GS.CASES)
INT_KEYS = ("counts", "n_children", "done", "ep_step", "episode", "root_n")
FP_KEYS = ("obs", "actions", "Q", "V_target", "action_taken", "reward", "env_state")


def run_oracle(name, math_mode=azo.MATH_DET, eval_mode=azo.EVAL_FP32):
    cfg, meta, g = GS.load(name)
    cfg.math_mode, cfg.eval_mode, cfg.use_eval_tape = math_mode, eval_mode, 0
    ref = osp.run(cfg, g["weights"], g["states0"], meta["steps"], meta["max_episode_length"], seed=cfg.seed, tree_id0=meta["tree_id0"],
                  deterministic=meta["deterministic"], by_value=meta["final_selection"] == "max_value", temperature=meta["temperature"])
    return cfg, meta, g, ref


def compare_with_golden(steps, g, discrete):
    for s, rec in enumerate(steps):
        for k in INT_KEYS:
            assert np.array_equal(rec[k], g[k][s]), f"step {s}: {k} differs from the reference"
        if discrete:
            assert np.array_equal(rec["action_taken"], g["action_taken"][s]), f"step {s}: chosen actions differ from the reference"
        for k in FP_KEYS:
            a, b = np.asarray(rec[k]), g[k][s]
            m = min(a.shape[-1], b.shape[-1]) if a.ndim == 2 else None
            if m is not None:
                a, b = a[:, :m], b[:, :m]
            assert close(a, b), f"step {s}: {k} out of tolerance"


@pytest.mark.parametrize("math_mode", [azo.MATH_LIBM, azo.MATH_DET])
@pytest.mark.parametrize("name", CASES)
def test_selfplay_oracle_reproduces_the_reference_agents(name, math_mode):
    cfg, meta, g, ref = run_oracle(name, math_mode)
    compare_with_golden(ref, g, cfg.variant == azo.DISCRETE)


def test_temperature_cases_are_sensitive():
    """The goldens distinguish x ** temp from x ** (1 / temp): temperature 0.5 and 2 choose different actions somewhere."""
    _, _, a = GS.load("selfplay_cartpole_t05")
    _, _, b = GS.load("selfplay_cartpole_t2")
    assert not np.array_equal(a["action_taken"], b["action_taken"])
    cfg, meta, g = GS.load("selfplay_cartpole_t2")
    cfg.math_mode = azo.MATH_DET
    wrong = osp.run(cfg, g["weights"], g["states0"], meta["steps"], meta["max_episode_length"], seed=cfg.seed, tree_id0=meta["tree_id0"],
                    temperature=1.0 / meta["temperature"])
    assert any(not np.array_equal(r["action_taken"], g["action_taken"][s]) for s, r in enumerate(wrong)), \
        "an inverted temperature exponent would pass the t2 golden"
