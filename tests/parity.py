"""Comparison helpers shared by the parity tests.

Tolerances (BASELINE.json north_star): integer work -- visit counts, child indices, tree topology --
bit-exact; floating point Q/V/W within 1e-5 relative.  `close()` implements the FP rule as
|a-b| <= RTOL*|ref| + ATOL with ATOL = 1e-6: network outputs are O(0.1) f32 sums with cancellation, so a
purely relative bound is meaningless for values that happen to land near zero (torch-vs-torch differs
there too); 1e-6 is 1e-5 of the typical magnitude.
"""
import numpy as np

RTOL = 1e-5
ATOL = 1e-6

DISCRETE_INT = ("n_nodes", "parent", "paction", "node_n", "terminal", "en", "echild")
DISCRETE_FP = ("V", "r", "state", "prior", "eW")
CONT_INT = ("n_rows", "parent", "en", "expanded", "node_n", "terminal")
CONT_FP = ("action", "V", "r", "state", "eW", "head")
RES_INT = ("n_children", "counts")
RES_FP = ("actions", "Q", "V_target")


def close(a, ref, rtol=RTOL, atol=ATOL):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return bool(np.all(np.abs(a - ref) <= rtol * np.abs(ref) + atol))


def max_rel(a, ref, floor=0.1):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), floor))) if a.size else 0.0


def _mask_root_r(x, key):
    # Node.r of the root is never read by the search (mcts.py:262 stops at parent_action None); after
    # tree reuse the reference's root still carries the r it was created with, ours is 0.
    if key == "r":
        x = np.array(x, copy=True)
        x[:, 0] = 0
    return x


def assert_tree_equal(got, ref, discrete, exact_fp=True, skip=(), results_only=False):
    """exact_fp=True: every array bit-identical.  False: ints bit-identical, FP within tolerance."""
    ints = (DISCRETE_INT if discrete else CONT_INT) + RES_INT
    fps = (DISCRETE_FP if discrete else CONT_FP) + RES_FP
    # a comparison that silently loses a field passes for the wrong reason: every key of the table layout must be on both sides
    # (callers that compare root results only pass dump-less dicts on BOTH sides and say so with results_only=True)
    want = (RES_INT + RES_FP) if results_only else ints + fps
    missing = [k for k in want if k not in skip and (k not in got or k not in ref)]
    assert not missing, f"arrays missing from the comparison: {missing}"
    if results_only:
        ints, fps = RES_INT, RES_FP
    for k in ints:
        if k in skip:
            continue
        assert np.array_equal(got[k], ref[k]), f"integer array {k} differs at {np.argwhere(got[k] != ref[k])[:4].tolist()}"
    for k in fps:
        if k in skip:
            continue
        a, b = _mask_root_r(got[k], k), _mask_root_r(ref[k], k)
        if exact_fp:
            same = np.array_equal(a.view(np.uint8), b.view(np.uint8)) if a.dtype == b.dtype else False
            assert same, f"fp array {k} not bit-identical at {np.argwhere(a != b)[:4].tolist()}"
        else:
            assert close(a, b), f"fp array {k} out of tolerance: max rel {max_rel(a, b)}"
