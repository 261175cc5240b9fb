#!/usr/bin/env python
"""Tiny whole-search runs against the oracle (kernel bring-up): python tools/wg_smoke.py [variant B N]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench  # noqa: E402
from oracle import azo  # noqa: E402
from alphazero_gym_b200.engine import SearchEngine  # noqa: E402


def one(variant, B, N):
    eng = SearchEngine(bench.engine_config(variant, B, N, 0, q8=True, fused=True))
    w = bench.make_weights(variant)
    eng.set_weights(w)
    roots = bench.make_roots(variant, B)
    eng.search(torch.from_numpy(roots).cuda(), N)
    torch.cuda.synchronize()
    eng.status()
    got = {k: v.cpu().numpy() for k, v in eng.root_results().items()}
    cfg = bench.oracle_config(variant, N)
    cfg.eval_mode = azo.EVAL_Q8
    ref = azo.search(cfg, w, roots, dump=False, n_threads=8)
    c = ref["counts"].shape[1]
    ok = all(np.array_equal(got[k][:, :c] if got[k].ndim == 2 else got[k], ref[k]) for k in ("counts", "Q", "V_target", "actions", "n_children"))
    print(variant, B, N, "OK" if ok else "MISMATCH", "counts sum", int(got["counts"].sum()), "expected", B * N, flush=True)
    if not ok:
        bad = np.argwhere(got["counts"][:, :c] != ref["counts"])
        print(" first bad rows", bad[:5].tolist(), "of", len(bad))
    eng.close()


if __name__ == "__main__":
    if len(sys.argv) > 3:
        one(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
    else:
        for args in (("continuous", 64, 2), ("continuous", 512, 5), ("continuous", 443 * 3 + 7, 25), ("discrete", 100, 8), ("discrete", 4096, 20),
                     ("continuous", 148 * 16 * 128 + 1000, 6)):
            one(*args)
