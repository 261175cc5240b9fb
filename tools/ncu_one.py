#!/usr/bin/env python
"""One warmed-up whole-search launch for ncu: python tools/ncu_one.py [workload] (ncu -k regex:k_search_wg -s 1 -c 1 ...)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from alphazero_gym_b200.engine import SearchEngine  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "pendulum_65536x100"
variant, B, N = bench.WORKLOADS[wl]
eng = SearchEngine(bench.engine_config(variant, B, N, 0, q8=True, fused=True))
eng.set_weights(bench.make_weights(variant))
roots = torch.from_numpy(bench.make_roots(variant, B)).cuda()
for _ in range(3):
    eng.search(roots, N)
torch.cuda.synchronize()
eng.close()
