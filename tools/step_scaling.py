#!/usr/bin/env python
"""Average launch time of the tree-step and evaluation kernels as a function of the number of trees (latency- or
throughput-bound?):  python tools/step_scaling.py [N]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
from alphazero_gym_b200.engine import SearchEngine
w = bench.make_weights("continuous")
for B in (8192, 16384, 32768, 65536, 131072, 262144):
    eng = SearchEngine(bench.engine_config("continuous", B, N, 0, q8=True))
    eng.set_weights(w)
    roots = torch.from_numpy(bench.make_roots("continuous", B)).cuda()
    eng.profile_search(roots, N)
    pr = eng.profile_search(roots, N)
    c = eng.counters()
    print(json.dumps({"B": B, "N": N, "step_us": 1e3 * pr["tree_step"]["ms"] / pr["tree_step"]["launches"],
                      "eval_us": 1e3 * pr["evaluation"]["ms"] / pr["evaluation"]["launches"],
                      "step_ns_per_tree": 1e6 * pr["tree_step"]["ms"] / pr["tree_step"]["launches"] / B,
                      "eval_ns_per_tree": 1e6 * pr["evaluation"]["ms"] / pr["evaluation"]["launches"] / B}), flush=True)
    eng.close()
