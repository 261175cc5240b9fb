#!/usr/bin/env python
"""Cycle accounting of the whole-search kernel on one workload: python tools/fused_stats.py [workload] [reps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from alphazero_gym_b200.engine import SearchEngine  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "pendulum_65536x100"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    variant, B, N = bench.WORKLOADS[wl]
    eng = SearchEngine(bench.engine_config(variant, B, N, 0, q8=True, fused=True))
    eng.set_weights(bench.make_weights(variant))
    roots = torch.from_numpy(bench.make_roots(variant, B)).cuda()
    eng.search(roots, N)
    torch.cuda.synchronize()
    eng.fused_stats()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        eng.search(roots, N)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    st = eng.fused_stats()
    st.update(workload=wl, ms_per_search=ms, sims_per_s=B * N / ms * 1e3, us_per_sim=ms * 1e3 / (N + 1))
    print(json.dumps(st))
    eng.close()
    # the same search as one launch per simulation step
    eng = SearchEngine(bench.engine_config(variant, B, N, 0, q8=True, fused=False))
    eng.set_weights(bench.make_weights(variant))
    eng.search(roots, N)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        eng.search(roots, N)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    prof = eng.profile_search(roots, N)
    print(json.dumps(dict(per_simulation_launches=True, ms_per_search=ms, sims_per_s=B * N / ms * 1e3, us_per_sim=ms * 1e3 / (N + 1),
                          us_per_launch={k: v["ms"] * 1e3 / max(1, v["launches"]) for k, v in prof.items()})))


if __name__ == "__main__":
    main()
