#!/usr/bin/env python
"""Training-step throughput: Trainer.update on device-resident replay batches vs the same step on the host cores.

    python tools/train_bench.py [steps]

Prints one JSON line per batch size: samples/s on cuda:0 (CUDA events, batches already in HBM) and on the CPU (the same Trainer on
the host, torch CPU kernels, all cores) -- a baseline, not a target.  Network / loss / optimizer = config/run_continuous.yaml
(GMM K = 2, 3-128-128-128 ELU, A0CLossTuned, RMSprop)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_gym_b200.network import PolicyNet, init_policy_weights  # noqa: E402
from alphazero_gym_b200.train import LossConfig, Trainer  # noqa: E402


def batch(n, device, cmax=10, seed=0):
    rng = np.random.default_rng(seed)
    th = rng.uniform(-np.pi, np.pi, n)
    obs = np.stack([np.cos(th), np.sin(th), rng.uniform(-8, 8, n)], 1).astype(np.float32)
    counts = rng.multinomial(100, np.full(cmax, 1.0 / cmax), n).astype(np.int32)
    counts[counts == 0] = 1
    return dict(obs=torch.from_numpy(obs).to(device), actions=torch.from_numpy(rng.uniform(-1.99, 1.99, (n, cmax)).astype(np.float32)).to(device),
                counts=torch.from_numpy(counts).to(device), V_target=torch.from_numpy(rng.uniform(-3, 0, n)).to(device))


def run(device, n, steps, cuda_graph=False):
    net = PolicyNet(3, 128, 3, 6, "elu", num_components=2).load_flat(init_policy_weights(34, 3, 128, 3, 6)).to(device)
    tr = Trainer(net, LossConfig(tuned=True), optimizer="rmsprop", cuda_graph=cuda_graph)
    bs = [batch(n, device, seed=s) for s in range(4)]
    for i in range(5):
        tr.update(bs[i % 4])
    if device != "cpu":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        tr.update(bs[i % 4])
    if device != "cpu":
        torch.cuda.synchronize()
    return n * steps / (time.perf_counter() - t0)


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    for n in (32, 4096, 65536):
        gpu = run("cuda:0", n, steps)
        gpu_graph = run("cuda:0", n, steps, cuda_graph=True)
        cpu = run("cpu", n, max(5, steps // (1 if n <= 4096 else 20)))
        print(json.dumps({"metric": "training samples/s (Agent.update, A0CLossTuned + RMSprop)", "batch": n, "gpu_cuda_graph": gpu_graph,
                          "gpu_eager": gpu, "cpu": cpu, "cpu_threads": torch.get_num_threads(),
                          "gpu_graph_ms_per_step": 1e3 * n / gpu_graph, "gpu_eager_ms_per_step": 1e3 * n / gpu}), flush=True)


if __name__ == "__main__":
    main()
