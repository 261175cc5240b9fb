#!/usr/bin/env python
"""bench.py -- MCTS simulations/sec of the batched search hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

A "step" is one full batched search over one batch of synthetic roots: root evaluation, n_rollouts
simulations (backup + select + env step + leaf evaluation) and root-result extraction.
Workload (default, BASELINE.json configs[3]): 65536 Pendulum-v0 trees per GPU x 100
simulations, progressive widening, GMM K=2 policy, MLP 3-128-128-128 ELU, default-initialised weights
(torch.manual_seed(34)), roots from numpy default_rng(34).  Trees are independent, so N GPUs run N shards
with no data-path collective (weak scaling: 65536 trees per GPU).

One JSON line on stdout:
  value      whole-job sims/s, roots already resident in HBM, results left on the device (CUDA events)
  e2e        the same through the host-buffer C-ABI call azg_search_host: pinned H2D of the roots and D2H of
             (actions, counts, Q, V_target, n_children) inside the timed region
  roofline   for the dominant kernel, from per-launch CUDA-event times taken live (azg_profile_search)
  cpu_baseline  the CPU oracle (a C port of the reference search) on a bounded sample of the same workload
`--impl reference` times that CPU port alone (the reference itself is Python + gym and cannot travel to the
GPU box; its single-core numbers measured in the build container are in BASELINE.md / DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (variant, trees per GPU, n_rollouts)
    "pendulum_65536x100": ("continuous", 65536, 100),
    "cartpole_4096x50": ("discrete", 4096, 50),
    "pendulum_32768x200": ("continuous", 32768, 200),
    "pendulum_1024x25": ("continuous", 1024, 25),  # quick sanity size
    # BASELINE.json configs[4]: self-play loop, 32768 Pendulum environments per GPU (262144 on 8 GPUs) x 200 sims per env step,
    # weight broadcast every SELFPLAY_BROADCAST_EVERY steps, replay rows all-gathered every step
    "selfplay_pendulum_32768x200": ("continuous", 32768, 200),
    "selfplay_pendulum_2048x25": ("continuous", 2048, 25),  # quick sanity size
}
SELFPLAY_BROADCAST_EVERY = 4
FLOP_PER_EVAL = {"discrete": 2 * 17280, "continuous": 2 * 34048}


def make_roots(variant: str, B: int, seed: int = 34) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if variant == "discrete":
        return rng.uniform(-0.05, 0.05, size=(B, 4))
    return np.stack([rng.uniform(-np.pi, np.pi, B), rng.uniform(-1.0, 1.0, B)], 1)


def make_weights(variant: str) -> np.ndarray:
    from alphazero_gym_b200.network import init_policy_weights
    if variant == "discrete":
        return init_policy_weights(34, 4, 128, 2, 2)
    return init_policy_weights(34, 3, 128, 3, 6)


def engine_config(variant: str, B: int, N: int, device: int, q8: bool = False, fused=None):
    from alphazero_gym_b200.engine import EngineConfig
    from alphazero_gym_b200._cabi import ACT_ELU, ACT_RELU, CONTINUOUS, DISCRETE
    if variant == "discrete":  # run_discrete.yaml / MCTSDiscrete.yaml / DiscretePolicy.yaml
        return EngineConfig(variant=DISCRETE, max_rollouts=N, max_trees=B, num_actions=2, state_dim=4, hidden=128, n_hidden=2,
                            activation=ACT_RELU, c_uct=1.5, gamma=1.0, epsilon=0.1, device=device, seed=34, eval_q8=q8, fused=fused)
    return EngineConfig(variant=CONTINUOUS, max_rollouts=N, max_trees=B, num_components=2, state_dim=3, hidden=128, n_hidden=3,
                        activation=ACT_ELU, c_uct=0.05, c_pw=1.0, kappa=0.5, gamma=1.0, epsilon=0.0, action_bound=2.0,
                        device=device, seed=34, eval_q8=q8, fused=fused)


def oracle_config(variant: str, N: int):
    from oracle import azo
    if variant == "discrete":
        return azo.discrete_config(n_rollouts=N, epsilon=0.1, math_mode=azo.MATH_DET)
    return azo.continuous_config(n_rollouts=N, math_mode=azo.MATH_DET)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc, self.thread = device, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    """(HBM GB/s, SM max MHz, dense bf16 TFLOP/s burst, source)"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), d.get("bf16_tflops", 1650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, 1650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(variant: str, c: dict) -> float:
    """Algorithmic node-table bytes of the tree-step kernels for one search:
    select  sum_levels [4 node.n + 4 V + 4 chosen link] + children scanned * (8 W + 4 n [+4 prior]);
    backup  sum_levels [8 r + 4 parent link + 2*(8+4) edge W,n rw + 2*4 parent n rw];
    expansion one node row (env state + r 8 + V 4 + n 4 + links 8 [+ priors]) + edge init."""
    per_child = 16 if variant == "discrete" else 12
    expansions = c["sims"] - c["terminal_leaf_sims"]
    exp_bytes = (32 + 8 + 4 + 4 + 8 + 8 + 24) if variant == "discrete" else (16 + 8 + 4 + 4 + 8 + 16)
    return c["levels"] * (12 + 44) + c["children_scanned"] * per_child + expansions * exp_bytes


def cpu_port_throughput(variant: str, N: int, roots: np.ndarray, weights: np.ndarray, seconds: float, threads: int):
    """Time the CPU oracle (C port of the reference search) on a bounded sample; returns (sims/s, sample text)."""
    from oracle import azo
    cfg = oracle_config(variant, N)
    n0 = min(len(roots), 32 * threads)
    t0 = time.perf_counter()
    azo.search(cfg, weights, roots[:n0], dump=False, n_threads=threads)
    rate = n0 * N / (time.perf_counter() - t0)
    n = int(max(n0, min(len(roots), rate * seconds / N)))
    passes = max(1, int(round(rate * seconds / (n * N))))  # the whole batch fits the budget several times: repeat it
    t0 = time.perf_counter()
    for i in range(passes):
        azo.search(cfg, weights, roots[:n], dump=False, n_threads=threads, tree_id0=i * n)
    dt = time.perf_counter() - t0
    return passes * n * N / dt, f"{passes} x first {n} trees x {N} sims of the workload, {threads} threads, {dt:.1f} s"


def run_reference(args, variant, B, N, rank, world):
    """--impl reference: the CPU port on all host cores, each step a bounded sample of the workload."""
    if rank != 0:
        return
    from oracle import azo
    threads = os.cpu_count() or 1
    roots, weights = make_roots(variant, B), make_weights(variant)
    cfg = oracle_config(variant, N)
    # size one step to ~ (120 s budget) / (steps + warmup)
    n0 = min(B, 32 * threads)
    t0 = time.perf_counter()
    azo.search(cfg, weights, roots[:n0], dump=False, n_threads=threads)
    rate = n0 * N / (time.perf_counter() - t0)
    budget = 120.0 / max(1, args.steps + args.warmup)
    n = int(max(n0, min(B, rate * budget / N)))
    for _ in range(args.warmup):
        azo.search(cfg, weights, roots[:n], dump=False, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        azo.search(cfg, weights, roots[:n], dump=False, n_threads=threads)
    dt = time.perf_counter() - t0
    v = n * N * args.steps / dt
    sample = f"{n} of {B} trees x {N} sims per step, {threads} threads"
    line = {"impl": "reference", "metric": "MCTS simulations/sec (batched trees)", "value": v, "unit": "sims/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 tree statistics / f32 network", "data": "synthetic",
            "config": {"workload": args.workload, "trees_per_step": n, "n_rollouts": N,
                       "note": "CPU C port (oracle/azg_oracle.c) of the reference's python search; the python reference cannot travel"},
            "cpu_baseline": {"value": v, "unit": "sims/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


_JSON_OUT = None


def _claim_stdout():
    """stdout carries the ONE JSON line and nothing else: keep a private handle on the real stdout for it and point file descriptor 1 at
    stderr, so that whatever a library prints there (NCCL's version banner under NCCL_DEBUG=VERSION, which ignores NCCL_DEBUG_FILE)
    ends up on stderr."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line: dict) -> None:
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pendulum_65536x100", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fused", default="auto", choices=["auto", "on", "off"],
                    help="whole-search persistent kernel (AZG_FLAG_FUSED); auto = the engine's default")
    ap.add_argument("--eval", default="q8", choices=["fp32", "q8"],
                    help="leaf evaluation arithmetic: int8-sliced tcgen05 products (qmlp2.cuh, default) or FP32 FMA (mlp.cuh)")
    args = ap.parse_args()
    variant, B, N = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, variant, B, N, rank, world)
        return

    import torch
    import torch.distributed as dist
    from alphazero_gym_b200.engine import SearchEngine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries the one JSON line and nothing else: NCCL's own log (its version banner under NCCL_DEBUG=VERSION/INFO) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tree_id0 = rank * B  # global tree ids: shard r owns trees [r*B, (r+1)*B)
    roots_h = make_roots(variant, B * world)[rank * B:(rank + 1) * B].copy()
    weights = make_weights(variant)
    eng = SearchEngine(engine_config(variant, B, N, local, q8=args.eval == "q8", fused={"auto": None, "on": True, "off": False}[args.fused]))
    eng.set_weights(weights)
    roots_d = torch.from_numpy(roots_h).cuda()

    selfplay = args.workload.startswith("selfplay")
    api = "azg_search_host via SearchEngine.search_host, page-locked host buffers (wall clock between syncs)"
    if selfplay:
        # ---- self-play loop (BASELINE config 5): search -> final action -> real env step, replay rows all-gathered every
        # step (C2), weights re-broadcast from rank 0 and re-loaded every SELFPLAY_BROADCAST_EVERY steps (C1)
        from alphazero_gym_b200.selfplay import DeviceReplayBuffer, ROW_KEYS, SelfPlayDriver
        drv = SelfPlayDriver(eng, B, N, max_episode_length=200, tree_id0=tree_id0, total_envs=B * world, seed=34)
        replay = DeviceReplayBuffer(max_size=2 * B * world, batch_size=256, obs_dim=3 if variant == "continuous" else 4,
                                    cmax=eng.cmax, device=eng.device)
        wflat = torch.from_numpy(weights).cuda()
        w_pinned = torch.from_numpy(weights).pin_memory()
        rows_pinned = {k: torch.empty_like(drv.t[k], device="cpu").pin_memory() for k in ROW_KEYS}

        def device_step(i):
            if i % SELFPLAY_BROADCAST_EVERY == 0:
                drv.sync_weights(wflat)
            drv.step(store=False)
            replay.store(drv.gathered_rows())

        def host_step(i):
            if i % SELFPLAY_BROADCAST_EVERY == 0:
                wflat.copy_(w_pinned, non_blocking=True)  # the trainer's new weights arrive from the host
                drv.sync_weights(wflat)
            drv.step(store=False)
            replay.store(drv.gathered_rows())
            for k in ROW_KEYS:  # this step's replay rows leave for a host-side consumer
                rows_pinned[k].copy_(drv.t[k], non_blocking=True)
            torch.cuda.synchronize()

        api = "SelfPlayDriver.step (azg_selfplay_step) + per-step D2H of the replay rows, H2D of the weights at broadcast steps"
        h2d = weights.nbytes / SELFPLAY_BROADCAST_EVERY
        d2h = sum(rows_pinned[k].numel() * rows_pinned[k].element_size() for k in ROW_KEYS)
        extra_launches = 2  # root results + the self-play kernel (+ set_seed)
    else:
        def device_step(i):
            eng.search(roots_d, N, tree_id0=tree_id0)
            return eng.root_results()

        host_out = eng.host_buffers(B)  # page-locked, as the contract asks: results arrive by DMA, no staging copy
        roots_pinned = torch.from_numpy(roots_h).pin_memory().numpy()

        def host_step(i):
            return eng.search_host(roots_pinned, N, tree_id0=tree_id0, out=host_out)

        host_out2 = eng.host_buffers(B)

        def host_pipeline(steps):
            """The same K searches through the pipelined entry points (azg_search_host_begin / _end, two slots in flight): every
            step still copies its roots in from pinned host memory and its results out; the D2H of step i runs under step i + 1."""
            outs = (host_out, host_out2)
            for i in range(steps):
                if i >= 2:
                    eng.search_host_end(i & 1)
                eng.search_host_begin(i & 1, roots_pinned, N, outs[i & 1], tree_id0=tree_id0)
            for i in range(max(0, steps - 2), steps):
                eng.search_host_end(i & 1)
            return outs[(steps - 1) & 1]

        h2d = roots_h.nbytes
        extra_launches = 1  # the root-results kernel

    # ---- device-resident throughput ("value") ---------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        device_step(i)
    eng.status()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        device_step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    eng.status()
    counters = eng.counters()
    launches_per_step = counters["launches"] + extra_launches

    # ---- end to end through the host-facing entry point ("e2e") -------------------------------------
    for i in range(2):
        out = host_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        out = host_step(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_blocking_s = None
    if not selfplay:  # the pipelined entry points: the number reported as e2e; the blocking call's stays next to it
        host_pipeline(2)
        barrier()
        t0 = time.perf_counter()
        out = host_pipeline(args.steps)
        e2e_blocking_s, e2e_s = e2e_s, time.perf_counter() - t0
        api = ("azg_search_host_begin / azg_search_host_end via SearchEngine.search_host_begin / _end, two searches in flight, page-locked "
               "host buffers (wall clock from the first begin to the last end); blocking_value = azg_search_host, one call per step")
    clk = clocks.stop()
    checksum = int(drv.t["counts"].sum()) if selfplay else int(out["counts"].sum())
    assert checksum == B * N, f"visit counts do not add up: {checksum} != {B * N}"
    if selfplay:
        roots_d = drv.env_state.clone()
    else:
        d2h = sum(out[k].nbytes for k in ("actions", "counts", "Q", "V_target", "n_children"))

    # ---- per-kernel times, live, for the roofline ----------------------------------------------------
    eng.profile_search(roots_d, N, tree_id0=tree_id0)
    prof = eng.profile_search(roots_d, N, tree_id0=tree_id0)
    pc = eng.counters()
    hbm_peak, sm_max_mhz, bf16_peak, peak_src = measured_peaks()
    sm_count = torch.cuda.get_device_properties(local).multi_processor_count
    ev, tr = prof["evaluation"], prof["tree_step"]
    fp32_peak = sm_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    ev_avg_ms = ev["ms"] / max(1, ev["launches"])
    tr_avg_ms = tr["ms"] / max(1, tr["launches"])
    ev_tflops = FLOP_PER_EVAL[variant] * B / (ev_avg_ms * 1e-3) / 1e12
    tr_bytes = algorithmic_bytes(variant, pc) / max(1, tr["launches"])
    tr_gbs = tr_bytes / (tr_avg_ms * 1e-3) / 1e9
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get(args.workload, {})
    ev_share = ev["ms"] / (ev["ms"] + tr["ms"] + prof["setup"]["ms"])
    if args.eval == "q8":
        # tensor-core evaluation: algorithmic FLOPs (2 x MACs of the network) against the measured dense bf16 peak.
        # The exact fixed-point arithmetic issues 6 int8 MMAs per algorithmic H x H product, and the kernel is bound by its
        # CUDA-core epilogue (dequantise, activation, quantise), not by the tensor pipe -- DESIGN.md section 4.2.
        hh = 2 * 128 * 128 * ((3 if variant == "continuous" else 2) - 1)
        roof_eval = {"kernel": "k_qmlp2 (leaf evaluation, tcgen05 kind::i8)", "bound": "tensor", "achieved": ev_tflops, "peak": bf16_peak,
                     "unit": "TFLOP/s", "frac": ev_tflops / bf16_peak, "traffic": traffic.get("k_qmlp2"),
                     "peak_source": peak_src + " dense bf16; int8 digits: 6 MMAs per algorithmic product",
                     "avg_launch_ms": ev_avg_ms, "launches": ev["launches"], "flop_per_launch": FLOP_PER_EVAL[variant] * B,
                     "tensor_ops_issued_per_launch": 6 * hh * B, "tensor_tops_issued": 6 * hh * B / (ev_avg_ms * 1e-3) / 1e12,
                     "frac_of_fp32_cuda_core_peak": ev_tflops / fp32_peak, "share_of_step": ev_share}
    else:
        roof_eval = {"kernel": "k_mlp (leaf evaluation)", "bound": "fp32", "achieved": ev_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                     "frac": ev_tflops / fp32_peak, "traffic": traffic.get("k_mlp"),
                     "peak_source": f"{sm_count} SMs x 128 FMA/clk x 2 x {sm_max_mhz:.0f} MHz (CUDA-core FP32)",
                     "avg_launch_ms": ev_avg_ms, "launches": ev["launches"], "flop_per_launch": FLOP_PER_EVAL[variant] * B,
                     "share_of_step": ev_share}
    roof_tree = {"kernel": "k_step (backup + select + expansion/env step)", "bound": "hbm", "achieved": tr_gbs, "peak": hbm_peak,
                 "unit": "GB/s", "frac": tr_gbs / hbm_peak, "traffic": traffic.get("k_step"), "peak_source": peak_src,
                 "avg_launch_ms": tr_avg_ms, "launches": tr["launches"], "algorithmic_bytes_per_launch": tr_bytes,
                 "algorithmic_bytes_per_sim": algorithmic_bytes(variant, pc) / max(1, pc["sims"]),
                 "share_of_step": tr["ms"] / (ev["ms"] + tr["ms"] + prof["setup"]["ms"])}
    dominant = roof_eval if ev["ms"] >= tr["ms"] else roof_tree
    roof_all = {"evaluation": roof_eval, "tree_step": roof_tree}
    if eng.cfg.is_fused():
        # The search runs as ONE persistent kernel (AZG_FLAG_FUSED); the per-kernel numbers above are those of the same search as one
        # launch per kernel and simulation (azg_profile_search always takes that path) and stay in roofline_all for comparison.
        eng.fused_stats()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, args.steps)
        f0.record()
        for _ in range(reps):
            eng.search(roots_d, N, tree_id0=tree_id0)
        f1.record()
        torch.cuda.synchronize()
        k_ms = f0.elapsed_time(f1) / reps / max(1, counters["launches"])
        st = eng.fused_stats()
        fc = eng.counters()
        flop = FLOP_PER_EVAL[variant] * fc["evals"] / max(1, counters["launches"])
        tflops = flop / (k_ms * 1e-3) / 1e12
        ev_ms, tree_ms = k_ms * (1.0 - st["tree_phase"]), k_ms * st["tree_phase"]
        tbytes = algorithmic_bytes(variant, fc) / max(1, counters["launches"])
        hh = 2 * 128 * 128 * ((3 if variant == "continuous" else 2) - 1)
        dominant = {
            "kernel": "k_qmlp2<FUSED> (whole search in one persistent kernel: per simulation an evaluation phase on tcgen05 kind::i8 "
                      "and a tree phase, backup + select + expansion, one thread per tree)",
            "bound": "tensor", "achieved": tflops, "peak": bf16_peak, "unit": "TFLOP/s", "frac": tflops / bf16_peak,
            "traffic": traffic.get("k_search"), "peak_source": peak_src + " dense bf16; int8 digits: 6 MMAs per algorithmic product",
            "avg_launch_ms": k_ms, "launches": counters["launches"], "flop_per_launch": flop,
            "tensor_tops_issued": 6 * hh * fc["evals"] / max(1, counters["launches"]) / (k_ms * 1e-3) / 1e12,
            "frac_of_fp32_cuda_core_peak": tflops / fp32_peak, "share_of_step": 1.0,
            "phases": {
                "evaluation": {"share": 1.0 - st["tree_phase"], "ms_per_launch": ev_ms, "tflops_in_phase": flop / (ev_ms * 1e-3) / 1e12,
                               "frac_of_tensor_peak_in_phase": flop / (ev_ms * 1e-3) / 1e12 / bf16_peak,
                               "frac_of_fp32_cuda_core_peak_in_phase": flop / (ev_ms * 1e-3) / 1e12 / fp32_peak},
                "tree": {"share": st["tree_phase"], "ms_per_launch": tree_ms, "algorithmic_bytes_per_launch": tbytes,
                         "hbm_gbs_in_phase": tbytes / (tree_ms * 1e-3) / 1e9, "frac_of_hbm_peak_in_phase": tbytes / (tree_ms * 1e-3) / 1e9 / hbm_peak,
                         "bound": "dependent-load latency chain + L1 tag lookups of per-thread scattered accesses (profiles/README.md r1f/r1g)"},
                "source": "in-kernel cycle counters (azg_fused_stats)"},
        }
        roof_all = {"whole_search_kernel": dominant, "per_simulation_launches": {"evaluation": roof_eval, "tree_step": roof_tree}}

    # ---- CPU baseline (rank 0, N=1 only) ------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, sample = cpu_port_throughput(variant, N, roots_h, weights, args.cpu_seconds, threads)
        if selfplay:
            sample += " (searches only; the env step and bookkeeping are negligible next to them)"
        cpu = {"value": v, "unit": "sims/s", "cores": threads, "kind": "port", "sample": sample}

    # ---- max over ranks ---------------------------------------------------------------------------------------
    t = torch.tensor([ms, e2e_s * 1e3, (e2e_blocking_s or 0.0) * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max, e2e_blocking_ms_max = float(t[0]), float(t[1]), float(t[2])
    total_sims = world * B * N * args.steps
    if rank == 0:
        line = {
            "metric": "MCTS simulations/sec (batched trees)", "value": total_sims / (ms_max * 1e-3), "unit": "sims/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 tree statistics + env dynamics / " + ("f32 network with exact int8-sliced tensor-core products" if args.eval == "q8" else "f32 network"),
            "data": "synthetic",
            "config": {"workload": args.workload, "eval": args.eval, "whole_search_kernel": bool(eng.cfg.is_fused()), "env": "Pendulum-v0" if variant == "continuous" else "CartPole-v0",
                       "trees_per_gpu": B, "global_trees": B * world, "n_rollouts": N, "parallelism": f"tree-sharded x{world}, no data-path collective",
                       "weights": "default init, torch.manual_seed(34)", "roots": "numpy default_rng(34)",
                       "l2": "node tables per GPU (%.0f MB) exceed the 126 MB L2; no explicit flush" % (eng.rows * B * (32 + 16 + 32) / 1e6)
                       if variant == "continuous" else "tables are L2-resident at this size (SURVEY 8d config 3); no explicit flush"},
            "e2e": {"value": total_sims / (e2e_ms_max * 1e-3), "unit": "sims/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms_max / args.steps, "api": api,
                    **({"blocking_value": total_sims / (e2e_blocking_ms_max * 1e-3)} if e2e_blocking_ms_max > 0 else {})},
            "gpu_launches": launches_per_step * args.steps,
            **({"selfplay": {"max_episode_length": 200, "weight_broadcast_every_steps": SELFPLAY_BROADCAST_EVERY,
                             "collectives_per_step": "all_gather of 5 replay-row tensors (C2); broadcast of the flat weights every k steps (C1)",
                             "backend": "nccl" if world > 1 else "none (1 GPU)"}} if selfplay else {}),
            "clocks": clk,
            "roofline": dominant,
            "roofline_all": roof_all,
            "counters_per_sim": {k: pc[k] / max(1, pc["sims"]) for k in ("levels", "children_scanned", "pw_inserts", "evals")},
            "cpu_baseline": cpu,
        }
        _emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
