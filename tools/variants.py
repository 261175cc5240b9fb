#!/usr/bin/env python
"""Kernel experiments: build several variants of libazg.so (extra -D flags) and time the whole-search kernel with each.

    python tools/variants.py build name1=-DFOO name2="-DBAR -DBAZ=3" ...     (build container: nvcc only)
    python tools/variants.py run [workload ...]                              (GPU box: one subprocess per variant, AZG_LIB_PATH)

Variants land in alphazero_gym_b200/lib/variants/ (git-ignored, travels with gpurun).  `run` prints one JSON line per
(variant, workload): sims/s, in-kernel phase split (azg_fused_stats) and a checksum of every root result, so that a variant
whose results differ from the shipped library is visible at once."""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "alphazero_gym_b200", "lib", "variants")


def build(specs):
    from alphazero_gym_b200 import build as B
    os.makedirs(VDIR, exist_ok=True)
    procs = []
    for spec in specs:
        name, _, flags = spec.partition("=")
        out = os.path.join(VDIR, f"libazg_{name}.so")
        cmd = ["nvcc"] + B.NVCC_FLAGS + flags.split() + ["-o", out, os.path.join(B.CSRC, "engine.cu")]
        procs.append((name, out, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, out, p in procs:
        log, _ = p.communicate()
        with open(out + ".log", "w") as f:
            f.write(log)
        info = []
        lines = log.splitlines()
        for i, l in enumerate(lines):
            if any(k in l for k in ("Compiling entry function '_Z7k_qmlp2ILi3ELi1ELi2ELb1ELb0E", "Compiling entry function '_Z7k_qmlp2ILi4ELi0ELi1ELb1ELb1E",
                                    "Compiling entry function '_Z11k_search_wgILi3ELi1ELi2E")):
                info.append(" | ".join(x.strip().replace("ptxas info    : ", "") for x in lines[i + 1:i + 4]))
        print(name, "rc", p.returncode, *info, sep="\n   ")


def run_one(workload, reps):
    import numpy as np
    import torch
    import bench
    from alphazero_gym_b200.engine import SearchEngine
    variant, B, N = bench.WORKLOADS[workload]
    eng = SearchEngine(bench.engine_config(variant, B, N, 0, q8=True, fused=True))
    eng.set_weights(bench.make_weights(variant))
    roots = torch.from_numpy(bench.make_roots(variant, B)).cuda()
    eng.search(roots, N)
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
    eng.status()
    res = eng.root_results()
    h = hashlib.sha1()
    for k in ("counts", "actions", "Q", "V_target", "n_children"):
        h.update(res[k].cpu().numpy().tobytes())
    us = ms * 1e3 / (N + 1)
    print(json.dumps(dict(lib=os.path.basename(os.environ.get("AZG_LIB_PATH", "libazg.so")), workload=workload, msims=round(B * N / ms / 1e3, 1),
                          us_per_sim=round(us, 2), eval_us=round(us * (1 - st["tree_phase"]), 2), tree_us=round(us * st["tree_phase"], 2),
                          wait_post=round(st["wait_for_post_processing"], 4), extra={k: v for k, v in st.items() if k.startswith("x_")},
                          sha=h.hexdigest()[:12])), flush=True)
    eng.close()


def run(workloads):
    libs = [None] + sorted(os.path.join(VDIR, f) for f in os.listdir(VDIR) if f.endswith(".so")) if os.path.isdir(VDIR) else [None]
    only = os.environ.get("VARIANTS")
    for lib in libs:
        if only and lib and not any(o in lib for o in only.split(",")):
            continue
        env = dict(os.environ)
        if lib:
            env["AZG_LIB_PATH"] = lib
        for wl in workloads:
            subprocess.run([sys.executable, os.path.abspath(__file__), "_one", wl], env=env)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "run":
        run(sys.argv[2:] or ["pendulum_65536x100"])
    elif sys.argv[1] == "_one":
        run_one(sys.argv[2], 3)
