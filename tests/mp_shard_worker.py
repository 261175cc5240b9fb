"""Worker of tests/test_multigpu_gpu.py (launched with torch.distributed.run, one process per rank).

Every rank searches its contiguous shard of a batch of Pendulum trees on ITS GPU (tensor-core evaluation, whole-search kernel,
global tree ids rank * B ...), the root-result rows are all-gathered, and rank 0 compares them bit for bit with (1) the same batch
searched as ONE batch on one GPU and (2) the CPU oracle.  With fewer GPUs than ranks (the driver's single-GPU test box) the ranks
share cuda:0 and gather over gloo; with one GPU per rank the collective is NCCL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    B, N = int(sys.argv[1]), int(sys.argv[2])
    ngpu = torch.cuda.device_count()
    nccl = ngpu >= world
    dev = rank if nccl else 0
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if nccl else "gloo", **({"device_id": torch.device("cuda", dev)} if nccl else {}))
    import bench
    from oracle import azo
    from alphazero_gym_b200.engine import SearchEngine
    from alphazero_gym_b200.parallel import allgather_results, shard_range
    total = B * world
    roots = bench.make_roots("continuous", total)
    w = bench.make_weights("continuous")
    lo, hi = shard_range(total, rank, world)
    assert (lo, hi) == (rank * B, (rank + 1) * B)
    eng = SearchEngine(bench.engine_config("continuous", B, N, dev, q8=True))
    eng.set_weights(w)
    eng.search(torch.from_numpy(roots[lo:hi].copy()).cuda(), N, tree_id0=lo)
    eng.status()
    local = eng.root_results()
    eng.close()
    if not nccl:
        local = {k: v.cpu() for k, v in local.items()}
    gathered = allgather_results(local, total)
    if rank == 0:
        got = {k: v.cpu().numpy() for k, v in gathered.items()}
        eng = SearchEngine(bench.engine_config("continuous", total, N, dev, q8=True))
        eng.set_weights(w)
        eng.search(torch.from_numpy(roots).cuda(), N, tree_id0=0)
        eng.status()
        one = {k: v.cpu().numpy() for k, v in eng.root_results().items()}
        eng.close()
        for k in got:
            assert np.array_equal(got[k], one[k]), f"{k}: {world}-rank result differs from the single-GPU run"
        n = min(total, 2048)
        sel = np.r_[0:n // 2, total - n // 2:total]  # trees of the first and of the last rank
        cfg = bench.oracle_config("continuous", N, q8=True)
        for part, t0 in ((sel[:n // 2], 0), (sel[n // 2:], total - n // 2)):
            ref = azo.search(cfg, w, roots[part], tree_id0=t0, dump=False, n_threads=os.cpu_count() or 4)
            c = ref["counts"].shape[1]
            for k in ("counts", "actions", "Q", "V_target", "n_children"):
                a = got[k][part]
                assert np.array_equal(a[:, :c] if a.ndim == 2 else a, ref[k]), f"{k}: differs from the oracle"
        print(f"OK world={world} backend={'nccl' if nccl else 'gloo'} trees={total} sims={N}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
