"""N>1 path on CPU: two processes over gloo shard a batch of trees, each runs its shard with the global
tree ids of its range, the root-result rows are all-gathered and must equal the single-process run bit for bit.
The per-shard searcher here is the CPU oracle (the CUDA engine takes its place on GPUs; bench.py --gpus N
uses the same shard_range / tree_id0 logic).  Also covers the weight broadcast."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import azo, gen_golden as G
from alphazero_gym_b200.parallel import allgather_results, broadcast_weights, results_to_torch, shard_range

KEYS = ("actions", "counts", "Q", "V_target", "n_children")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg, g = G.load("pendulum_n25_k2")
        roots = G.pendulum_roots(total, seed=11)
        # (C1) rank 0 owns the weights; the others start from zeros
        w = torch.from_numpy(g["weights"].copy()) if rank == 0 else torch.zeros(len(g["weights"]))
        broadcast_weights(w, src=0)
        assert np.array_equal(w.numpy(), g["weights"])
        lo, hi = shard_range(total, rank, world)
        local = azo.search(cfg, w.numpy(), roots[lo:hi], tree_id0=lo, dump=False)
        gathered = allgather_results(results_to_torch({k: local[k] for k in KEYS}), total)  # (C2)
        if rank == 0:
            np.savez(out_path, **{k: v.numpy() for k, v in gathered.items()})
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_search_equals_single_process(tmp_path):
    total, world = 21, 2  # odd on purpose: ranks own 11 and 10 trees
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, _free_port(), total, out), nprocs=world, join=True)
    got = np.load(out)
    cfg, g = G.load("pendulum_n25_k2")
    ref = azo.search(cfg, g["weights"], G.pendulum_roots(total, seed=11), tree_id0=0, dump=False)
    for k in KEYS:
        assert np.array_equal(got[k], ref[k]), k


def _selfplay_worker(rank, world, port, total, steps, out_path):
    """Self-play sharded over two ranks: each rank advances its environments (global ids keep the streams), the replay rows
    of every step are all-gathered (C2) into a DeviceReplayBuffer on CPU tensors."""
    from oracle import selfplay as osp
    from alphazero_gym_b200.selfplay import DeviceReplayBuffer, ROW_KEYS, initial_states
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg, g = G.load("pendulum_n25_k2")
        cfg.use_eval_tape = 0
        lo, hi = shard_range(total, rank, world)
        states0 = initial_states(cfg.variant, hi - lo, seed=3, tree_id0=lo, total=total)
        recs = osp.run(cfg, g["weights"], states0, steps, 2, seed=cfg.seed, tree_id0=lo, n_threads=1)
        rb = DeviceReplayBuffer(max_size=total * steps, batch_size=8, obs_dim=3, cmax=cfg.cmax, device="cpu")
        for r in recs:
            rb.store(allgather_results(results_to_torch({k: r[k] for k in ROW_KEYS}), total))
        if rank == 0:
            np.savez(out_path, **{k: rb.data[k][: len(rb)].numpy() for k in ROW_KEYS})
    finally:
        dist.destroy_process_group()


def test_two_rank_selfplay_replay_equals_single_process(tmp_path):
    from oracle import selfplay as osp
    from alphazero_gym_b200.selfplay import ROW_KEYS, initial_states
    total, world, steps = 13, 2, 3
    out = str(tmp_path / "replay.npz")
    mp.spawn(_selfplay_worker, args=(world, _free_port(), total, steps, out), nprocs=world, join=True)
    got = np.load(out)
    cfg, g = G.load("pendulum_n25_k2")
    cfg.use_eval_tape = 0
    recs = osp.run(cfg, g["weights"], initial_states(cfg.variant, total, seed=3), steps, 2, seed=cfg.seed, n_threads=1)
    for k in ROW_KEYS:
        assert np.array_equal(got[k], np.concatenate([r[k] for r in recs], 0)), k


class _FakeEngine:
    """Stands in for SearchEngine.set_weights on the CPU (the CUDA engine takes its place on GPUs)."""

    def __init__(self):
        self.w = None

    def set_weights(self, flat):
        self.w = flat.detach().clone()


def _train_worker(rank, world, port, out_path):
    """Rank 0 trains (one reference-pinned update on its replay batch), then every rank calls Trainer.push_weights: the flat weights
    are broadcast from rank 0 (C1) and loaded into each rank's engine, so all ranks search with the trainer's new weights."""
    from alphazero_gym_b200.network import PolicyNet
    from alphazero_gym_b200.train import LossConfig, Trainer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_a0c_adam_clip.npz"))
        torch.set_num_threads(1)
        net = PolicyNet(3, 128, 3, 6, "elu", num_components=2).load_flat(g["w0"])
        tr = Trainer(net, LossConfig(tuned=False, tau=0.1, policy_coeff=1, value_coeff=1, alpha=1), optimizer="adam", grad_clip=0.5)
        if rank == 0:
            for s in range(3):
                tr.update(dict(obs=torch.from_numpy(g[f"states_{s}"]), actions=torch.from_numpy(g[f"actions_{s}"]),
                               counts=torch.from_numpy(g[f"counts_{s}"]), V_target=torch.from_numpy(g[f"V_{s}"])))
        eng = _FakeEngine()
        tr.push_weights(eng)
        np.save(out_path + f".{rank}.npy", eng.w.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_trainer_pushes_rank0_weights_everywhere(tmp_path):
    out = str(tmp_path / "w")
    mp.spawn(_train_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    w0, w1 = np.load(out + ".0.npy"), np.load(out + ".1.npy")
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_a0c_adam_clip.npz"))
    assert np.array_equal(w0, w1)                       # rank 1 never trained: it holds rank 0's weights
    assert np.abs(w0 - g["w_3"]).max() <= 2e-6          # = the reference's weights after the same three updates
    assert np.abs(w0 - g["w0"]).max() > 1e-4


def _packed_worker(rank, world, port, B, out_path):
    """(C2) as ONE collective: the step's replay rows live in one byte buffer per rank (parallel.PackedRows) and a single
    all_gather_into_tensor moves them; the gathered views must equal the per-tensor gather in global environment order."""
    from alphazero_gym_b200.parallel import PackedRows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cm, S = 5, 3
        pk = PackedRows([("Q", (cm,), torch.float64), ("V_target", (), torch.float64), ("obs", (S,), torch.float32),
                         ("actions", (cm,), torch.float32), ("counts", (cm,), torch.int32)], B, "cpu")
        rng = np.random.default_rng(100 + rank)
        pk.views["Q"].copy_(torch.from_numpy(rng.standard_normal((B, cm))))
        pk.views["V_target"].copy_(torch.from_numpy(rng.standard_normal(B)))
        pk.views["obs"].copy_(torch.from_numpy(rng.standard_normal((B, S)).astype(np.float32)))
        pk.views["actions"].copy_(torch.from_numpy(rng.standard_normal((B, cm)).astype(np.float32)))
        pk.views["counts"].copy_(torch.from_numpy(rng.integers(0, 25, (B, cm)).astype(np.int32)))
        one = pk.gathered()
        five = allgather_results({k: v.clone() for k, v in pk.views.items()}, world * B)
        for k in one:
            assert one[k].shape == five[k].shape and torch.equal(one[k], five[k]), k
        if rank == 0:
            np.savez(out_path, **{k: v.numpy() for k, v in one.items()})
    finally:
        dist.destroy_process_group()


def test_packed_rows_single_collective_equals_per_tensor_gather(tmp_path):
    out = str(tmp_path / "packed.npz")
    B = 7  # odd: the blocks of the byte buffer need their 16-byte padding
    mp.spawn(_packed_worker, args=(2, _free_port(), B, out), nprocs=2, join=True)
    got = np.load(out)
    rng0, rng1 = np.random.default_rng(100), np.random.default_rng(101)
    assert np.array_equal(got["Q"], np.concatenate([rng0.standard_normal((B, 5)), rng1.standard_normal((B, 5))]))
    assert got["counts"].shape == (2 * B, 5) and got["obs"].dtype == np.float32
