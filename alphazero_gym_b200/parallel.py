"""Multi-GPU plumbing: one process per GPU, trees sharded, no collective inside the search.

Trees are independent, so N GPUs run N shards of the batch.  A tree's result depends only on its
GLOBAL id (Philox streams are keyed by tree_id0 + i), never on which rank or batch it ran in.  The only
exchanges are outside the search: (C1) broadcast of the flat weight vector from the trainer rank and (C2)
all-gather of fixed-size root-result rows for the replay buffer.  Backend: "nccl" on GPUs (NVLink/NVSwitch),
"gloo" in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total_trees: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous tree-id range [lo, hi) of `rank`; sizes differ by at most one."""
    base, rem = divmod(total_trees, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_weights(flat: torch.Tensor, src: int = 0) -> torch.Tensor:
    """(C1) weight broadcast: 70 KB (CartPole net) / 138 KB (Pendulum net) of f32, latency-bound."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src=src)
    return flat


def allgather_results(local: Dict[str, torch.Tensor], total_trees: int) -> Dict[str, torch.Tensor]:
    """(C2) gather every rank's root-result rows into global tree order.  Ranks may own different numbers of
    trees (shard_range), so rows are padded to the largest shard for the collective and trimmed after."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(total_trees, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    out: Dict[str, torch.Tensor] = {}
    for k, v in local.items():
        buf = torch.zeros((pad,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
        buf[: v.shape[0]] = v
        gathered = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(gathered, buf)
        out[k] = torch.cat([g[: hi - lo] for g, (lo, hi) in zip(gathered, sizes)], 0)
    return out


class PackedRows:
    """The tensors of one step's replay rows as views into ONE contiguous byte buffer per rank, so that (C2) is a single
    `all_gather_into_tensor` of that buffer: no per-tensor collectives, no packing copy on the way in, no pad-and-trim.

    Layout per rank (blocks 16-byte aligned): for every key in order, the [B, ...] tensor of that key, row-major.  After the
    gather the buffer of rank r sits at [r]; `gathered()` returns per-key views of shape [world, B, ...] (global environment
    order is rank-major because shards are contiguous tree-id ranges of equal size)."""

    def __init__(self, spec, B: int, device):
        # spec: sequence of (key, trailing shape tuple, dtype)
        self.B, self.device = int(B), torch.device(device)
        self.spec, self.off, off = list(spec), {}, 0
        for key, shape, dtype in self.spec:
            n = int(np.prod((B,) + tuple(shape))) * torch.empty((), dtype=dtype).element_size()
            self.off[key] = (off, n)
            off += (n + 15) // 16 * 16
        self.nbytes = off
        self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=self.device)
        self.views = {key: self.buf[self.off[key][0]:self.off[key][0] + self.off[key][1]].view(dtype).view((B,) + tuple(shape))
                      for key, shape, dtype in self.spec}
        self._out = None

    def gathered(self) -> Dict[str, torch.Tensor]:
        """(C2) one collective for all keys; returns [world * B, ...] tensors in global environment order."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return dict(self.views)
        world = dist.get_world_size()
        if self._out is None or self._out.numel() != world * self.nbytes:
            self._out = torch.empty(world * self.nbytes, dtype=torch.uint8, device=self.device)  # flat: gloo insists on it
        dist.all_gather_into_tensor(self._out, self.buf)
        g = self._out.view(world, self.nbytes)
        out = {}
        for key, shape, dtype in self.spec:
            o, n = self.off[key]
            out[key] = g[:, o:o + n].view(dtype).reshape((world * self.B,) + tuple(shape))
        return out


def results_to_torch(res: Dict[str, np.ndarray], device="cpu") -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in res.items()}
