"""Hardware test of the tree-sharded N > 1 path: rank r's rows equal rows [r B, (r + 1) B) of a one-GPU run and the
CPU oracle, bit for bit.  Launched like bench.py is (torch.distributed.run, one process per rank, 127.0.0.1 rendezvous); uses one
GPU per rank over NCCL when the box has them, otherwise both ranks share cuda:0 and gather over gloo."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


# 3000 trees per rank: the two-phase whole-search kernel; 57000 per rank: the warpgroup kernel (engine.cu launch_fused_t)
@pytest.mark.parametrize("B,N", [(3000, 25), (57000, 40)])
def test_two_rank_sharded_search_equals_single_gpu_and_oracle(B, N):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "mp_shard_worker.py"), str(B), str(N)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "OK world=2" in res.stdout
