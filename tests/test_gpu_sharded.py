"""Multi-rank GPU test of the document-sharded /retrieve: one process per GPU under
torch.distributed.run, each holding a contiguous shard; the peer-memory exchange (krag_p2p_*), the NCCL all-gather path and
the single-shard oracle over the whole corpus must agree bit for bit (scripts/sharded_gpu_check.py does the asserting).
Runs with 1 rank on a single-GPU box and with 2 (and 4 when present) ranks where the box has them."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, script="sharded_gpu_check.py", marker="sharded gpu check ok"):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", script)]
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-4000:]
    assert f"{marker}: world={world}" in p.stdout, p.stdout[-2000:]


@pytest.mark.parametrize("world", [1, 2, 4])
def test_sharded_retrieve_equals_single_shard_oracle(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
    _run(world)


@pytest.mark.parametrize("world", [1, 2, 4])
def test_sharded_service_equals_single_gpu_service(world):
    """the multi-GPU SERVICE (kaito_b200.sharded_engine: round-robin shards, rank 0 = HTTP host + coalescer, workers on the other
    GPUs) returns the single-GPU service's ids and scores -- scripts/sharded_service_check.py asserts it end to end"""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
    _run(world, "sharded_service_check.py", "sharded service check ok")
