"""Multi-rank GPU correctness (NCCL, one process per GPU; skipped on boxes with fewer than 2 GPUs): sharded-batch
training reproduces the single-process update and keeps all replicas bit-identical."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("reduce_kind", ["peer_memory", "nccl"])
@pytest.mark.parametrize("precision,mode", [("fp32", "eager"), ("bf16", "graph")])
def test_two_rank_nccl_matches_single_process(precision, mode, reduce_kind):
    """reduce_kind: the NVLink peer-memory all-reduce kernel (pcrl_p2p_allreduce, the default on one node) or NCCL."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_nccl_worker.py"), precision, mode, reduce_kind]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "NCCL_2RANK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_peer_memory_allreduce_kernel():
    """pcrl_p2p_allreduce alone (all GPUs of the box, at least 2): ragged ranges, repeated calls, two channels in flight,
    graph replay; result equals NCCL's and is bit-identical on all ranks."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    port = 29300 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 8)), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_p2p_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "P2P_ALLREDUCE PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
