"""NVLink peer-memory exchange (cor_b200/peer.py, csrc/peer.cu) against the NCCL collectives it replaces.  The NCCL
comparison needs two GPUs; on a single-GPU box it is skipped and the one-GPU emulation below still drives every peer
kernel (two processes sharing cuda:0 over CUDA IPC, gloo for the plumbing, results checked against CPU copies and
the ATen port).  The host-side logic of the multi-rank step is also covered by tests/test_dist_cpu.py under gloo."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_peer_exchange_matches_nccl():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_peer_worker.py")]
    env = dict(os.environ, NCCL_DEBUG="WARN")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "PEER_OK" in r.stdout, r.stdout[-3000:] + "\n" + r.stderr[-3000:]


def test_peer_exchange_two_ranks_on_one_gpu():
    """World size 2 on ONE device: gather / reduce / wait_exit kernels, the fused step with gathered negatives against the
    ATen port on both ranks' inputs, and graph replay interleaved with an eager forward-only step."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_peer_worker_1gpu.py")]
    env = dict(os.environ, COR_PEER_TIMEOUT_S="30")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=420)
    assert r.returncode == 0 and "PEER1GPU_OK" in r.stdout, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
