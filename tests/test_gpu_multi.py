"""Two-GPU tier (skipped on a one-GPU box): the peer-memory all-reduce kernels of csrc/peer.cu against NCCL, and the
synced-BatchNorm trainer -- 2 ranks x 32 with all-reduced statistics against ONE GPU at batch 64 (losses, gradients),
captured graphs against eager steps, replicated parameters identical on both ranks (tests/multi_gpu_worker.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_allreduce_and_synced_batchnorm_on_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "multi_gpu_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out.decode(errors="replace"))
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-3000:] for o in outs)
    assert "MULTI_OK" in outs[0], outs[0][-3000:]
