"""Mesh-partitioned mode: 2 or 4 ranks on a multi-GPU box (`gpurun --gpus 2`); on a 1-GPU box the same worker runs with
ONE rank, which still drives the whole partitioned code path — local reduction to one relation, the peer-memory push /
wait kernels (a rank pushes into its own exchange buffer), the NCCL flavour, the interface solve, CUDA-graph replay — so
the path is never skipped for lack of a second GPU."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partitioned_newton_matches_single_gpu():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 1:
        pytest.skip("needs a GPU")
    world = 1 if ngpu < 2 else 2 if ngpu < 4 else 4
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "partition_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stderr[-3000:]
    assert "PARTITION_OK" in r.stdout
