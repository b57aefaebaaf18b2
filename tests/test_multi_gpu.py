"""Multi-GPU tests (`-m gpu`, need >= 2 visible GPUs: `gpurun --gpus 2 -- python -m pytest tests -m gpu`).

One process per GPU through torch.distributed.run, like bench.py.  Skipped on a single-GPU box."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_shard_equivalence_and_fused_allgather(world):
    """B_global instances seeded / driven by global index, sharded over `world` GPUs, gathered by the fused step +
    all-gather (peer stores over NVLink) and by NCCL: bit-identical to the single-GPU batch (tests/mgpu_worker.py)."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_worker.py"), "1024", "40"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, (res.stdout[-3000:], res.stderr[-3000:])
    assert "MGPU_OK world=%d" % world in res.stdout
