"""-m gpu, needs >= 2 GPUs on the box (skipped otherwise): tensor-parallel engine vs the single-GPU engine.
One process per GPU under torchrun; the ranks talk through CUDA IPC peer memory inside the kernels."""
import os
import subprocess
import sys

import pytest
import torch

from gpu_util import gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@gpu
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.timeout(900)
def test_tensor_parallel_matches_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(HERE, "tp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "TP_RESULT" in r.stdout
