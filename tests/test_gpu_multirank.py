"""-m gpu, needs >= 2 GPUs (skipped otherwise): the particle-sharded launcher over real NCCL ranks.  SURVEY.md section 8(e):
shards are contiguous row blocks of a globally defined, chunk-seeded x0, so the concatenated outputs of G ranks must be
bit-identical to one GPU solving all rows; the MMD collectives (one all_gather of the evaluation rows, one of three fp64
sums) must reproduce the single-process value."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_nccl_ranks_reproduce_one_gpu(cuda_device):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    ws = 2 if n < 4 else 4
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ws}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "nccl_shard_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTIRANK_OK" in out.stdout, out.stdout[-2000:]
