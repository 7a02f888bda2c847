"""Global fit sharded over 2 GPUs: NCCL host loop and the fused peer-memory exchange
inside the persistent kernel, through tests/multi_gpu_check.py under torchrun.  Needs >= 2 GPUs
(skipped on a single-GPU box; run by hand with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
def test_global_fit_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(HERE, "multi_gpu_check.py"), "200001"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTI_GPU_OK world=2" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
