"""2-GPU data-parallel step == single-GPU step on the global batch (needs >= 2 GPUs; skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_step_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(out.stdout[-3000:], out.stderr[-2000:])
    assert out.returncode == 0 and "DP_CHECK OK" in out.stdout
