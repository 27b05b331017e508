"""GPU, >= 2 devices: NCCL data-parallel step equals the single-process step on the concatenated batch."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_step_matches_single_rank(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(29600 + os.getpid() % 300), os.path.join(ROOT, 'scripts', 'multi_gpu_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert 'MULTI_GPU_OK' in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]

