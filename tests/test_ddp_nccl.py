"""NCCL gradient-equality test of the data-parallel path: needs >= 2 GPUs on the box (skipped otherwise;
`gpurun --gpus 2 -- python -m pytest tests/test_ddp_nccl.py -m gpu`).  The work is in tests/ddp_nccl_worker.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_rank_nccl_gradients_equal_single_gpu():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29641', os.path.join(ROOT, 'tests', 'ddp_nccl_worker.py')]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'DDP_NCCL_OK' in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
