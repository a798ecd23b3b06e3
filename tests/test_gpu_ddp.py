"""-m gpu, needs >= 2 GPUs: 2-rank NCCL DDP over the fused fp16 layer == single-GPU gradient of the combined batch
(multi-GPU correctness of the PRODUCT was untested; tests/test_dist_gloo.py covers the oracle on CPU)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.gpu
def test_two_rank_ddp_matches_single_gpu_gradient():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run under gpurun --gpus 2)')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'ddp_worker.py')]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert 'DDP_PARITY_OK world=2' in p.stdout, p.stdout[-2000:]
