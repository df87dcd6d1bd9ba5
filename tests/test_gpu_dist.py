"""Multi-GPU invariance as a -m gpu test: spawns tests/dist_gpu_check.py under torchrun with 2 ranks when the box has at
least two GPUs (sharded == single-GPU bit for bit: plain fit, temporal fit with the in-kernel NVLink halo inside one CUDA
graph, temporal fit with the host-driven NCCL fallback).  On a one-GPU box the multi-process part is skipped; the halo
kernels themselves are covered on one GPU by test_gpu_parity.py::test_nvlink_halo_shards_of_one_gpu."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('ranks', [2])
def test_sharded_fit_equals_single_gpu(ranks):
    if torch.cuda.device_count() < ranks:
        pytest.skip('needs %d GPUs, box has %d' % (ranks, torch.cuda.device_count()))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(ranks), '--master-addr', '127.0.0.1',
           '--master-port', '29731', os.path.join(ROOT, 'tests', 'dist_gpu_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(out.stdout[-3000:])
    assert out.returncode == 0, out.stderr[-3000:]
    assert 'DIST CHECK OK' in out.stdout
