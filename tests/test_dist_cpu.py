"""N>1 host logic on CPU: world_size-2 `gloo` run: rendezvous + NCCL-id exchange plumbing,
batch sharding, and `sharded inference + all-gather == full-batch inference` on the oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world_size_2_gloo():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(ROOT, 'tests', '_dist_worker.py')]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='', OMP_NUM_THREADS='1')
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert 'rank 0 OK' in res.stdout and 'rank 1 OK' in res.stdout


def test_shard_bounds():
    from vkjax_b200 import dist as vdist
    assert vdist.shard_bounds(2048, 3, 8) == (768, 1024)
    with pytest.raises(ValueError):
        vdist.shard_bounds(10, 0, 4)
