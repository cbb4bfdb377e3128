"""Multi-GPU result parity ("no GPU test of the multi-GPU result"): N = 2 ranks under torchrun, one process per
GPU over NCCL; the all-gathered logits must equal single-GPU logits row for row.  Needs >= 2 GPUs: skipped on a
one-GPU box (run it with `gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        out = subprocess.run(['nvidia-smi', '-L'], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith('GPU '))
    except Exception:                                          # noqa: BLE001
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize('world', [2])
def test_gathered_logits_equal_single_gpu(world):
    if _n_gpus() < world:
        pytest.skip(f'needs {world} GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', '29541', os.path.join(ROOT, 'tests', '_dist_gpu_worker.py')]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-5000:]
    for r in range(world):
        assert f'rank {r} OK' in res.stdout
