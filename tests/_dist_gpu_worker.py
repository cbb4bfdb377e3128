"""Worker of tests/test_dist_gpu.py: N ranks (torchrun, one process per GPU) run batch-sharded ResNet inference through
vkModel with the NCCL all-gather of the logits; rank 0 checks the gathered result against single-GPU inference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vkjax_b200 import dist as vdist, nets, runtime as rt            # noqa: E402
from vkjax_b200.elegy import vkModel                                 # noqa: E402


def main():
    rank, local_rank, world = vdist.env()
    ctx = rt.Context.get(local_rank)
    vdist.init(ctx, 'nccl')
    import torch.distributed as dist
    per = 4
    x_all = np.random.default_rng(0).random((per * world, 64, 64, 3), np.float32)       # the same global batch on every rank
    xs = vdist.shard_batch(x_all, rank, world)
    module = nets.ResNet18()
    for mode in (True, 'root'):
        for precision in ('fp32', 'tf32'):
            model = vkModel(module, precision=precision, allgather_outputs=mode)
            model.init(seed=0, host=True)                                              # same seed: replicated weights
            y = model.predict_on_batch(xs)
            local = vkModel(module, precision=precision)
            local.states, local.initialized = model.states, True
            y_local = local.predict_on_batch(xs)                                         # this rank's shard, no collective
            y_full = local.predict_on_batch(x_all)                                       # single-GPU inference of the whole batch
            if mode is True or rank == 0:
                assert y.shape == (per * world, 1000), y.shape
                # every rank's rows arrive unchanged: the gathered block of rank r is rank r's own result, bit for bit
                assert np.array_equal(y[rank * per:(rank + 1) * per], y_local)
                # gathered logits == single-GPU logits, row for row (the per-sample computation does not depend on the batch
                # it sits in; tile shapes may differ between batch sizes, the summation order over k does not)
                assert np.allclose(y, y_full, rtol=1e-5, atol=1e-5), float(np.abs(y - y_full).max())
                assert (y.argmax(-1) == y_full.argmax(-1)).all()
            else:
                assert y.shape == (per, 1000) and np.array_equal(y, y_local)               # 'root': other ranks keep their shard
            # pipelined path (Function.map) gives the same
            outs = model.call_pred_step_jit.map([(xs, model.states, False, False)] * 3)
            assert all(np.array_equal(o[0], y) for o in outs)
            dist.barrier()
    print(f'rank {rank} OK', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
