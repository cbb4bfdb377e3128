"""Worker of tests/test_dist_cpu.py: world_size-2 gloo run of the batch-sharding host logic (no GPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vkjax_b200 import dist as vdist, nets, tree_util            # noqa: E402
from vkjax_b200.frontend import make_jaxpr                        # noqa: E402
from oracle.eval_jaxpr import eval_jaxpr                          # noqa: E402


class FakeCtx:
    """Stands in for runtime.Context: records what the rendezvous hands to b2j_comm_init."""
    def __init__(self):
        self.got = None

    def nccl_unique_id(self):
        return bytes(range(128))

    def comm_init(self, world, rank, uid):
        self.got = (world, rank, uid)


def main():
    ctx = FakeCtx()
    rank, world = vdist.init(ctx, 'gloo')
    assert world == 2 and ctx.got == (2, rank, bytes(range(128))), ctx.got

    # batch-sharded inference == full-batch inference (the jaxpr has no cross-sample op)
    model = nets.ConvNet()
    states = model.init(0, in_shape=(16, 16, 3))
    x = np.random.default_rng(0).random((8, 16, 16, 3), np.float32)
    f = lambda x, s: model.apply(s, x)
    xs = vdist.shard_batch(x, rank, world)
    y_shard = eval_jaxpr(make_jaxpr(f)(xs, states), *tree_util.tree_leaves((xs, states)))[0]
    y_all = vdist.all_gather_host(y_shard)
    y_full = eval_jaxpr(make_jaxpr(f)(x, states), *tree_util.tree_leaves((x, states)))[0]
    assert y_all.shape == y_full.shape == (8, 10)
    assert np.allclose(y_all, y_full, rtol=1e-6, atol=1e-6)

    # the graph each rank records ends in an all-gather when asked to (dry run: analysis + planning only)
    from vkjax_b200 import JaxprInterpreter
    it = JaxprInterpreter(make_jaxpr(f)(xs, states), dry_run=True, allgather_outputs=True)
    assert it.allgather_outputs and len(it.all_ops) > 0
    print(f'rank {rank} OK')


if __name__ == '__main__':
    main()
