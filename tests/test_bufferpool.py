"""Restatement of reference tests/test_bufferpool.py: the liveness/reuse planner.  Planning is pure host logic,
so these run without a GPU (BufferPool(ctx=None) / JaxprInterpreter(dry_run=True))."""
import numpy as np

from vkjax_b200 import core, JaxprInterpreter
from vkjax_b200.buffers import BufferPool
from vkjax_b200.frontend import make_jaxpr


def test_lowlevel():
    """≙ reference tests/test_bufferpool.py:13-40, verbatim expectations."""
    pool = BufferPool(None, 1)
    v0 = core.Var(0, '', core.ShapedArray((10, 10), 'float32'))
    pool.get_buffer(v0, increment_op_counter=True)
    assert pool.op_counter == 1
    v1 = core.Var(1, '', core.ShapedArray((10, 10), 'float32'))
    pool.get_buffer(v1, increment_op_counter=True)
    assert pool.op_counter == 2
    pool.get_buffer(v0, increment_op_counter=True)
    # this one should re-use a previous tensor
    assert pool.op_counter == 3
    v3 = core.Var(3, '', core.ShapedArray((10, 10), 'float32'))
    pool.get_buffer(v3, increment_op_counter=True)

    assert pool.buffers[v0].accesses == [0, 2]
    assert pool.buffers[v1].accesses == [1]
    assert pool.buffers[v3].accesses == [3]
    pool.create_tensors()
    # only 2 tensors should have been created
    assert len(pool.unique_tensors()) == 2


def func(a, b):
    c = a + 5
    d = c * b
    e = d - 2
    f = e ** 3
    g = f - 1
    h = g * 5
    return h


def test_highlevel():
    """≙ reference :43-63: "11 buffers overall, -3 re-used".  Literals are immediates here (no tensors), so the
    count is a, b, h (own) + 5 intermediates sharing 2 arena slots = 5 tensors; the invariant is the 3 re-uses."""
    jaxpr = make_jaxpr(func)(65, 5)
    it = JaxprInterpreter(jaxpr, dry_run=True, reuse_buffers=True, fuse=False)
    pool = it.bufferpool
    inter = [b for b in pool.buffers.values() if b is not None and not b.tensor.own]
    assert len(inter) == 5
    assert len({id(b.tensor) for b in inter}) == 2          # 5 buffers - 3 re-used
    assert len(pool.unique_tensors()) == 5
    # without reuse every intermediate gets its own slot
    it = JaxprInterpreter(jaxpr, dry_run=True, reuse_buffers=False, fuse=False)
    assert len(it.bufferpool.unique_tensors()) == 8
    # with fusion the whole function is one elementwise chain: nothing intermediate is materialised
    it = JaxprInterpreter(jaxpr, dry_run=True, fuse=True)
    assert len(it.all_ops) == 1 and it.bufferpool.arena_bytes == 0


def test_plan_never_overlaps_live_buffers():
    """Interval-overlap assertion over a real plan (SURVEY §5 'race detection'): two buffers that share a tensor
    must have disjoint access intervals."""
    from vkjax_b200 import nets
    model = nets.ResNet18()
    states = model.init(0)
    x = np.zeros((2, 64, 64, 3), np.float32)
    jaxpr = make_jaxpr(lambda x, s: model.apply(s, x))(x, states)
    for fuse in (True, False):
        it = JaxprInterpreter(jaxpr, dry_run=True, fuse=fuse)
        by_tensor = {}
        for b in it.bufferpool.buffers.values():
            if b is None or b.tensor is None or b.tensor.own or not b.accesses:
                continue
            by_tensor.setdefault(id(b.tensor), {})[id(b.accesses)] = (min(b.accesses), max(b.accesses), b)
        n_shared = 0
        for users in by_tensor.values():
            iv = sorted((lo, hi) for lo, hi, _ in users.values())
            n_shared += len(iv) > 1
            for (lo0, hi0), (lo1, hi1) in zip(iv, iv[1:]):
                assert hi0 < lo1, (iv,)
            assert all(b.nbytes() <= b.tensor.nbytes for _, _, b in users.values())
        assert n_shared > 0


def _hidden_features(x, w):
    from vkjax_b200.frontend import nn, jnp
    h = nn.relu(jnp.dot(x, w))
    u = nn.relu(jnp.dot(h, w))
    v = nn.relu(jnp.dot(u, w))
    return h, jnp.dot(v, w)


def _reshaped_intermediate(x):
    from vkjax_b200.frontend import jnp
    a = (x + 1.0).reshape(-1)
    b = jnp.sum((x * 2.0) * 3.0)
    return a, b


def _output_slots_are_live_to_the_end(it):
    """No buffer whose interval starts after an output's producer may share the output's arena tensor."""
    n_ops = len(it.all_ops)
    outs = [b for b in it.output_buffers if b is not None]
    for ob in outs:
        assert max(ob.accesses) >= n_ops, 'an output must stay live until the download'
        for b in it.bufferpool.buffers.values():
            if b is None or b.tensor is None or b.same_storage(ob) or b.tensor is not ob.tensor or not b.accesses:
                continue
            assert max(b.accesses) < min(ob.accesses), (b, ob, b.accesses, ob.accesses)


def test_outputs_rebound_to_intermediates_stay_live():
    """an output that is the result of an inlined call (nn.relu -> custom_jvp_call_jaxpr) or a
    reshape view of an intermediate lives in the arena; after fusion its liveness was rebuilt from the op list only,
    without the end-of-program read (reference kompute_jaxpr_interpreter.py:45), so a later op could take its slot."""
    x, w = np.zeros((64, 64), np.float32), np.zeros((64, 64), np.float32)
    for fuse in (True, False):
        for reuse in (True, False):
            it = JaxprInterpreter(make_jaxpr(_hidden_features)(x, w), dry_run=True, fuse=fuse, reuse_buffers=reuse)
            _output_slots_are_live_to_the_end(it)
            it = JaxprInterpreter(make_jaxpr(_reshaped_intermediate)(np.zeros((32, 32), np.float32)), dry_run=True,
                                  fuse=fuse, reuse_buffers=reuse)
            _output_slots_are_live_to_the_end(it)
