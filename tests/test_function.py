"""Restatement of reference tests/test_function.py: known answers (:13-16), trace-cache behaviour (:19-46),
dtype preservation (:48-60), plus the static-argument fix for quirk Q1 and DeviceArray inputs."""
import numpy as np
import pytest

import vkjax_b200 as vkjax
from vkjax_b200.frontend import Tracer

pytestmark = pytest.mark.gpu
GLOBAL_VAR = 0


def func0(x):
    global GLOBAL_VAR
    GLOBAL_VAR = x
    return x + 1


def test_function_basic():
    vk_func = vkjax.Function(func0)
    assert vk_func(65) == 66
    assert vk_func(-5) == -4


def test_shape_checking():
    global GLOBAL_VAR
    vk_func = vkjax.Function(func0)
    assert len(vk_func._jaxpr_interpreters) == 0
    vk_func(65)
    assert len(vk_func._jaxpr_interpreters) == 1
    assert isinstance(GLOBAL_VAR, Tracer)
    GLOBAL_VAR = 0
    vk_func(77)
    assert len(vk_func._jaxpr_interpreters) == 1
    assert GLOBAL_VAR == 0
    vk_func(np.zeros([4, 4]))
    assert len(vk_func._jaxpr_interpreters) == 2
    assert isinstance(GLOBAL_VAR, Tracer)
    GLOBAL_VAR = 0
    vk_func(np.zeros([4, 4]))
    assert len(vk_func._jaxpr_interpreters) == 2
    assert GLOBAL_VAR == 0
    vk_func(np.zeros([4, 5]))
    assert len(vk_func._jaxpr_interpreters) == 3
    assert isinstance(GLOBAL_VAR, Tracer)
    GLOBAL_VAR = 0


def test_dtype_checking():
    vk_func = vkjax.Function(func0)
    x = np.arange(128).astype('uint32')
    y0 = vk_func(x)
    y1 = vk_func(x.astype('float32'))
    assert np.allclose(y0, y1)
    assert y0.dtype == np.uint32
    assert y1.dtype == np.float32
    assert len(vk_func._jaxpr_interpreters) == 2


def test_static_args_are_part_of_the_key():
    """quirk Q1: the reference keys its cache on shapes/dtypes only; static values must select the trace."""
    def f(x, training):
        return x * 2.0 if training else x + 1.0
    vk = vkjax.wrap(f, static_argnums=[1])
    x = np.full(8, 3.0, np.float32)
    assert np.array_equal(vk(x, True), x * 2)
    assert np.array_equal(vk(x, False), x + 1)
    assert len(vk._jaxpr_interpreters) == 2


def test_pytree_io_and_passthrough():
    def f(x, state):
        return {'y': x * state['w'] + state['b']}, state
    vk = vkjax.wrap(f)
    x = np.arange(12, dtype=np.float32).reshape(3, 4)
    state = {'w': np.full((3, 4), 2.0, np.float32), 'b': np.ones((3, 4), np.float32)}
    out, state2 = vk(x, state)
    assert np.array_equal(out['y'], x * 2 + 1)
    assert sorted(state2) == ['b', 'w'] and np.array_equal(state2['w'], state['w'])


def test_device_array_inputs_stay_resident():
    w = np.random.RandomState(0).random_sample((64, 64)).astype(np.float32)
    dw = vkjax.device_put(w)
    vk = vkjax.wrap(lambda x, w: x * w)
    x = np.ones((64, 64), np.float32)
    y = vk(x, dw)
    assert np.array_equal(y, w)
    interp = list(vk._jaxpr_interpreters.values())[0]
    y = vk(x * 3, dw)
    assert np.array_equal(y, w * 3)
    assert interp.h2d_bytes == x.nbytes          # only x travelled on the second call


def test_profiling_info():
    """≙ reference function.py:21-24 + kompute_jaxpr_interpreter.py:91-95: (output, [(label, dt)])."""
    vk = vkjax.wrap(lambda x, y: (x + y) * 2.0, profiling=True, fuse=False)
    x = np.ones(1024, np.float32)
    out, info = vk(x, x)
    assert np.array_equal(out, x * 4)
    assert [l for l, _ in info] == ['add', 'mul'] and all(dt >= 0 for _, dt in info)


def test_hoisted_weight_work_follows_rebinding():
    """Weight-only work is hoisted into a prologue for DeviceArray inputs; rebinding a weight must replay it."""
    from vkjax_b200.frontend import lax
    from vkjax_b200.core import ConvDimensionNumbers
    from common import oracle
    dn = ConvDimensionNumbers((0, 3, 1, 2), (3, 2, 0, 1), (0, 3, 1, 2))

    def f(x, w, var, scale):
        y = lax.conv_general_dilated(x, w, (1, 1), 'SAME', dimension_numbers=dn)
        return y * (scale * lax.rsqrt(var + 1e-5))

    rs = np.random.RandomState(3)
    x = rs.normal(0, 1, (2, 12, 12, 32)).astype(np.float32)
    ws = [rs.normal(0, 0.1, (3, 3, 32, 64)).astype(np.float32) for _ in range(2)]
    vs = [rs.uniform(0.5, 1.5, (1, 1, 1, 64)).astype(np.float32) for _ in range(2)]
    scale = np.ones((1, 1, 1, 64), np.float32)
    vk = vkjax.wrap(f, precision='fp32')
    dws, dvs, dscale = [vkjax.device_put(w) for w in ws], [vkjax.device_put(v) for v in vs], vkjax.device_put(scale)
    for wi, vi in [(0, 0), (0, 0), (1, 0), (1, 1), (0, 1)]:
        y = vk(x, dws[wi], dvs[vi], dscale)
        ytrue, _ = oracle(f, [x, ws[wi], vs[vi], scale])
        assert np.allclose(y, ytrue, rtol=1e-5, atol=1e-5), (wi, vi)
    interp = list(vk._jaxpr_interpreters.values())[0]
    assert len(vk._jaxpr_interpreters) == 1 and interp.prologue is not None and interp.n_hoisted == 2
    # host-array weights: same function, separate trace, nothing hoisted
    y = vk(x, ws[1], vs[1], scale)
    assert np.allclose(y, oracle(f, [x, ws[1], vs[1], scale])[0], rtol=1e-5, atol=1e-5)
    assert len(vk._jaxpr_interpreters) == 2


def test_map_pipelines_batches_like_sequential_calls():
    """Function.map (upload of batch i+1 overlaps the replay of batch i) returns what one call per batch returns,
    for pinned and pageable host arrays and with resident weights."""
    from vkjax_b200 import runtime as rt
    rs = np.random.RandomState(5)
    w = vkjax.device_put(rs.normal(0, 1, (64, 32)).astype(np.float32))
    vk = vkjax.wrap(lambda x, w, k: {'y': (x @ w) * 2.0 + 1.0, 'k': k})
    ctx = rt.Context.get()
    big = ctx.pinned_empty((5 * 16, 64), np.float32)
    big[...] = rs.normal(0, 1, big.shape)
    ks = [np.int32(i) for i in range(5)]
    pinned = [(big[i * 16:(i + 1) * 16], w, ks[i]) for i in range(5)]
    pageable = [(np.array(a[0]), w, a[2]) for a in pinned]
    seq = [vk(*a) for a in pageable]
    for batches in (pinned, pageable):
        for lanes in (1, 2, 3):
            outs = vk.map(batches, lanes=lanes)
            assert len(outs) == 5
            for o, s_, a in zip(outs, seq, batches):
                assert np.array_equal(o['y'], s_['y']) and o['k'] == a[2]
    assert vk.map([]) == []
    with pytest.raises(TypeError):
        vk.map([pinned[0], (np.zeros((3, 64), np.float32), w, ks[0])])
