"""Pins the CPU oracle (oracle/) against every fixed vector available without JAX: Random123
Threefry KATs, JAX's documented PRNG outputs through the whole traced chain, the reference's own known answers,
and C-restated shaders vs numpy-f64 vs torch-CPU for the contractions and pooling."""
import json
import os

import numpy as np
import pytest

from oracle import eval_jaxpr as E, shader_ref
from vkjax_b200.core import ConvDimensionNumbers
from vkjax_b200.frontend import make_jaxpr, random, lax, jnp

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'jax_random.json')))
NHWC = ConvDimensionNumbers((0, 3, 1, 2), (3, 2, 0, 1), (0, 3, 1, 2))


def test_threefry_kat():
    for v in GOLDEN['threefry2x32_kat']:
        o0, o1 = E.threefry2x32(np.uint32(v['key'][0]), np.uint32(v['key'][1]), np.uint32([v['ctr'][0]]), np.uint32([v['ctr'][1]]))
        assert [int(o0[0]), int(o1[0])] == v['out']


def run(f, *args):
    return E.eval_jaxpr(make_jaxpr(f)(*args), *args)[0]


def test_jax_documented_prng_values():
    k0 = random.PRNGKey(0)
    assert np.array_equal(run(lambda k: random.split(k), k0), np.array(GOLDEN['split_key0'], np.uint32))
    assert np.allclose(run(lambda k: random.uniform(k), k0), GOLDEN['uniform_key0'], rtol=1e-7)
    assert np.allclose(run(lambda k: random.normal(k), k0), GOLDEN['normal_key0'], rtol=2e-6)
    assert np.allclose(run(lambda k: random.normal(k, (3,)), random.PRNGKey(42)), GOLDEN['normal_key42_shape3'], rtol=2e-6)


def test_reference_known_answers():
    """f(65)==66, f(-5)==-4 (reference tests/test_function.py:13-16)."""
    f = lambda x: x + 1
    assert run(f, 65) == 66 and run(f, -5) == -4


def test_nextafter_properties():
    x = np.random.RandomState(0).random_sample((77, 101)).astype(np.float32) * 2 - 1
    assert np.all(E.BINARY['nextafter'](x, np.float32(np.inf)) > x)
    assert np.all(E.BINARY['nextafter'](x, np.float32(-np.inf)) < x)


@pytest.mark.parametrize('stride,pad,lhs_dil,rhs_dil', [((1, 1), ((0, 0), (0, 0)), (1, 1), (1, 1)), ((2, 2), ((1, 1), (1, 1)), (1, 1), (1, 1)),
                                                       ((1, 1), ((2, 0), (0, 3)), (1, 1), (1, 1)), ((2, 2), ((0, 0), (0, 0)), (1, 1), (2, 2)),
                                                       ((1, 1), ((2, 2), (3, 3)), (2, 2), (1, 1)), ((2, 2), ((0, 0), (0, 0)), (2, 2), (1, 1))])
def test_conv_three_ways(stride, pad, lhs_dil, rhs_dil):
    rs = np.random.RandomState(1)
    x, w = rs.random_sample((3, 17, 12, 5)).astype(np.float32), rs.random_sample((3, 3, 5, 7)).astype(np.float32)
    y64 = E.conv_general_dilated(x, w, stride, pad, lhs_dil, rhs_dil, NHWC)
    yc = shader_ref.conv2d(x, w, y64.shape, NHWC, (pad[0][0], pad[1][0]), stride, lhs_dil, rhs_dil)
    assert np.allclose(yc, y64, rtol=1e-5, atol=1e-6)
    if lhs_dil == (1, 1):
        yt = E.conv_general_dilated(x, w, stride, pad, lhs_dil, rhs_dil, NHWC, backend='torch')
        assert np.allclose(yt, y64, rtol=1e-5, atol=1e-6)


def test_conv_nchw_spec():
    rs = np.random.RandomState(2)
    dn = ConvDimensionNumbers((0, 1, 2, 3), (0, 1, 2, 3), (0, 1, 2, 3))
    x, w = rs.random_sample((2, 8, 15, 11)).astype(np.float32), rs.random_sample((9, 8, 3, 3)).astype(np.float32)
    y64 = E.conv_general_dilated(x, w, (1, 1), ((0, 0), (0, 0)), (1, 1), (1, 1), dn)
    yc = shader_ref.conv2d(x, w, y64.shape, dn, (0, 0), (1, 1))
    assert y64.shape == (2, 9, 13, 9) and np.allclose(yc, y64, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('ca,cb', [(1, 0), (0, 0), (1, 1), (0, 1)])
def test_dot_two_ways(ca, cb):
    rs = np.random.RandomState(3)
    a = rs.random_sample((2, 100) if ca == 1 else (100, 2)).astype(np.float32)
    b = rs.random_sample((100, 32) if cb == 0 else (32, 100)).astype(np.float32)
    assert np.allclose(shader_ref.dot_general(a, b, ca, cb), E.dot_general(a, b, (((ca,), (cb,)), ((), ()))), rtol=1e-5)


def test_reduce_window_max_and_q3():
    x = np.random.RandomState(4).random_sample((7, 10, 99, 17)).astype(np.float32)
    pads = ((0, 0), (0, 1), (1, 1), (0, 0))
    lax_out = E.reduce_window(x, 'max', (1, 3, 3, 1), (1, 2, 2, 1), pads)
    shader = shader_ref.reduce_window_max(x, lax_out.shape, (0, 0, 1, 0), (1, 2, 2, 1), (1, 3, 3, 1), q3=True)
    assert np.array_equal(lax_out, shader)              # equal on non-negative inputs (all the reference tests use)
    xn = x - 2.0                                        # all negative: the shader's 0.0 padding wins at the borders (quirk Q3)
    lax_n = E.reduce_window(xn, 'max', (1, 3, 3, 1), (1, 2, 2, 1), pads)
    shader_n = shader_ref.reduce_window_max(xn, lax_n.shape, (0, 0, 1, 0), (1, 2, 2, 1), (1, 3, 3, 1), q3=True)
    fixed = shader_ref.reduce_window_max(xn, lax_n.shape, (0, 0, 1, 0), (1, 2, 2, 1), (1, 3, 3, 1), q3=False)
    assert np.array_equal(lax_n, fixed) and not np.array_equal(lax_n, shader_n)


def test_shifts_follow_xla():
    a = np.arange(-777, 777).astype(np.int32)
    assert np.array_equal(E.BINARY['shift_right_logical'](np.int32(1), np.int32(32)), 0)
    assert np.array_equal(E.BINARY['shift_right_arithmetic'](a.view(np.uint32), np.uint32(1)), (a >> 1).view(np.uint32))
    assert np.array_equal(E.BINARY['shift_left'](a, np.int32(1)), a << 1)


def test_resnet_oracle_torch_vs_numpy_conv():
    from vkjax_b200 import nets, tree_util
    model = nets.ResNet18()
    s = model.init(0)
    x = np.random.default_rng(0).random((1, 64, 64, 3), np.float32)
    jaxpr = make_jaxpr(lambda x, s: model.apply(s, x))(x, s)
    leaves = tree_util.tree_leaves((x, s))
    y = E.eval_jaxpr(jaxpr, *leaves)[0]
    os.environ['ORACLE_CONV_BACKEND'] = 'torch'
    try:
        yt = E.eval_jaxpr(jaxpr, *leaves)[0]
    finally:
        del os.environ['ORACLE_CONV_BACKEND']
    assert y.shape == (1, 1000) and np.allclose(y, yt, rtol=1e-4, atol=1e-4)
