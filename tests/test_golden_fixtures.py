"""Integer-exact golden fixtures (tests/golden/contractions_int.npz, generator tests/golden/make_golden.py): the 13 conv
cases of reference tests/test_conv.py:66-85, both max-pools of tests/test_reduce_window.py:14-21 and the 6 dot_general
variants of tests/test_basic_ops.py:197-203, computed by plain-Python restatements of conv2d.comp / reduce_window_max_2d.comp /
dot_general.comp with hand-written padding tuples -- no tracer, no oracle, no numpy arithmetic.

The inputs are small integers, so every product and partial sum is exact in fp32 and in TF32: the numpy oracle, the C
restatement AND the CUDA path in all three contraction modes (3xTF32, single-pass TF32, fp32 FMA) must reproduce the
fixtures BIT FOR BIT.  The functions below are the reference test's own (`conv2` passes the string 'SAME', etc.), so the
tracer's padding / dimension-spec arithmetic is pinned too.
"""
import json
import os

import numpy as np
import pytest

from vkjax_b200 import tree_util
from vkjax_b200.core import ConvDimensionNumbers
from vkjax_b200.frontend import make_jaxpr, lax, jnp
from oracle.eval_jaxpr import eval_jaxpr

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'contractions_int.npz'))
META = json.loads(bytes(G['meta']).decode())

NHWC = ConvDimensionNumbers((0, 3, 1, 2), (3, 2, 0, 1), (0, 3, 1, 2))
NCHW = ConvDimensionNumbers((0, 1, 2, 3), (0, 1, 2, 3), (0, 1, 2, 3))


# the functions of reference tests/test_conv.py:13-63, tests/test_reduce_window.py:14-15
def conv1(x, k): return lax.conv_general_dilated(x, k, (1, 1), 'VALID', dimension_numbers=NHWC)
def conv1a(x, k): return lax.conv_general_dilated(x, k, (1, 1), 'VALID', dimension_numbers=NCHW)
def conv2(x, k): return lax.conv_general_dilated(x, k, (1, 1), 'SAME', dimension_numbers=NHWC)
def conv3(x, k): return lax.conv_general_dilated(x, k, (1, 1), [(2, 0), (0, 3)], dimension_numbers=NHWC)
def conv4(x, k): return lax.conv_general_dilated(x, k, (2, 2), 'SAME', dimension_numbers=NHWC)
def conv5(x, k): return lax.conv_general_dilated(x, k, (2, 2), 'VALID', rhs_dilation=(2, 2), dimension_numbers=NHWC)
def conv6a(x, k): return lax.conv_general_dilated(x, k, (1, 1), [(2, 2), (3, 3)], lhs_dilation=(2, 2), dimension_numbers=NHWC)
def conv6b(x, k): return lax.conv_general_dilated(x, k, (2, 2), [(0, 0), (0, 0)], lhs_dilation=(2, 2), dimension_numbers=NHWC)
def reduce_window_max0(x): return lax.reduce_window(x, -jnp.inf, lax.max, (1, 2, 2, 1), (1, 1, 1, 1), 'VALID')
def reduce_window_max1(x): return lax.reduce_window(x, -jnp.inf, lax.max, (1, 3, 3, 1), (1, 2, 2, 1), 'SAME')


CONV_FNS = dict(conv0=conv1, conv1=conv1, conv1a=conv1a, conv2=conv2, conv3=conv3, conv4=conv4, conv5=conv5, conv6a=conv6a, conv6b=conv6b)
POOL_FNS = dict(reduce_window_max0=reduce_window_max0, reduce_window_max1=reduce_window_max1)


def dot_fn(ca, cb):
    return lambda a, b: lax.dot_general(a, b, (((ca,), (cb,)), ((), ())))


def cases():
    out = []
    for i, m in enumerate(META['conv']):
        x, k, y = G[f'conv{i}_x'].astype(np.float32), G[f'conv{i}_k'].astype(np.float32), G[f'conv{i}_y'].astype(np.float32)
        f = CONV_FNS[m['fn']]
        if m['fn'] == 'conv0':                      # closed-over constant kernel (reference tests/test_conv.py:13-16)
            f = (lambda kk: (lambda x: conv1(x, kk)))(k)
            out.append((f'conv{i:02d}-{m["name"]}', f, [x], y))
        else:
            out.append((f'conv{i:02d}-{m["name"]}', f, [x, k], y))
    for i, m in enumerate(META['pool']):
        out.append((f'pool{i}-{m["name"]}', POOL_FNS[m['fn']], [G[f'pool{i}_x'].astype(np.float32)], G[f'pool{i}_y'].astype(np.float32)))
    out.append(('pool-negative-inputs-lax-semantics', reduce_window_max1, [G['pool_neg_x'].astype(np.float32)], G['pool_neg_y'].astype(np.float32)))
    for i, m in enumerate(META['dot']):
        a, b, y = G[f'dot{i}_a'].astype(np.float32), G[f'dot{i}_b'].astype(np.float32), G[f'dot{i}_y'].astype(np.float32)
        if 'const' in m['name']:
            f = (lambda bb: (lambda a: lax.dot_general(a, bb, (((1,), (0,)), ((), ())))))(b)
            out.append((f'dot{i}-{m["name"]}', f, [a], y))
        elif 'reshape' in m['name']:                # x.reshape(-1, 4) @ y with x [2,77,102] (reference tests/test_basic_ops.py:203)
            out.append((f'dot{i}-{m["name"]}', lambda x, y: jnp.dot(x.reshape(-1, 4), y), [a.reshape(2, 77, 102), b], y))
        else:
            out.append((f'dot{i}-{m["name"]}', dot_fn(m['cdim_a'], m['cdim_b']), [a, b], y))
    return out


CASES = cases()
IDS = [c[0] for c in CASES]


@pytest.mark.parametrize('name,f,args,ytrue', CASES, ids=IDS)
def test_oracle_matches_golden(name, f, args, ytrue):
    """pins oracle/eval_jaxpr.py AND the tracer (SAME padding, dimension specs) on reference-derived known answers"""
    jaxpr = make_jaxpr(f)(*args)
    y = eval_jaxpr(jaxpr, *tree_util.tree_leaves(args))[0]
    assert y.shape == ytrue.shape and y.dtype == np.float32
    assert np.array_equal(y, ytrue)


@pytest.mark.parametrize('i', range(len(META['conv'])), ids=[m['name'] for m in META['conv']])
def test_c_restatement_matches_golden(i):
    """pins oracle/shader_ref.c (the C restatement of conv2d.comp) on the same fixtures"""
    from oracle import shader_ref
    m = META['conv'][i]
    x, k, y = G[f'conv{i}_x'].astype(np.float32), G[f'conv{i}_k'].astype(np.float32), G[f'conv{i}_y'].astype(np.float32)
    dn = ConvDimensionNumbers(tuple(m['lhs_spec']), tuple(m['rhs_spec']), tuple(m['out_spec']))
    got = shader_ref.conv2d(x, k, y.shape, dn, tuple(m['pad_lo']), tuple(m['strides']), tuple(m['lhs_dilation']), tuple(m['rhs_dilation']))
    assert np.array_equal(got, y)


def test_traced_padding_equals_hand_written():
    """the low padding the tracer derives from 'SAME' / 'VALID' / explicit pairs == the hand-written tuples of the fixtures"""
    for i, m in enumerate(META['conv']):
        x, k = G[f'conv{i}_x'].astype(np.float32), G[f'conv{i}_k'].astype(np.float32)
        jaxpr = make_jaxpr(CONV_FNS[m['fn']])(x, k)
        eq = [e for e in jaxpr.jaxpr.eqns if e.primitive.name == 'conv_general_dilated'][0]
        assert tuple(p[0] for p in eq.params['padding']) == tuple(m['pad_lo']), m['name']
        assert tuple(eq.outvars[0].aval.shape) == G[f'conv{i}_y'].shape, m['name']
    for i, m in enumerate(META['pool']):
        jaxpr = make_jaxpr(POOL_FNS[m['fn']])(G[f'pool{i}_x'].astype(np.float32))
        eq = [e for e in jaxpr.jaxpr.eqns if e.primitive.name == 'reduce_window_max'][0]
        assert tuple(p[0] for p in eq.params['padding']) == tuple(m['pad_lo']), m['name']


# ---- the CUDA path -----------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('name,f,args,ytrue', CASES, ids=IDS)
@pytest.mark.parametrize('precision', ['fp32', 'tf32', 'simt'])
@pytest.mark.parametrize('force_tc', [False, True], ids=['default_path', 'tensor_cores_forced'])
def test_gpu_matches_golden(name, f, args, ytrue, precision, force_tc, monkeypatch):
    """bit-exact in every contraction mode; `tensor_cores_forced` drops the minimum-FLOPs gate so that these small
    problems run on the tcgen05 kernels (conv_tc2 / conv_patch) wherever the planner can put them there"""
    import vkjax_b200 as vkjax
    from vkjax_b200 import ops
    if force_tc:
        if precision == 'simt' or name.startswith('pool'):
            pytest.skip('no tensor-core path to force')
        monkeypatch.setattr(ops, 'TC_MIN_FLOPS', 0)
    fn = vkjax.Function(f, precision=precision)
    y = fn(*args)
    assert y.shape == ytrue.shape and y.dtype == np.float32
    assert np.array_equal(y, ytrue), f'{int(np.sum(y != ytrue))} of {y.size} elements differ, max abs err {np.abs(y - ytrue).max()}'
    if force_tc and not name.startswith('pool'):
        interp = list(fn._jaxpr_interpreters.values())[0]
        paths = [op.path for op in interp.all_ops if isinstance(op, ops.ContractionOp)]
        assert paths == ['tc'], paths
