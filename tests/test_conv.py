"""Restatement of reference tests/test_conv.py: the same 13 conv_general_dilated cases (:66-85), in every
contraction mode.  Truth = numpy oracle (float64 accumulation) on the same jaxpr and inputs; the 'simt'
mode is additionally compared with the C restatement of conv2d.comp (same summation order)."""
import numpy as np
import pytest

import vkjax_b200 as vkjax
from vkjax_b200.frontend import lax
from vkjax_b200.core import ConvDimensionNumbers
from common import check, as_device_dtype

pytestmark = pytest.mark.gpu
rng = np.random.RandomState(7)
R = rng.random_sample
NHWC = ConvDimensionNumbers((0, 3, 1, 2), (3, 2, 0, 1), (0, 3, 1, 2))
NCHW = ConvDimensionNumbers((0, 1, 2, 3), (0, 1, 2, 3), (0, 1, 2, 3))

kernel_1x1 = R([1, 1, 5, 33])
def conv0(x): return lax.conv_general_dilated(x, kernel_1x1, (1, 1), 'VALID', dimension_numbers=NHWC)
def conv1(x, k): return lax.conv_general_dilated(x, k, (1, 1), 'VALID', dimension_numbers=NHWC)
def conv1a(x, k): return lax.conv_general_dilated(x, k, (1, 1), 'VALID', dimension_numbers=NCHW)
def conv2(x, k): return lax.conv_general_dilated(x, k, (1, 1), 'SAME', dimension_numbers=NHWC)
def conv3(x, k): return lax.conv_general_dilated(x, k, (1, 1), [(2, 0), (0, 3)], dimension_numbers=NHWC)
def conv4(x, k): return lax.conv_general_dilated(x, k, (2, 2), 'SAME', dimension_numbers=NHWC)
def conv5(x, k): return lax.conv_general_dilated(x, k, (2, 2), 'VALID', rhs_dilation=(2, 2), dimension_numbers=NHWC)
def conv6a(x, k): return lax.conv_general_dilated(x, k, (1, 1), [(2, 2), (3, 3)], lhs_dilation=(2, 2), dimension_numbers=NHWC)
def conv6b(x, k): return lax.conv_general_dilated(x, k, (2, 2), [(0, 0), (0, 0)], lhs_dilation=(2, 2), dimension_numbers=NHWC)

param_matrix = [
    (conv0, 'conv0 1x1 const kernel no pad', [R([11, 100, 100, 5])]),
    (conv1, 'conv1 1x1 var kernel no pad', [R([11, 100, 100, 33]), R([1, 1, 33, 11])]),
    (conv1, 'conv1 3x3 var kernel no pad', [R([40, 65, 33, 5]), R([3, 3, 5, 7])]),
    (conv1a, 'conv1a 3x3 convdims 0123', [R([40, 8, 65, 35]), R([39, 8, 3, 3])]),
    (conv2, 'conv2 1x1 var kernel +pad', [R([77, 17, 9, 12]), R([1, 1, 12, 11])]),
    (conv2, 'conv2 3x3 var kernel +pad', [R([23, 44, 19, 7]), R([3, 3, 7, 38])]),
    (conv2, 'conv2 7x7 var kernel +pad', [R([8, 12, 19, 3]), R([7, 7, 3, 4])]),
    (conv3, 'conv3 3x3 uneven pad', [R([15, 67, 42, 11]), R([3, 3, 11, 38])]),
    (conv4, 'conv4 1x1 window strides=2', [R([2, 67, 42, 3]), R([1, 1, 3, 2])]),
    (conv4, 'conv4 3x3 window strides=2', [R([15, 67, 42, 11]), R([3, 3, 11, 7])]),
    (conv5, 'conv5 3x3 window strides=2 + rhs_dilate=2', [R([15, 67, 42, 11]), R([3, 3, 11, 7])]),
    (conv6a, 'conv6a 3x3 window strides=2 + lhs_dilate=2', [R([15, 67, 42, 11]), R([3, 3, 11, 7])]),
    (conv6b, 'conv6b 3x3 window strides=2 + lhs_dilate=2', [R([15, 67, 42, 11]), R([3, 3, 11, 7])]),
    # tensor-core eligible variants of the same functions (O % 4 == 0): exercise the tcgen05 path
    (conv1, 'tc 1x1 C=33->O=12', [R([11, 100, 100, 33]), R([1, 1, 33, 12])]),
    (conv2, 'tc 3x3 SAME C=7->O=40', [R([23, 44, 19, 7]), R([3, 3, 7, 40])]),
    (conv2, 'tc 7x7 SAME C=3->O=64 (stem-like)', [R([4, 56, 56, 3]), R([7, 7, 3, 64])]),
    (conv4, 'tc 3x3 s2 SAME C=64->O=128', [R([6, 30, 29, 64]), R([3, 3, 64, 128])]),
    (conv5, 'tc 3x3 s2 rhs_dil 2 C=16->O=256', [R([3, 33, 42, 16]), R([3, 3, 16, 256])]),
    (conv3, 'tc 3x3 uneven pad C=32->O=36', [R([5, 37, 22, 32]), R([3, 3, 32, 36])]),
    (conv1, 'tc 1x1 C=256->O=64 (bottleneck)', [R([4, 28, 28, 256]), R([1, 1, 256, 64])]),
    # TMA-fed persistent kernel (conv_tc2): tiled-2D A for 1x1/s1, im2col A for k x k with C % 32 == 0
    (conv1, 'tma 1x1 C=64->O=256, 400 tiles (persistent loop)', [R([16, 40, 40, 64]), R([1, 1, 64, 256])]),
    (conv1, 'tma 1x1 C=36->O=72 (K, N tails)', [R([3, 17, 19, 36]), R([1, 1, 36, 72])]),
    (conv2, 'tma 3x3 SAME C=32->O=64', [R([5, 23, 31, 32]), R([3, 3, 32, 64])]),
    (conv2, 'tma 3x3 SAME C=128->O=128, M tail', [R([3, 13, 13, 128]), R([3, 3, 128, 128])]),
    (conv4, 'tma 3x3 s2 SAME C=64->O=192', [R([6, 30, 29, 64]), R([3, 3, 64, 192])]),
    (conv4, 'tma 1x1 s2 SAME C=64->O=128 (projection)', [R([4, 28, 28, 64]), R([1, 1, 64, 128])]),
    (conv5, 'tma 3x3 s2 VALID rhs_dil 2 C=32->O=64', [R([3, 33, 42, 32]), R([3, 3, 32, 64])]),
    (conv3, 'tma 3x3 uneven pad (2,0),(0,3) C=32->O=36', [R([5, 37, 22, 32]), R([3, 3, 32, 36])]),
    (conv1, 'tma 3x3 VALID C=64->O=64', [R([2, 21, 20, 64]), R([3, 3, 64, 64])]),
    (conv2, 'tma 7x7 SAME C=32->O=32', [R([2, 20, 21, 32]), R([7, 7, 32, 32])]),
    # patch kernel (conv_patch.cuh): stride-1 k x k, N <= 128, one activation fetch shared by all filter taps
    (conv2, 'patch 3x3 SAME C=64->O=64 56x56 (ResNet stage 0)', [R([3, 56, 56, 64]), R([3, 3, 64, 64])]),
    (conv2, 'patch 3x3 SAME C=32->O=128 28x28', [R([5, 28, 28, 32]), R([3, 3, 32, 128])]),
    (conv1, 'patch 3x3 VALID C=32->O=64 30x30', [R([2, 30, 30, 32]), R([3, 3, 32, 64])]),
    (conv3, 'patch 3x3 uneven pad (2,0),(0,3) C=32->O=36', [R([2, 26, 27, 32]), R([3, 3, 32, 36])]),
    (conv2, 'patch 5x5 SAME C=32->O=32 27x27, ragged last row block', [R([3, 27, 27, 32]), R([5, 5, 32, 32])]),
    # CTA pairs (tcgen05 cta_group::2): 256 x 256 tiles need >= ~60 of them to be chosen
    (conv1, 'pair 1x1 C=64->O=256, 64 tiles of 256x256', [R([4, 64, 64, 64]), R([1, 1, 64, 256])]),
    (conv1, 'pair 1x1 C=32->O=256, M tail (15477 rows)', [R([3, 77, 67, 32]), R([1, 1, 32, 256])]),
    (conv2, 'pair 3x3 SAME C=32->O=512, 256x256 tiles', [R([4, 64, 64, 32]), R([3, 3, 32, 512])]),
    (conv4, 'pair 3x3 s2 SAME C=32->O=384 (N tail of a 256 tile)', [R([9, 84, 84, 32]), R([3, 3, 32, 384])]),
]
IDS = [f'{i:02d}-{p[1]}' for i, p in enumerate(param_matrix)]


@pytest.mark.parametrize('f,desc,args', param_matrix, ids=IDS)
@pytest.mark.parametrize('precision', ['fp32', 'simt'])
def test_convmatrix(f, desc, args, precision):
    """reference tolerance: default np.allclose, rtol=1e-5, atol=1e-8 (tests/test_conv.py:115)."""
    check(f, args, 1e-5, 1e-8, precision=precision)


@pytest.mark.parametrize('f,desc,args', param_matrix, ids=IDS)
def test_convmatrix_tf32(f, desc, args):
    """single-pass TF32: north-star tolerance rtol=2e-3 for tensor-core contractions."""
    check(f, args, 2e-3, 1e-6, precision='tf32')


@pytest.mark.parametrize('f,desc,args', param_matrix[:13], ids=IDS[:13])
def test_conv_simt_matches_shader_order(f, desc, args):
    """the fp32-FMA kernel keeps conv2d.comp's (kh,kw,c) summation order: compare with oracle/shader_ref.c"""
    from oracle import shader_ref
    from vkjax_b200.frontend import make_jaxpr
    args = [as_device_dtype(a) for a in args]
    jaxpr = make_jaxpr(f)(*args)
    eq = [e for e in jaxpr.jaxpr.eqns if e.primitive.name == 'conv_general_dilated'][0]
    p = eq.params
    lhs = args[0]
    rhs = args[1] if len(args) > 1 else jaxpr.consts[0]
    ref = shader_ref.conv2d(lhs, rhs, eq.outvars[0].aval.shape, p['dimension_numbers'],
                            (p['padding'][0][0], p['padding'][1][0]), p['window_strides'], p['lhs_dilation'], p['rhs_dilation'])
    y = vkjax.Function(f, precision='simt')(*args)
    assert np.array_equal(np.asarray(y), ref) or np.allclose(y, ref, rtol=1e-6, atol=0)


def test_fused_block_tma_kernels():
    """conv -> BatchNorm -> (+ residual) -> ReLU fused into the tcgen05 epilogue, at shapes that take the TMA path;
    checked against the unfused, fp32-FMA execution of the same jaxpr and against the oracle."""
    from vkjax_b200.frontend import nn
    from vkjax_b200 import nets
    rs = np.random.RandomState(5)
    C = 64
    def block(x, w1, w2, w3, bn1, bn2, bn3):
        y = nn.relu(nets.batch_norm(nets.conv2d(x, w1, 1), bn1))
        y = nn.relu(nets.batch_norm(nets.conv2d(y, w2, 1, [(1, 1), (1, 1)]), bn2))
        y = nets.batch_norm(nets.conv2d(y, w3, 1), bn3)
        return nn.relu(x + y)
    def bn(c):
        return {'scale': rs.uniform(0.5, 1.5, (1, 1, 1, c)).astype(np.float32), 'offset': rs.normal(0, 0.1, (1, 1, 1, c)).astype(np.float32),
                'mean': rs.normal(0, 0.1, (1, 1, 1, c)).astype(np.float32), 'var': rs.uniform(0.5, 1.5, (1, 1, 1, c)).astype(np.float32)}
    args = [rs.normal(0, 1, (6, 28, 28, 4 * C)).astype(np.float32), rs.normal(0, (2 / (4 * C)) ** .5, (1, 1, 4 * C, C)).astype(np.float32),
            rs.normal(0, (2 / (9 * C)) ** .5, (3, 3, C, C)).astype(np.float32), rs.normal(0, (2 / C) ** .5, (1, 1, C, 4 * C)).astype(np.float32),
            bn(C), bn(C), bn(4 * C)]
    y_ref = vkjax.Function(block, precision='simt', fuse=False)(*args)
    y_tf32 = vkjax.Function(block, precision='tf32')(*args)
    y_fp32 = vkjax.Function(block, precision='fp32')(*args)
    from common import oracle
    y_true, _ = oracle(block, args)
    assert np.allclose(y_ref, y_true, rtol=1e-4, atol=1e-5)
    assert np.allclose(y_fp32, y_true, rtol=1e-4, atol=1e-5)
    assert np.allclose(y_tf32, y_true, rtol=2e-2, atol=2e-2)      # three chained TF32 convs
