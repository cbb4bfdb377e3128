"""Edge cases of the kernels added in round 2, each against the numpy oracle on the same jaxpr and inputs (through the C ABI):

  * raw-row few-channel convolutions (B2J_CT_ROWS, conv_tc2_kernel<64, A_ROWS / A_ROWS_U8>): channel counts 1..7, filter sizes,
    strides, padding, output rows longer than one 128-pixel tile, K with / without padding to 32, float32 and packed-uint8
    sources, single-pass and 3xTF32 -- and equality with the re-layout + im2col path it replaces (B2J_ENABLE_ROWS=0);
  * the TMA-store epilogue (programs without a residual): ragged M and N tails, bias / bias + ReLU / BN programs -- and
    bit-equality with the st.global epilogue (B2J_TMA_STORE=0);
  * the elementwise kernel's narrow-operand instantiation and its fallbacks, vectorised transposes, sliced column reductions
    (incl. NaN and tie semantics of argmax / argmin).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import vkjax_b200 as vkjax
from vkjax_b200.frontend import lax, jnp, nn
from vkjax_b200.core import ConvDimensionNumbers
from common import check, oracle

pytestmark = pytest.mark.gpu
NHWC = ConvDimensionNumbers((0, 3, 1, 2), (3, 2, 0, 1), (0, 3, 1, 2))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _conv(stride, padding):
    def f(x, w):
        return lax.conv_general_dilated(x, w, stride, padding, dimension_numbers=NHWC)
    return f


def _conv_u8(stride, padding):
    def f(x, w):
        return lax.conv_general_dilated(x.astype(jnp.float32) / 255.0, w, stride, padding, dimension_numbers=NHWC)
    return f


# (name, input NHWC, filter HWIO, stride, padding)
# The raw-row path needs W * C to be a multiple of 16 elements (TMA row pitch for float32 AND packed uint8 sources), O <= 64 and
# Kpad <= 256 (single pass) / 160 (3xTF32: the split weight matrix stays resident in shared memory).
ROWS_CASES = [
    ('stem 7x7/2 C3 O64', (3, 40, 48, 3), (7, 7, 3, 64), (2, 2), 'SAME'),
    ('7x7/2 C3 O64, 304-pixel rows (3 tiles per row, ragged last tile)', (2, 20, 608, 3), (7, 7, 3, 64), (2, 2), 'SAME'),
    ('3x3/1 C3 O32 VALID', (4, 33, 48, 3), (3, 3, 3, 32), (1, 1), 'VALID'),
    ('3x3/1 C4 O64 SAME, 160-pixel rows', (2, 18, 160, 4), (3, 3, 4, 64), (1, 1), 'SAME'),
    ('5x5/2 C1 O16 SAME (K = 25)', (5, 31, 48, 1), (5, 5, 1, 16), (2, 2), 'SAME'),
    ('3x3/2 C5 O40 uneven pad', (3, 29, 32, 5), (3, 3, 5, 40), (2, 2), [(2, 0), (0, 3)]),
    ('7x7/1 C4 O8 SAME (K = 196: single pass only)', (2, 21, 36, 4), (7, 7, 4, 8), (1, 1), 'SAME'),
    ('1x7/1 C2 O64 (K = 14)', (3, 9, 64, 2), (1, 7, 2, 64), (1, 1), 'SAME'),
    ('3x3/3 C3 O16 stride 3', (4, 60, 96, 3), (3, 3, 3, 16), (3, 3), 'VALID'),
]


def _rows_expected(ws, precision):
    kpad = (int(np.prod(ws[:3])) + 31) // 32 * 32
    return kpad <= (160 if precision == 'fp32' else 256)


def _uses_rows(fn):
    interp = list(fn._jaxpr_interpreters.values())[0]
    return [bool(o.attrs.get('rows')) for o in interp.all_ops if hasattr(o, 'attrs') and getattr(o, 'path', '') == 'tc']


@pytest.mark.parametrize('precision', ['fp32', 'tf32'])
@pytest.mark.parametrize('case', ROWS_CASES, ids=[c[0] for c in ROWS_CASES])
def test_rows_conv_float32(case, precision):
    _, xs, ws, stride, padding = case
    rng = np.random.default_rng(11)
    x = rng.random(xs, np.float32)
    w = rng.normal(0, (2.0 / np.prod(ws[:3])) ** 0.5, ws).astype(np.float32)
    f = vkjax.wrap(_conv(stride, padding), precision=precision)
    y = f(x, w)
    ytrue, _ = oracle(_conv(stride, padding), [x, w])
    assert _uses_rows(f) == [_rows_expected(ws, precision)], 'the case is meant to exercise the raw-row kernel'
    tol = dict(rtol=1e-5, atol=1e-6) if precision == 'fp32' else dict(rtol=2e-3, atol=2e-3)
    assert np.allclose(y, ytrue, **tol), float(np.abs(y - ytrue).max())


@pytest.mark.parametrize('precision', ['fp32', 'tf32'])
@pytest.mark.parametrize('case', ROWS_CASES[:5], ids=[c[0] for c in ROWS_CASES[:5]])
def test_rows_conv_uint8_bit_equal_to_float_image(case, precision):
    """uint8 pixels + astype + /255 folded into the gather (lookup table) == the same conv fed the float32 image, bit for bit"""
    _, xs, ws, stride, padding = case
    rng = np.random.default_rng(12)
    xu8 = rng.integers(0, 256, xs, dtype=np.uint8)
    xf = xu8.astype(np.float32) / np.float32(255.0)
    w = rng.normal(0, (2.0 / np.prod(ws[:3])) ** 0.5, ws).astype(np.float32)
    f8 = vkjax.wrap(_conv_u8(stride, padding), precision=precision)
    ff = vkjax.wrap(_conv(stride, padding), precision=precision)
    y8, yf = f8(xu8, w), ff(xf, w)
    interp = list(f8._jaxpr_interpreters.values())[0]
    assert interp.n_input_chains_fused == 1 and _uses_rows(f8) == [True]
    assert np.array_equal(y8, yf)


@pytest.mark.parametrize('precision', ['fp32', 'tf32'])
def test_rows_conv_nan_and_inf_inputs_stay_local(precision):
    """A NaN pixel may only reach the outputs whose window covers it (the K padding columns and the rows beyond a ragged tile are
    assembled from finite data or zeros, never from a neighbour's values).  Inf: exact in the single-pass mode; 3xTF32 turns an
    infinite operand into NaN (lo = x - hi = inf - inf), as every split-precision emulation does -- there the outputs whose window
    covers the Inf must be non-finite and all others exact."""
    rng = np.random.default_rng(13)
    x = rng.random((8, 16, 160, 3), np.float32)
    x[0, 5, 70, 1] = np.nan
    x[1, 9, 20, 2] = np.inf
    w = rng.normal(0, 0.1, (3, 3, 3, 8)).astype(np.float32)
    f = vkjax.wrap(_conv((1, 1), 'SAME'), precision=precision)
    y = f(x, w)
    ytrue, _ = oracle(_conv((1, 1), 'SAME'), [x, w])
    assert _uses_rows(f) == [True]
    assert np.array_equal(np.isfinite(y), np.isfinite(ytrue))
    assert np.array_equal(np.isnan(y[0]), np.isnan(ytrue[0]))
    if precision == 'tf32':
        assert np.array_equal(np.isinf(y), np.isinf(ytrue))
    fin = np.isfinite(ytrue)
    tol = dict(rtol=1e-5, atol=1e-6) if precision == 'fp32' else dict(rtol=2e-3, atol=2e-3)
    assert np.allclose(y[fin], ytrue[fin], **tol)


def _run_py(code, env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd=ROOT, env=e, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout


_AB_CODE = r'''
import sys, hashlib
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import vkjax_b200 as vkjax
from vkjax_b200 import nets
m = nets.ResNet18()
st = m.init(3)
x = np.random.default_rng(4).random((4, 64, 64, 3), np.float32)
for prec in ('fp32', 'tf32'):
    y = vkjax.wrap(lambda x, s: m.apply(s, x), precision=prec)(x, st)
    print(prec, hashlib.sha256(np.ascontiguousarray(y).tobytes()).hexdigest())
'''


def test_tma_store_epilogue_bit_equal_to_st_global():
    """The TMA-store epilogue evaluates the same fp32 operations in the same order: whole-network logits are bit-identical."""
    a = _run_py(_AB_CODE, {'B2J_TMA_STORE': '1'})
    b = _run_py(_AB_CODE, {'B2J_TMA_STORE': '0'})
    assert a == b and 'fp32' in a


def test_rows_kernel_equals_relayout_path():
    """fp32-exact mode: the raw-row stem and the re-layout + im2col stem split and sum the same products in the same k order
    within a k-block chunk, but pad K differently (160 vs 256), so chunk boundaries differ: equality to the reference tolerance."""
    code = _AB_CODE.replace("print(prec, hashlib.sha256(np.ascontiguousarray(y).tobytes()).hexdigest())",
                            "np.save('/tmp/_rows_ab_%s_%s.npy' % (prec, __import__('os').environ['B2J_ENABLE_ROWS']), y)")
    _run_py(code, {'B2J_ENABLE_ROWS': '1'})
    _run_py(code, {'B2J_ENABLE_ROWS': '0'})
    for prec, tol in (('fp32', dict(rtol=1e-4, atol=1e-5)), ('tf32', dict(rtol=5e-2, atol=5e-2))):
        y1, y0 = np.load(f'/tmp/_rows_ab_{prec}_1.npy'), np.load(f'/tmp/_rows_ab_{prec}_0.npy')
        assert np.allclose(y1, y0, **tol), (prec, float(np.abs(y1 - y0).max()))


# ---- TMA-store epilogue: ragged tails and every program it covers ---------------------------------------------------------------
def _dense(act):
    def f(x, w, b):
        y = jnp.dot(x, w) + b
        return nn.relu(y) if act else y
    return f


@pytest.mark.parametrize('precision', ['fp32', 'tf32'])
@pytest.mark.parametrize('act', [False, True])
@pytest.mark.parametrize('m,k,n', [(1000, 256, 520), (129, 64, 36), (4096, 96, 1000), (257, 512, 4)])
def test_tma_store_bias_programs_ragged(m, k, n, act, precision):
    rng = np.random.default_rng(21)
    args = [rng.random((m, k), np.float32), rng.normal(0, k ** -0.5, (k, n)).astype(np.float32), rng.normal(0, 1, (n,)).astype(np.float32)]
    tol = dict(rtol=1e-5, atol=1e-5) if precision == 'fp32' else dict(rtol=5e-3, atol=5e-3)
    check(_dense(act), args, precision=precision, **tol)


@pytest.mark.parametrize('precision', ['fp32', 'tf32'])
def test_tma_store_bn_relu_conv_ragged(precision):
    def f(x, w, mean, var, scale, offset):
        y = lax.conv_general_dilated(x, w, (1, 1), 'SAME', dimension_numbers=NHWC)
        return nn.relu((y - mean) * (scale * lax.rsqrt(var + 1e-5)) + offset)
    rng = np.random.default_rng(22)
    c, o = 32, 72                      # O % 32 != 0: the last 32-column box is clipped by the TMA unit
    args = [rng.random((3, 19, 23, c), np.float32), rng.normal(0, (2 / (9 * c)) ** 0.5, (3, 3, c, o)).astype(np.float32),
            rng.normal(0, 0.1, (1, 1, 1, o)).astype(np.float32), rng.uniform(0.5, 1.5, (1, 1, 1, o)).astype(np.float32),
            rng.uniform(0.5, 1.5, (1, 1, 1, o)).astype(np.float32), rng.normal(0, 0.1, (1, 1, 1, o)).astype(np.float32)]
    tol = dict(rtol=1e-5, atol=1e-5) if precision == 'fp32' else dict(rtol=5e-3, atol=5e-3)
    check(f, args, precision=precision, **tol)


# ---- residual 1x1 layers: cp.async residual prefetch of the single-pass pair kernel (Tc2Cfg RES2) ----------------------------------
def _bottleneck_tail(x, w, mean, var, scale, offset, res):
    y = lax.conv_general_dilated(x, w, (1, 1), 'VALID', dimension_numbers=NHWC)
    return nn.relu((y - mean) * (scale * lax.rsqrt(var + 1e-5)) + offset + res)


def _bottleneck_args(n, h, w_, c, o, seed=23):
    rng = np.random.default_rng(seed)
    ch = lambda lo, hi: rng.uniform(lo, hi, (1, 1, 1, o)).astype(np.float32)
    return [rng.random((n, h, w_, c), np.float32), rng.normal(0, (2 / c) ** 0.5, (1, 1, c, o)).astype(np.float32),
            ch(-0.1, 0.1), ch(0.5, 1.5), ch(0.5, 1.5), ch(-0.1, 0.1), rng.normal(0, 1, (n, h, w_, o)).astype(np.float32)]


@pytest.mark.parametrize('precision', ['tf32', 'fp32'])
@pytest.mark.parametrize('n,h,w_,c,o', [(3, 57, 59, 64, 256),       # M = 10089: ragged against the 256-row pair tile, 2 column tiles
                                        (2, 56, 56, 128, 512),      # 4 column tiles, K = 128: the residual-prefetch instantiation (RES2)
                                        (2, 57, 59, 96, 520),       # RES2 with a ragged last row tile AND a partial last column tile (520 = 4 x 128 + 8)
                                        (1, 101, 97, 256, 384),     # K = 256 (the largest the prefetching instantiation takes), 3 column tiles
                                        (2, 40, 40, 512, 256)])     # K = 512: stays on the register-load epilogue
def test_residual_pointwise_layers_ragged(n, h, w_, c, o, precision):
    """1x1 expand + BatchNorm + residual + ReLU: sizes that put the layer on 256 x 128 CTA-pair tiles with a ragged last row tile"""
    tol = dict(rtol=1e-5, atol=1e-5) if precision == 'fp32' else dict(rtol=5e-3, atol=5e-3)
    check(_bottleneck_tail, _bottleneck_args(n, h, w_, c, o), precision=precision, **tol)


_RES2_CODE = r'''
import sys, hashlib
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import vkjax_b200 as vkjax
from test_round2_kernels import _bottleneck_tail, _bottleneck_args
for shape in ((2, 56, 56, 128, 512), (2, 57, 59, 96, 520), (3, 57, 59, 64, 256)):      # the first two take the prefetching instantiation by default
    y = vkjax.wrap(_bottleneck_tail, precision='tf32')(*_bottleneck_args(*shape))
    print(shape, hashlib.sha256(np.ascontiguousarray(y).tobytes()).hexdigest())
'''


def test_residual_prefetch_epilogue_bit_equal_to_register_loads():
    """The prefetching epilogue evaluates the same fp32 operations in the same order as the register-load one."""
    a = _run_py(_RES2_CODE, {'B2J_TF32_RES2': '1'})
    b = _run_py(_RES2_CODE, {'B2J_TF32_RES2': '0'})
    assert a == b and '512' in a


# ---- elementwise: narrow-operand instantiation and its fallbacks ---------------------------------------------------------------
@pytest.mark.parametrize('shape,c', [((8, 14, 14, 64), 64), ((3, 7, 5, 256), 256), ((2, 9, 11, 96), 96), ((5, 33, 4), 4), ((2, 3, 1030), 1030),
                                     ((7, 1, 13, 1024), 1024), ((1, 1, 1, 68), 68)])
def test_bn_relu_chain_unfused(shape, c):
    """sub, mul, add, max with per-channel operands: periods that divide 1024 take the narrow 8-vector instantiation, the others the
    general one; sizes that are not a multiple of the tile exercise the tails"""
    rng = np.random.default_rng(31)
    bshape = (1,) * (len(shape) - 1) + (c,)

    def f(x, mean, inv, offset):
        return jnp.maximum((x - mean) * inv + offset, 0.0)
    check(f, [rng.normal(0, 1, shape).astype(np.float32)] + [rng.normal(0, 1, bshape).astype(np.float32) for _ in range(3)])


def test_scalar_and_immediate_chain_tail():
    rng = np.random.default_rng(32)
    x = rng.normal(0, 1, (3, 1001)).astype(np.float32)
    s = np.float32(1.7)
    check(lambda x, s: (x * s + 0.25) / 3.0 - s, [x, s])


@pytest.mark.parametrize('shape', [(64, 128), (68, 132), (8, 4, 36), (1000, 4), (4, 1000), (3, 257, 64), (130, 66)])
def test_transpose_vectorised_and_fallback(shape):
    rng = np.random.default_rng(33)
    x = rng.integers(-1000, 1000, shape).astype(np.int32)
    perm = (1, 0) if len(shape) == 2 else (0, 2, 1)
    check(lambda x: jnp.transpose(x, perm), [x])


# ---- sliced column reductions ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('rows,cols', [(4096, 96), (513, 40), (1000, 7), (2048, 1)])
def test_column_max_min_sliced(rows, cols):
    rng = np.random.default_rng(41)
    x = rng.normal(0, 1, (rows, cols)).astype(np.float32)
    x[rows // 3, cols // 2] = np.nan
    check(lambda x: (jnp.max(x, axis=0), jnp.min(x, axis=0)), [x])


@pytest.mark.parametrize('rows,cols', [(4096, 96), (777, 33)])
def test_column_argmax_argmin_ties_and_nan(rows, cols):
    rng = np.random.default_rng(42)
    x = rng.integers(0, 5, (rows, cols)).astype(np.float32)          # many ties: the lowest index has to win across slices
    x[rows - 2, 1] = np.nan                                          # first NaN wins
    x[5, 1] = np.nan
    check(lambda x: (jnp.argmax(x, axis=0), jnp.argmin(x, axis=0)), [x])


def test_row_argmax_ties_and_nan():
    rng = np.random.default_rng(43)
    x = rng.integers(0, 3, (300, 1000)).astype(np.float32)
    x[7, 900] = np.nan
    x[7, 901] = np.nan
    check(lambda x: (jnp.argmax(x, axis=-1), jnp.argmin(x, axis=-1)), [x])
