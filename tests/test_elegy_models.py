"""Restatement of the reference's integration tests (tests/test_elegy_mlp.py:37-58, tests/test_elegy_conv.py:27-47,
tests/test_elegy_resnet.py:18-32) and of BASELINE.json's configs C1-C4 at their stated sizes.

Elegy is not installable here: the models are vkjax_b200.nets (same primitive sequences), wrapped
by vkjax_b200.elegy.vkModel exactly as the reference wraps elegy.Model.  Truth = the numpy oracle (float64 accumulation)
evaluating the same jaxpr on the same inputs and weights.  Tolerances are the reference's: MLP / ConvNet inference
atol 1e-6 (test_elegy_mlp.py:57, test_elegy_conv.py:46), ResNet rtol 1e-4 / atol 1e-5 (test_elegy_resnet.py:32) for the
fp32 path; single-pass TF32 is held to the north star's rtol 2e-3 per contraction, i.e. a relative-L2 bound on logits.
"""
import numpy as np
import pytest

import vkjax_b200 as vkjax
from vkjax_b200 import nets, tree_util
from vkjax_b200.elegy import vkModel
from vkjax_b200.frontend import jnp, lax
from common import oracle

pytestmark = pytest.mark.gpu


def _predict_oracle(module, states, x):
    y, _ = oracle(lambda x, s: module.apply(s, x), [x, states])
    return np.asarray(y)


def test_c1_readme_example():
    """BASELINE configs[0] / reference README.md:7-22: jnp.dot(x, W) + b through vkjax.wrap, float64 host inputs."""
    rng = np.random.default_rng(0)
    x, W, b = rng.random((8, 128)), rng.random((128, 16)), rng.random(16)
    f = lambda x, W, b: jnp.dot(x, W) + b
    y = vkjax.wrap(f)(x, W, b)
    assert y.dtype == np.float32 and y.shape == (8, 16)
    assert np.allclose(y, (x.astype(np.float32) @ W.astype(np.float32)) + b.astype(np.float32), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('batch', [2, 4096], ids=['reference_shape_b2', 'C2_b4096'])
def test_mlp_inference(batch):
    """≙ reference tests/test_elegy_mlp.py:37-58 ([2,32,32,3]); batch 4096 is BASELINE configs[1]."""
    module = nets.MLP()
    model = vkModel(module)
    model.init(seed=3)
    x = np.random.default_rng(1).integers(0, 256, (batch, 32, 32, 3)).astype(np.float32)
    y = model.predict(x)
    states = tree_util.tree_map(np.asarray, model.states)
    ytrue = _predict_oracle(module, states, x)
    assert y.shape == (batch, 10) and y.dtype == np.float32
    assert np.allclose(y, ytrue, atol=1e-6 if batch == 2 else 2e-6, rtol=1e-5)
    # fast path: same function, single-pass TF32 contractions
    y32 = vkModel(module, precision='tf32')
    y32.states, y32.initialized = model.states, True
    yt = y32.predict(x)
    assert np.linalg.norm(yt - ytrue) / np.linalg.norm(ytrue) < 2e-3


def test_convnet_inference():
    """≙ reference tests/test_elegy_conv.py:27-47: 2 x (Conv2D 32, 3x3, stride 2, ReLU) + Linear on [5,32,32,3]."""
    module = nets.ConvNet()
    model = vkModel(module)
    model.init(seed=5)
    x = np.random.default_rng(2).random((5, 32, 32, 3), np.float32)
    y = model.predict(x)
    ytrue = _predict_oracle(module, tree_util.tree_map(np.asarray, model.states), x)
    assert np.allclose(y, ytrue, atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize('arch,batch', [('ResNet18', 1), ('ResNet50', 4)], ids=['resnet18_b1_reference_shape', 'resnet50_b4'])
def test_resnet_inference(arch, batch):
    """≙ reference tests/test_elegy_resnet.py:18-32 (ResNet18, [1,224,224,3], rtol 1e-4 / atol 1e-5), random-init weights
    instead of the downloaded ones; ResNet-50 is the BASELINE model."""
    module = getattr(nets, arch)()
    model = vkModel(module)                                   # precision='fp32': 3xTF32 with chunked promotion
    model.init(seed=0)
    x = np.random.default_rng(4).random((batch, 224, 224, 3), np.float32)
    y = model.predict(x)
    ytrue = _predict_oracle(module, tree_util.tree_map(np.asarray, model.states), x)
    assert y.shape == (batch, 1000)
    # the reference's tolerance, verbatim (tests/test_elegy_resnet.py:32); logits are O(1..10) with the synthetic init
    assert 1.0 < float(np.abs(ytrue).max()) < 100.0
    assert np.allclose(y, ytrue, rtol=1e-4, atol=1e-5), float(np.abs(y - ytrue).max())
    fast = vkModel(module, precision='tf32')
    fast.states, fast.initialized = model.states, True
    yt = fast.predict(x)
    assert np.linalg.norm(yt - ytrue) / np.linalg.norm(ytrue) < 5e-3
    assert (yt.argmax(-1) == ytrue.argmax(-1)).all()


@pytest.mark.parametrize('precision', ['tf32', 'fp32'])
def test_c4_resnet50_batch256_full_size(precision):
    """BASELINE configs[3] at its stated size (ResNet-50, [256,224,224,3] fp32 -- the bench workload), where the oracle is
    too slow to evaluate the whole batch.  (1) Parity on a sample: rows 0, 1, 254, 255 of the batch-256 result against the
    oracle evaluating those four images.  (2) Size-independent properties of a batch-parallel path: every image's logits
    are independent of its position and of its batch mates (reversing the batch reverses the rows, bit for bit; the first
    four rows equal a batch-4 run within rounding -- tile shapes differ between the two problem sizes, the per-element
    summation order over k does not)."""
    module = nets.ResNet50()
    model = vkModel(module, precision=precision)
    model.init(seed=0)
    x = np.random.default_rng(8).random((256, 224, 224, 3), np.float32)
    y = model.predict_on_batch(x)
    assert y.shape == (256, 1000) and np.isfinite(y).all()
    rows = [0, 1, 254, 255]
    ytrue = _predict_oracle(module, tree_util.tree_map(np.asarray, model.states), x[rows])
    if precision == 'fp32':
        assert np.allclose(y[rows], ytrue, rtol=1e-4, atol=1e-5), float(np.abs(y[rows] - ytrue).max())   # reference tolerance, verbatim
    else:
        assert np.linalg.norm(y[rows] - ytrue) / np.linalg.norm(ytrue) < 5e-3
        assert (y[rows].argmax(-1) == ytrue.argmax(-1)).all()
    y_rev = model.predict_on_batch(np.ascontiguousarray(x[::-1]))
    assert np.array_equal(y_rev[::-1], y)
    y4 = model.predict_on_batch(x[:4])
    assert np.allclose(y4, y[:4], rtol=1e-5, atol=1e-5)


def test_predict_batches_pipelined_equals_per_batch():
    """vkModel.predict(x, batch_size) pipelines uploads (Function.map); results equal one call per batch, ragged tail included."""
    module = nets.ResNet18()
    model = vkModel(module, precision='tf32')
    model.init(seed=1)
    x = np.random.default_rng(6).random((22, 64, 64, 3), np.float32)
    y = model.predict(x, batch_size=8)
    ref = np.concatenate([model.predict_on_batch(x[i:i + 8]) for i in range(0, 22, 8)])
    assert y.shape == (22, 1000) and np.array_equal(y, ref)


# ---- C3: the conv / pool shapes of the reference tests with the batch dimension scaled to 256 -----------------------
from vkjax_b200.core import ConvDimensionNumbers
NHWC = ConvDimensionNumbers((0, 3, 1, 2), (3, 2, 0, 1), (0, 3, 1, 2))
NCHW = ConvDimensionNumbers((0, 1, 2, 3), (0, 1, 2, 3), (0, 1, 2, 3))
# all 13 cases of reference tests/test_conv.py:66-85 with the batch dimension -> 256:
# (description, x, kernel, strides, padding, rhs_dilation, lhs_dilation, dimension numbers)
C3_CONVS = [
    ('conv0 1x1 VALID C5->O33', (256, 100, 100, 5), (1, 1, 5, 33), (1, 1), 'VALID', None, None, NHWC),
    ('conv1 1x1 VALID C33->O11', (256, 100, 100, 33), (1, 1, 33, 11), (1, 1), 'VALID', None, None, NHWC),
    ('conv1 3x3 VALID', (256, 65, 33, 5), (3, 3, 5, 7), (1, 1), 'VALID', None, None, NHWC),
    ('conv1a 3x3 NCHW OIHW', (256, 8, 65, 35), (39, 8, 3, 3), (1, 1), 'VALID', None, None, NCHW),
    ('conv2 1x1 SAME', (256, 17, 9, 12), (1, 1, 12, 11), (1, 1), 'SAME', None, None, NHWC),
    ('conv2 3x3 SAME', (256, 44, 19, 7), (3, 3, 7, 38), (1, 1), 'SAME', None, None, NHWC),
    ('conv2 7x7 SAME', (256, 12, 19, 3), (7, 7, 3, 4), (1, 1), 'SAME', None, None, NHWC),
    ('conv3 uneven pad', (256, 67, 42, 11), (3, 3, 11, 38), (1, 1), [(2, 0), (0, 3)], None, None, NHWC),
    ('conv4 1x1 s2 SAME', (256, 67, 42, 3), (1, 1, 3, 2), (2, 2), 'SAME', None, None, NHWC),
    ('conv4 3x3 s2 SAME', (256, 67, 42, 11), (3, 3, 11, 7), (2, 2), 'SAME', None, None, NHWC),
    ('conv5 s2 rhs_dil 2', (256, 67, 42, 11), (3, 3, 11, 7), (2, 2), 'VALID', (2, 2), None, NHWC),
    ('conv6a lhs_dil 2 pad', (256, 67, 42, 11), (3, 3, 11, 7), (1, 1), [(2, 2), (3, 3)], None, (2, 2), NHWC),
    ('conv6b lhs_dil 2 s2', (256, 67, 42, 11), (3, 3, 11, 7), (2, 2), [(0, 0), (0, 0)], None, (2, 2), NHWC),
]


def c3_conv_fn(stride, pad, dil, ldil, dn):
    return lambda x, w: lax.conv_general_dilated(x, w, stride, pad, lhs_dilation=ldil, rhs_dilation=dil, dimension_numbers=dn)


@pytest.mark.parametrize('desc,xs,ws,stride,pad,dil,ldil,dn', C3_CONVS, ids=[c[0] for c in C3_CONVS])
def test_c3_conv_sweep_batch256(desc, xs, ws, stride, pad, dil, ldil, dn, monkeypatch):
    """BASELINE configs[2]: ALL 13 conv cases of reference tests/test_conv.py:66-85 at batch 256, each on the tcgen05
    kernels (asserted).  Truth: oracle with the torch-CPU fp32 conv backend (the float64 numpy path needs minutes at this
    size), hence rtol 2e-5 instead of 1e-5; plus linearity in the filter, a size-independent property:
    conv(x, 2w) == 2 conv(x, w) bit for bit."""
    monkeypatch.setenv('ORACLE_CONV_BACKEND', 'torch')
    rs = np.random.RandomState(len(desc))
    x, w = rs.random_sample(xs).astype(np.float32), rs.random_sample(ws).astype(np.float32)
    f = c3_conv_fn(stride, pad, dil, ldil, dn)
    vk = vkjax.wrap(f)
    y = vk(x, w)
    ytrue, _ = oracle(f, [x, w])
    assert y.shape == ytrue.shape
    assert np.allclose(y, ytrue, rtol=2e-5, atol=1e-6), float(np.abs(y - ytrue).max())
    assert np.array_equal(vk(x, 2.0 * w), 2.0 * y)
    from vkjax_b200.ops import ContractionOp
    interp = list(vk._jaxpr_interpreters.values())[0]
    assert [op.path for op in interp.all_ops if isinstance(op, ContractionOp)] == ['tc']
    assert any('conv_general_dilated' == l for l in interp.labels)


@pytest.mark.parametrize('shape,win,stride,pad', [((256, 100, 111, 5), (1, 2, 2, 1), (1, 1, 1, 1), 'VALID'),
                                                   ((256, 10, 99, 17), (1, 3, 3, 1), (1, 2, 2, 1), 'SAME'),
                                                   ((256, 112, 112, 64), (1, 3, 3, 1), (1, 2, 2, 1), 'SAME')],
                         ids=['2x2_s1_VALID', '3x3_s2_SAME', 'resnet_stem_pool'])
def test_c3_pool_sweep_batch256(shape, win, stride, pad):
    """BASELINE configs[2]: reference tests/test_reduce_window.py:14-21 shapes at batch 256 (+ the ResNet stem pool);
    max is exact, so the result is bit-equal; avg-pool (reduce_window_sum / window) within rtol 1e-6."""
    x = (np.random.RandomState(9).random_sample(shape) - 0.5).astype(np.float32)
    fmax = lambda x: lax.reduce_window(x, -jnp.inf, lax.max, win, stride, pad)
    y, ytrue = vkjax.wrap(fmax)(x), oracle(fmax, [x])[0]
    assert np.array_equal(y, ytrue)
    favg = lambda x: lax.reduce_window(x, 0.0, lax.add, win, stride, pad) / float(win[1] * win[2])
    y, ytrue = vkjax.wrap(favg)(x), oracle(favg, [x])[0]
    assert np.allclose(y, ytrue, rtol=1e-5, atol=1e-6)
