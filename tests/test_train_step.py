"""Restatement of the reference's training-step tests (tests/test_elegy_mlp.py:62-129: SCCE loss, its value_and_grad, one
SGD step of the LeNet-300-100 MLP with every intermediate compared; tests/test_elegy_conv.py:51-94).

JAX is not installable here, so there is no autodiff: the backward pass is written out by hand with the primitives
`jax.value_and_grad` emits for these models (dot_general with contracting dims (0,0) / (1,1), transpose-free weight
gradients, reduce_max / exp / log / reduce_sum for the softmax cross-entropy, gather for take_along_axis and its
scatter-add transpose, select for the ReLU mask, the SGD update chain).  The differential test is the reference's:
the same jaxpr on the same inputs through vkjax.wrap and through the numpy oracle, outputs at atol 1e-6/rtol 1e-5
(reference: atol 1e-7 against XLA:CPU fp32; the oracle accumulates in float64), every intermediate at atol 1e-5
with reuse_buffers=False / return_all=True (reference :120-129).
"""
import numpy as np
import pytest

import vkjax_b200 as vkjax
from vkjax_b200 import JaxprInterpreter, tree_util, nets
from vkjax_b200.core import GatherDimensionNumbers, ScatterDimensionNumbers
from vkjax_b200.frontend import make_jaxpr, lax, jnp, nn
from oracle.eval_jaxpr import eval_jaxpr
from common import oracle, assert_tree_close

pytestmark = pytest.mark.gpu
LR = 0.1


def scce_mean(logits, labels):
    """elegy.losses.sparse_categorical_crossentropy(labels, logits, from_logits=True).mean()"""
    logp = nn.log_softmax(logits, axis=-1)
    picked = jnp.take_along_axis(logp, labels.reshape(-1, 1), axis=-1)
    return -jnp.mean(picked)


def scce_value_and_grad(logits, labels):
    """what jax.value_and_grad(scce_mean) computes: d loss / d logits = (softmax - onehot) / B"""
    B = logits.shape[0]
    logp = nn.log_softmax(logits, axis=-1)
    picked = jnp.take_along_axis(logp, labels.reshape(-1, 1), axis=-1)
    loss = -jnp.mean(picked)
    rows = lax.broadcast_in_dim(lax.iota(np.int32, B), (B, 1, 1), (0,))
    idx = lax.concatenate([rows, lax.reshape(labels.astype(jnp.int32), (B, 1, 1))], 2)
    dn = ScatterDimensionNumbers(update_window_dims=(), inserted_window_dims=(0, 1), scatter_dims_to_operand_dims=(0, 1))
    ct = jnp.broadcast_to(jnp.asarray(np.float32(-1.0 / B)), (B, 1))
    onehot_term = lax.scatter_add(jnp.broadcast_to(jnp.asarray(np.float32(0.0)), logits.shape), idx, ct, dn)   # -onehot / B
    grad = jnp.exp(logp) * np.float32(1.0 / B) + onehot_term
    return loss, grad


def mlp_train_step(x, y, params):
    """forward + SCCE + backward + SGD(0.1) of nets.MLP (reference tests/test_elegy_mlp.py:87-118)."""
    (W1, b1), (W2, b2), (W3, b3) = [(p['w'], p['b']) for p in params]
    B = x.shape[0]
    h0 = (x.astype(jnp.float32) / 255.0).reshape(B, -1)
    z1 = jnp.dot(h0, W1) + jnp.broadcast_to(b1, (B, W1.shape[1]))
    a1 = nn.relu(z1)
    z2 = jnp.dot(a1, W2) + jnp.broadcast_to(b2, (B, W2.shape[1]))
    a2 = nn.relu(z2)
    z3 = jnp.dot(a2, W3) + jnp.broadcast_to(b3, (B, W3.shape[1]))
    loss, dz3 = scce_value_and_grad(z3, y)
    dot_t = lambda a, b, ca, cb: lax.dot_general(a, b, (((ca,), (cb,)), ((), ())))
    dW3, db3 = dot_t(a2, dz3, 0, 0), jnp.sum(dz3, axis=0)
    dz2 = lax.select(z2 > 0.0, dot_t(dz3, W3, 1, 1), jnp.broadcast_to(jnp.asarray(np.float32(0.0)), z2.shape))
    dW2, db2 = dot_t(a1, dz2, 0, 0), jnp.sum(dz2, axis=0)
    dz1 = lax.select(z1 > 0.0, dot_t(dz2, W2, 1, 1), jnp.broadcast_to(jnp.asarray(np.float32(0.0)), z1.shape))
    dW1, db1 = dot_t(h0, dz1, 0, 0), jnp.sum(dz1, axis=0)
    sgd = lambda w, g: w - LR * g
    new = [{'w': sgd(W1, dW1), 'b': sgd(b1, db1)}, {'w': sgd(W2, dW2), 'b': sgd(b2, db2)}, {'w': sgd(W3, dW3), 'b': sgd(b3, db3)}]
    return {'loss': loss, 'logits': z3}, new


def test_scce():
    """≙ reference tests/test_elegy_mlp.py:62-71"""
    rs = np.random.RandomState(0)
    X = [rs.random_sample((8, 10)).astype(np.float32), rs.randint(0, 10, size=8).astype(np.int32)]
    y, ytrue = vkjax.Function(scce_mean)(*X), oracle(scce_mean, X)[0]
    assert np.allclose(y, ytrue)


def test_scce_value_and_grad():
    """≙ reference tests/test_elegy_mlp.py:73-83"""
    rs = np.random.RandomState(1)
    X = [rs.random_sample((8, 10)).astype(np.float32), rs.randint(0, 10, size=8).astype(np.int32)]
    y, ytrue = vkjax.Function(scce_value_and_grad)(*X), oracle(scce_value_and_grad, X)[0]
    assert_tree_close(y, ytrue, rtol=1e-5, atol=1e-7)
    assert abs(float(np.sum(y[1]))) < 1e-6                    # rows of (softmax - onehot) sum to zero


@pytest.mark.parametrize('precision,batch', [('fp32', 8), ('simt', 8), ('fp32', 256)], ids=['fp32_b8', 'simt_b8', 'fp32_b256'])
def test_mlp_training_step(precision, batch):
    """≙ reference tests/test_elegy_mlp.py:87-118: outputs and updated states of one SGD step."""
    rs = np.random.RandomState(2)
    x = (rs.random_sample((batch, 32, 32, 3)) * 255).astype(np.float32)
    y = rs.randint(0, 10, size=batch).astype(np.int32)
    params = nets.MLP().init(3)
    out, new = vkjax.wrap(mlp_train_step, precision=precision)(x, y, params)
    (out_t, new_t), _ = oracle(mlp_train_step, [x, y, params])
    assert_tree_close(out, out_t, rtol=1e-5, atol=1e-6)
    assert_tree_close(new, new_t, rtol=1e-5, atol=1e-6)
    # the step actually moved the weights, by LR * gradient
    assert float(np.abs(new[2]['w'] - params[2]['w']).max()) > 1e-4


def test_mlp_training_step_every_intermediate():
    """≙ reference tests/test_elegy_mlp.py:120-129: no buffer reuse, every jaxpr variable compared with the oracle."""
    rs = np.random.RandomState(4)
    x = (rs.random_sample((8, 32, 32, 3)) * 255).astype(np.float32)
    y = rs.randint(0, 10, size=8).astype(np.int32)
    params = nets.MLP().init(5)
    jaxpr = make_jaxpr(mlp_train_step)(x, y, params)
    leaves = tree_util.tree_leaves((x, y, params))
    _, envtrue = eval_jaxpr(jaxpr, *leaves, return_env=True)
    interp = JaxprInterpreter(jaxpr, reuse_buffers=False, fuse=False, precision='simt')
    _, envpred = interp.run(*leaves, return_all=True)
    checked = 0
    for var, vtrue in envtrue.items():
        from vkjax_b200 import core
        got = envpred.get(core.hashable(var))
        if got is None:
            continue
        vtrue = np.asarray(vtrue)
        assert got.shape == vtrue.shape, (var, got.shape, vtrue.shape)
        if vtrue.dtype.kind == 'f':
            assert np.allclose(got, vtrue, rtol=1e-5, atol=1e-5, equal_nan=True), (str(var), float(np.abs(got - vtrue).max()))
        else:
            assert np.array_equal(got, vtrue), str(var)
        checked += 1
    assert checked > 40


# ---- ConvNet: reference tests/test_elegy_conv.py:51-94 (2 x (Conv2D 32, 3x3, stride 2, SAME) + ReLU, Linear, SCCE, SGD) -----
from vkjax_b200.core import ConvDimensionNumbers
FWD = ConvDimensionNumbers((0, 3, 1, 2), (3, 2, 0, 1), (0, 3, 1, 2))          # NHWC x HWIO -> NHWC
# the dimension numbers jax's transpose rules give the two gradients of such a conv:
WGRAD = ConvDimensionNumbers((3, 0, 1, 2), (3, 0, 1, 2), (2, 3, 0, 1))        # lhs: C is "batch", N is contracted; out = HWIO
XGRAD = ConvDimensionNumbers((0, 3, 1, 2), (2, 3, 0, 1), (0, 3, 1, 2))        # rhs: flipped filter with I/O swapped


def _same_pads(size, k, s):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv_fwd(x, w, s):
    pads = [_same_pads(x.shape[1], w.shape[0], s), _same_pads(x.shape[2], w.shape[1], s)]
    return lax.conv_general_dilated(x, w, (s, s), pads, dimension_numbers=FWD), pads


def conv_wgrad(x, dy, pads, k, s):
    """d loss / d w: the stride becomes the dilation of the cotangent 'filter' (jax _conv_general_dilated_transpose_rhs)."""
    return lax.conv_general_dilated(x, dy, (1, 1), pads, rhs_dilation=(s, s), dimension_numbers=WGRAD)


def conv_xgrad(dy, w, pads, s):
    """d loss / d x: cotangent dilated by the stride, filter reversed (jax _conv_general_dilated_transpose_lhs)."""
    kh, kw = w.shape[0], w.shape[1]
    p = [(kh - 1 - pads[0][0], kh - 1 - pads[0][1]), (kw - 1 - pads[1][0], kw - 1 - pads[1][1])]
    return lax.conv_general_dilated(dy, lax.rev(w, (0, 1)), (1, 1), p, lhs_dilation=(s, s), dimension_numbers=XGRAD)


def convnet_train_step(x, y, st):
    B = x.shape[0]
    zero = lambda like: jnp.broadcast_to(jnp.asarray(np.float32(0.0)), like.shape)
    y1, p1 = conv_fwd(x, st['c1']['w'], 2)
    y1 = y1 + jnp.broadcast_to(st['c1']['b'], y1.shape)
    a1 = nn.relu(y1)
    y2, p2 = conv_fwd(a1, st['c2']['w'], 2)
    y2 = y2 + jnp.broadcast_to(st['c2']['b'], y2.shape)
    a2 = nn.relu(y2)
    f = a2.reshape(B, -1)
    z = jnp.dot(f, st['fc']['w']) + jnp.broadcast_to(st['fc']['b'], (B, 10))
    loss, dz = scce_value_and_grad(z, y)
    dot_t = lambda a, b, ca, cb: lax.dot_general(a, b, (((ca,), (cb,)), ((), ())))
    dWf, dbf = dot_t(f, dz, 0, 0), jnp.sum(dz, axis=0)
    dy2 = lax.select(y2 > 0.0, dot_t(dz, st['fc']['w'], 1, 1).reshape(a2.shape), zero(y2))
    dw2, db2 = conv_wgrad(a1, dy2, p2, 3, 2), jnp.sum(dy2, axis=(0, 1, 2))
    dy1 = lax.select(y1 > 0.0, conv_xgrad(dy2, st['c2']['w'], p2, 2), zero(y1))
    dw1, db1 = conv_wgrad(x, dy1, p1, 3, 2), jnp.sum(dy1, axis=(0, 1, 2))
    sgd = lambda w, g: w - LR * g
    new = {'c1': {'w': sgd(st['c1']['w'], dw1), 'b': sgd(st['c1']['b'], db1)},
           'c2': {'w': sgd(st['c2']['w'], dw2), 'b': sgd(st['c2']['b'], db2)},
           'fc': {'w': sgd(st['fc']['w'], dWf), 'b': sgd(st['fc']['b'], dbf)}}
    return {'loss': loss, 'logits': z}, new, {'dw1': dw1, 'dw2': dw2, 'dy1': dy1}


def test_convnet_training_step():
    """≙ reference tests/test_elegy_conv.py:51-94 ([5,32,32,3]); outputs/states atol 1e-6, gradients (the reference's
    intermediates) atol 5e-5 (:94).  The backward pass is validated against a central-difference gradient of the oracle."""
    rs = np.random.RandomState(6)
    x = rs.random_sample((5, 32, 32, 3)).astype(np.float32)
    y = rs.randint(0, 10, size=5).astype(np.int32)
    st = nets.ConvNet().init(7)
    out, new, grads = vkjax.wrap(convnet_train_step)(x, y, st)
    (out_t, new_t, grads_t), _ = oracle(convnet_train_step, [x, y, st])
    assert_tree_close(out, out_t, rtol=1e-5, atol=1e-6)
    assert_tree_close(new, new_t, rtol=1e-5, atol=1e-6)
    assert_tree_close(grads, grads_t, rtol=1e-5, atol=5e-5)
    assert grads['dw1'].shape == (3, 3, 3, 32) and grads['dw2'].shape == (3, 3, 32, 32) and grads['dy1'].shape == (5, 16, 16, 32)



def test_conv_gradient_formulas_are_the_gradients():
    """conv is linear, so for L = sum(conv(x, w) * R) a unit step in one entry changes L by exactly dL/d(entry): checks the
    hand-written weight / input gradient convolutions (lhs-dilated, reversed filter, permuted dimension numbers) against
    the forward conv, through the GPU path and through the oracle."""
    rs = np.random.RandomState(0)
    for H, C, O in ((16, 32, 32), (32, 3, 32)):
        x, w = rs.normal(size=(2, H, H, C)).astype(np.float32), rs.normal(size=(3, 3, C, O)).astype(np.float32)
        R = rs.normal(size=(2, H // 2, H // 2, O)).astype(np.float32)
        pads = [_same_pads(H, 3, 2)] * 2
        fwd = vkjax.wrap(lambda a, b: conv_fwd(a, b, 2)[0])
        L = lambda x_, w_: float((np.asarray(fwd(x_, w_), np.float64) * R).sum())
        dw = vkjax.wrap(lambda a, r: conv_wgrad(a, r, pads, 3, 2))(x, R)
        dx = vkjax.wrap(lambda r, b: conv_xgrad(r, b, pads, 2))(R, w)
        assert_tree_close(dw, oracle(lambda a, r: conv_wgrad(a, r, pads, 3, 2), [x, R])[0], rtol=1e-5, atol=2e-4)
        assert_tree_close(dx, oracle(lambda r, b: conv_xgrad(r, b, pads, 2), [R, w])[0], rtol=1e-5, atol=2e-4)
        base = L(x, w)
        for idx in ((0, 0, 0, 0), (2, 1, C - 1, 3), (1, 2, 1, O - 1)):
            d = np.zeros_like(w); d[idx] = 1.0
            assert abs((L(x, w + d) - base) - dw[idx]) < 1e-3 * max(1.0, abs(dw[idx]))
        for idx in ((0, 0, 0, 0), (1, H - 1, 5, C - 1), (0, 7, H - 1, 1)):
            d = np.zeros_like(x); d[idx] = 1.0
            assert abs((L(x + d, w) - base) - dx[idx]) < 1e-3 * max(1.0, abs(dx[idx]))
