"""Training-step gradients against an independent automatic differentiation (same motive as tests/test_resnet_independent.py):
the front end's reverse-mode rules (vkjax_b200/frontend/autodiff.py) restate jax.lax's transpose rules, and until now their only
judges were the repo's own hand-derived backward passes and central differences of the repo's own oracle.  Here the loss and the
gradients of the reference's two trainable models -- the LeNet-300-100 MLP of tests/test_elegy_mlp.py:14-33 and the strided ConvNet
of tests/test_elegy_conv.py:14-23, sparse categorical cross-entropy on logits -- are computed by torch.autograd on float64 from
functions written in plain torch (no tracer, no oracle, no nets.apply), and the jaxpr that `value_and_grad` traces must reproduce
them through the oracle.  The ResNet-18 case adds BatchNorm, max-pool (select_and_scatter_add), residual adds and the mean."""
import numpy as np
import pytest

from vkjax_b200 import nets, tree_util
from vkjax_b200.elegy import losses
from vkjax_b200.frontend import make_jaxpr, value_and_grad
from oracle.eval_jaxpr import eval_jaxpr
from test_resnet_independent import _same_pad, torch_resnet


def _torch():
    import torch
    return torch, torch.nn.functional


def _t(a, grad=False):
    torch, _ = _torch()
    return torch.tensor(np.asarray(a, np.float64), requires_grad=grad)


def _conv_same(x, w, stride):
    """NCHW x, OIHW w, XLA SAME padding"""
    _, F = _torch()
    ph, pw = _same_pad(x.shape[2], w.shape[2], stride), _same_pad(x.shape[3], w.shape[3], stride)
    return F.conv2d(F.pad(x, (pw[0], pw[1], ph[0], ph[1])), w, stride=stride)


def _scce(logits, labels):
    torch, F = _torch()
    return F.cross_entropy(logits, torch.from_numpy(labels.astype(np.int64)))            # mean over the batch of -log softmax[label]


def torch_mlp(params, x, labels):
    torch, _ = _torch()
    ws = [(_t(l['w'], True), _t(l['b'], True)) for l in params]
    h = _t(x).reshape(x.shape[0], -1) / 255.0
    for i, (w, b) in enumerate(ws):
        h = h @ w + b
        if i < len(ws) - 1:
            h = torch.relu(h)
    loss = _scce(h, labels)
    loss.backward()
    return loss.item(), h.detach().numpy(), [{'w': w.grad.numpy(), 'b': b.grad.numpy()} for w, b in ws]


def torch_convnet(params, x, labels):
    torch, _ = _torch()
    leaf = {k: {'w': _t(v['w'], True), 'b': _t(v['b'], True)} for k, v in params.items()}
    h = _t(x).permute(0, 3, 1, 2)
    for k in ('c1', 'c2'):
        h = torch.relu(_conv_same(h, leaf[k]['w'].permute(3, 2, 0, 1), 2) + leaf[k]['b'].reshape(1, -1, 1, 1))
    h = h.permute(0, 2, 3, 1).reshape(x.shape[0], -1)                                       # Flatten of the NHWC activation
    logits = h @ leaf['fc']['w'] + leaf['fc']['b']
    loss = _scce(logits, labels)
    loss.backward()
    return loss.item(), logits.detach().numpy(), {k: {'w': v['w'].grad.numpy(), 'b': v['b'].grad.numpy()} for k, v in leaf.items()}


def _frontend_step(module, x, labels, params):
    """loss, logits and gradients through the repo's tracer + autodiff, evaluated by the oracle"""
    def f(x, y, p):
        def loss_fn(p):
            logits = module.apply(p, x)
            return losses.SparseCategoricalCrossentropy(from_logits=True)(y, logits), logits
        (loss, logits), g = value_and_grad(loss_fn, has_aux=True)(p)
        return loss, logits, g
    jaxpr, shapes = make_jaxpr(f, return_shape=True)(x, labels, params)
    out = eval_jaxpr(jaxpr, *[np.asarray(l) for l in tree_util.tree_leaves((x, labels, params))])
    return tree_util.tree_unflatten(tree_util.tree_structure(shapes), out)


def _assert_grads(got, want, rtol, atol):
    gl, wl = tree_util.tree_leaves(got), tree_util.tree_leaves(want)
    assert len(gl) == len(wl)
    for a, b in zip(gl, wl):
        assert a.shape == b.shape
        assert np.abs(b).max() > 0                                           # the gradient carries signal
        assert np.allclose(a, b, rtol=rtol, atol=atol * max(1.0, float(np.abs(b).max()))), float(np.abs(a - b).max())


def test_mlp_gradients_match_torch_autograd():
    rs = np.random.RandomState(11)
    x = (rs.random_sample((16, 32, 32, 3)) * 255).astype(np.float32)
    labels = rs.randint(0, 10, size=16).astype(np.int32)
    params = nets.MLP().init(3)
    loss, logits, g = _frontend_step(nets.MLP(), x, labels, params)
    tl, tlogits, tg = torch_mlp(params, x, labels)
    assert np.isclose(loss, tl, rtol=1e-5)
    assert np.allclose(logits, tlogits, rtol=1e-4, atol=1e-5)
    _assert_grads(g, tg, rtol=1e-4, atol=1e-6)


def test_convnet_gradients_match_torch_autograd():
    rs = np.random.RandomState(12)
    x = rs.random_sample((6, 32, 32, 3)).astype(np.float32)
    labels = rs.randint(0, 10, size=6).astype(np.int32)
    params = nets.ConvNet().init(7)
    loss, logits, g = _frontend_step(nets.ConvNet(), x, labels, params)
    tl, tlogits, tg = torch_convnet(params, x, labels)
    assert np.isclose(loss, tl, rtol=1e-5)
    assert np.allclose(logits, tlogits, rtol=1e-4, atol=1e-5)
    _assert_grads(g, tg, rtol=1e-4, atol=1e-6)


def test_convnet_gradients_odd_image_size():
    """29 x 23 images: both stride-2 SAME convolutions pad asymmetrically, forward and in the transposed convolutions of the backward pass"""
    rs = np.random.RandomState(13)
    x = rs.random_sample((3, 29, 23, 3)).astype(np.float32)
    labels = rs.randint(0, 10, size=3).astype(np.int32)
    params = nets.ConvNet().init(8, in_shape=(29, 23, 3))
    loss, logits, g = _frontend_step(nets.ConvNet(), x, labels, params)
    tl, tlogits, tg = torch_convnet(params, x, labels)
    assert np.isclose(loss, tl, rtol=1e-5)
    _assert_grads(g, tg, rtol=1e-4, atol=1e-6)


def test_resnet18_gradients_match_torch_autograd():
    """BatchNorm on running statistics, 3x3/2 max-pool (-> select_and_scatter_add), residual adds, global mean: every trainable leaf"""
    torch, _ = _torch()
    rs = np.random.RandomState(14)
    x = rs.random_sample((2, 40, 36, 3)).astype(np.float32)
    labels = rs.randint(0, 1000, size=2).astype(np.int32)
    m = nets.ResNet18()
    params = m.init(9)
    loss, logits, g = _frontend_step(m, x, labels, params)
    leaves = {}

    def t(a):                                    # one torch leaf per weight array, found again by identity
        if id(a) not in leaves:
            leaves[id(a)] = (a, _t(a, True))
        return leaves[id(a)][1]
    tlogits = torch_resnet(params, x, (2, 2, 2, 2), False, t=t, as_numpy=False)
    tloss = _scce(tlogits, labels)
    tloss.backward()
    assert np.isclose(loss, tloss.item(), rtol=1e-5)
    assert np.allclose(logits, tlogits.detach().numpy(), rtol=1e-4, atol=1e-5)
    pl, gl = tree_util.tree_leaves(params), tree_util.tree_leaves(g)
    assert len(pl) == len(gl)
    checked = 0
    for w, gw in zip(pl, gl):
        tw = leaves[id(w)][1]
        want = tw.grad.numpy().reshape(np.shape(w))
        assert np.allclose(gw, want, rtol=1e-4, atol=1e-6 * max(1.0, float(np.abs(want).max()))), float(np.abs(gw - want).max())
        checked += 1
    assert checked == len(leaves) > 60
