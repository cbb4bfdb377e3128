"""The reference's training / test / initialisation tests restated THROUGH THE MODEL:
tests/test_elegy_mlp.py:87-148 (train_on_batch of the LeNet-300-100 MLP with SCCE + SGD(0.1), every intermediate compared;
initialisation with a seed) and tests/test_elegy_conv.py:51-94 (the same for the strided ConvNet), i.e.
vkModel(module, loss=..., optimizer=...).train_on_batch(x, y) -> call_train_step_jit -> vkjax.wrap -> jaxpr -> GPU.

Elegy / optax / jax are not installable here: the step functions are the stand-in's (vkjax_b200/elegy.py), the gradients
come from the front end's reverse-mode rules (vkjax_b200/frontend/autodiff.py, restating jax.lax's transpose rules) and the
truth is the numpy oracle evaluating the SAME jaxpr on the same inputs.  The CPU tests below additionally pin the autodiff
against the hand-derived backward passes of tests/test_train_step.py and against central differences of the oracle.
"""
import numpy as np
import pytest

import vkjax_b200 as vkjax
from vkjax_b200 import nets, tree_util, JaxprInterpreter, core
from vkjax_b200.elegy import vkModel, losses, optimizers, TrainStates
from vkjax_b200.frontend import make_jaxpr, value_and_grad, grad, lax, jnp, nn, random
from oracle.eval_jaxpr import eval_jaxpr
from common import assert_tree_close

LR = 0.1


def _scce(logits, labels):
    return losses.SparseCategoricalCrossentropy(from_logits=True)(labels, logits)


def _oracle_of(fn_jit, args):
    """evaluates the jaxpr the wrapped step function was traced to, on the same (flattened, non-static) inputs"""
    interp = list(fn_jit._jaxpr_interpreters.values())[-1]
    dyn = [a for i, a in enumerate(args) if i not in fn_jit._static_argnums]
    leaves = [np.asarray(l) for l in tree_util.tree_leaves(dyn)]
    return eval_jaxpr(interp.jaxpr, *leaves), interp


# ---- CPU: the autodiff itself -----------------------------------------------------------------------
def test_autodiff_matches_hand_derived_backward_passes():
    import test_train_step as T
    rs = np.random.RandomState(2)

    def step(module):
        def f(x, y, params):
            def loss_fn(p):
                logits = module.apply(p, x)
                return _scce(logits, y), logits
            (loss, logits), g = value_and_grad(loss_fn, has_aux=True)(params)
            return {'loss': loss, 'logits': logits}, tree_util.tree_map(lambda w, gw: w - LR * gw, params, g)
        return f
    x = (rs.random_sample((8, 32, 32, 3)) * 255).astype(np.float32)
    y = rs.randint(0, 10, size=8).astype(np.int32)
    params = nets.MLP().init(3)
    leaves = tree_util.tree_leaves((x, y, params))
    got = eval_jaxpr(make_jaxpr(step(nets.MLP()))(x, y, params), *leaves)
    want = eval_jaxpr(make_jaxpr(T.mlp_train_step)(x, y, params), *leaves)
    for a, b in zip(got, want):
        assert np.allclose(a, b, rtol=1e-6, atol=1e-7)
    x = rs.random_sample((5, 32, 32, 3)).astype(np.float32)
    y = rs.randint(0, 10, size=5).astype(np.int32)
    st = nets.ConvNet().init(7)
    leaves = tree_util.tree_leaves((x, y, st))
    got = eval_jaxpr(make_jaxpr(step(nets.ConvNet()))(x, y, st), *leaves)
    want = eval_jaxpr(make_jaxpr(T.convnet_train_step)(x, y, st), *leaves)
    for a, b in zip(got, want[:len(got)]):
        assert np.allclose(a, b, rtol=1e-6, atol=1e-7)


def test_autodiff_rules_against_central_differences():
    """every transpose rule used by a model: composite function, gradient vs central differences of the oracle (float64-ish)"""
    rs = np.random.RandomState(0)
    dn = core.ConvDimensionNumbers((0, 3, 1, 2), (3, 2, 0, 1), (0, 3, 1, 2))

    def f(x, w, v, b):
        h = lax.conv_general_dilated(x, w, (2, 2), 'SAME', dimension_numbers=dn)           # conv (strided, padded)
        h = nn.relu(h + jnp.broadcast_to(b, h.shape))                                       # add, broadcast, relu
        h = lax.reduce_window(h, -jnp.inf, lax.max, (1, 2, 2, 1), (1, 2, 2, 1), 'VALID')    # max-pool -> select_and_scatter_add
        h = jnp.tanh(h) * lax.logistic(h) - jnp.exp(-h) / (1.0 + h * h)                     # tanh, logistic, exp, div, mul, sub, neg
        g = jnp.mean(h, axis=(1, 2))                                                        # reduce_sum, div
        z = jnp.dot(g, v)                                                                   # dot_general
        z = z - jnp.max(z, axis=-1, keepdims=True)                                          # reduce_max, reshape
        return jnp.sum(jnp.log(1.0 + z * z) * jnp.sqrt(2.0 + z * z) + lax.rsqrt(3.0 + z * z))   # log, sqrt, rsqrt
    args = [rs.normal(size=(2, 9, 8, 3)), rs.normal(size=(3, 3, 3, 4)) * 0.5, rs.normal(size=(4, 5)), rs.normal(size=(4,)) * 0.1]
    args = [a.astype(np.float32) for a in args]
    gj = make_jaxpr(grad(f, argnums=(0, 1, 2, 3)))(*args)
    grads = eval_jaxpr(gj, *args)
    fj = make_jaxpr(f)(*args)
    F = lambda a: float(eval_jaxpr(fj, *a)[0])
    for k, g in enumerate(grads):
        assert g.shape == args[k].shape
        for _ in range(4):
            idx = tuple(rs.randint(0, s) for s in args[k].shape)
            eps = 1e-2
            hi = [a.copy() for a in args]; hi[k][idx] += eps
            lo = [a.copy() for a in args]; lo[k][idx] -= eps
            num = (F(hi) - F(lo)) / (2 * eps)
            assert abs(num - g[idx]) < 2e-2 * max(1.0, abs(num)), (k, idx, num, g[idx])


def test_train_step_plans_without_a_gpu():
    """host logic only: trace call_train_step through the stand-in, fuse and plan it (dry run)"""
    m = vkModel(nets.MLP(), loss=losses.SparseCategoricalCrossentropy(), optimizer=optimizers.sgd(LR))
    x = np.zeros((8, 32, 32, 3), np.float32)
    y = np.zeros((8,), np.int32)
    st = TrainStates(nets.MLP().init(0), ())
    jaxpr = make_jaxpr(m.call_train_step, static_argnums=[5, 6])(x, y, None, None, st, False, True)
    names = {e.primitive.name for e in jaxpr.jaxpr.eqns}
    assert {'dot_general', 'select', 'scatter-add', 'gather', 'custom_jvp_call_jaxpr'} <= names
    it = JaxprInterpreter(jaxpr, dry_run=True)
    assert len(it.all_ops) < len(jaxpr.jaxpr.eqns)


# ---- GPU: through the model ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('module_cls,xshape', [(nets.MLP, (8, 32, 32, 3)), (nets.ConvNet, (5, 32, 32, 3))], ids=['mlp', 'convnet'])
def test_basic_training(module_cls, xshape):
    """≙ reference tests/test_elegy_mlp.py:87-118 / tests/test_elegy_conv.py:51-83: logs and the updated states of one
    train_on_batch (reference atol 1e-7 against XLA:CPU fp32; the oracle accumulates in float64: atol 1e-6)."""
    rs = np.random.RandomState(11)
    x = (rs.random_sample(xshape) * (255 if module_cls is nets.MLP else 1)).astype(np.float32)
    y = rs.randint(0, 10, size=xshape[0]).astype(np.int32)
    model = vkModel(module_cls(), loss=losses.SparseCategoricalCrossentropy(from_logits=True), optimizer=optimizers.sgd(LR))
    model.init(x, y, seed=3, host=True)
    before = tree_util.tree_map(np.asarray, model.states)
    logs = model.train_on_batch(x, y)
    args = (x, y, None, None, TrainStates(before, ()), False, True)
    outs, interp = _oracle_of(model.call_train_step_jit, args)
    assert np.allclose(logs['loss'], outs[0], rtol=1e-5, atol=1e-6)
    after = tree_util.tree_leaves(tree_util.tree_map(np.asarray, model.states))
    truth = outs[-len(after):]
    assert len(after) == len(tree_util.tree_leaves(before))
    for a, b, w0 in zip(after, truth, tree_util.tree_leaves(before)):
        assert a.shape == b.shape == w0.shape
        assert np.allclose(a, b, rtol=1e-5, atol=1e-6)
    moved = max(float(np.abs(a - w0).max()) for a, w0 in zip(after, tree_util.tree_leaves(before)))
    assert moved > 1e-4                                        # the step did move the weights
    # a second step runs on the updated, device-resident states and lowers the loss on the same batch
    loss2 = model.train_on_batch(x, y)['loss']
    assert float(loss2) < float(logs['loss'])


@pytest.mark.gpu
@pytest.mark.parametrize('module_cls,xshape', [(nets.MLP, (8, 32, 32, 3)), (nets.ConvNet, (5, 32, 32, 3))], ids=['mlp', 'convnet'])
def test_training_every_intermediate(module_cls, xshape):
    """≙ reference tests/test_elegy_mlp.py:120-129 / tests/test_elegy_conv.py:85-94: the train-step jaxpr re-instantiated
    without buffer reuse, every variable compared (reference atol 1e-6 MLP / 5e-5 ConvNet)."""
    rs = np.random.RandomState(12)
    x = (rs.random_sample(xshape) * (255 if module_cls is nets.MLP else 1)).astype(np.float32)
    y = rs.randint(0, 10, size=xshape[0]).astype(np.int32)
    model = vkModel(module_cls(), loss=losses.SparseCategoricalCrossentropy(from_logits=True), optimizer=optimizers.sgd(LR))
    st = TrainStates(module_cls().init(5), ())
    jaxpr = make_jaxpr(model.call_train_step, static_argnums=[5, 6])(x, y, None, None, st, False, True)
    leaves = tree_util.tree_leaves((x, y, st))
    _, envtrue = eval_jaxpr(jaxpr, *leaves, return_env=True)
    interp = JaxprInterpreter(jaxpr, reuse_buffers=False, fuse=False, precision='simt')
    _, envpred = interp.run(*leaves, return_all=True)
    checked = 0
    for var, vtrue in envtrue.items():
        got = envpred.get(core.hashable(var))
        if got is None:
            continue
        vtrue = np.asarray(vtrue)
        assert got.shape == vtrue.shape, (var, got.shape, vtrue.shape)
        if vtrue.dtype.kind == 'f':
            assert np.allclose(got, vtrue, rtol=1e-5, atol=5e-5, equal_nan=True), (str(var), float(np.abs(got - vtrue).max()))
        else:
            assert np.array_equal(got, vtrue), str(var)
        checked += 1
    assert checked > 40


@pytest.mark.gpu
def test_test_step():
    rs = np.random.RandomState(13)
    x = (rs.random_sample((8, 32, 32, 3)) * 255).astype(np.float32)
    y = rs.randint(0, 10, size=8).astype(np.int32)
    model = vkModel(nets.MLP(), loss=losses.SparseCategoricalCrossentropy(from_logits=True))
    model.init(x, seed=2, host=True)
    logs = model.test_on_batch(x, y)
    outs, _ = _oracle_of(model.call_test_step_jit, (x, y, None, None, tree_util.tree_map(np.asarray, model.states), False, False))
    assert np.allclose(logs['loss'], outs[0], rtol=1e-5, atol=1e-6)
    assert abs(float(logs['loss']) - np.log(10)) < 1.5       # ~ ln(10) for random weights


@pytest.mark.gpu
@pytest.mark.parametrize('module_cls', [nets.MLP, nets.ConvNet, nets.ResNet18], ids=['mlp', 'convnet', 'resnet18'])
def test_basic_initialization(module_cls):
    """≙ reference tests/test_elegy_mlp.py:134-148: vkModel(module, seed=1).init(x) runs the initialisers on the device
    (threefry2x32 / erf_inv chains); compared with the oracle on the same init jaxpr.  The reference needs atol 2e-3
    ("limited by erf_inv", its Winitzki approximation); erfinvf here is accurate: rtol 1e-5 / atol 1e-6."""
    x = np.random.RandomState(14).random_sample((2, 32, 32, 3)).astype(np.float32)
    model = vkModel(module_cls(), seed=1)
    model.init(x)
    states = tree_util.tree_map(np.asarray, model.states)
    outs, interp = _oracle_of(model.call_init_step_jit, (x, random.PRNGKey(1)))
    leaves = tree_util.tree_leaves(states)
    assert len(leaves) == len(outs)
    for a, b in zip(leaves, outs):
        assert a.shape == b.shape and np.allclose(a, b, rtol=1e-5, atol=1e-6)
    # a different seed gives different weights; the same seed the same ones (bit for bit)
    again = vkModel(module_cls(), seed=1)
    again.init(x)
    other = vkModel(module_cls(), seed=2)
    other.init(x)
    la, lo = tree_util.tree_leaves(tree_util.tree_map(np.asarray, again.states)), tree_util.tree_leaves(tree_util.tree_map(np.asarray, other.states))
    assert all(np.array_equal(a, b) for a, b in zip(leaves, la))
    assert any(a.std() > 0 and not np.array_equal(a, b) for a, b in zip(leaves, lo))
    # and the initialised model predicts
    y = model.predict(x if module_cls is not nets.ResNet18 else np.zeros((1, 64, 64, 3), np.float32))
    assert np.all(np.isfinite(y))
