"""SURVEY.md §8(f) rows that go beyond the reference's handlers: select_and_scatter_add (max / min pool gradient, f2),
dot_general with batch dimensions (f1; the reference asserts them away at ops.py:280), and the modern-JAX jaxpr dialect
(f1) read from a checked-in text dump that the repo's own tracer did not emit (tests/golden/modern_jax_jaxpr.txt)."""
import os

import numpy as np
import pytest

import vkjax_b200 as vkjax
from vkjax_b200.frontend import lax, jnp
from common import check, oracle


def _maxpool_grad(x, g, window, strides, padding, select):
    return lax.select_and_scatter_add(g, x, select, window, strides, padding)


POOLS = [
    ('3x3 s2 SAME max (ResNet stem pool)', (2, 12, 13, 8), (1, 3, 3, 1), (1, 2, 2, 1), 'SAME', 'ge'),
    ('2x2 s2 VALID max', (3, 8, 10, 5), (1, 2, 2, 1), (1, 2, 2, 1), 'VALID', 'ge'),
    ('2x2 s1 VALID max (overlapping windows)', (2, 7, 6, 4), (1, 2, 2, 1), (1, 1, 1, 1), 'VALID', 'ge'),
    ('3x3 s2 SAME min', (2, 9, 9, 4), (1, 3, 3, 1), (1, 2, 2, 1), 'SAME', 'le'),
]


def test_select_and_scatter_add_oracle_is_the_transpose_of_pooling():
    """CPU: <pool_grad(g), dx> == d/deps <g, pool(x + eps dx)> for distinct inputs (the oracle's own sanity, no GPU)."""
    from oracle.eval_jaxpr import select_and_scatter_add, reduce_window
    rs = np.random.RandomState(0)
    x = rs.permutation(2 * 8 * 9 * 3).reshape(2, 8, 9, 3).astype(np.float64)       # distinct values: unique maxima
    g = rs.randn(2, 4, 5, 3)
    win, st, pad = (1, 3, 3, 1), (1, 2, 2, 1), [(0, 0), (0, 1), (1, 1), (0, 0)]
    dx = rs.randn(*x.shape)
    eps = 1e-3
    grad = select_and_scatter_add(g, x, 'ge', win, st, pad)
    num = (reduce_window(x + eps * dx, 'max', win, st, pad) - reduce_window(x - eps * dx, 'max', win, st, pad)) / (2 * eps)
    assert np.allclose((grad * dx).sum(), (g * num).sum(), rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize('desc,xs,window,strides,padding,select', POOLS, ids=[p[0] for p in POOLS])
@pytest.mark.parametrize('ties', [False, True], ids=['distinct', 'with_ties'])
def test_select_and_scatter_add(desc, xs, window, strides, padding, select, ties):
    rs = np.random.RandomState(3)
    x = rs.randint(0, 4, xs).astype(np.float32) if ties else rs.random_sample(xs).astype(np.float32)
    pooled_shape = np.asarray(oracle(lambda x: lax.reduce_window(x, -jnp.inf, lax.max, window, strides, padding), [x])[0]).shape
    g = rs.random_sample(pooled_shape).astype(np.float32)
    f = lambda x, g: _maxpool_grad(x, g, window, strides, padding, select)
    y, ytrue = check(f, [x, g], 1e-6, 1e-7)
    assert np.isclose(y.sum(), g.sum(), rtol=1e-5)            # every window hands its value to exactly one element


BATCHED = [
    ('bmk,bkn', (3, 17, 20), (3, 20, 12), 2, 1),
    ('bmk,bnk', (4, 9, 33), (4, 6, 33), 2, 2),
    ('bkm,bkn', (2, 40, 8), (2, 40, 16), 1, 1),
    ('2 batch dims', (2, 3, 8, 16), (2, 3, 16, 4), 3, 2),
    ('tensor cores: 2 x 128x256x64', (2, 128, 256), (2, 256, 64), 2, 1),
]


@pytest.mark.gpu
@pytest.mark.parametrize('desc,sa,sb,ca,cb', BATCHED, ids=[b[0] for b in BATCHED])
@pytest.mark.parametrize('precision', ['fp32', 'tf32', 'simt'])
def test_dot_general_batch_dims(desc, sa, sb, ca, cb, precision):
    rs = np.random.RandomState(5)
    a, b = rs.random_sample(sa).astype(np.float32), rs.random_sample(sb).astype(np.float32)
    nb = len(sa) - 2
    f = lambda a, b: lax.dot_general(a, b, (((ca,), (cb,)), (tuple(range(nb)), tuple(range(nb))))) + 1.0
    tol = (2e-3, 1e-6) if precision == 'tf32' else (1e-5, 1e-8)
    check(f, [a, b], *tol, precision=precision)


# ---- modern-JAX jaxpr dumps (text), not emitted by this repo's tracer ----------------------------------------------
def _modern_cases():
    rs = np.random.RandomState(0)
    sig = lambda x: 1.0 / (1.0 + np.exp(-x.astype(np.float64)))

    def softmax(q, k):
        s = np.einsum('bqd,bkd->bqk', q.astype(np.float64), k.astype(np.float64)) / 4.0
        e = np.exp(s - s.max(-1, keepdims=True))
        return e / e.sum(-1, keepdims=True)

    def conv_avgpool(x, w):
        import torch
        y = torch.nn.functional.conv2d(torch.from_numpy(x).double().permute(0, 3, 1, 2), torch.from_numpy(w).double().permute(3, 2, 0, 1), padding=1)
        y = torch.nn.functional.avg_pool2d(torch.relu(y), 2, 2)
        return y.permute(0, 2, 3, 1).numpy()
    f4 = lambda *s: rs.randn(*s).astype(np.float32)
    return {
        'mlp_relu': ([f4(8, 128), f4(128, 16), f4(16)], lambda x, W, b: np.maximum(x.astype(np.float64) @ W + b, 0)),
        'gelu_where_norm': ([f4(4, 32)], lambda x: np.where(x > 0, x, np.float32(0.01) * x) * sig(x) / np.sqrt((x.astype(np.float64) ** 2).mean(-1, keepdims=True) + 1e-6)),
        'batched_matmul_softmax': ([f4(3, 8, 16), f4(3, 12, 16)], softmax),
        'conv_avgpool': ([f4(2, 8, 8, 4), f4(3, 3, 4, 8)], conv_avgpool),
    }


def test_modern_jaxpr_text_parses_and_oracle_matches_numpy():
    """CPU: the reader + the oracle on the modern-dialect dumps against a direct numpy / torch evaluation of the documented function"""
    from vkjax_b200 import jaxpr_text, JaxprInterpreter
    from oracle.eval_jaxpr import eval_jaxpr
    js = jaxpr_text.parse_file(os.path.join(os.path.dirname(__file__), 'golden', 'modern_jax_jaxpr.txt'))
    cases = _modern_cases()
    assert set(js) == set(cases)
    names = {e.primitive.name for j in js.values() for e in j.jaxpr.eqns}
    assert {'pjit', 'custom_jvp_call', 'select_n', 'logistic', 'sqrt', 'integer_pow', 'reduce_window_sum', 'copy', 'stop_gradient'} <= names
    for name, (args, ref) in cases.items():
        y = eval_jaxpr(js[name], *args)[0]
        assert np.allclose(y, ref(*args), rtol=1e-5, atol=1e-5), name
        JaxprInterpreter(js[name], dry_run=True)                # and the executor plans it


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['mlp_relu', 'gelu_where_norm', 'batched_matmul_softmax', 'conv_avgpool'])
@pytest.mark.parametrize('fuse', [True, False], ids=['fused', 'unfused'])
def test_modern_jaxpr_text_on_gpu(name, fuse):
    from vkjax_b200 import jaxpr_text, JaxprInterpreter
    from oracle.eval_jaxpr import eval_jaxpr
    js = jaxpr_text.parse_file(os.path.join(os.path.dirname(__file__), 'golden', 'modern_jax_jaxpr.txt'))
    args, ref = _modern_cases()[name]
    y = JaxprInterpreter(js[name], fuse=fuse).run(*args)[0]
    ytrue = eval_jaxpr(js[name], *args)[0]
    assert y.shape == ytrue.shape
    assert np.allclose(y, ytrue, rtol=1e-5, atol=1e-6)
    assert np.allclose(y, ref(*args), rtol=1e-5, atol=1e-5)


# ---- typed PRNG keys / random_bits: modern jax.random dumps, pinned on JAX's documented values ---------
def _random_cases():
    import json
    G = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'jax_random.json')))
    key = lambda seed: np.array([0, seed], np.uint32)             # key data of jax.random.key(seed), seed < 2**32
    return {
        'normal_key42_shape3': ([key(42)], np.array(G['normal_key42_shape3'], np.float32), 2e-6),
        'uniform_from_seed': ([np.int32(0)], np.float32(G['uniform_key0']), 1e-7),
        'split_raw_key': ([key(0)], np.array(G['split_key0'], np.uint32), 0),
        'fold_in_bits': ([key(42)], None, 0),                      # no documented value: oracle vs the 0.2.x chain vs the GPU
    }


def _random_jaxprs():
    from vkjax_b200 import jaxpr_text
    return jaxpr_text.parse_file(os.path.join(os.path.dirname(__file__), 'golden', 'modern_jax_random_jaxpr.txt'))


def test_typed_key_jaxprs_oracle_matches_documented_values():
    from vkjax_b200 import JaxprInterpreter
    from vkjax_b200.frontend import make_jaxpr, random
    from oracle.eval_jaxpr import eval_jaxpr
    js, cases = _random_jaxprs(), _random_cases()
    assert set(js) == set(cases)
    names = {e.primitive.name for j in js.values() for e in j.jaxpr.eqns}
    assert {'random_seed', 'random_bits', 'random_split', 'random_fold_in', 'random_wrap', 'random_unwrap'} <= names
    for name, (args, documented, rtol) in cases.items():
        y = eval_jaxpr(js[name], *args)[0]
        if documented is not None:
            assert y.shape == documented.shape and y.dtype == documented.dtype, name
            assert np.array_equal(y, documented) if rtol == 0 else np.allclose(y, documented, rtol=rtol, atol=0), name
        JaxprInterpreter(js[name], dry_run=True)
    # the same function through this repo's own tracer (threefry2x32 + iota + slices, the JAX-0.2.x formulation)
    k = np.array([0, 42], np.uint32)
    f = lambda key: random._random_bits(random.fold_in(key, 7), (5,))
    assert np.array_equal(eval_jaxpr(make_jaxpr(f)(k), k)[0], eval_jaxpr(js['fold_in_bits'], k)[0])
    # other seeds: random_seed + uniform against the 0.2.x chain, incl. a negative seed (key data = [0, seed mod 2**32])
    for seed in (1, 12345, -7):
        kd = np.array([0, seed & 0xFFFFFFFF], np.uint32)
        assert eval_jaxpr(js['uniform_from_seed'], np.int32(seed))[0] == eval_jaxpr(make_jaxpr(lambda key: random.uniform(key))(kd), kd)[0]


def test_typed_key_unsupported_forms_raise():
    from vkjax_b200 import jaxpr_text, JaxprInterpreter
    wide = jaxpr_text.parse_jaxpr('{ lambda ; a:key<fry>[]. let b:u8[4] = random_bits[bit_width=8 shape=(4,)] a in (b,) }')
    with pytest.raises(NotImplementedError):
        JaxprInterpreter(wide, dry_run=True)
    rbg = jaxpr_text.parse_jaxpr('{ lambda ; a:i32[]. let b:key<fry>[] = random_seed[impl=rbg] a in (b,) }')
    with pytest.raises(NotImplementedError):
        JaxprInterpreter(rbg, dry_run=True)


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['normal_key42_shape3', 'uniform_from_seed', 'split_raw_key', 'fold_in_bits'])
@pytest.mark.parametrize('fuse', [True, False], ids=['fused', 'unfused'])
def test_typed_key_jaxprs_on_gpu(name, fuse):
    from vkjax_b200 import JaxprInterpreter
    from oracle.eval_jaxpr import eval_jaxpr
    j = _random_jaxprs()[name]
    args, documented, rtol = _random_cases()[name]
    y = JaxprInterpreter(j, fuse=fuse).run(*args)[0]
    ytrue = eval_jaxpr(j, *args)[0]
    assert y.shape == ytrue.shape and y.dtype == ytrue.dtype
    if ytrue.dtype == np.uint32:
        assert np.array_equal(y, ytrue)                              # bits: exact
    else:
        assert np.allclose(y, ytrue, rtol=2e-6, atol=0)             # erf_inv / float chain
    if documented is not None:
        assert np.array_equal(y, documented) if rtol == 0 else np.allclose(y, documented, rtol=rtol, atol=0)


# ---- uint8 inputs (the host/wire side of the call) ----------------------------------------------------
def test_uint8_input_traces_and_plans():
    """CPU: a uint8 image batch is an accepted INPUT dtype; convert_element_type widens it; the convert + /255 chain in
    front of a tensor-core convolution is folded into the conv's re-layout (no f32 image tensor); anything else on
    uint8 raises NotImplementedError as the reference does for unsupported dtypes (ops.py:19-25)."""
    from vkjax_b200 import nets, JaxprInterpreter
    from vkjax_b200.frontend import make_jaxpr
    from oracle.eval_jaxpr import eval_jaxpr
    m = nets.ResNet18()
    st = m.init(0)
    x = np.random.RandomState(0).randint(0, 256, (2, 64, 64, 3)).astype(np.uint8)
    f = lambda x, s: m.apply(s, x.astype(jnp.float32) / 255.0)
    jaxpr = make_jaxpr(f)(x, st)
    assert jaxpr.jaxpr.invars[0].aval.dtype == np.uint8
    it = JaxprInterpreter(jaxpr, dry_run=True, precision='tf32')
    assert it.n_input_chains_fused == 1
    assert it.input_buffers[0].nbytes() == x.size
    y = eval_jaxpr(jaxpr, x, *[l for l in __import__('vkjax_b200').tree_util.tree_leaves(st)])[0]
    assert np.isfinite(y).all()
    with pytest.raises(NotImplementedError):
        JaxprInterpreter(make_jaxpr(lambda x: x + x)(x), dry_run=True)


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(4, 32, 32, 3), (3, 7, 5, 3), (1000,), (13,)])
def test_uint8_convert_on_device(shape):
    x = np.random.RandomState(1).randint(0, 256, shape).astype(np.uint8)
    f = lambda x: x.astype(jnp.float32) / 255.0
    y, ytrue = check(f, [x], 1e-6, 1e-7)
    assert y.dtype == np.float32 and np.array_equal(y, x.astype(np.float32) / np.float32(255.0))
    g = lambda x: x.astype(jnp.int32) * 2 + 1
    y, _ = check(g, [x])
    assert np.array_equal(y, x.astype(np.int32) * 2 + 1)


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['fp32', 'tf32'])
def test_uint8_images_through_resnet_stem(precision):
    """uint8 pixels + astype + /255 fused into the stem's operand load == the same network fed the f32 image, bit for bit"""
    from vkjax_b200 import nets
    from vkjax_b200.elegy import vkModel
    m = nets.ResNet18()
    st = m.init(3)
    xu8 = np.random.RandomState(2).randint(0, 256, (4, 64, 64, 3)).astype(np.uint8)
    xf = xu8.astype(np.float32) / np.float32(255.0)
    f8 = vkjax.wrap(lambda x, s: m.apply(s, x.astype(jnp.float32) / 255.0), precision=precision)
    ff = vkjax.wrap(lambda x, s: m.apply(s, x), precision=precision)
    y8, yf = f8(xu8, st), ff(xf, st)
    interp = list(f8._jaxpr_interpreters.values())[0]
    interp_f = list(ff._jaxpr_interpreters.values())[0]
    assert interp.n_input_chains_fused == 1
    assert interp.h2d_bytes + 3 * xu8.size == interp_f.h2d_bytes                  # the image went up as bytes: 4x less
    assert np.array_equal(y8, yf)
    ytrue, _ = oracle(lambda x, s: m.apply(s, x.astype(jnp.float32) / 255.0), [xu8, st])
    tol = dict(rtol=1e-4, atol=1e-5) if precision == 'fp32' else dict(rtol=5e-2, atol=5e-2)
    assert np.allclose(y8, ytrue, **tol)
