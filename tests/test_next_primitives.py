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
