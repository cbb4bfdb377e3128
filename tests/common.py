"""Shared helpers of the parity tests (≙ reference tests/common.py).

`check(f, args)` is the reference's differential test (tests/test_basic_ops.py:336-362) with the JAX-CPU
truth replaced by the numpy oracle evaluating the *same jaxpr on the same inputs*: pytree structure,
shapes and dtypes must be equal, values within the stated tolerance (bit-exact for integer/bool results).
"""
import numpy as np

import vkjax_b200 as vkjax
from vkjax_b200 import tree_util
from vkjax_b200.frontend import make_jaxpr
from oracle.eval_jaxpr import eval_jaxpr


def as_device_dtype(x):
    """What jnp.asarray does to host values without x64 (reference tests/test_basic_ops.py:340)."""
    x = np.asarray(x)
    if x.dtype == np.float64:
        return x.astype(np.float32)
    if x.dtype == np.int64:
        return x.astype(np.int32)
    if x.dtype == np.uint64:
        return x.astype(np.uint32)
    return x


def oracle(f, args, static_argnums=()):
    jaxpr, shapes = make_jaxpr(f, static_argnums, return_shape=True)(*args)
    dyn = [a for i, a in enumerate(args) if i not in static_argnums]
    outs = eval_jaxpr(jaxpr, *tree_util.tree_leaves(dyn))
    flat_shapes = tree_util.tree_leaves(shapes)
    outs = [np.asarray(o).reshape(s.shape) for o, s in zip(outs, flat_shapes)]
    return tree_util.tree_unflatten(tree_util.tree_structure(shapes), outs), jaxpr


def assert_tree_close(y, ytrue, rtol=1e-5, atol=1e-8, exact_ints=True):
    assert tree_util.tree_structure(y) == tree_util.tree_structure(ytrue)
    for a, b in zip(tree_util.tree_leaves(y), tree_util.tree_leaves(ytrue)):
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape, (a.shape, b.shape)
        assert a.dtype == b.dtype, (a.dtype, b.dtype)
        if exact_ints and a.dtype.kind in 'iub':
            assert np.array_equal(a, b), f'integer/bool result differs in {np.sum(a != b)} of {a.size} elements'
        else:
            ok = np.allclose(a, b, rtol, atol, equal_nan=True)
            if not ok:
                err = np.abs(a.astype(np.float64) - b.astype(np.float64))
                rel = err / np.maximum(np.abs(b.astype(np.float64)), 1e-30)
                raise AssertionError(f'max abs err {np.nanmax(err):.3e}, max rel err {np.nanmax(rel):.3e} '
                                     f'(rtol={rtol}, atol={atol})')


def check(f, args, rtol=1e-5, atol=1e-8, **wrap_kwargs):
    args = tree_util.tree_map(as_device_dtype, list(args))
    ytrue, jaxpr = oracle(f, args)
    vkfunc = vkjax.Function(f, **wrap_kwargs)
    y = vkfunc(*args)
    assert_tree_close(y, ytrue, rtol, atol)
    return y, ytrue
