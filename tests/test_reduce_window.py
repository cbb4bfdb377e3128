"""Restatement of reference tests/test_reduce_window.py (:14-21) + negative inputs (expose quirk Q3: lax pads
with -inf, the reference shader with 0.0) + the avg-pool building block reduce_window_sum (no reference handler)."""
import numpy as np
import pytest

import vkjax_b200 as vkjax
from vkjax_b200.frontend import lax, jnp
from common import check

pytestmark = pytest.mark.gpu
rng = np.random.RandomState(11)
R = rng.random_sample


def reduce_window_max0(x): return lax.reduce_window(x, -jnp.inf, lax.max, (1, 2, 2, 1), window_strides=(1, 1, 1, 1), padding='VALID')
def reduce_window_max1(x): return lax.reduce_window(x, -jnp.inf, lax.max, (1, 3, 3, 1), window_strides=(1, 2, 2, 1), padding='SAME')
def reduce_window_sum0(x): return lax.reduce_window(x, 0.0, lax.add, (1, 3, 3, 1), window_strides=(1, 2, 2, 1), padding='SAME')
def avg_pool_2x2(x): return lax.reduce_window(x, 0.0, lax.add, (1, 2, 2, 1), window_strides=(1, 2, 2, 1), padding='VALID') / 4.0
def reduce_window_min0(x): return lax.reduce_window(x, jnp.inf, lax.min, (1, 3, 3, 1), window_strides=(1, 2, 2, 1), padding='SAME')
def reduce_window_max_s1(x): return lax.reduce_window(x, -jnp.inf, lax.max, (1, 3, 3, 1), window_strides=(1, 1, 1, 1), padding='SAME')
def reduce_window_c(x): return lax.reduce_window(x, -jnp.inf, lax.max, (1, 1, 2, 2), window_strides=(1, 1, 1, 2), padding='VALID')

param_matrix = [
    (reduce_window_max0, '2x2 no-pad', [R([11, 100, 111, 5])], 0),
    (reduce_window_max1, '3x3 +pad', [R([77, 10, 99, 17])], 0),
    (reduce_window_max1, '3x3 +pad, negative inputs (Q3)', [R([7, 10, 99, 17]) - 2.0], 0),
    (reduce_window_max1, '3x3 +pad, C%4==0 (vector path, ResNet stem shape)', [R([3, 112, 112, 64]) - 0.5], 0),
    (reduce_window_max0, '2x2 no-pad C%4==0', [R([5, 33, 18, 8])], 0),
    (reduce_window_sum0, 'sum 3x3 s2 SAME', [R([9, 10, 99, 16])], 1e-6),
    (avg_pool_2x2, 'avg 2x2 s2 VALID', [R([9, 10, 98, 17])], 1e-6),
    (reduce_window_min0, 'min 3x3 s2 SAME', [R([9, 10, 99, 17])], 0),
    (reduce_window_c, 'window over W and C', [R([4, 9, 10, 12])], 0),
    # the two-columns-per-thread pooling kernel (pool2d_pair_kernel): odd output widths, stride 1, every window kind
    (reduce_window_max1, '3x3 s2 SAME C%4==0, odd OW', [R([2, 9, 13, 8]) - 0.5], 0),
    (reduce_window_max_s1, '3x3 s1 SAME C%4==0', [R([2, 9, 14, 8]) - 0.5], 0),
    (reduce_window_max_s1, '3x3 s1 SAME C%4==0, odd OW', [R([3, 5, 7, 4]) - 0.5], 0),
    (reduce_window_min0, 'min 3x3 s2 SAME C%4==0, odd OW', [R([3, 11, 21, 12])], 0),
    (avg_pool_2x2, 'avg 2x2 s2 VALID C%4==0', [R([9, 10, 98, 16])], 1e-6),
    (reduce_window_sum0, 'sum 3x3 s2 SAME C%4==0, odd OW', [R([4, 7, 9, 32])], 1e-6),
    (reduce_window_max1, '3x3 s2 SAME with NaN and -inf inputs', [np.where(R([2, 8, 8, 8]) < 0.02, np.nan, np.where(R([2, 8, 8, 8]) < 0.85, -np.inf, R([2, 8, 8, 8])))], 0),
]


@pytest.mark.parametrize('f,desc,args,rtol', param_matrix, ids=[p[1] for p in param_matrix])
def test_reduce_window_matrix(f, desc, args, rtol):
    y, ytrue = check(f, args, rtol, 0 if rtol == 0 else 1e-8)
    if rtol == 0:
        assert np.array_equal(np.asarray(y), np.asarray(ytrue), equal_nan=True)      # max/min are exact: bit-equal
