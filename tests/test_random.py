"""Restatement of reference tests/test_random.py (:7-60): jax.random.normal over 51 keys + shape [2,34],
randint exact equality -- the whole threefry -> bits -> erf_inv chain on the device.  Also pinned against
JAX's own documented values for PRNGKey(0)/PRNGKey(42) (golden vectors, tests/golden/jax_random.json)."""
import json
import os

import numpy as np
import pytest

import vkjax_b200 as vkjax
from vkjax_b200.frontend import random
from common import check, oracle

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'jax_random.json')))
rng = np.random.RandomState(3)


def test_normal0():
    vkfunc = vkjax.wrap(random.normal)
    for i in [0] + list(rng.randint(10000, size=50)):
        key = random.PRNGKey(i)
        ypred = vkfunc(key)
        ytrue, _ = oracle(random.normal, [key])
        assert np.allclose(ytrue, ypred, rtol=1e-5, atol=1e-6)


def test_normal1():
    key = random.PRNGKey(int(rng.randint(10000)))
    fn = lambda k: random.normal(k, shape=[2, 34])
    check(fn, [key], 1e-5, 1e-6)


def test_randint0():
    fn = lambda k: random.randint(k, shape=[1], minval=0, maxval=9999)
    vkfunc = vkjax.wrap(fn)
    for i in [0] + list(rng.randint(10000, size=50)):
        key = random.PRNGKey(i)
        ypred = vkfunc(key)
        ytrue, _ = oracle(fn, [key])
        assert np.all(ypred == ytrue)


def test_golden_jax_values_on_device():
    g = GOLDEN
    key0 = random.PRNGKey(0)
    assert np.array_equal(vkjax.wrap(lambda k: random.split(k))(key0), np.array(g['split_key0'], np.uint32))
    assert np.allclose(vkjax.wrap(lambda k: random.uniform(k))(key0), g['uniform_key0'], rtol=1e-6, atol=0)
    assert np.allclose(vkjax.wrap(lambda k: random.normal(k))(key0), g['normal_key0'], rtol=2e-6, atol=0)
    y = vkjax.wrap(lambda k: random.normal(k, (3,)))(random.PRNGKey(42))
    assert np.allclose(y, g['normal_key42_shape3'], rtol=2e-6, atol=0)
