"""Threefry-based PRNG chains in the JAX-0.2.x formulation (what `jax.random.normal/uniform/
randint/split` traced to when the reference's tests/test_random.py was written).  Every step is
a recorded primitive (threefry2x32, shifts, or, bitcast, erf_inv, rem, select ...), so the whole
chain runs on the GPU through the handlers in vkjax_b200/ops.py.
"""
import numpy as np

from . import lax, jnp
from .tracing import abstractify


def PRNGKey(seed: int) -> np.ndarray:
    """Host-side: keys are plain uint32[2] arrays, [hi, lo] of the 64-bit seed."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def _unpack(key):
    k0 = lax.reshape(lax.slice(key, (0,), (1,)), ())
    k1 = lax.reshape(lax.slice(key, (1,), (2,)), ())
    return k0, k1


def threefry_2x32(key, count):
    """Hash `count` (any shape, uint32) with `key` (uint32[2])."""
    k0, k1 = _unpack(key)
    a = abstractify(count)
    flat = lax.reshape(count, (a.size,))
    odd = a.size % 2
    if odd:
        flat = lax.concatenate([flat, np.zeros((1,), np.uint32)], 0)
    n = a.size + odd
    x0 = lax.slice(flat, (0,), (n // 2,))
    x1 = lax.slice(flat, (n // 2,), (n,))
    y0, y1 = lax.threefry2x32(k0, k1, x0, x1)
    out = lax.concatenate([y0, y1], 0)
    if odd:
        out = lax.slice(out, (0,), (n - 1,))
    return lax.reshape(out, a.shape)


def split(key, num: int = 2):
    counts = lax.iota(np.uint32, num * 2)
    return lax.reshape(threefry_2x32(key, counts), (num, 2))


def fold_in(key, data: int):
    return threefry_2x32(key, PRNGKey(data))


def _random_bits(key, shape):
    size = int(np.prod(shape, dtype=np.int64))
    bits = threefry_2x32(key, lax.iota(np.uint32, size))
    return lax.reshape(bits, tuple(shape))


def uniform(key, shape=(), dtype=np.float32, minval=0.0, maxval=1.0):
    shape = tuple(shape)
    minval, maxval = np.float32(minval), np.float32(maxval)
    bits = _random_bits(key, shape)
    float_bits = lax.bitwise_or(lax.shift_right_logical(bits, np.uint32(32 - 23)),
                                np.float32(1.0).view(np.uint32))
    floats = lax.sub(lax.bitcast_convert_type(float_bits, np.float32), np.float32(1.0))
    scaled = lax.add(lax.mul(floats, np.float32(maxval - minval)), minval)
    return lax.max(minval, scaled)


def normal(key, shape=(), dtype=np.float32):
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0), dtype=np.float32)
    u = uniform(key, shape, np.float32, lo, 1.0)
    return lax.mul(np.float32(np.sqrt(2)), lax.erf_inv(u))


def truncated_normal(key, lower, upper, shape=(), dtype=np.float32):
    from math import erf, sqrt
    a = np.float32(erf(lower / sqrt(2)))
    b = np.float32(erf(upper / sqrt(2)))
    u = uniform(key, shape, np.float32, a, b)
    return lax.mul(np.float32(np.sqrt(2)), lax.erf_inv(u))


def randint(key, shape, minval: int, maxval: int, dtype=np.int32):
    shape = tuple(shape)
    keys = split(key)
    k1 = lax.reshape(lax.slice(keys, (0, 0), (1, 2)), (2,))
    k2 = lax.reshape(lax.slice(keys, (1, 0), (2, 2)), (2,))
    higher_bits, lower_bits = _random_bits(k1, shape), _random_bits(k2, shape)
    span = np.uint32(maxval - minval) if maxval > minval else np.uint32(1)
    multiplier = lax.rem(np.uint32(2 ** 16), span)
    multiplier = lax.rem(lax.mul(multiplier, multiplier), span)
    random_offset = lax.add(lax.mul(lax.rem(higher_bits, span), multiplier), lax.rem(lower_bits, span))
    random_offset = lax.rem(random_offset, span)
    return lax.add(np.int32(minval), lax.convert_element_type(random_offset, np.int32))
