#!/usr/bin/env python
"""Generator of the committed golden fixtures under tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

Why these exist ("parity unpinned"): the numpy oracle (oracle/eval_jaxpr.py) and the CUDA path both
consume jaxprs from the repo's own tracer, so an error shared by tracer + oracle (SAME-padding arithmetic, dimension-spec
handling, ...) would be invisible to a purely differential test.  The fixtures below are produced WITHOUT the tracer,
WITHOUT the oracle and WITHOUT numpy arithmetic:

  contractions_int.npz   integer-exact known answers for
      * the 13 conv_general_dilated cases of reference tests/test_conv.py:66-85 (same channel counts, filter sizes,
        strides, paddings, dilations and dimension specs; batch / spatial extents shrunk so that pure-Python loops
        finish), computed by `conv2d_comp`, a plain-Python restatement of reference vkjax/shaders/conv2d.comp:44-95
        (same loop nest, same is_valid logic) with HAND-WRITTEN padding tuples and output shapes (no SAME/VALID helper);
      * both reduce_window_max cases of reference tests/test_reduce_window.py:14-21, by `reduce_window_max_comp`
        (reduce_window_max_2d.comp:29-52) -- inputs are non-negative as in the reference test, where the shader's
        "padding contributes 0.0" (quirk Q3) and lax's "-inf" agree; the generator asserts that they do;
      * the 6 dot_general variants of reference tests/test_basic_ops.py:197-203, by `dot_general_comp`
        (dot_general.comp:10-34).
    Inputs are small integers drawn from Python's `random.Random` (so every product and partial sum is exactly
    representable in fp32 AND in TF32: any correct implementation, whatever its summation order or tensor-core
    precision mode, must reproduce the answers BIT FOR BIT).

  jax_random.json        JAX's documented PRNG values (written by hand from the JAX documentation / jax.random
      docstrings, see `JAX_RANDOM_DOC`); regenerated verbatim by this script so that the file has a committed source.

numpy is used only to serialise (np.savez_compressed); no array arithmetic happens here.
"""
import json
import os
import random

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


# =================================================================================================
# reference vkjax/shaders/common.glsl:4-29
def unravel_index(index, shape):
    coords = [0] * len(shape)
    for i in range(len(shape) - 1, -1, -1):
        coords[i] = index % shape[i]
        index //= shape[i]
    return coords


def ravel_coords(coords, shape):
    index, stride = 0, 1
    for i in range(len(shape) - 1, -1, -1):
        index += coords[i] * stride
        stride *= shape[i]
    return index


def is_out_of_bounds(coords, shape):
    return any(c < 0 or c >= s for c, s in zip(coords, shape))


def size_of(shape):
    n = 1
    for s in shape:
        n *= s
    return n


# =================================================================================================
def conv2d_comp(in_a, in_b, shape_a, shape_b, shape_out, spec_lhs, spec_rhs, spec_out, padding, strides, dilate_lhs, dilate_rhs):
    """reference vkjax/shaders/conv2d.comp:44-95, one "invocation" per output element, flat row-major buffers."""
    d_shape_a, d_shape_b = list(shape_a), list(shape_b)
    d_shape_a[spec_lhs[2]] *= dilate_lhs[0]
    d_shape_a[spec_lhs[3]] *= dilate_lhs[1]
    d_shape_b[spec_rhs[2]] *= dilate_rhs[0]
    d_shape_b[spec_rhs[3]] *= dilate_rhs[1]
    result = [0] * size_of(shape_out)
    for index in range(len(result)):
        coords_out = unravel_index(index, shape_out)
        total = 0
        for i0 in range(0, d_shape_b[spec_rhs[2]], dilate_rhs[0]):                       # spatial dimension 0
            j0 = coords_out[spec_out[2]] * strides[0] + i0 - padding[0]
            for i1 in range(0, d_shape_b[spec_rhs[3]], dilate_rhs[1]):                   # spatial dimension 1
                j1 = coords_out[spec_out[3]] * strides[1] + i1 - padding[1]
                for c in range(shape_b[spec_rhs[1]]):                                    # in feature dimension
                    coords_a, coords_b = [0] * 4, [0] * 4
                    coords_a[spec_lhs[0]] = coords_out[spec_out[0]]
                    coords_a[spec_lhs[1]] = c
                    coords_a[spec_lhs[2]] = j0
                    coords_a[spec_lhs[3]] = j1
                    coords_b[spec_rhs[0]] = coords_out[spec_out[1]]
                    coords_b[spec_rhs[1]] = c
                    coords_b[spec_rhs[2]] = i0
                    coords_b[spec_rhs[3]] = i1
                    if is_out_of_bounds(coords_a, d_shape_a) or is_out_of_bounds(coords_b, d_shape_b):
                        continue
                    if coords_a[spec_lhs[2]] % dilate_lhs[0] > 0 or coords_a[spec_lhs[3]] % dilate_lhs[1] > 0:
                        continue                                                          # falls between dilated lhs samples
                    if coords_b[spec_rhs[2]] % dilate_rhs[0] > 0 or coords_b[spec_rhs[3]] % dilate_rhs[1] > 0:
                        continue
                    coords_a[spec_lhs[2]] //= dilate_lhs[0]
                    coords_a[spec_lhs[3]] //= dilate_lhs[1]
                    coords_b[spec_rhs[2]] //= dilate_rhs[0]
                    coords_b[spec_rhs[3]] //= dilate_rhs[1]
                    total += in_a[ravel_coords(coords_a, shape_a)] * in_b[ravel_coords(coords_b, shape_b)]
        result[index] = total
    return result


def dot_general_comp(in_a, in_b, n, c, m, cdim_a, cdim_b):
    """reference vkjax/shaders/dot_general.comp:10-34: out[N, M]."""
    result = [0] * (n * m)
    stride_a = 1 * cdim_a + n * (1 - cdim_a)
    stride_b = 1 * cdim_b + m * (1 - cdim_b)
    for index in range(n * m):
        row, col = index // m, index % m
        offset_a = row * c * cdim_a + row * (1 - cdim_a)
        offset_b = col * c * cdim_b + col * (1 - cdim_b)
        total = 0
        for i in range(c):
            total += in_a[offset_a + i * stride_a] * in_b[offset_b + i * stride_b]
        result[index] = total
    return result


def reduce_window_max_comp(in_a, shape_a, shape_out, padding, strides, window, pad_value):
    """reference vkjax/shaders/reduce_window_max_2d.comp:29-52.  pad_value = 0 is the shader (`in_a[0] * 0.0`),
    pad_value = None means lax semantics (padding is the identity, -inf)."""
    result = [0] * size_of(shape_out)
    for index in range(len(result)):
        coords_out = unravel_index(index, shape_out)
        acc = None                                                                       # -inf
        for i in range(size_of(window)):
            cw = unravel_index(i, window)
            coords_a = [coords_out[d] * strides[d] + cw[d] - padding[d] for d in range(4)]
            if is_out_of_bounds(coords_a, shape_a):
                v = pad_value
            else:
                v = in_a[ravel_coords(coords_a, shape_a)]
            if v is not None and (acc is None or v > acc):
                acc = v
        result[index] = acc
    return result


# =================================================================================================
NHWC = ((0, 3, 1, 2), (3, 2, 0, 1), (0, 3, 1, 2))          # lhs_spec, rhs_spec, out_spec (reference tests/test_conv.py:19)
NCHW = ((0, 1, 2, 3), (0, 1, 2, 3), (0, 1, 2, 3))          # reference tests/test_conv.py:25-28

# name (≙ the description in reference tests/test_conv.py:66-84), function name in that file, x shape, kernel shape,
# out shape, specs, low padding, strides, lhs dilation, rhs dilation, the `padding` argument the test passes to lax.
# Output shapes and low paddings are written out by hand:
#   VALID: out = (in - (k-1)*rhs_dil - 1) // stride + 1, pad 0
#   SAME:  out = ceil(in / stride), total = max((out-1)*stride + (k-1)*rhs_dil + 1 - in, 0), low = total // 2
#   explicit [(lo, hi), ...] with lhs dilation: in' = (in-1)*lhs_dil + 1, out = (in' + lo + hi - k) // stride + 1
CONV_CASES = [
    ('conv0 1x1 const kernel no pad', 'conv0', (2, 12, 11, 5), (1, 1, 5, 33), (2, 12, 11, 33), NHWC, (0, 0), (1, 1), (1, 1), (1, 1), 'VALID'),
    ('conv1 1x1 var kernel no pad', 'conv1', (2, 9, 10, 33), (1, 1, 33, 11), (2, 9, 10, 11), NHWC, (0, 0), (1, 1), (1, 1), (1, 1), 'VALID'),
    ('conv1 3x3 var kernel no pad', 'conv1', (3, 13, 11, 5), (3, 3, 5, 7), (3, 11, 9, 7), NHWC, (0, 0), (1, 1), (1, 1), (1, 1), 'VALID'),
    ('conv1a 3x3 convdims 0123', 'conv1a', (2, 8, 9, 8), (39, 8, 3, 3), (2, 39, 7, 6), NCHW, (0, 0), (1, 1), (1, 1), (1, 1), 'VALID'),
    ('conv2 1x1 var kernel +pad', 'conv2', (4, 17, 9, 12), (1, 1, 12, 11), (4, 17, 9, 11), NHWC, (0, 0), (1, 1), (1, 1), (1, 1), 'SAME'),
    ('conv2 3x3 var kernel +pad', 'conv2', (2, 9, 7, 7), (3, 3, 7, 38), (2, 9, 7, 38), NHWC, (1, 1), (1, 1), (1, 1), (1, 1), 'SAME'),
    ('conv2 7x7 var kernel +pad', 'conv2', (2, 12, 19, 3), (7, 7, 3, 4), (2, 12, 19, 4), NHWC, (3, 3), (1, 1), (1, 1), (1, 1), 'SAME'),
    ('conv3 3x3 uneven pad', 'conv3', (1, 9, 8, 11), (3, 3, 11, 38), (1, 9, 9, 38), NHWC, (2, 0), (1, 1), (1, 1), (1, 1), [(2, 0), (0, 3)]),
    ('conv4 1x1 window strides=2', 'conv4', (2, 67, 42, 3), (1, 1, 3, 2), (2, 34, 21, 2), NHWC, (0, 0), (2, 2), (1, 1), (1, 1), 'SAME'),
    ('conv4 3x3 window strides=2', 'conv4', (2, 17, 12, 11), (3, 3, 11, 7), (2, 9, 6, 7), NHWC, (1, 0), (2, 2), (1, 1), (1, 1), 'SAME'),
    ('conv5 3x3 window strides=2 + rhs_dilate=2', 'conv5', (2, 17, 12, 11), (3, 3, 11, 7), (2, 7, 4, 7), NHWC, (0, 0), (2, 2), (1, 1), (2, 2), 'VALID'),
    ('conv6a 3x3 window strides=2 + lhs_dilate=2', 'conv6a', (2, 9, 7, 11), (3, 3, 11, 7), (2, 19, 17, 7), NHWC, (2, 3), (1, 1), (2, 2), (1, 1), [(2, 2), (3, 3)]),
    ('conv6b 3x3 window strides=2 + lhs_dilate=2', 'conv6b', (2, 9, 7, 11), (3, 3, 11, 7), (2, 8, 6, 7), NHWC, (0, 0), (2, 2), (2, 2), (1, 1), [(0, 0), (0, 0)]),
]

# reference tests/test_reduce_window.py:14-21 (batch / spatial extents shrunk; windows, strides, padding kind kept)
POOL_CASES = [
    ('2x2 no-pad', 'reduce_window_max0', (2, 10, 11, 5), (2, 9, 10, 5), (0, 0, 0, 0), (1, 1, 1, 1), (1, 2, 2, 1), 'VALID'),
    # SAME, 3x3 stride 2: H 10 -> 5, total pad (5-1)*2+3-10 = 1 -> low 0; W 9 -> 5, total (5-1)*2+3-9 = 2 -> low 1
    ('3x3 +pad', 'reduce_window_max1', (3, 10, 9, 17), (3, 5, 5, 17), (0, 0, 1, 0), (1, 2, 2, 1), (1, 3, 3, 1), 'SAME'),
]

# reference tests/test_basic_ops.py:197-203: (description, a shape, b shape, contracting dim of a, of b)
DOT_CASES = [
    ('dot_general(2x100 @ 100x32)', (2, 100), (100, 32), 1, 0),
    ('dot_general const rhs', (2, 100), (100, 32), 1, 0),
    ('dot_general contraction (0,0)', (100, 2), (100, 32), 0, 0),
    ('dot_general contraction (1,1)', (2, 100), (32, 100), 1, 1),
    ('dot_general contraction (0,1)', (100, 2), (32, 100), 0, 1),
    ('reshape(-1,4) @ 4x4', (2 * 77 * 102 // 4, 4), (4, 4), 1, 0),
]


def ints(rng, n, lo, hi):
    return [rng.randint(lo, hi) for _ in range(n)]


def make_contractions(path):
    rng = random.Random(20261017)
    arrays, meta = {}, {'conv': [], 'pool': [], 'dot': []}
    for i, (name, fn, xs, ks, os_, specs, pad, strides, ldil, rdil, padding_arg) in enumerate(CONV_CASES):
        x = ints(rng, size_of(xs), 0, 3)                   # reference inputs are U[0,1): non-negative
        k = ints(rng, size_of(ks), -2, 2)
        y = conv2d_comp(x, k, xs, ks, os_, specs[0], specs[1], specs[2], pad, strides, ldil, rdil)
        assert max(abs(v) for v in y) < 2 ** 20
        arrays[f'conv{i}_x'] = np.array(x, np.int8).reshape(xs)
        arrays[f'conv{i}_k'] = np.array(k, np.int8).reshape(ks)
        arrays[f'conv{i}_y'] = np.array(y, np.int32).reshape(os_)
        meta['conv'].append(dict(name=name, fn=fn, lhs_spec=specs[0], rhs_spec=specs[1], out_spec=specs[2], pad_lo=pad, strides=strides,
                                 lhs_dilation=ldil, rhs_dilation=rdil, padding=padding_arg))
        print(f'conv {i:2d} {name}: x{xs} k{ks} -> y{os_}')
    for i, (name, fn, xs, os_, pad, strides, window, padding_arg) in enumerate(POOL_CASES):
        x = ints(rng, size_of(xs), 0, 1000)
        y = reduce_window_max_comp(x, xs, os_, pad, strides, window, 0)
        y_lax = reduce_window_max_comp(x, xs, os_, pad, strides, window, None)
        assert y == y_lax, 'quirk Q3 must be invisible on non-negative inputs'
        arrays[f'pool{i}_x'] = np.array(x, np.int16).reshape(xs)
        arrays[f'pool{i}_y'] = np.array(y, np.int16).reshape(os_)
        meta['pool'].append(dict(name=name, fn=fn, pad_lo=pad, strides=strides, window=window, padding=padding_arg))
        print(f'pool {i} {name}: x{xs} -> y{os_}')
    # negative-valued max-pool input: lax semantics (padding = -inf), where the reference shader is WRONG (quirk Q3); the
    # fixture pins the lax behaviour that the north star asks for
    name, fn, xs, os_, pad, strides, window, padding_arg = POOL_CASES[1]
    x = ints(rng, size_of(xs), -1000, -1)
    arrays['pool_neg_x'] = np.array(x, np.int16).reshape(xs)
    arrays['pool_neg_y'] = np.array(reduce_window_max_comp(x, xs, os_, pad, strides, window, None), np.int16).reshape(os_)
    for i, (name, sa, sb, ca, cb) in enumerate(DOT_CASES):
        a = ints(rng, size_of(sa), -3, 3)
        b = ints(rng, size_of(sb), -3, 3)
        n, m, c = sa[1 - ca], sb[1 - cb], sa[ca]
        y = dot_general_comp(a, b, n, c, m, ca, cb)
        arrays[f'dot{i}_a'] = np.array(a, np.int8).reshape(sa)
        arrays[f'dot{i}_b'] = np.array(b, np.int8).reshape(sb)
        arrays[f'dot{i}_y'] = np.array(y, np.int32).reshape(n, m)
        meta['dot'].append(dict(name=name, cdim_a=ca, cdim_b=cb))
        print(f'dot  {i} {name}: a{sa} b{sb} -> y({n}, {m})')
    arrays['meta'] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(path, **arrays)
    print('wrote', path, os.path.getsize(path), 'bytes')


# =================================================================================================
# JAX's documented PRNG values + the Random123 Threefry-2x32-20 known-answer vectors.  jax cannot be imported in this
# image, so these are TRANSCRIBED (from jax.random's module documentation / "Pseudo random numbers in JAX" / the README
# quickstart, and from Random123's kat_vectors for threefry2x32_20 -- the kernel reference vkjax/shaders/threefry2x32.comp:1
# cites); this dict is their committed source, written to jax_random.json verbatim.
JAX_RANDOM_DOC = {
    '_source': "Values printed in JAX's own documentation (jax.random module docs / 'Pseudo random numbers in JAX' / README quickstart) "
               "for the default threefry2x32 PRNG: jax.random.split(PRNGKey(0)), jax.random.uniform(PRNGKey(0)), "
               "jax.random.normal(PRNGKey(0)), jax.random.normal(PRNGKey(42), (3,)). jax cannot be imported in this image, so these "
               "are transcribed, not regenerated; tests/golden/make_golden.py (JAX_RANDOM_DOC) is their committed source and "
               "regenerates everything that CAN be generated here.",
    'split_key0': [[4146024105, 967050713], [2718843009, 1272950319]],
    'uniform_key0': 0.41845703,
    'normal_key0': -0.20584226,
    'normal_key42_shape3': [0.18693547, -1.2806505, -1.5593132],
    'threefry2x32_kat': [
        {'key': [0, 0], 'ctr': [0, 0], 'out': [1797259609, 2579123966]},
        {'key': [4294967295, 4294967295], 'ctr': [4294967295, 4294967295], 'out': [481924860, 3137350631]},
        {'key': [320440878, 57701188], 'ctr': [608135816, 2242054355], 'out': [3297917596, 1212020640]},
    ],
}


def main():
    make_contractions(os.path.join(HERE, 'contractions_int.npz'))
    json.dump(JAX_RANDOM_DOC, open(os.path.join(HERE, 'jax_random.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
