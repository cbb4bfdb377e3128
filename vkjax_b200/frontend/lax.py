"""`lax`-level primitives with their shape/dtype rules (≙ the subset of jax.lax the reference's
handlers accept, SURVEY.md Appendix A, plus the "next" names of §8f).

Each function only *records* an equation in the current trace (frontend/tracing.py).
Parameter names and conventions follow JAX 0.2.x so that vkjax_b200/ops.py handlers read the
same `equation.params` a real jaxpr would carry.
"""
import builtins as _b
import math
import typing as tp

import numpy as np

from .. import core
from ..core import ConvDimensionNumbers, GatherDimensionNumbers, ScatterDimensionNumbers, ShapedArray
from .tracing import Tracer, bind, abstractify, canonicalize_dtype

_prims: tp.Dict[str, core.Primitive] = {}


def prim(name, multiple_results=False) -> core.Primitive:
    if name not in _prims:
        _prims[name] = core.Primitive(name, multiple_results)
    return _prims[name]


def _bcast_shape(*shapes):
    """Old-lax implicit broadcasting: scalars and size-1 dims (same rank) broadcast."""
    shapes = [s for s in shapes if s != ()]
    if not shapes:
        return ()
    rank = len(shapes[0])
    if any(len(s) != rank for s in shapes):
        raise TypeError(f'Incompatible shapes for broadcasting: {shapes}')
    out = []
    for dims in zip(*shapes):
        ds = set(dims) - {1}
        if len(ds) > 1:
            raise TypeError(f'Incompatible shapes for broadcasting: {shapes}')
        out.append(ds.pop() if ds else 1)
    return tuple(out)


def _check_same_dtype(name, *avals):
    dts = {a.dtype for a in avals}
    if len(dts) != 1:
        raise TypeError(f'lax.{name} requires arguments to have the same dtypes, got {[a.dtype.name for a in avals]}')


# ------------------------------------------------------------------------------ elementwise
def _binary(name, out_dtype=None, allowed='fiub'):
    p = prim(name)

    def fn(x, y):
        ax, ay = abstractify(x), abstractify(y)
        # weak python scalars adopt the other operand's dtype
        if ax.weak_type and not ay.weak_type:
            x = _weak_cast(x, ay.dtype); ax = abstractify(x)
        elif ay.weak_type and not ax.weak_type:
            y = _weak_cast(y, ax.dtype); ay = abstractify(y)
        _check_same_dtype(name, ax, ay)
        if ax.dtype.kind not in allowed:
            raise TypeError(f'lax.{name} does not accept dtype {ax.dtype.name}')
        shape = _bcast_shape(ax.shape, ay.shape)
        dt = out_dtype or ax.dtype
        return bind(p, x, y, out_avals=[ShapedArray(shape, dt)])
    fn.__name__ = name
    return fn


def _weak_cast(x, dtype):
    if isinstance(x, Tracer):
        return x
    return np.asarray(x).astype(dtype)[()]


add = _binary('add')
sub = _binary('sub')
mul = _binary('mul')
div = _binary('div', allowed='fiu')
max = _binary('max')
min = _binary('min')
rem = _binary('rem', allowed='fiu')
pow = _binary('pow', allowed='f')
nextafter = _binary('nextafter', allowed='f')
atan2 = _binary('atan2', allowed='f')
add_any = _binary('add_any')
gt = _binary('gt', np.bool_)
ge = _binary('ge', np.bool_)
lt = _binary('lt', np.bool_)
le = _binary('le', np.bool_)
eq = _binary('eq', np.bool_)
ne = _binary('ne', np.bool_)
bitwise_and = _binary('and', allowed='iub')
bitwise_or = _binary('or', allowed='iub')
bitwise_xor = _binary('xor', allowed='iub')
shift_left = _binary('shift_left', allowed='iu')
shift_right_logical = _binary('shift_right_logical', allowed='iu')
shift_right_arithmetic = _binary('shift_right_arithmetic', allowed='iu')


def _unary(name, allowed='f'):
    p = prim(name)

    def fn(x):
        a = abstractify(x)
        if a.dtype.kind not in allowed:
            raise TypeError(f'lax.{name} does not accept dtype {a.dtype.name}')
        return bind(p, x, out_avals=[ShapedArray(a.shape, a.dtype)])
    fn.__name__ = name
    return fn


exp = _unary('exp'); log = _unary('log'); rsqrt = _unary('rsqrt'); sqrt = _unary('sqrt')
erf = _unary('erf'); erf_inv = _unary('erf_inv'); erfc = _unary('erfc')
neg = _unary('neg', 'fiu'); abs = _unary('abs', 'fiu'); sign = _unary('sign', 'fiu')
cos = _unary('cos'); sin = _unary('sin'); tan = _unary('tan')
cosh = _unary('cosh'); sinh = _unary('sinh'); tanh = _unary('tanh')
acos = _unary('acos'); asin = _unary('asin'); atan = _unary('atan')
acosh = _unary('acosh'); asinh = _unary('asinh'); atanh = _unary('atanh')
ceil = _unary('ceil'); floor = _unary('floor'); round = _unary('round')
log1p = _unary('log1p'); expm1 = _unary('expm1'); logistic = _unary('logistic')
bitwise_not = _unary('not', 'iub')
stop_gradient = _unary('stop_gradient', 'fiub')


def integer_pow(x, y: int):
    a = abstractify(x)
    return bind(prim('integer_pow'), x, out_avals=[ShapedArray(a.shape, a.dtype)], y=int(y))


def convert_element_type(x, new_dtype):
    a = abstractify(x)
    new_dtype = canonicalize_dtype(new_dtype)
    if a.dtype == new_dtype and isinstance(x, Tracer):
        return x
    return bind(prim('convert_element_type'), x, out_avals=[ShapedArray(a.shape, new_dtype)],
                new_dtype=new_dtype, weak_type=False)


def bitcast_convert_type(x, new_dtype):
    a = abstractify(x)
    new_dtype = canonicalize_dtype(new_dtype)
    return bind(prim('bitcast_convert_type'), x, out_avals=[ShapedArray(a.shape, new_dtype)], new_dtype=new_dtype)


def select(pred, on_true, on_false):
    ap, a, b = abstractify(pred), abstractify(on_true), abstractify(on_false)
    if ap.dtype != np.bool_:
        raise TypeError('select predicate must be boolean')
    _check_same_dtype('select', a, b)
    if not (ap.shape == a.shape == b.shape):
        raise TypeError(f'select requires equal shapes, got {ap.shape} {a.shape} {b.shape}')
    return bind(prim('select'), pred, on_true, on_false, out_avals=[ShapedArray(a.shape, a.dtype)])


# ------------------------------------------------------------------------------ structural
def broadcast_in_dim(x, shape, broadcast_dimensions):
    a = abstractify(x)
    shape = tuple(int(s) for s in shape)
    bd = tuple(int(d) for d in broadcast_dimensions)
    if len(bd) != len(a.shape):
        raise TypeError('broadcast_dimensions must have one entry per operand dimension')
    for i, d in enumerate(bd):
        if a.shape[i] not in (1, shape[d]):
            raise TypeError(f'broadcast_in_dim: operand dim {i} ({a.shape[i]}) incompatible with target {shape[d]}')
    return bind(prim('broadcast_in_dim'), x, out_avals=[ShapedArray(shape, a.dtype)],
                broadcast_dimensions=bd, shape=shape)


def broadcast(x, sizes):
    a = abstractify(x)
    sizes = tuple(sizes)
    return broadcast_in_dim(x, sizes + a.shape, tuple(range(len(sizes), len(sizes) + len(a.shape))))


def reshape(x, new_sizes, dimensions=None):
    a = abstractify(x)
    new_sizes = tuple(int(s) for s in new_sizes)
    if int(np.prod(new_sizes, dtype=np.int64)) != a.size:
        raise TypeError(f'reshape: cannot reshape {a.shape} to {new_sizes}')
    return bind(prim('reshape'), x, out_avals=[ShapedArray(new_sizes, a.dtype)],
                dimensions=dimensions, new_sizes=new_sizes)


def squeeze(x, dimensions):
    a = abstractify(x)
    dims = tuple(d % len(a.shape) for d in dimensions)
    assert all(a.shape[d] == 1 for d in dims)
    shape = tuple(s for i, s in enumerate(a.shape) if i not in dims)
    return bind(prim('squeeze'), x, out_avals=[ShapedArray(shape, a.dtype)], dimensions=dims)


def transpose(x, permutation):
    a = abstractify(x)
    perm = tuple(int(p) for p in permutation)
    assert sorted(perm) == list(range(len(a.shape)))
    return bind(prim('transpose'), x, out_avals=[ShapedArray(tuple(a.shape[p] for p in perm), a.dtype)],
                permutation=perm)


def rev(x, dimensions):
    a = abstractify(x)
    return bind(prim('rev'), x, out_avals=[ShapedArray(a.shape, a.dtype)], dimensions=tuple(dimensions))


def slice(x, start_indices, limit_indices, strides=None):
    a = abstractify(x)
    start = tuple(int(s) for s in start_indices)
    limit = tuple(int(s) for s in limit_indices)
    st = tuple(int(s) for s in strides) if strides is not None else None
    shape = tuple(-(-(l - s) // (st[i] if st else 1)) for i, (s, l) in enumerate(zip(start, limit)))
    assert all(0 <= s <= l <= d for s, l, d in zip(start, limit, a.shape)), (start, limit, a.shape)
    return bind(prim('slice'), x, out_avals=[ShapedArray(shape, a.dtype)],
                start_indices=start, limit_indices=limit, strides=st)


def concatenate(operands, dimension):
    avals = [abstractify(o) for o in operands]
    _check_same_dtype('concatenate', *avals)
    d = dimension % len(avals[0].shape)
    shape = list(avals[0].shape)
    shape[d] = sum(a.shape[d] for a in avals)
    return bind(prim('concatenate'), *operands, out_avals=[ShapedArray(shape, avals[0].dtype)], dimension=d)


def iota(dtype, size, dimension=0, shape=None):
    shape = (int(size),) if shape is None else tuple(shape)
    dtype = canonicalize_dtype(dtype)
    return bind(prim('iota'), out_avals=[ShapedArray(shape, dtype)], dtype=dtype, shape=shape, dimension=dimension)


# ------------------------------------------------------------------------------ reductions
def _reduce(name):
    p = prim(name)

    def fn(x, axes):
        a = abstractify(x)
        axes = tuple(sorted(int(ax) % _b.max(len(a.shape), 1) for ax in axes))
        shape = tuple(s for i, s in enumerate(a.shape) if i not in axes)
        return bind(p, x, out_avals=[ShapedArray(shape, a.dtype)], axes=axes)
    fn.__name__ = name
    return fn


reduce_sum = _reduce('reduce_sum'); reduce_max = _reduce('reduce_max')
reduce_min = _reduce('reduce_min'); reduce_prod = _reduce('reduce_prod')


def _argreduce(name):
    p = prim(name)

    def fn(x, axis, index_dtype=np.int32):
        a = abstractify(x)
        axis = int(axis) % len(a.shape)
        shape = tuple(s for i, s in enumerate(a.shape) if i != axis)
        return bind(p, x, out_avals=[ShapedArray(shape, canonicalize_dtype(index_dtype))],
                    axes=(axis,), index_dtype=canonicalize_dtype(index_dtype))
    fn.__name__ = name
    return fn


argmax = _argreduce('argmax'); argmin = _argreduce('argmin')


def padtype_to_pads(in_shape, window_shape, window_strides, padding: str):
    """'SAME'/'VALID' → explicit (lo, hi) pairs, XLA convention."""
    if padding.upper() == 'VALID':
        return [(0, 0)] * len(in_shape)
    if padding.upper() == 'SAME':
        pads = []
        for i, k, s in zip(in_shape, window_shape, window_strides):
            out = -(-i // s)
            total = builtins_max((out - 1) * s + k - i, 0)
            pads.append((total // 2, total - total // 2))
        return pads
    raise ValueError(padding)


builtins_max = _b.max
builtins_min = _b.min


def _window_out_shape(in_shape, window, strides, padding, base_dilation=None, window_dilation=None):
    out = []
    for i, (n, k, s, (lo, hi)) in enumerate(zip(in_shape, window, strides, padding)):
        bd = base_dilation[i] if base_dilation else 1
        wd = window_dilation[i] if window_dilation else 1
        n_d = 0 if n == 0 else (n - 1) * bd + 1
        k_d = 0 if k == 0 else (k - 1) * wd + 1
        out.append(builtins_max((n_d + lo + hi - k_d) // s + 1, 0))
    return tuple(out)


def _reduce_window(name, x, window_dimensions, window_strides, padding):
    a = abstractify(x)
    wd = tuple(int(w) for w in window_dimensions)
    ws = tuple(int(s) for s in window_strides)
    if isinstance(padding, str):
        padding = padtype_to_pads(a.shape, wd, ws, padding)
    padding = tuple((int(lo), int(hi)) for lo, hi in padding)
    ones = (1,) * len(a.shape)
    shape = _window_out_shape(a.shape, wd, ws, padding)
    return bind(prim(name), x, out_avals=[ShapedArray(shape, a.dtype)], window_dimensions=wd,
                window_strides=ws, padding=padding, base_dilation=ones, window_dilation=ones)


def reduce_window(x, init_value, computation, window_dimensions, window_strides, padding):
    """≙ jax.lax.reduce_window for the monoids JAX special-cases (reference tests/test_reduce_window.py:14-15)."""
    init = float(np.asarray(init_value))
    if computation is max and init == -math.inf:
        return _reduce_window('reduce_window_max', x, window_dimensions, window_strides, padding)
    if computation is min and init == math.inf:
        return _reduce_window('reduce_window_min', x, window_dimensions, window_strides, padding)
    if computation is add and init == 0:
        return _reduce_window('reduce_window_sum', x, window_dimensions, window_strides, padding)
    raise NotImplementedError('reduce_window with a general computation')


def select_and_scatter_add(source, operand, select_prim, window_dimensions, window_strides, padding):
    """≙ jax.lax's select_and_scatter_add_p (what jax.grad of a max / min pool emits): `select_prim` is lax.ge_p / lax.le_p
    there; here the primitive object, its name, or the lax.ge / lax.le function."""
    a, sa = abstractify(operand), abstractify(source)
    _check_same_dtype('select_and_scatter_add', a, sa)
    name = getattr(select_prim, 'name', getattr(select_prim, '__name__', select_prim))
    wd = tuple(int(w) for w in window_dimensions)
    ws = tuple(int(s) for s in window_strides)
    if isinstance(padding, str):
        padding = padtype_to_pads(a.shape, wd, ws, padding)
    padding = tuple((int(lo), int(hi)) for lo, hi in padding)
    assert _window_out_shape(a.shape, wd, ws, padding) == tuple(sa.shape), (a.shape, sa.shape)
    return bind(prim('select_and_scatter_add'), source, operand, out_avals=[ShapedArray(a.shape, a.dtype)],
                select_prim=prim(name), window_dimensions=wd, window_strides=ws, padding=padding)


# ------------------------------------------------------------------------------ contractions
def dot_general(lhs, rhs, dimension_numbers, precision=None):
    a, b = abstractify(lhs), abstractify(rhs)
    _check_same_dtype('dot_general', a, b)
    (lc, rc), (lb, rb) = dimension_numbers
    lc, rc, lb, rb = tuple(lc), tuple(rc), tuple(lb), tuple(rb)
    assert [a.shape[i] for i in lc] == [b.shape[i] for i in rc], (a.shape, b.shape, dimension_numbers)
    assert [a.shape[i] for i in lb] == [b.shape[i] for i in rb]
    batch = tuple(a.shape[i] for i in lb)
    lfree = tuple(s for i, s in enumerate(a.shape) if i not in lc + lb)
    rfree = tuple(s for i, s in enumerate(b.shape) if i not in rc + rb)
    return bind(prim('dot_general'), lhs, rhs, out_avals=[ShapedArray(batch + lfree + rfree, a.dtype)],
                dimension_numbers=((lc, rc), (lb, rb)), precision=precision)


def conv_dimension_numbers(lhs_shape, rhs_shape, dimension_numbers) -> ConvDimensionNumbers:
    if dimension_numbers is None:
        n = len(lhs_shape)
        iota_ = tuple(range(n))
        return ConvDimensionNumbers(iota_, iota_, iota_)
    if isinstance(dimension_numbers, ConvDimensionNumbers):
        return dimension_numbers
    lhs_s, rhs_s, out_s = dimension_numbers   # strings like ('NHWC','HWIO','NHWC')
    def spec(s, order):
        spatial = [c for c in s if c not in order]
        # spatial dims are ordered as they appear in the rhs spec
        return tuple(s.index(c) for c in order) , spatial
    rhs_sp = [c for c in rhs_s if c not in 'OI']
    lhs_spec = (lhs_s.index('N'), lhs_s.index('C')) + tuple(lhs_s.index(c) for c in rhs_sp)
    rhs_spec = (rhs_s.index('O'), rhs_s.index('I')) + tuple(rhs_s.index(c) for c in rhs_sp)
    out_spec = (out_s.index('N'), out_s.index('C')) + tuple(out_s.index(c) for c in rhs_sp)
    return ConvDimensionNumbers(lhs_spec, rhs_spec, out_spec)


def conv_general_dilated(lhs, rhs, window_strides, padding, lhs_dilation=None, rhs_dilation=None,
                         dimension_numbers=None, feature_group_count=1, batch_group_count=1, precision=None):
    a, b = abstractify(lhs), abstractify(rhs)
    _check_same_dtype('conv_general_dilated', a, b)
    dn = conv_dimension_numbers(a.shape, b.shape, dimension_numbers)
    nsp = len(a.shape) - 2
    strides = tuple(int(s) for s in window_strides)
    lhs_dil = tuple(lhs_dilation) if lhs_dilation is not None else (1,) * nsp
    rhs_dil = tuple(rhs_dilation) if rhs_dilation is not None else (1,) * nsp
    lhs_sp = [a.shape[i] for i in dn.lhs_spec[2:]]
    rhs_sp = [b.shape[i] for i in dn.rhs_spec[2:]]
    if isinstance(padding, str):
        eff_k = [(k - 1) * d + 1 for k, d in zip(rhs_sp, rhs_dil)]
        eff_in = [(n - 1) * d + 1 for n, d in zip(lhs_sp, lhs_dil)]
        padding = padtype_to_pads(eff_in, eff_k, strides, padding)
    padding = tuple((int(lo), int(hi)) for lo, hi in padding)
    out_sp = _window_out_shape(lhs_sp, rhs_sp, strides, padding, lhs_dil, rhs_dil)
    assert a.shape[dn.lhs_spec[1]] == b.shape[dn.rhs_spec[1]] * feature_group_count
    out_shape = [0] * len(a.shape)
    out_shape[dn.out_spec[0]] = a.shape[dn.lhs_spec[0]]
    out_shape[dn.out_spec[1]] = b.shape[dn.rhs_spec[0]]
    for i, d in enumerate(dn.out_spec[2:]):
        out_shape[d] = out_sp[i]
    return bind(prim('conv_general_dilated'), lhs, rhs, out_avals=[ShapedArray(out_shape, a.dtype)],
                window_strides=strides, padding=padding, lhs_dilation=lhs_dil, rhs_dilation=rhs_dil,
                dimension_numbers=dn, feature_group_count=feature_group_count,
                batch_group_count=batch_group_count, lhs_shape=a.shape, rhs_shape=b.shape, precision=precision)


# ------------------------------------------------------------------------------ indexing
def gather(operand, start_indices, dimension_numbers: GatherDimensionNumbers, slice_sizes,
           indices_are_sorted=False, unique_indices=False):
    a, idx = abstractify(operand), abstractify(start_indices)
    if idx.dtype.kind not in 'iu':
        raise TypeError('gather indices must be integers')
    dn = dimension_numbers
    slice_sizes = tuple(int(s) for s in slice_sizes)
    offset_sizes = [s for i, s in enumerate(slice_sizes) if i not in dn.collapsed_slice_dims]
    batch = list(idx.shape[:-1])
    rank = len(dn.offset_dims) + len(batch)
    out, oi, bi = [], 0, 0
    for d in range(rank):
        if d in dn.offset_dims:
            out.append(offset_sizes[oi]); oi += 1
        else:
            out.append(batch[bi]); bi += 1
    return bind(prim('gather'), operand, start_indices, out_avals=[ShapedArray(out, a.dtype)],
                dimension_numbers=dn, slice_sizes=slice_sizes,
                indices_are_sorted=indices_are_sorted, unique_indices=unique_indices)


def scatter_add(operand, scatter_indices, updates, dimension_numbers: ScatterDimensionNumbers,
                indices_are_sorted=False, unique_indices=False):
    a = abstractify(operand)
    return bind(prim('scatter-add'), operand, scatter_indices, updates, out_avals=[ShapedArray(a.shape, a.dtype)],
                dimension_numbers=dimension_numbers, indices_are_sorted=indices_are_sorted,
                unique_indices=unique_indices, update_consts=(), update_jaxpr=None)


# ------------------------------------------------------------------------------ PRNG
def threefry2x32(key0, key1, data0, data1):
    """≙ jax.random.threefry2x32_p.bind (reference tests/test_basic_ops.py:141-143)."""
    avals = [abstractify(v) for v in (key0, key1, data0, data1)]
    if any(a.dtype != np.uint32 for a in avals):
        raise TypeError('threefry2x32 operands must be uint32')
    shape = avals[2].shape
    return bind(prim('threefry2x32', multiple_results=True), key0, key1, data0, data1,
                out_avals=[ShapedArray(shape, np.uint32), ShapedArray(shape, np.uint32)])
