"""`jax.numpy`-flavoured helpers over frontend.lax, restricted to what the reference's tests and
model code use (reference tests/test_basic_ops.py:16-157).  dtype/shape promotion follows
JAX-0.2.x conventions: python scalars are weak, rank promotion inserts a size-preserving
`broadcast_in_dim`, size-1 broadcasting is left implicit on the binary primitive.
"""
import builtins
import typing as tp

import numpy as np

from . import lax
from .tracing import Tracer, abstractify, canonicalize_dtype
from ..core import GatherDimensionNumbers

float32 = np.float32
int32 = np.int32
uint32 = np.uint32
bool_ = np.bool_
inf = np.inf
newaxis = None

_RANK = {'b': 0, 'u': 1, 'i': 2, 'f': 3}


def _result_dtype(ax, ay):
    if ax.dtype == ay.dtype:
        return ax.dtype
    if ax.weak_type != ay.weak_type:
        weak, strong = (ax, ay) if ax.weak_type else (ay, ax)
        # a python scalar adopts the array's dtype unless it is of a "higher kind" (float over int over bool)
        order = {'b': 0, 'u': 1, 'i': 1, 'f': 2}
        if order[weak.dtype.kind] <= order[strong.dtype.kind]:
            return strong.dtype
        return canonicalize_dtype(weak.dtype)            # e.g. int array + python float -> float32
    return ax.dtype if _RANK[ax.dtype.kind] >= _RANK[ay.dtype.kind] else ay.dtype


def _cast(x, dtype):
    a = abstractify(x)
    if a.dtype == dtype:
        return x
    if isinstance(x, Tracer):
        return lax.convert_element_type(x, dtype)
    return np.asarray(x).astype(dtype)[()] if np.ndim(x) == 0 else np.asarray(x).astype(dtype)


def _promote(x, y):
    ax, ay = abstractify(x), abstractify(y)
    dt = _result_dtype(ax, ay)
    x, y = _cast(x, dtype=dt), _cast(y, dtype=dt)
    # rank promotion (scalars are left alone: the binary primitives broadcast them)
    if ax.shape and ay.shape and len(ax.shape) != len(ay.shape):
        nd = builtins.max(len(ax.shape), len(ay.shape))
        def up(v, a):
            if len(a.shape) == nd:
                return v
            k = nd - len(a.shape)
            return lax.broadcast_in_dim(v, (1,) * k + a.shape, tuple(range(k, nd)))
        x, y = up(x, ax), up(y, ay)
    return x, y


def _binop(fn):
    def op(x, y):
        x, y = _promote(x, y)
        return fn(x, y)
    return op


add = _binop(lax.add); subtract = _binop(lax.sub); multiply = _binop(lax.mul)
maximum = _binop(lax.max); minimum = _binop(lax.min)
greater = _binop(lax.gt); greater_equal = _binop(lax.ge); less = _binop(lax.lt); less_equal = _binop(lax.le)
equal = _binop(lax.eq); not_equal = _binop(lax.ne)
bitwise_and = _binop(lax.bitwise_and); bitwise_or = _binop(lax.bitwise_or); bitwise_xor = _binop(lax.bitwise_xor)
left_shift = _binop(lax.shift_left)


def true_divide(x, y):
    x, y = _promote(x, y)
    if abstractify(x).dtype.kind != 'f':
        x, y = _cast(x, np.float32), _cast(y, np.float32)
    return lax.div(x, y)


def power(x, y):
    if isinstance(y, int) and not isinstance(y, bool):
        return lax.integer_pow(x, y)
    x, y = _promote(x, y)
    return lax.pow(x, y)


def _float(x):
    return x if abstractify(x).dtype.kind == 'f' else _cast(x, np.float32)


exp = lambda x: lax.exp(_float(x))
log = lambda x: lax.log(_float(x))
sqrt = lambda x: lax.sqrt(_float(x))
tanh = lambda x: lax.tanh(_float(x))
abs = lambda x: lax.abs(x)
negative = lambda x: lax.neg(x)


def asarray(x, dtype=None):
    if dtype is not None:
        return _cast(x, canonicalize_dtype(dtype))
    return x


def reshape(x, *shape):
    if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
        shape = tuple(shape[0])
    a = abstractify(x)
    if -1 in shape:
        known = int(np.prod([s for s in shape if s != -1], dtype=np.int64))
        shape = tuple(a.size // builtins.max(known, 1) if s == -1 else s for s in shape)
    return lax.reshape(x, shape)


def ravel(x):
    return reshape(x, (-1,))


def squeeze(x, axis=None):
    a = abstractify(x)
    dims = tuple(i for i, s in enumerate(a.shape) if s == 1) if axis is None else \
        ((axis,) if isinstance(axis, int) else tuple(axis))
    return lax.squeeze(x, dims)


def transpose(x, axes=None):
    a = abstractify(x)
    axes = tuple(reversed(range(len(a.shape)))) if axes is None else tuple(axes)
    return lax.transpose(x, axes)


def broadcast_to(x, shape):
    a = abstractify(x)
    shape = tuple(shape)
    if a.shape == shape:
        return x
    k = len(shape) - len(a.shape)
    return lax.broadcast_in_dim(x, shape, tuple(range(k, len(shape))))


def _axes(a, axis):
    if axis is None:
        return tuple(range(len(a.shape)))
    if isinstance(axis, int):
        return (axis % len(a.shape),)
    return tuple(sorted(ax % builtins.max(len(a.shape), 1) for ax in axis))


def _reduction(lax_fn):
    def red(x, axis=None, keepdims=False):
        a = abstractify(x)
        axes = _axes(a, axis)
        out = lax_fn(x, axes)
        if keepdims:
            out = lax.reshape(out, tuple(1 if i in axes else s for i, s in enumerate(a.shape)))
        return out
    return red


sum = _reduction(lax.reduce_sum); max = _reduction(lax.reduce_max)
min = _reduction(lax.reduce_min); prod = _reduction(lax.reduce_prod)
amax, amin = max, min


def mean(x, axis=None, keepdims=False):
    a = abstractify(x)
    axes = _axes(a, axis)
    n = int(np.prod([a.shape[i] for i in axes], dtype=np.int64))
    s = sum(_float(x), axis=axes, keepdims=keepdims)
    return lax.div(s, np.float32(n))


def argmax(x, axis=None):
    if axis is None:
        x, axis = ravel(x), 0
    return lax.argmax(x, axis, np.int32)


def argmin(x, axis=None):
    if axis is None:
        x, axis = ravel(x), 0
    return lax.argmin(x, axis, np.int32)


def where(cond, x, y):
    x, y = _promote(x, y)
    shape = lax._bcast_shape(*[abstractify(v).shape for v in (cond, x, y)])
    return lax.select(broadcast_to(cond, shape), broadcast_to(x, shape), broadcast_to(y, shape))


def concatenate(arrays, axis=0):
    return lax.concatenate(list(arrays), axis)


def arange(n, dtype=np.int32):
    return lax.iota(dtype, int(n))


def dot(x, y):
    x, y = _promote_dtypes_only(x, y)
    ax, ay = abstractify(x), abstractify(y)
    if len(ax.shape) == 0 or len(ay.shape) == 0:
        return multiply(x, y)
    rc = 0 if len(ay.shape) == 1 else len(ay.shape) - 2
    return lax.dot_general(x, y, (((len(ax.shape) - 1,), (rc,)), ((), ())))


matmul = dot


def _promote_dtypes_only(x, y):
    ax, ay = abstractify(x), abstractify(y)
    dt = _result_dtype(ax, ay)
    return _cast(x, dt), _cast(y, dt)


def take_along_axis(arr, indices, axis=-1):
    """≙ jnp.take_along_axis for 2-D arr / (B,1) indices along the last axis: lowers to the gather
    the reference accepts (`gather_fn0`, reference tests/test_basic_ops.py:93-101)."""
    a, i = abstractify(arr), abstractify(indices)
    assert len(a.shape) == 2 and axis in (-1, 1) and i.shape == (a.shape[0], 1), 'only the reference-tested case'
    rows = lax.iota(np.int32, a.shape[0])
    rows = lax.broadcast_in_dim(rows, (a.shape[0], 1, 1), (0,))
    idx = lax.concatenate([rows, lax.reshape(_cast(indices, np.int32), (a.shape[0], 1, 1))], 2)
    dn = GatherDimensionNumbers(offset_dims=(), collapsed_slice_dims=(0, 1), start_index_map=(0, 1))
    return lax.gather(arr, idx, dn, slice_sizes=(1, 1))


def _index_static(x, idx):
    """Static int / contiguous-slice indexing → one `gather` (how JAX 0.2.x lowers `x[5,:]`,
    `x[:,:,4:7,:]`, reference tests/test_basic_ops.py:88-90)."""
    a = abstractify(x)
    if not isinstance(idx, tuple):
        idx = (idx,)
    if builtins.any(i is Ellipsis for i in idx):
        k = idx.index(Ellipsis)
        idx = idx[:k] + (slice(None),) * (len(a.shape) - (len(idx) - 1)) + idx[k + 1:]
    idx = idx + (slice(None),) * (len(a.shape) - len(idx))
    starts, index_map, collapsed, slice_sizes = [], [], [], []
    for d, (i, n) in enumerate(zip(idx, a.shape)):
        if isinstance(i, (int, np.integer)):
            i = int(i) % n
            starts.append(i); index_map.append(d); collapsed.append(d); slice_sizes.append(1)
        elif isinstance(i, slice):
            s, e, st = i.indices(n)
            if st != 1:
                raise NotImplementedError('strided indexing: use lax.slice')
            if (s, e) == (0, n):
                slice_sizes.append(n)
            else:
                starts.append(s); index_map.append(d); slice_sizes.append(builtins.max(e - s, 0))
        else:
            raise NotImplementedError(f'index of type {type(i)}')
    if not starts:
        return x
    offset_dims = tuple(range(len(a.shape) - len(collapsed)))
    dn = GatherDimensionNumbers(offset_dims=offset_dims, collapsed_slice_dims=tuple(collapsed),
                                start_index_map=tuple(index_map))
    return lax.gather(x, np.asarray(starts, np.int32), dn, tuple(slice_sizes))


# ----------------------------------------------------------------------- Tracer operators
def _swap(f):
    return lambda a, b: f(b, a)


Tracer.__add__ = add; Tracer.__radd__ = _swap(add)
Tracer.__sub__ = subtract; Tracer.__rsub__ = _swap(subtract)
Tracer.__mul__ = multiply; Tracer.__rmul__ = _swap(multiply)
Tracer.__truediv__ = true_divide; Tracer.__rtruediv__ = _swap(true_divide)
Tracer.__pow__ = power
Tracer.__matmul__ = dot; Tracer.__rmatmul__ = _swap(dot)
Tracer.__neg__ = negative
Tracer.__abs__ = abs
Tracer.__gt__ = greater; Tracer.__ge__ = greater_equal
Tracer.__lt__ = less; Tracer.__le__ = less_equal
Tracer.__eq__ = equal; Tracer.__ne__ = not_equal
Tracer.__hash__ = lambda self: id(self)
Tracer.__and__ = bitwise_and; Tracer.__rand__ = _swap(bitwise_and)
Tracer.__or__ = bitwise_or; Tracer.__ror__ = _swap(bitwise_or)
Tracer.__xor__ = bitwise_xor
Tracer.__lshift__ = left_shift
Tracer.__rshift__ = _binop(lax.shift_right_arithmetic)
Tracer.__getitem__ = _index_static
Tracer.T = property(lambda self: transpose(self))
Tracer.reshape = reshape
Tracer.ravel = ravel
Tracer.astype = lambda self, dtype: lax.convert_element_type(self, dtype)
Tracer.sum = sum; Tracer.max = max; Tracer.min = min; Tracer.mean = mean; Tracer.prod = prod
Tracer.argmax = argmax; Tracer.argmin = argmin
Tracer.squeeze = squeeze
Tracer.transpose = lambda self, *axes: transpose(self, axes[0] if len(axes) == 1 and not isinstance(axes[0], int) else (axes or None))
