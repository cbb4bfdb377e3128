"""Duck-typed jaxpr IR.

The reference's front end is `jax.make_jaxpr` (reference vkjax/function.py:9) and
every layer below only *reads attributes* from JAX objects.
JAX is not installable in this image, so this module provides objects with the
same attribute surface (`ClosedJaxpr.consts/.jaxpr/.literals`,
`Jaxpr.constvars/.invars/.outvars/.eqns`, `JaxprEqn.primitive.name/.invars/
.outvars/.params`, `Var.aval.shape/.dtype/.count`, `Literal.val/.hash/.aval`).
A real `jax.core.ClosedJaxpr` passes through the interpreter unchanged because
only these attributes (never isinstance checks against this module) are used
downstream -- see `is_literal`, `is_unit`, `is_dropvar`.
"""
import itertools
import string
import typing as tp

import numpy as np


SUPPORTED_DTYPES = (np.bool_, np.int32, np.uint32, np.float32)


class ShapedArray:
    """≙ jax.core.ShapedArray(shape, dtype) (reference vkjax/ops.py:162)."""
    __slots__ = ('shape', 'dtype', 'weak_type')

    def __init__(self, shape, dtype, weak_type=False):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.weak_type = bool(weak_type)

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    def __eq__(self, other):
        return (isinstance(other, ShapedArray) and self.shape == other.shape
                and self.dtype == other.dtype)

    def __hash__(self):
        return hash((self.shape, self.dtype))

    def str_short(self):
        dt = {'float32': 'f32', 'int32': 'i32', 'uint32': 'u32', 'bool': 'bool'}.get(self.dtype.name, self.dtype.name)
        return f'{dt}[{",".join(map(str, self.shape))}]'

    def __repr__(self):
        return f'ShapedArray({self.str_short()})'


class AbstractUnit:
    """≙ jax.core.AbstractUnit: a variable without storage (reference buffers.py:102)."""
    shape = ()
    dtype = np.dtype('float32')

    def __repr__(self):
        return 'AbstractUnit()'


abstract_unit = AbstractUnit()


class ShapeDtypeStruct:
    """≙ jax.ShapeDtypeStruct, what make_jaxpr(..., return_shape=True) returns."""
    __slots__ = ('shape', 'dtype')

    def __init__(self, shape, dtype):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)

    def __repr__(self):
        return f'ShapeDtypeStruct(shape={self.shape}, dtype={self.dtype.name})'


def _var_name(count):
    letters = string.ascii_lowercase
    s = ''
    count += 1
    while count:
        count, r = divmod(count - 1, 26)
        s = letters[r] + s
    return s


class Var:
    """≙ jax.core.Var(count, suffix, aval) (reference tests/test_bufferpool.py:17)."""
    __slots__ = ('count', 'suffix', 'aval')

    def __init__(self, count: int, suffix: str, aval):
        self.count = count
        self.suffix = suffix
        self.aval = aval if not isinstance(aval, tuple) else ShapedArray(*aval)

    def __repr__(self):
        return _var_name(self.count) + self.suffix


class DropVar(Var):
    """An output nobody reads; prints as `_` (reference ops.py:56)."""
    def __init__(self, aval):
        super().__init__(-1, '', aval)

    def __repr__(self):
        return '_'


class UnitVar(Var):
    """The unit variable; prints as `*` (reference ops.py:226,238)."""
    def __init__(self):
        super().__init__(-2, '', abstract_unit)

    def __repr__(self):
        return '*'


unitvar = UnitVar()


class Literal:
    """≙ jax.core.Literal: a value baked into an equation (reference buffers.py:80-86)."""
    __slots__ = ('val', 'aval', 'hash')
    _ids = itertools.count(1)

    def __init__(self, val, aval=None):
        self.val = val
        arr = np.asarray(val)
        self.aval = aval if aval is not None else ShapedArray(arr.shape, arr.dtype)
        # jax literals are not hashable but carry a `.hash`; identity-like is enough
        self.hash = hash((next(Literal._ids), 'literal'))

    def __repr__(self):
        return str(np.asarray(self.val).tolist()) if np.ndim(self.val) == 0 else f'Literal{np.shape(self.val)}'


class Primitive:
    """≙ jax.core.Primitive; only `.name` and `.multiple_results` are read downstream."""
    def __init__(self, name, multiple_results=False):
        self.name = name
        self.multiple_results = multiple_results

    def __repr__(self):
        return self.name


class JaxprEqn(tp.NamedTuple):
    invars: list
    outvars: list
    primitive: Primitive
    params: dict

    def __repr__(self):
        ps = ''
        if self.params:
            items = []
            for k, v in self.params.items():
                if isinstance(v, (Jaxpr, ClosedJaxpr)):
                    v = '{...}'
                items.append(f'{k}={v}')
            ps = '[' + ' '.join(items) + ']'
        outs = ' '.join(map(str, self.outvars))
        ins = ' '.join(map(str, self.invars))
        return f'{outs} = {self.primitive.name}{ps} {ins}'


class Jaxpr:
    def __init__(self, constvars, invars, outvars, eqns):
        self.constvars = list(constvars)
        self.invars = list(invars)
        self.outvars = list(outvars)
        self.eqns = list(eqns)

    def __repr__(self):
        lines = ['{ lambda ' + ' '.join(map(str, self.constvars)) + ' ; ' + ' '.join(map(str, self.invars)) + '.', '  let']
        lines += ['    ' + repr(e) for e in self.eqns]
        lines += ['  in (' + ' '.join(map(str, self.outvars)) + ') }']
        return '\n'.join(lines)


class ClosedJaxpr:
    def __init__(self, jaxpr: Jaxpr, consts):
        self.jaxpr = jaxpr
        self.consts = list(consts)

    @property
    def literals(self):
        return self.consts

    def __repr__(self):
        return repr(self.jaxpr)


class ConvDimensionNumbers(tp.NamedTuple):
    """≙ jax.lax.ConvDimensionNumbers (reference tests/test_conv.py:19)."""
    lhs_spec: tp.Tuple[int, ...]
    rhs_spec: tp.Tuple[int, ...]
    out_spec: tp.Tuple[int, ...]


class GatherDimensionNumbers(tp.NamedTuple):
    """≙ jax.lax.GatherDimensionNumbers (reference tests/test_basic_ops.py:95-99)."""
    offset_dims: tp.Tuple[int, ...]
    collapsed_slice_dims: tp.Tuple[int, ...]
    start_index_map: tp.Tuple[int, ...]


class ScatterDimensionNumbers(tp.NamedTuple):
    """≙ jax.lax.ScatterDimensionNumbers (reference vkjax/ops.py:406-407)."""
    update_window_dims: tp.Tuple[int, ...]
    inserted_window_dims: tp.Tuple[int, ...]
    scatter_dims_to_operand_dims: tp.Tuple[int, ...]


# ---------------------------------------------------------------------------------------------
# duck-type predicates: these, not isinstance(), are what the executor uses so that real
# jax.core objects behave identically.

def is_literal(v) -> bool:
    return hasattr(v, 'val')


def is_unit(v) -> bool:
    return str(v) == '*' or type(getattr(v, 'aval', None)).__name__ == 'AbstractUnit'


def is_dropvar(v) -> bool:
    return str(v) == '_'


def hashable(v):
    """Literals are not hashable in JAX but carry `.hash` (reference buffers.py:80-86)."""
    if is_literal(v):
        try:
            return v.hash
        except AttributeError:
            return id(v)
    return v
