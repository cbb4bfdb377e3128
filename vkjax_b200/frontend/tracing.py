"""A small JAX-free tracer that emits jaxprs in the JAX-0.2.x dialect the reference consumes.

≙ `jax.make_jaxpr(fun, static_argnums, return_shape=True)` at reference vkjax/function.py:9.
`jax` cannot be installed in this image, so `vkjax.wrap` needs its own
producer of jaxprs.  Design points:

* Every primitive application inside a trace is *recorded*, even when all operands are
  concrete constants.  Nothing is ever evaluated on the host: there is no CPU path in the
  product (north star: "no CPU fallback").
* Python / numpy scalars become `Literal`s; closed-over arrays become constvars
  (≙ `dot1_const`, reference tests/test_basic_ops.py:38-39).
* `jit(f)` inside a trace records an `xla_call` equation with an open inner jaxpr
  (≙ reference ops.py:218-244); `custom_jvp_call_jaxpr` is used by `nn.relu`
  (≙ reference ops.py:246-260).
"""
import typing as tp

import numpy as np

from .. import core
from .. import tree_util


def canonicalize_dtype(dtype) -> np.dtype:
    """JAX without x64: float64→float32, int64→int32, uint64→uint32 (≙ reference buffers.py:73-76)."""
    dtype = np.dtype(dtype)
    # the dtypes the reference accepts on upload (reference buffers.py:69); anything else is NotImplementedError
    table = {'float32': 'float32', 'float64': 'float32', 'int32': 'int32', 'int64': 'int32',
             'uint32': 'uint32', 'uint64': 'uint32', 'bool': 'bool',
             # extension: packed 8-bit pixels as an INPUT dtype; the only primitive that accepts them is
             # convert_element_type (images are uploaded as bytes -- 4x less PCIe traffic -- and widened on the device)
             'uint8': 'uint8'}
    if dtype.name not in table:
        raise NotImplementedError(f'{dtype} data types currently not supported')
    return np.dtype(table[dtype.name])


class Trace:
    def __init__(self, parent: tp.Optional['Trace'] = None):
        self.parent = parent
        self.level = 0 if parent is None else parent.level + 1
        self.eqns: tp.List[core.JaxprEqn] = []
        self.counter = parent.counter if parent is not None else [0]   # shared var numbering
        self.constvars: tp.List[core.Var] = []
        self.consts: tp.List[np.ndarray] = []
        self._const_ids: tp.Dict[int, 'Tracer'] = {}
        self.lifted: tp.List[tp.Tuple['Tracer', core.Var]] = []        # closed-over outer tracers
        self._lift_ids: tp.Dict[int, 'Tracer'] = {}

    def new_var(self, aval) -> core.Var:
        v = core.Var(self.counter[0], '', aval)
        self.counter[0] += 1
        return v

    def new_tracer(self, aval) -> 'Tracer':
        return Tracer(self, self.new_var(aval), aval)

    # -- operands -----------------------------------------------------------------------------
    def to_operand(self, x):
        """Returns (var_or_literal, aval) usable as an equation invar in *this* trace."""
        if isinstance(x, Tracer):
            if x.trace is self:
                return x.var, x.aval
            # tracer of an enclosing trace captured by closure: lift as an implicit argument
            if id(x) not in self._lift_ids:
                inner = self.new_tracer(x.aval)
                self._lift_ids[id(x)] = inner
                self.lifted.append((x, inner.var))
            t = self._lift_ids[id(x)]
            return t.var, t.aval
        if isinstance(x, (bool, int, float, np.generic)) or (isinstance(x, np.ndarray) and x.ndim == 0):
            weak = isinstance(x, (bool, int, float))
            arr = np.asarray(x)
            arr = arr.astype(canonicalize_dtype(arr.dtype))
            lit = core.Literal(arr[()], core.ShapedArray((), arr.dtype, weak_type=weak))
            return lit, lit.aval
        arr = np.asarray(x)
        if arr.dtype == object:
            raise TypeError(f'Cannot interpret value of type {type(x)} as an abstract array')
        key = id(x)
        if key not in self._const_ids:
            arr = np.ascontiguousarray(arr.astype(canonicalize_dtype(arr.dtype)))
            t = self.new_tracer(core.ShapedArray(arr.shape, arr.dtype))
            self.constvars.append(t.var)
            self.consts.append(arr)
            self._const_ids[key] = t
            t._keepalive = x
        t = self._const_ids[key]
        return t.var, t.aval

    def add_eqn(self, prim, invars, out_avals, params) -> tp.List['Tracer']:
        outs = [self.new_tracer(a) for a in out_avals]
        self.eqns.append(core.JaxprEqn(list(invars), [o.var for o in outs], prim, dict(params)))
        return outs


_trace_stack: tp.List[Trace] = []


def current_trace() -> Trace:
    if not _trace_stack:
        raise RuntimeError('vkjax_b200.frontend primitives can only be used inside vkjax.wrap / make_jaxpr: '
                           'there is no host-side (CPU) evaluation path.')
    return _trace_stack[-1]


class Tracer:
    __array_priority__ = 1000
    __slots__ = ('trace', 'var', 'aval', '_keepalive')

    def __init__(self, trace, var, aval):
        self.trace = trace
        self.var = var
        self.aval = aval

    shape = property(lambda self: self.aval.shape)
    dtype = property(lambda self: self.aval.dtype)
    ndim = property(lambda self: len(self.aval.shape))
    size = property(lambda self: self.aval.size)
    weak_type = property(lambda self: getattr(self.aval, 'weak_type', False))

    def __len__(self):
        if not self.aval.shape:
            raise TypeError('len() of unsized object')
        return self.aval.shape[0]

    def __repr__(self):
        return f'Traced<{self.aval.str_short()}>'

    def __bool__(self):
        raise TypeError('Abstract tracer value encountered where concrete value is expected')

    def __iter__(self):
        if not self.aval.shape:
            raise TypeError('iteration over a 0-d array')
        from . import jnp
        return iter([jnp._index_static(self, (i,)) for i in range(self.aval.shape[0])])

    # operators are attached by frontend.jnp (avoids a circular import)


def bind(prim: core.Primitive, *args, out_avals, **params):
    trace = current_trace()
    invars = [trace.to_operand(a)[0] for a in args]
    outs = trace.add_eqn(prim, invars, out_avals, params)
    return outs if prim.multiple_results else outs[0]


def abstractify(x) -> core.ShapedArray:
    if isinstance(x, Tracer):
        return x.aval
    weak = isinstance(x, (bool, int, float))
    arr = np.asarray(x)
    return core.ShapedArray(arr.shape, canonicalize_dtype(arr.dtype), weak_type=weak)


def _trace_to_jaxpr(fun, in_avals_flat, in_tree, static_args: dict, parent=None):
    """Runs fun with tracers; returns (Trace, in_tracers, out_leaves(operands), out_tree)."""
    trace = Trace(parent)
    _trace_stack.append(trace)
    try:
        in_tracers = [trace.new_tracer(a) for a in in_avals_flat]
        dyn_args = tree_util.tree_unflatten(in_tree, in_tracers)
        n = len(dyn_args) + len(static_args)
        args, it = [], iter(dyn_args)
        for i in range(n):
            args.append(static_args[i] if i in static_args else next(it))
        out = fun(*args)
        out_leaves, out_tree = tree_util.tree_flatten(out)
        out_ops = [trace.to_operand(o) for o in out_leaves]
    finally:
        _trace_stack.pop()
    return trace, in_tracers, out_ops, out_tree


def make_jaxpr(fun: tp.Callable, static_argnums=(), return_shape: bool = False):
    """≙ jax.make_jaxpr.  Returns f(*args) -> ClosedJaxpr [, pytree of ShapeDtypeStruct]."""
    if isinstance(static_argnums, int):
        static_argnums = (static_argnums,)
    static_argnums = tuple(static_argnums)

    def jaxpr_maker(*args):
        static_args = {i: a for i, a in enumerate(args) if i in static_argnums}
        dyn_args = tuple(a for i, a in enumerate(args) if i not in static_argnums)
        leaves, in_tree = tree_util.tree_flatten(dyn_args)
        in_avals = [abstractify(x) for x in leaves]
        in_avals = [core.ShapedArray(a.shape, a.dtype) for a in in_avals]    # arguments are never weak
        trace, in_tracers, out_ops, out_tree = _trace_to_jaxpr(fun, in_avals, in_tree, static_args)
        if trace.lifted:
            raise RuntimeError('leaked tracer from an enclosing trace')
        jaxpr = core.Jaxpr(trace.constvars, [t.var for t in in_tracers], [v for v, _ in out_ops], trace.eqns)
        closed = core.ClosedJaxpr(jaxpr, trace.consts)
        if not return_shape:
            return closed
        shapes = tree_util.tree_unflatten(out_tree, [core.ShapeDtypeStruct(a.shape, a.dtype) for _, a in out_ops])
        return closed, shapes

    jaxpr_maker.__name__ = f'make_jaxpr({getattr(fun, "__name__", "fun")})'
    return jaxpr_maker


xla_call_p = core.Primitive('xla_call', multiple_results=True)
custom_jvp_call_jaxpr_p = core.Primitive('custom_jvp_call_jaxpr', multiple_results=True)


def _call_subjaxpr(fun, args, kind: str, name: str):
    outer = current_trace()
    leaves, in_tree = tree_util.tree_flatten(tuple(args))
    in_avals = [abstractify(x) for x in leaves]
    inner, in_tracers, out_ops, out_tree = _trace_to_jaxpr(fun, in_avals, in_tree, {}, parent=outer)
    # lift inner constants + closed-over outer tracers to explicit arguments (call jaxprs have no constvars)
    extra_invars = list(inner.constvars) + [v for _, v in inner.lifted]
    extra_args = list(inner.consts) + [t for t, _ in inner.lifted]
    jaxpr = core.Jaxpr([], [t.var for t in in_tracers] + extra_invars, [v for v, _ in out_ops], inner.eqns)
    operands = [outer.to_operand(a)[0] for a in list(leaves) + extra_args]
    out_avals = [a for _, a in out_ops]
    if kind == 'xla_call':
        params = dict(backend=None, call_jaxpr=jaxpr, device=None,
                      donated_invars=(False,) * len(operands), name=name)
        outs = outer.add_eqn(xla_call_p, operands, out_avals, params)
    else:
        assert not extra_invars, 'custom_jvp functions must not close over values'
        params = dict(fun_jaxpr=core.ClosedJaxpr(jaxpr, []), jvp_jaxpr_thunk=None, num_consts=0)
        outs = outer.add_eqn(custom_jvp_call_jaxpr_p, operands, out_avals, params)
    return tree_util.tree_unflatten(out_tree, outs)


def jit(fun: tp.Callable):
    """≙ jax.jit used *inside* a traced function (reference tests/test_basic_ops.py:20-21)."""
    def jitted(*args):
        return _call_subjaxpr(fun, args, 'xla_call', getattr(fun, '__name__', 'fun'))
    jitted.__name__ = getattr(fun, '__name__', 'fun')
    return jitted


def custom_jvp(fun: tp.Callable):
    """≙ jax.custom_jvp as seen by the executor: a `custom_jvp_call_jaxpr` equation."""
    def wrapped(*args):
        return _call_subjaxpr(fun, args, 'custom_jvp', getattr(fun, '__name__', 'fun'))
    wrapped.__name__ = getattr(fun, '__name__', 'fun')
    return wrapped
