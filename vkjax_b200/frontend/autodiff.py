"""Reverse-mode differentiation for the JAX-free front end: `grad`, `value_and_grad`, `vjp`.

Why it exists: the reference's `vkModel` wraps Elegy's `call_train_step` (reference vkjax/elegy.py:20-23), whose jaxpr is
what `jax.value_and_grad` of the loss emits (reference tests/test_elegy_mlp.py:73-118, tests/test_elegy_conv.py:51-94).
JAX cannot be installed in this image, so to trace a training step *through the model* the front end needs its own
transpose rules.  They follow jax.lax's (JAX 0.2.x) rule by rule, so the backward equations are the ones a real jaxpr
would carry: dot_general gradients as dot_generals over the other contracting pair (no transposes), conv gradients as
lhs-dilated / rhs-dilated conv_general_dilated with permuted dimension numbers and a reversed filter
(_conv_general_dilated_transpose_lhs / _rhs), gather -> scatter-add, reduce_window_max -> select_and_scatter_add,
relu's custom JVP -> select(x > 0, g, 0), cotangent sums as add_any.

Mechanism: the function is traced into the *current* trace (a tape is simply the slice of equations it appended); the
backward pass walks that slice in reverse and records the transposed equations into the same trace.  Nothing is ever
evaluated on the host.
"""
import typing as tp

import numpy as np

from .. import core, tree_util
from ..core import ConvDimensionNumbers, ScatterDimensionNumbers
from . import lax, jnp
from .tracing import Tracer, bind, current_trace, abstractify


# ---- helpers ----------------------------------------------------------------------------------------
def _is_float(aval):
    return np.dtype(aval.dtype).kind == 'f'


def _val(trace, v):
    """equation operand -> something the lax functions accept (a Tracer of `trace`, or the literal's value)"""
    if core.is_literal(v):
        return v.val
    return Tracer(trace, v, v.aval)


def _shape(x):
    return tuple(abstractify(x).shape)


def _zeros(shape, dtype=np.float32):
    return jnp.broadcast_to(np.asarray(0, dtype)[()], tuple(shape))


def _unbroadcast(ct, shape):
    """cotangent of an operand that lax's implicit broadcasting (scalars, size-1 dims of equal rank) expanded"""
    cs, shape = _shape(ct), tuple(shape)
    if cs == shape:
        return ct
    if shape == ():
        return lax.reduce_sum(ct, tuple(range(len(cs))))
    assert len(shape) == len(cs), (shape, cs)
    axes = tuple(i for i, (a, b) in enumerate(zip(shape, cs)) if a == 1 and b != 1)
    return lax.reshape(lax.reduce_sum(ct, axes), shape)


def _add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    return lax.add_any(a, b)


RULES: tp.Dict[str, tp.Callable] = {}


def rule(*names):
    def deco(fn):
        for n in names:
            RULES[n] = fn
        return fn
    return deco


# ---- elementwise ------------------------------------------------------------------------------------
@rule('add', 'add_any')
def _add_rule(eq, ct, ins, outs):
    return [_unbroadcast(ct[0], _shape(x)) for x in ins]


@rule('sub')
def _sub_rule(eq, ct, ins, outs):
    return [_unbroadcast(ct[0], _shape(ins[0])), _unbroadcast(lax.neg(ct[0]), _shape(ins[1]))]


@rule('mul')
def _mul_rule(eq, ct, ins, outs):
    x, y = ins
    return [_unbroadcast(lax.mul(ct[0], y), _shape(x)), _unbroadcast(lax.mul(x, ct[0]), _shape(y))]


@rule('div')
def _div_rule(eq, ct, ins, outs):
    x, y = ins
    gx = _unbroadcast(lax.div(ct[0], y), _shape(x))
    gy = _unbroadcast(lax.mul(lax.mul(lax.neg(ct[0]), x), lax.integer_pow(y, -2)), _shape(y))
    return [gx, gy]


def _balanced_eq(x, z, y):
    one, half, zero = np.float32(1), np.float32(0.5), np.float32(0)
    return jnp.where(jnp.equal(x, z), jnp.where(jnp.equal(y, z), half, one), zero)


@rule('max', 'min')
def _max_rule(eq, ct, ins, outs):
    x, y = ins
    return [_unbroadcast(jnp.multiply(ct[0], _balanced_eq(x, outs[0], y)), _shape(x)),
            _unbroadcast(jnp.multiply(ct[0], _balanced_eq(y, outs[0], x)), _shape(y))]


@rule('neg')
def _neg_rule(eq, ct, ins, outs):
    return [lax.neg(ct[0])]


@rule('exp')
def _exp_rule(eq, ct, ins, outs):
    return [lax.mul(ct[0], outs[0])]


@rule('log')
def _log_rule(eq, ct, ins, outs):
    return [lax.div(ct[0], ins[0])]


@rule('tanh')
def _tanh_rule(eq, ct, ins, outs):
    return [lax.mul(ct[0], lax.sub(np.float32(1), lax.mul(outs[0], outs[0])))]


@rule('logistic')
def _logistic_rule(eq, ct, ins, outs):
    return [lax.mul(ct[0], lax.mul(outs[0], lax.sub(np.float32(1), outs[0])))]


@rule('sqrt')
def _sqrt_rule(eq, ct, ins, outs):
    return [lax.div(lax.mul(ct[0], np.float32(0.5)), outs[0])]


@rule('rsqrt')
def _rsqrt_rule(eq, ct, ins, outs):
    return [lax.mul(ct[0], lax.mul(np.float32(-0.5), lax.div(outs[0], ins[0])))]


@rule('sin')
def _sin_rule(eq, ct, ins, outs):
    return [lax.mul(ct[0], lax.cos(ins[0]))]


@rule('cos')
def _cos_rule(eq, ct, ins, outs):
    return [lax.neg(lax.mul(ct[0], lax.sin(ins[0])))]


@rule('integer_pow')
def _ipow_rule(eq, ct, ins, outs):
    y = int(eq.params['y'])
    if y == 0:
        return [None]
    return [lax.mul(ct[0], lax.mul(np.float32(y), lax.integer_pow(ins[0], y - 1) if y != 1 else _ones_like(ins[0])))]


def _ones_like(x):
    return jnp.broadcast_to(np.float32(1), _shape(x))


@rule('convert_element_type')
def _convert_rule(eq, ct, ins, outs):
    a = abstractify(ins[0])
    return [lax.convert_element_type(ct[0], a.dtype) if _is_float(a) else None]


@rule('select')
def _select_rule(eq, ct, ins, outs):
    pred = ins[0]
    z = _zeros(_shape(ct[0]))
    return [None, lax.select(pred, ct[0], z), lax.select(pred, z, ct[0])]


@rule('stop_gradient', 'iota', 'gt', 'ge', 'lt', 'le', 'eq', 'ne', 'argmax', 'argmin', 'sign', 'floor', 'ceil', 'round',
      'threefry2x32', 'and', 'or', 'not', 'xor', 'shift_left', 'shift_right_logical', 'shift_right_arithmetic', 'rem')
def _no_grad(eq, ct, ins, outs):
    return [None] * len(ins)


# ---- structural -------------------------------------------------------------------------------------
@rule('reshape', 'squeeze', 'expand_dims', 'copy')
def _reshape_rule(eq, ct, ins, outs):
    return [lax.reshape(ct[0], _shape(ins[0]))]


@rule('broadcast_in_dim')
def _bcast_rule(eq, ct, ins, outs):
    in_shape, out_shape = _shape(ins[0]), _shape(outs[0])
    bdims = tuple(eq.params['broadcast_dimensions'])
    axes = tuple(d for d in range(len(out_shape)) if d not in bdims) + \
        tuple(d for i, d in enumerate(bdims) if in_shape[i] == 1 and out_shape[d] != 1)
    g = lax.reduce_sum(ct[0], tuple(sorted(axes))) if axes else ct[0]
    return [lax.reshape(g, in_shape)]


@rule('reduce_sum')
def _rsum_rule(eq, ct, ins, outs):
    in_shape = _shape(ins[0])
    axes = tuple(eq.params['axes'])
    kept = tuple(d for d in range(len(in_shape)) if d not in axes)
    return [lax.broadcast_in_dim(ct[0], in_shape, kept)]


@rule('reduce_max', 'reduce_min')
def _rmax_rule(eq, ct, ins, outs):
    """jax.lax _reduce_chooser_jvp_rule: the cotangent is shared equally between the locations that attain the extremum"""
    x = ins[0]
    in_shape = _shape(x)
    axes = tuple(eq.params['axes'])
    kept = tuple(d for d in range(len(in_shape)) if d not in axes)
    loc = lax.convert_element_type(lax.eq(x, lax.broadcast_in_dim(outs[0], in_shape, kept)), np.float32)
    counts = lax.reduce_sum(loc, axes)
    return [lax.mul(lax.broadcast_in_dim(lax.div(ct[0], counts), in_shape, kept), loc)]


@rule('transpose')
def _transpose_rule(eq, ct, ins, outs):
    perm = tuple(eq.params['permutation'])
    return [lax.transpose(ct[0], tuple(int(i) for i in np.argsort(perm)))]


@rule('rev')
def _rev_rule(eq, ct, ins, outs):
    return [lax.rev(ct[0], tuple(eq.params['dimensions']))]


@rule('gather')
def _gather_rule(eq, ct, ins, outs):
    """jax.lax _gather_transpose_rule"""
    dn = eq.params['dimension_numbers']
    sdn = ScatterDimensionNumbers(update_window_dims=tuple(dn.offset_dims), inserted_window_dims=tuple(dn.collapsed_slice_dims),
                                  scatter_dims_to_operand_dims=tuple(dn.start_index_map))
    return [lax.scatter_add(_zeros(_shape(ins[0])), ins[1], ct[0], sdn), None]


# ---- contractions -----------------------------------------------------------------------------------
@rule('dot_general')
def _dot_rule(eq, ct, ins, outs):
    (lc, rc), (lb, rb) = eq.params['dimension_numbers']
    if tuple(lb) or tuple(rb) or len(_shape(ins[0])) != 2 or len(_shape(ins[1])) != 2:
        raise NotImplementedError('grad of dot_general with batch dimensions / rank != 2')
    a, b = ins
    ca, cb = int(lc[0]), int(rc[0])
    g = ct[0]                                               # [n, m]
    dn = lambda x, y: (((x,), (y,)), ((), ()))
    # dA[n, c] = sum_m g[n, m] B[c, m]; stored [c, n] when A is (contracting dim 0)
    da = lax.dot_general(g, b, dn(1, 1 - cb)) if ca == 1 else lax.dot_general(b, g, dn(1 - cb, 1))
    # dB[c, m] = sum_n A[n, c] g[n, m]; stored [m, c] when B contracts on dim 1
    db = lax.dot_general(a, g, dn(1 - ca, 0)) if cb == 0 else lax.dot_general(g, a, dn(0, 1 - ca))
    return [da, db]


def _dilate(shape, dil):
    return [0 if s == 0 else 1 + d * (s - 1) for s, d in zip(shape, dil)]


def _spec_transpose(spec):
    return (spec[1], spec[0]) + tuple(spec[2:])


@rule('conv_general_dilated')
def _conv_rule(eq, ct, ins, outs):
    """jax.lax _conv_general_dilated_transpose_lhs / _rhs (feature_group_count == batch_group_count == 1)"""
    p = eq.params
    assert p['feature_group_count'] == 1 and p['batch_group_count'] == 1
    lhs, rhs = ins
    g = ct[0]
    dn = p['dimension_numbers']
    lhs_spec, rhs_spec, out_spec = tuple(dn.lhs_spec), tuple(dn.rhs_spec), tuple(dn.out_spec)
    strides, padding = tuple(p['window_strides']), tuple(p['padding'])
    lhs_dil, rhs_dil = tuple(p['lhs_dilation']), tuple(p['rhs_dilation'])
    ls, rs, os_ = _shape(lhs), _shape(rhs), _shape(g)
    in_sp = [ls[d] for d in lhs_spec[2:]]
    win = [rs[d] for d in rhs_spec[2:]]
    out_sp = [os_[d] for d in out_spec[2:]]
    lhs_d, rhs_d, out_d = _dilate(in_sp, lhs_dil), _dilate(win, rhs_dil), _dilate(out_sp, strides)
    # d lhs: the cotangent dilated by the stride, convolved with the reversed filter with I / O swapped
    pad_before = [r - lo - 1 for r, (lo, _) in zip(rhs_d, padding)]
    pad_after = [l + r - 1 - o - pb for l, r, o, pb in zip(lhs_d, rhs_d, out_d, pad_before)]
    t_dn = ConvDimensionNumbers(out_spec, _spec_transpose(rhs_spec), lhs_spec)
    dlhs = lax.conv_general_dilated(g, lax.rev(rhs, rhs_spec[2:]), lhs_dil, list(zip(pad_before, pad_after)),
                                    lhs_dilation=strides, rhs_dilation=rhs_dil, dimension_numbers=t_dn)
    # d rhs: the lhs convolved with the cotangent as a filter dilated by the stride; batch is contracted
    pads_hi = [(o - l) + (r - lo - 1) for o, l, r, (lo, _) in zip(out_d, lhs_d, rhs_d, padding)]
    t_dn = ConvDimensionNumbers(_spec_transpose(lhs_spec), _spec_transpose(out_spec), _spec_transpose(rhs_spec))
    drhs = lax.conv_general_dilated(lhs, g, rhs_dil, [(lo, hi) for (lo, _), hi in zip(padding, pads_hi)],
                                    lhs_dilation=lhs_dil, rhs_dilation=strides, dimension_numbers=t_dn)
    return [dlhs, drhs]


@rule('reduce_window_max', 'reduce_window_min')
def _rwmax_rule(eq, ct, ins, outs):
    p = eq.params
    sel = lax.ge if eq.primitive.name.endswith('max') else lax.le
    return [lax.select_and_scatter_add(ct[0], ins[0], sel, p['window_dimensions'], p['window_strides'], p['padding'])]


# ---- call primitives: re-record the callee inline (rematerialisation), then differentiate that -------------
def _replay(jaxpr, consts, in_vals):
    """records the equations of `jaxpr` into the current trace with `in_vals` as arguments; returns the outputs"""
    env = {}

    def read(v):
        return v.val if core.is_literal(v) else env[core.hashable(v)]
    for v, x in zip(jaxpr.invars, in_vals):
        env[core.hashable(v)] = x
    for v, c in zip(getattr(jaxpr, 'constvars', []), consts):
        env[core.hashable(v)] = c
    for e in jaxpr.eqns:
        outs = bind(e.primitive, *[read(v) for v in e.invars], out_avals=[v.aval for v in e.outvars], **e.params)
        outs = outs if e.primitive.multiple_results else [outs]
        for v, o in zip(e.outvars, outs):
            env[core.hashable(v)] = o
    return [read(v) for v in jaxpr.outvars]


def _is_relu(jaxpr):
    if len(jaxpr.eqns) != 1 or jaxpr.eqns[0].primitive.name != 'max':
        return False
    lits = [v for v in jaxpr.eqns[0].invars if core.is_literal(v)]
    return len(lits) == 1 and float(lits[0].val) == 0.0


@rule('custom_jvp_call_jaxpr', 'custom_jvp_call', 'xla_call', 'pjit', 'closed_call', 'core_call')
def _call_rule(eq, ct, ins, outs):
    inner = eq.params.get('fun_jaxpr', eq.params.get('call_jaxpr', eq.params.get('jaxpr')))
    consts = getattr(inner, 'consts', ())
    jaxpr = getattr(inner, 'jaxpr', inner)
    if eq.primitive.name.startswith('custom_jvp') and _is_relu(jaxpr):
        # jax.nn.relu's own rule: relu.defjvps(lambda g, ans, x: lax.select(x > 0, g, lax.full_like(g, 0)))
        x = ins[0]
        return [lax.select(lax.gt(x, np.float32(0.0)), ct[0], _zeros(_shape(ct[0])))]
    trace = current_trace()
    start = len(trace.eqns)
    re_outs = _replay(jaxpr, consts, ins)
    cts = {}
    for o, c in zip(re_outs, ct):
        if c is not None and isinstance(o, Tracer):
            cts[core.hashable(o.var)] = c
    cts = backward(trace, trace.eqns[start:], cts)
    return [cts.get(core.hashable(x.var)) if isinstance(x, Tracer) else None for x in ins]


# ---- the backward pass --------------------------------------------------------------------------------
def backward(trace, eqns, cts):
    """`cts`: {hashable(var): cotangent tracer}.  Walks `eqns` in reverse, returns the completed cotangent map."""
    cts = dict(cts)
    for eq in reversed(list(eqns)):
        out_cts = [cts.get(core.hashable(v)) if not core.is_dropvar(v) else None for v in eq.outvars]
        if all(c is None for c in out_cts):
            continue
        name = eq.primitive.name.replace('-', '_')
        if name not in RULES:
            raise NotImplementedError(f'no transpose rule for primitive {eq.primitive.name}')
        ins = [_val(trace, v) for v in eq.invars]
        outs = [Tracer(trace, v, v.aval) for v in eq.outvars]
        if len(out_cts) == 1 or not any(c is None for c in out_cts):
            pass
        else:                                              # multiple results with some zero cotangents
            out_cts = [c if c is not None else _zeros(v.aval.shape, v.aval.dtype) for c, v in zip(out_cts, eq.outvars)]
        in_cts = RULES[name](eq, out_cts, ins, outs)
        for v, c in zip(eq.invars, in_cts):
            if c is None or core.is_literal(v) or not _is_float(v.aval):
                continue
            k = core.hashable(v)
            cts[k] = _add(cts.get(k), c)
    return cts


def _prune(trace, start, roots):
    """Dead-code elimination over the equations recorded since `start` (the backward pass): rules compute the cotangent of
    every operand, also of literals / integer operands nobody asks for; whatever does not reach `roots` is dropped."""
    live = {core.hashable(r.var) for r in roots if isinstance(r, Tracer)}
    kept = []
    for eq in reversed(trace.eqns[start:]):
        if any(core.hashable(v) in live for v in eq.outvars):
            kept.append(eq)
            for v in eq.invars:
                if not core.is_literal(v):
                    live.add(core.hashable(v))
    trace.eqns[start:] = kept[::-1]


def _run_backward(trace, tape_start, tape_end, out_leaves, ct_leaves, primals):
    cts = {}
    for o, c in zip(out_leaves, ct_leaves):
        if c is not None and isinstance(o, Tracer) and o.trace is trace:
            cts[core.hashable(o.var)] = _add(cts.get(core.hashable(o.var)), c)
    cts = backward(trace, trace.eqns[tape_start:tape_end], cts)
    found = []

    def pick(x):
        if isinstance(x, Tracer) and x.trace is trace:
            g = cts.get(core.hashable(x.var))
            if g is not None:
                found.append(g)
                return g
            return _zeros(x.shape, x.dtype)
        return _zeros(_shape(x), abstractify(x).dtype)
    grads = tuple(tree_util.tree_map(pick, p) for p in primals)
    _prune(trace, tape_end, [g for g in tree_util.tree_leaves(grads) if isinstance(g, Tracer)])
    return grads


def vjp(fun, *primals):
    """≙ jax.vjp inside a trace: returns (outputs, vjp_fn); vjp_fn(cotangents) -> cotangents of `primals` (pytrees).
    vjp_fn may be called once (the backward equations are appended right behind the forward ones)."""
    trace = current_trace()
    start = len(trace.eqns)
    out = fun(*primals)
    end = len(trace.eqns)
    out_leaves = tree_util.tree_leaves(out)

    def vjp_fn(cotangents):
        ct_leaves = tree_util.tree_leaves(cotangents)
        assert len(ct_leaves) == len(out_leaves)
        assert len(trace.eqns) == end, 'vjp_fn must be called before anything else is traced'
        return _run_backward(trace, start, end, out_leaves, ct_leaves, primals)
    return out, vjp_fn


def value_and_grad(fun, argnums: tp.Union[int, tp.Sequence[int]] = 0, has_aux: bool = False):
    """≙ jax.value_and_grad for use INSIDE a traced function (vkjax.wrap / make_jaxpr)."""
    single = isinstance(argnums, int)
    nums = (argnums,) if single else tuple(argnums)

    def wrapped(*args):
        trace = current_trace()
        start = len(trace.eqns)
        out = fun(*args)
        end = len(trace.eqns)
        value, aux = out if has_aux else (out, None)
        if _shape(value) != ():
            raise TypeError(f'grad requires a scalar-output function, got shape {_shape(value)}')
        one = np.asarray(1, abstractify(value).dtype)[()]
        grads = _run_backward(trace, start, end, [value], [one], [args[i] for i in nums])
        grads = grads[0] if single else grads
        return ((value, aux), grads) if has_aux else (value, grads)
    return wrapped


def grad(fun, argnums=0, has_aux=False):
    vg = value_and_grad(fun, argnums, has_aux)

    def wrapped(*args):
        out, g = vg(*args)
        return (g, out[1]) if has_aux else g
    return wrapped
