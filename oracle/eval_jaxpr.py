"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under vkjax_b200/ may import this package.

A numpy restatement of what every primitive on the hot path computes, one function per
primitive name, plus `eval_jaxpr` (≙ reference tests/common.py:12-51, which evaluates a jaxpr
with `primitive.bind` on the JAX CPU backend and records every variable).

PARITY PINNING STATUS: *partially pinned*.  The reference's own tests hold almost no fixed
vectors: truth there is a live JAX-CPU run, and neither jax nor kp/Vulkan can
run in this image.  What *is* pinned (tests/test_oracle_golden.py): the Random123
Threefry-2x32-20 known-answer vectors; JAX's documented `random.split/uniform/normal(PRNGKey(0))`
values through the whole threefry→bits→erf_inv chain; `f(65)==66`, `f(-5)==-4`
(reference tests/test_function.py:13-16); the nextafter direction properties
(reference tests/test_basic_ops.py:367-382); conv/dot/pool against the C restatement of the
reference shaders (oracle/shader_ref.c) and float64 recomputation.  Everything else is a
restatement of `jax.lax` semantics in float32, cross-checked in float64 -- "parity unpinned"
for those rows.

Semantics follow jax.lax (what the reference's tests call `ytrue`); where the reference's
GLSL deviates (quirks Q3-Q7) the lax behaviour is implemented and the
deviation noted at the function.
"""
import math
import os

import numpy as np

F32 = np.float32


def _is_literal(v):
    return hasattr(v, 'val')


# ------------------------------------------------------------------------------ elementwise
def _bc(x, y):
    return np.asarray(x), np.asarray(y)


def _int_div(a, b):
    # lax.div on integers truncates toward zero (C semantics)
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == 'u':
        return (a // np.where(b == 0, 1, b)).astype(a.dtype)
    q = np.abs(a.astype(np.int64)) // np.maximum(np.abs(b.astype(np.int64)), 1)
    return (q * np.sign(a.astype(np.int64)) * np.sign(b.astype(np.int64))).astype(a.dtype)


def _int_rem(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == 'f':
        return np.fmod(a, b).astype(a.dtype)
    if a.dtype.kind == 'u':
        return (a % np.where(b == 0, 1, b)).astype(a.dtype)
    return np.fmod(a.astype(np.int64), np.where(b == 0, 1, b).astype(np.int64)).astype(a.dtype)


def _shift(kind):
    def fn(a, b):
        a, b = np.asarray(a), np.asarray(b)
        bits = 32
        bu = b.astype(np.int64) & 0xFFFFFFFF if b.dtype.kind == 'i' else b.astype(np.int64)
        over = bu >= bits
        sh = np.where(over, 0, bu)
        if kind == 'left':
            r = (a.astype(np.int64) << sh).astype(np.int64) & 0xFFFFFFFF
            r = np.where(over, 0, r)
            return r.astype(np.uint32).view(np.uint32).astype(np.uint32).view(a.dtype) if a.dtype != np.uint32 else r.astype(np.uint32)
        if kind == 'logical':
            # lax semantics: shifting by >= 32 gives 0 (reference clamps to 31: quirk Q7)
            r = (a.view(np.uint32).astype(np.int64) >> sh)
            r = np.where(over, 0, r)
            return r.astype(np.uint32).view(a.dtype)
        if kind == 'arith':
            # XLA emits AShr on the bit pattern whatever the signedness (elemental_ir_emitter), shifts
            # >= 32 saturate to the sign bit; the reference shader is hard-typed int (shift_right_arithmetic.comp)
            ai = a.view(np.int32) if a.dtype == np.uint32 else a
            r = np.where(over, np.where(ai < 0, -1, 0), ai.astype(np.int64) >> sh).astype(np.int32)
            return r.view(a.dtype) if a.dtype == np.uint32 else r
    return fn


def _nextafter(a, b):
    return np.nextafter(np.asarray(a, F32), np.asarray(b, F32)).astype(F32)


def _pow(a, b):
    with np.errstate(all='ignore'):
        return np.power(np.asarray(a, F32), np.asarray(b, F32)).astype(F32)


def _logical_or_bitwise(npop):
    def fn(a, b):
        a, b = np.asarray(a), np.asarray(b)
        return npop(a, b)
    return fn


BINARY = {
    'add': lambda a, b: np.add(*_bc(a, b)),
    'add_any': lambda a, b: np.add(*_bc(a, b)),
    'sub': lambda a, b: np.subtract(*_bc(a, b)),
    'mul': lambda a, b: np.multiply(*_bc(a, b)),
    'div': lambda a, b: (np.divide(*_bc(a, b)) if np.asarray(a).dtype.kind == 'f' else _int_div(a, b)),
    'max': lambda a, b: np.maximum(*_bc(a, b)),
    'min': lambda a, b: np.minimum(*_bc(a, b)),
    'gt': lambda a, b: np.greater(*_bc(a, b)),
    'ge': lambda a, b: np.greater_equal(*_bc(a, b)),
    'lt': lambda a, b: np.less(*_bc(a, b)),
    'le': lambda a, b: np.less_equal(*_bc(a, b)),
    'eq': lambda a, b: np.equal(*_bc(a, b)),        # reference compares bit patterns as float (Q5); lax: typed
    'ne': lambda a, b: np.not_equal(*_bc(a, b)),
    'and': _logical_or_bitwise(np.bitwise_and),
    'or': _logical_or_bitwise(np.bitwise_or),
    'xor': _logical_or_bitwise(np.bitwise_xor),
    'rem': _int_rem,
    'pow': _pow,
    'nextafter': _nextafter,
    'atan2': lambda a, b: np.arctan2(np.asarray(a, F32), np.asarray(b, F32)).astype(F32),
    'shift_left': _shift('left'),
    'shift_right_logical': _shift('logical'),
    'shift_right_arithmetic': _shift('arith'),
}


def _erf(x):
    from scipy import special
    return special.erf(np.asarray(x, np.float64)).astype(F32)     # reference: A&S 7.1.26, 1.5e-7 (Q7)


def _erf_inv(x):
    from scipy import special
    return special.erfinv(np.asarray(x, np.float64)).astype(F32)  # reference: Winitzki approx, 2e-3 (Q7)


def _f64(fn):
    def wrapped(x):
        with np.errstate(all='ignore'):
            return fn(np.asarray(x, np.float64)).astype(F32)
    return wrapped


def _sign(x):
    x = np.asarray(x)
    return np.sign(x).astype(x.dtype)


UNARY = {
    'exp': _f64(np.exp), 'log': _f64(np.log), 'log1p': _f64(np.log1p), 'expm1': _f64(np.expm1),
    'neg': lambda x: np.negative(np.asarray(x)), 'abs': lambda x: np.abs(np.asarray(x)),
    'rsqrt': lambda x: (F32(1.0) / np.sqrt(np.asarray(x, F32))).astype(F32),   # reference rsqrt.comp:11: 1.0/sqrt(x)
    'sqrt': lambda x: np.sqrt(np.asarray(x, F32)).astype(F32),
    'erf': _erf, 'erf_inv': _erf_inv,
    'erfc': lambda x: (1.0 - _erf(x).astype(np.float64)).astype(F32),
    'cos': _f64(np.cos), 'sin': _f64(np.sin), 'tan': _f64(np.tan),
    'cosh': _f64(np.cosh), 'sinh': _f64(np.sinh), 'tanh': _f64(np.tanh),
    'acos': _f64(np.arccos), 'asin': _f64(np.arcsin), 'atan': _f64(np.arctan),
    'acosh': _f64(np.arccosh), 'asinh': _f64(np.arcsinh), 'atanh': _f64(np.arctanh),
    'ceil': _f64(np.ceil), 'floor': _f64(np.floor), 'sign': _sign,
    'round': lambda x: np.where(np.asarray(x) >= 0, np.floor(np.asarray(x, np.float64) + 0.5),
                                np.ceil(np.asarray(x, np.float64) - 0.5)).astype(F32),   # lax.round: away from zero
    'logistic': _f64(lambda x: 1.0 / (1.0 + np.exp(-x))),
    'not': lambda x: (np.logical_not(x) if np.asarray(x).dtype == np.bool_ else np.bitwise_not(np.asarray(x))),
    'stop_gradient': lambda x: np.asarray(x),
}


def integer_pow(x, y):
    x = np.asarray(x)
    if y >= 0:
        out = np.ones_like(x)
        for _ in range(y):      # reference integer_pow.comp:21-23: repeated multiply (only y>=0 there)
            out = out * x
        return out.astype(x.dtype)
    return (F32(1.0) / integer_pow(x, -y)).astype(x.dtype)


def convert_element_type(x, new_dtype):
    x = np.asarray(x)
    new_dtype = np.dtype(new_dtype)
    if new_dtype == np.bool_:
        return x != 0
    if x.dtype.kind == 'f' and new_dtype.kind in 'iu':
        with np.errstate(all='ignore'):
            return np.trunc(x).astype(np.int64).astype(new_dtype)   # float→int truncates (GLSL T(x), XLA)
    return x.astype(new_dtype)


def select(pred, a, b):
    return np.where(np.asarray(pred), a, b)      # true select; reference blends arithmetically (Q4)


# ------------------------------------------------------------------------------ structural
def broadcast_in_dim(x, shape, broadcast_dimensions):
    x = np.asarray(x)
    view = [1] * len(shape)
    for i, d in enumerate(broadcast_dimensions):
        view[d] = x.shape[i]
    return np.broadcast_to(x.reshape(view), shape).copy()


def gather(operand, indices, dimension_numbers, slice_sizes):
    """XLA gather (≙ reference gather.comp:62-103; no index clamping there -- XLA clamps starts)."""
    operand, indices = np.asarray(operand), np.asarray(indices)
    dn = dimension_numbers
    offset_dims = tuple(dn.offset_dims)
    collapsed = tuple(dn.collapsed_slice_dims)
    sim = tuple(dn.start_index_map)
    batch_shape = indices.shape[:-1]
    offset_sizes = [s for i, s in enumerate(slice_sizes) if i not in collapsed]
    rank = len(offset_dims) + len(batch_shape)
    out_shape, oi, bi = [], 0, 0
    for d in range(rank):
        if d in offset_dims:
            out_shape.append(offset_sizes[oi]); oi += 1
        else:
            out_shape.append(batch_shape[bi]); bi += 1
    out = np.empty(out_shape, operand.dtype)
    noncollapsed = [i for i in range(operand.ndim) if i not in collapsed]
    batch_out_dims = [d for d in range(rank) if d not in offset_dims]
    for out_idx in np.ndindex(*out_shape) if out_shape else [()]:
        b = tuple(out_idx[d] for d in batch_out_dims)
        start = [0] * operand.ndim
        for k, od in enumerate(sim):
            s = int(indices[b + (k,)])
            start[od] = min(max(s, 0), operand.shape[od] - slice_sizes[od])
        full = list(start)
        for k, d in enumerate(offset_dims):
            full[noncollapsed[k]] += out_idx[d]
        out[out_idx] = operand[tuple(full)]
    return out


def gather_fast(operand, indices, dimension_numbers, slice_sizes):
    """Vectorised version of `gather` (same semantics) for large shapes."""
    operand, indices = np.asarray(operand), np.asarray(indices)
    dn = dimension_numbers
    offset_dims = tuple(dn.offset_dims)
    collapsed = tuple(dn.collapsed_slice_dims)
    sim = tuple(dn.start_index_map)
    batch_shape = indices.shape[:-1]
    offset_sizes = [s for i, s in enumerate(slice_sizes) if i not in collapsed]
    rank = len(offset_dims) + len(batch_shape)
    out_shape, oi, bi = [], 0, 0
    for d in range(rank):
        if d in offset_dims:
            out_shape.append(offset_sizes[oi]); oi += 1
        else:
            out_shape.append(batch_shape[bi]); bi += 1
    grids = np.indices(out_shape) if out_shape else np.zeros((0,), np.int64)
    batch_out_dims = [d for d in range(rank) if d not in offset_dims]
    noncollapsed = [i for i in range(operand.ndim) if i not in collapsed]
    bidx = tuple(grids[d] for d in batch_out_dims)
    full = [np.zeros(out_shape, np.int64) for _ in range(operand.ndim)]
    for k, od in enumerate(sim):
        s = indices[bidx + (np.full(out_shape, k),)] if batch_out_dims or True else indices[k]
        full[od] = np.clip(s.astype(np.int64), 0, operand.shape[od] - slice_sizes[od])
    for k, d in enumerate(offset_dims):
        full[noncollapsed[k]] = full[noncollapsed[k]] + grids[d]
    return operand[tuple(full)]


def scatter_add(operand, indices, updates, dimension_numbers):
    """General XLA scatter-add; the reference only accepts two dimension-number cases
    (reference vkjax/ops.py:406-407, scatter0.comp / scatter1.comp)."""
    operand, indices, updates = np.asarray(operand), np.asarray(indices), np.asarray(updates)
    dn = dimension_numbers
    out = operand.copy()
    uwd = tuple(dn.update_window_dims)
    iwd = tuple(dn.inserted_window_dims)
    sd2od = tuple(dn.scatter_dims_to_operand_dims)
    scatter_dims = [d for d in range(updates.ndim) if d not in uwd]
    window_operand_dims = [d for d in range(operand.ndim) if d not in iwd]
    assert indices.ndim - 1 == len(scatter_dims), 'index_vector_dim must be the last dim of scatter_indices'
    for uidx in np.ndindex(*updates.shape) if updates.shape else [()]:
        sidx = tuple(uidx[d] for d in scatter_dims)
        ivec = indices[sidx]
        full = [0] * operand.ndim
        for k, od in enumerate(sd2od):
            full[od] += int(np.asarray(ivec).reshape(-1)[k])
        for k, d in enumerate(uwd):
            full[window_operand_dims[k]] += uidx[d]
        if all(0 <= f < s for f, s in zip(full, operand.shape)):
            out[tuple(full)] += updates[uidx]
    return out


def slice_(x, start_indices, limit_indices, strides):
    x = np.asarray(x)
    strides = strides or (1,) * x.ndim
    return x[tuple(slice(s, l, st) for s, l, st in zip(start_indices, limit_indices, strides))].copy()


def iota(dtype, shape, dimension):
    n = shape[dimension]
    view = [1] * len(shape); view[dimension] = n
    return np.broadcast_to(np.arange(n).astype(dtype).reshape(view), shape).copy()


# ------------------------------------------------------------------------------ reductions
def _reduce(npfn):
    def fn(x, axes):
        x = np.asarray(x)
        if tuple(axes) == ():
            return x
        if npfn in (np.sum, np.prod) and x.dtype == F32:
            return npfn(x.astype(np.float64), axis=tuple(axes)).astype(F32)
        return npfn(x, axis=tuple(axes))
    return fn


def _argreduce(npfn):
    def fn(x, axes, index_dtype=np.int32):
        x = np.asarray(x)
        assert len(axes) == 1
        return npfn(x, axis=axes[0]).astype(index_dtype)     # first occurrence wins, as argmax.comp:50-52
    return fn


REDUCE = {
    'reduce_sum': _reduce(np.sum), 'reduce_max': _reduce(np.max),
    'reduce_min': _reduce(np.min), 'reduce_prod': _reduce(np.prod),
    'argmax': _argreduce(np.argmax), 'argmin': _argreduce(np.argmin),
}


def reduce_window(x, kind, window_dimensions, window_strides, padding):
    """lax.reduce_window for max / min / sum.  lax semantics: padding is the monoid identity
    (-inf for max).  Reference shader pads with 0.0 (reduce_window_max_2d.comp:43-46, quirk Q3) --
    identical on non-negative inputs, which is all the reference tests use."""
    x = np.asarray(x)
    ident = {'max': -np.inf, 'min': np.inf, 'sum': 0.0}[kind]
    comb = {'max': np.maximum, 'min': np.minimum, 'sum': np.add}[kind]
    acc_dtype = np.float64 if (kind == 'sum' and x.dtype == F32) else x.dtype
    if x.dtype.kind in 'iu' and kind != 'sum':
        ident = np.iinfo(x.dtype).min if kind == 'max' else np.iinfo(x.dtype).max
    xp = np.pad(x.astype(acc_dtype), [(lo, hi) for lo, hi in padding], constant_values=ident)
    out_shape = tuple((xp.shape[i] - window_dimensions[i]) // window_strides[i] + 1 for i in range(x.ndim))
    out = np.full(out_shape, ident, acc_dtype)
    for w in np.ndindex(*window_dimensions):
        sl = tuple(slice(w[i], w[i] + (out_shape[i] - 1) * window_strides[i] + 1, window_strides[i]) for i in range(x.ndim))
        out = comb(out, xp[sl])
    return out.astype(x.dtype)


def select_and_scatter_add(source, operand, select, window_dimensions, window_strides, padding):
    """XLA SelectAndScatter with scatter = add (the transpose of max / min pooling; JAX's select_and_scatter_add_p).
    Plain loops over windows: within a window the selected element starts as the first in-bounds one in row-major
    order and is replaced by a candidate whenever select(selected, candidate) is false (select = ge: first maximum)."""
    source, operand = np.asarray(source), np.asarray(operand)
    keep = {'ge': lambda a, b: a >= b, 'le': lambda a, b: a <= b}[select]
    out = np.zeros(operand.shape, np.float64)
    for o in np.ndindex(*source.shape):
        best = None
        for w in np.ndindex(*window_dimensions):
            q = tuple(o[d] * window_strides[d] + w[d] - padding[d][0] for d in range(operand.ndim))
            if any(c < 0 or c >= operand.shape[d] for d, c in enumerate(q)):
                continue
            if best is None or not keep(operand[best], operand[q]):
                best = q
        if best is not None:
            out[best] += source[o]
    return out.astype(operand.dtype)


# ------------------------------------------------------------------------------ contractions
def dot_general(a, b, dimension_numbers, accum=np.float64):
    """lax.dot_general: output dims = batch dims, then lhs free dims, then rhs free dims.  The reference accepts no batch
    dims (ops.py:280); they are a SURVEY §8 f1 extension."""
    a, b = np.asarray(a), np.asarray(b)
    (lc, rc), (lb, rb) = dimension_numbers
    lc, rc, lb, rb = tuple(lc), tuple(rc), tuple(lb), tuple(rb)
    if not lb and not rb:
        out = np.tensordot(a.astype(accum), b.astype(accum), axes=(lc, rc))
        return out.astype(a.dtype)
    lfree = [d for d in range(a.ndim) if d not in lc + lb]
    rfree = [d for d in range(b.ndim) if d not in rc + rb]
    at = np.transpose(a, lb + tuple(lfree) + lc).astype(accum)
    bt = np.transpose(b, rb + tuple(rfree) + rc).astype(accum)
    nb, nc = len(lb), len(lc)
    bshape = at.shape[:nb]
    a2 = at.reshape(bshape + (int(np.prod(at.shape[nb:nb + len(lfree)], dtype=np.int64)), -1))
    b2 = bt.reshape(bshape + (int(np.prod(bt.shape[nb:nb + len(rfree)], dtype=np.int64)), -1))
    out = np.einsum('...mk,...nk->...mn', a2, b2)
    return out.reshape(bshape + tuple(a.shape[d] for d in lfree) + tuple(b.shape[d] for d in rfree)).astype(a.dtype)


def conv_general_dilated(lhs, rhs, window_strides, padding, lhs_dilation, rhs_dilation, dimension_numbers,
                         accum=np.float64, backend=None):
    """lax.conv_general_dilated, 2-D, groups == 1.  Computes in NHWC/HWIO internally, any specs in/out.
    `backend='torch'` uses torch-CPU float32 (threads) -- the CPU-baseline arm of bench.py;
    default is numpy with float64 accumulation (high-precision truth)."""
    lhs, rhs = np.asarray(lhs), np.asarray(rhs)
    dn = dimension_numbers
    x = np.transpose(lhs, (dn.lhs_spec[0], dn.lhs_spec[2], dn.lhs_spec[3], dn.lhs_spec[1]))     # NHWC
    w = np.transpose(rhs, (dn.rhs_spec[2], dn.rhs_spec[3], dn.rhs_spec[1], dn.rhs_spec[0]))     # HWIO
    N, H, W, C = x.shape
    KH, KW, _, O = w.shape
    sh, sw = window_strides
    (pt, pb), (pl, pr) = padding
    lh, lw = lhs_dilation
    rh, rw = rhs_dilation
    if backend == 'torch' and (lh, lw) == (1, 1) and min(pt, pb, pl, pr) >= 0:
        import torch
        import torch.nn.functional as Fnn
        xt = torch.from_numpy(np.ascontiguousarray(x)).permute(0, 3, 1, 2)
        wt = torch.from_numpy(np.ascontiguousarray(w)).permute(3, 2, 0, 1)
        xt = Fnn.pad(xt, (pl, pr, pt, pb))
        y = Fnn.conv2d(xt, wt, stride=(sh, sw), dilation=(rh, rw)).permute(0, 2, 3, 1).contiguous().numpy()
    else:
        Hd, Wd = (H - 1) * lh + 1, (W - 1) * lw + 1
        xd = np.zeros((N, Hd, Wd, C), accum)
        xd[:, ::lh, ::lw, :] = x
        # negative padding crops
        xd = xd[:, max(-pt, 0):Hd - max(-pb, 0), max(-pl, 0):Wd - max(-pr, 0), :]
        xd = np.pad(xd, [(0, 0), (max(pt, 0), max(pb, 0)), (max(pl, 0), max(pr, 0)), (0, 0)])
        KHd, KWd = (KH - 1) * rh + 1, (KW - 1) * rw + 1
        OH = (xd.shape[1] - KHd) // sh + 1
        OW = (xd.shape[2] - KWd) // sw + 1
        y = np.zeros((N, OH, OW, O), accum)
        wa = w.astype(accum)
        for i in range(KH):
            for j in range(KW):
                patch = xd[:, i * rh: i * rh + (OH - 1) * sh + 1: sh, j * rw: j * rw + (OW - 1) * sw + 1: sw, :]
                y += patch.reshape(-1, C) .dot(wa[i, j]).reshape(N, OH, OW, O)
        y = y.astype(lhs.dtype)
    # NHWC → out_spec
    inv = [0] * 4
    inv[dn.out_spec[0]], inv[dn.out_spec[1]], inv[dn.out_spec[2]], inv[dn.out_spec[3]] = 0, 3, 1, 2
    return np.ascontiguousarray(np.transpose(y, inv))


# ------------------------------------------------------------------------------ PRNG
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds (≙ reference threefry2x32.comp:24-90, itself following jaxlib's
    cuda_prng_kernels).  Scalar or per-element keys."""
    k0, k1 = np.asarray(k0, np.uint32), np.asarray(k1, np.uint32)
    x0, x1 = np.asarray(x0, np.uint32).copy(), np.asarray(x1, np.uint32).copy()
    with np.errstate(over='ignore'):
        ks = [k0, k1, k0 ^ k1 ^ np.uint32(0x1BD11BDA)]

        def rotl(v, d):
            return (v << np.uint32(d)) | (v >> np.uint32(32 - d))

        def rounds(x0, x1, rots):
            for r in rots:
                x0 = x0 + x1
                x1 = rotl(x1, r)
                x1 = x1 ^ x0
            return x0, x1

        x0 = x0 + ks[0]; x1 = x1 + ks[1]
        for i in range(5):
            x0, x1 = rounds(x0, x1, _ROT[i % 2])
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + np.uint32(i + 1)
    return x0.astype(np.uint32), x1.astype(np.uint32)


def threefry_hash(key, counts):
    """jax._src.prng.threefry_2x32 (JAX's original, non-partitionable layout): the flattened counters are padded to an even
    length, the first half feeds x0 and the second half x1, and the two output halves are concatenated again."""
    key = np.asarray(key, np.uint32).reshape(2)
    c = np.asarray(counts, np.uint32)
    flat = c.reshape(-1)
    odd = flat.size % 2
    if odd:
        flat = np.concatenate([flat, np.zeros(1, np.uint32)])
    h = flat.size // 2
    y0, y1 = threefry2x32(key[0], key[1], flat[:h], flat[h:])
    out = np.concatenate([y0, y1])
    return (out[:-1] if odd else out).reshape(c.shape)


def random_bits(key, shape):
    """jax._src.prng._threefry_random_bits_original for bit_width 32: hash iota(size) with the key"""
    size = int(np.prod(shape, dtype=np.int64))
    return threefry_hash(key, np.arange(size, dtype=np.uint32)).reshape(tuple(shape))


def random_seed(seed):
    """jax._src.prng.threefry_seed for a 32-bit seed: key data [0, seed mod 2**32]"""
    s = np.asarray(seed)
    lo = s.astype(np.int64).astype(np.uint64).astype(np.uint32) if s.dtype.kind == 'i' else s.astype(np.uint32)
    return np.stack([np.zeros(s.shape, np.uint32), lo], axis=-1)


# ------------------------------------------------------------------------------ evaluator
def eval_eqn(name, invals, params, eval_inner):
    name = name.replace('-', '_')
    if name in BINARY:
        return [np.asarray(BINARY[name](*invals))]
    if name in UNARY:
        return [np.asarray(UNARY[name](invals[0]))]
    if name in REDUCE:
        if name.startswith('arg'):
            return [REDUCE[name](invals[0], params['axes'], params.get('index_dtype', np.int32))]
        return [REDUCE[name](invals[0], params['axes'])]
    if name == 'integer_pow':
        return [integer_pow(invals[0], params['y'])]
    if name == 'convert_element_type':
        return [convert_element_type(invals[0], params['new_dtype'])]
    if name == 'bitcast_convert_type':
        return [np.asarray(invals[0]).view(np.dtype(params['new_dtype']))]
    if name == 'select':
        return [select(*invals)]
    if name == 'select_n':
        return [np.where(np.asarray(invals[0]).astype(bool), invals[2], invals[1])]
    if name == 'broadcast_in_dim':
        return [broadcast_in_dim(invals[0], params['shape'], params['broadcast_dimensions'])]
    if name == 'reshape':
        return [np.asarray(invals[0]).reshape(params['new_sizes'])]
    if name == 'squeeze':
        return [np.squeeze(np.asarray(invals[0]), axis=tuple(params['dimensions']))]
    if name in ('stop_gradient', 'copy', 'copy_p'):
        return [np.asarray(invals[0])]
    if name == 'transpose':
        return [np.ascontiguousarray(np.transpose(invals[0], params['permutation']))]
    if name == 'rev':
        return [np.ascontiguousarray(np.flip(invals[0], axis=tuple(params['dimensions'])))]
    if name == 'slice':
        return [slice_(invals[0], params['start_indices'], params['limit_indices'], params['strides'])]
    if name == 'concatenate':
        return [np.concatenate([np.asarray(v) for v in invals], axis=params['dimension'])]
    if name == 'iota':
        return [iota(params['dtype'], params['shape'], params['dimension'])]
    if name == 'gather':
        return [gather_fast(invals[0], invals[1], params['dimension_numbers'], params['slice_sizes'])]
    if name == 'scatter_add':
        return [scatter_add(invals[0], invals[1], invals[2], params['dimension_numbers'])]
    if name == 'dot_general':
        return [dot_general(invals[0], invals[1], params['dimension_numbers'])]
    if name == 'conv_general_dilated':
        return [conv_general_dilated(invals[0], invals[1], params['window_strides'], params['padding'],
                                     params['lhs_dilation'], params['rhs_dilation'], params['dimension_numbers'],
                                     backend=os.environ.get('ORACLE_CONV_BACKEND'))]
    if name in ('reduce_window_max', 'reduce_window_min', 'reduce_window_sum'):
        return [reduce_window(invals[0], name.rsplit('_', 1)[1], params['window_dimensions'],
                              params['window_strides'], params['padding'])]
    if name == 'select_and_scatter_add':
        sel = getattr(params['select_prim'], 'name', params['select_prim'])
        return [select_and_scatter_add(invals[0], invals[1], sel, params['window_dimensions'], params['window_strides'], params['padding'])]
    if name == 'threefry2x32':
        return list(threefry2x32(*invals))
    # typed PRNG keys (today's jax.random): a key<fry>[dims] value is its uint32[dims + (2,)] key data
    if name in ('random_wrap', 'random_unwrap'):
        return [np.asarray(invals[0], np.uint32)]
    if name == 'random_seed':
        return [random_seed(invals[0])]
    if name == 'random_bits':
        assert int(params.get('bit_width', 32)) == 32
        return [random_bits(invals[0], params['shape'])]
    if name == 'random_split':
        shape = tuple(int(d) for d in params['shape'])
        n = int(np.prod(shape, dtype=np.int64))
        return [threefry_hash(invals[0], np.arange(2 * n, dtype=np.uint32)).reshape(shape + (2,))]
    if name == 'random_fold_in':
        return [threefry_hash(invals[0], random_seed(invals[1]))]
    if name in ('xla_call', 'pjit', 'core_call', 'closed_call'):
        inner = params.get('call_jaxpr', params.get('jaxpr'))
        consts = getattr(inner, 'consts', [])
        inner = getattr(inner, 'jaxpr', inner)
        return eval_inner(inner, consts, invals)
    if name in ('custom_jvp_call_jaxpr', 'custom_jvp_call'):
        inner = params.get('fun_jaxpr', params.get('call_jaxpr'))
        return eval_inner(inner.jaxpr, inner.consts, invals)
    raise NotImplementedError(name)


def eval_jaxpr(closed_jaxpr, *args, return_env=False):
    """≙ reference tests/common.py:12-51.  `args` are the flattened (non-static) input leaves."""
    env = {}

    def run(jaxpr, consts, invals):
        def read(v):
            return np.asarray(v.val) if _is_literal(v) else env[v]
        assert len(jaxpr.invars) == len(invals), (len(jaxpr.invars), len(invals))
        for v, x in zip(jaxpr.invars, invals):
            env[v] = x
        for v, x in zip(jaxpr.constvars, consts):
            env[v] = np.asarray(x)
        for eqn in jaxpr.eqns:
            invals_ = [read(v) for v in eqn.invars]
            outs = eval_eqn(eqn.primitive.name, invals_, eqn.params, run)
            for v, o in zip(eqn.outvars, outs):
                env[v] = o
        return [read(v) for v in jaxpr.outvars]

    coerced = []
    for v, x in zip(closed_jaxpr.jaxpr.invars, args):
        coerced.append(np.asarray(x).astype(v.aval.dtype).reshape(v.aval.shape))
    outs = run(closed_jaxpr.jaxpr, closed_jaxpr.consts, coerced)
    return (outs, env) if return_env else outs
