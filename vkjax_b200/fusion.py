"""Jaxpr-level fusion over the op list produced by ops.analyze_jaxpr.

The reference executes one dispatch per equation plus one per materialised broadcast
(SURVEY.md §2.3 "Fused ops: None").  On B200 the elementwise traffic around the contractions is
what bounds ResNet-50 (SURVEY.md §7 "hard parts"), so before buffers are planned:

  1. broadcast elision   a broadcast_in_dim copy whose only readers are elementwise chains is
                         dropped; the readers index the small source directly.
  2. chain fusion        producer chain -> consumer chain (single reader, same iteration space)
                         become one kernel launch.
  3. epilogue fusion     conv/dot -> chain of {add,sub,mul,div,max,min} with immediate /
                         per-channel / same-shape operands (BatchNorm, bias, residual add, ReLU)
                         moves into the contraction kernel's epilogue.

`fuse=False` keeps one op per equation (needed by run(return_all=True), which must be able to read
every intermediate -- reference kompute_jaxpr_interpreter.py:83-89).
"""
import typing as tp

import numpy as np

from . import runtime as rt
from .ops import ChainOp, ContractionOp, KernelOp, Operand, Step, OP

_EPI_OPS = {OP[n] for n in ('ADD_F', 'SUB_F', 'MUL_F', 'DIV_F', 'MAX_F', 'MIN_F')}


def _sid(buf):
    return id(buf._ph)


def aligned_shape(in_shape, out_shape):
    return (1,) * (len(out_shape) - len(in_shape)) + tuple(in_shape)


def classify_operand(in_shape, out_shape):
    """How an elementwise operand of shape `in_shape` is indexed for an output of `out_shape`.
    Returns (kind, mod, strides) -- see B2J_OPK_* in include/b2jax.h."""
    out_shape = tuple(out_shape)
    n_in = int(np.prod(in_shape, dtype=np.int64))
    n_out = int(np.prod(out_shape, dtype=np.int64))
    if n_in == 1:
        return rt.OPK_SCALAR, 0, None
    al = aligned_shape(in_shape, out_shape)
    if al == out_shape or n_in == n_out:
        return rt.OPK_FULL, 0, None
    for k in range(len(out_shape) + 1):
        if all(d == 1 for d in al[:k]) and al[k:] == out_shape[k:]:
            return rt.OPK_MOD, int(np.prod(al[k:], dtype=np.int64)), None
    for k in range(len(out_shape), -1, -1):
        if all(d == 1 for d in al[k:]) and al[:k] == out_shape[:k]:
            return rt.OPK_DIV, int(np.prod(out_shape[k:], dtype=np.int64)), None
    strides, acc = [0] * len(al), 1
    for d in range(len(al) - 1, -1, -1):
        strides[d] = 0 if al[d] == 1 else acc
        acc *= al[d]
    return rt.OPK_STRIDED, 0, strides


def epilogue_operand_kind(op: ContractionOp, operand: Operand):
    """'imm' | 'channel' | 'full' | None (not expressible in the contraction epilogue)."""
    if operand is None:
        return None
    if operand.kind == 'imm':
        return 'imm'
    if operand.kind != 'buf':
        return None
    out_shape = tuple(op.out.shape)
    shp = tuple(operand.buf.shape)
    if shp == out_shape:
        return 'full'
    feat_dim = op.attrs['out_spec'][1] if op.what == 'conv' else len(out_shape) - 1
    n_feat = out_shape[feat_dim]
    al = aligned_shape(shp, out_shape)
    if len(al) == len(out_shape) and all((d == 1) if i != feat_dim else (d == n_feat) for i, d in enumerate(al)):
        return 'channel'
    if int(np.prod(shp, dtype=np.int64)) == 1:
        return None     # a device scalar: leave to the elementwise kernel
    return None


def _chain_distinct_inputs(init, steps):
    seen = []
    for o in [init] + [s.operand for s in steps] + [s.operand2 for s in steps]:
        if o is not None and o.kind == 'buf' and not any(o.buf.same_storage(b) and o.buf.shape == b.shape for b in seen):
            seen.append(o.buf)
    return len(seen)


def fuse(ops: tp.List, keep: tp.Set[int]) -> tp.List:
    """`keep` = storage ids (id of the tensor placeholder) that must be materialised (jaxpr outputs)."""
    # reader counts per storage
    readers: tp.Dict[int, int] = {}
    for op in ops:
        for sid in {_sid(b) for b in op.inputs()}:
            readers[sid] = readers.get(sid, 0) + 1

    def single_use(buf):
        return readers.get(_sid(buf), 0) == 1 and _sid(buf) not in keep

    out: tp.List = []
    producer: tp.Dict[int, tp.Any] = {}

    def reads_once(op, buf):
        return sum(1 for b in op.inputs() if b.same_storage(buf)) == 1

    for op in ops:
        if isinstance(op, ChainOp):
            # ---- 1. broadcast elision --------------------------------------------------------
            def elide(o: tp.Optional[Operand]):
                if o is None or o.kind != 'buf':
                    return o
                prod = producer.get(_sid(o.buf))
                src = getattr(prod, 'bcast_src', None)
                if src is None or not single_use(o.buf) or tuple(o.buf.shape) != tuple(op.out.shape):
                    return o
                inbuf, view_shape = src
                out.remove(prod)
                del producer[_sid(o.buf)]
                return Operand('buf', inbuf.view(inbuf.dtype, view_shape))
            op.init = elide(op.init)
            op.steps = [s._replace(operand=elide(s.operand), operand2=elide(s.operand2)) for s in op.steps]

            # ---- 2./3. absorb the producer of the accumulator --------------------------------
            merged = True
            while merged:
                merged = False
                cands = []
                if op.init.kind == 'buf':
                    cands.append(('init', op.init.buf))
                if op.steps and op.steps[0].operand is not None and op.steps[0].operand.kind == 'buf' \
                        and op.steps[0].op != OP['SELECT']:
                    cands.append(('step0', op.steps[0].operand.buf))
                # prefer fusing into a contraction
                cands.sort(key=lambda c: 0 if isinstance(producer.get(_sid(c[1])), ContractionOp) else 1)
                for where, buf in cands:
                    prod = producer.get(_sid(buf))
                    if prod is None or not isinstance(prod, (ChainOp, ContractionOp)):
                        continue
                    if not single_use(buf) or not reads_once(op, buf):
                        continue
                    if tuple(prod.out.shape) != tuple(op.out.shape) or tuple(buf.shape) != tuple(op.out.shape):
                        continue
                    steps = list(op.steps)
                    if where == 'step0':
                        # acc and operand trade places on the first step
                        s0 = steps[0]
                        steps[0] = s0._replace(operand=op.init, swap=not s0.swap)
                    if isinstance(prod, ChainOp):
                        if len(prod.steps) + len(steps) > rt.ELT_MAX_STEPS:
                            continue
                        if _chain_distinct_inputs(prod.init, prod.steps + steps) > rt.ELT_MAX_IN:
                            continue
                        out.remove(prod)
                        op.init, op.steps = prod.init, prod.steps + steps
                        op.equations = prod.equations + op.equations
                        merged = True
                        break
                    else:
                        if len(prod.epilogue) + len(steps) > rt.EPI_MAX_STEPS:
                            continue
                        if not all(s.op in _EPI_OPS and epilogue_operand_kind(prod, s.operand) is not None for s in steps):
                            continue
                        out.remove(prod)
                        prod.epilogue = prod.epilogue + steps
                        prod.out = op.out
                        prod.equations = prod.equations + op.equations
                        op = prod
                        break
                if not isinstance(op, ChainOp):
                    break
        out.append(op)
        for b in op.outputs():
            producer[_sid(b)] = op
    return out
